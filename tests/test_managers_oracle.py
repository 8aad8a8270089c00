"""oracle/managers_oracle.py against the golden vectors made by the reference's own manager classes
(tests/golden/make_golden_managers.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import managers_oracle as MO

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'managers.npz'))
N = G['ramp_w0'].shape[0]
TOL = 1e-11


def close(a, b, tol=TOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max()), (np.abs(a - b).max(), a, b)


def quat_close(a, b, tol=TOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    close(a, b, tol)  # same sign convention too: the managers hand these to update_desired as they are


def test_ramps():
    for i in range(N):
        a = (G['ramp_t0'][i], G['ramp_dur'][i], G['ramp_t'][i])
        close(MO.ramp(G['ramp_w0'][i], float(G['ramp_wmin']), *a), G['ramp_w_to_min'][i])
        close(MO.ramp(G['ramp_w0'][i], float(G['ramp_wmax']), *a), G['ramp_w_to_max'][i])
        close(MO.ramp(G['ramp_rfz0'][i], 0.001, *a), G['ramp_rf_to_min'][i])
        close(MO.ramp(G['ramp_rfz0'][i], float(G['ramp_rfzmax']), *a), G['ramp_rf_to_max'][i])


def test_foot_use_current():
    for i in range(N):
        out = MO.foot_use_current(G['foot_iso'][i], G['foot_vel'][i])
        for k, v in zip(('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc'), out):
            close(v, G['uc_' + k][i])


def test_swing_foot():
    for i in range(N):
        t = G['sw_t0'][i] + G['sw_frac'][i] * G['sw_dur'][i]
        out = MO.swing_foot_desired(G['foot_iso'][i], G['land_pos'][i], G['land_quat'][i], G['sw_t0'][i], G['sw_dur'][i],
                                    float(G['sw_height']), t)
        for k, v in zip(('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc'), out):
            close(v, G['sw_' + k][i], 1e-10)


def test_floating_base():
    for i in range(N):
        out = MO.floating_base_desired(G['com'][i], G['base_iso'][i][:3, :3], G['tgt_com'][i], G['tgt_quat'][i],
                                       G['fb_t0'][i], G['fb_dur'][i], G['fb_t'][i])
        for k, v in zip(('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc'), out):
            close(v, G['fb_' + k][i], 1e-10)
        p, v, a = MO.sinusoid(G['fb_t0'][i], G['com'][i], G['sway_amp'], G['sway_freq'], G['fb_t'][i])
        close(p, G['fb_sway_pos'][i]); close(v, G['fb_sway_vel'][i]); close(a, G['fb_sway_acc'][i])


def test_dcm():
    for i in range(N):
        d, p, v = MO.dcm_update(G['com'][i], G['dcm_vcom'][i], G['dcm_state'][i], float(G['dcm_dt']))
        close(d, G['dcm_out'][i]); close(p, G['dcm_prev_out'][i]); close(v, G['dcm_vel_out'][i])


def test_hand_trajectory():
    """HandTrajectoryManager (hand_trajectory_manager.py): use_current, the plain and the keypoint trajectory, start
    angular velocities on both sides of HermiteCurveQuat's 1e-6 switch, times before the start and after the end."""
    for i in range(N):
        t = G['hand_t0'][i] + G['hand_frac'][i] * G['hand_dur'][i]
        uc = MO.foot_use_current(G['hand_iso'][i], G['hand_vel6'][i])
        close(uc[0], G['hand_uc_pos'][i]); close(uc[1], G['hand_uc_vel'][i]); close(uc[3], G['hand_uc_quat'][i])
        close(uc[4], G['hand_uc_angvel'][i])
        for key, pre in ((None, ''), (G['hand_key'][i], 'kp_')):
            out = MO.hand_desired(G['hand_iso'][i], G['hand_vel6'][i], G['hand_tgt_pos'][i], G['hand_tgt_quat'][i],
                                  G['hand_t0'][i], G['hand_dur'][i], t, key)
            for k, v in zip(('pos', 'vel', 'acc', 'quat', 'angvel', 'angacc'), out):
                close(v, G['hand_' + pre + k][i], 1e-10)
