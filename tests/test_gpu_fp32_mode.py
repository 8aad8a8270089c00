"""The optional fp32 mode (BASELINE.json north_star: "1e-4 in an optional fp32 mode"): pnc_set_precision(
PNC_PREC_FP32_DYNAMICS) runs k_update_system in fp32 arithmetic; the factorisations and the active-set iterations stay
fp64 (hard parts: cond(H) ~ 1e9..1e11 with lambda_q = 1e-8, config/atlas_config.py:65-66).

Tolerances, written out:
* rigid-body quantities (M, g, c, CoM, Jacobians, frames): 2e-5 of the state's largest entry against the fp64 kernel
  (measured 1e-7 .. 4e-6; Jdot_com qdot 1.4e-5: a difference of terms a hundred times its size); q exact.
* trq / acc / rf against the fp64 ORACLE: 1e-4 relative (the north star's figure) on every state but the
  ill-conditioned tail, whose share is bounded per robot by what merely ROUNDING the fp64 rigid-body quantities to fp32
  does (scripts/fp32_feasibility.py: the floor of any fp32 dynamics).  With lambda_rf = 1e-7 the
  split of the reaction forces over Laikago's four point contacts is fixed by the regulariser alone, and a 1e-7
  perturbation of the data moves it by more than 1e-4 on ~5 % of the states.
"""
import numpy as np
import pytest

from pypnc_b200 import robots

pytestmark = pytest.mark.gpu

KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
FIELDS = ('M', 'grav', 'cori', 'com', 'vcom', 'jcom', 'jcomdot', 'jcom_jdqd', 'frame_iso', 'frame_vel', 'frame_jac',
          'frame_jdqd', 'qdot')
# (robot, states, min share of states within 1e-4 of the fp64 oracle, max median error)
CASES = [('atlas', 4096, 0.998, 1e-5), ('laikago', 4096, 0.92, 1e-5), ('valkyrie', 2048, 0.97, 2e-5),
         ('draco3', 2048, 0.99, 2e-5), ('draco3_lb', 2048, 0.985, 2e-5)]


@pytest.mark.parametrize("name,B,min_share,max_median", CASES)
def test_fp32_mode_against_fp64_path_and_oracle(name, B, min_share, max_median, oracle_mod):
    import torch
    from pypnc_b200 import engine
    spec = robots.PROBLEMS[name]()
    op = oracle_mod.OracleProblem(spec)
    st = robots.synth_states(name, spec, B, seed=33)
    des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=33)
    ref = op.tick_batch(st, des, w, rfz, nthreads=8)
    dm = engine.Model(spec.model, 0)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)
    d_st = [torch.from_numpy(np.ascontiguousarray(st[k])).cuda() for k in KEYS]
    d_in = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (des, w, rfz)]
    assert engine.get_precision() == engine.PREC_FP64
    ws.update_system(*d_st)
    f64 = {f: ws.view(f).clone() for f in FIELDS + ('q',)}
    out64 = {k: v.cpu().numpy() for k, v in engine.ihwbc_solve(pb, ws, *d_in).items()}
    engine.set_precision(engine.PREC_FP32_DYNAMICS)
    try:
        assert engine.get_precision() == engine.PREC_FP32_DYNAMICS
        ws.update_system(*d_st)
        f32 = {f: ws.view(f).clone() for f in FIELDS + ('q',)}
        out32 = {k: v.cpu().numpy() for k, v in engine.ihwbc_solve(pb, ws, *d_in).items()}
        torch.cuda.synchronize()
    finally:
        engine.set_precision(engine.PREC_FP64)
    # the mode really ran in fp32: the results differ from the fp64 kernel's, and q (a copy of the inputs) does not
    assert torch.equal(f32['q'], f64['q'])
    assert not torch.equal(f32['M'], f64['M'])
    for f in FIELDS:
        a, b = f64[f].reshape(B, -1), f32[f].reshape(B, -1)
        e = ((a - b).abs().amax(1) / a.abs().amax(1).clamp_min(1e-30)).max().item()
        assert e < 2e-5, (f, e)
    M32 = f32['M']
    assert torch.equal(M32, M32.transpose(1, 2))  # still exactly symmetric
    # outputs: same status everywhere, 1e-4 of the fp64 oracle on all but the ill-conditioned tail
    assert (out32['status'] == ref['status']).all() and (out64['status'] == ref['status']).all()
    ok = ref['status'] == 0
    err = np.zeros(B)
    for k in ('trq', 'acc', 'rf'):
        den = np.maximum(np.abs(ref[k]).max(1), 1e-3 * np.abs(ref[k]).max())
        err = np.maximum(err, np.abs(out32[k] - ref[k]).max(1) / den)
    e = err[ok]
    share = float((e <= 1e-4).mean())
    print('fp32 mode %s: %d states, median %.2e p99 %.2e max %.2e, within 1e-4: %.4f' %
          (name, e.size, np.median(e), np.quantile(e, 0.99), e.max(), share))
    assert np.median(e) < max_median
    assert share >= min_share, share
    # back in fp64: bit for bit the first answer
    ws.update_system(*d_st)
    again = engine.ihwbc_solve(pb, ws, *d_in)
    assert np.array_equal(again['trq'].cpu().numpy(), out64['trq'], equal_nan=True)


def test_fp32_mode_through_tick_host(oracle_mod):
    """pnc_wbc_tick_host honours the mode (host buffers in and out) and equals the device path of the same mode."""
    import ctypes as C
    import torch
    from pypnc_b200 import engine
    name, B = 'atlas', 512
    spec = robots.PROBLEMS[name]()
    op = oracle_mod.OracleProblem(spec)
    st = robots.synth_states(name, spec, B, seed=5)
    des, w, rfz = robots.synth_desireds(name, spec, st, op.task_actuals(st), seed=5)
    dm = engine.Model(spec.model, 0)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)
    lib = dm.lib
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    h_in = [np.ascontiguousarray(st[k]) for k in KEYS] + [np.ascontiguousarray(a) for a in (des, w, rfz)]
    res = {}
    for mode in (engine.PREC_FP64, engine.PREC_FP32_DYNAMICS):
        engine.set_precision(mode)
        try:
            trq, acc = np.empty((B, spec.nact)), np.empty((B, spec.nact))
            rf = np.empty((B, pb.nc))
            status, iters = np.empty(B, np.int32), np.empty(B, np.int32)
            active = np.empty((B, pb.mw), np.int64)
            rc = lib.pnc_wbc_tick_host(pb.h, ws.h, B, *[hp(a) for a in h_in], hp(trq), hp(acc), hp(rf), hp(status),
                                       hp(iters), hp(active))
            assert rc == 0
            ws.update_system(*[torch.from_numpy(a).cuda() for a in h_in[:6]])
            dev = engine.ihwbc_solve(pb, ws, *[torch.from_numpy(a).cuda() for a in h_in[6:]])
            assert np.array_equal(dev['trq'].cpu().numpy(), trq, equal_nan=True)
            res[mode] = trq
        finally:
            engine.set_precision(engine.PREC_FP64)
    d = np.abs(res[0] - res[1]).max() / np.abs(res[0]).max()
    assert 0 < d < 1e-3, d
