#!/usr/bin/env python
"""bench.py -- WBC solves/s (robot_system update + IHWBC QP), Atlas DCM-walking task set.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--robot atlas] [--batch 65536] [--scaling weak|strong] [--no-configs]

One "step" = one pass of the hot path over one batch of synthetic states: pnc_update_system
(FK, Jacobians, CoM, CRBA, RNEA terms) followed by pnc_ihwbc_solve (task/contact evaluation, QP
assembly, dual active-set solve, torque recovery) for the batch of independent robot states.

Headline (the JSON line's `value`): BASELINE.json's metric, Atlas, 64 K states per GPU ("weak":
per-GPU work fixed; `--scaling strong` shards one 64 K batch over the N GPUs instead).
Multi-GPU: one process per GPU (torchrun), the batch dimension is sharded (pypnc_b200/sharding.py),
no data-path collective; only the final stats gather uses NCCL.

`configs` in the same line carries the other BASELINE.json configurations, each timed the same way
(CUDA events, barrier on both sides, max over ranks, clocks sampled during its own timed region):
  config 1  three-link manipulator OSC step, 1 K states (rank 0)
  config 2  Laikago stance, 16 K states global
  config 4  Draco3 with the rolling-joint constraint, 64 K states GLOBAL, sharded over the N GPUs
  config 5  Valkyrie, 256 K states GLOBAL, sharded over the N GPUs
(global batches: these are strong-scaling points by the configs' own definition).

--impl reference times the CPU restatement of the reference loop (oracle/, -O3, all host threads)
on a bounded sample of the same workload; see DESIGN.md section 6.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "wbc_solves_per_sec"
UNIT = "solves/s"
KEYS = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--robot', default='atlas')
    ap.add_argument('--batch', type=int, default=65536, help='states per GPU (weak) / in total (strong)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    ap.add_argument('--cpu-sample', type=int, default=0, help='states in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='headline only: skip BASELINE configs 1, 2, 4, 5')
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'fp32'],
                    help="fp32 = the optional fp32 mode for the whole run (pnc_set_precision: update_system in fp32 "
                         "arithmetic, factorisations and active-set iterations fp64); the default fp64 run reports the "
                         "mode's speed and accuracy on the headline batch under `fp32_mode`")
    ap.add_argument('--same-state', action='store_true', help='diagnostic: every state of the batch is state 0')
    return ap.parse_args()


WORKLOADS = {'atlas': 'Atlas DCM-walking IHWBC (atlas_tci_container task set, 2 surface contacts, torque limits)',
             'laikago': 'Laikago stance IHWBC, 4 point contacts',
             'draco3': 'Draco3 IHWBC with rolling-joint internal constraint',
             'draco3_lb': 'Draco3 lower body IHWBC, internal constraint + torque limits',
             'valkyrie': 'Valkyrie full-body IHWBC (synthesised task set)'}


def make_config(robot, batch_per_gpu, world, scaling, qp_dims):
    """The `config` object -- one function for both arms, so the two lines carry the same keys."""
    return {"workload": "%s, batch %d per GPU" % (WORKLOADS.get(robot, robot), batch_per_gpu), "robot": robot,
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * world if scaling == 'weak' else None,
            "qp_dims_n_p_m": list(qp_dims),
            "l2": "inputs + per-state workspace of one step exceed the 126 MB L2 (no flush needed)",
            "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world}


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
                 nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake_slowdown'}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_rank):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_reference_run(robot, sample, threads, repeats=1, seed=1):
    """Times the CPU oracle loop (update + assemble + QuadProg++ port) on `sample` states and counts
    the algorithmic flops per solve on the same states."""
    key = (robot, sample, threads, seed)
    if key in _CPU_CACHE:
        return _CPU_CACHE[key]
    from pypnc_b200 import robots
    from oracle import oracle
    spec = robots.PROBLEMS[robot]()
    st = robots.synth_states(robot, spec, sample, seed=seed)
    op = oracle.OracleProblem(spec, fast=True)
    act = op.task_actuals(st)
    des, w, rfz = robots.synth_desireds(robot, spec, st, act, seed=seed)
    best, out = None, None
    for _ in range(max(1, repeats)):
        out = op.tick_batch(st, des, w, rfz, nthreads=threads)
        best = out['seconds'] if best is None else min(best, out['seconds'])
    r = dict(solves_per_s=sample / best, seconds=best, mean_flops=float(out['flops'].mean()),
             ok=int((out['status'] == 0).sum()), mean_iters=float(out['iters'].mean()))
    _CPU_CACHE[key] = r
    return r


def cpu_baseline_for(robot, sample_multi):
    """cpu_baseline object (+ oracle-counted flops per solve) of one robot: one thread on <= 1024 states
    (task_actuals is a Python loop, kept short), all host threads on `sample_multi` states."""
    threads = os.cpu_count() or 1
    n1 = min(sample_multi, 1024)
    c1 = cpu_reference_run(robot, n1, 1)
    cN = cpu_reference_run(robot, sample_multi, threads) if threads > 1 else c1
    return ({"value": cN['solves_per_s'], "unit": UNIT, "cores": threads, "kind": "port",
             "sample": "%d states of the same synthetic generator (seed 1) on %d host threads; single thread: "
                       "%.0f solves/s on %d states; oracle/ C restatement, gcc -O3 -march=x86-64-v3"
                       % (sample_multi, threads, c1['solves_per_s'], n1),
             "single_thread_value": c1['solves_per_s']}, cN['mean_flops'])


def run_reference(args):
    """--impl reference: the reference path's CPU restatement on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', '1'))
    threads = os.cpu_count() or 1
    sample = args.cpu_sample or 4096 * max(1, min(threads, 16))
    # inputs are generated once, outside the timed steps
    from pypnc_b200 import robots, sharding
    from oracle import oracle
    spec = robots.PROBLEMS[args.robot]()
    st = robots.synth_states(args.robot, spec, sample, seed=1)
    op = oracle.OracleProblem(spec, fast=True)
    des, w, rfz = robots.synth_desireds(args.robot, spec, st, op.task_actuals(st), seed=1)
    for _ in range(args.warmup):
        op.tick_batch(st, des, w, rfz, nthreads=threads)
    t0 = time.perf_counter()
    secs = 0.0
    for _ in range(args.steps):
        secs += op.tick_batch(st, des, w, rfz, nthreads=threads)['seconds']
    wall = time.perf_counter() - t0
    value = sample * args.steps / secs
    B = args.batch if args.scaling == 'weak' else sharding.shard_range(args.batch, 0, world)[1]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args.robot, B, world, args.scaling, spec.qp_dims),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d states per step of the same synthetic generator (seed 1), oracle -O3 "
                                       "build, %d host threads over contiguous shards" % (sample, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
class Env:
    """torch / distributed / library handles of this rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from pypnc_b200 import _cabi
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
        self.dev = torch.device('cuda', self.local_rank)
        self.affinity = self._bind_to_gpu_numa_node()
        self.lib = _cabi.load_library()
        self.stream = torch.cuda.current_stream()

    def _bind_to_gpu_numa_node(self):
        """Run this rank (and first-touch its pinned staging buffers) on the CPUs next to its GPU: with N ranks
        on one node the host side of the end-to-end path otherwise shares one NUMA node's DRAM and PCIe root."""
        if self.world <= 1 or os.environ.get('PNC_NO_AFFINITY'):
            return None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(self.local_rank))
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = [64 * wi + b for wi, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1 and 64 * wi + b < ncpu]
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
            if allowed:
                os.sched_setaffinity(0, allowed)
                return len(allowed)
        except Exception:
            pass
        return None

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


def run_wbc_config(env, robot, B, K, warmup, seed, same_state=False, want_e2e=True, min_ms=0.0, want_fp32=False):
    """Times K steps of update_system + IHWBC solve on this rank's B states.  Returns a dict with the
    max-over-ranks time, per-stage times, solver stats, clocks, launches and (optionally) the end-to-end
    numbers through pnc_wbc_tick_host."""
    import ctypes as C
    torch, lib = env.torch, env.lib
    from pypnc_b200 import robots, engine, sharding
    dev = env.dev
    spec = robots.PROBLEMS[robot]()
    dm = engine.Model(spec.model, env.local_rank)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)
    st = robots.synth_states(robot, spec, B, seed=seed)
    if same_state:
        st = {k: np.ascontiguousarray(np.repeat(v[:1], B, axis=0)) for k, v in st.items()}
    d_st = [torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in KEYS]
    ws.update_system(*d_st)
    torch.cuda.synchronize()
    des, w, rfz = robots.synth_desireds(robot, spec, st, engine.task_actuals_from_workspace(spec, ws), seed=seed)
    if same_state:
        des, w, rfz = [np.ascontiguousarray(np.repeat(a[:1], B, axis=0)) for a in (des, w, rfz)]
    d_des, d_w, d_rfz = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (des, w, rfz)]
    nact, nc, mw = spec.nact, pb.nc, pb.mw
    f64 = dict(dtype=torch.float64, device=dev)
    o_trq, o_acc = torch.empty((B, nact), **f64), torch.empty((B, nact), **f64)
    o_rf = torch.empty((B, nc), **f64)
    o_status = torch.empty(B, dtype=torch.int32, device=dev)
    o_iters = torch.empty(B, dtype=torch.int32, device=dev)
    o_active = torch.empty((B, mw), dtype=torch.int64, device=dev)
    P = engine._ptr
    stream = env.stream
    sp = engine.vp(stream.cuda_stream)

    def step():
        rc = lib.pnc_update_system(ws.h, B, *[P(t) for t in d_st], sp)
        rc |= lib.pnc_ihwbc_solve(pb.h, ws.h, B, P(d_des), P(d_w), P(d_rfz), P(o_trq), P(o_acc), P(o_rf), None,
                                  P(o_status), P(o_iters), P(o_active), sp)
        assert rc == 0, rc

    for _ in range(max(3, warmup)):
        step()
    env.barrier()
    if min_ms > 0:  # short configurations: time at least `min_ms` of work so the clock sampler sees the region
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        step()
        t1.record(stream)
        torch.cuda.synchronize()
        est = sharding.max_over_ranks(t0.elapsed_time(t1), device=dev)
        K = int(max(K, min(400, np.ceil(min_ms / max(est, 1e-3)))))
    sampler = ClockSampler(physical_gpu_index(env.local_rank))
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = lib.pnc_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for k in range(K):
        ev[k][0].record(stream)
        rc = lib.pnc_update_system(ws.h, B, *[P(t) for t in d_st], sp)
        ev[k][1].record(stream)
        rc |= lib.pnc_ihwbc_solve(pb.h, ws.h, B, P(d_des), P(d_w), P(d_rfz), P(o_trq), P(o_acc), P(o_rf), None,
                                  P(o_status), P(o_iters), P(o_active), sp)
        ev[k][2].record(stream)
        assert rc == 0, rc
    t_end.record(stream)
    env.barrier()
    launches = lib.pnc_launch_count() - launches0
    clocks = sampler.finish()
    ms_total = sharding.max_over_ranks(t_start.elapsed_time(t_end), device=dev)
    ms_dyn = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    ms_qp = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    status, iters = o_status.cpu().numpy(), o_iters.cpu().numpy()
    stats = sharding.gather_stats(sharding.solver_stats(status, iters), device=dev)  # the only collective: a stats gather
    res = dict(spec=spec, B=B, K=K, ms_total=ms_total, ms_per_step=ms_total / K, ms_dyn=ms_dyn, ms_qp=ms_qp, stats=stats,
               clocks=clocks, launches=int(launches), ws_gb=ws.nbytes / 1e9, mw=mw, nact=nact, nc=nc, e2e=None)

    # ---- the optional fp32 mode on the same batch: speed, and distance of its outputs from the fp64 path's ----
    if want_fp32 and lib.pnc_get_precision() == 0:
        ref = {k: t.clone() for k, t in (('trq', o_trq), ('acc', o_acc), ('rf', o_rf))}
        ref_active = o_active.clone()
        assert lib.pnc_set_precision(1) == 0
        try:
            for _ in range(3):
                step()
            env.barrier()
            f0, f1, f2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            f0.record(stream)
            for k in range(K):
                rc = lib.pnc_update_system(ws.h, B, *[P(t) for t in d_st], sp)
                assert rc == 0, rc
            f1.record(stream)
            for k in range(K):
                step()
            f2.record(stream)
            env.barrier()
        finally:
            assert lib.pnc_set_precision(0) == 0
        ms32 = sharding.max_over_ranks(f1.elapsed_time(f2), device=dev) / K
        st32 = o_status.cpu().numpy()
        ok = torch.from_numpy((status == 0) & (st32 == 0)).to(dev)
        err = torch.zeros(B, **f64)
        for k, t in (('trq', o_trq), ('acc', o_acc), ('rf', o_rf)):
            den = torch.maximum(ref[k].abs().amax(1), 1e-3 * ref[k].abs().max())
            err = torch.maximum(err, (t - ref[k]).abs().amax(1) / den)
        e = err[ok].cpu().numpy()
        same = ((o_active == ref_active).all(1) & ok).sum().item()
        res['fp32'] = dict(ms_per_step=ms32, ms_dyn=f0.elapsed_time(f1) / K, n=int(e.size),
                           status_mismatch=int((st32 != status).sum()), same_final_active_set=int(same),
                           median=float(np.median(e)), p99=float(np.quantile(e, 0.99)), max=float(e.max()),
                           within_1e4=int((e <= 1e-4).sum()))
        step()  # leave the fp64 outputs in place for the end-to-end cross-check below
        torch.cuda.synchronize()

    # ---- end to end through the host-buffer C ABI call (pinned host memory), every timed step ----
    if want_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = [pin(st[k]) for k in KEYS] + [pin(des), pin(w), pin(rfz)]
        h_trq = torch.empty((B, nact), dtype=torch.float64).pin_memory()
        h_acc = torch.empty((B, nact), dtype=torch.float64).pin_memory()
        h_rf = torch.empty((B, nc), dtype=torch.float64).pin_memory()
        h_status = torch.empty(B, dtype=torch.int32).pin_memory()
        h_iters = torch.empty(B, dtype=torch.int32).pin_memory()
        h_active = torch.empty((B, mw), dtype=torch.int64).pin_memory()
        HP = lambda t: engine.vp(t.data_ptr())

        def tick_host():
            rc = lib.pnc_wbc_tick_host(pb.h, ws.h, B, *[HP(t) for t in h_in], HP(h_trq), HP(h_acc), HP(h_rf),
                                       HP(h_status), HP(h_iters), HP(h_active))
            assert rc == 0, rc

        for _ in range(3):
            tick_host()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            tick_host()  # synchronous: returns when the outputs are in host memory
        torch.cuda.synchronize()
        dt = sharding.max_over_ranks(time.perf_counter() - t0, device=dev)
        h2d = sum(t.numel() * t.element_size() for t in h_in)
        d2h = sum(t.numel() * t.element_size() for t in (h_trq, h_acc, h_rf, h_status, h_iters, h_active))
        res['e2e'] = dict(dt=dt, h2d=int(h2d), d2h=int(d2h), steps=K)
        # the host path must agree with the device-resident path
        assert (h_status.numpy() == status).all()
        assert np.array_equal(h_trq.numpy(), o_trq.cpu().numpy(), equal_nan=True)
    del ws, pb, dm
    return res


def alg_bytes_per_solve(spec, nact, nc, mw):
    """inputs + outputs of one tick."""
    na = spec.model.na
    bytes_in = (3 + 4 + 3 + 3 + 2 * na + spec.ndes + len(spec.tasks) + len(spec.contacts)) * 8
    bytes_out = (2 * nact + nc) * 8 + 8 * mw + 8
    return bytes_in + bytes_out


def roofline_of(res, mean_flops, flops_src, fp64_peak, hbm_peak, hbm_src, traffic):
    B = res['B']
    alg_bytes = alg_bytes_per_solve(res['spec'], res['nact'], res['nc'], res['mw'])
    qp_tflops = B * mean_flops / (res['ms_qp'] * 1e-3) * 1e-12
    return {"bound": "fp64",
            "kernel": "pnc_ihwbc_solve = k_task_prepare (+ k_ic_prepare) + k_ihwbc_setup + k_ihwbc_reg + k_ihwbc_finish",
            "achieved": qp_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": qp_tflops / fp64_peak if fp64_peak else None,
            "traffic": traffic,
            "peak_source": "measured in this run: DFMA-chain microbenchmark, all SMs (pnc_measure_fp64_peak)",
            "flops_per_solve": mean_flops, "flops_source": flops_src, "kernel_ms": res['ms_qp'],
            "hbm": {"achieved": B * alg_bytes / (res['ms_per_step'] * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": B * alg_bytes / (res['ms_per_step'] * 1e-3) * 1e-9 / hbm_peak,
                    "bytes_per_solve": alg_bytes, "peak_source": hbm_src}}


def e2e_of(res, world_states):
    e = res['e2e']
    if not e:
        return None
    return {"value": world_states * e['steps'] / e['dt'], "unit": UNIT, "h2d_bytes_per_step": e['h2d'],
            "d2h_bytes_per_step": e['d2h'], "steps": e['steps'], "ms_per_step": 1e3 * e['dt'] / e['steps'],
            "api": "pnc_wbc_tick_host (pinned host buffers in, host buffers out, synchronous)"}


def run_osc_config(env, K, warmup):
    """BASELINE config 1: three-link manipulator OSC control step (manipulator_interface.py:77-114), 1 K joint
    states -- the reference's CPU-runnable case; rank 0 only."""
    import ctypes as C
    torch, lib = env.torch, env.lib
    from pypnc_b200 import robots, engine
    m = robots.load_model('three_link_manipulator')
    B = 1024
    rng = np.random.default_rng(11)
    jp, jv = rng.uniform(-1.5, 1.5, (B, 3)), rng.normal(0, 1.0, (B, 3))
    pos_des, vel_des = rng.uniform(-2, 2, (B, 3)), rng.normal(0, 0.3, (B, 3))
    pos_des[:, 2] = 0.0
    kp, kd = np.array([10., 10., 10.]), np.array([1., 1., 1.])
    ee = m.frame_id('ee')
    dm = engine.Model(m, env.local_rank)
    ws = engine.Workspace(dm, B, [ee])
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(env.dev)
    d_jp, d_jv, d_pd, d_vd, d_kp, d_kd = d(jp), d(jv), d(pos_des), d(vel_des), d(kp), d(kd)
    trq = torch.zeros((B, 3), dtype=torch.float64, device=env.dev)
    P = engine._ptr
    sp = engine.vp(env.stream.cuda_stream)

    def step():
        rc = lib.pnc_update_system(ws.h, B, None, None, None, None, P(d_jp), P(d_jv), sp)
        rc |= lib.pnc_osc_step(ws.h, B, ee, P(d_pd), P(d_vd), P(d_kp), P(d_kd), 1.0, P(trq), None, sp)
        assert rc == 0, rc

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(env.local_rank))
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.pnc_launch_count()
    Kk = max(K, 50)  # a 1 K-state step is two ~10 us launches: time enough of them
    t0.record(env.stream)
    for _ in range(Kk):
        step()
    t1.record(env.stream)
    torch.cuda.synchronize()
    clocks = sampler.finish()
    ms = t0.elapsed_time(t1) / Kk
    # end to end: pinned host inputs -> device -> kernels -> torques back on the host, every step
    h_jp, h_jv = [torch.from_numpy(a).pin_memory() for a in (jp, jv)]
    h_trq = torch.empty((B, 3), dtype=torch.float64).pin_memory()

    def step_host():
        d_jp.copy_(h_jp, non_blocking=True)
        d_jv.copy_(h_jv, non_blocking=True)
        step()
        h_trq.copy_(trq, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        step_host()
    w0 = time.perf_counter()
    for _ in range(Kk):
        step_host()
    dt = (time.perf_counter() - w0) / Kk
    from oracle import oracle
    om = oracle.OracleModel(m, fast=True)
    q = jp.copy()
    reps = 20
    sec = min(om.osc_step_batch(ee, q, jv, pos_des, vel_des, kp, kd)[2] for _ in range(reps))
    t_ref = om.osc_step_batch(ee, q, jv, pos_des, vel_des, kp, kd)[0]
    err = float(np.abs(trq.cpu().numpy() - t_ref).max() / np.abs(t_ref).max())
    return {"workload": "three-link manipulator OSC control step (manipulator_interface.py:77-114), 1024 joint states, rank 0",
            "metric": "osc_steps_per_sec", "unit": "control steps/s", "value": B / (ms * 1e-3), "ms_per_step": ms, "steps": Kk,
            "batch": B, "n_gpus": 1, "gpu_launches": int(lib.pnc_launch_count() - n0), "clocks": clocks,
            "e2e": {"value": B / dt, "unit": "control steps/s", "h2d_bytes_per_step": 2 * B * 3 * 8, "d2h_bytes_per_step": B * 3 * 8,
                    "ms_per_step": 1e3 * dt},
            "roofline": {"bound": "launch latency", "note": "two launches of a few microseconds of work each; 1 K states "
                         "occupy 8 of 148 SMs -- the configuration is the reference's CPU-runnable case, not a GPU load"},
            "cpu_baseline": {"value": B / sec, "unit": "control steps/s", "cores": 1, "kind": "port",
                             "sample": "the same 1024 states, oracle C loop, one thread, best of %d" % reps},
            "max_rel_err_vs_oracle": err}


# ---------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    env = Env()
    from pypnc_b200 import sharding
    world, rank = env.world, env.rank
    if args.precision == 'fp32':
        assert env.lib.pnc_set_precision(1) == 0
    fp32_run = env.lib.pnc_get_precision() == 1
    K = args.steps
    robot = args.robot
    if args.scaling == 'weak':
        B = args.batch
        world_states = B * world
    else:
        lo, hi = sharding.shard_range(args.batch, rank, world)
        B = hi - lo
        world_states = args.batch
    res = run_wbc_config(env, robot, B, K, args.warmup, seed=1 + rank, same_state=args.same_state, want_e2e=not args.no_e2e,
                         want_fp32=not fp32_run)
    value = world_states * K / (res['ms_total'] * 1e-3)

    # ---- the other BASELINE configs (every rank takes part in the sharded ones) ----
    extra = {}
    if not args.no_configs and robot == 'atlas':
        Kc = max(3, min(K, 10))
        for key, rb, gbatch in (('config2_laikago_16k', 'laikago', 16384), ('config4_draco3_64k', 'draco3', 65536),
                                ('config5_valkyrie_256k', 'valkyrie', 262144)):
            lo, hi = sharding.shard_range(gbatch, rank, world)
            r = run_wbc_config(env, rb, hi - lo, Kc, args.warmup, seed=1 + rank, want_e2e=not args.no_e2e, min_ms=150.0)
            extra[key] = (rb, gbatch, r['K'], r)
    osc = None
    if rank == 0 and not args.no_configs and robot == 'atlas':
        osc = run_osc_config(env, K, args.warmup)

    if rank != 0:
        if world > 1:
            env.dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the QP solve), fp64 pipe + HBM ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'
    fp64_peak = C_double_peak(env.lib, env.local_rank)
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass

    def traffic_and_flops(rb, Bx):
        t = traffic_tab.get(rb) or {}
        return (t.get('solve_stage_bytes') if t.get('batch') == Bx else None), t.get('flops_per_solve')

    def finish(rb, r, want_cpu):
        cpu, mean_flops, src = None, None, None
        if want_cpu:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or (2048 * max(1, min(threads, 16)) if rb == robot else 256 * max(1, min(threads, 16)))
            cpu, mean_flops = cpu_baseline_for(rb, sample)
            src = "instrumented CPU oracle, mean over %d states of the same generator" % sample
        traffic, fl = traffic_and_flops(rb, r['B'])
        if mean_flops is None and fl:
            mean_flops, src = float(fl), "profiles/r2_traffic.json (oracle count recorded with the ncu capture)"
        if mean_flops is None:
            mean_flops, src = r['spec'].algorithmic_flops_assembly(), "assembly only (lower bound; no oracle count available)"
        return cpu, roofline_of(r, mean_flops, src, fp64_peak, hbm_peak, hbm_src, traffic)

    cpu, roofline = finish(robot, res, not args.no_cpu_baseline)
    st = res['stats']
    cfg = make_config(robot, B, world, args.scaling, res['spec'].qp_dims)
    if args.scaling == 'strong':
        cfg['global_batch'] = args.batch
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": res['ms_per_step'], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64 (update_system in f32: the optional fp32 mode)" if fp32_run else "f64", "data": "synthetic",
            "config": cfg, "workspace_gb_per_gpu": res['ws_gb'],
            "kernels_ms": {"pnc_update_system": res['ms_dyn'], "pnc_ihwbc_solve": res['ms_qp']},
            "solver": {"status_counts": [int(x) for x in st[:5]], "mean_active_set_iterations": st[5] / st[6]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e_of(res, world_states), "gpu_launches": res['launches'],
            "clocks": res['clocks'], "host_cpus_bound_per_rank": env.affinity}
    f32 = res.get('fp32')
    if f32:
        line["fp32_mode"] = {
            "what": "pnc_set_precision(PNC_PREC_FP32_DYNAMICS) on the same batch: k_update_system in fp32 arithmetic, "
                    "factorisations and active-set iterations fp64",
            "value": world_states / (f32['ms_per_step'] * 1e-3), "unit": UNIT, "ms_per_step": f32['ms_per_step'],
            "kernels_ms": {"pnc_update_system": f32['ms_dyn']},
            "rel_err_vs_fp64": {"states": f32['n'], "median": f32['median'], "p99": f32['p99'], "max": f32['max'],
                                "frac_within_1e-4": f32['within_1e4'] / max(1, f32['n']),
                                "same_final_active_set": f32['same_final_active_set'],
                                "status_mismatch": f32['status_mismatch'], "scope": "rank 0's shard"}}
    if extra or osc:
        configs = {}
        if osc:
            configs['config1_manipulator_osc_1k'] = osc
        for key, (rb, gbatch, Kc, r) in extra.items():
            c_cpu, c_roof = finish(rb, r, not args.no_cpu_baseline)
            s2 = r['stats']
            configs[key] = {"workload": "%s, %d states in total sharded over %d GPU(s)" % (WORKLOADS[rb], gbatch, world),
                            "metric": METRIC, "unit": UNIT, "value": gbatch * Kc / (r['ms_total'] * 1e-3), "scaling": "strong",
                            "global_batch": gbatch, "batch_per_gpu": r['B'], "n_gpus": world, "steps": Kc,
                            "ms_per_step": r['ms_per_step'], "qp_dims_n_p_m": list(r['spec'].qp_dims),
                            "kernels_ms": {"pnc_update_system": r['ms_dyn'], "pnc_ihwbc_solve": r['ms_qp']},
                            "solver": {"status_counts": [int(x) for x in s2[:5]], "mean_active_set_iterations": s2[5] / s2[6]},
                            "roofline": c_roof, "cpu_baseline": c_cpu, "e2e": e2e_of(r, gbatch), "gpu_launches": r['launches'],
                            "clocks": r['clocks'], "workspace_gb_per_gpu": r['ws_gb']}
        line["configs"] = configs
    print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


def C_double_peak(lib, device):
    import ctypes as C
    v = C.c_double(0)
    rc = lib.pnc_measure_fp64_peak(device, C.byref(v))
    return float(v.value) if rc == 0 else None


if __name__ == '__main__':
    main()
