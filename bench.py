#!/usr/bin/env python
"""bench.py -- WBC solves/s (robot_system update + IHWBC QP), Atlas DCM-walking task set.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--robot atlas] [--batch 65536]

One "step" = one pass of the hot path over one batch of synthetic states: pnc_update_system
(FK, Jacobians, CoM, CRBA, RNEA terms) followed by pnc_ihwbc_solve (task/contact evaluation, QP
assembly, dual active-set solve, torque recovery) for `batch` independent robot states per GPU.
Multi-GPU: one process per GPU (torchrun), the batch dimension is sharded, no data-path
collective; only the final stats gather uses NCCL ("scaling": "weak": `batch` states per GPU).

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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--robot', default='atlas')
    ap.add_argument('--batch', type=int, default=65536, help='states per GPU')
    ap.add_argument('--cpu-sample', type=int, default=0, help='states in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--same-state', action='store_true', help='diagnostic: every state of the batch is state 0')
    return ap.parse_args()


def workload_name(robot, batch):
    names = {'atlas': 'Atlas DCM-walking IHWBC (atlas_tci_container task set, 2 surface contacts, torque limits)',
             'laikago': 'Laikago stance IHWBC, 4 point contacts',
             'draco3': 'Draco3 IHWBC with rolling-joint internal constraint',
             'draco3_lb': 'Draco3 lower body IHWBC, internal constraint + torque limits',
             'valkyrie': 'Valkyrie full-body IHWBC (synthesised task set)'}
    return "%s, batch %d per GPU" % (names.get(robot, robot), batch)


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
            self._stop_evt.wait(0.01)

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
def cpu_reference_run(robot, sample, threads, repeats=1, seed=1):
    """Times the CPU oracle loop (update + assemble + QuadProg++ port) on `sample` states."""
    from pypnc_b200 import robots
    from oracle import oracle
    spec = robots.PROBLEMS[robot]()
    st = robots.synth_states(robot, spec, sample, seed=seed)
    op_strict = oracle.OracleProblem(spec, fast=True)
    act = op_strict.task_actuals(st)
    des, w, rfz = robots.synth_desireds(robot, spec, st, act, seed=seed)
    best, out = None, None
    for _ in range(max(1, repeats)):
        out = op_strict.tick_batch(st, des, w, rfz, nthreads=threads)
        best = out['seconds'] if best is None else min(best, out['seconds'])
    return dict(solves_per_s=sample / best, seconds=best, mean_flops=float(out['flops'].mean()),
                ok=int((out['status'] == 0).sum()), mean_iters=float(out['iters'].mean()))


def run_reference(args):
    """--impl reference: the reference path's CPU restatement on the host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.cpu_sample or 4096 * max(1, min(threads, 16))
    # inputs are generated once, outside the timed steps
    from pypnc_b200 import robots
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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.robot, args.batch), "robot": args.robot,
                       "batch_per_gpu": args.batch},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d states per step of the same synthetic generator (seed 1), oracle -O3 "
                                       "build, %d host threads over contiguous shards" % (sample, threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from pypnc_b200 import robots, engine, _cabi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    lib = _cabi.load_library()

    robot, B = args.robot, args.batch
    spec = robots.PROBLEMS[robot]()
    dm = engine.Model(spec.model, local_rank)
    pb = engine.Problem(dm, spec)
    ws = engine.Workspace(dm, B, spec.frames)

    # ---- synthetic inputs, seed + rank per shard ----
    st = robots.synth_states(robot, spec, B, seed=1 + rank)
    if args.same_state:
        st = {k: np.ascontiguousarray(np.repeat(v[:1], B, axis=0)) for k, v in st.items()}
    keys = ('base_pos', 'base_quat', 'base_lin_vel', 'base_ang_vel', 'joint_pos', 'joint_vel')
    d_st = [torch.from_numpy(np.ascontiguousarray(st[k])).to(dev) for k in keys]
    ws.update_system(*d_st)
    torch.cuda.synchronize()
    des, w, rfz = robots.synth_desireds(robot, spec, st, engine.task_actuals_from_workspace(spec, ws), seed=1 + rank)
    if args.same_state:
        des, w, rfz = [np.ascontiguousarray(np.repeat(a[:1], B, axis=0)) for a in (des, w, rfz)]
    d_des, d_w, d_rfz = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (des, w, rfz)]

    nact, nc, nv, mw = spec.nact, pb.nc, spec.model.nv, pb.mw
    f64 = dict(dtype=torch.float64, device=dev)
    o_trq, o_acc = torch.empty((B, nact), **f64), torch.empty((B, nact), **f64)
    o_rf = torch.empty((B, nc), **f64)
    o_status = torch.empty(B, dtype=torch.int32, device=dev)
    o_iters = torch.empty(B, dtype=torch.int32, device=dev)
    o_active = torch.empty((B, mw), dtype=torch.int64, device=dev)
    P = lambda t: engine._ptr(t)
    stream = torch.cuda.current_stream()
    sp = engine.vp(stream.cuda_stream)

    def step():
        rc = lib.pnc_update_system(ws.h, B, *[P(t) for t in d_st], sp)
        rc |= lib.pnc_ihwbc_solve(pb.h, ws.h, B, P(d_des), P(d_w), P(d_rfz), P(o_trq), P(o_acc), P(o_rf), None,
                                  P(o_status), P(o_iters), P(o_active), sp)
        assert rc == 0, rc

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = lib.pnc_launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
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
    barrier()
    launches = lib.pnc_launch_count() - launches0
    clocks = sampler.finish()
    ms_total = t_start.elapsed_time(t_end)
    ms_dyn = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    ms_qp = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    ms_per_step = ms_total_max / K
    value = world * B * K / (ms_total_max * 1e-3)

    status = o_status.cpu().numpy()
    iters = o_iters.cpu().numpy()
    stats = torch.tensor([float((status == s).sum()) for s in range(5)] + [float(iters.sum()), float(B)],
                         dtype=torch.float64, device=dev)
    if world > 1:  # the only collective of the run: a stats gather
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        stats = torch.stack(gathered).sum(0)
    stats = stats.cpu().numpy()

    # ---- end-to-end through the host-buffer C ABI call (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_in = [pin(st[k]) for k in keys] + [pin(des), pin(w), pin(rfz)]
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

        for _ in range(2):
            tick_host()
        barrier()
        Ke = max(2, min(K, 5))
        t0 = time.perf_counter()
        for _ in range(Ke):
            tick_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = sum(t.numel() * t.element_size() for t in h_in)
        d2h = sum(t.numel() * t.element_size() for t in (h_trq, h_acc, h_rf, h_status, h_iters, h_active))
        e2e = {"value": world * B * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": Ke, "ms_per_step": 1e3 * dt / Ke,
               "api": "pnc_wbc_tick_host (pinned host buffers in, host buffers out, synchronous)"}
        # the host path must agree with the device-resident path
        assert (h_status.numpy() == status).all()
        assert np.array_equal(h_trq.numpy(), o_trq.cpu().numpy(), equal_nan=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
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
    fp64_peak = C_double_peak(lib, local_rank)
    cpu = None
    mean_flops = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        sample = args.cpu_sample or 2048 * max(1, min(threads, 16))
        c1 = cpu_reference_run(robot, min(sample, 2048), 1)
        cN = cpu_reference_run(robot, sample, threads) if threads > 1 else c1
        mean_flops = cN['mean_flops']
        cpu = {"value": cN['solves_per_s'], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "%d states of the same synthetic generator (seed 1) on %d host threads; single thread: "
                         "%.0f solves/s on %d states; oracle/ C restatement, gcc -O3 -march=x86-64-v3"
                         % (sample, threads, c1['solves_per_s'], min(sample, 2048)),
               "single_thread_value": c1['solves_per_s']}
    # algorithmic bytes per solve: inputs + outputs of one tick
    na = spec.model.na
    bytes_in = (3 + 4 + 3 + 3 + 2 * na + spec.ndes + len(spec.tasks) + len(spec.contacts)) * 8
    bytes_out = (2 * nact + nc) * 8 + 8 * mw + 8
    alg_bytes = bytes_in + bytes_out
    traffic = None  # DRAM bytes of the solve stage per launch, from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as f:
            tj = json.load(f)
        if tj.get('robot') == robot and tj.get('batch') == B:
            traffic = tj['solve_stage_bytes']
        if mean_flops is None and tj.get('robot') == robot:
            mean_flops = float(tj['flops_per_solve'])  # the oracle's count, recorded with the capture
    except Exception:
        pass
    if mean_flops is None:
        mean_flops = spec.algorithmic_flops_assembly()  # lower bound: assembly only (no oracle count on record)
    qp_tflops = B * mean_flops / (ms_qp * 1e-3) * 1e-12
    roofline = {"bound": "fp64", "kernel": "pnc_ihwbc_solve = k_task_prepare + k_ihwbc_setup + k_ihwbc_reg + k_ihwbc_finish",
                "achieved": qp_tflops,
                "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": qp_tflops / fp64_peak if fp64_peak else None, "traffic": traffic,
                "peak_source": "measured in this run: DFMA-chain microbenchmark, all SMs (pnc_measure_fp64_peak)",
                "flops_per_solve": mean_flops, "kernel_ms": ms_qp,
                "hbm": {"achieved": B * alg_bytes / (ms_per_step * 1e-3) * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": B * alg_bytes / (ms_per_step * 1e-3) * 1e-9 / hbm_peak,
                        "bytes_per_solve": alg_bytes, "peak_source": hbm_src}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(robot, B), "robot": robot, "batch_per_gpu": B,
                       "global_batch": B * world, "qp_dims_n_p_m": list(spec.qp_dims),
                       "l2": "inputs + per-state workspace (%.2f GB per GPU) exceed the 126 MB L2" % (ws.nbytes / 1e9),
                       "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world},
            "kernels_ms": {"pnc_update_system": ms_dyn, "pnc_ihwbc_solve": ms_qp},
            "solver": {"status_counts": [int(x) for x in stats[:5]], "mean_active_set_iterations": stats[5] / stats[6]},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def C_double_peak(lib, device):
    import ctypes as C
    v = C.c_double(0)
    rc = lib.pnc_measure_fp64_peak(device, C.byref(v))
    return float(v.value) if rc == 0 else None


if __name__ == '__main__':
    main()
