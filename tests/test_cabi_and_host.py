"""CPU-side checks: the C-ABI library loads and exports every symbol include/pnc_b200.h declares
(no compute without a GPU), the model compiler / problem tables are consistent, and the
multi-GPU host logic works over gloo with world_size 2."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest

from pypnc_b200 import _cabi, robots

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return C.CDLL(_cabi.LIB_PATH)


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, 'include', 'pnc_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(pnc_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol %s declared in include/pnc_b200.h is not exported" % name
    assert declared == set(_cabi.EXPORTED_SYMBOLS)


def test_version_and_launch_counter_without_gpu(lib):
    lib.pnc_version.restype = C.c_char_p
    assert b'sm_100a' in lib.pnc_version()
    lib.pnc_launch_count.restype = C.c_int64
    assert lib.pnc_launch_count() == 0


def test_precision_and_termination_setters_validate_their_argument(lib):
    """Process-wide mode setters (no device needed): defaults, round trip, PNC_E_ARG on anything else."""
    assert lib.pnc_get_precision() == 0 and lib.pnc_get_termination() == 0
    assert lib.pnc_set_precision(1) == 0 and lib.pnc_get_precision() == 1
    assert lib.pnc_set_precision(2) < 0 and lib.pnc_set_precision(-1) < 0 and lib.pnc_get_precision() == 1
    assert lib.pnc_set_precision(0) == 0 and lib.pnc_get_precision() == 0
    assert lib.pnc_set_termination(7) < 0 and lib.pnc_get_termination() == 0


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a GPU the compute entry points fail loudly (PNC_E_NOGPU), they do not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    spec = robots.atlas_problem()
    desc, keep = _cabi.make_model_desc(spec.model)
    h = C.c_void_p()
    assert lib.pnc_model_create(C.byref(desc), 0, C.byref(h)) == -4
    v = C.c_double(0)
    assert lib.pnc_measure_fp64_peak(0, C.byref(v)) == -4


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under pypnc_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, 'pypnc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, flags=re.M), f
                assert 'pnc_oracle' not in txt and 'orc_' not in txt, f


@pytest.mark.parametrize("name,dims", [('atlas', (48, 6, 96)), ('laikago', (30, 6, 24)), ('draco3', (45, 8, 36)),
                                       ('draco3_lb', (32, 8, 60)), ('valkyrie', (44, 6, 88))])
def test_problem_tables_match_the_survey(name, dims):
    spec = robots.PROBLEMS[name]()
    assert spec.qp_dims == dims
    d, keep = _cabi.make_problem_desc(spec)
    assert d.ntask == len(spec.tasks) and d.ndes == spec.ndes
    assert len(spec.frames) <= 12


def test_shard_range_partitions_the_batch():
    from pypnc_b200.sharding import shard_range
    for total in (0, 1, 7, 65536, 262144 + 3):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from pypnc_b200.sharding import shard_range, solver_stats, gather_stats, max_over_ranks
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    total = 1001
    lo, hi = shard_range(total, rank, world)
    rng = np.random.default_rng(5)
    status = rng.integers(0, 5, total)
    iters = rng.integers(1, 30, total)
    st = gather_stats(solver_stats(status[lo:hi], iters[lo:hi]))
    mx = max_over_ranks(10.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, st.tolist(), mx, solver_stats(status, iters).tolist()))


def test_stats_gather_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, st, mx, full in res:
        assert st == full          # every rank sees the whole-job stats
        assert mx == 11.0          # max over ranks


def test_manager_record_sizes_match_header():
    import re
    from pypnc_b200 import managers
    hdr = open(os.path.join(ROOT, 'include', 'pnc_b200.h')).read()
    assert int(re.search(r'#define PNC_SWING_REC (\d+)', hdr).group(1)) == managers.SWING_REC
    assert int(re.search(r'#define PNC_FB_REC (\d+)', hdr).group(1)) == managers.FB_REC
    assert int(re.search(r'#define PNC_HAND_REC (\d+)', hdr).group(1)) == managers.HAND_REC
