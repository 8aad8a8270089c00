"""bench.py contract checks that need no GPU: the reference arm (--impl reference) prints one JSON line with the keys the
driver reads, and both arms build their `config` object with the same function (same key set; a key-set mismatch made
the driver report same_config: false in round 1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_config_keys(oracle_mod):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--cpu-sample', '256'], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in line, k
    assert line['impl'] == 'reference' and line['metric'] == 'wbc_solves_per_sec' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value'] == line['cpu_baseline']['value']
    sys.path.insert(0, ROOT)
    import bench
    ours = bench.make_config('atlas', 65536, 1, 'weak', (48, 6, 96))
    assert set(ours) == set(line['config'])
    assert ours['workload'] == line['config']['workload'] and ours['global_batch'] == line['config']['global_batch']
    assert 'model' not in ours  # the hot-path tier's config names a workload, not a model
