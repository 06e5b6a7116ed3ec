"""bench.py contract pieces that run without a GPU: the reference arm (`--impl reference`: the reference's own CPU
implementation of the path on the host cores, rank 0 only), and the product arm's refusal to run without a CUDA device
(there is no CPU fallback).  The GPU legs of bench.py are exercised by the driver and by tests/test_gpu_parity.py
(`test_bench_workload_parity` runs the exact bench workload against the oracle)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, 'bench.py')


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    frames = 2                       # (the default, 32 frames per step, takes minutes on a small CI box)
    r = run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-frames', str(frames)])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert d['impl'] == 'reference'
    assert 'reenacted frames/sec' in d['metric'] and 'reenacted frames/sec' in base['metric']
    assert d['unit'] == 'frames/s' and d['higher_is_better'] is True and d['data'] == 'synthetic'
    assert d['steps'] == 1 and d['n_gpus'] == 1 and d['vs_baseline'] is None
    assert d['value'] > 0 and abs(d['value'] - frames / (d['ms_per_step'] * 1e-3)) <= 1e-6 * d['value']
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] == (os.cpu_count() or 1) and cb['value'] == d['value']
    assert cb['sample']
    # baseline/_ref (the unmodified reference) is what runs wherever the copy exists; the oracle port is the fallback
    if os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'libs')):
        assert cb['kind'] == 'reference'
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
            env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU refusal')
def test_product_arm_refuses_to_run_without_a_gpu():
    r = run(['--steps', '1', '--warmup', '3'], timeout=300)
    assert r.returncode != 0
    assert 'CUDA' in (r.stderr + r.stdout) and 'no CPU fallback' in (r.stderr + r.stdout)
