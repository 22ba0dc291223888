"""bench.py's contract on a machine without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads, and the product arm refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from _util import ROOT


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--size', '64', '--steps', '1', '--warmup', '0')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'sample' in d['cpu_baseline']
    assert d['steps'] == 1 and d['warmup'] == 0          # exactly the K / W the driver asked for
    # the reference arm reports on the product arm's config object (same function builds both)
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ns = argparse.Namespace(size=64, batch=2, frames=4, gpus=1, generator_only=False)
    assert d['config'] == bench.bench_config(ns) and d['metric'] == bench.METRIC


def test_reference_arm_other_ranks_exit_without_work():
    r = _run('--impl', 'reference', '--gpus', '2', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run('--steps', '1', '--warmup', '0', '--no-cpu-baseline')
    assert r.returncode != 0
    assert 'no CPU fallback' in (r.stderr + r.stdout)
