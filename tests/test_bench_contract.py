"""The bench.py contract that can be checked without a GPU: the reference arm's JSON line, and the product arm's refusal
to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), cwd=ROOT, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--cores', '2')
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'].startswith('oamix+oaloss images/sec') and d['unit'] == 'images/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['vs_baseline'] is None and d['data'] == 'synthetic' and 'workload' in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] == 2 and cb['value'] == d['value'] and 'frames' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour on a machine without a GPU')
def test_product_arm_refuses_to_run_without_a_gpu():
    out = _run('--steps', '1', '--warmup', '1', '--no-cpu-baseline')
    assert out.returncode != 0 and 'no CUDA device' in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith('{')]      # no number is printed
