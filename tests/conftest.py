import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

OAMIX_CFG = dict(num_views=2, keep_orig=True, severity=10, random_box_ratio=(3, 1 / 3),
                 random_box_scale=(0.01, 0.1), oa_random_box_scale=(0.005, 0.1),
                 oa_random_box_ratio=(3, 1 / 3), spatial_ratio=4, sigma_ratio=0.3)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def libpath():
    """Build libOADG.so if it is missing (nvcc cross-compiles without a GPU)."""
    from oadg_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def cuda(libpath):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def sampler_cfg(cfg):
    """Keys the oracle's sample_plan / the product ctor share."""
    return {k: v for k, v in cfg.items() if k not in ('num_views', 'keep_orig', 'severity')}


def build_hostsim():
    """(Re)build tests/hostsim/libhostsim.so when it is missing or older than its sources; returns its path.
    Test infrastructure only: the device bodies + host scheduler compiled for the CPU."""
    import subprocess
    src = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim.cpp')
    lib = os.path.join(ROOT, 'tests', 'hostsim', 'libhostsim.so')
    deps = [src, os.path.join(ROOT, 'include', 'oadg.h')] + [
        os.path.join(ROOT, 'oadg_b200', 'csrc', f) for f in ('oamix_math.h', 'oamix_body.h', 'oamix_exec.h', 'oamix_tile.h')]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-std=c++17', '-shared', '-fPIC',
                               '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(ROOT, 'oadg_b200', 'csrc'),
                               src, '-o', lib])
    return lib


# ---- seeded inputs of the consistency-loss goldens (tests/golden/consistency.npz, scripts/make_golden_f2.py)
def f2_inputs(kind):
    """Same seeded tensors as scripts/make_golden_f2.py::inputs."""
    import torch
    g = torch.Generator().manual_seed({'roi': 11, 'rpn': 12, 'reg': 13}[kind])
    if kind == 'roi':
        n = 256
        pred = torch.randn(n, 9, generator=g, dtype=torch.float64) * 3
        half = torch.randint(0, 9, (n // 2,), generator=g)
        return pred, torch.cat([half, half]), torch.ones(n, dtype=torch.float64), float(n)
    if kind == 'rpn':
        n = 4096
        pred = torch.randn(n, 1, generator=g, dtype=torch.float64) * 4
        half = torch.randint(0, 2, (n // 2,), generator=g)
        w = (torch.rand(n // 2, generator=g) < 0.25).double()
        return pred, torch.cat([half, half]), torch.cat([w, w]), 512.0
    n = 512
    pred = torch.randn(n, 4, generator=g, dtype=torch.float64)
    tgt = torch.randn(n // 2, 4, generator=g, dtype=torch.float64)
    w = (torch.rand(n // 2, 4, generator=g) < 0.8).double()
    return pred, torch.cat([tgt, tgt]), torch.cat([w, w]), float(n)


CE_CFG = {
    'roi': dict(type='CrossEntropyLossPlus', use_sigmoid=False, loss_weight=1.0, num_views=2,
                additional_loss='jsdv1_3_2aug', lambda_weight=10, wandb_name='roi_cls', log_pos_ratio=True),
    'rpn': dict(type='CrossEntropyLossPlus', use_sigmoid=True, loss_weight=1.0, num_views=2,
                additional_loss='jsdv1_3_2aug', lambda_weight=0.1, wandb_name='rpn_cls'),
}


