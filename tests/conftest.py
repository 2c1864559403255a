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
