"""Build libOADG.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

    python -m oadg_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libOADG.so')
STAMP = os.path.join(HERE, '.libOADG.stamp')
SOURCES = ['api.cu', 'saliency.cu', 'oamix.cu', 'oamix_sampler.cpp', 'oaloss.cu', 'oaloss_tc.cu', 'jsd.cu', 'peer.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-ffp-contract=off', '-shared', '-fmad=false', '-Xptxas', '-v',
         '-I', os.path.join(ROOT, 'include'), '-I', CSRC] + os.environ.get('OADG_NVCC_EXTRA', '').split()
# -fmad=false: the OA-Mix float stages must not contract a*b+c (oamix_math.h); kernels that
# want FMA (the loss GEMMs) call fmaf() explicitly.


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = _sources() + [os.path.join(ROOT, 'include', 'oadg.h')]
    files += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.h', '.cuh'))]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build_test_variant(path, extra_flags):
    """A second library built from the same sources with extra -D flags (test infrastructure: the FFMA cross-check
    of the OA-Loss, tests/test_gpu_oaloss.py); rebuilt when older than the product library."""
    build()
    if os.path.exists(path) and os.path.getmtime(path) >= os.path.getmtime(LIB):
        return path
    cmd = [NVCC] + [f for f in FLAGS if f not in ('-Xptxas', '-v')] + list(extra_flags) + _sources() + ['-o', path]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building %s' % path)
    return path


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    cmd = [NVCC] + FLAGS + _sources() + ['-o', LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libOADG.so')
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, 'w') as fh:
        fh.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
