"""Per-op timing of the OA-Mix kernels: forces every region of every step to one op kind and reports the
CUDA-event time of each kernel kind (oadg_oamix_execute_profiled) per lane step at 1024x2048."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402
from oadg_b200.oamix import _invert_affine  # noqa: E402

dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(2)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
H, W = 1024, 2048


def forced(name):
    def sample(self, gt, img_size, gt_int=None):
        if name in ('autocontrast', 'equalize'):
            return (name,)
        if name == 'posterize':
            return (name, 2)
        if name in ('color', 'sharpness', 'contrast', 'brightness'):
            return (name, 1.3)
        if name == 'invert':
            return (name, 1, -1)
        where, geo = name.split('_', 1)
        lvl = 7.0
        if where == 'bg':
            return ('bg_affine', _invert_affine(OAMix._forward_affine(geo, lvl, False, img_size, None, img_size)))
        chain = []
        for k, b in enumerate(gt):
            x1, y1, x2, y2 = int(b[0]), int(b[1]), int(b[2]), int(b[3])
            fwd = OAMix._forward_affine(geo, lvl, False, (x2 - x1 + 1, y2 - y1 + 1), ((x1 + x2) / 2., (y1 + y2) / 2.), img_size)
            chain.append((k, _invert_affine(fwd)))
        return ('bbo_affine', chain)
    return sample


names = ['posterize', 'autocontrast', 'equalize', 'bg_translate_x', 'bg_shear_x', 'bg_shear_y', 'bg_rotate',
         'bbo_translate_x', 'bbo_shear_x', 'bbo_rotate', 'invert', 'color', 'sharpness']
if len(sys.argv) > 1:
    names = sys.argv[1:]
for name in names:
    mix = OAMix(**dict(bench.OAMIX_CFG, version='augmix.all'))
    mix._sample_op = forced(name).__get__(mix)
    np.random.seed(0)
    mix.oamix_batch(imgs, gts)           # warm
    prof = {}
    for it in range(3):
        mix.oamix_batch(imgs, gts, profile=prof)
    lanes = prof['step_bytes'] / (2 * 3 * H * W)
    ms = {k[:-3]: round(v, 3) for k, v in prof.items() if k.endswith('_ms') and v > 0}
    n = {k[:-2]: v for k, v in prof.items() if k.endswith('_n') and v > 0}
    gbs = prof['step_bytes'] / (prof['step_ms'] / 1e3) / 1e9
    print('%-16s lanes %3d  step %.1f us/lane (%.0f GB/s)  kernels ms %s  launches %s' %
          (name, lanes, prof['step_ms'] * 1e3 / lanes, gbs, ms, n), flush=True)
