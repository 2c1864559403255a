"""Mismatch statistics of the edge-case views of tests/test_gpu_oamix.py (GPU box): differing values / total, max |d|,
and of the saliency scores on the randomized sweep's boxes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from conftest import OAMIX_CFG, sampler_cfg  # noqa: E402
from oracle import oamix_np, synth  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

dev = torch.device('cuda:0')
for (h, w, gt, seed) in [(97, 131, np.zeros((0, 4), np.float32), 1),
                         (97, 131, np.float32([[10, 10, 12, 40], [50, 20, 90, 23]]), 2),
                         (128, 96, np.float32([[0, 0, 96, 128]]), 3),
                         (65, 67, np.float32([[60.5, 50.2, 66.9, 64.7], [1.2, 1.9, 30.3, 20.8]]), 4)]:
    img, _ = synth.make_image(seed, h, w, 0)
    cfg = dict(OAMIX_CFG, version='augmix.all')
    worst = (0, 0)
    for rep in range(8):
        np.random.seed(seed + 100 * rep)
        ref, plan = oamix_np.oamix_view(img, gt, **sampler_cfg(cfg))
        np.random.seed(seed + 100 * rep)
        out = OAMix(**cfg).oamix_batch([torch.from_numpy(img).to(dev)], [gt])[0][0].cpu().numpy()
        d = np.abs(out.astype(np.int16) - ref.astype(np.int16))
        worst = max(worst, (int((d > 0).sum()), int(d.max())))
    print('edge case %dx%d gt=%d: worst of 8 plans: %d of %d values differ (%.2e), max |d| %d' % (
        h, w, len(gt), worst[0], out.size, worst[0] / out.size, worst[1]))
# saliency: bench frames
t = OAMix()
errs = []
for s in range(6):
    img, gt = synth.make_image(s)
    got = t.saliency_scores([torch.from_numpy(img).to(dev)], [gt])[0]
    ref = oamix_np.fg_scores(img, gt)
    errs += [abs(a - b) for a, b in zip(got, ref) if a >= 0]
print('saliency |d| on 6 bench frames (48 boxes): max %.2e, mean %.2e' % (max(errs), float(np.mean(errs))))
small, sgt = synth.make_image(7, 96, 160, 3)
gts = np.float32([[130, 36, 136, 59], [141, 62, 144, 69], [98, 30, 106, 53]])
got = t.saliency_scores([torch.from_numpy(small).to(dev)], [gts])[0]
ref = oamix_np.fg_scores(small, gts)
print('saliency on the tiny boxes of the 96x160 frame:', [(round(a, 4), round(b, 4)) for a, b in zip(got, ref)])
