"""Small driver for ncu: `python scripts/profile_path.py oamix|loss [iters] [frames per launch]` runs the hot path a few
times on the bench workload (1024x2048 frames / [2088,256] embeddings) with nothing else in the process.  8 frames per
launch is what the loader loop's full groups execute (4 steps of 2 frames in one plan / chain launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix, ContrastiveLossPlus  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'oamix'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda:0')
if what == 'oamix':
    per = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    frames = [bench.make_image(s) for s in range(max(4, 2 * per))]
    imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
    gts = [g for _, g in frames]
    mix = OAMix(**bench.OAMIX_CFG)
    np.random.seed(1000)
    for i in range(iters):
        j = (per * i) % len(imgs)
        mix.oamix_batch(imgs[j:j + per], gts[j:j + per])
    torch.cuda.synchronize()
    print('oamix launches/iter ~', mix.last_launches)
else:
    x, labels = bench.make_roi_set()
    xd = x.to(dev).requires_grad_(True)
    fn = ContrastiveLossPlus(**bench.LOSS_CFG)
    for i in range(iters):
        xd.grad = None
        fn(xd, labels.to(dev)).backward()
    torch.cuda.synchronize()
    print('loss launches', fn.stats)
