"""Where the host time of one bench step goes: cProfile over `python scripts/host_profile.py [steps]` (GPU box)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix, ContrastiveLossPlus  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(8)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
x, labels = bench.make_roi_set()
xd = x.to(dev).requires_grad_(True)
ld = labels.to(dev)
fn = ContrastiveLossPlus(**bench.LOSS_CFG)
outs = [torch.empty_like(imgs[0]) for _ in range(2)]


def step(i):
    j = (2 * i) % 8
    mix.oamix_batch(imgs[j:j + 2], gts[j:j + 2], outs=outs, inputs_ready=True)
    xd.grad = None
    fn(xd, ld).backward()


np.random.seed(1)
for i in range(5):
    step(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(steps):
    step(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(35)
