"""Run a few un-profiled OA-Mix launches (for ncu): python scripts/run_one.py [views_per_launch] [n_launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

dev = torch.device('cuda:0')
group = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = [bench.make_image(s) for s in range(max(group, 8))]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
np.random.seed(1000)
for i in range(n):
    mix.oamix_batch(imgs[:group], gts[:group])
torch.cuda.synchronize()
