"""Do two half-width chain launches on two streams beat one full-width launch after the other?
usage: OADG_CTAS_PER_SM=2 python scripts/overlap_probe.py 2   |   python scripts/overlap_probe.py 1"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_batches = 24
dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(bench.POOL)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mixes = [OAMix(**bench.OAMIX_CFG) for _ in range(n_streams)]
streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
outs = [[torch.empty_like(imgs[0]) for _ in range(2)] for _ in range(n_streams)]

# plans sampled up front (same seeds for every mode), so only the kernels are timed
np.random.seed(1000)
plans = []
for i in range(n_batches):
    j = (2 * i) % bench.POOL
    b = [imgs[j], imgs[j + 1]]
    g = [np.asarray(x, np.float32).reshape(-1, 4) for x in gts[j:j + 2]]
    sc = mixes[0].saliency_scores(b, g)
    plans.append((mixes[0].sample_plan([(1024, 2048)] * 2, g, sc).blob, b))


def run():
    for i, (blob, b) in enumerate(plans):
        k = i % n_streams
        mixes[k].execute(blob, b, outs=outs[k], stream=streams[k])


for _ in range(3):
    run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    run()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print('streams=%d OADG_CTAS_PER_SM=%s: %.3f ms per %d batches (%.1f us per batch)' % (
    n_streams, os.environ.get('OADG_CTAS_PER_SM', '-'), dt * 1e3, n_batches, dt / n_batches * 1e6))
