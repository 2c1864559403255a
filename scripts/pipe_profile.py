"""Loader-loop profile on device-resident frames (GPU box): ms per step and host microseconds per pipeline phase.

    python scripts/pipe_profile.py [steps] [group_batches] [with_loss 0/1]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix, ContrastiveLossPlus  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
group = int(sys.argv[2]) if len(sys.argv) > 2 else 4
with_loss = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(bench.POOL)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
mix.group_batches = group
x, labels = bench.make_roi_set()
xd = x.to(dev).requires_grad_(True)
ld = labels.to(dev)
fn = ContrastiveLossPlus(**bench.LOSS_CFG)


def batches(n):
    for i in range(n):
        j = (i * 2) % bench.POOL
        yield [dict(img=imgs[(j + b) % bench.POOL], gt_bboxes=gts[(j + b) % bench.POOL]) for b in range(2)]


def run(n):
    for _ in mix.iter_batches(batches(n)):
        if with_loss:
            xd.grad = None
            fn(xd, ld).backward()


np.random.seed(7)
run(8)
torch.cuda.synchronize()
for n in (steps, 3 * steps):
    mix.pipe_profile = {}
    np.random.seed(1000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%d steps, group %d, loss %d: %.3f ms/step (%.0f images/s)' % (n, group, with_loss, dt / n * 1e3, 2 * n / dt))
    print('   host us per step: ' + ', '.join('%s %.0f' % (k, v / n * 1e6) for k, v in mix.pipe_profile.items()
                                               if not k.endswith('.max')))

# arrival time of every batch of a 20-step loop (ms since the loop started)
np.random.seed(1000)
torch.cuda.synchronize()
t0 = time.perf_counter()
marks = []
for _ in mix.iter_batches(batches(20)):
    if with_loss:
        xd.grad = None
        fn(xd, ld).backward()
    marks.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
print('batch arrival ms: ' + ' '.join('%.2f' % m for m in marks) + '  | end %.2f' % ((time.perf_counter() - t0) * 1e3))
