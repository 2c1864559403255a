"""Host time per phase of the pipelined loader loop (OAMix.iter_batches + the loss with loss.item() per step) on the
bench workload.  usage: python scripts/e2e_profile.py [steps] [threaded 0|1]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix, ContrastiveLossPlus  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
threaded = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(bench.POOL)]
host = [torch.from_numpy(f).pin_memory() for f, _ in frames]
gts = [g for _, g in frames]
x, labels = bench.make_roi_set()
x_host, labels_dev = x.pin_memory(), labels.to(dev)
mix, loss_fn = OAMix(**bench.OAMIX_CFG), ContrastiveLossPlus(**bench.LOSS_CFG)


def batches(n):
    for i in range(n):
        j = (i * bench.BS) % bench.POOL
        yield [dict(img=host[(j + b) % bench.POOL].numpy(), gt_bboxes=gts[(j + b) % bench.POOL]) for b in range(bench.BS)]


def run(n, prof=None):
    mix.pipe_profile = prof
    t_loss = t_item = 0.0
    for _ in mix.iter_batches(batches(n), threaded=threaded):
        t0 = time.perf_counter()
        xd = x_host.to(dev, non_blocking=True).requires_grad_(True)
        loss = loss_fn(xd, labels_dev)
        loss.backward()
        t1 = time.perf_counter()
        loss.item()
        t2 = time.perf_counter()
        t_loss += t1 - t0
        t_item += t2 - t1
    return t_loss, t_item


np.random.seed(7)
run(5)
np.random.seed(1000)
prof = {}
torch.cuda.synchronize()
t0 = time.perf_counter()
t_loss, t_item = run(steps, prof)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print('%d steps: %.3f ms/step wall (%.0f images/s)' % (steps, wall / steps * 1e3, bench.BS * steps / wall))
for k, v in prof.items():
    if not k.endswith('.max'):
        print('  %-18s %7.1f us/step (max %.0f)' % (k, v / steps * 1e6, prof[k + '.max'] * 1e6))
print('  %-18s %7.1f us/step (inside consumer)' % ('loss enqueue', t_loss / steps * 1e6))
print('  %-18s %7.1f us/step (inside consumer)' % ('loss.item() wait', t_item / steps * 1e6))
