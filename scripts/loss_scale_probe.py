"""Per-rank cost of the gathered OA-Loss as the world grows, on ONE GPU: the packed rows of W ranks are fabricated
(W independent random RoI sets), rank 0's four library calls are timed with CUDA events.  No collective is timed."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oadg_b200 import distributed as D  # noqa: E402
from oadg_b200.contrastive_loss import reference_pair_map  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device('cuda:0')
n = 2088
pair_local = reference_pair_map(n)
worlds = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 2, 4, 8]
n_warm, n_iter = (1, 2) if len(sys.argv) > 2 else (5, 20)   # a second argument: short run for ncu
for world in worlds:
    be = D.CudaBackend()
    sets = [synth.make_roi_set(n, seed=100 + r) for r in range(world)]
    xs = [s[0].to(dev) for s in sets]
    ys = [s[1].to(dev).view(-1) for s in sets]
    pair_all = D._pair_all_on(dev, pair_local, world)
    g = torch.ones((), device=dev) * world

    def step(timed=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        send = be.pack(xs[0], ys[0], world * n, True)
        ev[1].record()
        recv = torch.cat([send] + [be.pack(xs[r], ys[r], world * n, True) for r in range(1, world)]) if world > 1 else send
        torch.cuda.synchronize()
        ev[1].record()
        tail = be.forward_packed(recv, pair_all, 0, n, 0.06, 0.01, 10)
        ev[2].record()
        tail_all = torch.cat([tail] * world) if world > 1 else tail
        torch.cuda.synchronize()
        ev[2].record() if False else None
        e3 = torch.cuda.Event(enable_timing=True)
        e3.record()
        loss = be.finish(tail_all, world, n)
        ev[3].record()
        gx = be.backward_packed(xs[0], pair_all, 0, 0.06, True, g)
        ev[4].record()
        torch.cuda.synchronize()
        if timed is not None:
            timed.append((ev[1].elapsed_time(ev[2]), e3.elapsed_time(ev[3]), ev[3].elapsed_time(ev[4])))
        return loss, gx

    for _ in range(n_warm):
        step()
    t = []
    for _ in range(n_iter):
        step(t)
    t = np.median(np.array(t), axis=0) * 1e3
    print('W=%d  rows %5d  forward_packed %7.1f us  finish %5.1f us  backward_packed %7.1f us  sum %7.1f us' %
          (world, world * n, t[0], t[1], t[2], t.sum()))
