"""Device time of the OA-Loss C-ABI sequence (CUDA events, warm): local loss at n = 2088 and one rank's share of a
W-rank gathered contrast set.  usage: python scripts/loss_stats.py [iters] [W ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth   # noqa: E402  (input generator only)
from oadg_b200 import ContrastiveLossPlus, reference_pair_map   # noqa: E402
from oadg_b200.distributed import CudaBackend, gathered_pair_map   # noqa: E402


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    worlds = [int(a) for a in sys.argv[2:]] or [1, 8]
    dev = torch.device('cuda:0')
    n = 2088
    x, labels = synth.make_roi_set(n)
    xd, ld = x.to(dev).requires_grad_(True), labels.to(dev)
    fn = ContrastiveLossPlus(loss_weight=0.01, num_views=2, temperature=0.06)

    def fwd_bwd():
        xd.grad = None
        fn(xd, ld).backward()
    with torch.no_grad():
        print('local n=%d: forward %.1f us' % (n, timed(lambda: fn(xd, ld), iters)))
    print('local n=%d: forward+backward %.1f us' % (n, timed(fwd_bwd, iters)))
    for w in worlds:
        if w == 1:
            continue
        be = CudaBackend()
        lab = labels.view(-1)
        lab = torch.cat([lab, lab[-1:].repeat(n - lab.shape[0])])
        xs = [synth.make_roi_set(n, seed=60 + r)[0].to(dev) for r in range(w)]
        f_all = torch.cat([be.normalize(v, w * n, True) for v in xs]).contiguous()
        labels_all = lab.repeat(w).to(dev)
        pair_all = torch.from_numpy(gathered_pair_map(reference_pair_map(n), w)).to(dev)
        be.normalize(xs[0], w * n, True)
        st = {}

        def f():
            st['s'] = be.forward(f_all, labels_all, pair_all, 0, n, 0.06, 0.01, 10)[1]
        tf = timed(f, iters)
        stats_all = st['s'].repeat(w, 1).contiguous()
        one = torch.ones((), device=dev)
        tb = timed(lambda: be.backward(xs[0], f_all, labels_all, pair_all, stats_all, 0, 0.06, True, one), iters)
        print('gathered W=%d (n_total=%d): forward %.1f us, backward %.1f us' % (w, w * n, tf, tb))


if __name__ == '__main__':
    main()
