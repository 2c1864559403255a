"""CTA-busy time per work-item kind of the OA-Mix chain kernel for a few bench batches (GPU box).

    python scripts/chain_stats.py [n_launches] [views_per_launch] [quiet]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

dev = torch.device('cuda:0')
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
group = int(sys.argv[2]) if len(sys.argv) > 2 else 2
quiet = len(sys.argv) > 3
pool = 24
frames = [bench.make_image(s) for s in range(pool)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
np.random.seed(1000)


def pick(i):
    idx = [(group * i + k) % pool for k in range(group)]
    return [imgs[j] for j in idx], [gts[j] for j in idx]


for i in range(3):
    mix.oamix_batch(*pick(i))
tot_chain = tot_mix = 0.0
step_bytes = view_bytes = 0
for i in range(n_launch):
    prof = {}
    mix.oamix_batch(*pick(i), profile=prof)
    tot_chain += prof['chain_ms'] * 1e3
    tot_mix += prof['mix_ms'] * 1e3
    step_bytes += prof['step_bytes']
    view_bytes += prof['view_bytes']
    if quiet:
        continue
    print('launch %d: chain %.1f us, mix %.1f us, %d items, %d tiles' % (
        i, prof['chain_ms'] * 1e3, prof['mix_ms'] * 1e3, prof['items'], prof['tiles']))
    tot = 0.0
    for k, (us, n, mx) in prof['kind_busy_us_and_tiles'].items():
        tot += us
        print('   kind %-16s busy %9.1f CTA-us over %6d tiles = %7.2f us/tile, longest %7.1f us' % (k, us, n, us / max(n, 1), mx))
    print('   total %.1f CTA-us = %.1f us on 444 CTAs (3 per SM)' % (tot, tot / 444))
print('%d launches x %d views: chain %.1f us/view, mix %.1f us/view; chain %.0f GB/s of lane-step bytes, '
      'chain+mix %.0f GB/s of whole-view bytes (profiled build: timers on)' % (
          n_launch, group, tot_chain / (n_launch * group), tot_mix / (n_launch * group),
          step_bytes / tot_chain / 1e3, view_bytes / (tot_chain + tot_mix) / 1e3))
# the production kernel (no timers): CUDA events around un-profiled calls
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
np.random.seed(1000)
for i in range(3):
    mix.oamix_batch(*pick(i))
np.random.seed(1000)
torch.cuda.synchronize()
t_tot = 0.0
for i in range(n_launch):
    im, g = pick(i)
    scores = mix.saliency_scores(im, g)
    plan = mix.sample_plan([(int(t.shape[0]), int(t.shape[1])) for t in im], g, scores)
    torch.cuda.synchronize()
    e0.record()
    mix.execute(plan.blob, im)
    e1.record()
    torch.cuda.synchronize()
    t_tot += e0.elapsed_time(e1) * 1e3
print('production kernels (upload + chain + mix, CUDA events): %.1f us/view' % (t_tot / (n_launch * group)))
