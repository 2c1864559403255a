"""CTA-busy time per work-item kind of the OA-Mix chain kernel for a few bench batches (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

dev = torch.device('cuda:0')
frames = [bench.make_image(s) for s in range(8)]
imgs = [torch.from_numpy(f).to(dev) for f, _ in frames]
gts = [g for _, g in frames]
mix = OAMix(**bench.OAMIX_CFG)
np.random.seed(1000)
for i in range(3):
    mix.oamix_batch(imgs[0:2], gts[0:2])
tot_chain = tot_mix = 0.0
quiet = len(sys.argv) > 2
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    prof = {}
    j = (2 * i) % 8
    mix.oamix_batch(imgs[j:j + 2], gts[j:j + 2], profile=prof)
    tot_chain += prof['chain_ms'] * 1e3
    tot_mix += prof['mix_ms'] * 1e3
    if quiet:
        continue
    print('batch %d: chain %.1f us, mix %.1f us, %d items, %d tiles' % (
        i, prof['chain_ms'] * 1e3, prof['mix_ms'] * 1e3, prof['items'], prof['tiles']))
    tot = 0.0
    for k, (us, n, mx) in prof['kind_busy_us_and_tiles'].items():
        tot += us
        print('   kind %-16s busy %9.1f CTA-us over %6d tiles = %7.2f us/tile, longest %7.1f us' % (k, us, n, us / max(n, 1), mx))
    print('   total %.1f CTA-us = %.1f us on 592 CTAs' % (tot, tot / 592))
print('chain kernel total over the batches: %.1f us, mix kernel total: %.1f us' % (tot_chain, tot_mix))
