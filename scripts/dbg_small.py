"""Debug aid (GPU box): one small OA-Mix view through the plugin vs the oracle; prints the mismatch statistics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from conftest import OAMIX_CFG, sampler_cfg  # noqa: E402
from oracle import oamix_np, synth  # noqa: E402
from oadg_b200 import OAMix  # noqa: E402

h, w, n_gt = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (96, 160, 3)))
seeds = range(int(sys.argv[4]) if len(sys.argv) > 4 else 4)
dev = torch.device('cuda:0')
cfg = dict(OAMIX_CFG, version='augmix')
bad = 0
for s in seeds:
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(100 + s)
    ref, plan = oamix_np.oamix_view(img, gt, **sampler_cfg(cfg))
    np.random.seed(100 + s)
    t = OAMix(**cfg)
    outs, _, _ = t.oamix_batch([torch.from_numpy(img).to(dev)], [gt])
    torch.cuda.synchronize()
    out = outs[0].cpu().numpy()
    d = np.abs(out.astype(int) - ref.astype(int))
    print('seed %d: max diff %d, mismatching %.5f %%' % (s, d.max(), 100.0 * (d != 0).mean()), flush=True)
    bad += int(d.max() > 1 or (d != 0).mean() > 1e-3)
print('BAD' if bad else 'OK')
