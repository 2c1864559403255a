import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, torch
from conftest import OAMIX_CFG, sampler_cfg
from oracle import oamix_np, synth
from oadg_b200 import OAMix
for case in [('augmix', 96, 160, 3, 0, 100, {}), ('augmix', 96, 160, 3, 1, 101, {}), ('augmix', 600, 1067, 8, 4, 44, {})][:int(os.environ.get('NCASE', '3'))]:
    version, h, w, n_gt, s, seed, extra = case
    cfg = dict(OAMIX_CFG, version=version, **extra)
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed)
    ref, plan = oamix_np.oamix_view(img, gt, **sampler_cfg(cfg))
    kinds = [op['name'] for br in plan['branches'] for regs in br for op in regs]
    for dbg in (sys.argv[1:] or ('0', '1', '2', '3')):
        os.environ['OADG_DEBUG'] = dbg
        np.random.seed(seed)
        t = OAMix(**cfg)
        try:
            out = t.oamix_batch([torch.from_numpy(img).cuda()], [gt])[0][0].cpu().numpy()
        except Exception as e:
            print(case[:6], 'debug', dbg, 'EXC', str(e)[:60]); break
        d = np.abs(out.astype(int) - ref.astype(int))
        ys, xs = np.nonzero(d.max(axis=2))
        print(case[:6], 'debug', dbg, 'max', d.max(), 'frac', (d != 0).mean(), 'bbox', (xs.min(), ys.min(), xs.max(), ys.max()) if len(xs) else None)
    if os.environ.get('MAP'):
        m = (d.max(axis=2) != 0)
        for by in range(0, h, 8):
            print(''.join('#' if m[by:by + 8, bx:bx + 8].mean() > 0.5 else ('+' if m[by:by + 8, bx:bx + 8].any() else '.') for bx in range(0, w, 8)))
    print('  ops', kinds, 'ml', plan['ml_boxes'].tolist())
