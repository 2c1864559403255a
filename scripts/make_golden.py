"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

    PYTHONPATH=. python scripts/make_golden.py

The reference modules are imported from /root/reference through oracle.ref_loader (mmcv stub +
cv2.saliency restatement, see its header).  Inputs come from oracle.synth (seeded); the plan RNG
is np.random.seed(seed) immediately before the call.  Outputs are what the reference returned.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
CFG = dict(num_views=2, keep_orig=True, severity=10, random_box_ratio=(3, 1 / 3), random_box_scale=(0.01, 0.1),
           oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3), spatial_ratio=4, sigma_ratio=0.3)

SMALL = [  # (name, version, h, w, n_gt, image seed, rng seed, extra cfg)
    ('augmix_a', 'augmix', 96, 160, 3, 0, 100, {}),
    ('augmix_b', 'augmix', 96, 160, 3, 3, 103, {}),
    ('augmix_c', 'augmix', 101, 203, 5, 1, 401, {}),
    ('augmix_d', 'augmix', 128, 256, 8, 2, 502, {}),
    ('all_a', 'augmix.all', 120, 200, 4, 1, 201, {}),
    ('all_b', 'augmix.all', 120, 200, 4, 6, 206, {}),
    ('all_c', 'augmix.all', 120, 200, 4, 7, 207, {}),
    ('all_nogt', 'augmix.all', 64, 64, 0, 1, 301, dict(mixture_width=1)),
    ('dwd_w1', 'augmix.all', 150, 267, 6, 4, 604, dict(mixture_width=1, mixture_depth=-1)),
]
FULL = [(s, 1000 + s) for s in range(4)]  # BASELINE config 1: 1024x2048, 8 gt


def main():
    R = ref_loader.load_reference()
    os.makedirs(OUT, exist_ok=True)
    small = {}
    for name, version, h, w, n_gt, s, seed, extra in SMALL:
        cfg = dict(CFG, version=version, **extra)
        t = R['OAMix'](**cfg)
        img, gt = synth.make_image(s, h, w, n_gt)
        np.random.seed(seed)
        res = t(dict(img=img.copy(), gt_bboxes=gt.copy()))
        small[name + '/img2'] = res['img2']
        small[name + '/oamix_boxes'] = res['oamix_boxes']
        small[name + '/multilevel_boxes'] = res['multilevel_boxes']
        small[name + '/scores'] = np.array(t._history['fg_score_list'], dtype=np.float64)
        small[name + '/meta'] = np.array([h, w, n_gt, s, seed], dtype=np.int64)
        print(name, res['img2'].shape, int(res['img2'].astype(np.int64).sum()))
    np.savez_compressed(os.path.join(OUT, 'oamix_small.npz'), **small)

    full = {}
    for s, seed in FULL:
        t = R['OAMix'](**dict(CFG, version='augmix'))
        img, gt = synth.make_image(s)
        np.random.seed(seed)
        res = t(dict(img=img.copy(), gt_bboxes=gt.copy()))
        out = res['img2']
        full['s%d/sum' % s] = np.int64(out.astype(np.int64).sum())
        full['s%d/sha256' % s] = np.frombuffer(hashlib.sha256(out.tobytes()).digest(), np.uint8)
        full['s%d/thumb' % s] = out[::16, ::16].copy()  # 64x128x3 sub-sample for <=1 LSB comparisons
        full['s%d/oamix_boxes' % s] = res['oamix_boxes']
        full['s%d/multilevel_boxes' % s] = res['multilevel_boxes']
        full['s%d/scores' % s] = np.array(t._history['fg_score_list'], dtype=np.float64)
        print('full', s, int(full['s%d/sum' % s]))
    np.savez_compressed(os.path.join(OUT, 'oamix_full.npz'), **full)

    loss = {}
    for n in (2048, 2088, 2085):
        x, labels = synth.make_roi_set(n)
        L = R['ContrastiveLossPlus'](loss_weight=0.01, num_views=2, temperature=0.06)
        for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
            xr = x.to(dt).clone().requires_grad_(True)
            l = L(xr, labels)
            l.backward()
            g = xr.grad.numpy()
            loss['n%d/%s/loss' % (n, tag)] = np.array(l.item(), dtype=np.float64)
            loss['n%d/%s/grad_norm' % (n, tag)] = np.array(np.linalg.norm(g.astype(np.float64)))
            loss['n%d/%s/grad_rows' % (n, tag)] = np.concatenate([g[0:8], g[1024:1032], g[n - 8:n]]).astype(np.float64)
        print('loss', n, float(loss['n%d/f32/loss' % n]), float(loss['n%d/f64/loss' % n]))
    x, labels = synth.make_roi_set(2048, n_fg=5)
    L = R['ContrastiveLossPlus'](loss_weight=0.01, num_views=2, temperature=0.06)
    loss['fewfg/loss'] = np.array(float(L(x, labels)))
    np.savez_compressed(os.path.join(OUT, 'supcon.npz'), **loss)


if __name__ == '__main__':
    main()
