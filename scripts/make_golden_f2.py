"""Generate tests/golden/consistency.npz by running the UNMODIFIED reference losses (dev container only).

    PYTHONPATH=. python scripts/make_golden_f2.py

CrossEntropyLossPlus (+ jsdv1_3_2aug) for the RoI head (softmax, 9 classes) and the RPN head (sigmoid, 1 logit),
SmoothL1LossPlus / L1LossPlus, and supcontrast_yolo, on seeded inputs, in float64 (loss + gradient) and float32 (loss);
imported from /root/reference through oracle.ref_loader."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader, synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'consistency.npz')


def inputs(kind):
    g = torch.Generator().manual_seed({'roi': 11, 'rpn': 12, 'reg': 13}[kind])
    if kind == 'roi':      # 2 views x 2 images x 64 RoIs, 8 classes + background
        n = 256
        pred = torch.randn(n, 9, generator=g, dtype=torch.float64) * 3
        half = torch.randint(0, 9, (n // 2,), generator=g)
        return pred, torch.cat([half, half]), torch.ones(n, dtype=torch.float64), float(n)
    if kind == 'rpn':      # objectness logits of both views' anchors, label 0 = fg, 1 = bg (mmdet RPN convention)
        n = 4096
        pred = torch.randn(n, 1, generator=g, dtype=torch.float64) * 4
        half = torch.randint(0, 2, (n // 2,), generator=g)
        w = (torch.rand(n // 2, generator=g) < 0.25).double()
        return pred, torch.cat([half, half]), torch.cat([w, w]), 512.0
    n = 512                # box regression deltas of the positives of both views
    pred = torch.randn(n, 4, generator=g, dtype=torch.float64)
    tgt = torch.randn(n // 2, 4, generator=g, dtype=torch.float64)
    w = (torch.rand(n // 2, 4, generator=g) < 0.8).double()
    return pred, torch.cat([tgt, tgt]), torch.cat([w, w]), float(n)


def main():
    R = ref_loader.load_reference()
    out = {}
    cfgs = {
        'roi': dict(use_sigmoid=False, loss_weight=1.0, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=10,
                    wandb_name='roi_cls', log_pos_ratio=True),
        'rpn': dict(use_sigmoid=True, loss_weight=1.0, num_views=2, additional_loss='jsdv1_3_2aug', lambda_weight=0.1,
                    wandb_name='rpn_cls'),
    }
    for kind, cfg in cfgs.items():
        pred, label, weight, avg = inputs(kind)
        for dt in (torch.float64, torch.float32):
            p = pred.detach().clone().to(dt).requires_grad_(True)
            loss = R['CrossEntropyLossPlus'](**cfg)(p, label, weight.to(dt), avg_factor=avg)
            loss.backward()
            tag = '%s/%s' % (kind, 'f64' if dt == torch.float64 else 'f32')
            out[tag + '/loss'] = np.float64(loss.item())
            out[tag + '/grad'] = p.grad.numpy().copy()
        pj = pred.detach().clone().requires_grad_(True)
        j = R['jsdv1_3_2aug'](pj, label, None, reduction='mean', avg_factor=None)
        j.backward()
        out[kind + '/jsd'] = np.float64(j.item())
        out[kind + '/jsd_grad'] = pj.grad.numpy().copy()
    pred, tgt, w, avg = inputs('reg')
    for name, cls, kw in (('smoothl1', 'SmoothL1LossPlus', dict(beta=1.0)), ('l1', 'L1LossPlus', {})):
        p = pred.detach().clone().requires_grad_(True)
        loss = R[cls](loss_weight=1.0, num_views=2, additional_loss='None', lambda_weight=0.0, wandb_name='x', **kw)(
            p, tgt, w, avg_factor=avg)
        loss.backward()
        out[name + '/loss'] = np.float64(loss.item())
        out[name + '/grad'] = p.grad.numpy().copy()
    # supcontrast_yolo on 2 x 900 sampled cells (+ an odd trailing row)
    for n in (1800, 1801):
        x, _ = synth.make_roi_set(max(n, 2048), seed=n)
        x = x[:n].double()
        g = torch.Generator().manual_seed(n)
        half = torch.full((n // 2,), 8, dtype=torch.int64)
        idx = torch.randperm(n // 2, generator=g)[:150]
        half[idx] = torch.randint(0, 8, (150,), generator=g)
        labels = torch.cat([half, half, torch.full((n - 2 * (n // 2),), 8, dtype=torch.int64)]).view(-1, 1)
        xr = torch.nn.functional.normalize(x, dim=1).requires_grad_(True)
        loss = R['supcontrast_yolo'](xr, labels, temper=0.06, min_samples=10)
        loss.backward()
        out['yolo%d/loss' % n] = np.float64(loss.item())
        out['yolo%d/grad_rows' % n] = np.concatenate([xr.grad.numpy()[:8], xr.grad.numpy()[n // 2:n // 2 + 8]])
        out['yolo%d/labels' % n] = labels.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, sorted(out))


if __name__ == '__main__':
    main()
