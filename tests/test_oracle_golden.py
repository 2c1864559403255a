"""The oracle against the committed goldens (made by the reference itself, scripts/make_golden.py)
and -- in the dev container where /root/reference exists -- against the reference run live."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN, OAMIX_CFG, sampler_cfg
from oracle import oamix_np, supcon_np, synth, ref_loader

SMALL = np.load(GOLDEN + '/oamix_small.npz')
FULL = np.load(GOLDEN + '/oamix_full.npz')
SUPCON = np.load(GOLDEN + '/supcon.npz')
SMALL_CASES = {'augmix_a': ('augmix', {}), 'augmix_b': ('augmix', {}), 'augmix_c': ('augmix', {}),
               'augmix_d': ('augmix', {}), 'all_a': ('augmix.all', {}), 'all_b': ('augmix.all', {}),
               'all_c': ('augmix.all', {}), 'all_nogt': ('augmix.all', dict(mixture_width=1)),
               'dwd_w1': ('augmix.all', dict(mixture_width=1))}


@pytest.mark.parametrize('name', sorted(SMALL_CASES))
def test_oamix_oracle_matches_reference_golden_small(name):
    version, extra = SMALL_CASES[name]
    h, w, n_gt, s, seed = (int(v) for v in SMALL[name + '/meta'])
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed)
    res = oamix_np.oamix_call(dict(img=img.copy(), gt_bboxes=gt.copy()), num_views=2, keep_orig=True,
                              **sampler_cfg(dict(OAMIX_CFG, version=version, **extra)))
    assert np.array_equal(res['img2'], SMALL[name + '/img2'])          # bit-exact
    assert np.array_equal(res['oamix_boxes'], SMALL[name + '/oamix_boxes'])
    assert np.array_equal(res['multilevel_boxes'], SMALL[name + '/multilevel_boxes'])
    assert res['oamix_boxes'].dtype == np.int64 and res['multilevel_boxes'].dtype == np.int64
    assert res['custom_field'] == ['img2', 'gt_bboxes2', 'oamix_boxes', 'multilevel_boxes']
    assert res['img_fields'] == ['img', 'img2']


@pytest.mark.parametrize('s', [1])
def test_oamix_oracle_matches_reference_golden_full(s):
    img, gt = synth.make_image(s)
    np.random.seed(1000 + s)
    out, plan = oamix_np.oamix_view(img, gt, **sampler_cfg(dict(OAMIX_CFG, version='augmix')))
    assert int(out.astype(np.int64).sum()) == int(FULL['s%d/sum' % s])
    assert hashlib.sha256(out.tobytes()).digest() == FULL['s%d/sha256' % s].tobytes()
    assert np.array_equal(np.stack(plan['oa_boxes']), FULL['s%d/oamix_boxes' % s])
    assert np.allclose(plan['scores'], FULL['s%d/scores' % s], rtol=0, atol=0)


@pytest.mark.parametrize('n', [2048, 2088, 2085])
def test_supcon_oracle_matches_reference_golden(n):
    x, labels = synth.make_roi_set(n)
    loss, grad = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
    ref64 = float(SUPCON['n%d/f64/loss' % n])
    assert abs(loss - ref64) <= 1e-12 * abs(ref64)
    assert abs(loss - float(SUPCON['n%d/f32/loss' % n])) <= 1e-6 * abs(ref64)   # the f32 reference itself
    rows = np.concatenate([grad[0:8], grad[1024:1032], grad[n - 8:n]])
    g64 = SUPCON['n%d/f64/grad_rows' % n]
    assert np.linalg.norm(rows - g64) <= 1e-10 * np.linalg.norm(g64)
    assert abs(np.linalg.norm(grad) - float(SUPCON['n%d/f64/grad_norm' % n])) <= 1e-10 * np.linalg.norm(grad)


@pytest.mark.parametrize('n', [2048, 2088])
def test_supcon_torch_restatement_matches_reference_golden(n):
    """oracle/supcon_torch.py (the stock-torch op sequence bench.py times on the GPU as the reference's own path) in
    f64 on the CPU: loss and gradient rows against the fixtures the reference itself produced."""
    import torch
    from oracle import supcon_torch
    x, labels = synth.make_roi_set(n)
    xr = x.double().requires_grad_(True)
    loss = supcon_torch.contrastive_loss_plus_torch(xr, labels, loss_weight=0.01, temperature=0.06)
    loss.backward()
    ref64 = float(SUPCON['n%d/f64/loss' % n])
    assert abs(float(loss) - ref64) <= 1e-12 * abs(ref64)
    g = xr.grad.numpy()
    rows = np.concatenate([g[0:8], g[1024:1032], g[n - 8:n]])
    g64 = SUPCON['n%d/f64/grad_rows' % n]
    assert np.linalg.norm(rows - g64) <= 1e-10 * np.linalg.norm(g64)
    few, lab = synth.make_roi_set(2048, n_fg=5)
    assert float(supcon_torch.contrastive_loss_plus_torch(few, lab, 0.01, 0.06)) == 0.0


def test_supcon_oracle_quirks():
    x, labels = synth.make_roi_set(2048, n_fg=5)
    assert supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01) == float(SUPCON['fewfg/loss']) == 0.0
    with pytest.raises(RuntimeError):   # the reference crashes below 2048 rows (contrastive_loss.py:205)
        supcon_np.supcon_loss(np.zeros((1024, 256)), np.zeros(1024, np.int64))
    assert supcon_np.supcon_loss(np.zeros((0, 256)), np.zeros(0, np.int64)) == 0.0


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree only exists in the dev container')
def test_oracle_matches_live_reference():
    """Run in a subprocess: ref_loader installs stub mmcv/mmdet modules and must not meet the shim."""
    import os
    import subprocess
    import sys
    code = r'''
import numpy as np, torch
from oracle import ref_loader, synth, oamix_np, supcon_np
R = ref_loader.load_reference()
cfg = dict(num_views=2, keep_orig=True, severity=10, random_box_ratio=(3,1/3), random_box_scale=(0.01,0.1),
           oa_random_box_scale=(0.005,0.1), oa_random_box_ratio=(3,1/3), spatial_ratio=4, sigma_ratio=0.3)
for version, h, w, n_gt, s, seed, extra in [('augmix', 96, 160, 3, 5, 905, {}), ('augmix.all', 120, 200, 4, 8, 908, {}),
                                            ('augmix.all', 90, 90, 2, 9, 909, dict(mixture_width=2, mixture_depth=2)),
                                            ('augmix', 80, 120, 0, 2, 902, {}),
                                            ('augmix', 96, 160, 3, 6, 906, dict(num_views=3, keep_orig=False))]:
    c = dict(cfg, version=version, **extra)
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed); a = R['OAMix'](**c)(dict(img=img.copy(), gt_bboxes=gt.copy()))
    st_a = np.random.get_state()
    np.random.seed(seed); b = oamix_np.oamix_call(dict(img=img.copy(), gt_bboxes=gt.copy()), **{k: v for k, v in c.items() if k != 'severity'})
    st_b = np.random.get_state()
    assert sorted(a) == sorted(b), (sorted(a), sorted(b))
    for k in a:
        if isinstance(a[k], np.ndarray): assert np.array_equal(a[k], b[k]), k
        else: assert a[k] == b[k], k
    assert st_a[2] == st_b[2] and np.array_equal(st_a[1], st_b[1])
for n in (2048, 2088, 2085):
    x, labels = synth.make_roi_set(n, seed=n)
    xr = x.double().requires_grad_(True)
    l = R['ContrastiveLossPlus'](loss_weight=0.01, num_views=2, temperature=0.06)(xr, labels); l.backward()
    lo, go = supcon_np.supcon_loss(x.numpy(), labels.numpy(), 0.06, 10, 0.01, want_grad=True)
    assert abs(lo - l.item()) <= 1e-12 * abs(lo)
    assert np.linalg.norm(go - xr.grad.numpy()) <= 1e-10 * np.linalg.norm(go)
print("LIVE-OK")
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True,
                         env=dict(os.environ, PYTHONPATH=root))
    assert 'LIVE-OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
