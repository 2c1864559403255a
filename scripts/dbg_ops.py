"""GPU executor vs the host arithmetic check (tests/hostsim) on hand-made single-step plans (debug aid)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT + '/tests')
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import synth  # noqa: E402
from oadg_b200.oamix import OAMix, _ViewPlan, _invert_affine  # noqa: E402

hs = ctypes.CDLL(ROOT + '/tests/hostsim/libhostsim.so')
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 160)
img, gt = synth.make_image(0, h, w, 3)
t = OAMix(version='augmix.all')


def run(ops_by_branch, ml, label):
    vp = _ViewPlan()
    vp.h, vp.w = h, w
    nb = len(ops_by_branch)
    vp.ws = np.float32([1.0 / nb] * nb)
    vp.ml_boxes = np.array(ml, dtype=np.int64)
    vp.depths = [len(s) for s in ops_by_branch]
    vp.ops = ops_by_branch
    vp.scores, vp.oa_low, vp.oa_boxes, vp.m, vp.m_oa = [50.0] * len(gt), [], [], 1.0, []
    jobs = [(vp, gt, 0)]
    blob = t._pack(jobs)
    ref = np.zeros_like(img)
    src = (ctypes.c_void_p * 1)(img.ctypes.data)
    dst = (ctypes.c_void_p * 1)(ref.ctypes.data)
    n = ctypes.c_int(0)
    rc = hs.hostsim_oamix_execute(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, 1, dst, ctypes.byref(n))
    assert rc == 0, rc
    out = t.execute(blob, [torch.from_numpy(img).cuda()])[0].cpu().numpy()
    d = np.abs(out.astype(int) - ref.astype(int)).max(axis=2)
    ys, xs = np.nonzero(d)
    print('%-40s max %3d frac %.4f bbox %s' % (label, d.max(), (d != 0).mean(),
                                                (xs.min(), ys.min(), xs.max(), ys.max()) if len(xs) else None))
    if d.max() and os.environ.get('MAP'):
        for by in range(0, h, 8):
            print('   ' + ''.join('#' if (d[by:by + 8, bx:bx + 8] != 0).mean() > 0.5 else
                                  ('+' if d[by:by + 8, bx:bx + 8].any() else '.') for bx in range(0, w, 8)))


ml = [[14, 34, 29, 56]]
gi = [(int(b[0]), int(b[1]), int(b[2]), int(b[3])) for b in gt]
tr = ('bg_affine', _invert_affine([1.0, 0.0, -7.0, 0.0, 1.0, 0.0]))
sh = ('bg_affine', _invert_affine([1.0, 0.21, 0.0, 0.0, 1.0, 0.0]))
rot = ('bg_affine', _invert_affine(OAMix._forward_affine('rotate', 7.0, False, (w, h), None, (w, h))))
bbo = ('bbo_affine', [(k, _invert_affine(OAMix._forward_affine('rotate', 6.0, k % 2 == 0, (x2 - x1 + 1, y2 - y1 + 1),
                                                             ((x1 + x2) / 2., (y1 + y2) / 2.), (w, h))))
                      for k, (x1, y1, x2, y2) in enumerate(gi)])
for name, a, b in [('autocontrast | solarize', ('autocontrast',), ('solarize', 100)),
                   ('autocontrast | bg translate', ('autocontrast',), tr),
                   ('bg shear | posterize', sh, ('posterize', 3)),
                   ('bg rotate | bg rotate', rot, rot),
                   ('bbo rotate | autocontrast', bbo, ('autocontrast',)),
                   ('equalize | bbo rotate', ('equalize',), bbo),
                   ('invert | color', ('invert', 1, -1), ('color', 0.7)),
                   ('sharpness | contrast', ('sharpness', 1.5), ('contrast', 0.4))]:
    run([[[a, b]]], ml, name)
run([[[('autocontrast',), tr]], [[bbo, ('autocontrast',)], [('posterize', 3), ('solarize', 99)]], [[sh, bbo], [bbo, ('solarize', 77)]]],
    ml, 'case0-like 3 branches')
