"""Kernel arithmetic + host orchestration on the CPU (tests/hostsim: the device bodies compiled for the
host, test infrastructure only) against the oracle.  Lets the dev container catch parity bugs without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import OAMIX_CFG, ROOT, sampler_cfg
from oracle import oamix_np, synth


@pytest.fixture(scope='module')
def hostsim():
    from conftest import build_hostsim
    return ctypes.CDLL(build_hostsim())


def run(hs, t, jobs, imgs):
    blob = t._pack(jobs)
    outs = [np.zeros_like(imgs[j[2]]) for j in jobs]
    src = (ctypes.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
    dst = (ctypes.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    n = ctypes.c_int(0)
    rc = hs.hostsim_oamix_execute(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, len(imgs),
                                  dst, ctypes.byref(n))
    assert rc == 0, rc
    return outs


CASES = [('augmix', 96, 160, 3, s, 100 + s, {}) for s in range(4)] + \
        [('augmix.all', 120, 200, 4, s, 200 + s, {}) for s in range(5)] + \
        [('augmix.all', 64, 64, 0, 1, 301, dict(mixture_width=1)), ('augmix', 101, 203, 5, 1, 401, {}),
         ('augmix', 99, 131, 6, 3, 77, dict(mixture_width=4, mixture_depth=3))]


@pytest.mark.parametrize('case', CASES)
def test_kernel_arithmetic_matches_oracle(hostsim, case):
    from oadg_b200.oamix import OAMix
    version, h, w, n_gt, s, seed, extra = case
    cfg = sampler_cfg(dict(OAMIX_CFG, version=version, **extra))
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed)
    ref, plan = oamix_np.oamix_view(img, gt, **cfg)
    np.random.seed(seed)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    out, = run(hostsim, t, [(vp, gt, 0)], [img])
    d = np.abs(out.astype(int) - ref.astype(int))
    # float stages (blurred masks are ~1e-7 off cv2's) may flip a truncation by 1 LSB at a handful of pixels
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3, (int(d.max()), float((d != 0).mean()))


@pytest.mark.parametrize('rep', range(8))
def test_frame_sized_box_saturated_mask_is_exact(hostsim, rep):
    """A gt box that spans the frame: every blur window is fully covered through the reflected border, the mask
    saturates at cv2's float32 constant (1 - 2^-24, 1 or 1 + 2^-23 depending on the kernel) and `img*(1-m) + aug*m`
    sits on an integer at EVERY pixel, so the last bit of that constant decides every truncation.  With the constant
    summed in OpenCV's order (cv_saturated_row / _col, oamix_math.h) the views are exact; with 1.0 in its place up to
    43 % of the values were off by 1-3 LSB."""
    from oadg_b200.oamix import OAMix
    h, w, gt = 128, 96, np.float32([[0, 0, 96, 128]])
    img, _ = synth.make_image(3, h, w, 0)
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix.all'))
    np.random.seed(3 + 100 * rep)
    ref, plan = oamix_np.oamix_view(img, gt, **cfg)
    np.random.seed(3 + 100 * rep)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    out, = run(hostsim, t, [(vp, gt, 0)], [img])
    assert np.array_equal(out, ref)


def test_two_views_one_batch(hostsim):
    from oadg_b200.oamix import OAMix
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix'))
    imgs, gts, refs, jobs = [], [], [], []
    t = OAMix(**cfg)
    np.random.seed(5)
    st = np.random.get_state()
    for i, (h, w) in enumerate([(96, 160), (80, 120)]):   # ragged batch
        img, gt = synth.make_image(10 + i, h, w, 3)
        imgs.append(img)
        gts.append(gt)
    for img, gt in zip(imgs, gts):
        ref, plan = oamix_np.oamix_view(img, gt, **cfg)
        refs.append((ref, plan))
    np.random.set_state(st)
    for i, (img, gt) in enumerate(zip(imgs, gts)):
        vp = t._sample_head(img.shape[0], img.shape[1], gt)
        t._sample_tail(vp, gt, refs[i][1]['scores'])
        jobs.append((vp, gt, i))
    outs = run(hostsim, t, jobs, imgs)
    for o, (ref, _) in zip(outs, refs):
        d = np.abs(o.astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 1e-3


def run_ex(hs, t, jobs, imgs, n_cta):
    blob = t._pack(jobs)
    outs = [np.zeros_like(imgs[j[2]]) for j in jobs]
    src = (ctypes.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
    dst = (ctypes.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
    stats = (ctypes.c_int * 3)()
    rc = hs.hostsim_oamix_execute_ex(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, len(imgs),
                                     dst, n_cta, stats)
    assert rc == 0, rc
    return outs, list(stats)


@pytest.mark.parametrize('seed', [11, 12, 13, 14])
def test_chain_scheduler_medium_frames(hostsim, seed):
    """8 gt boxes on a 320x576 frame: bboxes-only chains with several dependency levels, dead boxes pruned, phases
    executed in queue order (1) and in adversarial dependency-only orders (7, 148) -- the output must not depend on the order and must match the oracle."""
    from oadg_b200.oamix import OAMix
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix'))
    img, gt = synth.make_image(seed, 320, 576, 8)
    np.random.seed(seed)
    ref, plan = oamix_np.oamix_view(img, gt, **cfg)
    np.random.seed(seed)
    t = OAMix(**cfg)
    vp = t._sample_head(320, 576, gt)
    t._sample_tail(vp, gt, plan['scores'])
    outs = []
    for n_cta in (1, 7, 148):
        (o,), stats = run_ex(hostsim, t, [(vp, gt, 0)], [img], n_cta)
        outs.append(o)
        assert stats[1] >= 1
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    d = np.abs(outs[0].astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3, (int(d.max()), float((d != 0).mean()))
    n_boxes = sum(len(op[1]) for steps in vp.ops for regs in steps for op in regs if op[0] == 'bbo_affine')
    assert stats[2] <= n_boxes      # dead-box elimination never adds work


def test_many_boxes_deep_chains(hostsim):
    """40 gt boxes on a small frame: long bboxes-only chains (many dependency levels, hundreds of work items) and a
    crowded union mask -- scheduler limits and dependency tables under stress, against the oracle."""
    from oadg_b200.oamix import OAMix
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix'))
    img, gt = synth.make_image(5, 192, 256, 40)
    for seed in (3, 4):
        np.random.seed(seed)
        ref, plan = oamix_np.oamix_view(img, gt, **cfg)
        np.random.seed(seed)
        t = OAMix(**cfg)
        vp = t._sample_head(192, 256, gt)
        t._sample_tail(vp, gt, plan['scores'])
        (a,), _ = run_ex(hostsim, t, [(vp, gt, 0)], [img], 1)
        (b,), st = run_ex(hostsim, t, [(vp, gt, 0)], [img], 64)
        assert np.array_equal(a, b) and st[1] >= 10
        d = np.abs(a.astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d != 0).mean() <= 2e-3, (int(d.max()), float((d != 0).mean()))


def test_frames_beyond_the_compiled_limit_are_rejected(hostsim):
    """The profile tile keeps <= 1024 low-res samples / taps in shared memory: a 4100-px-wide frame must come back as
    OADG_E_LIMIT (-3) from the executor, not as a crash."""
    from oadg_b200.oamix import OAMix
    t = OAMix(**sampler_cfg(dict(OAMIX_CFG, version='augmix')))
    h, w = 96, 4104
    img = np.zeros((h, w, 3), np.uint8)
    gt = np.float32([[10, 2, 300, 60]])
    for seed in range(50):   # a seed whose random boxes fit the flat frame
        np.random.seed(seed)
        try:
            vp = t._sample_head(h, w, gt)
            t._sample_tail(vp, gt, [np.float64(20.0)])
            break
        except ValueError:
            continue
    blob = t._pack([(vp, gt, 0)])
    out = np.zeros_like(img)
    src = (ctypes.c_void_p * 1)(img.ctypes.data)
    dst = (ctypes.c_void_p * 1)(out.ctypes.data)
    rc = hostsim.hostsim_oamix_execute(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, 1, dst, None)
    assert rc == -3, rc


def _sweep_cases(n=16, seed=5):
    rng = np.random.RandomState(seed)
    cases = []
    for k in range(n):
        version = 'augmix.all' if k % 2 else 'augmix'
        h, w = int(rng.randint(40, 260)), int(rng.randint(40, 400))
        n_gt, s, draw = int(rng.randint(0, 9)), int(rng.randint(0, 1000)), int(rng.randint(0, 100000))
        extra = {} if k % 3 else dict(mixture_width=int(rng.randint(1, 5)), mixture_depth=int(rng.choice([-1, 1, 2, 3])))
        cases.append((version, h, w, n_gt, s, draw, extra))
    return cases


@pytest.mark.parametrize('case', _sweep_cases())
def test_randomized_sweep_against_oracle_in_two_execution_orders(hostsim, case):
    """Random frame sizes (odd widths, tiles cut by the frame edge), box counts 0..8, both op sets and random
    mixture widths / depths: oracle parity and independence of the work-queue order."""
    from oadg_b200.oamix import OAMix
    version, h, w, n_gt, s, seed, extra = case
    cfg = sampler_cfg(dict(OAMIX_CFG, version=version, **extra))
    img, gt = synth.make_image(s, h, w, n_gt)
    np.random.seed(seed)
    try:
        ref, plan = oamix_np.oamix_view(img, gt, **cfg)
    except ValueError:
        pytest.skip('no random box fits this frame (the reference raises as well)')
    np.random.seed(seed)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    outs = [run_ex(hostsim, t, [(vp, gt, 0)], [img], n_cta)[0][0] for n_cta in (1, 5)]
    assert np.array_equal(outs[0], outs[1])
    d = np.abs(outs[0].astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3, (int(d.max()), float((d != 0).mean()))


def _adversarial_boxes(n=48, seed=77):
    rng = np.random.RandomState(seed)
    cases = []
    for k in range(n):
        h, w = int(rng.randint(48, 280)), int(rng.randint(48, 400))
        n_gt, s, draw = int(rng.randint(1, 7)), int(rng.randint(0, 1000)), int(rng.randint(0, 100000))
        cases.append((k, h, w, n_gt, s, draw, int(rng.randint(n_gt))))
    return cases


@pytest.mark.parametrize('case', _adversarial_boxes())
def test_adversarial_gt_boxes_against_oracle(hostsim, case):
    """gt boxes the synthetic generator never draws: beyond the frame, narrower than spatial_ratio, sub-pixel, inverted,
    strips along a border, spanning an axis, all but a one-pixel frame.  Boxes that touch two borders have zones where
    the blurred mask is within ~1e-6 of 1 and a 1e-7 difference to cv2's float32 filter flips a truncation, and the
    flips of successive depth steps can add up: the bound here (3 LSB; more than 1 LSB on <= 3e-4 and any difference on
    <= 0.3 % of the values; measured worst: 3 LSB on 5e-5, 1 LSB on 2.6e-3) is looser than the 1 LSB / 0.1 % every
    other test holds, and is stated as a known deviation in DESIGN.md section 7."""
    from oadg_b200.oamix import OAMix
    k, h, w, n_gt, s, seed, j = case
    img, gt = synth.make_image(s, h, w, n_gt)
    gt = gt.copy()
    gt[j] = [[w * 0.5, h * 0.3, w + 30.7, h + 12.2], [10, 10, 13.9, 40], [30.2, 20.7, 30.9, 21.1],
             [w * 0.6, h * 0.6, w * 0.3, h * 0.2], [0, 0, w, 4], [w - 5, 0, w, h], [1, 1, w - 1, h - 1],
             [0, h * 0.25, w, h * 0.75]][k % 8]
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix.all' if k % 2 else 'augmix'))
    np.random.seed(seed)
    try:
        ref, plan = oamix_np.oamix_view(img, gt, **cfg)
    except ValueError:
        pytest.skip('no random box fits this frame (the reference raises as well)')
    np.random.seed(seed)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    out, = run(hostsim, t, [(vp, gt, 0)], [img])
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 3 and (d != 0).mean() <= 3e-3 and (d > 1).mean() <= 3e-4, (int(d.max()), float((d != 0).mean()))


class _FusedOut(ctypes.Structure):   # oadg_fused_out_t
    _fields_ = [('mean', ctypes.c_float * 3), ('std', ctypes.c_float * 3), ('to_rgb', ctypes.c_int32),
                ('size_divisor', ctypes.c_int32), ('view_f32', ctypes.c_void_p), ('src_f32', ctypes.c_void_p)]


NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375])


@pytest.mark.parametrize('hw,to_rgb,div', [((96, 160), True, 32), ((101, 203), True, 32), ((64, 64), False, 1)])
def test_fused_normalize_pad_chw_epilogue_is_exact(hostsim, hw, to_rgb, div):
    """The mix's opt-in epilogue (Normalize + Pad + HWC->CHW float32, oadg_fused_out_t) against the oracle's restatement
    of mmcv.imnormalize / impad_to_multiple / DefaultFormatBundle with the installed cv2: bit-exact floats, for the
    generated view (computed from the uint8 view the same call returns) and for the untouched source frame."""
    from oadg_b200.oamix import OAMix
    from oracle import prims_np
    h, w = hw
    cfg = sampler_cfg(dict(OAMIX_CFG, version='augmix'))
    img, gt = synth.make_image(5, h, w, 3)
    np.random.seed(31)
    _, plan = oamix_np.oamix_view(img, gt, **cfg)
    np.random.seed(31)
    t = OAMix(**cfg)
    vp = t._sample_head(h, w, gt)
    t._sample_tail(vp, gt, plan['scores'])
    blob = t._pack([(vp, gt, 0)])
    out = np.zeros_like(img)
    hp, wp = -(-h // div) * div, -(-w // div) * div
    v32 = np.full((3, hp, wp), np.nan, np.float32)
    s32 = np.full((3, hp, wp), np.nan, np.float32)
    fo = _FusedOut()
    fo.mean = (ctypes.c_float * 3)(*NORM['mean'])
    fo.std = (ctypes.c_float * 3)(*NORM['std'])
    fo.to_rgb, fo.size_divisor = int(to_rgb), div
    vt = (ctypes.c_void_p * 1)(v32.ctypes.data)
    st = (ctypes.c_void_p * 1)(s32.ctypes.data)
    fo.view_f32, fo.src_f32 = ctypes.cast(vt, ctypes.c_void_p), ctypes.cast(st, ctypes.c_void_p)
    src = (ctypes.c_void_p * 1)(img.ctypes.data)
    dst = (ctypes.c_void_p * 1)(out.ctypes.data)
    rc = hostsim.hostsim_oamix_execute_fused(ctypes.c_void_p(blob.ctypes.data), ctypes.c_size_t(blob.nbytes), src, 1, dst,
                                             ctypes.byref(fo))
    assert rc == 0, rc
    assert np.array_equal(v32, prims_np.imnormalize_pad_chw(out, NORM['mean'], NORM['std'], to_rgb, div))
    assert np.array_equal(s32, prims_np.imnormalize_pad_chw(img, NORM['mean'], NORM['std'], to_rgb, div))
