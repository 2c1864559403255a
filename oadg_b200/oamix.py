"""OA-Mix pipeline transform on B200: host plan sampler + libOADG plan executor.

Drop-in for the reference ``@PIPELINES.register_module() class OAMix``
(mmdet/datasets/pipelines/oa_mix.py:32-313): same constructor keys, same
``transform(results) -> results`` contract and result keys (oa_mix.py:187-204), and
it consumes the global legacy ``np.random`` stream draw for draw in the reference
order, so a seeded run picks the same boxes, ops and mixing
coefficients as the reference.  All pixel work (saliency, colour/geometric ops,
region composite, object-aware mixing) runs in CUDA kernels behind the C ABI in
include/oadg.h; there is no CPU pixel path.

Host/device split: the host draws a *plan* (a few KB), the device executes it.
The one device->host dependency is the per-box saliency score, which gates the
object-aware target list (oa_mix.py:249,295) and therefore later RNG draws.
"""
import math
import os
import warnings

import numpy as np

from . import _lib, plan as P
from .registry import PIPELINES

_AUG = {
    'augmix': ['autocontrast', 'equalize', 'posterize', 'solarize',
               'bbo_rotate', 'bbo_shear_xy', 'bbo_translate_xy',
               'bg_rotate', 'bg_shear_xy', 'bg_translate_xy'],
    'augmix.all': ['autocontrast', 'equalize', 'posterize', 'solarize', 'invert',
                   'color', 'contrast', 'brightness', 'sharpness',
                   'bbo_rotate', 'bbo_shear_xy', 'bbo_translate_xy',
                   'bg_rotate', 'bg_shear_xy', 'bg_translate_xy'],
}


ITEM_KINDS = ('profile', 'mask', 'hist', 'lut', 'copy', 'bbo_blend', 'bbo_catchup', 'step')   # chain-kernel work item kinds


def get_aug_list(version):
    """Names of the ops of reference oa_mix.py:15-29 (same order => same np.random.choice index)."""
    if version not in _AUG:
        raise NotImplementedError
    return _AUG[version]


def _f32(v):
    return float(np.float32(v))


def _invert_affine(m):
    """Forward 2x3 (6 python floats) -> inverse map in doubles, as cv::warpAffine computes it."""
    m = list(m)
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m


def _rotation(center, deg):
    """cv2.getRotationMatrix2D(center, deg, 1.0): Point2f centre, double math (augmix.py:91)."""
    cx, cy = _f32(center[0]), _f32(center[1])
    a = deg * (math.pi / 180.0)
    al, be = math.cos(a), math.sin(a)
    return [al, be, (1 - al) * cx - be * cy, -be, al, be * cx + (1 - al) * cy]


def _iou_1xk(box, boxes):
    """float32 IoU of one box against k boxes (core/evaluation/bbox_overlaps.py:5-65)."""
    b2 = np.asarray(boxes).astype(np.float32)
    if b2.size == 0:
        return np.zeros((1, 0), np.float32)
    b1 = np.asarray(box, dtype=np.float32)
    b2 = b2.reshape(-1, 4)
    a1 = (b1[2] - b1[0]) * (b1[3] - b1[1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    ov = np.maximum(np.minimum(b1[2], b2[:, 2]) - np.maximum(b1[0], b2[:, 0]), 0) * \
        np.maximum(np.minimum(b1[3], b2[:, 3]) - np.maximum(b1[1], b2[:, 1]), 0)
    union = np.maximum(a1 + a2 - ov, np.float32(1e-6))
    return (ov / union).reshape(1, -1)


def _slice(lo, hi, n):
    s, e, _ = slice(int(lo), int(hi)).indices(n)
    return s, max(e, s)


class _NativePlan:
    """What oadg_oamix_sample_plan returns for one batch: the packed blob and the boxes of the result dict."""
    __slots__ = ('blob', 'ml_boxes', 'oa_boxes', 'depth_sums', 'hw')


class _PlanSlice:
    """The result-dict boxes of images [lo, hi) of a (grouped) native plan."""
    __slots__ = ('ml_boxes', 'oa_boxes')

    def __init__(self, plan, lo, hi):
        self.ml_boxes = plan.ml_boxes[lo:hi]
        self.oa_boxes = plan.oa_boxes[lo:hi]


_RNG = None


def _global_rng():
    """oadg_rng_t over numpy's global legacy RandomState (np.random.seed / np.random.* share its MT19937 state)."""
    global _RNG
    import ctypes
    bg = np.random.mtrand._rand._bit_generator
    if _RNG is None or _RNG[0] is not bg:
        i = bg.ctypes

        class _Rng(ctypes.Structure):
            _fields_ = [('state', ctypes.c_void_p), ('next_uint32', ctypes.c_void_p), ('next_double', ctypes.c_void_p)]
        r = _Rng(i.state.value if hasattr(i.state, 'value') else int(i.state),
                 ctypes.cast(i.next_uint32, ctypes.c_void_p).value, ctypes.cast(i.next_double, ctypes.c_void_p).value)
        _RNG = (bg, r, i)   # keep the interface alive: it owns the function pointers
    return _RNG[1]


class _SamplerCfg(__import__('ctypes').Structure):
    _c = __import__('ctypes')
    _fields_ = [('version', _c.c_int32), ('severity', _c.c_int32), ('mixture_width', _c.c_int32),
                ('mixture_depth', _c.c_int32), ('spatial_ratio', _c.c_int32), ('score_thresh', _c.c_int32),
                ('random_box_scale', _c.c_double * 2), ('random_box_ratio', _c.c_double * 2),
                ('oa_random_box_scale', _c.c_double * 2), ('oa_random_box_ratio', _c.c_double * 2),
                ('sigma_ratio', _c.c_double)]


class _FusedOut(__import__('ctypes').Structure):     # oadg_fused_out_t (include/oadg.h)
    _c = __import__('ctypes')
    _fields_ = [('mean', _c.c_float * 3), ('std', _c.c_float * 3), ('to_rgb', _c.c_int32), ('size_divisor', _c.c_int32),
                ('view_f32', _c.c_void_p), ('src_f32', _c.c_void_p)]


class _ViewPlan:
    __slots__ = ('h', 'w', 'ws', 'ml_boxes', 'depths', 'ops', 'scores', 'oa_low', 'oa_boxes', 'm', 'm_oa')


@PIPELINES.register_module()
class OAMix:
    def __init__(self,
                 version='augmix',
                 num_views=2, keep_orig=True, severity=10,
                 mixture_width=3, mixture_depth=-1,
                 random_box_scale=(0.01, 0.1), random_box_ratio=(3, 1 / 3),
                 oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3), num_bboxes=(3, 5),
                 spatial_ratio=4, sigma_ratio=0.3,
                 fused_output=None,
                 **kwargs):
        """Reference keys (oa_mix.py:34-41) plus one opt-in extension: ``fused_output=dict(mean=..., std=...,
        to_rgb=True, size_divisor=32)`` (the config's ``img_norm_cfg`` + ``Pad``) makes the device paths also return
        what Normalize -> Pad -> DefaultFormatBundle produce from the frames: float32 [3, Hp, Wp] CUDA tensors
        ``img_norm`` / ``img2_norm`` (transforms.py:672-704,573-640, formating.py:217-234), written by the mix
        kernel itself."""
        self.aug_list = get_aug_list(version)
        self.fused_output = dict(fused_output) if fused_output else None
        self.last_fused = None
        self.version = version
        self.num_views = num_views
        self.keep_orig = keep_orig
        if self.num_views == 1 and self.keep_orig:
            warnings.warn('No augmentation will be applied since num_views=1 and keep_orig=True')
        self.severity = severity
        self.aug_prob_coeff = 1.0
        self.mixture_width = mixture_width
        self.mixture_depth = mixture_depth
        self.random_box_scale = random_box_scale
        self.random_box_ratio = random_box_ratio
        self.oa_random_box_scale = oa_random_box_scale
        self.oa_random_box_ratio = oa_random_box_ratio
        self.score_thresh = 10
        self.spatial_ratio = spatial_ratio
        self.sigma_ratio = sigma_ratio
        self._history = {}
        self.kwargs = kwargs
        if mixture_width > P.MAX_WIDTH or mixture_depth > P.MAX_DEPTH:
            raise ValueError('mixture_width/depth above the compiled limit (%d/%d)' % (P.MAX_WIDTH, P.MAX_DEPTH))
        if spatial_ratio != 4:
            raise NotImplementedError('libOADG is built for spatial_ratio=4 (all reference configs)')
        self._ws_cache = None
        self._ws_min_views = 0
        self._sal_state = None
        self._sal_prefetch = {}
        self._sal_slot = 0
        self._native_cfg = None
        self._host_state = dict(dev={}, pin={})
        self._streams = {}
        self.pipe_profile = None
        self.pipe_launches = 0
        # iter_batches puts the views of this many consecutive batches into one plan / one chain launch (their work
        # items are independent: a queue with more lanes starves less often; OA-Mix alone on 1024x2048 frames:
        # 265 / 204 / 182 us per view with 1 / 2 / 4 batches of two frames per launch)
        self.group_batches = int(os.environ.get('OADG_GROUP_BATCHES', '4'))
        # iter_batches: a group's chain launch waits for what the consumer has enqueued (on its own stream) for the
        # batches it was already handed, see _pipeline.launch_batches
        self.consumer_fence = os.environ.get('OADG_CONSUMER_FENCE', '1') != '0'
        # iter_batches: groups whose uploads + saliency are enqueued beyond the launched ones (1 measured 12 % slower:
        # the plan then waits for scores whose upload has only just been queued)
        self.stage_ahead = int(os.environ.get('OADG_STAGE_AHEAD', '2'))
        self.first_group = max(1, int(os.environ.get('OADG_FIRST_GROUP', '1')))   # batches in a loop's first group
        self.last_launches = 0

    def __repr__(self):
        return self.__class__.__name__

    # ------------------------------------------------------------------ sampling
    def _sample_regions(self, h, w, scale, ratio, num, gt=None, scores=None, max_iters=50, eps=1e-6):
        """oa_mix.py:122-184.  Draws: randint(*num) once, then per attempt randint(0,w), randint(0,h),
        uniform(*scale), uniform(*ratio); legacy uniform(a, b) is a + (b - a) * random_sample() bit for bit."""
        randint, u01 = np.random.randint, np.random.random_sample
        target = randint(*num) if isinstance(num, tuple) else num
        s_lo, s_w = scale[0], scale[1] - scale[0]
        r_lo, r_w = ratio[0], ratio[1] - ratio[0]
        boxes, bscores = [], []
        for _ in range(max_iters):
            if len(boxes) >= target:
                break
            x1, y1 = int(randint(0, w)), int(randint(0, h))
            area = (s_lo + s_w * u01()) * h * w
            r = r_lo + r_w * u01()
            bw, bh = int(math.sqrt(area / r)), int(math.sqrt(area * r))
            if x1 + bw > w or y1 + bh > h:
                continue
            box = np.array([x1, y1, min(x1 + bw, w), min(y1 + bh, h)])
            if boxes and np.sum(_iou_1xk(box, boxes)) > eps:
                continue
            if gt is not None:
                ious = _iou_1xk(box, gt)
                s = float('inf')
                if np.sum(ious) > eps:
                    for iou, fb, fs in zip(ious[0], gt, scores):
                        if iou == 0.0 or fb[2] - fb[0] < 1 or fb[3] - fb[1] < 1:
                            continue
                        if fs < s:
                            s = fs
                bscores.append(s)
            boxes.append(box)
        return boxes, bscores

    def _level_sign(self, geo):
        """augmix.py:61 sample_level = uniform(0.1, severity), then one sign draw (uniform()/random())."""
        u01 = np.random.random_sample
        level = 0.1 + (self.severity - 0.1) * u01()
        return level, u01() > 0.5

    @staticmethod
    def _forward_affine(geo, level, neg, size_for_level, center, img_size):
        """2x3 forward matrix as augmix.py:83-188 builds it (float32 entries except rotate)."""
        if geo == 'rotate':
            deg = int(level * 30 / 10)
            if neg:
                deg = -deg
            if center is None:
                center = (img_size[0] / 2, img_size[1] / 2)
            return _rotation(center, deg)
        if geo == 'shear_x':
            l = float(level) * 0.3 / 10.
            if neg:
                l = -l
            tx = 0 if center is None else -l * center[1]
            return [1.0, _f32(-l), _f32(-tx), 0.0, 1.0, 0.0]
        if geo == 'shear_y':
            l = float(level) * 0.3 / 10.
            if neg:
                l = -l
            ty = 0 if center is None else -l * center[0]
            return [1.0, 0.0, 0.0, _f32(-l), 1.0, _f32(-ty)]
        if geo == 'translate_x':
            l = int(level * (size_for_level[0] / 3) / 10)
            if neg:
                l = -l
            return [1.0, 0.0, _f32(-l), 0.0, 1.0, 0.0]
        l = int(level * (size_for_level[1] / 3) / 10)
        if neg:
            l = -l
        return [1.0, 0.0, 0.0, 0.0, 1.0, _f32(-l)]

    def _sample_op(self, gt, img_size, gt_int=None):
        """One OAMix.aug() (oa_mix.py:264-279): op index, then the op's own draws.
        np.random.choice(list) draws randint(0, len) (same legacy stream); sample_level as in _level_sign."""
        u01 = np.random.random_sample
        name = self.aug_list[int(np.random.randint(0, len(self.aug_list)))]
        if name == 'autocontrast' or name == 'equalize':
            return (name,)
        if name == 'posterize':
            return (name, 4 - int((0.1 + (self.severity - 0.1) * u01()) * 4 / 10))
        if name == 'solarize':
            return (name, 256 - int((0.1 + (self.severity - 0.1) * u01()) * 256 / 10))
        if name in ('color', 'contrast', 'brightness', 'sharpness'):
            return (name, float(0.1 + (self.severity - 0.1) * u01()) * 1.8 / 10. + 0.1)
        if name == 'invert':
            tx = 1 if u01() > 0.5 else -1
            ty = 1 if u01() > 0.5 else -1
            return (name, tx, ty)
        where, geo = name.split('_', 1)
        if geo == 'shear_xy':
            geo = 'shear_x' if u01() < 0.5 else 'shear_y'
        elif geo == 'translate_xy':
            geo = 'translate_x' if u01() < 0.5 else 'translate_y'
        if where == 'bg':
            level, neg = self._level_sign(geo)
            return ('bg_affine', _invert_affine(self._forward_affine(geo, level, neg, img_size, None, img_size)))
        chain = []
        if gt_int is None:
            gt_int = [(int(b[0]), int(b[1]), int(b[2]), int(b[3])) for b in gt]
        for k, (x1, y1, x2, y2) in enumerate(gt_int):
            if (x2 - x1) < 1 or (y2 - y1) < 1:
                continue  # bbox_augmentation.py:45-47: skipped before any draw
            level, neg = self._level_sign(geo)
            center = ((x1 + x2) / 2., (y1 + y2) / 2.)
            fwd = self._forward_affine(geo, level, neg, (x2 - x1 + 1, y2 - y1 + 1), center, img_size)
            chain.append((k, _invert_affine(fwd)))
        return ('bbo_affine', chain)

    def _sample_head(self, h, w, gt):
        """Draws of oamix() up to (not including) the object-aware part: oa_mix.py:212-234."""
        vp = _ViewPlan()
        vp.h, vp.w = h, w
        vp.ws = np.float32(np.random.dirichlet([self.aug_prob_coeff] * self.mixture_width))
        ml, _ = self._sample_regions(h, w, self.random_box_scale, self.random_box_ratio, (1, 3))
        vp.ml_boxes = np.stack(ml, axis=0)  # ValueError when no box could be placed (oa_mix.py:217)
        vp.depths, vp.ops = [], []
        gt_int = [(int(b[0]), int(b[1]), int(b[2]), int(b[3])) for b in gt]   # bbox_augmentation.py:44
        for _ in range(self.mixture_width):
            depth = self.mixture_depth if self.mixture_depth > 0 else int(np.random.randint(1, 4))
            vp.depths.append(depth)
            vp.ops.append([[self._sample_op(gt, (w, h), gt_int) for _r in range(len(ml) + 1)] for _d in range(depth)])
        return vp

    def _sample_tail(self, vp, gt, scores):
        """Object-aware targets and mixing coefficients: oa_mix.py:245-262,282,295-298."""
        vp.scores = scores
        vp.oa_low = [k for k, s in enumerate(scores) if s <= self.score_thresh]
        vp.oa_boxes, oa_scores = self._sample_regions(
            vp.h, vp.w, self.oa_random_box_scale, self.oa_random_box_ratio,
            min(max(len(vp.oa_low), 1), 5), gt=gt, scores=scores)
        vp.m = np.random.beta(self.aug_prob_coeff, self.aug_prob_coeff)
        tgt_scores = [scores[k] for k in vp.oa_low] + list(oa_scores)
        u01 = np.random.random_sample
        vp.m_oa = [np.float32(0.0 + 0.5 * u01()) if s <= self.score_thresh
                   else np.float32(0.0 + 1.0 * u01()) for s in tgt_scores]
        return vp

    # ------------------------------------------------------------------ native sampler
    def sample_plan(self, hw_list, gt_list, scores):
        """Draw + pack the plan of one batch in libOADG (oadg_oamix_sample_plan): consumes np.random draw for draw like
        _sample_head/_sample_tail below (which remain as the readable statement of the draw order and are checked
        against it byte for byte in tests/test_host.py)."""
        import ctypes
        lib = _lib.load()
        n = len(hw_list)
        cfg = self._native_cfg
        if cfg is None:
            cfg = self._native_cfg = _SamplerCfg(
                0 if self.version == 'augmix' else 1, int(self.severity), int(self.mixture_width),
                int(self.mixture_depth), int(self.spatial_ratio), int(self.score_thresh),
                (ctypes.c_double * 2)(*map(float, self.random_box_scale)),
                (ctypes.c_double * 2)(*map(float, self.random_box_ratio)),
                (ctypes.c_double * 2)(*map(float, self.oa_random_box_scale)),
                (ctypes.c_double * 2)(*map(float, self.oa_random_box_ratio)), float(self.sigma_ratio))
        hw = np.asarray(hw_list, dtype=np.int32).reshape(n, 2)
        gts = [np.ascontiguousarray(g, dtype=np.float32).reshape(-1, 4) for g in gt_list]
        scs = [np.ascontiguousarray([float(v) for v in s], dtype=np.float64) for s in scores]
        n_gt = np.asarray([len(g) for g in gts], dtype=np.int32)
        gt_ptrs = (ctypes.c_void_p * max(n, 1))(*[g.ctypes.data for g in gts])
        sc_ptrs = (ctypes.c_void_p * max(n, 1))(*[s_.ctypes.data for s_ in scs])
        n_bbo_max = int(sum(self.mixture_width * max(self.mixture_depth, 3) * 3 * len(g) for g in gts))
        cap = 4096 + n * (256 + P.OPS_PER_VIEW * P.OP_ST.size) + int(n_gt.sum()) * (P.GT_ST.size + P.TGT_ST.size) \
            + n * 8 * P.TGT_ST.size + n_bbo_max * P.BBO_ST.size
        blob = np.empty(cap, np.uint8)
        nbytes = ctypes.c_size_t(0)
        ml = np.zeros((n, 2, 4), np.int64)
        oa = np.zeros((n, 5, 4), np.int64)
        n_ml = np.zeros(n, np.int32)
        n_oa = np.zeros(n, np.int32)
        dsum = np.zeros(n, np.int32)
        rc = lib.oadg_oamix_sample_plan(ctypes.addressof(_global_rng()), ctypes.addressof(cfg), n, hw.ctypes.data,
                                        ctypes.addressof(gt_ptrs), n_gt.ctypes.data, ctypes.addressof(sc_ptrs),
                                        blob.ctypes.data, cap, ctypes.byref(nbytes), ml.ctypes.data, n_ml.ctypes.data,
                                        oa.ctypes.data, n_oa.ctypes.data, dsum.ctypes.data)
        if rc == -5:   # OADG_E_NOBOX: np.stack([]) in the reference (oa_mix.py:217)
            raise ValueError('need at least one array to stack')
        _lib.check(rc)
        out = _NativePlan()
        out.blob = blob[:nbytes.value]
        out.ml_boxes = [ml[i, :n_ml[i]].copy() for i in range(n)]
        out.oa_boxes = [oa[i, :n_oa[i]].copy() for i in range(n)]
        out.depth_sums = dsum
        out.hw = hw
        return out

    # ------------------------------------------------------------------ records
    def _gt_records(self, gt, h, w):
        """Per gt box: (lo x1,y1,x2,y2, blur, kx, ky, sigma_x, sigma_y, supp x0,y0,x1,y1) -- oa_mix.py:78-91."""
        sr = self.spatial_ratio
        h4, w4 = h // sr, w // sr
        lo_all = np.array(gt // sr, dtype=np.int32).reshape(-1, 4).tolist() if len(gt) else []   # oa_mix.py:79
        out = []
        for x1, y1, x2, y2 in lo_all:
            sx = (x2 - x1) * self.sigma_ratio / 3 * 2
            sy = (y2 - y1) * self.sigma_ratio / 3 * 2
            blur = not (sx <= 0 or sy <= 0)
            xs, xe = _slice(x1, x2, w4)
            ys, ye = _slice(y1, y2, h4)
            kx = (int(round(sx * 8 + 1)) | 1) if blur else 1   # cvRound(sigma*8+1)|1 for non-8U depth
            ky = (int(round(sy * 8 + 1)) | 1) if blur else 1
            supp = [0, 0, 0, 0]
            if xe > xs and ye > ys and w4 > 0 and h4 > 0:
                for ax, (s, e, n_lo, n_hi, ks) in enumerate(((xs, xe, w4, w, kx), (ys, ye, h4, h, ky))):
                    rad = ks // 2 if blur else 0
                    a, bnd = max(s - rad, 0), min(e - 1 + rad, n_lo - 1)
                    up = n_hi / n_lo
                    d0 = int(math.floor((a - 0.5) * up - 0.5)) - 1
                    d1 = int(math.ceil((bnd + 1.5) * up - 0.5)) + 1
                    supp[ax], supp[ax + 2] = max(d0, 0), min(d1, n_hi)
            out.append((xs, ys, xe, ye, int(blur), kx, ky, sx if blur else 1.0, sy if blur else 1.0, supp))
        return out

    def _pack(self, jobs):
        """jobs: list of (view_plan, gt, img_index).  Returns the plan blob (uint8 array)."""
        n_gt = sum(len(gt) for _, gt, _ in jobs)
        n_tgt = sum(len(vp.oa_low) + len(vp.oa_boxes) for vp, _, _ in jobs)
        n_bbo = sum(len(op[1]) for vp, _, _ in jobs for steps in vp.ops for regs in steps for op in regs
                    if op[0] == 'bbo_affine')
        B = P.BlobBuilder(len(jobs), n_gt, n_bbo, n_tgt)
        g0 = b0 = t0 = 0
        max_h = max_w = 1
        for v, (vp, gt, img_i) in enumerate(jobs):
            max_h, max_w = max(max_h, vp.h), max(max_w, vp.w)
            for k, (xs, ys, xe, ye, blur, kx, ky, sx, sy, supp) in enumerate(self._gt_records(gt, vp.h, vp.w)):
                B.gt(g0 + k, xs, ys, xe, ye, blur, kx, ky, v, sx, sy, *supp)
            base = v * P.OPS_PER_VIEW
            for b, steps in enumerate(vp.ops):
                for d, regs in enumerate(steps):
                    for r, op in enumerate(regs):
                        i = base + (b * P.MAX_DEPTH + d) * P.MAX_REGIONS + r
                        name = op[0]
                        if name == 'bbo_affine':
                            B.op(i, P.OP[name], bbo_first=b0, bbo_count=len(op[1]))
                            for k, minv in op[1]:
                                B.bbo(b0, g0 + k, minv)
                                b0 += 1
                        elif name == 'bg_affine':
                            B.op(i, P.OP[name], minv=op[1])
                        elif name == 'posterize' or name == 'solarize':
                            B.op(i, P.OP[name], p0=op[1])
                        elif name == 'invert':
                            B.op(i, P.OP[name], p0=op[1], p1=op[2])
                        elif len(op) > 1:
                            B.op(i, P.OP[name], factor=op[1])
                        else:
                            B.op(i, P.OP[name])
            nt = 0
            for k, m_oa in zip(vp.oa_low, vp.m_oa):
                B.tgt(t0 + nt, 0, g0 + k, (0, 0, 0, 0), float(m_oa))
                nt += 1
            for bx, m_oa in zip(vp.oa_boxes, vp.m_oa[len(vp.oa_low):]):
                xs, xe = _slice(bx[0], bx[2], vp.w)
                ys, ye = _slice(bx[1], bx[3], vp.h)
                B.tgt(t0 + nt, 1, -1, (xs, ys, xe, ye), float(m_oa))
                nt += 1
            ml = [int(c) for bx in vp.ml_boxes for c in bx] + [0] * (8 - 4 * len(vp.ml_boxes))
            nb = len(vp.depths)
            B.view(v, vp.h, vp.w, img_i, len(gt), g0, len(vp.ml_boxes), *ml, nb,
                   *(list(vp.depths) + [0] * (P.MAX_WIDTH - nb)),
                   *([float(x) for x in vp.ws] + [0.0] * (P.MAX_WIDTH - nb)),
                   base, nt, t0, 0, float(vp.m))
            g0 += len(gt)
            t0 += nt
        return B.finish(max_h, max_w)

    # ------------------------------------------------------------------ device
    def saliency_scores(self, imgs, gt_list, stream=None, inputs_ready=False):
        """Per image: list of saliency scores (oa_mix.py:100-111); one kernel + one D2H for the batch.

        The scores gate later RNG draws, so the host must wait for them.  With ``inputs_ready=True`` (the caller
        guarantees the frames are complete, e.g. they are resident from an earlier step) the kernel and its
        read-back run on a private stream, so the wait does not drain work queued on the compute stream."""
        return self._saliency_collect(self._saliency_launch(imgs, gt_list, stream, inputs_ready))

    @staticmethod
    def _saliency_key(imgs, gt_list):
        return tuple((int(t.data_ptr()), tuple(t.shape)) for t in imgs), tuple(np.asarray(g).tobytes() for g in gt_list)

    def _saliency_launch(self, imgs, gt_list, stream=None, inputs_ready=False, slot=0):
        """Enqueue the saliency kernel + score read-back of a batch; returns a handle for _saliency_collect."""
        torch = _lib.require_cuda()
        lib = _lib.load()
        sr = self.spatial_ratio
        rows, slots = [], []
        out = []
        for i, (img, gt) in enumerate(zip(imgs, gt_list)):
            h, w = int(img.shape[0]), int(img.shape[1])
            sc = []
            boxes_i = np.array(gt, dtype=np.int32).reshape(-1, 4).tolist() if len(gt) else []   # oa_mix.py:102
            for k, (x1, y1, x2, y2) in enumerate(boxes_i):
                if x2 - x1 < sr or y2 - y1 < sr:
                    sc.append(-1)
                    continue
                xs, xe = _slice(x1, x2, w)
                ys, ye = _slice(y1, y2, h)
                if xe <= xs or ye <= ys:
                    raise ValueError('empty crop for gt box %s (cv2 would raise on an empty image)' % (gt[k],))
                rows.append((i, xs, ys, xe, ye))
                slots.append((i, k))
                sc.append(None)
            out.append(sc)
        st = None
        if rows:
            dev = imgs[0].device
            cur = torch.cuda.current_stream(dev) if stream is None else stream
            n_img, n = len(imgs), len(rows)
            # one packed upload: [img pointers i64 | (H, W) i32 pairs | boxes n x 5 i32], 8-byte aligned sections
            o_hw = 8 * n_img
            o_box = o_hw + 8 * n_img
            nbytes = o_box + 20 * n
            if self._sal_state is None:
                self._sal_state = {}
            st = self._sal_state.get(slot)
            if st is None or st['cap'] < nbytes or st['dev'] != dev:
                cap = max(4096, 2 * nbytes)
                side = self._side_stream(dev)
                with torch.cuda.stream(side):   # device buffers owned by the stream that uses them (allocator pools
                    devbuf = torch.empty(cap, dtype=torch.uint8, device=dev)          # are per stream: no block that a
                    scores = torch.empty(cap // 8, dtype=torch.float64, device=dev)   # main-stream kernel still reads)
                st = self._sal_state[slot] = dict(
                    cap=cap, dev=dev, host=torch.empty(cap, dtype=torch.uint8).pin_memory(), devbuf=devbuf, scores=scores,
                    scores_host=torch.empty(cap // 8, dtype=torch.float64).pin_memory(), stream=side,
                    event=torch.cuda.Event())
            hb = st['host'].numpy()
            hb[:o_hw].view(np.int64)[:] = [int(t.data_ptr()) for t in imgs]
            hb[o_hw:o_box].view(np.int32)[:] = [v for t in imgs for v in (int(t.shape[0]), int(t.shape[1]))]
            hb[o_box:nbytes].view(np.int32)[:] = [v for r in rows for v in r]
            s = st['stream'] if inputs_ready else cur
            base, raw = st['devbuf'].data_ptr(), s.cuda_stream
            _lib.check(lib.oadg_memcpy_async(base, st['host'].data_ptr(), nbytes, 1, raw))
            _lib.check(lib.oadg_saliency_scores(base, base + o_hw, base + o_box, n, st['scores'].data_ptr(), raw))
            _lib.check(lib.oadg_memcpy_async(st['scores_host'].data_ptr(), st['scores'].data_ptr(), 8 * n, 0, raw))
            st['event'].record(s)
            self.last_launches += 1
        return dict(out=out, slots=slots, st=st, n=len(rows))

    def _side_stream(self, dev):
        """The private stream of the saliency kernels (and of iter_batches' uploads), one per device."""
        torch = _lib.require_cuda()
        side = self._streams.get(('side', str(dev)))
        if side is None:
            # high priority: the persistent chain kernels hold every SM slot until they finish, and the small
            # saliency kernel (whose scores the NEXT plan waits for) must get the first slot that frees up
            side = self._streams[('side', str(dev))] = torch.cuda.Stream(dev, priority=-1)
        return side

    def _saliency_collect(self, handle):
        out, st = handle['out'], handle['st']
        if st is not None:
            st['event'].synchronize()  # the one device->host sync of the path
            host = st['scores_host'][:handle['n']].numpy()
            for (i, k), v in zip(handle['slots'], host):
                out[i][k] = np.float64(v)
        return out

    def prefetch_saliency(self, imgs, gt_list, tag=None):
        """Enqueue the saliency scores of an UPCOMING batch (frames already resident on the device) so that its
        ``oamix_batch`` call finds them finished: the scores consume no random numbers and depend on nothing but
        the frames and boxes, so a loader that knows the next batch can hide the kernel and its read-back behind
        the current step.  Returns a handle to pass as ``oamix_batch(..., saliency=handle)``.

        Without a handle ``oamix_batch`` looks the request up by (frame pointers, shapes, boxes): a loader that
        REWRITES a frame buffer in place between the prefetch and the call must pass the handle or a ``tag`` of its
        own (e.g. the step number), otherwise the scores of the buffer's previous content would be used.  At most
        three requests are outstanding; a fourth first completes and drops the oldest (its staging slot is reused)."""
        gt_list = [np.asarray(g, dtype=np.float32).reshape(-1, 4) for g in gt_list]
        busy = {h['slot'] for h in self._sal_prefetch.values()}
        free = [k for k in (1, 2, 3) if k not in busy]       # staging slot 0 serves un-prefetched calls
        if not free:
            oldest = self._sal_prefetch.pop(next(iter(self._sal_prefetch)))
            self._saliency_collect(oldest)                     # its copies must be over before the slot is reused
            free = [oldest['slot']]
        handle = self._saliency_launch(imgs, gt_list, None, True, slot=free[0])
        handle['slot'] = free[0]
        key = (tag, self._saliency_key(imgs, gt_list))
        self._sal_prefetch[key] = handle
        return key

    def _workspace(self, nbytes, device, n_views=0, max_hw=(0, 0)):
        """Scratch for oadg_oamix_execute.  What a plan needs varies with its bboxes-only chains (two frames each);
        growing the buffer drains the device, so the first allocation already covers the largest plan this
        configuration can draw for as many views of this size (every op a chain), up to 8 GiB."""
        torch = _lib.require_cuda()
        if self._ws_cache is None or self._ws_cache.numel() < nbytes or self._ws_cache.device != device:
            # The old workspace may still be in use by kernels in flight; once released, the caching allocator may hand
            # its block to a buffer that is used on ANOTHER stream (the saliency stream) right away.  Growing is rare:
            # drain the device first.
            if self._ws_cache is not None:
                torch.cuda.synchronize(self._ws_cache.device)
            self._ws_cache = None
            depth = self.mixture_depth if self.mixture_depth > 0 else 3
            frame = (int(max_hw[0]) * int(max_hw[1]) * 3 + 255) // 256 * 256
            worst = n_views * (frame * (2 * self.mixture_width + 2 * self.mixture_width * depth * P.MAX_REGIONS) +
                               int(max_hw[0]) * int(max_hw[1]) * 5) + (8 << 20)
            want = max(int(nbytes * 1.25), min(worst, 8 << 30))
            self._ws_cache = torch.empty(want + 4096, dtype=torch.uint8, device=device)
        return self._ws_cache

    def execute(self, jobs, imgs, outs=None, stream=None, profile=None, fused=None):
        """Run the packed plans.  jobs: a packed plan blob (uint8 array, one view per image) or a list of
        (view_plan, gt, img_index); imgs: list of CUDA u8 HWC tensors; returns one output tensor per view.
        ``profile`` (a dict) switches to the event-timed entry point and receives per-kernel ms / counts."""
        import ctypes
        torch = _lib.require_cuda()
        lib = _lib.load()
        dev = imgs[0].device
        if isinstance(jobs, np.ndarray):
            blob = jobs
            jobs = [(None, None, i) for i in range(len(imgs))]
        else:
            blob = self._pack(jobs)
        need = ctypes.c_size_t(0)
        _lib.check(lib.oadg_oamix_workspace_bytes(blob.ctypes.data, blob.nbytes, ctypes.byref(need)))
        ws = self._workspace(need.value, dev, max(len(jobs), self._ws_min_views),
                             (max(int(t.shape[0]) for t in imgs), max(int(t.shape[1]) for t in imgs)))
        if outs is None:
            outs = [torch.empty_like(imgs[j[2]]) for j in jobs]
        src = (ctypes.c_void_p * len(imgs))(*[int(t.data_ptr()) for t in imgs])
        dst = (ctypes.c_void_p * len(outs))(*[int(t.data_ptr()) for t in outs])
        s_raw = _lib.raw_stream(dev) if stream is None else stream.cuda_stream
        base = (ws.data_ptr() + 255) // 256 * 256
        room = ws.numel() - (base - ws.data_ptr())
        if profile is not None:
            ms_chain, ms_mix = ctypes.c_float(0), ctypes.c_float(0)
            n_items, n_tiles = ctypes.c_int(0), ctypes.c_int(0)
            kstats = np.zeros(48, np.uint64)
            _lib.check(lib.oadg_oamix_execute_profiled(blob.ctypes.data, blob.nbytes, src, len(imgs), dst, base, room,
                                                       ctypes.byref(ms_chain), ctypes.byref(ms_mix), ctypes.byref(n_items),
                                                       ctypes.byref(n_tiles), kstats.ctypes.data, s_raw))
            profile['chain_ms'] = profile.get('chain_ms', 0.0) + float(ms_chain.value)
            profile['mix_ms'] = profile.get('mix_ms', 0.0) + float(ms_mix.value)
            profile['chain_n'] = profile.get('chain_n', 0) + 1
            profile['mix_n'] = profile.get('mix_n', 0) + 1
            profile['items'] = profile.get('items', 0) + int(n_items.value)
            profile['tiles'] = profile.get('tiles', 0) + int(n_tiles.value)
            names = ITEM_KINDS[:7] + ('step_stream', 'step_bg_staged', 'step_mixed', 'dependency_wait', 'claim_fast', 'claim_blocked', 'tile_end_sync')
            ks = profile.setdefault('kind_busy_us_and_tiles', {k: [0.0, 0, 0.0] for k in names})
            for i, k in enumerate(names):
                ks[k][0] += float(kstats[i]) / 1e3
                ks[k][1] += int(kstats[16 + i])
                ks[k][2] = max(ks[k][2], float(kstats[32 + i]) / 1e3)
            self.last_launches += 2
            return outs
        n = ctypes.c_int(0)
        if fused is not None:   # (view_f32 tensors, src_f32 tensors or None): the mix kernel's Normalize + Pad + CHW epilogue
            views32, srcs32 = fused
            cfg = self.fused_output
            fo = _FusedOut()
            fo.mean = (ctypes.c_float * 3)(*[float(v) for v in cfg['mean']])
            fo.std = (ctypes.c_float * 3)(*[float(v) for v in cfg['std']])
            fo.to_rgb = int(bool(cfg.get('to_rgb', True)))
            fo.size_divisor = int(cfg.get('size_divisor', 1) or 1)
            vt = (ctypes.c_void_p * len(views32))(*[int(t.data_ptr()) for t in views32])
            st_ = (ctypes.c_void_p * len(imgs))(*[int(t.data_ptr()) if t is not None else None for t in srcs32]) \
                if srcs32 is not None else None
            fo.view_f32 = ctypes.cast(vt, ctypes.c_void_p)
            fo.src_f32 = ctypes.cast(st_, ctypes.c_void_p) if st_ is not None else None
            _lib.check(lib.oadg_oamix_execute_fused(blob.ctypes.data, blob.nbytes, src, len(imgs), dst,
                                                    ctypes.addressof(fo), base, room, ctypes.byref(n), s_raw))
            self.last_launches += n.value
            return outs
        _lib.check(lib.oadg_oamix_execute(blob.ctypes.data, blob.nbytes, src, len(imgs), dst, base, room,
                                          ctypes.byref(n), s_raw))
        self.last_launches += n.value
        return outs

    def oamix_batch(self, imgs, gt_list, stream=None, profile=None, outs=None, inputs_ready=False, saliency=None):
        """Device fast path: one generated view per image (the ``num_views=2, keep_orig=True`` case).

        imgs: list of CUDA uint8 HWC tensors; gt_list: list of float32 [n,4] arrays.
        Returns (views, oamix_boxes, multilevel_boxes) with the reference's dtypes."""
        torch = _lib.require_cuda()
        self.last_launches = 0
        for t in imgs:
            if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3 and t.is_contiguous()):
                raise TypeError('images must be contiguous CUDA uint8 HWC tensors')
        gt_list = [np.asarray(g, dtype=np.float32).reshape(-1, 4) for g in gt_list]
        if saliency is not None:                      # the handle prefetch_saliency() returned
            if saliency not in self._sal_prefetch:
                raise KeyError('oamix_batch: unknown or already consumed saliency handle')
            if saliency[1] != self._saliency_key(imgs, gt_list):
                raise ValueError('oamix_batch: the saliency handle belongs to other frames / boxes')
            pre = self._sal_prefetch.pop(saliency)
        else:
            pre = self._sal_prefetch.pop((None, self._saliency_key(imgs, gt_list)), None) if self._sal_prefetch else None
        if pre is not None:
            scores = self._saliency_collect(pre)      # requested earlier by prefetch_saliency()
            self.last_launches = 1
        else:
            scores = self.saliency_scores(imgs, gt_list, stream, inputs_ready=inputs_ready)
        plan = self.sample_plan([(int(t.shape[0]), int(t.shape[1])) for t in imgs], gt_list, scores)
        fused = None
        if self.fused_output is not None and profile is None:
            fused = self.fused_buffers(imgs)
            self.last_fused = fused
        outs = self.execute(plan.blob, imgs, outs=outs, stream=stream, profile=profile, fused=fused)
        if profile is not None:  # algorithmic bytes: 1 read + 1 write of a frame per lane step (+ the mix, per view)
            for (h, w), dsum in zip(plan.hw, plan.depth_sums):
                profile['step_bytes'] = profile.get('step_bytes', 0) + 2 * 3 * int(h) * int(w) * int(dsum)
                profile['view_bytes'] = profile.get('view_bytes', 0) + \
                    3 * int(h) * int(w) * (2 * int(dsum) + self.mixture_width + 2)
        oamix_boxes = [np.stack(list(b), axis=0) for b in plan.oa_boxes]   # ValueError when none could be placed
        return outs, oamix_boxes, plan.ml_boxes

    def fused_buffers(self, imgs):
        """(view_f32, src_f32): float32 [3, Hp, Wp] CUDA tensors for the fused Normalize + Pad + CHW output of every
        generated view and of every source frame."""
        torch = _lib.require_cuda()
        d = int(self.fused_output.get('size_divisor', 1) or 1)

        def buf(t):
            h, w = int(t.shape[0]), int(t.shape[1])
            return torch.empty(3, -(-h // d) * d, -(-w // d) * d, dtype=torch.float32, device=t.device)
        return [buf(t) for t in imgs], [buf(t) for t in imgs]

    # ------------------------------------------------------------------ host buffers
    def _to_device(self, img, slot, stream=None):
        """Host uint8 HWC array -> CUDA tensor on the current stream (or ``stream``).  Page-locked arrays (e.g. a
        loader's pinned buffers) are copied asynchronously in place; pageable ones go through a persistent pinned
        staging buffer."""
        torch = _lib.require_cuda()
        img = np.ascontiguousarray(np.asarray(img, dtype=np.uint8))
        src = torch.from_numpy(img)
        st = self._host_state
        key = (slot, tuple(img.shape))
        dev = st['dev'].get(key)
        if dev is None:
            dev = st['dev'][key] = torch.empty(img.shape, dtype=torch.uint8, device='cuda')
            if isinstance(slot, tuple) and len(slot) == 3:   # (slot, image, slots): a loader's ring, allocated at once
                for k in range(slot[2]):
                    st['dev'].setdefault(((k, slot[1], slot[2]), tuple(img.shape)),
                                         torch.empty(img.shape, dtype=torch.uint8, device='cuda'))
        if not src.is_pinned():
            pin = st['pin'].get(key)
            if pin is None:
                pin = st['pin'][key] = torch.empty(img.shape, dtype=torch.uint8).pin_memory()
            pin.copy_(src)
            src = pin
        if stream is None:
            dev.copy_(src, non_blocking=True)
        else:
            _lib.check(_lib.load().oadg_memcpy_async(dev.data_ptr(), src.data_ptr(), src.numel(), 1, stream.cuda_stream))
        return dev, img

    def _pinned_out(self, shape, reserve=0):
        """(page-locked uint8 tensor, numpy array over it) for one generated view.  Pinning memory costs tens of
        milliseconds per view, so the buffers are recycled: when the caller drops the array (and every view of it), a
        finalizer puts the buffer back on the free list.  Buffers are flat and pooled by size class (1 MiB steps), so
        frames of different shapes share them; ``reserve`` buffers of the class are pinned at first use (the loader
        loop asks for as many as it can have in flight, so that it never pins in steady state), up to 2 GiB in all."""
        import weakref
        torch = _lib.require_cuda()
        n = int(np.prod(shape))
        cls = max(1, -(-n // (1 << 20))) << 20
        pool = self._host_state.setdefault('out_pool', dict(free={}, made={}, bytes=0))
        free = pool['free'].setdefault(cls, [])
        while not free or (pool['made'].get(cls, 0) < reserve and pool['bytes'] + cls <= (2 << 30)):
            free.append(torch.empty(cls, dtype=torch.uint8, pin_memory=True))
            pool['made'][cls] = pool['made'].get(cls, 0) + 1
            pool['bytes'] += cls
        t = free.pop()
        view = t[:n].view(tuple(shape))
        a = view.numpy()
        weakref.finalize(a, free.append, t)
        return view, a

    def _to_host(self, outs):
        """CUDA views -> numpy arrays backed by page-locked memory (one sync for all of them)."""
        torch = _lib.require_cuda()
        host = [self._pinned_out(o.shape) for o in outs]
        for (h_, _), o in zip(host, outs):
            h_.copy_(o, non_blocking=True)
        torch.cuda.current_stream(outs[0].device).synchronize()
        _lib.check(_lib.load().oadg_oamix_poll_fault(1))   # a launch that left its views incomplete raises here
        return [a for _, a in host]

    def oamix(self, img, gt_bboxes):
        """One view of one host image (reference oa_mix.py:207-243): H2D, kernels, D2H."""
        gt = np.asarray(gt_bboxes, dtype=np.float32).reshape(-1, 4)
        dimg, img = self._to_device(img, 0)
        scores = self.saliency_scores([dimg], [gt])[0]      # no RNG draw: may precede the plan head
        plan = self.sample_plan([img.shape[:2]], [gt], [scores])
        self._history.update(random_box_list=plan.ml_boxes[0], fg_box_list=gt, fg_score_list=scores,
                             oa_random_box_list=list(plan.oa_boxes[0]))
        out = self.execute(plan.blob, [dimg])[0]
        return self._to_host([out])[0]

    def call_batch(self, results_list):
        """The transform on a list of sample dicts at once (what a collate-level hook calls): the same result keys
        and the same np.random consumption as calling the transform on each dict in order (oa_mix.py:187-204), with
        one H2D / kernel chain / D2H for the whole list.  Falls back to per-sample calls for configurations other
        than num_views=2, keep_orig=True."""
        if not (self.num_views == 2 and self.keep_orig):
            return [self(r) for r in results_list]
        gts = [np.asarray(r['gt_bboxes'], dtype=np.float32).reshape(-1, 4) for r in results_list]
        staged = [self._to_device(r['img'], i) for i, r in enumerate(results_list)]
        dimgs = [d for d, _ in staged]
        scores = self.saliency_scores(dimgs, gts)
        plan = self.sample_plan([h.shape[:2] for _, h in staged], gts, scores)
        outs = self._to_host(self.execute(plan.blob, dimgs))
        return self._fill_results(results_list, outs, plan)

    def _fill_results(self, results_list, outs, plan):
        for r, out, oa, ml in zip(results_list, outs, plan.oa_boxes, plan.ml_boxes):
            r['custom_field'] = []
            r['img_fields'] = ['img', 'img2']
            r['img2'] = out
            r['gt_bboxes2'] = r['gt_bboxes'].copy()
            r['oamix_boxes'] = np.stack(list(oa), axis=0)
            r['custom_field'] += ['img2', 'gt_bboxes2', 'oamix_boxes']
            r['multilevel_boxes'] = ml
            r['custom_field'] += ['multilevel_boxes']
        return results_list

    def iter_batches(self, batches, threaded=True):
        """``call_batch`` for a loader loop, pipelined: yields ``call_batch(b)`` for every ``b`` of ``batches`` (lists
        of sample dicts), in order and with the same values, while the NEXT batch's kernel chain and the one after
        that's upload + saliency scores are already in flight, so host<->device copies, the score read-back and the
        host sampling overlap the kernels instead of adding to them.  The pipeline has its own CUDA streams (like a
        loader worker): what the caller enqueues on its stream between batches neither waits for nor delays it.
        The views of up to ``group_batches`` consecutive batches share one plan and one chain launch (see
        ``_pipeline``).  With ``threaded`` (default) the pipeline's host side runs in a worker thread,
        so its plan sampling and scheduling (native code, GIL released) overlap the caller's own host work.

        Samples whose ``img`` is a CUDA uint8 tensor (complete when the batch is read from ``batches``) stay on the
        device: no upload, ``img2`` is a CUDA tensor ordered on the consumer's current stream; the consumer must
        enqueue its work on a batch's views before it asks for the next batch (that work is fenced before the
        buffers are reused, 3 * group_batches + 4 batches later).

        Host views are numpy arrays over page-locked buffers that go back to the transform's pool when the last
        reference to them is dropped; the transform stores them in the sample dicts it was given (``img2``), so a
        caller that keeps all its dicts (a materialised list of batches) keeps all the buffers, and the loop pins new
        ones as it goes (milliseconds each).  An iterable with ``__len__`` lets the loop size its groups so that no
        short group is left for the end.

        Differences from calling ``call_batch`` in a loop: ``batches`` is read a few items ahead (the caller must
        leave a batch's input arrays alone until it is yielded), and the np.random draws of later batches are taken
        before earlier ones are yielded: the global stream is consumed in the same order, so results match as long
        as the consumer draws nothing from np.random (and does not use this transform) while iterating.  A batch
        that cannot be processed raises when it is its turn to be yielded.  Configurations other than
        num_views=2, keep_orig=True are served by ``call_batch`` without pipelining."""
        if not (self.num_views == 2 and self.keep_orig):
            for b in batches:
                yield self.call_batch(b)
            return
        torch = _lib.require_cuda()
        dev = torch.device('cuda', torch.cuda.current_device())
        released = {}     # batch index -> event on the consumer's stream: its work on that batch's device views
        self.pipe_launches = 0

        def hand_over(item):
            # consumer thread: device views become valid on its current stream; when it comes back for the next
            # batch, whatever it enqueued on them is fenced by an event the pipeline waits for before reuse
            res, job = item
            if job.get('device'):
                torch.cuda.current_stream(dev).wait_event(job['done'])
            return res, job

        def release(job):
            if job.get('device') or self.consumer_fence:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                if job.get('device'):
                    released[job['idx']] = ev
                    released.pop(job['idx'] - 128, None)
                released['latest'] = ev

        if not threaded:
            for item in self._pipeline(batches, dev, released):
                res, job = hand_over(item)
                yield res
                release(job)
            return
        import queue
        import threading
        q, stop = queue.Queue(maxsize=1), threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.05)
                    return True
                except queue.Full:
                    pass
            return False

        def work():
            try:
                torch.cuda.set_device(dev)
                for item in self._pipeline(batches, dev, released):
                    if not put(('ok', item)):
                        return
                put(('end', None))
            except BaseException as e:   # delivered to the consumer in order
                put(('err', e))

        th = threading.Thread(target=work, name='oamix-pipeline', daemon=True)
        th.start()
        try:
            while True:
                kind, val = q.get()
                if kind == 'ok':
                    res, job = hand_over(val)
                    yield res
                    release(job)
                elif kind == 'err':
                    raise val
                else:
                    return
        finally:
            stop.set()
            th.join(timeout=10)

    def _stream(self, name, dev):
        torch = _lib.require_cuda()
        st = self._streams.get((name, str(dev)))
        if st is None:
            st = self._streams[(name, str(dev))] = torch.cuda.Stream(dev)
        return st

    @staticmethod
    def _group_sizes(total, gmax, first=1):
        """Batches per group for a loop of known length: the ramp-up first, 2 * first, ... then full groups.  Batches
        that would be left over for a short last group (a launch with few lanes runs at a lower rate) are folded into
        the ramp-up groups instead."""
        sizes, left, g = [], total, max(1, first)
        while left > 0 and g < gmax:
            sizes.append(min(g, left))
            left -= sizes[-1]
            g *= 2
        n_ramp = len(sizes)
        full, rem = divmod(left, gmax)
        for k in range(n_ramp):            # fold the remainder into the ramp-up, first group first
            add = min(rem, gmax - sizes[k])
            sizes[k] += add
            rem -= add
        sizes += [gmax] * full
        if rem:
            sizes.append(rem)
        return sizes

    def _pipeline(self, batches, dev, released):
        """iter_batches' generator: yields (results, job) per batch, in order.

        Batches travel in GROUPS: the views of up to ``group_batches`` consecutive batches are sampled into ONE plan
        and executed by ONE chain launch (they are independent, oa_mix.py:187-204, and the more lanes the launch's
        work queue holds the fewer CTAs ever wait for a dependency).  The first groups of a loop are smaller (1, 2,
        then ``group_batches``) so that the first batch is not held back.  While the consumer works through group
        k, group k + 1's launch is in flight and group k + 2's upload + saliency kernel are enqueued.  np.random is
        consumed image by image in batch order, exactly as by per-batch calls."""
        import collections
        import time
        torch = _lib.require_cuda()
        lib = _lib.load()
        side, cout = self._side_stream(dev), self._stream('out', dev)
        pipe = self._stream('pipe', dev)
        it = iter(batches)
        staged, launched = collections.deque(), collections.deque()
        count = [0, 0]                    # batches read, groups staged
        exhausted = [False]
        prof = self.pipe_profile          # optional dict: host seconds per phase (scripts/e2e_profile.py)
        gmax = max(1, int(self.group_batches))
        sizes = self._group_sizes(len(batches), gmax, self.first_group) if hasattr(batches, '__len__') else None
        n_sets = 3 * gmax + 4             # view buffer sets: more than the batches launched ahead + being consumed
        n_in = 5 * gmax + 4               # frame staging slots (host frames): more than the batches staged + in flight

        def tick(name, t0):
            if prof is not None:
                dt = time.perf_counter() - t0
                prof[name] = prof.get(name, 0.0) + dt
                prof[name + '.max'] = max(prof.get(name + '.max', 0.0), dt)
            return time.perf_counter()

        def stage_group():
            """Read the next group's batches, upload their frames and enqueue ONE saliency kernel for all of them."""
            if exhausted[0]:
                return
            want = sizes[count[1]] if sizes is not None and count[1] < len(sizes) else \
                min(gmax, self.first_group << count[1])   # 1, 2, 4, ... batches per group
            jobs = []
            while len(jobs) < want:
                try:
                    results_list = next(it)
                except StopIteration:
                    exhausted[0] = True
                    break
                idx = count[0]
                count[0] += 1
                t0 = time.perf_counter()
                job = dict(results=results_list, idx=idx, error=None, device=False)
                try:   # a failure surfaces when the batch is yielded, after the batches before it
                    gts = [np.asarray(r['gt_bboxes'], dtype=np.float32).reshape(-1, 4) for r in results_list]
                    first = results_list[0]['img'] if results_list else None
                    if torch.is_tensor(first) and first.is_cuda:
                        dimgs = [r['img'] for r in results_list]
                        for t in dimgs:
                            if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 3 and t.shape[2] == 3 and
                                    t.is_contiguous()):
                                raise TypeError('images must be contiguous CUDA uint8 HWC tensors')
                        job.update(gts=gts, dimgs=dimgs, hw=[(int(t.shape[0]), int(t.shape[1])) for t in dimgs],
                                   device=True)
                    else:
                        ins = [self._to_device(r['img'], (idx % n_in, i, n_in), side) for i, r in enumerate(results_list)]
                        job.update(gts=gts, dimgs=[d for d, _ in ins], hw=[h.shape[:2] for _, h in ins])
                    tick('upload_enqueue', t0)
                except Exception as e:
                    job['error'] = e
                jobs.append(job)
                if job['error'] is not None:
                    break                          # the group ends at a failed batch
            if not jobs:
                return
            grp = dict(jobs=jobs, idx=count[1], ready=None, sal=None)
            count[1] += 1
            good = [j for j in jobs if j['error'] is None]
            if good:
                t0 = time.perf_counter()
                try:
                    if not all(j['device'] for j in good):
                        grp['ready'] = torch.cuda.Event()
                        grp['ready'].record(side)
                    imgs = [d for j in good for d in j['dimgs']]
                    gts = [g for j in good for g in j['gts']]
                    grp['sal'] = self._saliency_launch(imgs, gts, None, True, slot=4 + grp['idx'] % 3)
                    self.pipe_launches += 1 if grp['sal']['st'] is not None else 0
                except Exception as e:
                    for j in good:
                        j['error'] = e
                tick('saliency_enqueue', t0)
            staged.append(grp)

        def launch_batches(jobs, scores):
            """Sample ONE plan for `jobs` (consecutive batches) and enqueue its execution + the downloads."""
            t0 = time.perf_counter()
            hw = [x for j in jobs for x in j['hw']]
            gts = [g for j in jobs for g in j['gts']]
            imgs = [d for j in jobs for d in j['dimgs']]
            plan = self.sample_plan(hw, gts, scores)
            t0 = tick('sample_plan', t0)
            douts, k = [], 0
            for j in jobs:
                shapes = tuple(tuple(d.shape) for d in j['dimgs'])
                key = ('outs', j['idx'] % n_sets, n_sets, shapes)
                o = self._host_state['dev'].get(key)
                if o is None:   # all view sets of this batch shape at once: allocation belongs to the warm-up
                    for k2 in range(n_sets):
                        self._host_state['dev'].setdefault(('outs', k2, n_sets, shapes),
                                                           [torch.empty_like(d) for d in j['dimgs']])
                    o = self._host_state['dev'][key]
                douts.extend(o)
                # this batch's slice of the group's plan
                n = len(j['dimgs'])
                j['plan'] = _PlanSlice(plan, k, k + n)
                j['douts'] = o
                k += n
                if j['device'] and (j['idx'] - n_sets) in released:   # the consumer's work on this set's last views
                    pipe.wait_event(released[j['idx'] - n_sets])
            if self.consumer_fence and released.get('latest') is not None:
                # the persistent chain kernel takes every SM for its whole run: let what the consumer has enqueued on
                # the batches it already holds (its loss kernels, and in a multi-rank job the exchange they wait on)
                # go first, so that the step alternates [chain of group k + 2] [consumer's work on group k]
                pipe.wait_event(released['latest'])
            before = self.last_launches
            self._ws_min_views = max(self._ws_min_views, gmax * max(len(j['dimgs']) for j in jobs))
            self.execute(plan.blob, imgs, outs=douts, stream=pipe)
            self.pipe_launches += self.last_launches - before
            t0 = tick('execute_enqueue', t0)
            done = torch.cuda.Event()
            done.record(pipe)
            for j in jobs:
                j['done'] = done
                if j['device']:
                    j['views'] = j['douts']
            host_jobs = [j for j in jobs if not j['device']]
            if host_jobs:
                cout.wait_event(done)
                for j in host_jobs:
                    host = [self._pinned_out(o.shape, reserve=(3 * gmax + 6) * len(j['douts'])) for o in j['douts']]
                    for (h_, _), o in zip(host, j['douts']):
                        _lib.check(lib.oadg_memcpy_async(h_.data_ptr(), o.data_ptr(), o.numel(), 0, cout.cuda_stream))
                    j['out_ready'] = torch.cuda.Event()
                    j['out_ready'].record(cout)
                    j['host'] = host
            tick('download_enqueue', t0)

        def launch_group(grp):
            good = [j for j in grp['jobs'] if j['error'] is None]
            if good:
                try:
                    t0 = time.perf_counter()
                    scores = self._saliency_collect(grp['sal'])
                    tick('scores_wait', t0)
                    if grp['ready'] is not None:
                        pipe.wait_event(grp['ready'])
                    state = np.random.get_state() if len(good) > 1 else None
                    try:
                        launch_batches(good, scores)
                    except Exception:
                        if state is None:
                            raise
                        # a batch of the group cannot be sampled: redo the group batch by batch from the saved
                        # random state, so that the draws (and the batch that raises) match per-batch calls
                        np.random.set_state(state)
                        k = 0
                        for j in good:
                            n = len(j['dimgs'])
                            try:
                                launch_batches([j], scores[k:k + n])
                            except Exception as e:
                                j['error'] = e
                                for j2 in good[good.index(j) + 1:]:
                                    j2['error'] = e
                                break
                            k += n
                except Exception as e:
                    for j in good:
                        if j['error'] is None and 'done' not in j:
                            j['error'] = e
            for j in grp['jobs']:
                launched.append(j)

        stage_ahead = max(1, int(self.stage_ahead))

        def top_up():
            while len(staged) < stage_ahead and not exhausted[0]:
                stage_group()

        def launch_next():
            launch_group(staged.popleft())         # group k + 1: sampling + kernel chain + downloads
            top_up()                               # the groups behind it: uploads + saliency

        top_up()
        if staged:
            launch_next()
        while launched or staged:
            # keep the GPU fed: up to a full group's batches launched beyond the batch being handed over (handing the
            # first batches over earlier during the ramp-up of a loop was tried: the first step came 1-2 ms sooner and
            # the steady state lost more than that)
            if staged and len(launched) <= gmax:
                launch_next()
                continue
            job = launched.popleft()               # the next batch of group k
            if job['error'] is not None:
                raise job['error']
            t0 = time.perf_counter()
            if job['device']:
                views = job['views']
            else:
                job['out_ready'].synchronize()
                views = [a for _, a in job['host']]
            _lib.check(lib.oadg_oamix_poll_fault(0))   # a launch that left views incomplete raises here
            t0 = tick('views_wait', t0)
            res = self._fill_results(job['results'], views, job['plan'])
            t0 = tick('fill_results', t0)
            yield res, job
            tick('consumer', t0)

    def __call__(self, results, *args, **kwargs):
        """oa_mix.py:187-204."""
        results['custom_field'] = []
        for i in range(1, self.num_views + 1):
            if i == 1:
                self._history = {}
                if not self.keep_orig:
                    results['img'] = self.oamix(results['img'], results['gt_bboxes'].copy())   # the source is never mutated
                results['img_fields'] = ['img']
            else:
                results[f'img{i}'] = self.oamix(results['img'], results['gt_bboxes'].copy())
                results['img_fields'] += [f'img{i}']
                results[f'gt_bboxes{i}'] = results['gt_bboxes'].copy()
                results['oamix_boxes'] = np.stack(self._history['oa_random_box_list'], axis=0)
                results['custom_field'] += [f'img{i}', f'gt_bboxes{i}', 'oamix_boxes']
                results['multilevel_boxes'] = self._history['random_box_list']
                results['custom_field'] += ['multilevel_boxes']
        return results
