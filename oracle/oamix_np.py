"""CPU restatement of the reference OA-Mix transform.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows reference ``mmdet/datasets/pipelines/oa_mix.py:32-313`` with the op
library of ``augmix.py:32-212`` and ``bbox_augmentation.py:31-118,240-302``.
Pinned: ``tests/test_oracle_golden.py`` checks this file bit-exactly against the
unmodified reference (run under ``oracle.ref_loader``) and against the committed
goldens in ``tests/golden/`` that were produced by the reference itself
(``scripts/make_golden.py``).  The one un-pinned ingredient is
``cv2.saliency`` -> ``oracle.saliency_np`` (opencv-contrib absent; PARITY
UNPINNED for the saliency score).

Shape of this restatement (deliberately not the reference's): the transform is
written as *draw a plan, then execute it*.  The third-party arithmetic is still
the installed cv2 / Pillow / NumPy (the reference's own dependencies, which are
un-vendored: opencv-python 4.13.0, Pillow 12.2.0, NumPy 2.3.5 here), so the
pixel results are those of the reference; the multi-level composite is done by
rectangle selection instead of full-frame 0/1 float masks (exactly equal since
the masks are 0/1 and disjoint, oa_mix.py:152-154,228-234).

RNG: every draw goes through the legacy global ``np.random`` in the reference
order; the plan records each drawn parameter so tests can
compare the product's plan sampler draw for draw.
"""
import numpy as np
import cv2
from PIL import Image, ImageOps, ImageEnhance

from . import saliency_np

AUG_LIST = {
    'augmix': ['autocontrast', 'equalize', 'posterize', 'solarize',
               'bboxes_only_rotate', 'bboxes_only_shear_xy', 'bboxes_only_translate_xy',
               'bg_only_rotate', 'bg_only_shear_xy', 'bg_only_translate_xy'],
    'augmix.all': ['autocontrast', 'equalize', 'posterize', 'solarize', 'invert',
                   'color', 'contrast', 'brightness', 'sharpness',
                   'bboxes_only_rotate', 'bboxes_only_shear_xy', 'bboxes_only_translate_xy',
                   'bg_only_rotate', 'bg_only_shear_xy', 'bg_only_translate_xy'],
}  # oa_mix.py:15-29

SCORE_THRESH = 10  # oa_mix.py:64


def iou_1xk(box, boxes):
    """bbox_overlaps(box[None], boxes) (core/evaluation/bbox_overlaps.py:5-65), f32."""
    b1 = np.asarray(box, dtype=np.float32).reshape(1, 4)
    b2 = np.asarray(boxes).astype(np.float32)
    if b2.size == 0:
        return np.zeros((1, 0), np.float32)
    b2 = b2.reshape(-1, 4)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    xs = np.maximum(b1[0, 0], b2[:, 0])
    ys = np.maximum(b1[0, 1], b2[:, 1])
    xe = np.minimum(b1[0, 2], b2[:, 2])
    ye = np.minimum(b1[0, 3], b2[:, 3])
    ov = np.maximum(xe - xs, 0) * np.maximum(ye - ys, 0)
    union = np.maximum(a1[0] + a2 - ov, np.float32(1e-6))
    return (ov / union).astype(np.float32).reshape(1, -1)


# --------------------------------------------------------------------------
# masks (oa_mix.py:75-93)
# --------------------------------------------------------------------------
def hard_mask(box, shape):
    x1, y1, x2, y2 = box
    m = np.zeros(shape, np.float32)
    m[y1:y2, x1:x2, :] = 1.0
    return m


def blurred_mask(box, shape, spatial_ratio=4, sigma_ratio=0.3):
    h, w, c = shape
    x1, y1, x2, y2 = np.array(box // spatial_ratio, dtype=np.int32)
    m = np.zeros((h // spatial_ratio, w // spatial_ratio, c), np.float32)
    m[y1:y2, x1:x2, :] = 1.0
    sx = (x2 - x1) * sigma_ratio / 3 * 2
    sy = (y2 - y1) * sigma_ratio / 3 * 2
    if not (sx <= 0 or sy <= 0):
        m = cv2.GaussianBlur(m, (0, 0), sigmaX=sx, sigmaY=sy)
    return cv2.resize(m, (w, h))


# --------------------------------------------------------------------------
# plan sampling
# --------------------------------------------------------------------------
def _sample_regions(h, w, scale, ratio, num, fg_boxes=None, fg_scores=None, max_iters=50, eps=1e-6):
    """oa_mix.py:122-184.  Returns (boxes[list of int64[4]], scores or None, n_attempts)."""
    target = np.random.randint(*num) if isinstance(num, tuple) else num
    boxes, scores = [], []
    attempts = 0
    for _ in range(max_iters):
        if len(boxes) >= target:
            break
        attempts += 1
        x1, y1 = np.random.randint(0, w), np.random.randint(0, h)
        area = np.random.uniform(*scale) * h * w
        r = np.random.uniform(*ratio)
        bw, bh = int(np.sqrt(area / r)), int(np.sqrt(area * r))
        if x1 + bw > w or y1 + bh > h:
            continue
        box = np.array([x1, y1, min(x1 + bw, w), min(y1 + bh, h)])
        if np.sum(iou_1xk(box, np.asarray(boxes))) > eps:
            continue
        if fg_boxes is not None:
            ious = iou_1xk(box, fg_boxes)
            s = float('inf')
            if np.sum(ious) > eps:
                for iou, fb, fs in zip(ious[0], fg_boxes, fg_scores):
                    if iou == 0.0 or fb[2] - fb[0] < 1 or fb[3] - fb[1] < 1:
                        continue
                    if fs < s:
                        s = fs
            scores.append(s)
        boxes.append(box)
    return boxes, (scores if fg_boxes is not None else None), attempts


def _sample_level_sign(kind):
    """augmix.py:61,85,110,130,151,172: uniform(0.1,10) then a sign draw."""
    level = np.random.uniform(low=0.1, high=10)
    u = np.random.uniform() if kind in ('rotate', 'shear_x', 'shear_y') else np.random.random()
    return level, bool(u > 0.5)


def _affine(kind, level, neg, size_for_level, center, img_size):
    """2x3 forward matrix exactly as augmix.py:83-188 hands it to cv2.warpAffine."""
    if kind == 'rotate':
        deg = int(level * 30 / 10)
        if neg:
            deg = -deg
        if center is None:
            center = (img_size[0] / 2, img_size[1] / 2)
        return cv2.getRotationMatrix2D(center, deg, 1.0), deg
    if kind == 'shear_x':
        l = float(level) * 0.3 / 10.
        if neg:
            l = -l
        tx = 0 if center is None else -l * center[1]
        return np.float32([[1, -l, -tx], [0, 1, 0]]), l
    if kind == 'shear_y':
        l = float(level) * 0.3 / 10.
        if neg:
            l = -l
        ty = 0 if center is None else -l * center[0]
        return np.float32([[1, 0, 0], [-l, 1, -ty]]), l
    if kind == 'translate_x':
        l = int(level * (size_for_level[0] / 3) / 10)
        if neg:
            l = -l
        return np.float32([[1, 0, -l], [0, 1, 0]]), l
    if kind == 'translate_y':
        l = int(level * (size_for_level[1] / 3) / 10)
        if neg:
            l = -l
        return np.float32([[1, 0, 0], [0, 1, -l]]), l
    raise ValueError(kind)


def _sample_op(aug_names, gt_bboxes, img_size):
    """One ``OAMix.aug`` call (oa_mix.py:264-279): op choice + its parameter draws."""
    name = aug_names[np.random.choice(len(aug_names))]
    op = dict(name=name)
    if name in ('autocontrast', 'equalize'):
        pass
    elif name == 'posterize':
        op['bits'] = 4 - int(np.random.uniform(low=0.1, high=10) * 4 / 10)
    elif name == 'solarize':
        op['thr'] = 256 - int(np.random.uniform(low=0.1, high=10) * 256 / 10)
    elif name in ('color', 'contrast', 'brightness', 'sharpness'):
        op['factor'] = float(np.random.uniform(low=0.1, high=10)) * 1.8 / 10. + 0.1
    elif name == 'invert':
        op['tx'] = 1 if np.random.random() > 0.5 else -1
        op['ty'] = 1 if np.random.random() > 0.5 else -1
    else:
        where, geo = name.split('_only_')
        if geo == 'shear_xy':
            geo = 'shear_x' if np.random.rand() < 0.5 else 'shear_y'
        elif geo == 'translate_xy':
            geo = 'translate_x' if np.random.rand() < 0.5 else 'translate_y'
        op['geo'] = geo
        if where == 'bg':
            level, neg = _sample_level_sign(geo)
            op['M'], op['param'] = _affine(geo, level, neg, img_size, None, img_size)
        else:
            op['boxes'] = []
            for k, b in enumerate(gt_bboxes):
                x1, y1, x2, y2 = int(b[0]), int(b[1]), int(b[2]), int(b[3])
                if (x2 - x1) < 1 or (y2 - y1) < 1:
                    continue  # bbox_augmentation.py:45 (no draws)
                level, neg = _sample_level_sign(geo)
                center = ((x1 + x2) / 2., (y1 + y2) / 2.)
                M, p = _affine(geo, level, neg, (x2 - x1 + 1, y2 - y1 + 1), center, img_size)
                op['boxes'].append(dict(k=k, M=M, param=p))
    return op


def fg_scores(img, gt_bboxes, spatial_ratio=4):
    """oa_mix.py:100-111 saliency score per gt box (-1 when smaller than spatial_ratio)."""
    out = []
    for b in gt_bboxes:
        x1, y1, x2, y2 = np.array(b, dtype=np.int32)
        if x2 - x1 < spatial_ratio or y2 - y1 < spatial_ratio:
            out.append(-1)
        else:
            out.append(saliency_np.saliency_score(img[y1:y2, x1:x2]))
    return out


def sample_plan(img, gt_bboxes, version='augmix', mixture_width=3, mixture_depth=-1,
                random_box_scale=(0.01, 0.1), random_box_ratio=(3, 1 / 3),
                oa_random_box_scale=(0.005, 0.1), oa_random_box_ratio=(3, 1 / 3),
                spatial_ratio=4, sigma_ratio=0.3, scores=None, **_):
    """Everything random about one ``oamix()`` call (oa_mix.py:207-262,281-298)."""
    h, w, _c = img.shape
    names = AUG_LIST[version]
    plan = dict(h=h, w=w, spatial_ratio=spatial_ratio, sigma_ratio=sigma_ratio)
    plan['ws'] = np.float32(np.random.dirichlet([1.0] * mixture_width))
    ml_boxes, _s, _n = _sample_regions(h, w, random_box_scale, random_box_ratio, (1, 3))
    plan['ml_boxes'] = np.stack(ml_boxes, axis=0)  # ValueError if none placed (oa_mix.py:217)
    plan['scores'] = list(fg_scores(img, gt_bboxes, spatial_ratio) if scores is None else scores)
    plan['branches'] = []
    for _i in range(mixture_width):
        depth = mixture_depth if mixture_depth > 0 else np.random.randint(1, 4)
        steps = []
        for _d in range(depth):
            steps.append([_sample_op(names, gt_bboxes, (w, h)) for _r in range(len(ml_boxes) + 1)])
        plan['branches'].append(steps)
    # object-aware targets (oa_mix.py:245-262)
    low = [k for k, s in enumerate(plan['scores']) if s <= SCORE_THRESH]
    oa_boxes, oa_scores, _n = _sample_regions(
        h, w, oa_random_box_scale, oa_random_box_ratio, min(max(len(low), 1), 5),
        fg_boxes=gt_bboxes, fg_scores=plan['scores'])
    plan['oa_low_fg'] = low
    plan['oa_boxes'] = oa_boxes
    plan['oa_box_scores'] = oa_scores
    # mixing coefficients (oa_mix.py:282,295-298)
    plan['m'] = np.random.beta(1.0, 1.0)
    tgt_scores = [plan['scores'][k] for k in low] + list(oa_scores)
    plan['m_oa'] = [np.float32(np.random.uniform(0.0, 0.5)) if s <= SCORE_THRESH
                    else np.float32(np.random.uniform(0.0, 1.0)) for s in tgt_scores]
    return plan


# --------------------------------------------------------------------------
# plan execution
# --------------------------------------------------------------------------
def _apply_op(op, cur, gt_bboxes, fg_masks):
    """Full-frame result of one op on u8 HWC ``cur`` (augmix.py / bbox_augmentation.py)."""
    name = op['name']
    if name in ('autocontrast', 'equalize', 'posterize', 'solarize',
                'color', 'contrast', 'brightness', 'sharpness'):
        pil = Image.fromarray(cur, 'RGB')
        if name == 'autocontrast':
            pil = ImageOps.autocontrast(pil)
        elif name == 'equalize':
            pil = ImageOps.equalize(pil)
        elif name == 'posterize':
            pil = ImageOps.posterize(pil, op['bits'])
        elif name == 'solarize':
            pil = ImageOps.solarize(pil, op['thr'])
        else:
            enh = dict(color=ImageEnhance.Color, contrast=ImageEnhance.Contrast,
                       brightness=ImageEnhance.Brightness, sharpness=ImageEnhance.Sharpness)[name]
            pil = enh(pil).enhance(op['factor'])
        return np.asarray(pil)
    if name == 'invert':  # oa_mix.py:270-276
        M = np.float32([[1, 0, op['tx']], [0, 1, op['ty']]])
        return -cv2.warpAffine(cur, M, (0, 0))
    if name.startswith('bg_only'):  # bbox_augmentation.py:240-272
        if len(fg_masks) == 0:
            mask = np.zeros_like(cur)
        else:
            mask = np.max(fg_masks, axis=0)
        dsize = (cur.shape[1], cur.shape[0]) if op['geo'] == 'rotate' else (0, 0)
        aug = cv2.warpAffine(cur, op['M'], dsize)
        amask = cv2.warpAffine(np.asarray(mask * 255, dtype=np.uint8), op['M'], dsize) / 255
        keep = np.maximum(mask, amask)
        return np.asarray(keep * cur + (1.0 - keep) * aug, dtype=np.uint8)
    # bboxes_only (bbox_augmentation.py:31-88)
    out = cur
    for b in op['boxes']:
        dsize = (cur.shape[1], cur.shape[0]) if op['geo'] == 'rotate' else (0, 0)
        aug = cv2.warpAffine(out, b['M'], dsize)
        mask = 1.0 - fg_masks[b['k']]
        out = np.asarray(out * mask + aug * (1.0 - mask), dtype=np.uint8)
    return out


def execute_plan(img, gt_bboxes, plan, stages=None):
    """Pixels of one ``oamix()`` call given its plan.  ``stages`` (a list) receives
    ('step', branch, depth, u8 image) and ('mixed', f32 image) tuples when given."""
    img = np.asarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    fg_masks = [blurred_mask(b, img.shape, plan['spatial_ratio'], plan['sigma_ratio'])
                for b in gt_bboxes]
    ml = plan['ml_boxes']
    acc = np.zeros(img.shape, np.float32)
    for i, steps in enumerate(plan['branches']):
        cur = img.copy()
        for d, ops in enumerate(steps):
            nxt = _apply_op(ops[-1], cur, gt_bboxes, fg_masks).copy()  # outside region
            for box, op in zip(ml, ops[:-1]):
                x1, y1, x2, y2 = box
                nxt[y1:y2, x1:x2] = _apply_op(op, cur, gt_bboxes, fg_masks)[y1:y2, x1:x2]
            cur = nxt
            if stages is not None:
                stages.append(('step', i, d, cur.copy()))
        acc += plan['ws'][i] * np.asarray(cur, dtype=np.float32)
    if stages is not None:
        stages.append(('mixed', acc.copy()))
    # object-aware mixing (oa_mix.py:281-309), literal op order / dtypes
    masks = [fg_masks[k] for k in plan['oa_low_fg']] + [hard_mask(b, img.shape) for b in plan['oa_boxes']]
    m = plan['m']
    orig = np.zeros(img.shape, np.float32)
    aug = np.zeros(img.shape, np.float32)
    mask_sum = np.zeros(img.shape, np.float32)
    mask_max = None
    for mask, m_oa in zip(masks, plan['m_oa']):
        mask_sum += mask
        mask_max = mask.copy() if mask_max is None else np.maximum(mask_max, mask)
        overlap = mask_sum - mask_max
        orig += (1.0 - m_oa) * img * (mask - overlap * 0.5)
        aug += m_oa * acc * (mask - overlap * 0.5)
        mask_sum = mask_max.copy()
    out = orig + aug
    out += (1.0 - m) * img * (1.0 - mask_sum)
    out += m * acc * (1.0 - mask_sum)
    out = np.clip(out, 0, 255)
    return np.asarray(out, dtype=np.uint8)


def oamix_view(img, gt_bboxes, **cfg):
    """One generated view: returns (img_u8, plan)."""
    img = np.asarray(img, dtype=np.uint8)
    plan = sample_plan(img, gt_bboxes, **cfg)
    return execute_plan(img, gt_bboxes, plan), plan


def oamix_call(results, num_views=2, keep_orig=True, **cfg):
    """``OAMix.__call__`` (oa_mix.py:187-204) on a results dict."""
    results['custom_field'] = []
    for i in range(1, num_views + 1):
        if i == 1:
            if not keep_orig:
                results['img'], _ = oamix_view(results['img'].copy(), results['gt_bboxes'].copy(), **cfg)
            results['img_fields'] = ['img']
        else:
            out, plan = oamix_view(results['img'].copy(), results['gt_bboxes'].copy(), **cfg)
            results[f'img{i}'] = out
            results['img_fields'] += [f'img{i}']
            results[f'gt_bboxes{i}'] = results['gt_bboxes'].copy()
            results['oamix_boxes'] = np.stack(plan['oa_boxes'], axis=0)
            results['custom_field'] += [f'img{i}', f'gt_bboxes{i}', 'oamix_boxes']
            results['multilevel_boxes'] = plan['ml_boxes']
            results['custom_field'] += ['multilevel_boxes']
    return results
