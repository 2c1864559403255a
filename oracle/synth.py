"""Deterministic synthetic inputs shared by tests, smoke and bench.

ORACLE / TEST INFRASTRUCTURE ONLY.  (bench.py carries its own copy of the
generator so that the product arm does not import ``oracle``.)
"""
import numpy as np
import cv2


def make_image(seed, h=1024, w=2048, n_gt=8):
    """Natural-image-like u8 HWC "BGR" frame + n_gt float32 gt boxes."""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, (max(h // 32, 2), max(w // 32, 2), 3)).astype(np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    img = np.clip(img.astype(np.int16) + rng.randint(-12, 13, (h, w, 3)), 0, 255).astype(np.uint8)
    if n_gt == 0:
        return img, np.zeros((0, 4), np.float32)
    bw = rng.randint(max(w * 32 // 2048, 2), max(w * 400 // 2048, 4), n_gt)
    bh = rng.randint(max(h * 32 // 1024, 2), max(h * 300 // 1024, 4), n_gt)
    x1 = rng.randint(0, w - bw)
    y1 = rng.randint(0, h - bh)
    gt = np.stack([x1, y1, x1 + bw, y1 + bh], axis=1).astype(np.float32)
    return img, gt


def make_roi_set(n=2088, c=256, seed=0, n_fg=200, n_cls=8):
    """Two-view RoI embeddings + labels in the order ContrastiveRoIHead emits
    (reference contrastive_roi_head.py:96-97): [v1 img0|img1, v2 img0|img1, rp...].
    Labels cover only the first 2048 rows, shape [2048,1] int64 (the plugin pads)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, generator=g, dtype=torch.float32)
    base = torch.full((1024,), n_cls, dtype=torch.int64)
    idx = torch.randperm(1024, generator=g)[:n_fg]
    base[idx] = torch.randint(0, n_cls, (n_fg,), generator=g)
    labels = torch.cat([base, base]).view(-1, 1)
    return x, labels
