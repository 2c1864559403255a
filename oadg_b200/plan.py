"""Host-side OA-Mix plan: numpy mirrors of the C records in include/oadg.h and the packer.

A *plan* is everything random about one batch of ``OAMix.oamix`` calls (reference
oa_mix.py:207-262,281-298); the device executes it (csrc/oamix.cu).
"""
import numpy as np

MAX_WIDTH, MAX_DEPTH, MAX_REGIONS = 8, 8, 3
MAGIC = 0x4F414447

OP = dict(autocontrast=0, equalize=1, posterize=2, solarize=3, invert=4, color=5, contrast=6,
          brightness=7, sharpness=8, bg_affine=9, bbo_affine=10)

HEADER_DT = np.dtype([('magic', 'i4'), ('abi', 'i4'), ('n_views', 'i4'), ('n_gt', 'i4'), ('n_ops', 'i4'),
                      ('n_bbo', 'i4'), ('n_tgt', 'i4'), ('max_h', 'i4'), ('max_w', 'i4'),
                      ('off_views', 'i4'), ('off_gt', 'i4'), ('off_ops', 'i4'), ('off_bbo', 'i4'),
                      ('off_tgt', 'i4'), ('total_bytes', 'i4'), ('pad', 'i4')], align=True)
VIEW_DT = np.dtype([('H', 'i4'), ('W', 'i4'), ('img', 'i4'), ('n_gt', 'i4'), ('gt_first', 'i4'), ('n_ml', 'i4'),
                    ('ml_box', 'i4', (2, 4)), ('width', 'i4'), ('depth', 'i4', (MAX_WIDTH,)),
                    ('ws', 'f4', (MAX_WIDTH,)), ('op_first', 'i4'), ('n_tgt', 'i4'), ('tgt_first', 'i4'),
                    ('pad', 'i4'), ('m', 'f8')], align=True)
GT_DT = np.dtype([('lo', 'i4', (4,)), ('blur', 'i4'), ('kx', 'i4'), ('ky', 'i4'), ('view', 'i4'),
                  ('sigma_x', 'f8'), ('sigma_y', 'f8'), ('supp', 'i4', (4,))], align=True)
OP_DT = np.dtype([('kind', 'i4'), ('p0', 'i4'), ('p1', 'i4'), ('factor', 'f4'), ('minv', 'f8', (6,)),
                  ('bbo_first', 'i4'), ('bbo_count', 'i4'), ('lut', 'i4'), ('scratch', 'i4')], align=True)
BBO_DT = np.dtype([('gt', 'i4'), ('pad', 'i4'), ('minv', 'f8', (6,))], align=True)
TGT_DT = np.dtype([('kind', 'i4'), ('gt', 'i4'), ('box', 'i4', (4,)), ('m_oa', 'f4'), ('pad', 'i4')], align=True)

STRUCT_SIZES = [d.itemsize for d in (HEADER_DT, VIEW_DT, GT_DT, OP_DT, BBO_DT, TGT_DT)]
OPS_PER_VIEW = MAX_WIDTH * MAX_DEPTH * MAX_REGIONS


def _a8(n):
    return (n + 7) // 8 * 8


def pack(views, gts, ops, bbos, tgts, max_h, max_w):
    """Concatenate the record arrays into the blob oadg_oamix_execute consumes."""
    hdr = np.zeros(1, HEADER_DT)
    off = _a8(HEADER_DT.itemsize)
    parts = []
    for name, arr in (('views', views), ('gt', gts), ('ops', ops), ('bbo', bbos), ('tgt', tgts)):
        hdr['off_' + name] = off
        parts.append((off, arr))
        off = _a8(off + arr.nbytes)
    hdr['magic'], hdr['abi'] = MAGIC, 1
    hdr['n_views'], hdr['n_gt'], hdr['n_ops'] = len(views), len(gts), len(ops)
    hdr['n_bbo'], hdr['n_tgt'] = len(bbos), len(tgts)
    hdr['max_h'], hdr['max_w'] = max_h, max_w
    hdr['total_bytes'] = off
    blob = np.zeros(off, np.uint8)
    blob[:HEADER_DT.itemsize] = hdr.view(np.uint8)
    for o, arr in parts:
        if arr.nbytes:
            blob[o:o + arr.nbytes] = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    return blob


# ---- fast packers (struct formats mirror the dtypes above; tests/test_host.py checks both against the C sizes)
import struct as _struct

HEADER_ST = _struct.Struct('<16i')
VIEW_ST = _struct.Struct('<6i8ii8i8f4i4xd')
GT_ST = _struct.Struct('<4i4i2d4i')
OP_ST = _struct.Struct('<3if6d4i')
BBO_ST = _struct.Struct('<2i6d')
TGT_ST = _struct.Struct('<2i4ifi')
assert [st.size for st in (HEADER_ST, VIEW_ST, GT_ST, OP_ST, BBO_ST, TGT_ST)] == STRUCT_SIZES
_ZERO_MINV = (0.0,) * 6


class BlobBuilder:
    """Packs plan records straight into one bytearray (no per-field numpy assignments)."""

    def __init__(self, n_views, n_gt, n_bbo, n_tgt):
        self.n_views, self.n_gt, self.n_bbo, self.n_tgt = n_views, n_gt, n_bbo, n_tgt
        self.n_ops = n_views * OPS_PER_VIEW
        off = _a8(HEADER_ST.size)
        self.off_views = off
        off = _a8(off + n_views * VIEW_ST.size)
        self.off_gt = off
        off = _a8(off + n_gt * GT_ST.size)
        self.off_ops = off
        off = _a8(off + self.n_ops * OP_ST.size)
        self.off_bbo = off
        off = _a8(off + n_bbo * BBO_ST.size)
        self.off_tgt = off
        off = _a8(off + n_tgt * TGT_ST.size)
        self.total = off
        self.buf = bytearray(off)

    def view(self, v, *fields):
        VIEW_ST.pack_into(self.buf, self.off_views + v * VIEW_ST.size, *fields)

    def gt(self, g, *fields):
        GT_ST.pack_into(self.buf, self.off_gt + g * GT_ST.size, *fields)

    def op(self, i, kind, p0=0, p1=0, factor=0.0, minv=_ZERO_MINV, bbo_first=0, bbo_count=0):
        OP_ST.pack_into(self.buf, self.off_ops + i * OP_ST.size, kind, p0, p1, factor, *minv, bbo_first, bbo_count, -1, -1)

    def bbo(self, i, gt, minv):
        BBO_ST.pack_into(self.buf, self.off_bbo + i * BBO_ST.size, gt, 0, *minv)

    def tgt(self, i, kind, gt, box, m_oa):
        TGT_ST.pack_into(self.buf, self.off_tgt + i * TGT_ST.size, kind, gt, *box, m_oa, 0)

    def finish(self, max_h, max_w):
        HEADER_ST.pack_into(self.buf, 0, MAGIC, 1, self.n_views, self.n_gt, self.n_ops, self.n_bbo, self.n_tgt,
                            max_h, max_w, self.off_views, self.off_gt, self.off_ops, self.off_bbo, self.off_tgt,
                            self.total, 0)
        return np.frombuffer(self.buf, dtype=np.uint8)
