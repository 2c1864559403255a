"""Arithmetic restatements of the third-party primitives on the OA-Mix path.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference delegates its pixel arithmetic to un-vendored wheels
(opencv-python, Pillow; un-pinned in the reference, 4.13.0 / 12.2.0 in this
image).  These functions restate the *published algorithms* of exactly the calls
the reference makes, in plain numpy integer / float64 arithmetic, so that the
CUDA kernels have a line-by-line spec.  ``tests/test_prims.py`` checks each one
against the live library call (bit-exact for the integer ones).

  warp_affine_u8      cv2.warpAffine(8U, INTER_LINEAR, BORDER_CONSTANT 0)
                      call sites augmix.py:92,116,136,156,177; oa_mix.py:276
                      (OpenCV imgwarp.cpp WarpAffineInvoker + remapBilinear fixed point)
  rotation_matrix     cv2.getRotationMatrix2D            (augmix.py:91)
  lut_*               PIL.ImageOps.{autocontrast,equalize,posterize,solarize}
                      (augmix.py:64-75,103-105)
  enhance_*           PIL.ImageEnhance.{Color,Contrast,Brightness,Sharpness}
                      (augmix.py:192-212)
  mask_profiles       OAMix._get_mask blurred branch (oa_mix.py:78-91) as the
                      outer product of two 1-D profiles (float, <=1e-6 abs)
  gray_bgr / resize_linear_exact_u8   saliency front end (cvtColor BGR2GRAY, resize
                      INTER_LINEAR_EXACT) of StaticSaliencySpectralResidual
"""
import math
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
REMAP_COEF_BITS = 15


# ----------------------------------------------------------------------------
# affine
# ----------------------------------------------------------------------------
def rotation_matrix(center, angle_deg, scale=1.0):
    """cv2.getRotationMatrix2D: center is a Point2f (float32), math in double."""
    cx = float(np.float32(center[0]))
    cy = float(np.float32(center[1]))
    a = angle_deg * (math.pi / 180.0)
    alpha = math.cos(a) * scale
    beta = math.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def invert_affine(M):
    """Forward 2x3 -> inverse (dst->src) 2x3 in double, as cv::warpAffine does."""
    m = [float(v) for v in np.asarray(M, dtype=np.float64).reshape(6)]
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11 = m[4] * D
    A22 = m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2] = b1
    m[5] = b2
    return np.array(m, dtype=np.float64)


def warp_coords(Minv, h, w):
    """Fixed-point source coordinates of every destination pixel.
    Returns (sx, sy, fx, fy): integer tap origin and 5-bit fractions."""
    m = np.asarray(Minv, dtype=np.float64).reshape(6)
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    adelta = np.rint(m[0] * x * AB_SCALE).astype(np.int64)
    bdelta = np.rint(m[3] * x * AB_SCALE).astype(np.int64)
    round_delta = AB_SCALE // INTER_TAB_SIZE // 2
    X0 = np.rint((m[1] * y + m[2]) * AB_SCALE).astype(np.int64) + round_delta
    Y0 = np.rint((m[4] * y + m[5]) * AB_SCALE).astype(np.int64) + round_delta
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    return sx, sy, X & (INTER_TAB_SIZE - 1), Y & (INTER_TAB_SIZE - 1)


def warp_affine_u8(img, M, inverse_given=False):
    """cv2.warpAffine(img, M, (0,0)) for u8 HWC / HW input (bilinear, constant 0)."""
    img = np.asarray(img)
    h, w = img.shape[:2]
    src = img.reshape(h, w, -1).astype(np.int64)
    Minv = np.asarray(M, np.float64).reshape(6) if inverse_given else invert_affine(M)
    sx, sy, fx, fy = warp_coords(Minv, h, w)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return v * ok[..., None]

    w00 = ((INTER_TAB_SIZE - fx) * (INTER_TAB_SIZE - fy) * 32)[..., None]
    w01 = (fx * (INTER_TAB_SIZE - fy) * 32)[..., None]
    w10 = ((INTER_TAB_SIZE - fx) * fy * 32)[..., None]
    w11 = (fx * fy * 32)[..., None]
    acc = (tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 +
           tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11)
    out = (acc + (1 << (REMAP_COEF_BITS - 1))) >> REMAP_COEF_BITS
    return out.astype(np.uint8).reshape(img.shape)


# ----------------------------------------------------------------------------
# Pillow LUT ops (integer)
# ----------------------------------------------------------------------------
def histogram_u8c3(img):
    """PIL Image.histogram() for an RGB image: [3][256] counts."""
    return np.stack([np.bincount(img[..., c].ravel(), minlength=256) for c in range(3)])


def lut_autocontrast(hist):
    lut = np.zeros((3, 256), np.uint8)
    for c in range(3):
        nz = np.nonzero(hist[c])[0]
        lo, hi = (int(nz[0]), int(nz[-1])) if len(nz) else (255, 0)
        if hi <= lo:
            lut[c] = np.arange(256)
        else:
            scale = 255.0 / (hi - lo)
            offset = -lo * scale
            for ix in range(256):
                v = int(ix * scale + offset)
                lut[c, ix] = min(max(v, 0), 255)
    return lut


def lut_equalize(hist):
    lut = np.zeros((3, 256), np.uint8)
    for c in range(3):
        h = [int(v) for v in hist[c]]
        nz = [v for v in h if v]
        if len(nz) <= 1:
            lut[c] = np.arange(256)
            continue
        step = (sum(nz) - nz[-1]) // 255
        if not step:
            lut[c] = np.arange(256)
            continue
        n = step // 2
        for i in range(256):
            lut[c, i] = min(n // step, 255)  # Image.point clips list entries to u8
            n += h[i]
    return lut


def lut_posterize(bits):
    mask = ~(2 ** (8 - bits) - 1)
    return np.tile(np.array([i & mask for i in range(256)], np.uint8), (3, 1))


def lut_solarize(thr):
    return np.tile(np.array([i if i < thr else 255 - i for i in range(256)], np.uint8), (3, 1))


def apply_lut(img, lut):
    out = np.empty_like(img)
    for c in range(3):
        out[..., c] = lut[c][img[..., c]]
    return out


# ----------------------------------------------------------------------------
# Pillow enhance ops
# ----------------------------------------------------------------------------
def luma_u8(img):
    """PIL convert('L') on an "RGB" buffer: (19595 R + 38470 G + 7471 B + 0x8000) >> 16,
    where R,G,B are channels 0,1,2 of the buffer handed to Image.fromarray."""
    v = img.astype(np.int64)
    return ((19595 * v[..., 0] + 38470 * v[..., 1] + 7471 * v[..., 2] + 0x8000) >> 16).astype(np.uint8)


def smooth_u8(img):
    """PIL ImageFilter.SMOOTH: 3x3 kernel (1,1,1,1,5,1,1,1,1)/13, offset 0, borders copied."""
    v = img.astype(np.float32)
    h, w = img.shape[:2]
    out = img.copy()
    if h < 3 or w < 3:
        return out
    k = [[1, 1, 1], [1, 5, 1], [1, 1, 1]]
    acc = np.zeros((h - 2, w - 2, img.shape[2]), np.float32)
    # Pillow ImagingFilter3x3 (8-bit multi-band path): float sum of the 9 taps * kernel/div, + 0.5
    for dy in range(3):
        for dx in range(3):
            acc += v[dy:h - 2 + dy, dx:w - 2 + dx] * np.float32(k[dy][dx] / 13.0)
    res = np.clip(np.floor(acc + np.float32(0.5)), 0, 255)
    out[1:-1, 1:-1] = res.astype(np.uint8)
    return out


def blend_u8(deg, img, factor):
    """PIL Image.blend(deg, img, factor): float32 arithmetic, trunc (clip when factor outside [0,1])."""
    a = np.float32(factor)
    d = deg.astype(np.float32)
    t = (d + a * (img.astype(np.float32) - d)).astype(np.float32)
    if 0.0 <= factor <= 1.0:
        return t.astype(np.uint8)
    return np.clip(t, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# blurred box mask as an outer product of two 1-D profiles
# ----------------------------------------------------------------------------
def gaussian_kernel_f32(sigma):
    """cv::getGaussianKernel(ksize auto for 32F, sigma, CV_32F)."""
    n = int(round(sigma * 4 * 2 + 1)) | 1  # cvRound(sigma*8+1)|1
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    k = k * (1.0 / k.sum())
    return k.astype(np.float32)


def _profile_1d(lo, hi, n_lo, n_hi, sigma, blur):
    """indicator[lo:hi] on n_lo samples -> (blur) -> bilinear upsample to n_hi (cv2.resize f32)."""
    ind = np.zeros(n_lo, np.float32)
    ind[lo:hi] = 1.0  # python slice semantics, like the reference
    p = ind
    if blur:
        k = gaussian_kernel_f32(sigma)
        r = len(k) // 2
        # BORDER_REFLECT_101; cv2 limits reflection by clamping for very large kernels
        idx = np.arange(-r, n_lo + r)
        if n_lo == 1:
            idx = np.zeros_like(idx)
        else:
            period = 2 * (n_lo - 1)
            idx = np.abs(idx) % period
            idx = np.where(idx >= n_lo, period - idx, idx)
        padded = ind[idx].astype(np.float64)
        p = np.convolve(padded, k.astype(np.float64)[::-1], mode='valid').astype(np.float32)
    scale = n_lo / n_hi
    d = np.arange(n_hi)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    t = (f - s).astype(np.float32)
    t = np.where(s < 0, np.float32(0), t)
    s = np.where(s < 0, 0, s)
    t = np.where(s >= n_lo - 1, np.float32(0), t)
    s = np.where(s >= n_lo - 1, n_lo - 1, s)
    s1 = np.minimum(s + 1, n_lo - 1)
    return (p[s] * (np.float32(1) - t) + p[s1] * t).astype(np.float32)


def _fma32(a, b, c):
    """float32 fused multiply-add on arrays: the product of two float32 is exact in float64."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _reflect101(idx, n):
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    idx = np.abs(idx) % period
    return np.where(idx >= n, period - idx, idx)


def _resize_linear_f32(a, n_hi, axis):
    """One pass of cv2.resize(float32, INTER_LINEAR) along ``axis``: a + t * (b - a) with ONE fused multiply-add."""
    a = np.moveaxis(a, axis, 0)
    n_lo = a.shape[0]
    d = np.arange(n_hi)
    f = ((d + 0.5) * (n_lo / n_hi) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    t = (f - s).astype(np.float32)
    t = np.where(s < 0, np.float32(0), t)
    s = np.where(s < 0, 0, s)
    t = np.where(s >= n_lo - 1, np.float32(0), t)
    s = np.where(s >= n_lo - 1, n_lo - 1, s)
    s1 = np.minimum(s + 1, n_lo - 1)
    tt = t.reshape((-1,) + (1,) * (a.ndim - 1))
    out = _fma32(np.broadcast_to(tt, a[s].shape), (a[s1] - a[s]).astype(np.float32), a[s])
    return np.moveaxis(out, 0, axis)


def saturated_blur_value(sigma_x, sigma_y):
    """cv2.GaussianBlur(ones, (0, 0), sigma_x, sigma_y) as OpenCV 4.13 computes it in float32 -- the value a blurred box
    mask takes where its window is fully covered (1 - 2^-24, 1 or 1 + 2^-23): the row filter adds the taps in order
    (kernels of 3 and 5 taps: the symmetric form), the symmetric column filter starts at the centre tap and adds
    k[c + j] * (v + v) outward with fused multiply-adds.  BIT-EXACT against cv2 (tests/test_prims.py)."""
    def col(k, v):
        c = len(k) // 2
        s = np.float32(k[c] * v)
        v2 = np.float32(v + v)
        for j in range(1, c + 1):
            s = np.float32(np.float64(k[c + j]) * np.float64(v2) + np.float64(s))
        return s
    kx, ky = gaussian_kernel_f32(sigma_x), gaussian_kernel_f32(sigma_y)
    if len(kx) <= 5:
        rx = col(kx, np.float32(1))
    else:
        rx = np.float32(kx[0])
        for v in kx[1:]:
            rx = np.float32(rx + v)
    return col(ky, rx)


def mask_profiles(box, h, w, spatial_ratio=4, sigma_ratio=0.3):
    """(uy[h], ux[w]) with blurred_mask(box)[y,x,:] ~= uy[y]*ux[x]  (<=1e-6 abs)."""
    x1, y1, x2, y2 = [int(v) for v in np.array(np.asarray(box, np.float32) // spatial_ratio, dtype=np.int32)]
    h4, w4 = h // spatial_ratio, w // spatial_ratio
    sx = (x2 - x1) * sigma_ratio / 3 * 2
    sy = (y2 - y1) * sigma_ratio / 3 * 2
    blur = not (sx <= 0 or sy <= 0)
    ux = _profile_1d(x1, x2, w4, w, sx, blur)
    uy = _profile_1d(y1, y2, h4, h, sy, blur)
    return uy, ux


# ----------------------------------------------------------------------------
# saliency front end (integer, exact)
# ----------------------------------------------------------------------------
def gray_bgr(img):
    """cv2.cvtColor(BGR2GRAY) 8U: (B*3735 + G*19235 + R*9798 + 16384) >> 15."""
    v = img.astype(np.int64)
    return ((v[..., 0] * 3735 + v[..., 1] * 19235 + v[..., 2] * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def _exact_axis(n_in, n_out):
    """INTER_LINEAR_EXACT taps: index pair + weight in 1/(2*n_out) units (exact rationals)."""
    # f = (d + 0.5) * n_in / n_out - 0.5 = ((2d+1) n_in - n_out) / (2 n_out)
    d = np.arange(n_out, dtype=np.int64)
    num = (2 * d + 1) * n_in - n_out
    den = 2 * n_out
    i0 = np.floor_divide(num, den)
    t = num - i0 * den  # weight of tap i0+1, out of den
    lo = i0 < 0
    hi = i0 >= n_in - 1
    t = np.where(lo | hi, 0, t)
    i0 = np.where(lo, 0, np.where(hi, n_in - 1, i0))
    i1 = np.minimum(i0 + 1, n_in - 1)
    return i0, i1, t, den


def resize_linear_exact_u8(gray, n_out_h=64, n_out_w=64):
    """cv2.resize(gray u8, (w,h), INTER_LINEAR_EXACT): 8.8 fixed-point weights,
    horizontal pass then vertical pass, round half up."""
    g = gray.astype(np.int64)
    h, w = g.shape
    x0, x1, tx, denx = _exact_axis(w, n_out_w)
    y0, y1, ty, deny = _exact_axis(h, n_out_h)
    # weights in 1/256 (exact when den divides 256*k; rounded like ufixedpoint16 otherwise)
    ax = np.floor(tx * 256 / denx + 0.5).astype(np.int64)
    ay = np.floor(ty * 256 / deny + 0.5).astype(np.int64)
    rows = g[:, x0] * (256 - ax)[None, :] + g[:, x1] * ax[None, :]            # 8.8
    out = rows[y0, :] * (256 - ay)[:, None] + rows[y1, :] * ay[:, None]        # 8.16
    return ((out + (1 << 15)) >> 16).astype(np.uint8)


# ----------------------------------------------------------------------------
# saliency: arithmetic spec of the CUDA kernel (own FFT; cv2's f32 cartToPolar emulated)
# ----------------------------------------------------------------------------
_F32 = np.float32
_ATAN_P = [_F32(0.9997878412794807) * _F32(180 / np.pi), _F32(-0.3258083974640975) * _F32(180 / np.pi),
           _F32(0.1555786518463281) * _F32(180 / np.pi), _F32(-0.04432655554792128) * _F32(180 / np.pi)]


def _fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(_F32)


def cv_magnitude_f32(re, im):
    """cv2.cartToPolar magnitude for 64F input on an FMA host: sqrtf(fmaf(x,x,y*y)) in f32."""
    x = np.asarray(re).astype(_F32)
    y = np.asarray(im).astype(_F32)
    return np.sqrt(_fma32(x, x, (y * y).astype(_F32))).astype(np.float64)


def cv_fast_atan_f32(im, re):
    """cv2.cartToPolar angle (radians) for 64F input: f32 polynomial fastAtan with FMAs."""
    x = np.asarray(re).astype(_F32)
    y = np.asarray(im).astype(_F32)
    ax, ay = np.abs(x), np.abs(y)
    eps = _F32(2.220446049250313e-16)
    big = ax >= ay
    with np.errstate(all='ignore'):
        c = np.where(big, ay / (ax + eps), ax / (ay + eps)).astype(_F32)
    c2 = (c * c).astype(_F32)
    p1, p3, p5, p7 = _ATAN_P
    a = _fma32(_fma32(_fma32(np.full_like(c, p7), c2, np.full_like(c, p5)), c2, np.full_like(c, p3)),
               c2, np.full_like(c, p1))
    a = (a * c).astype(_F32)
    a = np.where(big, a, _F32(90) - a)
    a = np.where(x < 0, _F32(180) - a, a)
    a = np.where(y < 0, _F32(360) - a, a)
    return (a * _F32(np.pi / 180)).astype(_F32).astype(np.float64)


def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def resize_linear_f32(m, w, h):
    """cv2.resize(f32, INTER_LINEAR) generic path (IPP off): f32 weights, horizontal then vertical."""
    def axis(n_in, n_out):
        d = np.arange(n_out)
        f = ((d + 0.5) * (n_in / n_out) - 0.5).astype(_F32)
        s = np.floor(f).astype(np.int64)
        t = (f - s).astype(_F32)
        t = np.where(s < 0, _F32(0), t)
        s = np.where(s < 0, 0, s)
        t = np.where(s >= n_in - 1, _F32(0), t)
        s = np.where(s >= n_in - 1, n_in - 1, s)
        return s, np.minimum(s + 1, n_in - 1), t
    sx, sx1, tx = axis(m.shape[1], w)
    sy, sy1, ty = axis(m.shape[0], h)
    rows = (m[:, sx] * (_F32(1) - tx) + m[:, sx1] * tx).astype(_F32)
    return (rows[sy] * (_F32(1) - ty)[:, None] + rows[sy1] * ty[:, None]).astype(_F32)


def saliency_map64(crop):
    """64x64 f32 saliency map (before the final resize) with the kernel's arithmetic."""
    g = resize_linear_exact_u8(gray_bgr(crop) if crop.ndim == 3 else crop).astype(np.float64)
    F = np.fft.fft2(g)
    re, im = F.real, F.imag
    mag = cv_magnitude_f32(re, im)
    ang = cv_fast_atan_f32(im, re)
    with np.errstate(all='ignore'):
        la = np.log(mag)
        idx = _reflect101(np.arange(-1, 65), 64)
        pad = la[idx][:, idx]
        bl = sum(pad[dy:dy + 64, dx:dx + 64] for dy in range(3) for dx in range(3)) * (1.0 / 9)
        nm = np.exp(la - bl)
        G = np.fft.ifft2(nm * np.cos(ang) + 1j * nm * np.sin(ang)) * 4096
    m = cv_magnitude_f32(G.real, G.imag)
    x = np.arange(5) - 2.0
    k = np.exp(-0.5 / 64.0 * x * x)
    k = k / k.sum()
    idx = _reflect101(np.arange(-2, 66), 64)
    pad = m[:, idx]
    m = sum(pad[:, d:d + 64] * k[d] for d in range(5))
    pad = m[idx, :]
    m = sum(pad[d:d + 64, :] * k[d] for d in range(5))
    m = m * m
    with np.errstate(all='ignore'):
        m = m / m.max()
    return m.astype(_F32)


def saliency_score_emul(crop):
    m = saliency_map64(crop)
    h, w = crop.shape[:2]
    out = resize_linear_f32(m, w, h)
    with np.errstate(all='ignore'):
        v = np.where(np.isnan(out), 0, out * _F32(255)).astype(np.uint8)
    return float(v.astype(np.int64).sum() / (w * h))


# ---- Normalize -> Pad -> DefaultFormatBundle on one frame (the fused epilogue's oracle) ---------------------------
def imnormalize_pad_chw(img_u8, mean, std, to_rgb=True, size_divisor=32):
    """What the training pipeline makes of a uint8 HWC frame after OAMix:
    ``Normalize`` (mmdet/datasets/pipelines/transforms.py:672-704) calls ``mmcv.imnormalize`` -- mmcv is third party
    and absent from this image (PARITY UNPINNED against its binary); its published source (mmcv 1.x
    ``mmcv/image/photometric.py::imnormalize_``) is restated here with the cv2 calls it makes: float32 copy,
    ``cv2.cvtColor(BGR2RGB)`` in place, ``cv2.subtract(img, float64 mean)``, ``cv2.multiply(img, 1 / float64 std)``;
    ``Pad(size_divisor)`` (transforms.py:573-640 -> ``mmcv.impad_to_multiple``, zeros right / below) and the HWC ->
    CHW transpose of ``DefaultFormatBundle`` (formating.py:217-234).  Returns float32 [3, Hp, Wp]."""
    import cv2
    img = np.ascontiguousarray(img_u8).astype(np.float32)
    mean64 = np.float64(np.asarray(mean, np.float32).reshape(1, -1))
    stdinv = 1 / np.float64(np.asarray(std, np.float32).reshape(1, -1))
    if to_rgb:
        cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
    cv2.subtract(img, mean64, img)
    cv2.multiply(img, stdinv, img)
    h, w = img.shape[:2]
    hp, wp = -(-h // size_divisor) * size_divisor, -(-w // size_divisor) * size_divisor
    out = np.zeros((hp, wp, 3), np.float32)
    out[:h, :w] = img
    return np.ascontiguousarray(out.transpose(2, 0, 1))
