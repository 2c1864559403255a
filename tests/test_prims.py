"""oracle/prims_np.py (the arithmetic spec of the CUDA kernels) against the live cv2 / Pillow calls
the reference makes.  Integer restatements must be bit-exact."""
import cv2
import numpy as np
import pytest
from PIL import Image, ImageEnhance, ImageFilter, ImageOps

from oracle import oamix_np as O, prims_np as P, saliency_np as S, synth

IMG, GT = synth.make_image(3, 300, 500, 4)
PIL_IMG = Image.fromarray(IMG, 'RGB')
HIST = P.histogram_u8c3(IMG)


@pytest.mark.parametrize('M', [
    cv2.getRotationMatrix2D((250., 150.), 17, 1.0), cv2.getRotationMatrix2D((100.5, 30.5), -29, 1.0),
    np.float32([[1, -0.23, 0.23 * 150], [0, 1, 0]]), np.float32([[1, 0, 0], [0.29, 1, -0.29 * 250]]),
    np.float32([[1, 0, -37], [0, 1, 0]]), np.float32([[1, 0, 0], [0, 1, 55]]), np.float32([[1, 0, 1], [0, 1, -1]])])
def test_warp_affine_bit_exact(M):
    assert np.array_equal(cv2.warpAffine(IMG, M, (0, 0)), P.warp_affine_u8(IMG, M))
    assert np.array_equal(cv2.warpAffine(IMG[..., 0], M, (0, 0)), P.warp_affine_u8(IMG[..., 0], M))


def test_rotation_matrix_all_degrees():
    for deg in range(-30, 31):
        for c in [(250., 150.), (100.5, 30.5), (1023.5, 77.0), (1024.0, 512.0)]:
            assert np.array_equal(cv2.getRotationMatrix2D(c, deg, 1.0), P.rotation_matrix(c, deg))


def test_histogram_and_luts():
    assert np.array_equal(np.array(PIL_IMG.histogram()).reshape(3, 256), HIST)
    assert np.array_equal(np.asarray(ImageOps.autocontrast(PIL_IMG)), P.apply_lut(IMG, P.lut_autocontrast(HIST)))
    assert np.array_equal(np.asarray(ImageOps.equalize(PIL_IMG)), P.apply_lut(IMG, P.lut_equalize(HIST)))
    for bits in (1, 2, 3, 4):
        assert np.array_equal(np.asarray(ImageOps.posterize(PIL_IMG, bits)), P.apply_lut(IMG, P.lut_posterize(bits)))
    for thr in (1, 17, 128, 255, 256):
        assert np.array_equal(np.asarray(ImageOps.solarize(PIL_IMG, thr)), P.apply_lut(IMG, P.lut_solarize(thr)))
    low = (IMG // 4 + 40).astype(np.uint8)
    h2, p2 = P.histogram_u8c3(low), Image.fromarray(low, 'RGB')
    assert np.array_equal(np.asarray(ImageOps.autocontrast(p2)), P.apply_lut(low, P.lut_autocontrast(h2)))
    assert np.array_equal(np.asarray(ImageOps.equalize(p2)), P.apply_lut(low, P.lut_equalize(h2)))
    flat = np.full((8, 8, 3), 7, np.uint8)   # single-bin histogram: identity LUTs
    assert np.array_equal(P.apply_lut(flat, P.lut_equalize(P.histogram_u8c3(flat))), flat)
    assert np.array_equal(np.asarray(ImageOps.autocontrast(Image.fromarray(flat, 'RGB'))), flat)


@pytest.mark.parametrize('f', [0.118, 0.5, 1.0, 1.3, 1.9])
def test_enhance_ops(f):
    L = P.luma_u8(IMG)
    assert np.array_equal(np.asarray(ImageEnhance.Color(PIL_IMG).enhance(f)), P.blend_u8(np.stack([L] * 3, -1), IMG, f))
    mean = int(L.astype(np.float64).mean() + 0.5)
    assert np.array_equal(np.asarray(ImageEnhance.Contrast(PIL_IMG).enhance(f)), P.blend_u8(np.full_like(IMG, mean), IMG, f))
    assert np.array_equal(np.asarray(ImageEnhance.Brightness(PIL_IMG).enhance(f)), P.blend_u8(np.zeros_like(IMG), IMG, f))
    assert np.array_equal(np.asarray(ImageEnhance.Sharpness(PIL_IMG).enhance(f)), P.blend_u8(P.smooth_u8(IMG), IMG, f))
    assert np.array_equal(np.asarray(PIL_IMG.filter(ImageFilter.SMOOTH)), P.smooth_u8(IMG))


@pytest.mark.parametrize('hw', [(300, 500), (301, 503), (1024, 2048)])
def test_blurred_mask_is_outer_product_of_profiles(hw):
    h, w = hw
    for box in [np.float32([30, 40, 190, 170]), np.float32([0, 0, w, h]), np.float32([100, 100, 102, 103]),
                np.float32([10.7, 20.2, 55.9, 99.1]), np.float32([w - 60, h - 50, w, h])]:
        m = O.blurred_mask(box, (h, w, 3))
        uy, ux = P.mask_profiles(box, h, w)
        # the three channels agree to float rounding only (IPP's interleaved resize), so all are checked
        assert np.abs(m - (uy[:, None] * ux[None, :])[..., None]).max() <= 1e-6


def test_saliency_front_end_exact_and_score_spec():
    rng = np.random.RandomState(0)
    for (h, w) in [(5, 7), (64, 64), (100, 37), (300, 199), (1000, 999), (33, 64)]:
        c = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        g = cv2.cvtColor(c, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(g, P.gray_bgr(c))
        assert np.array_equal(cv2.resize(g, (64, 64), interpolation=cv2.INTER_LINEAR_EXACT), P.resize_linear_exact_u8(g))
    for s in range(2):
        img, gt = synth.make_image(s)
        for b in gt:
            x1, y1, x2, y2 = b.astype(np.int32)
            crop = img[y1:y2, x1:x2]
            # the kernel's arithmetic (own FFT, emulated cv2 f32 cartToPolar) vs the cv2-primitive oracle
            assert abs(S.saliency_score(crop) - P.saliency_score_emul(crop)) < 1e-3
    flat = np.full((20, 30, 3), 77, np.uint8)
    assert S.saliency_score(flat) == P.saliency_score_emul(flat) == 0.0


def test_resize_linear_f32_restatement_is_bit_exact():
    """cv2.resize(float32, INTER_LINEAR) = horizontal pass then vertical pass, each a + t * (b - a) with one fused
    multiply-add (prims_np._resize_linear_f32): what the profile tile's up-sampling follows."""
    rng = np.random.RandomState(2)
    for _ in range(12):
        n, m = int(rng.randint(5, 70)), int(rng.randint(5, 50))
        a = rng.rand(m, n, 3).astype(np.float32)
        if _ % 3 == 0:
            a = (1.0 - a * 1e-5).astype(np.float32)              # near-saturated values
        got = cv2.resize(a, (4 * n, 4 * m))
        mine = P._resize_linear_f32(P._resize_linear_f32(a, 4 * n, axis=1), 4 * m, axis=0)
        assert np.array_equal(got, mine)


def test_saturated_blur_value_is_bit_exact():
    """The value of a fully covered Gaussian window in OpenCV's float32 arithmetic (prims_np.saturated_blur_value), incl.
    the 3- and 5-tap kernels that take another code path: what the kernels put where a blurred box mask saturates."""
    rng = np.random.RandomState(3)
    ones = np.ones((16, 16, 3), np.float32)
    for i in range(160):
        sx = float(rng.uniform(0.2, 0.75)) if i % 2 else float(rng.uniform(0.2, 40))
        sy = float(rng.uniform(0.2, 0.75)) if i % 4 < 2 else float(rng.uniform(0.2, 40))
        got = np.unique(cv2.GaussianBlur(ones, (0, 0), sigmaX=sx, sigmaY=sy))
        assert len(got) == 1 and float(got[0]) == float(P.saturated_blur_value(sx, sy)), (sx, sy)
