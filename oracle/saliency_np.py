"""Spectral-residual static saliency, restated from cv2 core primitives.

ORACLE / TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED: the reference calls
``cv2.saliency.StaticSaliencySpectralResidual_create().computeSaliency(crop)``
(reference ``mmdet/datasets/pipelines/oa_mix.py:108-109``), which lives in
opencv-contrib-python (un-pinned in the reference, README.md:73-75) and is NOT
installed in this image (``hasattr(cv2, 'saliency') is False``), nor is its
source under /root/reference.  This file restates the published algorithm of
opencv_contrib ``modules/saliency/src/staticSaliencySpectralResidual.cpp``
(``computeSaliencyImpl``) out of the very cv2 *core* calls that function makes
(cvtColor, resize INTER_LINEAR_EXACT, dft, cartToPolar, log, blur, exp,
polarToCart, dft inverse, GaussianBlur, minMaxLoc, resize INTER_LINEAR), so that
everything except the glue is the installed OpenCV 4.13 arithmetic.
"""
import numpy as np
import cv2

RES_W = 64  # resImWidth  (contrib default)
RES_H = 64  # resImHeight (contrib default)


def compute_saliency(image):
    """Return (True, saliency_map f32 HxW) like ``computeSaliency``."""
    image = np.ascontiguousarray(image)
    if image.ndim == 3 and image.shape[2] == 3:
        gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
    else:
        gray = image
    gray_down = cv2.resize(gray, (RES_W, RES_H), interpolation=cv2.INTER_LINEAR_EXACT)
    real = gray_down.astype(np.float64)
    imag = np.zeros_like(real)
    combined = cv2.merge([real, imag])
    image_dft = cv2.dft(combined)
    re, im = cv2.split(image_dft)
    magnitude, angle = cv2.cartToPolar(re, im, angleInDegrees=False)
    with np.errstate(divide='ignore', invalid='ignore'):
        log_amplitude = cv2.log(magnitude)
        log_amplitude_blur = cv2.blur(log_amplitude, (3, 3), anchor=(-1, -1),
                                      borderType=cv2.BORDER_DEFAULT)
        magnitude = cv2.exp(log_amplitude - log_amplitude_blur)
    re, im = cv2.polarToCart(magnitude, angle, angleInDegrees=False)
    image_dft = cv2.merge([re, im])
    combined = cv2.dft(image_dft, flags=cv2.DFT_INVERSE)
    re, im = cv2.split(combined)
    magnitude, angle = cv2.cartToPolar(re, im, angleInDegrees=False)
    magnitude = cv2.GaussianBlur(magnitude, (5, 5), 8, None, 0, cv2.BORDER_DEFAULT)
    magnitude = magnitude * magnitude
    _, max_val, _, _ = cv2.minMaxLoc(magnitude)
    with np.errstate(divide='ignore', invalid='ignore'):
        magnitude = magnitude / max_val
    magnitude = magnitude.astype(np.float32)
    h, w = image.shape[:2]
    saliency_map = cv2.resize(magnitude, (w, h), interpolation=cv2.INTER_LINEAR)
    return True, saliency_map


def saliency_score(crop):
    """``np.mean((saliency_map * 255).astype('uint8'))`` (oa_mix.py:110)."""
    _, sal = compute_saliency(crop)
    with np.errstate(invalid='ignore'):
        return float(np.mean((sal * 255).astype('uint8')))


class _Shim:
    """Drop-in for the ``cv2.saliency`` namespace used at oa_mix.py:108."""

    class _Impl:
        @staticmethod
        def computeSaliency(img):
            return compute_saliency(img)

    @staticmethod
    def StaticSaliencySpectralResidual_create():
        return _Shim._Impl()


def install_cv2_shim():
    """Give the *reference* code a ``cv2.saliency`` (dev container only)."""
    if not hasattr(cv2, 'saliency'):
        cv2.saliency = _Shim
