"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the image half of the reference processor
(model/processing_spatialvla.py:174 -> HF SiglipImageProcessor pinned at transformers 4.47: uint8 HWC image -> PIL
`Image.resize((224, 224), resample=BICUBIC)` -> x * (1 / 255) in float64, stored float32 -> optional (x - mean) / std in float32
-> CHW).  The resampler is Pillow's `ImagingResample` (src/libImaging/Resample.c, Pillow >= 7; third-party code, absent from
/root/reference): separable two-pass convolution on uint8 with 22-bit fixed-point coefficients, horizontal pass first, every pass
rounded and clipped to uint8, filter support scaled by the down-sampling ratio (antialiasing).  Pinned against Pillow itself in
tests/test_image_preprocess.py."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (bounds int32 [out, 2] = (first tap, tap count), kk int32 [out, ksize])"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)                      # left-to-right double accumulation, as the C loop
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """one separable pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H, W, C] image"""
    src = img.astype(np.int64)
    out_n = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_n
    out = np.zeros(shape, dtype=np.uint8)
    for o in range(out_n):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        k = kk[o, :n].astype(np.int64)
        if axis == 1:
            acc = (src[:, lo:lo + n, :] * k[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, o, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        else:
            acc = (src[lo:lo + n, :, :] * k[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
            out[o, :, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_u8_ref(img: np.ndarray, out_h: int = 224, out_w: int = 224) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [out_h, out_w, C], Pillow's Image.resize(BICUBIC) bit for bit (horizontal pass, then vertical)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    if W != out_w:
        img = _pass(img, *resample_coeffs(W, out_w), axis=1)
    if H != out_h:
        img = _pass(img, *resample_coeffs(H, out_h), axis=0)
    return img


def preprocess_ref(images, out_h=224, out_w=224, rescale_factor=1 / 255, do_normalize=False, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """list of uint8 [H, W, 3] -> float32 [B, 3, out_h, out_w] exactly as SiglipImageProcessor (4.47) produces it."""
    out = []
    for im in images:
        r = resize_u8_ref(np.asarray(im), out_h, out_w)
        x = (r.astype(np.float64) * rescale_factor).astype(np.float32)
        if do_normalize:
            x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
        out.append(x.transpose(2, 0, 1))
    return np.stack(out)
