"""Device-side observation preprocessing: uint8 HWC camera frames -> float32 [B, 3, 224, 224] in [0, 1]
(or normalised) on the GPU, the work the reference does on the host in PIL / numpy (model/processing_spatialvla.py:174 -> HF
SiglipImageProcessor at transformers 4.47: `Image.resize((w, h), BICUBIC)`, `x * rescale_factor`, optional `(x - mean) / std`,
channels first).  Pillow's resampler is integer arithmetic on uint8 (22-bit fixed-point coefficients, two rounded passes), so the
CUDA kernels (csrc/image_ops.cu) reproduce it bit for bit; this module builds Pillow's coefficient tables per input size and the
256-entry value map, and launches the kernels."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _cubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=64)
def resample_tables(in_size: int, out_size: int):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the bicubic filter (support 2, scaled by the down-sampling ratio):
    -> (bounds int32 [out, 2] = (first source index, tap count), coefficients int32 [out, ksize], ksize)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 2.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - lo
        w = [_cubic((t + lo - center + 0.5) / fscale) for t in range(n)]
        tot = 0.0
        for v in w:
            tot += v
        if tot != 0.0:
            w = [v / tot for v in w]
        for t, v in enumerate(w):
            kk[o, t] = int(v * (1 << PRECISION_BITS) - 0.5) if v < 0 else int(v * (1 << PRECISION_BITS) + 0.5)
        bounds[o] = (lo, n)
    return bounds, kk, ksize


def value_lut(rescale_factor=1 / 255, do_rescale=True, do_normalize=False, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """float32 [3, 256]: what the HF image processor turns each uint8 level into, computed with its own numpy arithmetic
    (`image.astype(float64) * scale` stored as float32; `(image - mean) / std` in float32)."""
    lv = np.arange(256, dtype=np.float64)
    x = (lv * rescale_factor).astype(np.float32) if do_rescale else lv.astype(np.float32)
    lut = np.repeat(x[None], 3, 0)
    if do_normalize:
        lut = (lut - np.asarray(mean, dtype=np.float32)[:, None]) / np.asarray(std, dtype=np.float32)[:, None]
    return np.ascontiguousarray(lut.astype(np.float32))


class DeviceImageProcessor:
    """images uint8 [B, H, W, 3] (torch tensor on any device, or numpy) -> float32 [B, 3, out_h, out_w] on `ops.device`."""

    def __init__(self, ops, size=(224, 224), rescale_factor=1 / 255, do_rescale=True, do_normalize=False, mean=(0.5, 0.5, 0.5),
                 std=(0.5, 0.5, 0.5)):
        self.ops, self.oh, self.ow = ops, int(size[0]), int(size[1])
        self.lut = torch.from_numpy(value_lut(rescale_factor, do_rescale, do_normalize, mean, std)).to(ops.device)
        self._tables = {}

    def _dev_tables(self, n_in, n_out):
        key = (n_in, n_out)
        if key not in self._tables:
            b, k, ks = resample_tables(n_in, n_out)
            self._tables[key] = (torch.from_numpy(b).to(self.ops.device), torch.from_numpy(k).to(self.ops.device), ks)
        return self._tables[key]

    def __call__(self, images):
        if isinstance(images, (list, tuple)):
            images = np.stack([np.asarray(im) for im in images])
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images))
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError("DeviceImageProcessor expects uint8 images [B, H, W, 3]")
        x = images.to(self.ops.device, non_blocking=True).contiguous()
        B, H, W, _ = x.shape
        out = self.ops.empty((B, 3, self.oh, self.ow), torch.float32)
        th = self._dev_tables(W, self.ow) if W != self.ow else None
        tv = self._dev_tables(H, self.oh) if H != self.oh else None
        tmp = self.ops.empty((B, H, self.ow, 3), torch.uint8) if th is not None else None
        self.ops.image_preprocess(x, tmp, out, th, tv, self.lut)
        return out
