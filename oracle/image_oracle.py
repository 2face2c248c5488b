"""TEST INFRASTRUCTURE — CPU restatement of the reference's image path (mode 'resize'):
mmgpt/data/dataset/base_dataset.py:190-192 / mmgpt/utils/mm_utils.py:41-43

    image = image.resize((S, S))                                   # PIL, default resample = BICUBIC
    processor.preprocess(image, do_resize=False, do_center_crop=False)   # x / 255, (x - mean) / std

The arithmetic lives in third-party code that is not under /root/reference: Pillow (installed 12.2.0)
src/libImaging/Resample.c — `precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`,
`ImagingResampleVertical_8bpc` — and transformers `image_transforms.rescale` / `normalize`.  This file restates
them in numpy; tests/test_data_cpu.py pins the restatement against Pillow itself (bit-exact on the uint8 pixels)
and against the installed CLIPImageProcessor.  Only tests/, smoke() and bench.py's cpu_baseline may import it.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs (box = the whole axis) + normalize_coeffs_8bpc -> (bounds [out, 2], kk [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            v = k * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img: np.ndarray, bounds, kk) -> np.ndarray:
    """One resample pass along axis 1 of img [rows, in, 3] uint8 -> [rows, out, 3] uint8."""
    out = np.empty((img.shape[0], bounds.shape[0], 3), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        xmin, n = bounds[xx]
        acc = (src[:, xmin:xmin + n, :] * kk[xx, :n, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bicubic_u8(img: np.ndarray, size: int) -> np.ndarray:
    """img [H, W, 3] uint8 -> [size, size, 3] uint8, bit-exact with PIL `Image.resize((size, size))`: horizontal
    pass first (skipped when the width already matches), then the vertical pass on the uint8 intermediate."""
    H, W, _ = img.shape
    if W != size:
        img = _pass(img, *precompute_coeffs(W, size))
    if H != size:
        img = _pass(img.transpose(1, 0, 2), *precompute_coeffs(H, size)).transpose(1, 0, 2)
    return np.ascontiguousarray(img)


def rescale_normalize(img_u8: np.ndarray, mean, std) -> np.ndarray:
    """transformers image_transforms.rescale (float64 multiply by 1/255, cast to float32) then normalize
    ((x - mean) / std in float32); HWC uint8 -> CHW float32."""
    x = (img_u8.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def image_preprocess(img_u8: np.ndarray, size: int, mean, std) -> np.ndarray:
    return rescale_normalize(resize_bicubic_u8(img_u8, size), mean, std)
