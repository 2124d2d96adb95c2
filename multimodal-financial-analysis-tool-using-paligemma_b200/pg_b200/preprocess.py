"""Image pre-processing on the GPU: the reference's host recipe (processing_paligemma.py:13-50 --
`PIL.Image.resize(BICUBIC)` -> `* 1/255` -> float32 -> `(x - 0.5) / 0.5` -> CHW) reproduced bit for bit.

Pillow resizes 8-bit images with fixed-point arithmetic (libImaging/Resample.c): per output coordinate a window
[xmin, xmin+n) of the input and n 22-bit integer coefficients derived from the bicubic kernel (a = -0.5, support 2,
widened by the scale factor when shrinking), a horizontal pass into an 8-bit intermediate, then a vertical pass.
`resample_coeffs` restates that table computation in float64 exactly as the C code orders it; the CUDA kernels
(`pg_resample_u8`, `pg_u8_to_chw`) do the integer multiply-accumulates and the 256-entry value table.
`resample_u8_numpy` is the same arithmetic on the host: test infrastructure for the coefficient tables."""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2   # Resample.c


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=64)
def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """precompute_coeffs + normalize_coeffs_8bpc of Resample.c for the full-image box (in0 = 0, in1 = in_size):
    bounds int32 [out, 2] = (first input index, tap count), kk int32 [out, ksize], ksize."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def value_table(rescale_factor: float = 1 / 255.0, mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """float32 [256]: byte -> normalised value with the reference's dtypes (uint8 * python float -> float64 ->
    float32, then float32 subtraction and division: processing_paligemma.py:19-29)."""
    v = (np.arange(256, dtype=np.uint8) * rescale_factor).astype(np.float32)
    return ((v - np.array(mean, dtype=v.dtype)) / np.array(std, dtype=v.dtype)).astype(np.float32)


def resample_u8_numpy(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow's two-pass 8-bit bicubic resize of an (H, W, C) uint8 array, integer arithmetic on the host."""
    def one_axis(a, out_size):   # resamples axis 1 of (R, n, C)
        n = a.shape[1]
        if n == out_size:
            return a
        bounds, kk, _ = resample_coeffs(n, out_size)
        out = np.empty((a.shape[0], out_size, a.shape[2]), dtype=np.uint8)
        a64 = a.astype(np.int64)
        for xx in range(out_size):
            x0, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
            acc = (a64[:, x0:x0 + cnt, :] * kk[xx, :cnt].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        return out
    tmp = one_axis(img, out_w)                                            # horizontal pass first
    return one_axis(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2)     # then vertical


_dev_tables = {}


def _device_tables(in_size: int, out_size: int, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    key = (in_size, out_size, str(device))
    if key not in _dev_tables:
        bounds, kk, ksize = resample_coeffs(in_size, out_size)
        _dev_tables[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ksize)
    return _dev_tables[key]


def preprocess_images_cuda(images: Sequence[torch.Tensor], size: int, dtype: torch.dtype = torch.float32,
                           rescale_factor: float = 1 / 255.0, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """(H_i, W_i, 3) uint8 CUDA tensors -> (B, 3, size, size) `dtype` tensor, equal to the reference's
    `process_images` on the same pixels (bit-exact in float32; rounded once more for 16-bit dtypes, as the model's
    own `.to(dtype)` would).  Everything runs on the images' device through the C ABI; no host copy of pixel data."""
    from . import _cabi as cabi
    if not images:
        raise ValueError("no images")
    dev = images[0].device
    if dev.type != "cuda":
        raise RuntimeError("preprocess_images_cuda runs on CUDA tensors only (use PaliGemmaProcessor for host images)")
    L, st = cabi.lib(), cabi.stream()
    out = torch.empty((len(images), 3, size, size), dtype=dtype, device=dev)
    lut = torch.from_numpy(value_table(rescale_factor, mean, std)).to(dev)
    for i, im in enumerate(images):
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("images must be (H, W, 3) uint8 tensors")
        im = im.contiguous()
        H, W = int(im.shape[0]), int(im.shape[1])
        cur = im
        if W != size:    # horizontal pass: (H, W, 3) -> (H, size, 3)
            b, k, ks = _device_tables(W, size, dev)
            nxt = torch.empty((H, size, 3), dtype=torch.uint8, device=dev)
            cabi.check(L.pg_resample_u8(nxt.data_ptr(), cur.data_ptr(), b.data_ptr(), k.data_ptr(), ks, H, W, size, 3, 0, st),
                       "resample (horizontal)")
            cur = nxt
        if H != size:    # vertical pass: (H, size, 3) -> (size, size, 3)
            b, k, ks = _device_tables(H, size, dev)
            nxt = torch.empty((size, size, 3), dtype=torch.uint8, device=dev)
            cabi.check(L.pg_resample_u8(nxt.data_ptr(), cur.data_ptr(), b.data_ptr(), k.data_ptr(), ks, size, H, size, 3, 1, st),
                       "resample (vertical)")
            cur = nxt
        cabi.check(L.pg_u8_to_chw(out[i].data_ptr(), cur.data_ptr(), lut.data_ptr(), size, size, 3,
                                  cabi.DTYPE_CODE[dtype], st), "u8_to_chw")
    return out
