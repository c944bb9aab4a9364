"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's test-time frame
preprocessing, used to check siammot_b200/csrc/preprocess.cu.

The reference chain for one decoded RGB uint8 frame (demos/demo_inference.py:74-82 ->
siammot/data/adapters/augmentation/build_augmentation.py:52-66 with is_train=False):

  1. ImageResize.get_size + torchvision F.resize on a PIL image
     (siammot/data/adapters/augmentation/image_augmentation.py:21-50): PIL.Image.resize((ow, oh), BILINEAR)
  2. ToTensor: uint8 HWC -> float32 CHW / 255
  3. maskrcnn_benchmark transforms.Normalize(mean, std, to_bgr255): optional [2,1,0] * 255, then (x - mean) / std

Step 1's arithmetic lives in a third-party dependency that is not under /root/reference: Pillow (requirements_exact.txt:10
pins Pillow==10.0.1), src/libImaging/Resample.c: ImagingResample -> precompute_coeffs / normalize_coeffs_8bpc /
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc.  This file restates that published algorithm (separable
two-pass convolution, antialiased triangle filter, 22-bit fixed-point coefficients, uint8 intermediate image) and
tests/test_preprocess_cpu.py pins it bit-exactly against the Pillow installed in this image (12.2.0; the routine has not
changed since 10.0.1) on random frames, so parity for this row IS pinned.
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2  # Resample.c: coefficients are 22-bit fixed point


def get_size(w, h, min_size, max_size, size_divisibility):
    """ImageResize.get_size (image_augmentation.py:21-42) for a single test-time min_size. Returns (oh, ow)."""
    size = min_size
    if max_size is not None:
        mn, mx = float(min(w, h)), float(max(w, h))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if w < h:
        ow, oh = size, int(size * h / w)
    else:
        oh, ow = size, int(size * w / h)
    if size_divisibility > 0:
        oh = int(oh / size_divisibility) * size_divisibility
        ow = int(ow / size_divisibility) * size_divisibility
    return oh, ow


def _bilinear(x):
    x = -x if x < 0.0 else x
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs (box = the whole image) + normalize_coeffs_8bpc for the bilinear (triangle,
    support 1) filter.  Returns (bounds int32 [out,2] = (first tap, tap count), kk int32 [out, ksize])."""
    scale = filterscale = float(in_size) / float(out_size)
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
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
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img, bounds, kk):
    """One pass along axis 0 of a uint8 array (rows = resampled axis), 8bpc fixed-point like Resample.c."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for o in range(bounds.shape[0]):
        x0, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for t in range(n):
            acc += src[x0 + t] * int(kk[o, t])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear(img, oh, ow):
    """PIL.Image.resize((ow, oh), BILINEAR) on a uint8 HWC array: horizontal pass first, then vertical, each only
    if that dimension changes (ImagingResample)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    if ow != w:
        b, k = precompute_coeffs(w, ow)
        img = _resample_axis0(img.transpose(1, 0, 2), b, k).transpose(1, 0, 2)
    if oh != h:
        b, k = precompute_coeffs(h, oh)
        img = _resample_axis0(img, b, k)
    return np.ascontiguousarray(img)


def normalize(img_u8, mean, std, to_bgr255):
    """ToTensor + maskrcnn_benchmark Normalize, float32 op by op. img_u8: HWC uint8 -> (3,H,W) float32."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    if to_bgr255:
        x = x[[2, 1, 0]] * 255
    m = torch.tensor(mean, dtype=torch.float32)[:, None, None]
    s = torch.tensor(std, dtype=torch.float32)[:, None, None]
    return x.sub(m).div(s)


def preprocess(frame_u8, cfg):
    """The whole test-time transform for one RGB uint8 HWC frame -> normalised (3, oh, ow) float32 tensor."""
    h, w = frame_u8.shape[:2]
    I = cfg.INPUT
    oh, ow = get_size(w, h, I.MIN_SIZE_TEST, I.MAX_SIZE_TEST, cfg.DATALOADER.SIZE_DIVISIBILITY)
    return normalize(pil_resize_bilinear(frame_u8, oh, ow), I.PIXEL_MEAN, I.PIXEL_STD, I.TO_BGR255)
