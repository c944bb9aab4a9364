"""Test-time frame preprocessing on the GPU.

Host-side mirror of the reference's test transform -- ``build_siam_augmentation(cfg, is_train=False)``
(siammot/data/adapters/augmentation/build_augmentation.py:52-66) as used by ``DemoInference._preprocess``
(demos/demo_inference.py:74-82): ImageResize (PIL bilinear) -> ToTensor -> Normalize.  The reference does this on
the CPU and copies a float32 CHW tensor (10.8 MB at 720p) to the device; here the uint8 frame (2.8 MB) is copied and
the rest runs in two libsmot kernels, bit-identical to the CPU chain (tests/test_preprocess_gpu.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def get_size(w, h, min_size, max_size, size_divisibility):
    """ImageResize.get_size (image_augmentation.py:21-42) with one test-time min size.  Returns (oh, ow)."""
    size = min_size[0] if isinstance(min_size, (list, tuple)) else min_size
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


class _Geometry(object):
    """Device-resident resampling tables and staging buffers for one source frame size."""

    def __init__(self, h, w, oh, ow, device):
        L = lib()
        self.h, self.w, self.oh, self.ow = h, w, oh, ow
        self.frame = torch.empty((h, w, 3), dtype=torch.uint8, device=device)      # H2D target
        self.tmp = torch.empty((h, ow, 3), dtype=torch.uint8, device=device) if ow != w else None
        self.hb = self.hk = self.vb = self.vk = None
        self.hks = self.vks = 0
        if ow != w:
            self.hb, self.hk, self.hks = self._table(L, w, ow, device)
        if oh != h:
            self.vb, self.vk, self.vks = self._table(L, h, oh, device)

    @staticmethod
    def _table(L, n_in, n_out, device):
        ks = L.smot_resample_ksize(n_in, n_out)
        bounds = np.zeros((n_out, 2), dtype=np.int32)
        kk = np.zeros((n_out, ks), dtype=np.int32)
        check(L.smot_resample_coeffs(n_in, n_out, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p)),
              "smot_resample_coeffs")
        return torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ks


class FramePreprocessor(object):
    """``pre(frame_u8_hwc) -> (3, oh, ow) float32 CUDA tensor``; ``pre.into(frame, out)`` writes into a given buffer."""

    def __init__(self, cfg, device="cuda"):
        lib()  # fail loudly without the CUDA library: there is no CPU path
        I = cfg.INPUT
        self.min_size, self.max_size = I.MIN_SIZE_TEST, I.MAX_SIZE_TEST
        self.div = cfg.DATALOADER.SIZE_DIVISIBILITY
        self.mean = (C.c_float * 3)(*[float(v) for v in I.PIXEL_MEAN])
        self.std = (C.c_float * 3)(*[float(v) for v in I.PIXEL_STD])
        self.to_bgr255 = int(bool(I.TO_BGR255))
        self.device = torch.device(device)
        self._geo = {}

    def output_size(self, h, w):
        return get_size(w, h, self.min_size, self.max_size, self.div)

    def geometry(self, h, w, lane=0):
        """Staging buffers + coefficient tables of one input size; ``lane`` selects an independent copy of the staging
        buffers (callers that transform frames on several streams at once use one lane per stream)."""
        g = self._geo.get((h, w, lane))
        if g is None:
            oh, ow = self.output_size(h, w)
            g = self._geo[(h, w, lane)] = _Geometry(h, w, oh, ow, self.device)
        return g

    def into(self, frame, out, lane=0):
        """frame: uint8 (H, W, 3) RGB tensor / numpy array (host -- ideally pinned -- or device); out: float32
        (3, oh, ow) contiguous CUDA tensor.  Enqueues the copy and the kernels on the current stream."""
        if isinstance(frame, np.ndarray):
            frame = torch.from_numpy(frame)
        if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError("expected a uint8 (H, W, 3) RGB frame, got %s %s" % (frame.dtype, tuple(frame.shape)))
        g = self.geometry(frame.shape[0], frame.shape[1], lane)
        if tuple(out.shape) != (3, g.oh, g.ow) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("output must be a contiguous float32 (3, %d, %d) tensor" % (g.oh, g.ow))
        g.frame.copy_(frame, non_blocking=True)
        L, st = lib(), _lib.stream_ptr()
        src, pitch = g.frame, g.w * 3
        if g.tmp is not None:
            check(L.smot_resample_h_u8(src.data_ptr(), pitch, g.h, g.w, g.hb.data_ptr(), g.hk.data_ptr(), g.hks, g.ow,
                                       g.tmp.data_ptr(), g.ow * 3, st), "smot_resample_h_u8")
            src, pitch = g.tmp, g.ow * 3
        vb = g.vb.data_ptr() if g.vb is not None else None
        vk = g.vk.data_ptr() if g.vk is not None else None
        check(L.smot_resample_v_normalize(src.data_ptr(), pitch, g.h, g.ow, vb, vk, g.vks, g.oh, C.byref(self.mean),
                                          C.byref(self.std), self.to_bgr255, out.data_ptr(), st), "smot_resample_v_normalize")
        return out

    def __call__(self, frame):
        h, w = frame.shape[0], frame.shape[1]
        oh, ow = self.output_size(h, w)
        return self.into(frame, torch.empty((3, oh, ow), dtype=torch.float32, device=self.device))
