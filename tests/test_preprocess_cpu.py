"""CPU tests for the test-time preprocessing row: the oracle restatement is pinned
against the real third-party code the reference calls (Pillow's resize through torchvision F.resize, ToTensor,
F.normalize), and the library's host-side coefficient helper is checked against the oracle (no GPU needed)."""
import ctypes as C

import numpy as np
import pytest
import torch
from PIL import Image
from torchvision.transforms import functional as TF

from oracle import preprocess as opp
from siammot_b200 import _lib
from siammot_b200.config import get_cfg
from siammot_b200.preprocess import get_size

SIZES = [(720, 1280, 704, 1280), (1080, 1920, 704, 1280), (90, 160, 96, 160), (97, 131, 64, 96), (50, 70, 128, 192),
         (33, 47, 33, 96), (480, 640, 704, 928)]


@pytest.mark.parametrize("h,w,oh,ow", SIZES)
def test_oracle_resize_is_bit_exact_with_pillow(h, w, oh, ow):
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3] // 128) * 255  # saturated blocks: exercises the 0 / 255 clipping
    ref = np.asarray(TF.resize(Image.fromarray(img, "RGB"), (oh, ow)))  # what image_augmentation.py:46 calls
    assert np.array_equal(opp.pil_resize_bilinear(img, oh, ow), ref)


@pytest.mark.parametrize("to_bgr255", [False, True])
def test_oracle_transform_matches_torchvision_chain(to_bgr255):
    cfg = get_cfg()
    cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST, cfg.DATALOADER.SIZE_DIVISIBILITY = 96, 160, 32
    cfg.INPUT.TO_BGR255 = to_bgr255
    if not to_bgr255:
        cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(120, 200, 3), dtype=np.uint8)
    oh, ow = opp.get_size(200, 120, 96, 160, 32)
    assert (oh, ow) == (96, 160)
    t = TF.to_tensor(TF.resize(Image.fromarray(img, "RGB"), (oh, ow)))
    if to_bgr255:                                   # maskrcnn_benchmark transforms.Normalize.__call__
        t = t[[2, 1, 0]] * 255
    ref = TF.normalize(t, mean=cfg.INPUT.PIXEL_MEAN, std=cfg.INPUT.PIXEL_STD)
    assert torch.equal(opp.preprocess(img, cfg), ref)


def test_get_size_known_answers():
    # shipped yaml (MIN 800 / MAX 1280 / divisibility 32) and the MOT17 yaml (MIN 800 / MAX 1500)
    for fn in (opp.get_size, get_size):
        assert fn(1280, 720, 800, 1280, 32) == (704, 1280)
        assert fn(1920, 1080, 800, 1280, 32) == (704, 1280)
        assert fn(1920, 1080, 800, 1500, 32) == (800, 1408)
        assert fn(720, 1280, 800, 1280, 32) == (1280, 704)
        assert fn(640, 480, 800, 1333, 0) == (800, 1066)


@pytest.mark.parametrize("n_in,n_out", [(720, 704), (1080, 704), (1920, 1280), (50, 128), (97, 64), (33, 33)])
def test_library_coefficients_match_oracle(n_in, n_out):
    L = _lib.lib()
    ks = L.smot_resample_ksize(n_in, n_out)
    bounds = np.zeros((n_out, 2), dtype=np.int32)
    kk = np.zeros((n_out, ks), dtype=np.int32)
    assert L.smot_resample_coeffs(n_in, n_out, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p)) == 0
    rb, rk = opp.precompute_coeffs(n_in, n_out)
    assert ks == rk.shape[1] and np.array_equal(rb, bounds) and np.array_equal(rk, kk)


def test_oracle_matches_committed_golden_vectors():
    """tests/golden/preprocess.npz was written by Pillow / torchvision themselves (make_preprocess_golden.py): the oracle
    must reproduce it bit for bit whatever Pillow is installed where the tests run."""
    import os
    import sys
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    from make_preprocess_golden import CASES, frame
    gold = np.load(os.path.join(here, "preprocess.npz"))
    for (seed, h, w, oh, ow, mean, std, bgr) in CASES:
        img = frame(seed, h, w)
        resized = opp.pil_resize_bilinear(img, oh, ow)
        assert np.array_equal(resized, gold["resized_%d" % seed])
        assert np.array_equal(opp.normalize(resized, mean, std, bgr).numpy(), gold["tensor_%d" % seed])
