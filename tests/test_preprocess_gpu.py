"""GPU parity for the preprocessing kernels: bit-exact against the oracle (which is pinned to Pillow / torchvision
in tests/test_preprocess_cpu.py), and the raw-frame model path against the float-tensor path."""
import numpy as np
import pytest
import torch

import os

from helpers import CONFIG_DIR
from oracle import preprocess as opp
from siammot_b200.config import get_cfg
from siammot_b200.synth_clip import make_clip_u8
from siammot_b200.synthetic import make_state_dict

DEV = "cuda"

pytestmark = pytest.mark.gpu


def _cfg(min_size, max_size, bgr):
    cfg = get_cfg()
    cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST, cfg.DATALOADER.SIZE_DIVISIBILITY = min_size, max_size, 32
    cfg.INPUT.TO_BGR255 = bgr
    if not bgr:
        cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    return cfg


# (frame h, w, MIN_SIZE_TEST, MAX_SIZE_TEST): vertical pass only (720p), both passes down (1080p), both passes up,
# width only, nothing to resample, portrait
CASES = [(720, 1280, 800, 1280), (1080, 1920, 800, 1280), (100, 150, 256, 512), (96, 200, 96, 160), (96, 160, 96, 160),
         (333, 187, 160, 256)]


@pytest.mark.parametrize("bgr", [False, True])
@pytest.mark.parametrize("h,w,mn,mx", CASES)
def test_preprocess_is_bit_exact(h, w, mn, mx, bgr):
    from siammot_b200.preprocess import FramePreprocessor
    cfg = _cfg(mn, mx, bgr)
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 4] = (img[: h // 4] // 128) * 255
    ref = opp.preprocess(img, cfg)
    pre = FramePreprocessor(cfg, DEV)
    got = pre(torch.from_numpy(img).pin_memory())
    assert got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)
    # device-resident frame, numpy frame: same result
    assert torch.equal(pre(torch.from_numpy(img).to(DEV)).cpu(), ref)
    assert torch.equal(pre(img).cpu(), ref)


def test_model_accepts_raw_frames():
    from siammot_b200.modelling import build_siammot
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(CONFIG_DIR, "dla34_emm.yaml"))
    cfg.DTYPE = "float32"
    cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST = 192, 320   # 200x334 frames -> 192x320 network input
    sd = make_state_dict(cfg, 1)
    frames = make_clip_u8(3, 200, 334, 4, 7, cfg.INPUT.PIXEL_MEAN, cfg.INPUT.PIXEL_STD)   # seed 7: tracks start in frame 0
    assert opp.preprocess(frames[0].numpy(), cfg).shape == (3, 192, 320)

    def run(kind):
        model = build_siammot(cfg)
        model.load_state_dict(sd, strict=False)
        model = model.to(DEV).eval()
        if kind == "clip":
            return model.forward_clip([f for f in frames])
        return [model(f if kind == "raw" else opp.preprocess(f.numpy(), cfg).to(DEV))[0] for f in frames]

    ref, raw, clip = run("float"), run("raw"), run("clip")
    assert sum(len(r) for r in ref) > 0 and max(int(r.get_field("ids").max()) for r in ref if len(r)) >= 0
    for a, b, c in zip(ref, raw, clip):
        assert a.size == b.size == c.size == (320, 192)
        for o in (b, c):
            assert torch.equal(a.bbox, o.bbox) and torch.equal(a.get_field("scores"), o.get_field("scores"))
            assert torch.equal(a.get_field("ids"), o.get_field("ids"))
