"""GPU parity of the less-travelled branches of the path against the CPU oracle (fp32 arithmetic):
public detections (roi_heads.py:26-34), CrowdHuman-density track counts (BASELINE configs[2]: 80 tracks),
the TRACKTOR scoring switch (roi_heads.py:72-76) and the reference's failure mode when every track is lost."""
import numpy as np
import pytest
import torch

from helpers import scenario_cfg
from siammot_b200.synthetic import make_state_dict
from siammot_b200.synth_clip import make_clip

pytestmark = pytest.mark.gpu
BOX_TOL, SCORE_TOL = 1e-3, 1e-3


def _pair(cfg, seed=1):
    from oracle.siammot_oracle import OracleSiamMOT
    from siammot_b200.modelling import build_siammot
    sd = make_state_dict(cfg, seed)
    model = build_siammot(cfg)
    model.load_state_dict(sd, strict=False)
    return model.to("cuda").eval(), OracleSiamMOT(cfg, sd)


def _check(got, ref, t):
    assert got.bbox.shape[0] == ref["boxes"].shape[0], "frame %d: %d vs %d boxes" % (t, got.bbox.shape[0], ref["boxes"].shape[0])
    assert torch.equal(got.get_field("ids").cpu(), ref["ids"]), "frame %d ids" % t
    assert torch.equal(got.get_field("labels").cpu(), ref["labels"])
    if got.bbox.shape[0] == 0:
        return
    assert float((got.bbox.cpu() - ref["boxes"]).abs().max()) <= BOX_TOL
    assert float((got.get_field("scores").cpu() - ref["scores"]).abs().max()) <= SCORE_TOL


def test_given_detections_match_oracle():
    from siammot_b200.structures import BoxList
    cfg = scenario_cfg("emm_amodal_expire_192x320")
    model, orc = _pair(cfg, 2)
    clip = make_clip(4, 192, 320, 5, 3)
    g = torch.Generator().manual_seed(4)
    model.reset_siammot_status()
    orc.reset()
    for t in range(4):
        n = 0 if t == 2 else 24                       # one frame with no public detections at all
        xy = torch.rand(n, 2, generator=g) * torch.tensor([250., 120.])
        wh = torch.rand(n, 2, generator=g) * torch.tensor([50., 60.]) + 10
        boxes = torch.cat([xy, xy + wh], 1)
        bl = BoxList(boxes.clone(), (320, 192), mode="xyxy")
        bl.add_field("labels", torch.ones(n, dtype=torch.int64))
        bl.add_field("scores", torch.ones(n))
        bl.add_field("ids", torch.full((n,), -1, dtype=torch.int64))
        got = model(clip[t].to("cuda"), given_detection=[bl])[0]
        if n:
            given = dict(boxes=boxes, ids=torch.full((n,), -1, dtype=torch.int64), labels=torch.ones(n, dtype=torch.int64))
        else:
            given = dict(boxes=torch.zeros((0, 4)), scores=torch.zeros((0,)), ids=torch.zeros((0,), dtype=torch.int64),
                         labels=torch.zeros((0,), dtype=torch.int64))
        ref = orc.forward(clip[t], given_detection=given)
        _check(got, ref, t)


def test_given_detections_with_tracks_match_reference_golden():
    """The public-detection scenario whose expected outputs come from the reference itself
    (tests/golden/given_det_192x320.pt): tracks start, persist through a frame without detections, and lapse."""
    from helpers import load_golden
    from scenarios import GIVEN_SCENARIOS, given_boxes
    from siammot_b200.config import get_cfg
    from siammot_b200.structures import BoxList
    from helpers import CONFIG_DIR, YAML_MAP
    import os
    sc = GIVEN_SCENARIOS["given_det_192x320"]
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(CONFIG_DIR, YAML_MAP[sc["yaml"]]))
    cfg.merge_from_list(sc["overrides"])
    cfg.DTYPE = "float32"
    model, _ = _pair(cfg, sc["weight_seed"])
    gold = load_golden("given_det_192x320")["frames"]
    clip = make_clip(sc["frames"], sc["H"], sc["W"], sc["n_obj"], sc["clip_seed"])
    model.reset_siammot_status()
    for t, (boxes, g) in enumerate(zip(given_boxes(sc), gold)):
        n = boxes.shape[0]
        bl = BoxList(boxes.clone(), (sc["W"], sc["H"]), mode="xyxy")
        bl.add_field("labels", torch.ones(n, dtype=torch.int64))
        bl.add_field("scores", torch.ones(n))
        bl.add_field("ids", torch.full((n,), -1, dtype=torch.int64))
        got = model(clip[t].to("cuda"), given_detection=[bl])[0]
        _check(got, g, t)
        pool = model.roi_heads.track.track_pool
        assert sorted(pool.get_active_ids()) == g["active"] and sorted(pool._dormant_ids.keys()) == g["dormant"]


def test_eighty_tracks_match_oracle():
    """80 tracks in memory (CrowdHuman density, BASELINE configs[2]) seeded on a 384x640 frame pair."""
    from oracle.siammot_oracle import build_memory
    cfg = scenario_cfg("emm_256x384")
    model, orc = _pair(cfg, 1)
    clip = make_clip(2, 384, 640, 6, 7)
    g = torch.Generator().manual_seed(9)
    n = 80
    c = torch.rand(n, 2, generator=g) * torch.tensor([560., 300.]) + torch.tensor([40., 40.])
    wh = torch.rand(n, 2, generator=g) * torch.tensor([60., 120.]) + torch.tensor([12., 24.])
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    labels = torch.ones(n, dtype=torch.int64)
    # engine
    eng = model.engine()
    P = eng.run_static(clip[0].to("cuda"))
    pool = model.roi_heads.track.track_pool
    pool.reset()
    ids = torch.tensor([pool.start_track() for _ in range(n)])
    model.flush_memory(model.roi_heads._build_memory(P, boxes.numpy(), ids.numpy(), labels.numpy()))
    pool.increment_frame()
    got = model(clip[1].to("cuda"))[0]
    # oracle
    feats = orc.features(clip[0])
    orc.pool.reset()
    oids = torch.tensor([orc.pool.start() for _ in range(n)])
    orc.memory = build_memory(orc.P, cfg, orc.pool, feats, dict(boxes=boxes, scores=torch.full((n,), 0.9), ids=oids, labels=labels))
    orc.pool.frame += 1
    ref = orc.forward(clip[1])
    assert int((ref["ids"] >= 0).sum()) >= 20
    _check(got, ref, 1)


def test_tracktor_switch_matches_oracle():
    cfg = scenario_cfg("emm_256x384")
    cfg.MODEL.TRACK_HEAD.TRACKTOR = True
    model, orc = _pair(cfg, 1)
    clip = make_clip(3, 256, 384, 6, 0)
    model.reset_siammot_status()
    orc.reset()
    for t in range(3):
        _check(model(clip[t].to("cuda"))[0], orc.forward(clip[t]), t)


def test_all_tracks_lost_raises_like_the_reference():
    """roi_heads.py:64-65 returns a bare BoxList when every propagated track is clipped away and :44 then fails
    on list + BoxList; the engine reproduces the TypeError instead of silently continuing."""
    cfg = scenario_cfg("emm_256x384")
    model, _ = _pair(cfg, 1)
    clip = make_clip(2, 256, 384, 6, 0)
    eng = model.engine()
    P = eng.run_static(clip[0].to("cuda"))
    pool = model.roi_heads.track.track_pool
    pool.reset()
    # a template far outside the frame: the decoded box is clipped to an empty one
    boxes = np.array([[3000., 3000., 3050., 3100.]], dtype=np.float32)
    ids = np.array([pool.start_track()], dtype=np.int64)
    model.flush_memory(model.roi_heads._build_memory(P, boxes, ids, np.ones(1, dtype=np.int64)))
    pool.increment_frame()
    with pytest.raises(TypeError):
        model(clip[1].to("cuda"))


def test_integration_md_operator_stubs_run_as_written():
    """INTEGRATION.md section B shows the ctypes stubs a maintainer would put behind maskrcnn_benchmark's `_C.nms` /
    `_C.roi_align_forward`.  Execute that code block VERBATIM and compare with independent implementations."""
    import os
    import re
    import torchvision
    from oracle import prims
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(repo, "INTEGRATION.md")).read()
    block = [b for b in re.findall(r"```python\n(.*?)```", md, re.S) if "def nms(" in b and "def roi_align_forward(" in b]
    assert len(block) == 1
    code = block[0].replace('"siammot_b200/libsmot.so"', repr(os.path.join(repo, "siammot_b200", "libsmot.so")))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    g = torch.Generator().manual_seed(0)
    # ---- nms: score-descending keep list, IoU with +1 widths, suppress on IoU > thresh
    for n in (1, 57, 700):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 80 + 4
        dets, scores = torch.cat([xy, xy + wh], 1), torch.rand(n, generator=g)
        keep = ns["nms"](dets.cuda(), scores.cuda(), 0.5)
        assert keep.dtype == torch.int64 and torch.equal(keep.cpu(), prims.nms_legacy(dets, scores, 0.5))
    assert ns["nms"](torch.zeros((0, 4), device="cuda"), torch.zeros((0,), device="cuda"), 0.5).numel() == 0
    # ---- roi_align_forward: NCHW in, (K, C, ph, pw) out, legacy alignment, rois carry the batch index
    for dt, tol in ((torch.float32, 2e-5), (torch.float16, 2e-3)):
        x = torch.randn(2, 24, 40, 56, generator=g)
        b = torch.randint(0, 2, (19,), generator=g).float()
        xy = torch.rand(19, 2, generator=g) * torch.tensor([150., 100.])
        wh = torch.rand(19, 2, generator=g) * 80 + 2
        rois = torch.cat([b[:, None], xy, xy + wh], 1)
        got = ns["roi_align_forward"](x.to("cuda", dt), rois.cuda(), 0.25, 7, 7, 2)
        ref = torchvision.ops.roi_align(x.to(dt).float(), rois, (7, 7), 0.25, 2, False)
        assert tuple(got.shape) == (19, 24, 7, 7)
        assert float((got.float().cpu() - ref).abs().max()) <= tol * float(ref.abs().max())
