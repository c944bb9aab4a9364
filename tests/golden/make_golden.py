"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/siammot/modelling/**, operator patches, configs/defaults.py and the shipped yaml
files) on CPU over the maskrcnn_benchmark stand-in in oracle/shim.

Run in the authoring container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py [scenario ...]

Stored per scenario: for every frame the final boxes / scores / ids / labels returned by
``SiamMOT.forward`` (rcnn.py:68) plus three intermediates that localise a mismatch -- FPN feature
statistics, the first RPN proposals and the EMM track boxes before refinement.  Inputs and weights
are NOT stored: they are pure functions of the seeds (siammot_b200/synthetic.py, synth_clip.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from oracle import reference_loader  # noqa: E402
from scenarios import FULL_SCENARIOS, GIVEN_SCENARIOS, ORACLE_SCENARIOS, SCENARIOS, given_boxes, golden_path, inject_boxes  # noqa: E402
from siammot_b200.synthetic import make_state_dict  # noqa: E402
from siammot_b200.synth_clip import make_clip  # noqa: E402


def build_reference(sc):
    cfg0, build = reference_loader.load()
    cfg = cfg0.clone()
    cfg.merge_from_file(os.path.join(reference_loader.REFERENCE_ROOT, "configs", "dla", sc["yaml"]))
    cfg.merge_from_list(sc["overrides"])
    cfg.MODEL.DEVICE = "cpu"
    model = build(cfg).eval()
    sd = model.state_dict()
    new = make_state_dict(cfg, sc["weight_seed"])
    if sc.get("tweak"):
        from fp16_scene import apply_tweak
        new = apply_tweak(new, cfg, sc["tweak"])
    sd.update(new)
    model.load_state_dict(sd)
    return cfg, model


def run(name):
    sc = SCENARIOS.get(name) or ORACLE_SCENARIOS.get(name) or FULL_SCENARIOS[name]
    full = name in FULL_SCENARIOS
    cfg, model = build_reference(sc)
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    clip = make_clip(sc["frames"], sc["H"], sc["W"], sc["n_obj"], sc["clip_seed"])
    taps = {}
    model.backbone.register_forward_hook(lambda m, i, o: taps.__setitem__("feats", o))
    model.rpn.register_forward_hook(lambda m, i, o: taps.__setitem__("props", o[0][0]))
    model.roi_heads.track.register_forward_hook(lambda m, i, o: taps.__setitem__("tracks", o[1]))
    model.reset_siammot_status()
    frames = []
    start = 0
    with torch.no_grad():
        if full or sc["inject"] is not None:
            # seed the track table directly: active tracks whose templates come from frame 0
            feats = model.backbone(clip[0][None])
            head = model.roi_heads.track
            head.track_pool.reset()
            if full:
                from fp16_scene import track_table
                boxes = track_table(sc["tracks"], sc["H"], sc["W"])
            else:
                boxes = inject_boxes(sc["inject"])
            det = BoxList(boxes, (sc["W"], sc["H"]), mode="xyxy")
            det.add_field("ids", torch.tensor([head.track_pool.start_track() for _ in range(len(boxes))]))
            det.add_field("labels", torch.ones(len(boxes), dtype=torch.int64))
            det.add_field("scores", torch.full((len(boxes),), 0.9))
            model.flush_memory(head.get_track_memory(feats, [det]))
            head.track_pool.increment_frame()
            start = 1
        for t in range(start, sc["frames"]):
            taps.clear()
            out = model(clip[t])[0]
            rec = dict(boxes=out.bbox.clone(), scores=out.get_field("scores").clone(),
                       ids=out.get_field("ids").clone(), labels=out.get_field("labels").clone(),
                       feat_stats=torch.tensor([[f.mean(), f.abs().mean(), f[0, 0, 0, 0], f[0, -1, -1, -1]]
                                                for f in taps["feats"]]),
                       props=taps["props"].bbox[:32].clone(),
                       objectness=taps["props"].get_field("objectness")[:32].clone())
            trk = taps.get("tracks")
            if trk is not None:
                rec["track_boxes"] = trk[0].bbox.clone()
                rec["track_scores"] = trk[0].get_field("scores").clone()
                rec["track_ids"] = trk[0].get_field("ids").clone()
            pool = model.roi_heads.track.track_pool
            rec["active"] = sorted(pool._active_ids)
            rec["dormant"] = sorted(pool._dormant_ids.keys())
            frames.append(rec)
            print(name, "frame", t, "boxes", len(out), "tracked", int((rec["ids"] >= 0).sum()),
                  "active", len(rec["active"]), "dormant", len(rec["dormant"]))
    extra = {}
    if full:
        # the oracle's view of the same clip: its outputs must equal the reference's (checked here, so the fixture is only
        # written when they do) and its decision margins are stored for the GPU test to assert on
        from decisive import MarginOracle, min_margin
        from fp16_scene import build_scene
        scene = build_scene(sc["weight_seed"], sc["clip_seed"], sc["frames"] - 1, sc["tweak"], workload=sc["workload"], tracks=sc["tracks"],
                            n_obj=sc["n_obj"])
        mo = MarginOracle(scene["cfg"], scene["sd"])
        mo.inject(scene["clip"][0], scene["boxes"])
        margins = []
        for t in range(1, sc["frames"]):
            out, m = mo.step(scene["clip"][t], with_emm_gap=True)
            m.pop("top_logit_diff", None)
            margins.append(m)
            ref = frames[t - 1]
            assert torch.equal(out["ids"], ref["ids"]) and torch.equal(out["labels"], ref["labels"]), "oracle ids differ from the reference"
            assert float((out["boxes"] - ref["boxes"]).abs().max()) <= 1e-3 and float((out["scores"] - ref["scores"]).abs().max()) <= 1e-4
        extra = dict(margins=margins, min_margin=min_margin(margins))
        for f in frames:           # keep the fixture small: final outputs + track boxes only
            for k in ("feat_stats", "props", "objectness"):
                f.pop(k, None)
        print(name, "oracle == reference on every frame; min margin", extra["min_margin"])
    torch.save(dict(scenario=name, spec=sc, torch=torch.__version__, frames=frames, **extra), golden_path(name))


def run_given(name):
    """Public-detection path: model(frame, given_detection=[BoxList]) every frame (roi_heads.py:26-34)."""
    sc = GIVEN_SCENARIOS[name]
    cfg, model = build_reference(sc)
    from maskrcnn_benchmark.structures.bounding_box import BoxList
    clip = make_clip(sc["frames"], sc["H"], sc["W"], sc["n_obj"], sc["clip_seed"])
    model.reset_siammot_status()
    frames = []
    with torch.no_grad():
        for t, boxes in enumerate(given_boxes(sc)):
            n = boxes.shape[0]
            bl = BoxList(boxes.clone(), (sc["W"], sc["H"]), mode="xyxy")
            bl.add_field("labels", torch.ones(n, dtype=torch.int64))
            bl.add_field("scores", torch.ones(n))
            bl.add_field("ids", torch.full((n,), -1, dtype=torch.int64))
            out = model(clip[t], given_detection=[bl])[0]
            pool = model.roi_heads.track.track_pool
            frames.append(dict(boxes=out.bbox.clone(), scores=out.get_field("scores").clone(), ids=out.get_field("ids").clone(),
                               labels=out.get_field("labels").clone(), active=sorted(pool._active_ids),
                               dormant=sorted(pool._dormant_ids.keys())))
            print(name, "frame", t, "given", n, "boxes", len(out), "tracked", int((frames[-1]["ids"] >= 0).sum()))
    torch.save(dict(scenario=name, spec=sc, torch=torch.__version__, frames=frames), golden_path(name))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(SCENARIOS) + list(ORACLE_SCENARIOS) + list(GIVEN_SCENARIOS)):
        (run_given if n in GIVEN_SCENARIOS else run)(n)
