"""The full-size scenes of the fp16 end-to-end parity tests: BASELINE.json configs[1] / [2] / [4] geometry (3x704x1280 or
3x1056x1920 network input, 30 / 80 tracks injected into the memory at frame 0), seeded synthetic weights, a synthetic clip.

``tweak`` reshapes the synthetic HEAD weights (never the architecture or the configuration) so that a scene's id-deciding
comparisons are sparse: with the plain recipe of siammot_b200/synthetic.py every frame has ~300 box-head scores spread over
(0, 1) and ~100 candidates at the solver, so some comparison always sits within fp16 noise of its threshold.  tests/decisive.py
measures the margins; tools/parity_probe.py is the search tool."""
import os

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG_DIR = os.path.join(REPO, "siammot_b200", "configs")

WORKLOADS = {
    "720p30": dict(net=(704, 1280), yaml="dla34_emm.yaml", opts=[]),
    "r50_720p30": dict(net=(704, 1280), yaml="r50_emm.yaml", opts=[]),
    "1080p80": dict(net=(1056, 1920), yaml="dla34_emm.yaml", opts=["INPUT.MIN_SIZE_TEST", 1080, "INPUT.MAX_SIZE_TEST", 1920]),
    "small": dict(net=(256, 384), yaml="dla34_emm.yaml", opts=[]),
}


def track_table(n, H, W, seed=123, max_iou=0.2):
    """n pedestrian-like boxes (x1, y1, x2, y2) spread over the frame, hitting FPN levels 0..2 (the size ranges of bench.py's
    table), drawn by rejection so that no two overlap by more than ``max_iou``: the solver's NMS 0.5 between two injected
    tracks of near-equal score would otherwise be a coin toss by construction."""
    g = torch.Generator().manual_seed(seed)
    sx, sy = (W / 1280., H / 704.) if H < 400 else (1.0, 1.0)
    out = []
    while len(out) < n:
        cx = float(torch.rand(1, generator=g)) * (W / sx - 200) + 100
        cy = float(torch.rand(1, generator=g)) * (H / sy - 300) + 150
        h = float(torch.rand(1, generator=g)) * 260 + 60
        w = h * (0.3 + 0.2 * float(torch.rand(1, generator=g)))
        b = torch.tensor([(cx - w / 2) * sx, (cy - h / 2) * sy, (cx + w / 2) * sx, (cy + h / 2) * sy])
        ok = True
        for o in out:
            iw = min(b[2], o[2]) - max(b[0], o[0])
            ih = min(b[3], o[3]) - max(b[1], o[1])
            if iw > 0 and ih > 0:
                inter = iw * ih
                if inter / ((b[2] - b[0]) * (b[3] - b[1]) + (o[2] - o[0]) * (o[3] - o[1]) - inter) > max_iou:
                    ok = False
                    break
        if ok:
            out.append(b)
    return torch.stack(out)


def apply_tweak(sd, cfg, tweak):
    """Head-weight variants (a pure function of the state dict).  Tokens joined by '+':
      bgN      box-head background bias += N           -> fewer proposals clear SCORE_THRESH 0.05
      clsxN    box-head class weights x N              -> class scores pushed towards 0 / 1 (fewer in the middle band)
      emmN     EMM foreground bias += N                -> tracks' confidences pushed towards 1
      emmxN    EMM cls weights x N
    """
    sd = dict(sd)
    pre_b, pre_t = "roi_heads.box.predictor.", "roi_heads.track.tracker.predictor."
    for tok in [t for t in tweak.split("+") if t and t != "base"]:
        if tok.startswith("bg"):
            b = sd[pre_b + "cls_score.bias"].clone()
            b[0] += float(tok[2:])
            sd[pre_b + "cls_score.bias"] = b
        elif tok.startswith("clsx"):
            sd[pre_b + "cls_score.weight"] = sd[pre_b + "cls_score.weight"] * float(tok[4:])
        elif tok.startswith("emmx"):
            sd[pre_t + "cls.weight"] = sd[pre_t + "cls.weight"] * float(tok[4:])
        elif tok.startswith("emm"):
            b = sd[pre_t + "cls.bias"].clone()
            b[1] += float(tok[3:])
            sd[pre_t + "cls.bias"] = b
        else:
            raise ValueError("unknown tweak token %r" % tok)
    return sd


def build_scene(weight_seed, clip_seed, frames, tweak="base", workload="720p30", tracks=30, n_obj=12):
    from siammot_b200.config import get_cfg
    from siammot_b200.synth_clip import make_clip
    from siammot_b200.synthetic import make_state_dict
    w = WORKLOADS[workload]
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(CONFIG_DIR, w["yaml"]))
    if w["opts"]:
        cfg.merge_from_list(list(w["opts"]))
    H, W = w["net"]
    sd = apply_tweak(make_state_dict(cfg, weight_seed), cfg, tweak)
    clip = make_clip(frames + 1, H, W, n_obj, clip_seed)          # frame 0 primes the memory
    return dict(cfg=cfg, sd=sd, clip=clip, boxes=track_table(tracks, H, W), H=H, W=W, tracks=tracks,
                name="%s:w%d:c%d:%s" % (workload, weight_seed, clip_seed, tweak))


def run_engine(scene, dtype, clip_api=False):
    """The engine over the scene (GPU): 30 tracks injected at frame 0 exactly as the oracle does, then frames 1..F through
    model(frame) (or forward_clip).  Returns per-frame dict(boxes, scores, ids, labels) on the CPU."""
    from siammot_b200.config import get_cfg  # noqa: F401
    from siammot_b200.modelling import build_siammot
    cfg = scene["cfg"].clone()
    cfg.DTYPE = dtype
    model = build_siammot(cfg)
    model.load_state_dict(scene["sd"], strict=False)
    model = model.to("cuda:0").eval()
    eng = model.engine()
    pool = model.roi_heads.track.track_pool
    clip = scene["clip"]
    n = scene["tracks"]
    P = eng.run_static(clip[0].to("cuda:0"))
    pool.reset()
    ids = torch.tensor([pool.start_track() for _ in range(n)])
    mem = model.roi_heads._build_memory(P, scene["boxes"].numpy(), ids.numpy(), torch.ones(n, dtype=torch.int64).numpy())
    pool.increment_frame()
    model.flush_memory(mem)
    frames = [clip[t].to("cuda:0") for t in range(1, clip.shape[0])]
    if clip_api:
        res = model.forward_clip(frames)
    else:
        res = [model(f)[0] for f in frames]
    out = []
    for r in res:
        r = r.to("cpu")
        out.append(dict(boxes=r.bbox.clone(), scores=r.get_field("scores").clone(), ids=r.get_field("ids").clone(),
                        labels=r.get_field("labels").clone()))
    return out


def compare(ref, got):
    """Per-frame comparison of engine output with the oracle's: same ids in the same order?  max |box| / |score| difference
    over the rows (when the id sequences agree), else over the boxes matched by track id."""
    rep = {"frames": [], "first_id_mismatch": None, "max_box_err": 0.0, "max_score_err": 0.0}
    for t, (a, b) in enumerate(zip(ref, got)):
        same = a["ids"].shape == b["ids"].shape and bool(torch.equal(a["ids"], b["ids"])) and bool(torch.equal(a["labels"], b["labels"]))
        fr = {"ids_equal": same, "n_ref": int(a["ids"].numel()), "n_got": int(b["ids"].numel()),
              "tracked_ref": int((a["ids"] >= 0).sum()), "tracked_got": int((b["ids"] >= 0).sum())}
        if same and a["ids"].numel():
            fr["box_err"] = float((a["boxes"] - b["boxes"]).abs().max())
            fr["score_err"] = float((a["scores"] - b["scores"]).abs().max())
        else:
            ia = {int(i): k for k, i in enumerate(a["ids"].tolist()) if i >= 0}
            ib = {int(i): k for k, i in enumerate(b["ids"].tolist()) if i >= 0}
            common = sorted(set(ia) & set(ib))
            fr["common_ids"] = len(common)
            fr["only_ref"] = sorted(set(ia) - set(ib))[:8]
            fr["only_got"] = sorted(set(ib) - set(ia))[:8]
            if common:
                ka, kb = torch.tensor([ia[i] for i in common]), torch.tensor([ib[i] for i in common])
                fr["box_err"] = float((a["boxes"][ka] - b["boxes"][kb]).abs().max())
                fr["score_err"] = float((a["scores"][ka] - b["scores"][kb]).abs().max())
            if rep["first_id_mismatch"] is None:
                rep["first_id_mismatch"] = t
        rep["max_box_err"] = max(rep["max_box_err"], fr.get("box_err", 0.0))
        rep["max_score_err"] = max(rep["max_score_err"], fr.get("score_err", 0.0))
        rep["frames"].append(fr)
    rep["ids_exact_all_frames"] = rep["first_id_mismatch"] is None
    return rep
