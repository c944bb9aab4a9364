#!/usr/bin/env python
"""Developer probe behind tests/test_fp16_e2e_gpu.py: for each scene variant (weight seed : clip seed : head tweak) run the CPU
oracle over a 704x1280 clip with 30 injected tracks, report its decision margins (tests/decisive.py), and -- on a GPU box -- run
the engine in float16 and float32 over the same clip and report where ids / boxes / scores first differ.

  python tools/parity_probe.py --variants 1:0:base,1:0:sparse --frames 8 [--cpu-only] [--out gpurun_out/parity_probe.json]
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import torch  # noqa: E402


def calibrate(tops, b_diff, lo=4, hi=60, min_starts=1):
    """tops[t]: descending foreground-minus-background logits d of frame t's proposals (the largest 40); b_diff: the bias part of
    d.  Scaling the class weights by s and adding B to the background bias moves every d to s * (d - b_diff) + b_diff - B; a
    proposal is a detection when that exceeds logit(0.05) = -2.944 and may start a track above logit(0.6) = 0.405.  Returns the
    (s, B) whose nearest logit to either threshold is farthest away, among the settings with lo..hi detections and at least
    min_starts start candidates in the clip (and whose 40-deep lists are deep enough to have seen every near-threshold logit)."""
    import math
    t_det, t_start = math.log(0.05 / 0.95), math.log(0.6 / 0.4)
    best = None
    for sc in (1.0, 1.5, 2.0, 3.0, 4.0):
        fr = [[sc * (v - b_diff) + b_diff for v in f] for f in tops]
        allv = sorted(v for f in fr for v in f)
        B = allv[0] - t_det - 1.0
        while B < allv[-1] - t_start + 1.0:
            ndet = sum(1 for v in allv if v - B > t_det)
            nstart = sum(1 for v in allv if v - B > t_start)
            deep = all(len(f) < 40 or f[-1] - B < t_det - 1.0 for f in fr)
            if lo <= ndet <= hi and nstart >= min_starts and deep:
                gap = min(min(abs(v - B - t_det), abs(v - B - t_start)) for v in allv)
                if best is None or gap > best[0]:
                    best = (round(gap, 4), sc, round(B, 2), ndet, nstart)
            B += 0.01
    return None if best is None else dict(logit_gap=best[0], cls_scale=best[1], extra_bg_bias=best[2], detections_in_clip=best[3],
                                          above_start_thresh=best[4])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="1:0:base")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--tracks", type=int, default=30)
    ap.add_argument("--workload", default="720p30")
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--emm-gap", action="store_true")
    ap.add_argument("--calibrate", action="store_true", help="suggest an extra background-bias shift that puts SCORE_THRESH and "
                                                              "START_TRACK_THRESH into the widest gaps of the clip's class logits")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import fp16_scene as fs
    from decisive import MarginOracle, min_margin
    import bench
    torch.set_num_threads(bench.host_threads())
    report = []
    for v in args.variants.split(","):
        wseed, cseed, tweak = v.split(":")
        sc = fs.build_scene(int(wseed), int(cseed), args.frames, tweak, workload=args.workload, tracks=args.tracks)
        t0 = time.time()
        mo = MarginOracle(sc["cfg"], sc["sd"])
        mo.inject(sc["clip"][0], sc["boxes"])
        ref, margins = [], []
        for t in range(1, args.frames + 1):
            out, m = mo.step(sc["clip"][t], with_emm_gap=args.emm_gap)
            ref.append(out)
            margins.append(m)
        rec = {"variant": v, "oracle_s": round(time.time() - t0, 1), "margins": margins, "min_margin": min_margin(margins),
               "tracked": [int((r["ids"] >= 0).sum()) for r in ref], "boxes": [int(r["ids"].numel()) for r in ref]}
        if not args.cpu_only:
            for dtype in ("float16", "float32"):
                got = fs.run_engine(sc, dtype)
                rec[dtype] = fs.compare(ref, got)
        if args.calibrate and all("top_logit_diff" in m for m in margins):
            cb = sc["sd"]["roi_heads.box.predictor.cls_score.bias"]
            rec["calibration"] = calibrate([m["top_logit_diff"] for m in margins], float(cb[1] - cb[0]))
        print(json.dumps(rec))
        sys.stdout.flush()
        report.append(rec)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
