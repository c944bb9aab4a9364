#!/usr/bin/env python
"""Developer tool: run K natural frames of a bench workload through model(frame) (the harness restores the fixed track memory
before every frame, exactly as bench.py's per-frame arm does) -- the lean command to put under ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
      python tools/run_frames.py --frames 3 --eager
  ncu --set full --clock-control none --import-source on -k regex:xcorr -s 2 -c 1 -o gpurun_out/xcorr \
      python tools/run_frames.py --frames 4
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--workload", default="720p30")
    ap.add_argument("--dtype", default="float16")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs: every launch is a plain stream launch")
    ap.add_argument("--clip", action="store_true", help="forward_clip instead of model(frame)")
    args = ap.parse_args()
    import bench
    bench.select_workload(args.workload)
    torch.cuda.set_device(0)
    h = bench.Harness(args.dtype, torch.device("cuda", 0))
    if args.eager:
        h.eng.use_graph = False
        h.eng.frame_overlap = False
    h.model.results_on_host = True
    frames_u8 = bench.make_frames_u8(4, h.cfg)
    pre = h.eng.preprocessor()
    frames = torch.stack([pre(frames_u8[i]) for i in range(4)])
    h.prime(frames[0])
    if args.clip:
        res = h.model.forward_clip([frames[i % 4] for i in range(args.frames)], before_frame=lambda t: h.restore())
    else:
        res = [h.step(frames[i % 4]) for i in range(args.frames)]
    torch.cuda.synchronize()
    print("frames %d, tracked boxes per frame %s" % (args.frames, [int((r.get_field("ids") >= 0).sum()) for r in res]))


if __name__ == "__main__":
    main()
