"""Developer lab for the planar correlation kernel (one GPU):
  1. the channel-group ladder: us per launch (CUDA-graph replay of 20 back-to-back launches, as bench.py times the roofline
     kernel) for 2 / 4 / 8 / 16 planes per CTA x both MMA phases, at the three benchmark geometries;
  2. with tools/lab/libsmot_trace.so (python tools/lab/build_trace.py): per-warp phase stamps of the LAST launch of such a replay --
     where a CTA's time goes (template staging, wait for the predecessor grid, bulk-copy latency, MMA phase, result path).
Writes one JSON document to --out."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from siammot_b200 import _lib, ops  # noqa: E402

SLOTS = ["cta_start", "after_zero_fill_sync", "templates_staged|pdl_wait_returned(copy warp)", "after_staging_sync", "windows_arrived",
         "mma_done", "after_result_sync", "stores_issued"]


def planes(n, Cc, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, Cc, 30, 30, generator=g).half()
    k = (torch.randn(n, 15, 15, Cc, generator=g) / 15.).half().cuda()
    xp = torch.zeros((n, Cc, _lib.XCORR_PLANE), dtype=torch.float16, device="cuda")
    xp[:, :, :30 * _lib.XCORR_ROW_PITCH].view(n, Cc, 30, _lib.XCORR_ROW_PITCH)[..., :30] = x.cuda()
    return xp, k


def ladder(res):
    for n, Cc in ((30, 128), (80, 128), (30, 256), (16, 128), (8, 128)):
        xp, k = planes(n, Cc)
        out = torch.empty((n, 16, 16, Cc), dtype=torch.float16, device="cuda")
        nbytes = n * Cc * (30 * 30 + 15 * 15 + 16 * 16) * 2
        ref = ops.xcorr_planar(xp, k, mma_mode=1, channel_group=16).clone()
        for mode in (1, 0):
            for cg in (0, 2, 4, 8, 16):
                if cg == 0 and n * Cc > 28 * torch.cuda.get_device_properties(0).multi_processor_count:
                    continue
                L = _lib.lib()
                fn = lambda: _lib.check(L.smot_xcorr_planar_cfg(ops._ptr(xp), ops._ptr(k), ops._ptr(out), n, Cc, mode, cg, _lib.stream_ptr()), "x")
                t = bench.time_launches(fn)
                same = bool(torch.equal(out, ref)) if mode == 1 else None
                us = t.get("graph", {}).get("us_per_launch")
                res.append({"n": n, "C": Cc, "mma_mode": mode, "channel_group": cg, "ctas": (n * Cc // cg) if cg else min(n * Cc // 4, torch.cuda.get_device_properties(0).multi_processor_count), "graph_us": us,
                            "eager_us": t.get("eager", {}).get("us_per_launch"), "gbs": round(nbytes / (us * 1e-6) / 1e9, 1) if us else None,
                            "equal_to_cg16": same})
                print(res[-1], flush=True)


def trace(res, path):
    T = C.CDLL(path)
    T.smot_last_error.restype = C.c_char_p
    for n, Cc, cg in ((30, 128, 16),):
        xp, k = planes(n, Cc)
        out = torch.empty((n, 16, 16, Cc), dtype=torch.float16, device="cuda")
        ctas, warps = n * Cc // cg, cg + 1
        buf = torch.zeros((ctas, warps, 8, 2), dtype=torch.int64, device="cuda")
        assert T.smot_xcorr_trace_buffer(C.c_void_p(buf.data_ptr())) == 0
        args = (C.c_void_p(xp.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(out.data_ptr()), n, Cc, 1, cg)

        def launch():
            rc = T.smot_xcorr_planar_cfg(*args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            assert rc == 0, T.smot_last_error()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            launch()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                launch()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        b = buf.cpu()
        gt, ck = b[..., 0].double(), (b[..., 1] & ((1 << 48) - 1)).double()
        sm = (b[:, 0, 0, 1] >> 48) & 0xffff
        t0 = float(gt[:, :, 0].min())
        mma, cp = slice(0, cg), cg
        rel = gt - t0

        def st(x):
            x = x.flatten()
            return {"min": round(float(x.min()), 1), "median": round(float(x.median()), 1), "max": round(float(x.max()), 1)}
        per_sm = torch.bincount(sm.long(), minlength=148)
        doc = {"n": n, "C": Cc, "channel_group": cg, "ctas": ctas,
               "globaltimer_step_ns": float(torch.unique(gt.flatten()).diff().min()) if gt.numel() > 1 else None,
               "ctas_per_sm": {"min": int(per_sm[per_sm > 0].min()), "max": int(per_sm.max()), "sms_used": int((per_sm > 0).sum())},
               "timeline_ns_since_first_cta_start (MMA warps; copy warp where said)": {
                   "cta_start": st(rel[:, mma, 0]), "after_zero_fill_sync": st(rel[:, mma, 1]), "templates_staged": st(rel[:, mma, 2]),
                   "copy_warp_pdl_wait_returned": st(rel[:, cp, 2]), "after_staging_sync": st(rel[:, mma, 3]),
                   "windows_arrived": st(rel[:, mma, 4]), "mma_done": st(rel[:, mma, 5]), "after_result_sync": st(rel[:, mma, 6]),
                   "stores_issued": st(rel[:, mma, 7])},
               "phase_cycles_per_warp (clock64)": {
                   "template_load+zero_fill+sync": st(ck[:, mma, 1] - ck[:, mma, 0]), "template_scatter": st(ck[:, mma, 2] - ck[:, mma, 1]),
                   "pdl_wait+sync": st(ck[:, mma, 3] - ck[:, mma, 2]), "window_wait": st(ck[:, mma, 4] - ck[:, mma, 3]),
                   "mma_phase": st(ck[:, mma, 5] - ck[:, mma, 4]), "result_to_smem+sync": st(ck[:, mma, 6] - ck[:, mma, 5]),
                   "result_stores": st(ck[:, mma, 7] - ck[:, mma, 6]), "whole_warp": st(ck[:, mma, 7] - ck[:, mma, 0])},
               "copy_issue_to_arrival_ns": st(rel[:, mma, 4].max(dim=1).values - rel[:, cp, 2])}
        res.append(doc)
        print(json.dumps(doc), flush=True)
    T.smot_xcorr_trace_buffer(None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/xcorr_lab.json")
    ap.add_argument("--trace-lib", default=os.path.join(REPO, "tools", "lab", "libsmot_trace.so"))
    a = ap.parse_args()
    doc = {"ladder": [], "trace": []}
    ladder(doc["ladder"])
    if os.path.exists(a.trace_lib):
        trace(doc["trace"], a.trace_lib)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(doc, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
