#!/usr/bin/env python
"""Developer tool: pure host (Python) cost per frame of the clip pipeline, without a GPU.

The C-ABI emulator (tests/cabi_emulator.py) runs one real frame so that the pinned result block holds a realistic frame
(30 tracks in memory, ~80 tracked boxes); then every libsmot entry point is replaced by a no-op that returns 0, so the
following frames execute exactly the product's host code -- plan lookups, staging, ctypes marshalling, the solver, the
next-frame memory, result objects -- on that block, and cProfile shows where the per-frame Python time goes.

  python tools/host_overhead.py [--frames 200] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import pytest  # noqa: E402
import torch  # noqa: E402


class _Switch(object):
    """Proxy of the emulated library: while ``null`` is set every compute entry point returns 0 at once (size / version queries
    keep their real answers).  The wrappers are stable objects, so launch lists built before the switch see it too."""
    KEEP = ("smot_abi_version", "smot_last_error", "smot_resample_ksize", "smot_resample_coeffs", "smot_conv2d_algo")

    def __init__(self, real):
        self._real, self._w, self.null = real, {}, False

    def __getattr__(self, name):
        if not name.startswith("smot_"):
            return getattr(self._real, name)
        if name not in self._w:
            real = getattr(self._real, name)
            keep = "workspace" in name or name in self.KEEP

            def w(*a, _real=real, _keep=keep):
                return _real(*a) if (_keep or not self.null) else 0
            w.__name__ = name
            self._w[name] = w
        return self._w[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--workload", default="selftest")
    args = ap.parse_args()
    import cabi_emulator
    mp = pytest.MonkeyPatch()
    fake = cabi_emulator.install_for_bench(mp)
    import bench
    from siammot_b200 import _lib, engine, ops, preprocess
    sw = _Switch(fake)
    for mod in (_lib, engine, ops, preprocess):
        mp.setattr(mod, "lib", lambda: sw)
    bench.select_workload(args.workload)
    n_tracks = 30
    bench.N_TRACKS = n_tracks
    h = bench.Harness("float32", torch.device("cuda", 0))
    h.model.results_on_host = True
    frames_u8 = bench.make_frames_u8(4, h.cfg)
    pre = h.eng.preprocessor()
    frames = torch.stack([pre(frames_u8[i]) for i in range(4)])
    h.prime(frames[0])
    hook = lambda t: h.restore()
    res = h.model.forward_clip([frames[i % 4] for i in range(4)], before_frame=hook)      # real (emulated) frames
    print("emulated frame: %d boxes, %d tracked" % (res[-1].bbox.shape[0], int((res[-1].get_field("ids") >= 0).sum())))
    sw.null = True
    seq = [frames[i % 4] for i in range(args.frames)]
    h.model.forward_clip(seq[:8], before_frame=hook)
    t0 = time.perf_counter()
    h.model.forward_clip(seq, before_frame=hook)
    dt = time.perf_counter() - t0
    print("host time per frame (no GPU work, no waiting): %.1f us" % (dt / args.frames * 1e6))
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        h.model.forward_clip(seq, before_frame=hook)
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(32)
    mp.undo()


if __name__ == "__main__":
    main()
