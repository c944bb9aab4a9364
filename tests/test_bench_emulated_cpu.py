"""bench.py's GPU arm cannot run in a container without a GPU; its control flow can: over tests/cabi_emulator.py the whole
`run_ours` path executes on the CPU (fp32, 2 steps) -- harness priming, the device-resident clip arm, the from-host clip arm,
the per-frame arms, the xcorr brackets, the JSON line.  Timings are meaningless here; the test pins the line's contract and
that the from-host clip arm tracks exactly what the device-resident arm tracked (bench.py refuses to report it otherwise)."""
import contextlib
import io
import json
import sys

import cabi_emulator


def test_bench_gpu_arm_control_flow_and_json_contract(monkeypatch):
    cabi_emulator.install_for_bench(monkeypatch)
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "3", "--dtype", "float32", "--no-cpu-baseline"])
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            bench.main()
    finally:
        bench.select_workload("720p30")
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert key in line, key
    assert line["steps"] == 2 and line["n_gpus"] == 1 and line["config"]["tracks_in_memory"] == 30
    assert line["config"]["tracked_boxes_per_step"] > 30          # the 30 tracks in memory are tracked, plus new detections
    e2e = line["e2e"]
    assert e2e["clip_error"] is None and e2e["api"].startswith("model.forward_clip")
    assert e2e["h2d_bytes_per_step"] >= 3 * 720 * 1280 and e2e["d2h_bytes_per_step"] > 0
    assert {"value", "unit"} <= set(e2e["per_frame_call"]) and {"value", "unit"} <= set(e2e["float32_chw_host_input"])
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["algorithmic_bytes"] == 30 * 128 * 1381 * 4 and 0 < r["frac"]
    assert line["gpu_launches"] > 0
