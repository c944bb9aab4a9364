#!/bin/bash
# Round-2 GPU call I (1 GPU): persistent stem / level0 kernels, correlation kernel with per-copy mbarriers + 64-bit B words.
set +e
OUT=gpurun_out/r02i
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_more_gpu.py -q -m gpu -k "hires or xcorr or planar" > "$OUT/pytest_new.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_new.txt"
timeout 600 python tools/xcorr_lab.py --out "$OUT/xcorr_lab.json" > "$OUT/xcorr_lab.log" 2>&1
echo "rc=$?" >> "$OUT/xcorr_lab.log"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_HIRES_PERSIST=0 timeout 300 python bench.py $B > "$OUT/bench_nopersist.json" 2> "$OUT/bench_nopersist.err"
timeout 900 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 5 "$OUT/pytest_new.txt"; tail -n 5 "$OUT/pytest_gpu.txt"
grep "'mma_mode': 1" "$OUT/xcorr_lab.log" | cut -c1-200
python - "$OUT/xcorr_lab.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for t in d["trace"][:2]:
    print("TRACE n=%d C=%d cg=%d ctas=%d per_sm=%s" % (t["n"],t["C"],t["channel_group"],t["ctas"],t["ctas_per_sm"]))
    for k,v in t["timeline_ns_since_first_cta_start (MMA warps; copy warp where said)"].items(): print("   T %-32s %s" % (k,v))
    for k,v in t["phase_cycles_per_warp (clock64)"].items(): print("   C %-32s %s" % (k,v))
PY
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], "xcorr us", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"], d["e2e"]["clip_error"])
except Exception as e:
    print("ERR", e)
PY
done
head -12 "$OUT/launches_720p30_summary.txt"
