#!/bin/bash
# Round-2 GPU call H (1 GPU): planar correlation -- channel-group ladder + in-kernel phase trace; bench with the new default.
set +e
OUT=gpurun_out/r02h
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_more_gpu.py -q -m gpu -k "xcorr or planar" > "$OUT/pytest_xcorr.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_xcorr.txt"
timeout 600 python tools/xcorr_lab.py --out "$OUT/xcorr_lab.json" > "$OUT/xcorr_lab.log" 2>&1
echo "rc=$?" >> "$OUT/xcorr_lab.log"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_cg4.json" 2> "$OUT/bench_cg4.err"
SMOT_XCORR_CG=16 timeout 300 python bench.py $B > "$OUT/bench_cg16.json" 2> "$OUT/bench_cg16.err"
tail -n 4 "$OUT/pytest_xcorr.txt"
grep -v "^{\"n\": 30, \"C\": 128, \"channel_group\"" "$OUT/xcorr_lab.log" | cut -c1-400 | head -40
python - "$OUT/xcorr_lab.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for t in d["trace"]:
    print("TRACE n=%d C=%d cg=%d ctas=%d gt_step=%s per_sm=%s" % (t["n"],t["C"],t["channel_group"],t["ctas"],t["globaltimer_step_ns"],t["ctas_per_sm"]))
    for k,v in t["timeline_ns_since_first_cta_start (MMA warps; copy warp where said)"].items(): print("   T %-32s %s" % (k,v))
    for k,v in t["phase_cycles_per_warp (clock64)"].items(): print("   C %-32s %s" % (k,v))
    print("   copy_issue_to_arrival_ns", t["copy_issue_to_arrival_ns"])
PY
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "xcorr us", d["roofline"]["us_per_launch"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("ERR", e)
PY
done
