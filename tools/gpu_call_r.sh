#!/bin/bash
# Round-2 GPU call R (1 GPU): in-CTA K slices of the few-tile tcgen05 layers -- bit equality with the split path, suite, bench.
set +e
OUT=gpurun_out/r02r
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "k_slices or split_k or tcgen05" > "$OUT/pytest_slices.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_slices.txt"
timeout 1200 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
B="--steps 100 --warmup 10 --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_TC_SLICED=0 timeout 300 python bench.py $B > "$OUT/bench_split.json" 2> "$OUT/bench_split.err"
timeout 300 python bench.py $B --workload 1080p80 > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 300 python bench.py $B --workload r50_720p30 > "$OUT/bench_r50_720p30.json" 2> "$OUT/bench_r50_720p30.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 6 "$OUT/pytest_slices.txt"; tail -n 4 "$OUT/pytest_gpu.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
tail -16 "$OUT/launches_720p30_summary.txt"
