#!/bin/bash
# Round-2 GPU call E (1 GPU): unrolled row-wise ROIAlign, fine host timers, ncu of the SR ROIAlign.
set +e
OUT=gpurun_out/r02e
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 300 python bench.py $B > "$OUT/bench_default2.json" 2> "$OUT/bench_default2.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_rows_kernel -s 4 -c 3 -f -o "$OUT/roi_align_rows" \
    python tools/run_frames.py --frames 4 --eager > "$OUT/ncu_roi_rows.log" 2>&1
tail -n 6 "$OUT/pytest_gpu.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"])
    print("   host", d["stage_ms"].get("host_per_frame_ms"), "spread", d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
grep -n "roi_align" "$OUT/launches_720p30_summary.txt" | head
