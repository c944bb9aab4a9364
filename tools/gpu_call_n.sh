#!/bin/bash
# Round-2 GPU call N (1 GPU): ROIAlign rows kernel after the instruction diet (tap tables: element offsets + zeroed weights).
set +e
OUT=gpurun_out/r02n
mkdir -p "$OUT"
timeout 1200 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
B="--steps 100 --warmup 10 --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_ROI_UNROLL=1 timeout 300 python bench.py $B > "$OUT/bench_roi_unroll.json" 2> "$OUT/bench_roi_unroll.err"
SMOT_ROI_ROWS=0 timeout 300 python bench.py $B > "$OUT/bench_roi_old.json" 2> "$OUT/bench_roi_old.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
SMOT_ROI_UNROLL=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_unroll.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches2.log" 2>&1
python tools/launch_report.py "$OUT/launches_unroll.csv" > "$OUT/launches_unroll_summary.txt" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_rows_kernel -s 4 -c 4 -f -o "$OUT/roi_rows" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_roi.log" 2>&1
ncu -i "$OUT/roi_rows.ncu-rep" --page raw --csv > "$OUT/roi_rows_raw.csv" 2> /dev/null
python tools/ncu_summary.py "$OUT/roi_rows_raw.csv" > "$OUT/roi_rows_summary.csv" 2>&1
rm -f "$OUT/roi_rows.ncu-rep"
tail -n 4 "$OUT/pytest_gpu.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
grep -h "roi_align" "$OUT/launches_720p30_summary.txt" "$OUT/launches_unroll_summary.txt"
cat "$OUT/roi_rows_summary.csv" | cut -c1-250
