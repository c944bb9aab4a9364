#!/bin/bash
# Round-2 GPU call V (1 GPU): ROIAlign planar epilogue without divisions; rolled vs unrolled sampling loops, alternating runs.
set +e
OUT=gpurun_out/r02v
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_more_gpu.py -q -m gpu -k "roi_align or planar" > "$OUT/pytest_roi.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_roi.txt"
B="--steps 100 --warmup 10 --no-cpu-baseline"
for i in 1 2; do
  timeout 300 python bench.py $B > "$OUT/bench_rolled_$i.json" 2> "$OUT/bench_rolled_$i.err"
  SMOT_ROI_UNROLL=1 timeout 300 python bench.py $B > "$OUT/bench_unrolled_$i.json" 2> "$OUT/bench_unrolled_$i.err"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 3 "$OUT/pytest_roi.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
grep -h "roi_align" "$OUT/launches_720p30_summary.txt"
