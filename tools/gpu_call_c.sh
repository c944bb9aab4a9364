#!/bin/bash
# Round-2 GPU call C (1 GPU): frame-pair backbone + cluster split-K (improved finish) validation and A/Bs.
set +e
OUT=gpurun_out/r02c
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 300 python -m pytest tests/test_fp16_e2e_gpu.py -q -m gpu -s > "$OUT/pytest_fp16.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_fp16.txt"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_TC_CLUSTER=0 timeout 300 python bench.py $B > "$OUT/bench_pairs_nocluster.json" 2> "$OUT/bench_pairs_nocluster.err"
SMOT_CLIP_PAIRS=0 timeout 300 python bench.py $B > "$OUT/bench_nopairs_cluster.json" 2> "$OUT/bench_nopairs_cluster.err"
SMOT_CLIP_PAIRS=0 SMOT_TC_CLUSTER=0 timeout 300 python bench.py $B > "$OUT/bench_nopairs_nocluster.json" 2> "$OUT/bench_nopairs_nocluster.err"
timeout 300 python bench.py --steps 20 --warmup 5 --experimental off > "$OUT/bench_default_k20.json" 2> "$OUT/bench_default_k20.err"
timeout 300 python bench.py $B --workload 1080p80 > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 300 python bench.py $B --workload r50_720p30 > "$OUT/bench_r50_720p30.json" 2> "$OUT/bench_r50_720p30.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30_clip.csv" \
    python tools/run_frames.py --frames 6 --eager --clip > "$OUT/ncu_launches_clip.log" 2>&1
tail -n 8 "$OUT/pytest_gpu.txt" "$OUT/pytest_fp16.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "roofline", d["roofline"]["us_per_launch"], d["roofline"]["frac"], "stage", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"])
except Exception as e:
    print("ERR", e)
PY
done
grep -n "conv_tc_kernel<256" "$OUT/launches_720p30_summary.txt" | head -24
tail -n 16 "$OUT/launches_720p30_summary.txt"
