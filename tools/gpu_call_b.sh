#!/bin/bash
# Round-2 GPU call B (1 GPU): cluster split-K validation, fp16 e2e parity tests, pipeline A/Bs.
set +e
OUT=gpurun_out/r02b
mkdir -p "$OUT"
timeout 120 python -m pytest tests/test_e2e_gpu.py -q -m gpu -k "forward_clip_equals" > "$OUT/pytest_repro.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_repro.txt"
SMOT_FRAME_OVERLAP=0 timeout 120 python -m pytest tests/test_e2e_gpu.py -q -m gpu -k "forward_clip_equals" > "$OUT/pytest_repro_nooverlap.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_repro_nooverlap.txt"
timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 300 python -m pytest tests/test_fp16_e2e_gpu.py -q -m gpu -s > "$OUT/pytest_fp16.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_fp16.txt"
timeout 120 python __graft_entry__.py smoke > "$OUT/smoke.txt" 2>&1
echo "rc=$?" >> "$OUT/smoke.txt"
B="--steps 100 --warmup 10 --experimental off --no-cpu-baseline"
timeout 300 python bench.py $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_TC_CLUSTER=0 timeout 300 python bench.py $B > "$OUT/bench_nocluster.json" 2> "$OUT/bench_nocluster.err"
SMOT_CLIP_BACKBONE_STREAMS=2 SMOT_CLIP_SLOTS=3 timeout 300 python bench.py $B > "$OUT/bench_bb2_k3.json" 2> "$OUT/bench_bb2_k3.err"
SMOT_CLIP_BACKBONE_STREAMS=2 SMOT_CLIP_SLOTS=4 timeout 300 python bench.py $B > "$OUT/bench_bb2_k4.json" 2> "$OUT/bench_bb2_k4.err"
SMOT_CLIP_SLOTS=4 timeout 300 python bench.py $B > "$OUT/bench_k4.json" 2> "$OUT/bench_k4.err"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 6 "$OUT/pytest_repro.txt" "$OUT/pytest_repro_nooverlap.txt" "$OUT/pytest_gpu.txt" "$OUT/pytest_fp16.txt" "$OUT/smoke.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; cut -c 1-400 "$f"; done
tail -n 22 "$OUT/launches_720p30_summary.txt"
