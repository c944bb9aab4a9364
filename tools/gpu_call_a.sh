#!/bin/bash
# Round-2 GPU call A (1 GPU): validate the flipped defaults, take the first full set of numbers, probe fp16 parity scenes.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call_a.sh'
set +e
OUT=gpurun_out/r02a
mkdir -p "$OUT"
nvidia-smi > "$OUT/smi.txt" 2>&1
nproc > "$OUT/nproc.txt"
timeout 600 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 120 python __graft_entry__.py smoke > "$OUT/smoke.txt" 2>&1
echo "rc=$?" >> "$OUT/smoke.txt"
timeout 500 python tools/parity_probe.py --variants "$PROBE_VARIANTS" --frames 8 --calibrate --out "$OUT/parity_probe.json" > "$OUT/parity_probe.log" 2>&1
echo "rc=$?" >> "$OUT/parity_probe.log"
timeout 400 python bench.py --steps 100 --warmup 10 --experimental off > "$OUT/bench_720p30.json" 2> "$OUT/bench_720p30.err"
timeout 300 python bench.py --steps 20 --warmup 5 --experimental off --no-cpu-baseline > "$OUT/bench_720p30_k20.json" 2> "$OUT/bench_720p30_k20.err"
SMOT_STREAM_PRIORITY=0 timeout 300 python bench.py --steps 100 --warmup 10 --experimental off --no-cpu-baseline > "$OUT/bench_720p30_noprio.json" 2> "$OUT/bench_720p30_noprio.err"
SMOT_CLIP_SPLIT=0 timeout 300 python bench.py --steps 100 --warmup 10 --experimental off --no-cpu-baseline > "$OUT/bench_720p30_twostream.json" 2> "$OUT/bench_720p30_twostream.err"
timeout 400 python bench.py --steps 50 --warmup 5 --experimental off --workload 1080p80 > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 400 python bench.py --steps 50 --warmup 5 --experimental off --workload r50_720p30 > "$OUT/bench_r50_720p30.json" 2> "$OUT/bench_r50_720p30.err"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
# launch list (cold-cache, serialised: compare shares), eager launches of 3 natural frames
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/launches_720p30.csv" \
    python tools/run_frames.py --frames 3 --eager > "$OUT/ncu_launches.log" 2>&1
# one full capture of the roofline kernel and of its producer
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xcorr_planar_kernel -s 2 -c 1 -f -o "$OUT/xcorr_planar" \
    python tools/run_frames.py --frames 4 --eager > "$OUT/ncu_xcorr_planar.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_planar_kernel -s 2 -c 1 -f -o "$OUT/roi_align_planar" \
    python tools/run_frames.py --frames 4 --eager > "$OUT/ncu_roi_align_planar.log" 2>&1
python tools/launch_report.py "$OUT/launches_720p30.csv" > "$OUT/launches_720p30_summary.txt" 2>&1
tail -n 4 "$OUT/pytest_gpu.txt" "$OUT/smoke.txt"
tail -c 1500 "$OUT/parity_probe.log"
for f in "$OUT"/bench_*.json; do echo "== $f"; cut -c 1-900 "$f"; done
tail -n 20 "$OUT/launches_720p30_summary.txt"
