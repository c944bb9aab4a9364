#!/bin/bash
# Round-2 GPU call Q (1 GPU): compute-sanitizer over the kernels written this round (memcheck + racecheck + synccheck).
set +e
OUT=gpurun_out/r02q
mkdir -p "$OUT"
SEL="hires or roi_align or xcorr_planar or planar"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_more_gpu.py -q -m gpu -x -k "$SEL" > "$OUT/memcheck.txt" 2>&1
echo "rc=$?" >> "$OUT/memcheck.txt"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_more_gpu.py -q -m gpu -x -k "hires_persistent or xcorr_planar_equals or roi_align_planar" > "$OUT/racecheck.txt" 2>&1
echo "rc=$?" >> "$OUT/racecheck.txt"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_more_gpu.py -q -m gpu -x -k "hires_persistent or xcorr_planar_equals" > "$OUT/synccheck.txt" 2>&1
echo "rc=$?" >> "$OUT/synccheck.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/run_frames.py --frames 3 --eager > "$OUT/memcheck_frames.txt" 2>&1
echo "rc=$?" >> "$OUT/memcheck_frames.txt"
for f in memcheck racecheck synccheck memcheck_frames; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=|Invalid|hazard" "$OUT/$f.txt" | sort | uniq -c | head -12; done
