#!/bin/bash
# First GPU call after a stretch of work without one (run under gpurun from the repo root, ~12-15 min on one B200):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_checkin.sh'
# Everything lands in gpurun_out/checkin/.  Order: the validated suite first (a regression there is the first thing to know),
# then the pending cases verbosely (XPASS = ready to be promoted out of tests/test_zz_pending_gpu.py), then the bench lines
# (default path, planar exchange on, the other BASELINE workloads), then the ncu passes of B200_PROFILING.md.
set -u
OUT=gpurun_out/checkin
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
python - <<'PY' > "$OUT/env.txt" 2>&1
import torch, subprocess
print(torch.__version__, torch.cuda.get_device_name(0))
print(subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,power.draw", "--format=csv"], capture_output=True, text=True).stdout)
PY
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_zz_pending_gpu.py -x > "$OUT/pytest_validated.txt" 2>&1
echo "validated suite rc=$?" >> "$OUT/pytest_validated.txt"
timeout 600 python -m pytest tests/test_zz_pending_gpu.py -m gpu -rxX -v > "$OUT/pytest_pending.txt" 2>&1
echo "pending suite rc=$?" >> "$OUT/pytest_pending.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.txt" 2>&1
# the three-stage clip pipeline (developer switch): equality with frame-by-frame calls first, then its bench lines
for K in 2 3; do
  SMOT_CLIP_SPLIT=1 SMOT_CLIP_SLOTS=$K timeout 300 python -m pytest tests/test_e2e_gpu.py tests/test_preprocess_gpu.py -m gpu -q \
      -k "forward_clip or raw_frames" > "$OUT/pytest_clip_split_k$K.txt" 2>&1
  echo "rc=$?" >> "$OUT/pytest_clip_split_k$K.txt"
  SMOT_CLIP_SPLIT=1 SMOT_CLIP_SLOTS=$K timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --experimental off \
      > "$OUT/bench_clip_split_k$K.json" 2> "$OUT/bench_clip_split_k$K.err"
done
timeout 400 python bench.py --steps 200 --warmup 10 > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
SMOT_XCORR_PLANAR=1 timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --experimental off > "$OUT/bench_planar.json" 2> "$OUT/bench_planar.err"
SMOT_XCORR_PLANAR=2 timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --experimental off > "$OUT/bench_planar_trimmed.json" 2> "$OUT/bench_planar_trimmed.err"
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --experimental off --workload 1080p80 > "$OUT/bench_1080p80.json" 2> "$OUT/bench_1080p80.err"
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --experimental off --workload r50_720p30 > "$OUT/bench_r50.json" 2> "$OUT/bench_r50.err"
SMOT_FRAME_OVERLAP=1 timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --experimental off > "$OUT/bench_frame_overlap.json" 2> "$OUT/bench_frame_overlap.err"
SMOT_BODY_BRANCHES=1 timeout 400 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --experimental off > "$OUT/bench_body_branches.json" 2> "$OUT/bench_body_branches.err"
# launch lists (cold-cache, serialised: compare shares), default and planar
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_default.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --experimental off > "$OUT/ncu_default.log" 2>&1
SMOT_XCORR_PLANAR=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file "$OUT/launches_planar.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --experimental off > "$OUT/ncu_planar.log" 2>&1
# one full capture of each roofline-kernel candidate
SMOT_XCORR_PLANAR=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:xcorr_planar_kernel -c 1 -s 3 \
    -o "$OUT/xcorr_planar" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --experimental off > "$OUT/ncu_xcorr_planar.log" 2>&1
SMOT_XCORR_PLANAR=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:roi_align_planar_kernel -c 1 -s 3 \
    -o "$OUT/roi_align_planar" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --experimental off > "$OUT/ncu_roi_align_planar.log" 2>&1
# memcheck over the kernels that have never run on a GPU (small cases; the sanitizer is 10-50x slower)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_pending_gpu.py -m gpu -q --runxfail -x \
    -k "maxpool3x3s2 or track_combine_grouped or roi_align_planar or (xcorr_planar and 3-32)" > "$OUT/sanitizer_memcheck.txt" 2>&1
echo "memcheck rc=$?" >> "$OUT/sanitizer_memcheck.txt"
for f in "$OUT"/launches_*.csv; do python tools/launch_report.py "$f" > "${f%.csv}_summary.txt" 2>&1; done
tail -n 3 "$OUT"/pytest_validated.txt "$OUT"/pytest_pending.txt "$OUT"/smoke.txt
tail -n 2 "$OUT"/pytest_clip_split_k*.txt
cat "$OUT"/bench_clip_split_k2.json "$OUT"/bench_clip_split_k3.json "$OUT"/bench_default.json "$OUT"/bench_planar.json "$OUT"/bench_1080p80.json "$OUT"/bench_r50.json 2>/dev/null | cut -c 1-600
