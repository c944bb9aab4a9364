#!/bin/bash
# Round-2 GPU call S (1 GPU): K slices with 64-channel tiles against the split path, three workloads, same box.
set +e
OUT=gpurun_out/r02s
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fp16_e2e_gpu.py -q -m gpu -k "k_slices or split_k or fp16" > "$OUT/pytest_slices.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_slices.txt"
B="--steps 100 --warmup 10 --no-cpu-baseline"
for W in 720p30 1080p80 r50_720p30; do
  timeout 300 python bench.py $B --workload $W > "$OUT/bench_${W}_sliced64.json" 2> "$OUT/bench_${W}_sliced64.err"
  SMOT_TC_SLICED=0 timeout 300 python bench.py $B --workload $W > "$OUT/bench_${W}_split.json" 2> "$OUT/bench_${W}_split.err"
  SMOT_TC_SLICED=128 timeout 300 python bench.py $B --workload $W > "$OUT/bench_${W}_sliced128.json" 2> "$OUT/bench_${W}_sliced128.err"
done
tail -n 3 "$OUT/pytest_slices.txt"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
