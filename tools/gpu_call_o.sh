#!/bin/bash
# Round-2 GPU call O (1 GPU): widest split of the few-tile tcgen05 layers under the pipeline (SMOT_TC_MAXSPLIT), parity under the candidates.
set +e
OUT=gpurun_out/r02o
mkdir -p "$OUT"
B="--steps 100 --warmup 10 --no-cpu-baseline"
for M in 8 4 2 1; do
  SMOT_TC_MAXSPLIT=$M timeout 300 python bench.py $B > "$OUT/bench_maxsplit$M.json" 2> "$OUT/bench_maxsplit$M.err"
done
SMOT_TC_MAXSPLIT=1 timeout 300 python bench.py $B --workload 1080p80 > "$OUT/bench_1080p80_maxsplit1.json" 2> "$OUT/bench_1080p80_maxsplit1.err"
SMOT_TC_MAXSPLIT=1 timeout 300 python bench.py $B --workload r50_720p30 > "$OUT/bench_r50_maxsplit1.json" 2> "$OUT/bench_r50_maxsplit1.err"
SMOT_TC_MAXSPLIT=2 timeout 300 python bench.py $B --workload r50_720p30 > "$OUT/bench_r50_maxsplit2.json" 2> "$OUT/bench_r50_maxsplit2.err"
SMOT_TC_MAXSPLIT=1 timeout 900 python -m pytest tests -q -m gpu > "$OUT/pytest_gpu_maxsplit1.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu_maxsplit1.txt"
SMOT_TC_MAXSPLIT=2 timeout 600 python -m pytest tests/test_fp16_e2e_gpu.py tests/test_e2e_gpu.py tests/test_more_gpu.py -q -m gpu -s > "$OUT/pytest_sel_maxsplit2.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_sel_maxsplit2.txt"
tail -n 4 "$OUT/pytest_gpu_maxsplit1.txt"; grep -E "float16:|float32:|passed|failed" "$OUT/pytest_sel_maxsplit2.txt" | tail -8
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
