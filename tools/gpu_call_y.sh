#!/bin/bash
# Round-2 GPU call Y (1 GPU): clip pipeline with frame 0 alone ahead of the pairs; the driver's own flags (--steps 20 --warmup 5).
set +e
OUT=gpurun_out/r02y
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m gpu -k "clip or fp16 or pairs or given" > "$OUT/pytest_clip.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_clip.txt"
for i in 1 2; do
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_k20_$i.json" 2> "$OUT/bench_k20_$i.err"
done
SMOT_CLIP_PAIRS=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > "$OUT/bench_k20_nopairs.json" 2> "$OUT/bench_k20_nopairs.err"
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_k100.json" 2> "$OUT/bench_k100.err"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_reference_k20.json" 2> "$OUT/bench_reference_k20.err"
tail -n 3 "$OUT/pytest_clip.txt"
for f in "$OUT"/bench_k*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    sp=d["spread"]["value_fps"]
    print("steps", d["steps"], "value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "n_repeats", len(sp), "value min/max", min(sp), max(sp), d["e2e"]["clip_error"])
except Exception as e:
    print("ERR", e)
PY
done
cut -c1-400 "$OUT/bench_reference_k20.json"
