#!/bin/bash
# Round-2 GPU call T (1 GPU): K slices only where the split is 8 wide (default) vs everywhere vs never, same box.
set +e
OUT=gpurun_out/r02t
mkdir -p "$OUT"
B="--steps 100 --warmup 10 --no-cpu-baseline"
for W in 720p30 1080p80 r50_720p30; do
  timeout 300 python bench.py $B --workload $W > "$OUT/bench_${W}_min8.json" 2> "$OUT/bench_${W}_min8.err"
  SMOT_TC_SLICED=0 timeout 300 python bench.py $B --workload $W > "$OUT/bench_${W}_split.json" 2> "$OUT/bench_${W}_split.err"
done
SMOT_TC_SLICED_MIN=2 timeout 300 python bench.py $B > "$OUT/bench_720p30_min2.json" 2> "$OUT/bench_720p30_min2.err"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["e2e"]["clip_error"], d["spread"]["value_fps"])
except Exception as e:
    print("ERR", e)
PY
done
