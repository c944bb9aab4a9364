#!/bin/bash
# Round-2 GPU call X (1 GPU): repeatability of the default bench line (three runs in a row).
set +e
OUT=gpurun_out/r02x
mkdir -p "$OUT"
for i in 1 2 3; do
  timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > "$OUT/bench_$i.json" 2> "$OUT/bench_$i.err"
done
timeout 300 python bench.py > "$OUT/bench_full_default_flags.json" 2> "$OUT/bench_full_default_flags.err"
for f in "$OUT"/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "static", d["stage_ms"]["static_graph"], d["spread"]["value_fps"], d["spread"]["e2e_fps"])
except Exception as e:
    print("ERR", e)
PY
done
