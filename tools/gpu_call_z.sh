#!/bin/bash
# Round-2 GPU call Z (1 GPU): what the driver runs at round end, in its order -- GPU tests, smoke, reference arm, our arm (its flags).
set +e
OUT=gpurun_out/r02z
mkdir -p "$OUT"
timeout 1200 python -m pytest tests/ -x -q -m gpu > "$OUT/pytest_gpu.txt" 2>&1
echo "rc=$?" >> "$OUT/pytest_gpu.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.txt" 2>&1
echo "rc=$?" >> "$OUT/smoke.txt"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -n 3 "$OUT/pytest_gpu.txt"; tail -n 2 "$OUT/smoke.txt"
cut -c1-300 "$OUT/bench_reference.json"; echo
python - "$OUT/bench.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"], "clocks", d["clocks"], "launches", d["gpu_launches"])
PY
