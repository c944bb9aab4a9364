#!/bin/bash
# Round-2 scaling call (gpurun --gpus 8): the driver's launch line at N = 8 and N = 4 ranks (weak scaling: one stream set per GPU).
set +e
OUT=gpurun_out/r02_scale
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
P=29600
for N in "$@"; do
  P=$((P+1))
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline --experimental off > "$OUT/bench_${N}gpu.json" 2> "$OUT/bench_${N}gpu.err"
  echo "rc=$?" >> "$OUT/bench_${N}gpu.err"
done
for f in "$OUT"/bench_*gpu.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    L=[l for l in open(sys.argv[1]) if l.startswith("{")][-1]
    d=json.loads(L)
    print("n_gpus", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "per_frame", d["e2e"]["per_frame_call"]["value"], d["e2e"]["clip_error"])
    print("   per_rank value_region ms", [r[0] for r in d["per_rank_ms"]["rows"]], "gathered", d["per_rank_ms"]["gathered_tracks_per_rank"], "clocks", d.get("clocks"))
    print("   notes", {k:v for k,v in d["notes"].items() if "numa" in k})
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
