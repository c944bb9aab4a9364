#!/bin/bash
# Final 8-GPU run with the driver's launch line and flags.
set +e
OUT=gpurun_out/r02_8gpu_final
mkdir -p "$OUT"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 \
    bench.py --gpus 8 --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "rc=$?" >> "$OUT/bench.err"
tail -n 2 "$OUT/bench.err"
grep "^{" "$OUT/bench.json" | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'per_frame', d['e2e']['per_frame_call']['value'], [r[0] for r in d['per_rank_ms']['rows']], d['e2e']['clip_error'])"
