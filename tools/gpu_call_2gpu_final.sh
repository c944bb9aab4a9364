#!/bin/bash
# Final 2-GPU sanity run with the driver's launch line and flags (both arms).
set +e
OUT=gpurun_out/r02_2gpu_final
mkdir -p "$OUT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
echo "rc=$?" >> "$OUT/bench_reference.err"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
    bench.py --gpus 2 --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "rc=$?" >> "$OUT/bench.err"
tail -n 2 "$OUT"/*.err
for f in "$OUT"/bench*.json; do echo "== $f"; grep "^{" "$f" | tail -1 | cut -c1-500; done
