#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): the driver's own launch line for both arms, at N ranks.
set +e
N=${1:-2}
OUT=gpurun_out/r02_${N}gpu
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
echo "rc=$?" >> "$OUT/bench_reference.err"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/bench_k20.json" 2> "$OUT/bench_k20.err"
echo "rc=$?" >> "$OUT/bench_k20.err"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 100 --warmup 10 > "$OUT/bench_k100.json" 2> "$OUT/bench_k100.err"
echo "rc=$?" >> "$OUT/bench_k100.err"
timeout 300 python bench.py --steps 100 --warmup 10 --experimental off --no-cpu-baseline > "$OUT/bench_1gpu_same_box.json" 2> "$OUT/bench_1gpu_same_box.err"
tail -n 3 "$OUT"/*.err
for f in "$OUT"/bench_*.json; do echo "== $f"; cut -c 1-700 "$f"; done
