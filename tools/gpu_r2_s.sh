#!/bin/bash
# 4 GPUs: the big stages at per-rank sizes of the 8-GPU run (half the items and events), every stage
O=gpurun_out/r2s
mkdir -p $O
t0=$(date +%s)
MB200_BENCH_DEBUG=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py \
   --gpus 4 --steps 2 --warmup 3 --events 2e8 --e2e-events 16777216 --no-cosine --big on --big-timeout 380 --big-budget 300 \
   --c4-items 500000 --c4-events 1e9 --c4-check-rows 2048 --c4-check-rows-d4 64 --c4-d4 1 --c5-events 5e9 --c5-items 2e6 --c5-cos-events 4e9 --c5-check-rows 256 \
   > $O/bench_n4_emul.json 2> $O/bench_n4_emul.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
grep -E "bench_big|pipelined_cosine rank 0" $O/bench_n4_emul.err | tail -60
