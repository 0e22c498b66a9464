#!/bin/bash
# 2 GPUs: host-side trace (MB200_TRACE) of the certified finish of a configs[3]-shaped shard, depth 1 and 4
O=gpurun_out/r2u
mkdir -p $O
t0=$(date +%s)
MB200_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py \
   --gpus 2 --steps 1 --warmup 3 --events 1e8 --e2e-events 4194304 --no-cosine --big on --big-timeout 280 --big-budget 250 \
   --c4-items 250000 --c4-events 5e8 --c4-check-rows 64 --c4-check-rows-d4 16 --c4-d4 1 --c5-events 0 --c5-items 0 \
   > $O/bench.json 2> $O/bench.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
grep -E "bench_big|mb200 trace" $O/bench.err | tail -120
