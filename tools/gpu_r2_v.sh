#!/bin/bash
# 2 GPUs: the 2-GPU tests + traced certified finish of a configs[3]-shaped shard with local copies of the peers' int32 banks
O=gpurun_out/r2v
mkdir -p $O
t0=$(date +%s)
timeout 300 python -m pytest tests/test_itemsimilarity_gpu.py -q -m gpu --tb=short 2>&1 | tail -8 | tee $O/pytest_2gpu.log
echo "pytest t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
MB200_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py \
   --gpus 2 --steps 1 --warmup 3 --events 1e8 --e2e-events 4194304 --no-cosine --big on --big-timeout 280 --big-budget 250 \
   --c4-items 250000 --c4-events 5e8 --c4-check-rows 512 --c4-check-rows-d4 32 --c4-d4 1 --c5-events 0 --c5-items 0 \
   > $O/bench.json 2> $O/bench.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
grep -E "bench_big|mb200 trace\] (K3|rescore|band)" $O/bench.err | tail -60
