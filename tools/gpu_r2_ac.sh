#!/bin/bash
# 2 GPUs, final code: 2-GPU tests + the per-rank sizes of the 8-GPU big stages (configs[3] at depth 1 and 4, streamed configs[4] cosine)
O=gpurun_out/r2ac
mkdir -p $O
t0=$(date +%s)
timeout 200 python -m pytest tests/test_itemsimilarity_gpu.py tests/test_ingest_gpu.py -q -m gpu --tb=short 2>&1 | tail -5 | tee $O/pytest_2gpu.log
MB200_BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py \
   --gpus 2 --steps 2 --warmup 3 --events 2e8 --e2e-events 16777216 --no-cosine --big on --big-timeout 380 --big-budget 330 \
   --c4-items 250000 --c4-events 5e8 --c4-check-rows 1024 --c4-check-rows-d4 32 --c4-d4 1 --c5-events 0 --c5-items 1e6 --c5-cos-events 2e9 --c5-check-rows 256 \
   > $O/bench_n2_emul.json 2> $O/bench_n2_emul.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
grep -E "bench_big|pipelined_cosine rank 0" $O/bench_n2_emul.err | tail -40
