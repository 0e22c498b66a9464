#!/bin/bash
# 2-GPU session: sharded paths (routing, fused pull-gather, multi-GPU job behind the C ABI, CLI --numGpus), bench at N=2
mkdir -p gpurun_out/r2g
nvidia-smi -L | tee gpurun_out/r2g/gpus.txt
timeout 1200 python -m pytest tests/test_itemsimilarity_gpu.py tests/test_ingest_gpu.py -x -q -m gpu -rs > gpurun_out/r2g/pytest_2gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2g/summary.txt
tail -8 gpurun_out/r2g/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
   --big on --c4-items 80000 --c4-events 2e8 --c4-check-rows 256 --c4-check-rows-d4 64 \
   --c5-events 8e8 --c5-items 200000 --c5-cos-events 4e8 --c5-check-rows 128 > gpurun_out/r2g/bench_n2_big_small.json 2> gpurun_out/r2g/bench_n2_big_small.err
echo "bench n2 rc=$?" | tee -a gpurun_out/r2g/summary.txt
tail -5 gpurun_out/r2g/bench_n2_big_small.err
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g/bench_n1.json 2> gpurun_out/r2g/bench_n1.err
echo "bench n1 rc=$?" | tee -a gpurun_out/r2g/summary.txt
