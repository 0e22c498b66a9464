#!/bin/bash
# 8-GPU validation: the driver's own launch of bench.py (configs[3] / configs[4] stages on), then the sharded tests
mkdir -p gpurun_out/r2n8
nvidia-smi -L > gpurun_out/r2n8/gpus.txt; nproc >> gpurun_out/r2n8/gpus.txt; free -g | head -2 >> gpurun_out/r2n8/gpus.txt
t0=$(date +%s)
timeout 880 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 \
   > gpurun_out/r2n8/bench_n8.json 2> gpurun_out/r2n8/bench_n8.err
echo "bench n8 rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r2n8/summary.txt
tail -5 gpurun_out/r2n8/bench_n8.err
timeout 600 python -m pytest tests/test_itemsimilarity_gpu.py tests/test_ingest_gpu.py -x -q -m gpu -k "multi_gpu or sharded or num_gpus" -rs > gpurun_out/r2n8/pytest_8gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2n8/summary.txt
tail -5 gpurun_out/r2n8/pytest_8gpu.log
