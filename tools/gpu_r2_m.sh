#!/bin/bash
mkdir -p gpurun_out/r2m
timeout 1200 python -m pytest tests/test_cosine_gpu.py tests/test_itemsimilarity_gpu.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2m/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 3 \
   --big on --c4-items 80000 --c4-events 2e8 --c4-check-rows 256 --c4-check-rows-d4 64 \
   --c5-events 8e8 --c5-items 200000 --c5-cos-events 4e8 --c5-check-rows 128 > gpurun_out/r2m/bench_n2_big_small.json 2> gpurun_out/r2m/bench_n2_big_small.err
echo "bench n2 rc=$?" | tee -a gpurun_out/r2m/summary.txt
tail -5 gpurun_out/r2m/bench_n2_big_small.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_update_single_v2" -c 1 -o gpurun_out/r2m/k_update_single_v2 \
  python bench.py --steps 1 --warmup 1 --events 2.5e8 --e2e-events 1e6 --no-cosine > gpurun_out/r2m/ncu_single.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2m/summary.txt
