#!/bin/bash
# N=2 emulation of the per-rank costs of the 8-GPU big stages (same events, columns and sample rows per rank)
mkdir -p gpurun_out/r2o
t0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 5 --warmup 3 \
   --big on --c4-items 250000 --c4-events 5e8 --c4-check-rows 4096 --c5-events 2.5e9 --c5-items 1e6 --c5-cos-events 2e9 --c5-check-rows 512 \
   > gpurun_out/r2o/bench_n2_emul.json 2> gpurun_out/r2o/bench_n2_emul.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r2o/summary.txt
grep "bench_big" gpurun_out/r2o/bench_n2_emul.err | tail -40
timeout 600 python -m pytest tests/test_cosine_gpu.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2o/pytest.log
