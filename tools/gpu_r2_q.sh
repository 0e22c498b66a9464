#!/bin/bash
# 2 GPUs: the per-rank sizes of the 8-GPU big stages (configs[3] at depth 1 and the streamed configs[4] cosine)
O=gpurun_out/r2r
mkdir -p $O
t0=$(date +%s)
MB200_BENCH_DEBUG=1 timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py \
   --gpus 2 --steps 2 --warmup 3 --events 2e8 --e2e-events 16777216 --no-cosine --big on --big-timeout 400 --big-budget 330 \
   --c4-items 250000 --c4-events 5e8 --c4-check-rows 1024 --c4-d4 0 --c5-events 0 --c5-items 1e6 --c5-cos-events 2e9 --c5-check-rows 256 \
   > $O/bench_n2_emul.json 2> $O/bench_n2_emul.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" | tee -a $O/summary.txt
grep -E "bench_big|pipelined_cosine" $O/bench_n2_emul.err | tail -60
timeout 200 python -m pytest tests/test_cosine_gpu.py -q -m gpu --tb=short -k "bad_arguments or band_pass or beyond" 2>&1 | tail -5 | tee $O/pytest.log
