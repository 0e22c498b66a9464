#!/bin/bash
mkdir -p gpurun_out/r2e
timeout 900 python bench.py --steps 3 --warmup 3 --big on --c4-items 60000 --c4-events 1e8 --c4-check-rows 256 --c4-check-rows-d4 64 \
   --c5-events 4e8 --c5-items 100000 --c5-cos-events 2e8 --c5-check-rows 128 > gpurun_out/r2e/bench_n1_big_small.json 2> gpurun_out/r2e/bench_n1_big_small.err
echo "bench big-small rc=$?" | tee -a gpurun_out/r2e/summary.txt
tail -5 gpurun_out/r2e/bench_n1_big_small.err
timeout 900 python bench.py > gpurun_out/r2e/bench_n1.json 2> gpurun_out/r2e/bench_n1.err
echo "bench default rc=$?" | tee -a gpurun_out/r2e/summary.txt
tail -5 gpurun_out/r2e/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e/bench_ref.json 2> gpurun_out/r2e/bench_ref.err
echo "bench ref rc=$?" | tee -a gpurun_out/r2e/summary.txt
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2e/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2e/summary.txt
tail -5 gpurun_out/r2e/pytest_all.log
