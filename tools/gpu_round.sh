#!/bin/bash
# dev helper: one GPU visit = tests + K3 timings (+ optional extras); logs under gpurun_out/
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$tag.log 2>&1; tail -n 3 gpurun_out/t_$tag.log
python tools/cosine_perf.py --reps 3 > gpurun_out/cp_$tag.log 2>&1
python tools/cosine_perf.py --reps 2 --items 100000 --events 6e7 --k 100 > gpurun_out/cp100k_$tag.log 2>&1
MB200_TRACE=1 python tools/cosine_perf.py --reps 2 --precision rescored > gpurun_out/cpresc_$tag.log 2>&1
for f in gpurun_out/cp_$tag.log gpurun_out/cp100k_$tag.log gpurun_out/cpresc_$tag.log; do echo "== $f"; cut -c1-330 $f | tail -n 24; done
