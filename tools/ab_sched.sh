#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t_c.log 2>&1; tail -n 3 gpurun_out/t_c.log
out=gpurun_out/ab_sched.log; : > $out
for rep in 1 2; do
for cfg in "" "MB200_COS_CHUNKS=1" "MB200_COS_CHUNKS=3"; do
  echo "100k [$cfg]" >> $out
  env $cfg python tools/cosine_perf.py --reps 2 --items 100000 --events 6e7 --k 100 2>&1 | grep k3_ms | tail -n 1 | cut -c95-300 >> $out
done
for cfg in "" "MB200_COS_CHUNKS=2" "MB200_COS_CHUNKS=1"; do
  echo "c3 [$cfg]" >> $out
  env $cfg python tools/cosine_perf.py --reps 3 2>&1 | grep k3_ms | tail -n 1 | cut -c95-300 >> $out
done
done
cat $out
