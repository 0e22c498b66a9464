#!/bin/bash
out=gpurun_out/ab_main.log; : > $out
for rep in 1 2; do
for cfg in "" "MB200_COS_MAIN=1" "MB200_COS_MAIN=2" "MB200_COS_MAIN=3" "MB200_COS_MAIN=4"; do
  echo "100k d4 [$cfg]" >> $out
  env $cfg MB200_TRACE=1 python tools/cosine_perf.py --reps 2 --items 100000 --events 6e7 --k 100 2>&1 | grep -E "k3_ms|plan" | tail -n 2 | cut -c1-200 >> $out
done
done
for cfg in "" "MB200_COS_MAIN=1" "MB200_COS_MAIN=2"; do
  echo "100k d1 [$cfg]" >> $out
  env $cfg python tools/cosine_perf.py --reps 2 --items 100000 --events 6e7 --k 100 --depth 1 2>&1 | grep -E "k3_ms" | tail -n 1 | cut -c95-200 >> $out
  echo "c3 [$cfg]" >> $out
  env $cfg python tools/cosine_perf.py --reps 3 2>&1 | grep k3_ms | tail -n 1 | cut -c95-200 >> $out
done
cat $out
