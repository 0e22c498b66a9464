#!/bin/bash
# dev sweep: K3 time vs column chunks / L2 hints at a given item count
items=${1:-100000}; events=${2:-6e7}; out=gpurun_out/sweep_${items}.log; : > $out
for h in 0 3 1; do for s in 1 3 5 7 9 12 16; do
  echo "hints=$h chunks=$s" >> $out
  MB200_COS_HINTS=$h MB200_COS_CHUNKS=$s python tools/cosine_perf.py --reps 2 --items $items --events $events --k 100 2>&1 | grep k3_ms | tail -n 1 | cut -c1-300 >> $out
done; done
cat $out
