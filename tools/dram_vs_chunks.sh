#!/bin/bash
# K3 DRAM traffic and time against the number of column chunks per row block (config 3)
out=gpurun_out/dram_vs_chunks.log; : > $out
for s in 1 2 4 8 16; do
  echo "chunks=$s" >> $out
  MB200_COS_CHUNKS=$s ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_cosine -c 1 --csv \
     python tools/cosine_perf.py --reps 1 2>/dev/null | grep -E "dram__bytes_read|gpu__time_duration|hit_rate" | cut -d, -f13- >> $out
done
cat $out
