#!/bin/bash
mkdir -p gpurun_out/r2d
timeout 900 python -m pytest tests/test_sketch_gpu.py -x -q -m gpu > gpurun_out/r2d/pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2d/summary.txt
tail -3 gpurun_out/r2d/pytest.log
for cfg in "26744 2e7 grouped" "125000 2.5e8 grouped" "125000 2.5e8 grouped --no-prefetch" "125000 1e9 grouped" "125000 1e9 grouped --no-prefetch" "125000 1e9 csr"; do
  set -- $cfg
  timeout 600 python tools/k1_bank_bench.py --items $1 --events $2 --mode $3 $4 --parity-events 0 >> gpurun_out/r2d/k1_bank.jsonl 2>> gpurun_out/r2d/k1_bank.err
  echo "k1 $cfg rc=$?" | tee -a gpurun_out/r2d/summary.txt
done
cat gpurun_out/r2d/k1_bank.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_group|k_update_grouped|k_scan|k_tile|k_window" -c 20 --csv \
  --log-file gpurun_out/r2d/launches_k1_grouped.csv python tools/k1_bank_bench.py --items 125000 --events 2.5e8 --mode grouped --reps 1 --parity-events 0 \
  > gpurun_out/r2d/ncu_k1.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2d/summary.txt
