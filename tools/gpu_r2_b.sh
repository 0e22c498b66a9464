#!/bin/bash
# round-2 GPU session B: grouped K1 after the store-only flush / big sub-batch; ncu --set full of the four kernels
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_sketch_gpu.py tests/test_cosine_gpu.py -x -q -m gpu > gpurun_out/r2b/pytest_sketch.log 2>&1
echo "pytest_sketch rc=$?" | tee -a gpurun_out/r2b/summary.txt
tail -3 gpurun_out/r2b/pytest_sketch.log
for cfg in "26744 2e7 1 grouped" "125000 2.5e8 1 grouped" "125000 2.5e8 1 csr" "125000 1e9 1 grouped" "125000 1e9 1 csr" "125000 1e9 1 direct"; do
  set -- $cfg
  timeout 600 python tools/k1_bank_bench.py --items $1 --events $2 --calls $3 --mode $4 --parity-events 0 >> gpurun_out/r2b/k1_bank.jsonl 2>> gpurun_out/r2b/k1_bank.err
  echo "k1 $cfg rc=$?" | tee -a gpurun_out/r2b/summary.txt
done
cat gpurun_out/r2b/k1_bank.jsonl
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"k_group_hist|k_group_scatter|k_update_grouped" -c 4 \
  -o gpurun_out/r2b/k1_grouped_full python tools/k1_bank_bench.py --items 60000 --events 1e8 --mode grouped --reps 1 --parity-events 0 \
  > gpurun_out/r2b/ncu_full.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/r2b/summary.txt
ls -la gpurun_out/r2b
