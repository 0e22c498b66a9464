#!/bin/bash
# 1 GPU: band pass diagnostics, the GPU test suite, compute-sanitizer logs, the single-GPU bench line
O=gpurun_out/r2p
mkdir -p $O
t0=$(date +%s)
timeout 300 python tools/band_debug.py > $O/band_debug.log 2>&1
echo "band_debug rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
timeout 600 python -m pytest tests -q -m gpu --tb=short > $O/pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
tail -15 $O/pytest.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py all > $O/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
tail -4 $O/r2_sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py k1 > $O/r2_sanitizer_racecheck_k1.log 2>&1
echo "racecheck rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
tail -4 $O/r2_sanitizer_racecheck_k1.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$? t=$(( $(date +%s) - t0 ))" | tee -a $O/summary.txt
cat $O/band_debug.log
