#!/bin/bash
mkdir -p gpurun_out/r2i
MB200_BENCH_DEBUG=1 MB200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --events 2e8 --e2e-events 3e7 > gpurun_out/r2i/bench_dbg.json 2> gpurun_out/r2i/bench_dbg.err
echo "bench rc=$?"
grep -n "certified" gpurun_out/r2i/bench_dbg.err | head -5
grep -E "bench debug|trace" gpurun_out/r2i/bench_dbg.err | grep -B12 -A12 "certified" | head -150
timeout 600 python tools/pipelined_debug.py 60000 1.2e8 2>&1 | tee gpurun_out/r2i/pipelined_debug.log | grep -v trace
