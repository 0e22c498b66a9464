#!/bin/bash
mkdir -p gpurun_out/r2h
for v in plain stream stream_prof stream_prof_empty stream_prof_hold; do
  MB200_TRACE=1 timeout 300 python tools/cert_debug.py $v >> gpurun_out/r2h/cert_debug.log 2>&1
done
grep -E "wall ms|trace" gpurun_out/r2h/cert_debug.log | tail -150
