#!/bin/bash
# 1 GPU: candidate margin of the certified precision (k + 50 -> 160 kept, k + 92 -> 192 kept) on 500000 items, depth 1
O=gpurun_out/r2w2
mkdir -p $O
ARGS="--gpus 1 --steps 1 --warmup 3 --events 1e8 --e2e-events 4194304 --no-cosine --big on --c4-items 500000 --c4-events 1e9 --c4-check-rows 64 --c4-d4 0 --c5-events 0 --c5-items 0"
for m in 50 92; do
  MB200_MARGIN=$m MB200_TRACE=1 timeout 200 python bench.py $ARGS > $O/bench_margin_$m.json 2> $O/bench_margin_$m.err
  echo "margin $m rc=$?" | tee -a $O/summary.txt
  grep -E "bench_big.*cosine step|mb200 trace\] (K3|rescore|band)" $O/bench_margin_$m.err | tail -12
  python - <<PY
import json
l=json.loads(open('$O/bench_margin_$m.json').read().strip().splitlines()[-1])
c=l['config4']
print('margin',$m,'ms',c.get('ms_per_step'),'band',c.get('band_rows'),'fb',c.get('certified_fallback_rows'),c.get('kernels_ms_per_step_this_rank'),c.get('parity',{}).get('certified_topk_sets_equal_oracle'), c.get('error'))
PY
done
