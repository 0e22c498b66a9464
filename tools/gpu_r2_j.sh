#!/bin/bash
mkdir -p gpurun_out/r2j
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pipelined_debug2.py 100000 2e8 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tee gpurun_out/r2j/pipelined_debug2.log
for sk in 0 1 2; do
python - <<PY 2>&1 | tee -a gpurun_out/r2j/k1_single_variants.log
import torch, time, json
import mahout_b200 as mb
from mahout_b200 import _native as N, synth
ctx = mb.Context(0)
ctx.set_option(N.OPT_SINGLE_KERNEL, $sk)
n = 1_000_000_000
for s in (1.1, 1.5):
    cdf = torch.from_numpy(synth.zipf_cdf(10_000_000, s)).cuda()
    _, item, pref = synth.events_device(ctx, 20240002, 0, n, 1_000_000, cdf, None, want_user=False)
    bank = mb.SketchBank(1, 1 << 20, 4, 42, 1, ctx)
    ctx.set_profiling(True)
    for _ in range(3): bank.update(None, item, pref)
    ctx.reset_profile()
    for _ in range(5): bank.update(None, item, pref)
    ms, k = ctx.kernel_time(N.K_UPDATE)
    bank.check()
    print(json.dumps({"single_kernel": $sk, "zipf": s, "ms": ms / k, "G_events_per_s": n / (ms / k) / 1e6}), flush=True)
    bank.close(); del item, pref
PY
done
timeout 600 python -m pytest tests/test_sketch_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/r2j/pytest.log
