"""dev: why is the certified step slow inside bench.py but not in cosine_perf.py?"""
import os, sys, time, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N, synth
from mahout_b200.sketch import cosine_topk_blocks

ctx = mb.Context(0)
dev = torch.device("cuda:0")
variant = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "stream" in variant:
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
items, users, n = 26744, 138493, 20_000_000
cdf = torch.from_numpy(synth.zipf_cdf(items, 1.1)).to(dev)
perm = torch.from_numpy(synth.rank_permutation(items, 3) - 1).to(dev)
user, item, pref = synth.events_device(ctx, 20240003, 0, n, users, cdf, perm)
bank = mb.SketchBank(items, 4096, 4, 42, 1, ctx)
bank.update(item, user, pref)
bank.check()
if "hold" in variant:
    ballast = torch.empty(int(40e9), dtype=torch.uint8, device=dev)      # a process holding tens of GB
if "empty" in variant:
    del user, item, pref
    torch.cuda.empty_cache()
rows, valid = bank.normalize("f16")
cnt_t = bank.counters_tensor()
ctx.set_profiling("prof" in variant)
for prec in ("tensor", "certified", "certified", "certified"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), 4, 4096, 50, b_id=(1, items),
                       precision=prec, a_counters=cnt_t, b_counters=cnt_t)
    torch.cuda.synchronize()
    print(variant, prec, "wall ms %.2f" % ((time.perf_counter() - t0) * 1e3), flush=True)
