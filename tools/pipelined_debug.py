"""dev: certified fallback rows, one-shot vs chunked pushes, same data (single GPU)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N, synth
from mahout_b200.sketch import cosine_topk_blocks, CosineJob, last_fallback_rows

items = int(float(sys.argv[1])) if len(sys.argv) > 1 else 60000
events = int(float(sys.argv[2])) if len(sys.argv) > 2 else 120_000_000
k = 100
ctx = mb.Context(0)
dev = torch.device("cuda:0")
cdf = torch.from_numpy(synth.zipf_cdf(items, 1.1)).to(dev)
perm = torch.from_numpy(synth.rank_permutation(items, 4) - 1).to(dev)
user, item, pref = synth.events_device(ctx, 20240006, 0, events, 5_000_000, cdf, perm)
bank = mb.SketchBank(items, 4096, 1, 42, 1, ctx)
bank.update(item, user, pref)
bank.check()
del user, item, pref
rows, valid = bank.normalize("f16")
cnt_t = bank.counters_tensor()
ctx.set_profiling(True)
for prec in ("certified",):
    ctx.reset_profile()
    t0 = time.perf_counter()
    one = cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), 1, 4096, k, b_id=(1, items),
                             precision=prec, a_counters=cnt_t, b_counters=cnt_t)
    torch.cuda.synchronize()
    print("one-shot", prec, "wall ms %.1f" % ((time.perf_counter() - t0) * 1e3), "fallback", last_fallback_rows(ctx),
          "K3 %.1f K5 %.1f" % (ctx.kernel_time(N.K_COSINE)[0], ctx.kernel_time(N.K_RESCORE)[0]), flush=True)
    for chunk in (8192, 2048):
        ctx.reset_profile()
        t0 = time.perf_counter()
        job = CosineJob(ctx, rows, valid, 1, 4096, k, precision=prec)
        for c0 in range(0, items, chunk):
            c1 = min(items, c0 + chunk)
            rc = rows[:, c0:c1].contiguous().unsqueeze(0)
            vw = int(N.lib().mb200_valid_words(c1 - c0))
            vc = torch.zeros((1, 1, vw), dtype=torch.int32, device=dev)
            words = valid[:, c0 // 32:(c1 + 31) // 32]
            vc[0, :, :words.shape[1]] = words
            job.push(rc, vc, id_mul=1, id_add=0, id_base=c0)
        got = job.finish(a_counters=cnt_t, b_counters=cnt_t.unsqueeze(0), b_id=(1, items))
        torch.cuda.synchronize()
        same = all(torch.equal(x, y) for x, y in zip(got, one))
        sets = bool((got[2] == one[2]).all())
        print("chunk", chunk, prec, "wall ms %.1f" % ((time.perf_counter() - t0) * 1e3), "fallback", last_fallback_rows(ctx),
              "K3 %.1f K5 %.1f" % (ctx.kernel_time(N.K_COSINE)[0], ctx.kernel_time(N.K_RESCORE)[0]), "equal one-shot", same, sets, flush=True)
