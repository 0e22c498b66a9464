#!/usr/bin/env python
"""Stress of the fallbacks and limits of the cosine stage on one GPU: --booleanData (every
preference 1.0: exact tie groups), BF16 rows, k at the fused capacity -- timings, fallback rows and parity against
the oracle on sampled rows, for the MovieLens-100K and MovieLens-20M shapes.  Prints one JSON line per case."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N, synth
from mahout_b200.sketch import last_fallback_rows
from oracle import fast

ctx = mb.Context(0)
dev = torch.device("cuda:0")
shapes = {"ml100k": (943, 1682, 100_000, 1.0), "ml20m": (138_493, 26_744, 20_000_000, 1.1)}
cases = [("ml100k", True, "f16", 100, "rescored"), ("ml100k", True, "f16", 100, "certified"),
         ("ml100k", False, "bf16", 100, "certified"), ("ml100k", False, "f16", 192, "rescored"),
         ("ml20m", True, "f16", 50, "certified"), ("ml20m", True, "f16", 50, "rescored"),
         ("ml20m", False, "bf16", 50, "certified"), ("ml20m", False, "f16", 192, "certified"),
         ("ml20m", False, "f16", 500, "certified")]
for shape, boolean, dtype, k, precision in cases:
    users, items, n, s = shapes[shape]
    cdf = torch.from_numpy(synth.zipf_cdf(items, s)).to(dev)
    perm = torch.from_numpy(synth.rank_permutation(items, 3) - 1).to(dev)
    user, item, pref = synth.events_device(ctx, 20240003, 0, n, users, cdf, perm)
    if boolean:
        pref = torch.ones_like(pref)
    bank = mb.SketchBank(items, 4096, 4, 42, 1, ctx)
    bank.update(item, user, pref)
    bank.check()
    out = {"shape": shape, "boolean_data": boolean, "dtype": dtype, "k": k, "precision": precision}
    try:
        bank.cosine_topk(k, dtype=dtype, precision=precision, device=True)      # warm-up (workspaces)
        ctx.set_profiling(True)
        ctx.reset_profile()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx, sim, cnt = bank.cosine_topk(k, dtype=dtype, precision=precision, device=True)
        torch.cuda.synchronize()
        out.update(ms=1e3 * (time.perf_counter() - t0), k3_ms=ctx.kernel_time(N.K_COSINE)[0],
                   k5_ms=ctx.kernel_time(N.K_RESCORE)[0], fallback_rows=last_fallback_rows(ctx))
        ctx.set_profiling(False)
        rows = np.sort(np.random.Generator(np.random.PCG64(3)).choice(items, 64, replace=False))
        q = bank.read_i32()
        oi, osim, oc = fast.bank_rows_topk(q, rows, k, threads=max(1, len(os.sched_getaffinity(0))))
        gi, gc = idx.cpu().numpy()[rows], cnt.cpu().numpy()[rows]
        out["sets_equal_oracle_on_64_rows"] = bool((gc == oc).all() and all(
            set(gi[r, :gc[r]].tolist()) == set(oi[r, :oc[r]].tolist()) for r in range(64)))
        if precision == "rescored":
            out["sims_bit_equal"] = bool(sim.cpu().numpy()[rows].tobytes() == osim.tobytes())
    except Exception as ex:
        out["error"] = repr(ex)[:200]
    print(json.dumps(out), flush=True)
    bank.close()
    del user, item, pref
    torch.cuda.empty_cache()
