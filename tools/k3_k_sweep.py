#!/usr/bin/env python
"""K3 time against the number of candidates kept per row (is the sweep bound by its epilogue?).

    python tools/k3_k_sweep.py [--items 125000] [--events 2.5e8] [--depth 1] [--ks 10,50,100,150] [--precision tensor]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=125000)
    ap.add_argument("--events", type=float, default=2.5e8)
    ap.add_argument("--depth", type=int, default=1)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--ks", default="10,50,100,150")
    ap.add_argument("--precision", default="tensor")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--chunks", default="0", help="MB200_COS_CHUNKS values to try (0 = the planner's choice)")
    args = ap.parse_args()
    import torch
    import mahout_b200 as mb
    from mahout_b200 import _native as N
    from mahout_b200 import synth
    E, d, w, n = args.items, args.depth, args.width, int(args.events)
    dev = torch.device("cuda:0")
    ctx = mb.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    cdf = torch.from_numpy(synth.zipf_cdf(E, 1.1)).to(dev)
    perm = torch.from_numpy(synth.rank_permutation(E, 4) - 1).to(dev)
    user, item, pref = synth.events_device(ctx, 20240004, 0, n, 5_000_000, cdf, perm)
    bank = mb.SketchBank(E, w, d, 42, 1, ctx)
    bank.update(item, user, pref)
    del user, item, pref
    flops = 2.0 * d * float(E) ** 2 * int(N.lib().mb200_row_ld(w))
    for k, chunks in [(int(x), int(c)) for x in args.ks.split(",") for c in args.chunks.split(",")]:
        if chunks > 0:
            os.environ["MB200_COS_CHUNKS"] = str(chunks)
        else:
            os.environ.pop("MB200_COS_CHUNKS", None)
        bank.cosine_topk(k, precision=args.precision, device=True)   # warm-up (workspaces)
        ctx.set_profiling(True)
        ctx.reset_profile()
        for _ in range(args.reps):
            bank.cosine_topk(k, precision=args.precision, device=True)
        ms, cnt = ctx.kernel_time(N.K_COSINE)
        ms5, _ = ctx.kernel_time(N.K_RESCORE)
        ctx.set_profiling(False)
        ms /= args.reps
        print(json.dumps({"tool": "k3_k_sweep", "items": E, "depth": d, "k": k, "chunks": chunks, "precision": args.precision,
                          "K3_ms": ms, "K3_launches_per_call": cnt / args.reps, "K5_ms": ms5 / args.reps,
                          "TFLOPs": flops / (ms * 1e-3) / 1e12}), flush=True)
    bank.close()
    ctx.close()


if __name__ == "__main__":
    main()
