#!/usr/bin/env python
"""Quick single-GPU timing of the cosine stage on a config-3-shaped bank (dev tool)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N
from mahout_b200 import synth
from mahout_b200.sketch import cosine_topk_blocks, last_fallback_rows

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=26744)
ap.add_argument("--users", type=int, default=138493)
ap.add_argument("--events", type=float, default=2e7)
ap.add_argument("--depth", type=int, default=4)
ap.add_argument("--width", type=int, default=4096)
ap.add_argument("--k", type=int, default=50)
ap.add_argument("--block-n", type=int, default=0)
ap.add_argument("--precision", default="tensor")
ap.add_argument("--dtype", default="f16")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--threshold", type=float, default=0.0)
args = ap.parse_args()

ctx = mb.Context(0)
n = int(args.events)
cdf = torch.from_numpy(synth.zipf_cdf(args.items, 1.1)).cuda()
perm = torch.from_numpy(synth.rank_permutation(args.items, 3) - 1).cuda()
user, item, pref = synth.events_device(ctx, 20240003, 0, n, args.users, cdf, perm)
bank = mb.SketchBank(args.items, args.width, args.depth, 42, 1, ctx)
ctx.set_profiling(True)
t0 = time.perf_counter()
bank.update(item, user, pref)
bank.check()
t_upd = time.perf_counter() - t0
rows, valid = bank.normalize(args.dtype)
cnt_t = bank.counters_tensor()
out = {}
for rep in range(args.reps):
    ctx.reset_profile()
    t0 = time.perf_counter()
    idx, sim, cnt = cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), args.depth,
                                       args.width, args.k, b_id=(1, args.items), dtype=args.dtype,
                                       precision=args.precision, block_n=args.block_n, threshold=args.threshold,
                                       a_counters=cnt_t, b_counters=cnt_t)
    wall = time.perf_counter() - t0
    kc, _ = ctx.kernel_time(N.K_COSINE)
    kr, _ = ctx.kernel_time(N.K_RESCORE)
    flops = 2.0 * args.depth * args.items ** 2 * mb._native.lib().mb200_row_ld(args.width)
    out = dict(items=args.items, depth=args.depth, width=args.width, k=args.k, block_n=args.block_n,
               precision=args.precision, wall_s=wall, k3_ms=kc, k5_ms=kr, tflops=flops / kc / 1e9,
               pairs_per_s=args.items ** 2 / (kc * 1e-3), fallback_rows=last_fallback_rows(ctx),
               mean_cnt=float(cnt.float().mean()))
    print(json.dumps(out), flush=True)
ku, nu = ctx.kernel_time(N.K_UPDATE)
print(json.dumps(dict(update_wall_s=t_upd)), flush=True)
