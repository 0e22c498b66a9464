#!/usr/bin/env python
"""dev tool: throughput of the GPU ingest (parse + prepare) on MovieLens-20M-shaped CSV text"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N, ingest, synth

ap = argparse.ArgumentParser()
ap.add_argument("--events", type=float, default=2e7)
ap.add_argument("--users", type=int, default=138493)
ap.add_argument("--items", type=int, default=26744)
args = ap.parse_args()
n = int(args.events)
cdf = synth.zipf_cdf(args.items, 1.1)
user, item, pref = synth.events_numpy(20240003, 0, n, args.users, cdf, synth.rank_permutation(args.items, 3))
t0 = time.perf_counter()
# "user,item,pref,timestamp" like ratings.csv
import io
cols = np.char.add(np.char.add(np.char.add(user.astype(str), ","), np.char.add(item.astype(str), ",")),
                   np.char.add(np.char.mod("%.1f", pref), ",1112486027\n"))
text = "".join(cols.tolist()).encode()
gen_s = time.perf_counter() - t0
ctx = mb.Context(0)
ctx.set_profiling(True)
dtext = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
out = {"events": n, "text_bytes": len(text), "gen_s": gen_s}
for rep in range(3):
    ctx.reset_profile()
    t0 = time.perf_counter()
    ev = ingest.Events.parse(dtext, ctx=ctx)
    t1 = time.perf_counter()
    pm = ev.prepare(1)
    t2 = time.perf_counter()
    kp, _ = ctx.kernel_time(N.K_PARSE)
    kq, _ = ctx.kernel_time(N.K_PREPARE)
    out.update(parse_wall_ms=(t1 - t0) * 1e3, prepare_wall_ms=(t2 - t1) * 1e3, parse_kernels_ms=kp, prepare_kernels_ms=kq,
               parse_GBps=(len(text) * 2 + 20 * n) / kp / 1e6, parse_events_per_s=n / (kp * 1e-3),
               prepare_events_per_s=n / (kq * 1e-3), survivors=pm.n, num_items=pm.num_items, num_users=pm.num_users)
    pm.close(); ev.close()
t0 = time.perf_counter()
ev = ingest.Events.parse(text, ctx=ctx)
out["parse_from_host_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out))
