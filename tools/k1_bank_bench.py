#!/usr/bin/env python
"""Bank-mode K1 (sketch build into E x d x W counters) timed stage by stage on one GPU.

    python tools/k1_bank_bench.py [--items 125000] [--events 2.5e8] [--calls 1] [--depth 4] [--width 4096]
                                  [--mode grouped|direct|csr] [--reps 3] [--parity-events 2e6]

Shapes: config-4 shard (10^6 items x 2*10^9 events over 8 GPUs -> 125000 items, 2.5*10^8 events per GPU; --calls 4
makes it the >= 10^9-event bank), config 3 (26744 items, 2*10^7 events).  Events are the bench's synthetic stream:
Zipf(1.1) items, uniform users, half-star prefs.  Prints one JSON line: events/s for the whole update and per
stage (K_GROUP = histogram + scans + partition passes, K_UPDATE = the grouped / direct update kernel), the
84-byte-model and the stage-level roofline figures.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=125000)
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--events", type=float, default=2.5e8, help="events per update call")
    ap.add_argument("--calls", type=int, default=1, help="update calls per repetition (each on a fresh slice of the stream)")
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--zipf", type=float, default=1.1)
    ap.add_argument("--mode", default="grouped", choices=["grouped", "direct", "csr"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--parity-events", type=float, default=2e6)
    ap.add_argument("--no-prefetch", action="store_true")
    args = ap.parse_args()

    import torch
    import mahout_b200 as mb
    from mahout_b200 import _native as N
    from mahout_b200 import synth

    n, E, d, w = int(args.events), args.items, args.depth, args.width
    dev = torch.device("cuda:0")
    ctx = mb.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6544.3}
    cdf = torch.from_numpy(synth.zipf_cdf(E, args.zipf)).to(dev)
    perm = torch.from_numpy(synth.rank_permutation(E, 3) - 1).to(dev)
    ctx.set_option(N.OPT_GROUP_MIN_EVENTS, (1 << 62) if args.mode == "direct" else 0)
    if args.no_prefetch:
        ctx.set_option(N.OPT_GROUP_PREFETCH, 0)
    bank = mb.SketchBank(E, w, d, 42, 1, ctx)

    def slice_events(c):
        user, item, pref = synth.events_device(ctx, 20240004, c * n, n, args.users, cdf, perm)
        if args.mode == "csr":
            order = torch.sort(item, stable=True).indices           # plumbing: the caller's grouping
            item, user, pref = item[order], user[order], pref[order]
            rp = torch.searchsorted(item, torch.arange(E + 1, device=dev)).to(torch.int64)
            return rp, user.contiguous(), pref.contiguous()
        return item, user, pref

    ev = [slice_events(c) for c in range(min(args.calls, 2))]      # two resident slices, alternated

    def run():
        for c in range(args.calls):
            a, b, p = ev[c % len(ev)]
            if args.mode == "csr":
                bank.update_grouped(a, b, p)
            else:
                bank.update(a, b, p)

    run()                                                            # warm-up (allocates the workspaces)
    bank.check()
    ctx.sync()
    ctx.set_profiling(True)
    best = None
    for _ in range(args.reps):
        bank.clear()
        ctx.sync()
        ctx.reset_profile()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        g_ms, g_n = ctx.kernel_time(N.K_GROUP)
        u_ms, u_n = ctx.kernel_time(N.K_UPDATE)
        if best is None or ms < best[0]:
            best = (ms, g_ms, g_n, u_ms, u_n)
    ctx.set_profiling(False)
    bank.check()
    ms, g_ms, g_n, u_ms, u_n = best
    total = n * args.calls
    # parity on a prefix through the same path
    parity = None
    m = int(min(args.parity_events, n))
    if m > 0 and args.mode != "csr":
        import oracle as orc
        pe = min(E, 4096)
        a_, u_, p_ = ev[0]
        sel = a_[:m] < pe
        pb = mb.SketchBank(pe, w, d, 42, 1, ctx)
        pb.update(a_[:m][sel].contiguous(), u_[:m][sel].contiguous(), p_[:m][sel].contiguous())
        got = pb.read()
        ha, hb = orc.hash_params(42, d)
        want = np.zeros((pe, d, w))
        orc.bank_update(want, d, w, ha, hb, a_[:m][sel].cpu().numpy(), u_[:m][sel].cpu().numpy(), p_[:m][sel].cpu().numpy())
        parity = {"bit_exact": bool(got.tobytes() == want.tobytes()), "events_checked": int(sel.sum().item())}
        pb.close()
    model = 20 + 16 * d
    line = {
        "tool": "k1_bank_bench", "mode": args.mode, "prefetch": not args.no_prefetch, "items": E, "depth": d, "width": w, "zipf_s": args.zipf,
        "events_per_call": n, "calls": args.calls, "events": total, "bank_GB": E * d * w * 8 / 1e9,
        "ms_total": ms, "events_per_s": total / (ms * 1e-3),
        "ms_group_stage": g_ms, "ms_update_kernel": u_ms, "launch_spans": [int(g_n), int(u_n)],
        "model_bytes_per_event": model,
        "hbm_frac_by_model": model * total / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
        "hbm_frac_by_model_update_kernel_only": (model * total / (u_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if u_ms else None,
        "parity": parity,
    }
    print(json.dumps(line), flush=True)
    bank.close()
    ctx.close()


if __name__ == "__main__":
    main()
