#!/usr/bin/env python
"""Band pass on flat-similarity rows, with the differences from the oracle printed (development aid).

    python tools/band_debug.py [--items 1800,3000] [--precisions certified,rescored]
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
    ap.add_argument("--items", default="1800,3000")
    ap.add_argument("--precisions", default="certified,rescored")
    ap.add_argument("--k", type=int, default=100)
    args = ap.parse_args()
    import mahout_b200 as mb
    import oracle as orc
    from mahout_b200.sketch import last_band_rows, last_fallback_rows
    ctx = mb.Context(0)
    for E in [int(x) for x in args.items.split(",")]:
        rng = np.random.Generator(np.random.PCG64(5))
        d, w, k = 2, 128, args.k
        n = 60 * E
        item = rng.integers(0, E, n).astype(np.int64)
        user = rng.integers(1, 300, n).astype(np.int64)
        pref = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
        bank = mb.SketchBank(E, w, d, 42, 1, ctx)
        bank.update(item, user, pref)
        a, b = orc.hash_params(42, d)
        ref = np.zeros((E, d, w))
        orc.bank_update(ref, d, w, a, b, item, user, pref)
        oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
        for precision in args.precisions.split(","):
            idx, sim, cnt = bank.cosine_topk(k, dtype="bf16", precision=precision)
            bd, fb = last_band_rows(ctx), last_fallback_rows(ctx)
            bad_cnt = np.nonzero(cnt != ocnt)[0]
            bad_set = [r for r in range(E) if set(idx[r, :cnt[r]].tolist()) != set(oidx[r, :ocnt[r]].tolist())]
            bad_idx = np.nonzero((idx != oidx).any(axis=1))[0]
            bad_sim = np.nonzero((sim.view(np.int64) != osim.view(np.int64)).any(axis=1))[0]
            line = {"items": E, "precision": precision, "band_rows": int(bd), "fallback_rows": int(fb),
                    "rows_count_differs": int(bad_cnt.size), "rows_set_differs": len(bad_set),
                    "rows_order_differs": int(bad_idx.size), "rows_value_bits_differ": int(bad_sim.size)}
            print(json.dumps(line), flush=True)
            for r in (bad_set[:2] if bad_set else bad_idx[:2].tolist()):
                pos = np.nonzero(idx[r] != oidx[r])[0][:6]
                print("  row", r, "cnt", int(cnt[r]), int(ocnt[r]), "first differing ranks", pos.tolist(),
                      "ours", idx[r, pos].tolist(), [float(x) for x in sim[r, pos]],
                      "oracle", oidx[r, pos].tolist(), [float(x) for x in osim[r, pos]], flush=True)
        bank.close()
    ctx.close()


if __name__ == "__main__":
    main()
