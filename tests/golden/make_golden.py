"""Generates tests/golden/sketch_small.json from the CPU oracle (the reference's Java
cannot run in this image: no JVM).  The independent pure-Python BigInteger/Random
restatement in tests/test_oracle.py is what pins the oracle; this fixture freezes it."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as orc  # noqa: E402

rng = np.random.Generator(np.random.PCG64(20240000))
E, d, w, n, k, seed = 12, 3, 64, 400, 4, 42
ent = rng.integers(0, E - 1, n)           # entity E-1 stays empty (NaN row)
key = rng.integers(-50, 50, n)
inc = rng.integers(1, 11, n) * 0.5
a, b = orc.hash_params(seed, d)
bank = np.zeros((E, d, w))
orc.bank_update(bank, d, w, a, b, ent, key, inc.astype(np.float32))
idx, sim, cnt = orc.bank_cosine_topk(bank, k)
nz = np.flatnonzero(bank)
out = dict(seed=seed, E=E, d=d, w=w, k=k, a=[int(x) for x in a], b=[int(x) for x in b],
           entity=ent.tolist(), key=key.tolist(), inc=inc.tolist(),
           nonzero_cells=nz.tolist(), nonzero_values=bank.ravel()[nz].tolist(),
           topk_idx=idx.tolist(), topk_sim=sim.tolist(), topk_cnt=cnt.tolist())
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sketch_small.json"), "w"))
print("wrote sketch_small.json", len(nz), "non-zero cells")
