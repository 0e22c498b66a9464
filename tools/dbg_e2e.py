"""dev tool: per-step timing of the host-buffer (e2e) sketch update path"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mahout_b200 as mb
from mahout_b200 import _native as N, synth

ctx = mb.Context(0)
n = 1 << 27
big = int(float(os.environ.get("BIG", "0")))
cdf = torch.from_numpy(synth.zipf_cdf(10_000_000, 1.1)).cuda()
if big:
    _, item_big, pref_big = synth.events_device(ctx, 20240002, 0, big, 1_000_000, cdf, None, want_user=False)
_, item, pref = synth.events_device(ctx, 20240002, 0, n, 1_000_000, cdf, None, want_user=False)
hk = torch.empty(n, dtype=torch.int64).pin_memory()
hp = torch.empty(n, dtype=torch.float32).pin_memory()
hk.copy_(item)
hp.copy_(pref)
hout = torch.empty(4 << 20, dtype=torch.float64).pin_memory()
bank = mb.SketchBank(1, 1 << 20, 4, 42, 1, ctx)
hkn, hpn = hk.numpy(), hp.numpy()
for i in range(12):
    t0 = time.perf_counter()
    N.check(N.lib().mb200_bank_update(bank.handle, None, C.c_void_p(hkn.ctypes.data), C.c_void_p(hpn.ctypes.data),
                                      n, N.MEM_HOST), ctx.handle)
    t1 = time.perf_counter()
    N.check(N.lib().mb200_bank_read(bank.handle, 0, 1, C.c_void_p(hout.data_ptr()), N.MEM_HOST), ctx.handle)
    t2 = time.perf_counter()
    print("step %d update %.1f ms read %.1f ms  -> %.2f G ev/s" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3,
                                                                    n / (t2 - t0) / 1e9), flush=True)
