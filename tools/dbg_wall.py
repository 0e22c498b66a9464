import sys, os, time, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import mahout_b200 as mb
from mahout_b200 import _native as N, synth
from mahout_b200 import sketch as sk
ctx = mb.Context(0)
items=26744
cdf = torch.from_numpy(synth.zipf_cdf(items, 1.1)).cuda()
perm = torch.from_numpy(synth.rank_permutation(items, 3) - 1).cuda()
user, item, pref = synth.events_device(ctx, 20240003, 0, 20000000, 138493, cdf, perm)
bank = mb.SketchBank(items, 4096, 4, 42, 1, ctx)
bank.update(item, user, pref); bank.check()
rows, valid = bank.normalize("f16")
cnt_t = bank.counters_tensor()
orig_check = N.check
for rep in range(4):
    t0=time.perf_counter()
    torch.cuda.synchronize()
    t1=time.perf_counter()
    r = sk.cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), 4, 4096, 50, b_id=(1, items), precision="rescored", a_counters=cnt_t, b_counters=cnt_t)
    t2=time.perf_counter()
    print("sync %.3f ms  call %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3), flush=True)
# now time the inner parts by hand
import types
dev="cuda:0"
for rep in range(3):
    t0=time.perf_counter()
    idx = torch.empty((items, 50), dtype=torch.int64, device=dev)
    sim = torch.empty((items, 50), dtype=torch.float64, device=dev)
    cnt = torch.empty((items,), dtype=torch.int32, device=dev)
    t1=time.perf_counter()
    args = N.CosineArgs()
    args.a_rows, args.a_valid, args.a_count = rows.data_ptr(), valid.data_ptr(), items
    args.a_id_mul, args.a_id_off = 1, 0
    args.b_rows, args.b_valid, args.b_count, args.b_blocks = rows.data_ptr(), valid.data_ptr(), items, 1
    args.b_id_mul, args.b_id_add = 1, items
    args.depth, args.width, args.dtype, args.precision = 4, 4096, N.DTYPE_F16, N.PRECISION_RESCORED
    args.k, args.threshold, args.exclude_self, args.block_n = 50, 0.0, 1, 0
    args.a_counters = cnt_t.data_ptr(); args.b_counters = cnt_t.data_ptr()
    args.out_idx, args.out_sim, args.out_cnt = idx.data_ptr(), sim.data_ptr(), cnt.data_ptr()
    torch.cuda.synchronize(0)
    t2=time.perf_counter()
    rc = N.lib().mb200_cosine_topk(ctx.handle, C.byref(args))
    t3=time.perf_counter()
    ctx.sync()
    t4=time.perf_counter()
    print("alloc %.3f sync %.3f call %.3f ctxsync %.3f" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3), flush=True)
