"""dev (2+ GPUs, torchrun): certified fallback rows of the fused and the pipelined forms on the same sharded bank."""
import os, sys, time, ctypes as C
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mahout_b200 as mb
from mahout_b200 import _native as N, synth, similarity as sim
from mahout_b200.sketch import last_fallback_rows
import bench_big as bb

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
ctx = mb.Context(local)
ctx.set_option(N.OPT_MAX_FALLBACK_ROWS, 64)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
items, events, k = int(float(sys.argv[1])), int(float(sys.argv[2])), 100
env = bb.Env(ctx, stream, world, rank, local, dev, {"hbm": 6544.3, "bf16": 1381.3, "bf16_burst": 1663.9, "source": "x"})
plan = sim.ShardPlan(items, world, rank)
cdf = torch.from_numpy(synth.zipf_cdf(items, 1.1)).to(dev)
perm = torch.from_numpy(synth.rank_permutation(items, 4) - 1).to(dev)
bank = mb.SketchBank(plan.rows_per_shard, 4096, 1, 42, 1, ctx)
bb.routed_build(env, plan, bank, 20240006, events, 5_000_000, cdf, perm)
be = sim.GpuShardBackend(ctx)
be.bank = bank
peers = sim.PeerRows(ctx, plan, 1, 4096)
blocks = peers.map_counters(bank)
a_cnt = bank.counters_tensor()
N.check(N.lib().mb200_bank_normalize(bank.handle, N.DTYPE_F16, C.c_void_p(peers.rows.data_ptr()), C.c_void_p(peers.valid.data_ptr())), ctx.handle)
res = {}
for form in ("fused", "pipelined8192", "pipelined32768", "fused"):
  try:
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      if form == "fused":
          r = sim.fused_gather_cosine(be, plan, peers, k, None, "f16", "certified", a_counters=a_cnt, counter_blocks=blocks)
      else:
          peers.refresh_narrow()
          peers.barrier()
          r = sim.pipelined_cosine(be, plan, peers.rows, peers.valid, k, None, "f16", "certified", None, int(form[9:]), a_cnt,
                                   False, counter_blocks=blocks, counter_blocks32=peers.counter_blocks32)
          peers.barrier()
      torch.cuda.synchronize()
      fb = last_fallback_rows(ctx)
      print(f"rank {rank} {form}: {1e3 * (time.perf_counter() - t0):.1f} ms, fallback rows {fb}", flush=True)
      if "fused" in res and form != "fused":
          same = all(torch.equal(x, y) for x, y in zip(r, res["fused"]))
          cnt_same = bool((r[2] == res["fused"][2]).all())
          print(f"rank {rank} {form} equals fused: {same} (counts {cnt_same})", flush=True)
      res[form] = tuple(t.clone() for t in r)

  except Exception as ex:
    print(f'rank {rank} {form}: FAILED {ex}', flush=True)
    torch.cuda.synchronize()
    peers.barrier()
peers.close()
dist.barrier()
dist.destroy_process_group()
