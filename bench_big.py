"""bench_big.py -- the large configurations of BASELINE.json, run from bench.py at N = 8 (or, scaled by flags, at
any N for development):

  configs[3]  10^6-item all-pairs sketch cosine (width 4096), fused top-100 epilogue, items hash-sharded across the
              GPUs: routed sketch build (route.cu + grouped K1) -> K2 -> fused pull-gather K3 -> certified top-k
  configs[4]  heavy-skew stress: (a) 10^10 Zipf(1.5) events into replica sketches + all-reduce,
              (b) the 10^7-item cosine, scaled (default 4*10^6 items), in the STREAMED form: chunks of every shard's
              rows are gathered into two staging buffers while K3 consumes them (the gathered operand never exists)

Every object carries its parity leg: top-k SETS of seeded sample rows against the CPU oracle (oracle/fast.py, pinned
bit for bit to the loop-for-loop C oracle by tests/test_oracle.py) with every rank scoring the sample against its own
columns on its share of the host cores.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np


def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class Env:
    """what every stage needs: context, stream, ranks"""

    def __init__(self, ctx, stream, world, rank, local, dev, peaks):
        self.ctx, self.stream, self.world, self.rank, self.local, self.dev, self.peaks = ctx, stream, world, rank, local, dev, peaks
        self.t0 = time.perf_counter()

    def log(self, msg: str):
        """progress of the long stages, rank 0, stderr (stdout carries the one JSON line)"""
        if self.rank == 0:
            import sys
            print(f"[bench_big +{time.perf_counter() - self.t0:6.1f} s] {msg}", file=sys.stderr, flush=True)

    def elapsed(self) -> float:
        """seconds since the stages started, agreed by all ranks (the slowest clock)"""
        return self.max_over_ranks(time.perf_counter() - self.t0)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def gather_objects(self, obj):
        import torch.distributed as dist
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out


# ------------------------------------------------------------------------------------------------------------------
# routed sketch build
# ------------------------------------------------------------------------------------------------------------------
def routed_build(env: Env, plan, bank, seed, n_total, users, cdf, perm):
    """Every rank generates its 1/G slice of the stream on its GPU, the events travel to the owners of their items
    (route.cu: partition + peer scatter), the owner groups them by item and updates its shard bank (group.cu).
    Returns timings (max over ranks) and counts."""
    import torch
    from mahout_b200 import _native as N
    from mahout_b200 import similarity as sim
    from mahout_b200 import synth
    ctx, world, rank = env.ctx, env.world, env.rank
    n_mine = n_total // world
    user, item, pref = synth.events_device(ctx, seed, rank * n_mine, n_mine, users, cdf, perm)
    router = None
    t0 = time.perf_counter()
    if world > 1:
        # set-up (not timed as part of the step): receive columns sized by what the largest shard receives, mapped
        # by every peer.  The timed route() counts again -- the count kernel and the exchange belong to the step.
        probe = sim.EventRouter.__new__(sim.EventRouter)
        probe.ctx, probe.plan, probe.group = ctx, plan, None
        _, matrix = sim.EventRouter.counts(probe, item)
        router = sim.EventRouter(ctx, plan, int(matrix.sum(axis=0).max()))
    ctx.set_profiling(True)
    ctx.reset_profile()
    env.barrier()
    setup_s = time.perf_counter() - t0
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t0 = time.perf_counter()
    e[0].record(env.stream)
    if world > 1:
        lrow, luser, lpref = router.route(item, user, pref)
    else:
        lrow, luser, lpref = item, user, pref
    e[1].record(env.stream)
    bank.update(lrow, luser, lpref)
    e[2].record(env.stream)
    env.barrier()
    wall = env.max_over_ranks(time.perf_counter() - t0)
    route_ms = env.max_over_ranks(e[0].elapsed_time(e[1]))
    k1_ms = env.max_over_ranks(e[1].elapsed_time(e[2]))
    g_ms, _ = ctx.kernel_time(N.K_GROUP)
    u_ms, _ = ctx.kernel_time(N.K_UPDATE)
    r_ms, _ = ctx.kernel_time(N.K_ROUTE)
    n_recv = int(lrow.numel())
    bank.check()
    recv = env.gather_objects(n_recv)
    del user, item, pref, lrow, luser, lpref
    if router is not None:
        router.close()
    torch.cuda.empty_cache()
    model = 20 + 16 * bank.d
    hot = max(recv)
    return {
        "events": n_total, "events_received_per_gpu": recv,
        "route_ms": route_ms, "k1_ms": k1_ms, "wall_s": wall, "router_setup_s": setup_s,
        "events_per_s": n_total / ((route_ms + k1_ms) * 1e-3),
        "events_per_s_per_gpu": n_total / world / ((route_ms + k1_ms) * 1e-3),
        "kernels_ms_this_rank": {"route": r_ms, "group": g_ms, "update": u_ms},
        # the busiest shard (the owner of the hottest item) bounds the step: its K1 against the 84-byte model
        "k1_hbm_frac_by_model_busiest_gpu": model * hot / (k1_ms * 1e-3) / 1e9 / env.peaks["hbm"],
        "note": "route = count by owner + all-gather of the GxG counts + one partition/peer-scatter kernel + barrier; "
                "k1 = histogram + two partition passes + shared-memory-tile update (group.cu); item-hash sharding puts "
                "the hottest item's events on one GPU, so the shards are uneven by construction",
    }


# ------------------------------------------------------------------------------------------------------------------
# distributed parity: sampled rows against every column
# ------------------------------------------------------------------------------------------------------------------
def sampled_parity(env: Env, plan, bank, got, k, rows_total, seed, block=8192):
    """got = this rank's (idx, sim, cnt) device tensors in local row order.  A seeded sample of `rows_total` global
    rows is scored exactly (oracle/fast.py) by every rank against its own columns; partial top-k lists are merged on
    rank 0 and compared with the product's rows.  Returns the parity dict on rank 0 (None elsewhere)."""
    import torch
    from oracle import fast
    world, rank, G = env.world, env.rank, plan.G
    rng = np.random.Generator(np.random.PCG64(seed))
    sample = np.sort(rng.choice(plan.N, size=min(rows_total, plan.N), replace=False)).astype(np.int64)
    mine = sample[sample % G == rank]
    local = torch.from_numpy(mine // G).to(env.dev)
    cnt_t = bank.counters_tensor()                                   # [E_loc, d, W] int64 quanta
    my_q = cnt_t[local].to(torch.int32).cpu().numpy()
    my_got = tuple(t[local].cpu().numpy() for t in got)
    parts = env.gather_objects((mine, my_q, my_got))
    ids = np.concatenate([p[0] for p in parts])
    order = np.argsort(ids, kind="stable")
    ids = ids[order]
    sample_q = np.concatenate([p[1] for p in parts])[order]
    g_idx = np.concatenate([p[2][0] for p in parts])[order]
    g_sim = np.concatenate([p[2][1] for p in parts])[order]
    g_cnt = np.concatenate([p[2][2] for p in parts])[order]
    n_loc = plan.local_count(rank)
    col_ids = np.arange(n_loc, dtype=np.int64) * G + rank
    # FP64 BLAS threads per rank: the products are skinny ([rows, W] x [W, 8192]) and OpenBLAS scales poorly on them
    # (measured on the GPU boxes: 105 GFLOP/s on 12 threads, 48 on one) -- a few threads per rank, all ranks at once
    threads = max(1, min(2 if world >= 4 else 4, _host_threads() // max(world, 1)))

    def loader(c0, c1):
        return cnt_t[c0:c1].to(torch.int32).cpu().numpy()

    env.log(f"parity: {ids.shape[0]} sample rows x {n_loc} local columns per rank, {threads} BLAS thread(s) per rank")
    t0 = time.perf_counter()
    ps, pi = fast.rows_vs_columns_topk(sample_q, ids, None, col_ids, k, block=block, threads=threads, chunk_loader=loader)
    oracle_s = time.perf_counter() - t0
    env.log(f"parity: oracle done in {oracle_s:.1f} s on this rank")
    allp = env.gather_objects((ps, pi, oracle_s))
    if rank != 0:
        return None
    o_idx, o_sim, o_cnt = fast.merge_partials([(p[0], p[1]) for p in allp], k)
    sets_equal = bool((g_cnt == o_cnt).all() and all(
        set(g_idx[r, :g_cnt[r]].tolist()) == set(o_idx[r, :o_cnt[r]].tolist()) for r in range(ids.shape[0])))
    bad_rows = int(sum(1 for r in range(ids.shape[0]) if g_cnt[r] != o_cnt[r] or
                       set(g_idx[r, :g_cnt[r]].tolist()) != set(o_idx[r, :o_cnt[r]].tolist())))
    max_rel = 0.0
    for r in range(ids.shape[0]):
        o = dict(zip(o_idx[r, :o_cnt[r]].tolist(), o_sim[r, :o_cnt[r]].tolist()))
        for c, v in zip(g_idx[r, :g_cnt[r]].tolist(), g_sim[r, :g_cnt[r]].tolist()):
            if c in o and o[c] != 0.0:
                max_rel = max(max_rel, abs(v - o[c]) / abs(o[c]))
    slowest = max(p[2] for p in allp)
    return {"rows_checked": int(ids.shape[0]), "columns_per_row": int(plan.N), "certified_topk_sets_equal_oracle": sets_equal,
            "rows_with_a_different_set": bad_rows, "max_rel_err_of_returned_sims": max_rel, "tolerance": 1e-3,
            "oracle": "oracle/fast.py (exact integer dot products in FP64 blocks; pinned to the C loop oracle)",
            "oracle_s_slowest_rank": slowest, "oracle_threads_per_rank": threads,
            "oracle_pairs_per_s_all_ranks": ids.shape[0] * float(plan.N) / slowest}


# ------------------------------------------------------------------------------------------------------------------
# N-item all-pairs cosine, certified top-k, items sharded
# ------------------------------------------------------------------------------------------------------------------
def climb_down(ladder, attempt, on_abandon=None):
    """Try the modes of `ladder` (tuples (form, precision)) in order: attempt(mode) -> (result, error).  Returns
    (mode that worked, its result, [{"form", "precision", "error"} of the abandoned ones]); the last mode's error is
    raised.  Everything in attempt must be collective across ranks, errors included (run_steps)."""
    abandoned = []
    for i, m in enumerate(ladder):
        result, err = attempt(m)
        if err is None:
            return m, result, abandoned
        abandoned.append({"form": m[0], "precision": m[1], "error": repr(err)[:300]})
        if on_abandon is not None:
            on_abandon(m, err)
        if i == len(ladder) - 1:
            raise err
    raise ValueError("empty ladder")


def big_cosine(env: Env, name, workload, items, users, events, zipf, depth, width, k, form, check_rows, seed,
               chunk_rows=8192, reps=1, warmup=1, fallback_limit=256):
    import torch
    import mahout_b200 as mb
    from mahout_b200 import _native as N
    from mahout_b200 import similarity as sim
    from mahout_b200 import synth
    from mahout_b200.sketch import last_fallback_rows
    ctx, world, rank, dev = env.ctx, env.world, env.rank, env.dev
    plan = sim.ShardPlan(items, world, rank)
    E_loc = plan.rows_per_shard
    cdf = torch.from_numpy(synth.zipf_cdf(items, zipf)).to(dev)
    perm = torch.from_numpy(synth.rank_permutation(items, 4) - 1).to(dev)
    env.log(f"{name}: {items} items, {int(events)} events, depth {depth}, {form}")
    bank = mb.SketchBank(E_loc, width, depth, 42, 1, ctx)
    build = routed_build(env, plan, bank, seed, int(events), users, cdf, perm)
    env.log(f"{name}: sketch build done (route {build['route_ms']:.1f} ms + K1 {build['k1_ms']:.1f} ms)")
    del cdf, perm
    N.check(N.lib().mb200_release_workspace(ctx.handle), ctx.handle)   # the grouping workspaces (20 B / event)
    torch.cuda.empty_cache()

    be = sim.GpuShardBackend(ctx)
    be.bank = bank
    ld = int(N.lib().mb200_row_ld(width))
    a_cnt = bank.counters_tensor()
    ctx.set_profiling(True)
    peers = sim.PeerRows(ctx, plan, depth, width, staging=(form == "fused")) if world > 1 else None
    out_t = (torch.empty((E_loc, k), dtype=torch.int64, device=dev), torch.empty((E_loc, k), dtype=torch.float64, device=dev),
             torch.empty((E_loc,), dtype=torch.int32, device=dev))
    if world > 1:
        blocks = peers.map_counters(bank)
        rows, valid = peers.rows, peers.valid
    else:
        vw = int(N.lib().mb200_valid_words(E_loc))
        rows = torch.empty((depth, E_loc, ld), dtype=torch.float16, device=dev)
        valid = torch.empty((depth, vw), dtype=torch.int32, device=dev)

    def k2():
        N.check(N.lib().mb200_bank_normalize(bank.handle, N.DTYPE_F16, C.c_void_p(rows.data_ptr()),
                                             C.c_void_p(valid.data_ptr())), ctx.handle)

    k2()
    mixed = be.mixed_sign("certified")
    # a row that neither its candidate list nor the band pass can settle costs one exact dot product per COLUMN: a
    # handful is fine, thousands would take the stage past its budget -- then the stage reports the error and the
    # tensor-precision result instead of hanging
    ctx.set_option(N.OPT_MAX_FALLBACK_ROWS, max(8, int(fallback_limit) // max(depth, 1)))
    # The stage climbs down a ladder when a form fails on any rank: the fused pull-gather (DMA pulls + arrival flags)
    # is the fastest form but depends on the copy engines moving peer memory beside a kernel that holds every SM; if
    # a block does not arrive in time (the kernel reports it after ~4 s) the stage is repeated in the streamed form
    # (NCCL all-gathers of row chunks); if the certified precision itself fails (rows for the exact path beyond the
    # limit above), the tensor precision is reported instead, and the line says so.
    ladder = [(form, "certified")]
    if world > 1 and form == "fused":
        ladder.append(("pipelined", "certified"))
    ladder.append((ladder[-1][0], "tensor"))
    if getattr(env, "fused_failed", False) and world > 1:
        ladder = [m for m in ladder if m[0] != "fused"]
    mode = [ladder[0]]

    def step():
        cur_form, cur_prec = mode[0]
        r0 = sim.FUSED_RETRIES[0]
        out = step_in(cur_form, cur_prec)
        if sim.FUSED_RETRIES[0] != r0:
            # recovered by a second sweep after a time-out of seconds: as far as timing goes, this form failed
            raise RuntimeError("fused pull-gather: a peer block timed out (result recovered by a second sweep)")
        return out

    def step_in(cur_form, cur_prec):
        k2()
        if world == 1:
            from mahout_b200.sketch import cosine_topk_blocks
            kw = dict(a_counters=a_cnt, b_counters=a_cnt, mixed_sign=mixed) if cur_prec == "certified" else {}
            return cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), depth, width, k,
                                      a_id=(1, 0), b_id=(1, E_loc), precision=cur_prec, out=out_t, **kw)
        if cur_form == "fused":
            if cur_prec == "tensor":
                return sim.fused_gather_cosine(be, plan, peers, k, None, "f16", "tensor", out=out_t)
            return sim.fused_gather_cosine(be, plan, peers, k, None, "f16", "certified", a_counters=a_cnt,
                                           counter_blocks=blocks, out=out_t, mixed_sign=mixed)
        if cur_prec == "certified":
            peers.refresh_narrow()
        peers.barrier()
        if cur_prec == "tensor":
            r = sim.pipelined_cosine(be, plan, rows, valid, k, None, "f16", "tensor", None, chunk_rows, None)
        else:
            r = sim.pipelined_cosine(be, plan, rows, valid, k, None, "f16", "certified", None, chunk_rows, a_cnt, mixed,
                                     counter_blocks=blocks, counter_blocks32=peers.counter_blocks32)
        peers.barrier()
        return r

    def run_steps(nsteps):
        """nsteps steps; an error on any rank is an error on all of them (the steps are collective)"""
        err, out = None, None
        for _ in range(nsteps):
            try:
                out = step()
            except Exception as ex:
                err = ex
            if not sim.all_ranks_ok(err is None, dev):
                return None, err or RuntimeError("the step failed on another rank")
        return out, None

    def attempt(m):
        """warm-up + the timed steps in mode m -> ((result, start event, end event), error)"""
        mode[0] = m
        _, err = run_steps(max(warmup, 1))
        if err is not None:
            return None, err
        env.barrier()
        ctx.reset_profile()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(env.stream)
        out, err = run_steps(reps)
        e1.record(env.stream)
        env.barrier()
        return (out, e0, e1), err

    def on_abandon(m, err):
        env.log(f"{name}: {m[0]} / {m[1]} abandoned ({repr(err)[:200]})")
        if m[0] == "fused":
            env.fused_failed = True              # the later stages of this run start in the streamed form
        torch.cuda.synchronize(dev)

    done, (got, e0, e1), abandoned = climb_down(ladder, attempt, on_abandon)
    mode[0] = done
    for a_ in abandoned:
        a_["form"] = a_["form"] if world > 1 else "single GPU"
    form, precision = mode[0][0], [mode[0][1]]
    certified_error = abandoned[-1]["error"] if (abandoned and precision[0] == "tensor") else None
    ms = env.max_over_ranks(e0.elapsed_time(e1) / reps)
    env.log(f"{name}: cosine step {ms:.1f} ms")
    k2_ms, k2_n = ctx.kernel_time(N.K_NORMALIZE)
    k3_ms, k3_n = ctx.kernel_time(N.K_COSINE)
    k5_ms, k5_n = ctx.kernel_time(N.K_RESCORE)
    fallback = env.sum_over_ranks(last_fallback_rows(ctx))
    from mahout_b200.sketch import last_band_rows
    band_rows = env.sum_over_ranks(last_band_rows(ctx))
    ctx.set_option(N.OPT_MAX_FALLBACK_ROWS, -1)
    k3_s = env.max_over_ranks(k3_ms / max(reps, 1)) * 1e-3            # all K3 launches of one step, slowest rank
    ctx.set_profiling(False)
    parity = sampled_parity(env, plan, bank, got, k, check_rows, seed + 17) if check_rows > 0 else None
    # loop-for-loop CPU port on a few rows of this shard, for the extrapolated baseline
    cpu = None
    if rank == 0:
        import oracle as orc
        R, Ccols = 4, min(4096, plan.local_count(0))
        sub = a_cnt[:Ccols].cpu().numpy().astype(np.float64) * 0.5
        t0 = time.perf_counter()
        orc.bank_cosine_dense(sub, 0, R, nthreads=1)
        dt = time.perf_counter() - t0
        rate1 = R * Ccols / dt
        threads = _host_threads()
        cpu = {"value": rate1 * threads, "unit": "pairs/s", "cores": threads, "kind": "port",
               "single_thread_pairs_per_s": rate1,
               "sample": f"{R} rows x {Ccols} columns of shard 0 through the C loop oracle on one thread ({dt:.2f} s), "
                         f"multiplied by {threads} threads (rows are independent): EXTRAPOLATED, not run at this size",
               "extrapolated_s_for_this_config": float(items) ** 2 / (rate1 * threads)}
    flops = 2.0 * depth * float(E_loc * world) ** 2 * ld
    pairs = float(items) ** 2
    res = None
    if rank == 0:
        peak = env.peaks["bf16"] * world
        res = {
            "name": name, "metric": "item_pair_cosine_sims_per_sec", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": ms, "n_gpus": world,
            "precision": "certified (exact top-k sets, tensor-core values)" if precision[0] == "certified" else
                         "tensor (certified precision abandoned: see certified_error)",
            "certified_error": certified_error, "abandoned_forms": abandoned, "band_rows": int(band_rows),
            "form": form if world > 1 else "single GPU", "certified_fallback_rows": int(fallback), "mixed_sign": bool(mixed),
            "config": {"workload": workload, "items": items, "depth": depth, "width": width, "k": k, "events": int(events),
                       "zipf_s": zipf, "rows_per_gpu": E_loc, "chunk_rows": chunk_rows if form == "pipelined" else None,
                       "gathered_operand_bytes": world * depth * E_loc * ld * 2,
                       "staging_bytes": (2 * world * depth * min(chunk_rows, E_loc) * ld * 2) if form == "pipelined"
                       else world * depth * E_loc * ld * 2,
                       "parallelism": f"item-hash sharded x{world}", "reps": reps, "warmup": warmup},
            "kernels_ms_per_step_this_rank": {"K2_normalize": k2_ms / max(reps, 1), "K3_cosine_topk": k3_ms / max(reps, 1),
                                              "K3_launches": int(k3_n // max(reps, 1)), "K5_merge_certify": k5_ms / max(reps, 1)},
            "roofline": {"bound": "tensor", "kernel": "k_cosine<256,*>", "flops_per_step": flops,
                         "achieved": flops / (ms * 1e-3) / 1e12, "achieved_k3_only": flops / k3_s / 1e12,
                         "peak": peak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peak,
                         "frac_k3_only": flops / k3_s / 1e12 / peak,
                         "frac_of_burst_peak": flops / (ms * 1e-3) / 1e12 / (env.peaks["bf16_burst"] * world),
                         "peak_source": env.peaks["source"] + f" x{world} GPUs (sustained)", "traffic": None},
            "sketch_build": build, "parity": parity, "cpu_baseline": cpu,
        }
    if peers is not None:
        peers.close()
    bank.close()
    del out_t, rows, valid, got
    N.check(N.lib().mb200_release_workspace(ctx.handle), ctx.handle)
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------------
# config 5a: heavy-skew single-sketch update, replicas + all-reduce
# ------------------------------------------------------------------------------------------------------------------
def skew_update(env: Env, events_total, items, zipf, depth, width, steps, warmup, seed):
    import torch
    import torch.distributed as dist
    import mahout_b200 as mb
    from mahout_b200 import _native as N
    from mahout_b200 import synth
    from mahout_b200.sketch import _as_tensor
    ctx, world, rank, dev = env.ctx, env.world, env.rank, env.dev
    n = int(events_total) // world
    env.log(f"config5a: {int(events_total)} Zipf({zipf}) events into replica sketches")
    cdf_h = synth.zipf_cdf(items, zipf)
    cdf = torch.from_numpy(cdf_h).to(dev)
    _, item, pref = synth.events_device(ctx, seed, rank * n, n, 1_000_000, cdf, None, want_user=False)
    bank = mb.SketchBank(1, width, depth, 42, 1, ctx)
    cptr, cells = bank.counters_ptr()
    counters = _as_tensor(cptr, cells, env.local)
    glob = torch.empty_like(counters)

    def step():
        bank.update(None, item, pref)
        if world > 1:
            glob.copy_(counters)
            dist.all_reduce(glob, op=dist.ReduceOp.SUM)

    for _ in range(warmup):
        step()
    bank.check()
    bank.clear()
    env.barrier()
    ctx.set_profiling(True)
    ctx.reset_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(env.stream)
    for _ in range(steps):
        step()
    e1.record(env.stream)
    env.barrier()
    ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
    k_ms, k_n = ctx.kernel_time(N.K_UPDATE)
    ctx.set_profiling(False)
    bank.check()
    # properties: every row of the GLOBAL sketch holds the total mass of all ranks' events (x steps)
    total_q = env.sum_over_ranks(float((pref.double() * 2).sum().item())) * steps
    g = (glob if world > 1 else counters).view(depth, width)
    mass_ok = bool(all(float(x) == total_q for x in g.sum(dim=1).tolist()))
    res = None
    if rank == 0:
        import oracle as orc
        m = int(min(n, 1 << 24))
        a, b = orc.hash_params(42, depth)
        ref = np.zeros((1, depth, width))
        threads = _host_threads()
        hi, hp = item[:m].cpu().numpy(), pref[:m].cpu().numpy()
        t0 = time.perf_counter()
        orc.bank_update(ref, depth, width, a, b, None, hi, hp, nthreads=threads)
        cpu_dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref1 = np.zeros((1, depth, width))
        m1 = m >> 3
        orc.bank_update(ref1, depth, width, a, b, None, hi[:m1], hp[:m1], nthreads=1)
        cpu1_dt = time.perf_counter() - t0
        pb = mb.SketchBank(1, width, depth, 42, 1, ctx)
        pb.update(None, item[:m], pref[:m])
        exact = bool(pb.read().tobytes() == ref.tobytes())
        pb.close()
        kern_s = k_ms / max(k_n, 1) * 1e-3
        top = float(cdf_h[0])
        res = {
            "name": "config5a_skew_update", "metric": "sketch_updates_per_sec", "value": world * n / (ms * 1e-3),
            "unit": "events/s", "ms_per_step": ms, "n_gpus": world, "steps": steps, "warmup": warmup, "scaling": "weak",
            "config": {"workload": "configs[4]a: heavy-skew stress, Zipf(1.5) (user,item,pref) events into a depth=4 x "
                                   "width=2^20 sketch: replica sketches + all-reduce(int64 sum)",
                       "events_total_per_step": world * n, "events_per_step_per_gpu": n, "items": items, "zipf_s": zipf,
                       "hottest_key_share": top, "depth": depth, "width": width},
            "kernel_ms_per_launch": kern_s * 1e3, "kernel_events_per_s_per_gpu": n / kern_s,
            "parity": {"prefix_bit_exact": exact, "events_checked": m, "row_mass_conserved_global_sketch": mass_ok},
            "cpu_baseline": {"value": m / cpu_dt, "unit": "events/s", "cores": threads, "kind": "port",
                             "single_thread_events_per_s": m1 / cpu1_dt,
                             "sample": f"first {m} events of rank 0's stream on {threads} threads ({cpu_dt:.2f} s); "
                                       f"{m1} on one thread ({cpu1_dt:.2f} s)"},
        }
    bank.close()
    del item, pref, glob, counters
    torch.cuda.empty_cache()
    return res
