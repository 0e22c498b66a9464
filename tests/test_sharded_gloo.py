"""world_size-2 gloo run of the sharded driver's host logic (event routing, shard layout, index
mapping, all-gathers, result assembly).  The per-rank compute is injected from the test as an
oracle-backed stand-in -- the product has no CPU backend (tests/test_abi.py::test_no_cpu_fallback)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from mahout_b200 import similarity as sim


class OracleShardBackend:
    """Same interface as GpuShardBackend; tensors are CPU, arithmetic is the oracle's."""

    def build(self, plan, local_row, key, inc, width, depth, seed, frac_bits):
        self.d, self.w = depth, width
        a, b = orc.hash_params(seed, depth)
        self.bank = np.zeros((plan.rows_per_shard, depth, width))
        orc.bank_update(self.bank, depth, width, a, b, local_row, key, inc)

    def normalized(self, dtype):
        # the driver only moves these around; the stand-in ships the raw counters as "rows"
        rows = torch.from_numpy(self.bank.transpose(1, 0, 2).copy())          # [d, E_loc, w]
        valid = torch.from_numpy((np.abs(self.bank).sum(axis=2) > 0).T.astype(np.int32).copy())
        return rows, valid

    def counters(self):
        return torch.from_numpy(self.bank.copy())

    def cosine(self, plan, a_rows, a_valid, b_rows, b_valid, k, threshold, dtype, precision,
               a_counters=None, b_counters=None):
        G, d, E_loc, w = b_rows.shape
        # gathered blocks -> global bank in global row order (row = l * G + g)
        full = np.zeros((E_loc * G, d, w))
        for g in range(G):
            full[g::G] = b_rows[g].numpy().transpose(1, 0, 2)
        idx = np.zeros((E_loc, k), np.int64)
        s = np.zeros((E_loc, k))
        cnt = np.zeros(E_loc, np.int32)
        for l in range(E_loc):
            r = l * G + plan.rank
            i1, s1, c1 = orc.bank_cosine_topk(full, k, threshold if threshold else orc.NO_THRESHOLD,
                                              True, r0=r, r1=r + 1, nthreads=1)
            idx[l], s[l], cnt[l] = i1[0], s1[0], c1[0]
        return torch.from_numpy(idx), torch.from_numpy(s), torch.from_numpy(cnt)

    def begin(self, plan, a_rows, a_valid, k, threshold, dtype, precision):
        return _OracleJob(plan, k, threshold)

    def close(self):
        pass


class _OracleJob:
    """stand-in for sk.CosineJob: collects the pushed pieces, rebuilds the bank from the index mapping"""

    def __init__(self, plan, k, threshold):
        self.plan, self.k, self.threshold, self.pieces = plan, k, threshold, []

    def push(self, b_rows, b_valid, id_mul=1, id_add=0, id_base=0):
        self.pieces.append((b_rows.numpy().copy(), id_mul, id_add, id_base))

    def finish(self, a_counters=None, b_counters=None, b_id=(1, 0), out=None):
        plan, k = self.plan, self.k
        d, w = self.pieces[0][0].shape[1], self.pieces[0][0].shape[3]
        full = np.zeros((plan.rows_per_shard * plan.G, d, w))
        seen = np.zeros(full.shape[0], bool)
        for rows, mul, add, base in self.pieces:
            for g in range(rows.shape[0]):
                ids = np.arange(rows.shape[2]) * mul + g * add + base
                assert not seen[ids].any()
                seen[ids] = True
                full[ids] = rows[g].transpose(1, 0, 2)
        assert seen.all(), "some global rows were never pushed"
        if b_counters is not None:      # the re-score operand must be the same bank, laid out [G, E_loc, d, w]
            chk = np.zeros_like(full)
            for g in range(plan.G):
                chk[g::plan.G] = b_counters[g].numpy()
            assert (chk == full).all()
        E_loc = plan.rows_per_shard
        idx, s, cnt = np.zeros((E_loc, k), np.int64), np.zeros((E_loc, k)), np.zeros(E_loc, np.int32)
        for l in range(E_loc):
            r = l * plan.G + plan.rank
            i1, s1, c1 = orc.bank_cosine_topk(full, k, self.threshold if self.threshold else orc.NO_THRESHOLD,
                                              True, r0=r, r1=r + 1, nthreads=1)
            idx[l], s[l], cnt[l] = i1[0], s1[0], c1[0]
        return torch.from_numpy(idx), torch.from_numpy(s), torch.from_numpy(cnt)

    def abort(self):
        pass


def _events(seed, n, N, users):
    rng = np.random.Generator(np.random.PCG64(seed))
    row = (np.minimum(rng.zipf(1.3, n), N) - 1).astype(np.int64)
    row = (row * 31) % N
    return row, rng.integers(1, users, n).astype(np.int64), (rng.integers(1, 11, n) * 0.5).astype(np.float32)


def _worker(rank, world, port, N, k, d, w, out_q, chunk_rows=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    row, user, pref = _events(5, 4000, N, 200)
    mine = slice(rank, None, world)               # each rank holds an arbitrary slice of the stream
    idx, s, cnt = sim.sharded_item_similarity(row[mine], user[mine], pref[mine], N, k=k, width=w, depth=d,
                                              precision="rescored", backend=OracleShardBackend(),
                                              chunk_rows=chunk_rows)
    if rank == 0:
        out_q.put((idx, s, cnt))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(N, k, d, w, chunk_rows):
    world = 2
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, k, d, w, q, chunk_rows)) for r in range(world)]
    for p in procs:
        p.start()
    idx, s, cnt = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    row, user, pref = _events(5, 4000, N, 200)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((N, d, w))
    orc.bank_update(ref, d, w, a, b, row, user, pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all() and (idx == oidx).all()
    assert s.tobytes() == osim.tobytes()


@pytest.mark.timeout(300)
def test_sharded_driver_world2_gloo():
    _run_world2(45, 6, 3, 64, 0)


@pytest.mark.timeout(300)
def test_pipelined_driver_world2_gloo():
    """chunked all-gather + incremental job: two row chunks per shard ([0,256) and [256,301)), odd N so the
    last shard is one row short"""
    _run_world2(601, 6, 2, 32, 256)


def _collective_worker(rank, world, port, out_q):
    """a step that fails on rank 1 only (second call) must fail on every rank at the same point, in both modes"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    seen = {}
    for check_each in (True, False):
        calls = [0]

        def step():
            calls[0] += 1
            dist.barrier()                                  # every call meets its peers, like the fused step
            if rank == 1 and calls[0] == 2:
                raise ValueError("lost block")
            return calls[0]

        try:
            bench._collective_steps(step, 3, "cpu", check_each)
            seen[check_each] = ("no error", calls[0])
        except Exception as ex:
            seen[check_each] = (type(ex).__name__, calls[0])
    ok = sim.all_ranks_ok(True, "cpu")
    out_q.put((rank, seen, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_step_failures_are_collective_world2_gloo():
    world = 2
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_collective_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (seen, ok)) for r, seen, ok in (q.get(timeout=90) for _ in range(world)))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # checked after every call: both ranks stop after the failing (second) call; checked once at the end: both make all
    # three calls, then both fail -- the failing rank with its own error, the other with "failed on another rank"
    assert got[0][0] == {True: ("RuntimeError", 2), False: ("RuntimeError", 3)}
    assert got[1][0] == {True: ("ValueError", 2), False: ("ValueError", 3)}
    assert got[0][1] and got[1][1]
