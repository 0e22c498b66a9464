"""Job-level parity on the GPU: the ItemSimilarityJob mirror against the reference's own
end-to-end golden (ItemSimilarityJobTest.testCompleteJob) and against the oracle on the
MovieLens-100K-shaped configuration (BASELINE.json configs[0])."""
import os
import socket

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu


def _run_job(tmp_path, lines, *extra, measure="SIMILARITY_SKETCH_COSINE"):
    from mahout_b200.itemsimilarity import ItemSimilarityJob
    inp = tmp_path / "prefs.csv"
    inp.write_text("\n".join(lines) + "\n")
    out = tmp_path / "out.tsv"
    rc = ItemSimilarityJob().run(["--input", str(inp), "--output", str(out), "--similarityClassname",
                                  measure, *extra])
    assert rc == 0
    rows = []
    for ln in out.read_text().splitlines():
        a, b, s = ln.split("\t")
        rows.append((int(a), int(b), float(s)))
    return rows


def test_complete_job_reference_golden(tmp_path):
    """ItemSimilarityJobTest.testCompleteJob (:113-173): exactly two lines, 1-3 ~0.45, 2-3 ~0.89.
    With a sketch far wider than the 3 users no two keys collide, so sketch cosine == cosine."""
    rows = _run_job(tmp_path, ["2,1,1", "1,2,1", "3,4,1", "1,3,2", "2,3,1"], "--sketchWidth", "65536")
    assert len(rows) == 2
    assert rows[0][:2] == (1, 3) and abs(rows[0][2] - 0.45) < 0.01
    assert rows[1][:2] == (2, 3) and abs(rows[1][2] - 0.89) < 0.01
    assert abs(rows[0][2] - 1 / np.sqrt(5)) < 1e-12 and abs(rows[1][2] - 2 / np.sqrt(5)) < 1e-12


def test_complete_job_exact_measure_reference_golden(tmp_path):
    """-s SIMILARITY_COSINE is the exact measure: the same golden without any sketch width to choose"""
    rows = _run_job(tmp_path, ["2,1,1", "1,2,1", "3,4,1", "1,3,2", "2,3,1"], measure="SIMILARITY_COSINE")
    assert len(rows) == 2
    assert rows[0][:2] == (1, 3) and abs(rows[0][2] - 1 / np.sqrt(5)) < 1e-12
    assert rows[1][:2] == (2, 3) and abs(rows[1][2] - 2 / np.sqrt(5)) < 1e-12


def test_exact_measure_movielens_100k_shaped_equals_rowsimilarityjob_oracle():
    """configs[0]: ItemSimilarityJob -s SIMILARITY_COSINE on 943 x 1682 / 100K prefs against the oracle's
    restatement of RowSimilarityJob (normalise, dot, top-k).  The two formulas round differently
    (AB / (|A| |B|) vs sum of normalised products): values within 1e-12, sets equal up to such ties."""
    from mahout_b200 import similarity as sim
    from oracle import prep as oprep
    rng = np.random.Generator(np.random.PCG64(20240001))
    U, I, n, k = 943, 1682, 100000, 100
    user = rng.integers(1, U + 1, 2 * n)
    item = np.minimum(rng.zipf(1.3, 2 * n), I)
    pref = (rng.integers(1, 11, 2 * n) * 0.5).astype(np.float32)
    prep = oprep.PreferenceMatrix(user, item, pref)
    idx, s, cnt = sim.exact_item_similarity(prep.row, prep.user, prep.pref, prep.num_items, k=k)
    order = np.lexsort((prep.user, prep.row))
    rowptr = np.zeros(prep.num_items + 1, np.int64)
    np.add.at(rowptr, prep.row + 1, 1)
    rowptr = np.cumsum(rowptr)
    ucol = np.searchsorted(np.unique(prep.user), prep.user[order]).astype(np.int32)
    eidx, esim, ecnt = orc.rowsim_cosine_topk(prep.num_items, U, rowptr, ucol, prep.pref[order], k)
    assert (cnt == ecnt).all()
    for r in range(prep.num_items):
        c = cnt[r]
        assert np.allclose(s[r, :c], esim[r, :c], rtol=1e-12, atol=0)
        if (idx[r, :c] != eidx[r, :c]).any():
            # only entries tied (to rounding) may swap or straddle the cut
            kth = esim[r, c - 1]
            for col in set(idx[r, :c].tolist()) ^ set(eidx[r, :c].tolist()):
                # an item on one side of the cut only: its value must tie with the k-th to rounding
                v = s[r, :c][idx[r, :c] == col] if col in idx[r, :c] else esim[r, :c][eidx[r, :c] == col]
                assert abs(v[0] - kth) <= 1e-12 * kth


def test_max_similarities_per_item_and_threshold(tmp_path):
    lines = [f"{u},{i},{1 + (u * i) % 5}" for u in range(1, 12) for i in range(1, 9) if (u + i) % 3]
    all_rows = _run_job(tmp_path, lines, "--sketchWidth", "65536")
    capped = _run_job(tmp_path, lines, "--sketchWidth", "65536", "-m", "1")
    assert set(capped) <= set(all_rows) and len(capped) < len(all_rows)
    per_item = {}
    for a, b, s in all_rows:
        per_item.setdefault(a, []).append(s)
        per_item.setdefault(b, []).append(s)
    # with m = 1 every item still appears with its best partner
    best = {i: max(v) for i, v in per_item.items()}
    seen = {}
    for a, b, s in capped:
        seen[a] = max(seen.get(a, 0), s)
        seen[b] = max(seen.get(b, 0), s)
    assert all(abs(seen[i] - best[i]) < 1e-15 for i in best)
    thr = _run_job(tmp_path, lines, "--sketchWidth", "65536", "--threshold", "0.8")
    assert thr == [r for r in all_rows if r[2] >= 0.8]


def test_movielens_100k_shaped_config(tmp_path):
    """configs[0]: 943 users x 1682 items, 100K unique prefs.  GPU sketch path == oracle sketch path
    (exactly); the exact MapReduce-path oracle gives the recall of the sketch measure."""
    import mahout_b200 as mb
    from mahout_b200 import similarity as sim
    rng = np.random.Generator(np.random.PCG64(20240001))
    U, I, n, k, d, w = 943, 1682, 100000, 100, 4, 4096
    pairs = set()
    cdf = np.cumsum(1.0 / np.arange(1, I + 1))
    cdf /= cdf[-1]
    perm = rng.permutation(I)
    while len(pairs) < n:
        u = rng.integers(1, U + 1, n)
        it = perm[np.minimum(np.searchsorted(cdf, rng.random(n)), I - 1)] + 1
        pairs.update(zip(u.tolist(), it.tolist()))
    pairs = sorted(pairs)[:n]
    user = np.array([p[0] for p in pairs], np.int64)
    item = np.array([p[1] for p in pairs], np.int64)
    pref = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
    from oracle import prep as oprep
    prep = oprep.PreferenceMatrix(user, item, pref)
    idx, s, cnt = sim.item_similarity(prep.row, prep.user, prep.pref, prep.num_items, k=k, width=w, depth=d)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((prep.num_items, d, w))
    orc.bank_update(ref, d, w, a, b, prep.row, prep.user, prep.pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all() and (idx == oidx).all() and s.tobytes() == osim.tobytes()
    # exact path (RowSimilarityJob semantics) for the recall of the sketch measure
    order = np.lexsort((prep.user, prep.row))
    rowptr = np.zeros(prep.num_items + 1, np.int64)
    np.add.at(rowptr, prep.row + 1, 1)
    rowptr = np.cumsum(rowptr)
    ucol = np.searchsorted(np.unique(prep.user), prep.user[order]).astype(np.int32)
    eidx, esim, ecnt = orc.rowsim_cosine_topk(prep.num_items, U, rowptr, ucol, prep.pref[order], k)
    hit = tot = 0
    for r in range(prep.num_items):
        e = set(eidx[r, :ecnt[r]].tolist())
        hit += len(e & set(idx[r, :cnt[r]].tolist()))
        tot += len(e)
    assert hit / tot > 0.5, hit / tot


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from mahout_b200 import similarity as sim
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    rng = np.random.Generator(np.random.PCG64(12))
    N, n = 777, 60000
    row = ((np.minimum(rng.zipf(1.2, n), N) - 1) * 13 % N).astype(np.int64)
    user = rng.integers(1, 3000, n).astype(np.int64)
    pref = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
    res = {}
    # certified across GPUs: the undecided candidates are read from the peers' banks over NVLink (no gather)
    res["certified"] = sim.sharded_item_similarity(row[rank::world], user[rank::world], pref[rank::world], N,
                                                   k=20, width=1024, depth=4, precision="certified")
    # ... and from local copies of the peers' int32 banks, pulled behind the rows (what large shards do)
    keep, sim.LOCAL_COUNTERS_MIN_ROWS = sim.LOCAL_COUNTERS_MIN_ROWS, 0
    res["certified_local_copies"] = sim.sharded_item_similarity(row[rank::world], user[rank::world], pref[rank::world], N,
                                                                k=20, width=1024, depth=4, precision="certified")
    sim.LOCAL_COUNTERS_MIN_ROWS = keep
    for precision in ("rescored", "tensor"):
        res[precision] = sim.sharded_item_similarity(row[rank::world], user[rank::world], pref[rank::world], N,
                                                     k=20, width=1024, depth=4, precision=precision)
        # chunked all-gather overlapped with K3 through the incremental job: must give the same answer
        piped = sim.sharded_item_similarity(row[rank::world], user[rank::world], pref[rank::world], N,
                                            k=20, width=1024, depth=4, precision=precision, chunk_rows=256)
        for x, y in zip(piped, res[precision]):
            assert x.tobytes() == y.tobytes(), f"pipelined {precision} result differs"
    # the all-gather fused into K3 (DMA pulls over NVLink + arrival flags): same answer, three epochs
    import ctypes as C
    from mahout_b200 import _native as NV
    from mahout_b200 import sketch as sk
    ctx = sk.default_context()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    plan = sim.ShardPlan(N, world, rank)
    be = sim.GpuShardBackend(ctx)
    be.build(plan, *plan.my_events(row, user, pref), 1024, 4, 42, 1)
    peers = sim.PeerRows(ctx, plan, 4, 1024)
    fused_ok = True
    for precision in ("tensor", "rescored"):
        want = sim.sharded_item_similarity(row[rank::world], user[rank::world], pref[rank::world], N, k=20, width=1024,
                                           depth=4, precision=precision, gather_result=False, fused=False)
        for rep in range(3):
            NV.check(NV.lib().mb200_bank_normalize(be.bank.handle, NV.DTYPE_F16, C.c_void_p(peers.rows.data_ptr()),
                                                   C.c_void_p(peers.valid.data_ptr())), ctx.handle)
            kw = {}
            if precision == "rescored":
                a_cnt = be.counters()
                kw = dict(a_counters=a_cnt, b_counters=sim._all_gather(a_cnt, world, None))
            got = sim.fused_gather_cosine(be, plan, peers, 20, precision=precision, **kw)
            fused_ok &= all(torch.equal(x.cpu(), torch.as_tensor(y).cpu()) for x, y in zip(got, want))
    peers.close()
    be.close()
    res["fused_ok"] = bool(fused_ok)
    ok = torch.tensor([int(fused_ok)], device=f"cuda:{rank}")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    res["fused_ok"] = bool(ok.item())
    if rank == 0:
        q.put((row, user, pref, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_gpus_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    row, user, pref, res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = orc.hash_params(42, 4)
    ref = np.zeros((777, 4, 1024))
    orc.bank_update(ref, 4, 1024, a, b, row, user, pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, 20)
    idx, s, cnt = res["rescored"]
    assert (cnt == ocnt).all() and (idx == oidx).all() and s.tobytes() == osim.tobytes()
    assert res["fused_ok"], "fused pull-gather result differs from the all-gather + K3 result"
    for form in ("certified", "certified_local_copies"):
        idx, s, cnt = res[form]
        assert (cnt == ocnt).all()
        for r in range(777):
            assert set(idx[r, :cnt[r]].tolist()) == set(oidx[r, :cnt[r]].tolist()), (form, r)
    idx, s, cnt = res["tensor"]
    assert (cnt == ocnt).all()
    dense = orc.bank_cosine_dense(ref)
    for r in range(777):
        for t in range(cnt[r]):
            assert abs(s[r, t] - dense[r, idx[r, t]]) <= 1e-3 * abs(dense[r, idx[r, t]])


def _job_events(seed, n, U, I):
    rng = np.random.Generator(np.random.PCG64(seed))
    item = (np.minimum(rng.zipf(1.2, n), I) - 1).astype(np.int64)
    item = (item * 7919) % I
    user = rng.integers(1, U, n).astype(np.int64)
    pref = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
    return item, user, pref


@pytest.mark.parametrize("precision", ["rescored", "certified", "tensor"])
def test_multi_gpu_job_behind_the_c_abi(precision):
    """mb200_create_multi + mb200_job_item_similarity (csrc/job.cu): one process drives every visible GPU -- routing,
    grouped K1, K2, fused pull-gather K3, top-k -- and must agree with the oracle: re-scored bit for bit, certified by
    sets, tensor within tolerance; with 2+ GPUs also with the single-GPU result of the same call."""
    import torch
    import oracle as orc
    from mahout_b200.multi import MultiGpu
    I, d, w, k = 1500, 4, 1024, 20
    item, user, pref = _job_events(5, 120000, 3000, I)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((I, d, w))
    orc.bank_update(ref, d, w, a, b, item, user, pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    results = []
    for g in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        with MultiGpu(g) as m:
            idx, sim, cnt, st = m.item_similarity(item, user, pref, I, k=k, width=w, depth=d, precision=precision)
        assert st["n_gpus"] == g and st["events"] == item.shape[0] and st["rows"] == I
        assert st["similarities_kept"] == int(cnt.sum())
        assert (cnt == ocnt).all(), g
        if precision == "rescored":
            assert (idx == oidx).all() and sim.tobytes() == osim.tobytes(), g
        elif precision == "certified":
            assert all(set(idx[r, :cnt[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()) for r in range(I)), g
        else:
            dense = orc.bank_cosine_dense(ref)
            for r in range(0, I, 37):
                for t in range(cnt[r]):
                    assert abs(sim[r, t] - dense[r, idx[r, t]]) <= 1e-3 * abs(dense[r, idx[r, t]])
        results.append((idx, sim, cnt))
    for other in results[1:]:
        if precision != "tensor":
            assert (other[2] == results[0][2]).all()
        if precision == "rescored":
            assert (other[0] == results[0][0]).all() and other[1].tobytes() == results[0][1].tobytes()


def test_multi_gpu_job_errors_surface():
    from mahout_b200.multi import MultiGpu
    import mahout_b200 as mb
    with pytest.raises(ValueError):
        MultiGpu(1000)
    with MultiGpu(1) as m:
        item, user, pref = _job_events(6, 5000, 100, 50)
        with pytest.raises(ValueError):
            m.item_similarity(item, user, pref, 50, k=0)
        bad = item.copy()
        bad[7] = 50                                               # a row outside [0, num_items)
        with pytest.raises(ValueError):
            m.item_similarity(bad, user, pref, 50, k=5, width=256, depth=2)
        p = pref.copy()
        p[3] = 0.3                                                # not a multiple of the quantum
        with pytest.raises(mb.InexactError):
            m.item_similarity(item, user, p, 50, k=5, width=256, depth=2)
