"""GPU parity of the cosine stage (K2 normalise, K3 tcgen05 S.S^T + fused min/top-k, K5 merge and
exact re-score) against the CPU oracle, through the C ABI.

Bars: tensor-core similarities within 1e-3 relative of the oracle's FP64 cosine (BASELINE.json
north_star); re-scored similarities bit-equal; top-k index lists equal under (sim desc, index asc)."""
import json
import os

import numpy as np
import pytest

import oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-3


@pytest.fixture(scope="module")
def mb():
    import mahout_b200
    return mahout_b200


@pytest.fixture(scope="module")
def ctx(mb):
    c = mb.Context(0)
    yield c
    c.close()


def _make_bank(mb, ctx, E, d, w, n, seed, users=943, empty=(), zipf=1.2, frac_bits=1):
    """item-similarity mode: entity = item, key = user, inc = pref."""
    rng = np.random.Generator(np.random.PCG64(seed))
    item = (np.minimum(rng.zipf(zipf, n), E) - 1).astype(np.int64)
    item = (item * 7919) % E                       # spread the popular items over the index range
    if len(empty):
        keep = ~np.isin(item, np.array(empty))
        item = item[keep]
    user = rng.integers(1, users + 1, item.shape[0]).astype(np.int64)
    pref = (rng.integers(1, 11, item.shape[0]) * 0.5).astype(np.float32)
    bank = mb.SketchBank(E, w, d, 42, frac_bits, ctx)
    bank.update(item, user, pref)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((E, d, w))
    orc.bank_update(ref, d, w, a, b, item, user, pref)
    assert bank.read().tobytes() == ref.tobytes()
    return bank, ref


def _dense(mb, ctx, bank, dtype="f16", block_n=0):
    from mahout_b200.sketch import cosine_topk_blocks
    rows, valid = bank.normalize(dtype)
    idx, sim, cnt, dense = cosine_topk_blocks(
        ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), bank.d, bank.w, 8,
        b_id=(1, bank.E), dtype=dtype, precision="tensor", block_n=block_n, want_dense=True)
    return dense.cpu().numpy()[:, :bank.E]


@pytest.mark.parametrize("E,d,w,block_n", [(300, 4, 512, 0), (300, 1, 512, 0), (300, 4, 512, 128),
                                            (300, 1, 512, 128), (515, 3, 1000, 0), (130, 2, 64, 128)])
def test_tensor_core_similarities_match_oracle(mb, ctx, E, d, w, block_n):
    bank, ref = _make_bank(mb, ctx, E, d, w, 40 * E, seed=E + d + w, empty=(3, E - 1))
    got = _dense(mb, ctx, bank, block_n=block_n)
    want = orc.bank_cosine_dense(ref)
    assert (np.isnan(got) == np.isnan(want)).all(), "NaN pattern (no comparable row) differs"
    m = ~np.isnan(want) & (want != 0)
    rel = np.abs(got[m] - want[m]) / np.abs(want[m])
    assert rel.max() <= REL_TOL, rel.max()
    assert (got[~np.isnan(want) & (want == 0)] == 0).all()
    bank.close()


def test_bf16_rows_within_bf16_tolerance(mb, ctx):
    bank, ref = _make_bank(mb, ctx, 260, 2, 512, 20000, seed=77)
    got = _dense(mb, ctx, bank, dtype="bf16")
    want = orc.bank_cosine_dense(ref)
    m = ~np.isnan(want) & (want != 0)
    assert (np.abs(got[m] - want[m]) / np.abs(want[m])).max() <= 2.0 ** -6
    bank.close()


@pytest.mark.parametrize("E,d,w,k", [(300, 4, 512, 10), (1000, 4, 4096, 50), (257, 1, 256, 100), (129, 2, 192, 5)])
def test_topk_rescored_equals_oracle_exactly(mb, ctx, E, d, w, k):
    from mahout_b200.sketch import last_fallback_rows
    bank, ref = _make_bank(mb, ctx, E, d, w, 60 * E, seed=3 * E + k, empty=(0, 17))
    idx, sim, cnt = bank.cosine_topk(k)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all()
    assert (idx == oidx).all(), f"{(idx != oidx).sum()} index mismatches"
    assert sim.tobytes() == osim.tobytes(), "re-scored similarities are not bit-equal to the oracle"
    assert last_fallback_rows(ctx) <= E
    bank.close()


def test_topk_tensor_precision_within_tolerance(mb, ctx):
    E, d, w, k = 400, 4, 1024, 20
    bank, ref = _make_bank(mb, ctx, E, d, w, 80 * E, seed=5)
    idx, sim, cnt = bank.cosine_topk(k, precision="tensor")
    dense = orc.bank_cosine_dense(ref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all()
    for r in range(E):
        for t in range(cnt[r]):
            c = idx[r, t]
            assert c != r and c >= 0
            assert abs(sim[r, t] - dense[r, c]) <= REL_TOL * abs(dense[r, c])
        # every returned item is within tolerance of the true k-th value
        if cnt[r] == k:
            assert sim[r, :k].min() >= osim[r, k - 1] * (1 - 2 * REL_TOL)
    bank.close()


def test_golden_fixture_topk(mb, ctx):
    g = json.load(open(os.path.join(GOLD, "sketch_small.json")))
    bank = mb.SketchBank(g["E"], g["w"], g["d"], mb.HashFunctionBuilder(g["seed"]), 1, ctx)
    bank.update(np.array(g["entity"]), np.array(g["key"]), np.array(g["inc"], np.float32))
    idx, sim, cnt = bank.cosine_topk(g["k"])
    assert cnt.tolist() == g["topk_cnt"]
    assert idx.tolist() == g["topk_idx"]
    assert np.allclose(sim, np.array(g["topk_sim"]), rtol=0, atol=0)
    bank.close()


def test_threshold_self_and_ties(mb, ctx):
    """identical sketches tie exactly: the lower index must win; threshold and self handling."""
    E, d, w, k = 140, 2, 128, 3
    bank = mb.SketchBank(E, w, d, 42, 1, ctx)
    ent = np.repeat(np.arange(E), 3).astype(np.int64)
    key = np.tile(np.array([5, 9, 11]), E).astype(np.int64)
    key[3 * 100:] += 1000                            # items 100.. use different users
    inc = np.tile(np.array([1.0, 2.0, 0.5], np.float32), E)
    bank.update(ent, key, inc)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((E, d, w))
    orc.bank_update(ref, d, w, a, b, ent, key, inc)
    for kwargs in (dict(), dict(exclude_self=False), dict(threshold=0.5)):
        idx, sim, cnt = bank.cosine_topk(k, **kwargs)
        oidx, osim, ocnt = orc.bank_cosine_topk(
            ref, k, threshold=kwargs.get("threshold", orc.NO_THRESHOLD),
            exclude_self=kwargs.get("exclude_self", True))
        assert (cnt == ocnt).all() and (idx == oidx).all(), kwargs
        assert sim.tobytes() == osim.tobytes()
    idx, _, _ = bank.cosine_topk(k)
    assert idx[0].tolist() == [1, 2, 3] and idx[2].tolist() == [0, 1, 3]
    bank.close()


def test_sharded_blocks_on_one_gpu(mb, ctx):
    """Item-hash sharding (owner = index % G) emulated on one GPU: shard banks are normalised
    separately, concatenated like an all-gather, and every shard computes its block row."""
    import torch
    from mahout_b200.sketch import cosine_topk_blocks
    E, d, w, k, G = 600, 4, 512, 12, 3
    full, ref = _make_bank(mb, ctx, E, d, w, 50 * E, seed=9, empty=(5,))
    counters = full.counters_tensor()                     # [E, d, w] int64
    per = E // G
    rows, valid, cnts = [], [], []
    for g in range(G):
        sh = mb.SketchBank(per, w, d, 42, 1, ctx)
        sh.counters_tensor().copy_(counters[g::G])
        torch.cuda.synchronize()
        r, v = sh.normalize()
        rows.append(r)
        valid.append(v)
        cnts.append(sh.counters_tensor().clone())
        sh.close()
    b_rows = torch.stack(rows)                            # [G, d, per, ld]
    b_valid = torch.stack(valid)
    b_cnt = torch.stack(cnts)                             # [G, per, d, w]
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    for g in range(G):
        idx, sim, cnt = cosine_topk_blocks(ctx, rows[g], valid[g], b_rows, b_valid, d, w, k,
                                           a_id=(G, g), b_id=(G, 1), precision="rescored",
                                           a_counters=cnts[g], b_counters=b_cnt)
        assert (cnt.cpu().numpy() == ocnt[g::G]).all()
        assert (idx.cpu().numpy() == oidx[g::G]).all()
        assert sim.cpu().numpy().tobytes() == osim[g::G].tobytes()
    full.close()


def test_bad_arguments(mb, ctx):
    bank = mb.SketchBank(10, 64, 2, 42, 1, ctx)
    with pytest.raises(ValueError):
        bank.cosine_topk(0)
    with pytest.raises(mb.NativeError):
        bank.cosine_topk(5000)          # beyond the fused top-k capacity AND what the multi-pass selection holds
    bank.close()


def _shards(mb, ctx, full, G):
    """shard banks (owner = index % G) of one bank -> rows [G,d,per,ld], valid [G,d,vw], counters [G,per,d,w]"""
    import torch
    E, d, w = full.E, full.d, full.w
    counters = full.counters_tensor()
    per = E // G
    rows, valid, cnts = [], [], []
    for g in range(G):
        sh = mb.SketchBank(per, w, d, 42, 1, ctx)
        sh.counters_tensor().copy_(counters[g::G])
        torch.cuda.synchronize()
        r, v = sh.normalize()
        rows.append(r)
        valid.append(v)
        cnts.append(sh.counters_tensor().clone())
        sh.close()
    return torch.stack(rows), torch.stack(valid), torch.stack(cnts)


@pytest.mark.parametrize("precision", ["tensor", "rescored"])
def test_incremental_job_equals_one_shot_and_oracle(mb, ctx, precision):
    """mb200_cosine_begin / push / finish: the B side pushed block by block (ring order), and as row
    chunks of all blocks (a chunked all-gather), gives the one-shot result; re-scored == oracle."""
    import torch
    from mahout_b200.sketch import CosineJob, cosine_topk_blocks
    E, d, w, k, G = 1536, 4, 512, 20, 3
    full, ref = _make_bank(mb, ctx, E, d, w, 40 * E, seed=21, empty=(7, 100))
    b_rows, b_valid, b_cnt = _shards(mb, ctx, full, G)
    per = E // G
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    for g in range(G):
        kw = dict(a_counters=b_cnt[g], b_counters=b_cnt) if precision == "rescored" else {}
        one = cosine_topk_blocks(ctx, b_rows[g], b_valid[g], b_rows, b_valid, d, w, k, a_id=(G, g), b_id=(G, 1),
                                 precision=precision, **kw)
        # (1) one push per peer block, starting with the local one (ring order)
        job = CosineJob(ctx, b_rows[g], b_valid[g], d, w, k, a_id=(G, g), precision=precision)
        for s in range(G):
            src = (g + s) % G
            job.push(b_rows[src:src + 1], b_valid[src:src + 1], id_mul=G, id_add=0, id_base=src)
        ring = job.finish(b_id=(G, 1), **kw)
        # (2) row chunks of all blocks (what a chunked all-gather delivers); chunk rows % 256 == 0
        job = CosineJob(ctx, b_rows[g], b_valid[g], d, w, k, a_id=(G, g), precision=precision)
        for c0 in range(0, per, 256):
            c1 = min(per, c0 + 256)
            rows_c = b_rows[:, :, c0:c1].contiguous()
            valid_c = b_valid[:, :, c0 // 32:(c1 + 31) // 32].contiguous()
            vw = int(mb._native.lib().mb200_valid_words(c1 - c0))
            vpad = torch.zeros((G, d, vw), dtype=torch.int32, device=valid_c.device)
            vpad[:, :, :valid_c.shape[2]] = valid_c
            job.push(rows_c, vpad, id_mul=G, id_add=1, id_base=c0 * G)
        chunks = job.finish(b_id=(G, 1), **kw)
        for got in (ring, chunks):
            for x, y in zip(got, one):
                assert torch.equal(x, y)
        if precision == "rescored":
            assert (one[2].cpu().numpy() == ocnt[g::G]).all()
            assert (one[0].cpu().numpy() == oidx[g::G]).all()
            assert one[1].cpu().numpy().tobytes() == osim[g::G].tobytes()
    full.close()


def test_job_misuse(mb, ctx):
    from mahout_b200.sketch import CosineJob
    bank = mb.SketchBank(256, 64, 2, 42, 1, ctx)
    rows, valid = bank.normalize()
    job = CosineJob(ctx, rows, valid, 2, 64, 4)
    with pytest.raises(ValueError):
        CosineJob(ctx, rows, valid, 2, 64, 4)          # one job per context
    with pytest.raises(ValueError):
        job.finish()                                   # nothing pushed
    job = CosineJob(ctx, rows, valid, 2, 64, 4)         # the failed finish freed the job
    job.push(rows.unsqueeze(0), valid.unsqueeze(0))
    idx, sim, cnt = job.finish()
    assert idx.shape == (256, 4)
    bank.close()


@pytest.mark.parametrize("E,d,w,k", [(1, 2, 64, 3), (2, 1, 8, 5), (5, 20, 40, 4), (200, 24, 128, 7), (131, 3, 4100, 6)])
def test_topk_edge_shapes(mb, ctx, E, d, w, k):
    """one or two entities, k beyond the number of entities, tiny and non-multiple-of-64 widths, wide rows
    (two-pass normalise), depths up to CountMinSketchConfig's 24"""
    bank, ref = _make_bank(mb, ctx, E, d, w, 30 * E + 10, seed=E + d + w + k)
    for precision in ("rescored", "tensor"):
        idx, sim, cnt = bank.cosine_topk(k, precision=precision)
        oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
        assert (cnt == ocnt).all()
        if precision == "rescored":
            assert (idx == oidx).all() and sim.tobytes() == osim.tobytes()
        else:
            m = oidx >= 0
            assert np.allclose(np.sort(sim, axis=1), np.sort(osim, axis=1), rtol=2e-3, atol=0)
    bank.close()


@pytest.mark.parametrize("E,d,w,k", [(300, 4, 512, 10), (1000, 4, 4096, 50), (257, 1, 256, 100), (129, 2, 192, 5)])
def test_topk_certified_sets_equal_oracle(mb, ctx, E, d, w, k):
    """MB200_PRECISION_CERTIFIED: the top-k SETS are the oracle's, the similarities are tensor-core values
    (<= 1e-3 relative), ordered by the returned value"""
    bank, ref = _make_bank(mb, ctx, E, d, w, 60 * E, seed=3 * E + k, empty=(0, 17))
    idx, sim, cnt = bank.cosine_topk(k, precision="certified")
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    dense = orc.bank_cosine_dense(ref)
    assert (cnt == ocnt).all()
    for r in range(E):
        c = cnt[r]
        assert set(idx[r, :c].tolist()) == set(oidx[r, :c].tolist()), r
        want = dense[r, idx[r, :c]]
        assert (np.abs(sim[r, :c] - want) <= REL_TOL * np.abs(want)).all()
        assert (np.diff(sim[r, :c]) <= 0).all()
    bank.close()


def test_certified_ties_threshold_and_sharded_blocks(mb, ctx):
    """exact ties straddling the cut, a threshold inside the value range, and the sharded block layout"""
    import torch
    from mahout_b200.sketch import cosine_topk_blocks
    E, d, w, k = 140, 2, 128, 3
    bank = mb.SketchBank(E, w, d, 42, 1, ctx)
    ent = np.repeat(np.arange(E), 3).astype(np.int64)
    key = np.tile(np.array([5, 9, 11]), E).astype(np.int64)
    key[3 * 100:] += 1000
    inc = np.tile(np.array([1.0, 2.0, 0.5], np.float32), E)
    bank.update(ent, key, inc)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((E, d, w))
    orc.bank_update(ref, d, w, a, b, ent, key, inc)
    idx, sim, cnt = bank.cosine_topk(k, precision="certified")
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all() and (idx == oidx).all()           # all ties: exact values decide, lower index wins
    bank.close()
    E, d, w, k, G = 600, 4, 512, 12, 3
    full, ref = _make_bank(mb, ctx, E, d, w, 50 * E, seed=9, empty=(5,))
    b_rows, b_valid, b_cnt = _shards(mb, ctx, full, G)
    dense = orc.bank_cosine_dense(ref)
    for thr in (None, 0.3):
        oidx, osim, ocnt = orc.bank_cosine_topk(ref, k, threshold=thr if thr else orc.NO_THRESHOLD)
        for g in range(G):
            idx, sim, cnt = cosine_topk_blocks(ctx, b_rows[g], b_valid[g], b_rows, b_valid, d, w, k, a_id=(G, g),
                                               b_id=(G, 1), threshold=thr, precision="certified",
                                               a_counters=b_cnt[g], b_counters=b_cnt)
            idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
            assert (cnt == ocnt[g::G]).all()
            for l in range(E // G):
                assert set(idx[l, :cnt[l]].tolist()) == set(oidx[g::G][l, :cnt[l]].tolist())
    full.close()


@pytest.mark.parametrize("variant", ["16", "MB200_RESCORE32", "MB200_RESCORE64", "wide_counters"])
def test_rescore_kernel_variants_bit_equal(mb, ctx, variant, monkeypatch):
    """the 16-bit biased, 32-bit and 64-bit forms of the exact re-score give the oracle's bits; counters
    beyond 2^15 quanta select the 32-bit form by themselves"""
    E, d, w, k = 500, 3, 1024, 25
    if variant.startswith("MB200"):
        monkeypatch.setenv(variant, "1")
    frac_bits = 12 if variant == "wide_counters" else 1            # 2^12 quanta per unit: counters > 2^15
    bank, ref = _make_bank(mb, ctx, E, d, w, 60 * E, seed=77, empty=(3,), frac_bits=frac_bits)
    if variant == "wide_counters":
        assert np.abs(ref).max() * 2 ** frac_bits >= 2 ** 15
    idx, sim, cnt = bank.cosine_topk(k)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all() and (idx == oidx).all()
    assert sim.tobytes() == osim.tobytes()
    bank.close()


@pytest.mark.parametrize("precision", ["rescored", "certified"])
@pytest.mark.parametrize("E,d,w,k", [(400, 4, 512, 20), (300, 1, 256, 50)])
def test_mixed_sign_counters_exact_sets(mb, ctx, E, d, w, k, precision):
    """Negative increments (mb200_bank_update accepts them; ingest's rating_shift produces them) give
    mixed-sign counters: the tensor-core error is then absolute, similarities cluster around zero and many rows
    have fewer than k positive ones.  The exact-set precisions must still return the oracle's sets."""
    rng = np.random.Generator(np.random.PCG64(E + k))
    n = 50 * E
    item = rng.integers(0, E, n).astype(np.int64)
    user = rng.integers(1, 2000, n).astype(np.int64)
    pref = (rng.integers(-10, 11, n) * 0.5).astype(np.float32)
    bank = mb.SketchBank(E, w, d, 42, 1, ctx)
    bank.update(item, user, pref)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((E, d, w))
    orc.bank_update(ref, d, w, a, b, item, user, pref)
    assert (ref < 0).any()
    idx, sim, cnt = bank.cosine_topk(k, precision=precision)
    assert bank.sign_info()
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert (cnt == ocnt).all()
    if precision == "rescored":
        assert (idx == oidx).all() and sim.tobytes() == osim.tobytes()
    else:
        for r in range(E):
            assert set(idx[r, :cnt[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()), r
        m = oidx >= 0
        # values are the tensor-core ones where the set was decided without re-scoring: absolute tolerance
        assert np.abs(np.sort(sim, axis=1) - np.sort(osim, axis=1))[m.any(axis=1)].max() <= 2e-3
    bank.close()


@pytest.mark.parametrize("precision", ["rescored", "certified"])
def test_boolean_data_tie_groups_take_the_exact_path(mb, ctx, precision, monkeypatch):
    """--booleanData (every preference 1.0) with few users: many items have IDENTICAL sketches, so whole groups of
    candidates tie exactly at the k-th value and the candidate lists cannot certify their rows -- they take the exact
    full-row path (k_exact_rows_fast + k_exact_topk).  Ties are broken by index, as in the oracle; the memory-speed
    integer form and the loop-for-loop FP64 form (MB200_EXACT_ROWS_SEQ) of that path must give the same answer."""
    from mahout_b200.sketch import last_band_rows, last_fallback_rows
    rng = np.random.Generator(np.random.PCG64(12))
    E, d, w, k = 900, 2, 256, 10
    users = 40
    item = np.repeat(np.arange(E - 5), 2).astype(np.int64)           # the last 5 items stay empty
    user = rng.integers(1, users + 1, item.shape[0]).astype(np.int64)
    pref = np.ones(item.shape[0], np.float32)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((E, d, w))
    orc.bank_update(ref, d, w, a, b, item, user, pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    results = []
    # (band pass on: uncertified rows are settled by a second targeted sweep), (band pass off: they all take the exact
    # full-row path, memory-speed form), (the same through the loop-for-loop FP64 form)
    for band, seq in ((True, False), (False, False), (False, True)):
        for name, on in (("MB200_NO_BAND", not band), ("MB200_EXACT_ROWS_SEQ", seq)):
            if on:
                monkeypatch.setenv(name, "1")
            else:
                monkeypatch.delenv(name, raising=False)
        bank = mb.SketchBank(E, w, d, 42, 1, ctx)
        bank.update(item, user, pref)
        idx, sim, cnt = bank.cosine_topk(k, precision=precision)
        fb, bd = last_fallback_rows(ctx), last_band_rows(ctx)
        bank.close()
        assert fb + bd > 0, "the tie groups were expected to defeat certification"
        assert (bd > 0) == band, (fb, bd)
        assert (cnt == ocnt).all()
        assert (idx == oidx).all(), f"{(idx != oidx).sum()} index mismatches (fallback rows {fb})"
        if precision == "rescored":
            assert sim.tobytes() == osim.tobytes()
        results.append((idx, sim))
    for other in results[1:]:
        assert (results[0][0] == other[0]).all()
        if precision == "rescored":
            assert results[0][1].tobytes() == other[1].tobytes()


@pytest.mark.parametrize("dtype,k", [("bf16", 20), ("f16", 192), ("bf16", 192)])
def test_bf16_rows_and_largest_k_keep_the_exact_contract(mb, ctx, dtype, k):
    """BF16 rows (8-bit significand: a wide undecided band) and k at the fused top-k capacity (CAP - 64 = 192, no
    re-scoring margin left): RESCORED must still be bit-equal to the oracle and CERTIFIED must return its sets --
    through more fallback rows if that is what it takes."""
    from mahout_b200.sketch import last_fallback_rows
    E, d, w = 700, 2, 512
    bank, ref = _make_bank(mb, ctx, E, d, w, 80 * E, seed=91 + k, empty=(5,))
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    idx, sim, cnt = bank.cosine_topk(k, dtype=dtype, precision="rescored")
    assert (cnt == ocnt).all() and (idx == oidx).all() and sim.tobytes() == osim.tobytes()
    cidx, csim, ccnt = bank.cosine_topk(k, dtype=dtype, precision="certified")
    assert (ccnt == ocnt).all()
    assert all(set(cidx[r, :ccnt[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()) for r in range(E))
    assert last_fallback_rows(ctx) <= E
    bank.close()


@pytest.mark.parametrize("k,threshold", [(193, None), (500, None), (700, 0.2)])
def test_k_beyond_the_fused_capacity(mb, ctx, k, threshold):
    """The reference takes any --maxSimilaritiesPerItem (ItemSimilarityJob.java:105).  Beyond the fused capacity (192)
    the candidates are collected 128 ranks at a time (row ceilings): re-scored == the oracle bit for bit, certified ==
    its sets, tensor values within tolerance; rows with fewer than k admissible columns return them all."""
    E, d, w = 1500, 2, 512
    bank, ref = _make_bank(mb, ctx, E, d, w, 70 * E, seed=7 + k, empty=(3, 900))
    kw = {} if threshold is None else {"threshold": threshold}
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k, **kw)
    idx, sim, cnt = bank.cosine_topk(k, precision="rescored", **({} if threshold is None else {"threshold": threshold}))
    assert (cnt == ocnt).all()
    assert (idx == oidx).all() and sim.tobytes() == osim.tobytes()
    cidx, csim, ccnt = bank.cosine_topk(k, precision="certified", **({} if threshold is None else {"threshold": threshold}))
    assert (ccnt == ocnt).all()
    assert all(set(cidx[r, :ccnt[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()) for r in range(E))
    tidx, tsim, tcnt = bank.cosine_topk(k, precision="tensor", **({} if threshold is None else {"threshold": threshold}))
    if threshold is None:
        assert (tcnt == ocnt).all()
    dense = orc.bank_cosine_dense(ref)
    for r in range(0, E, 97):
        for t in range(tcnt[r]):
            assert abs(tsim[r, t] - dense[r, tidx[r, t]]) <= REL_TOL * abs(dense[r, tidx[r, t]])
    bank.close()


@pytest.mark.parametrize("precision", ["certified", "rescored"])
def test_band_pass_settles_flat_similarity_rows(mb, ctx, precision, monkeypatch):
    """Rows whose similarities are nearly flat around the k-th value (few distinct users, coarse sketches, BF16 rows:
    the undecided band is wider than the candidate margin) cannot be certified from their lists; the band pass --
    a second K3 sweep over those rows with a fixed cut, every column above it re-scored exactly -- must return the
    oracle's answer without falling back to the exact full-row path."""
    from mahout_b200.sketch import last_band_rows, last_fallback_rows
    monkeypatch.setenv("MB200_MARGIN", "28")       # 128 candidates kept for k = 100: many rows cannot be certified
    rng = np.random.Generator(np.random.PCG64(5))
    E, d, w, k = 3000, 2, 128, 100                 # (203 band rows, some with more than 2048 columns above their cut)
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
    idx, sim, cnt = bank.cosine_topk(k, dtype="bf16", precision=precision)
    bd, fb = last_band_rows(ctx), last_fallback_rows(ctx)
    if precision == "certified":      # (the exact k-th value of the re-scored lists certifies every row of this data)
        assert bd > 0, "expected uncertified rows on this data"
    assert fb == 0, (bd, fb)
    assert (cnt == ocnt).all()
    if precision == "rescored":
        assert (idx == oidx).all() and sim.tobytes() == osim.tobytes()
    else:
        assert all(set(idx[r, :cnt[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()) for r in range(E))
    st = ctx.stats()
    assert st["band_rows_total"] >= bd
    bank.close()


def test_deferred_band_pass_of_a_streamed_job(mb, ctx, monkeypatch):
    """A job whose B side arrives in pieces cannot re-sweep it by itself: finish(defer_uncertified) returns "pending",
    the caller pushes the same pieces once more (K3 sweeps them for the uncertified rows only) and finishes again.
    Result == the one-shot call == the oracle's sets; nothing takes the exact full-row path."""
    import torch
    from mahout_b200.sketch import CosineJob, cosine_topk_blocks, last_band_rows, last_fallback_rows
    monkeypatch.setenv("MB200_MARGIN", "28")       # (see test_band_pass_settles_flat_similarity_rows)
    rng = np.random.Generator(np.random.PCG64(6))
    E, d, w, k = 1792, 2, 128, 100
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
    rows, valid = bank.normalize("bf16")
    cnt_t = bank.counters_tensor()
    one = cosine_topk_blocks(ctx, rows, valid, rows.unsqueeze(0), valid.unsqueeze(0), d, w, k, b_id=(1, E), dtype="bf16",
                             precision="certified", a_counters=cnt_t, b_counters=cnt_t)
    assert last_band_rows(ctx) > 0 and last_fallback_rows(ctx) == 0
    job = CosineJob(ctx, rows, valid, d, w, k, dtype="bf16", precision="certified")

    def pushes():
        for c0 in range(0, E, 512):
            c1 = min(E, c0 + 512)
            rc = rows[:, c0:c1].contiguous().unsqueeze(0)
            vw = int(mb._native.lib().mb200_valid_words(c1 - c0))
            vc = torch.zeros((1, d, vw), dtype=torch.int32, device=rows.device)
            words = valid[:, c0 // 32:(c1 + 31) // 32]
            vc[0, :, :words.shape[1]] = words
            job.push(rc, vc, id_mul=1, id_add=0, id_base=c0)

    pushes()
    res = job.finish(a_counters=cnt_t, b_counters=cnt_t.unsqueeze(0), b_id=(1, E), defer_uncertified=True)
    assert res is None and last_band_rows(ctx) > 0          # rows are waiting for the second round
    pushes()
    got = job.finish()
    assert last_fallback_rows(ctx) == 0
    gi, gc = got[0].cpu().numpy(), got[2].cpu().numpy()
    assert (gc == ocnt).all() and (gc == one[2].cpu().numpy()).all()
    assert all(set(gi[r, :gc[r]].tolist()) == set(oidx[r, :ocnt[r]].tolist()) for r in range(E))
    bank.close()
