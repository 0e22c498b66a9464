"""Pins the CPU oracle against (a) every golden vector the reference's own tests hold
for this path, (b) the Java SE definitions it depends on (java.util.Random,
BigInteger.mod) re-derived independently in pure Python, (c) committed fixtures."""
import json
import math
import os

import numpy as np
import pytest

import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")
P = orc.PRIME


# ---- independent pure-Python restatements (Java SE spec) -------------------
def _java_random_longs(seed, n):
    mask = (1 << 48) - 1
    s = (seed ^ 0x5DEECE66D) & mask
    out = []

    def nxt32():
        nonlocal s
        s = (s * 0x5DEECE66D + 0xB) & mask
        v = s >> 16
        return v - (1 << 32) if v >= (1 << 31) else v

    for _ in range(n):
        hi, lo = nxt32(), nxt32()
        v = ((hi << 32) + lo) & ((1 << 64) - 1)
        out.append(v - (1 << 64) if v >= (1 << 63) else v)
    return out


def _py_hash(a, b, w, k):
    return ((a * k + b) % P) % w  # Python % == BigInteger.mod for positive modulus


def test_java_random_known_answers():
    # well-known JDK values
    assert orc.lib().orc_java_random_next_int(42) == -1170105035
    assert orc.lib().orc_java_random_next_long(42) == -5025562857975149833
    assert _java_random_longs(42, 1)[0] == -5025562857975149833


def test_prime_is_2_63_minus_25():
    assert P == 2 ** 63 - 25


@pytest.mark.parametrize("seed", [0, 42, -7, 20240002, 2 ** 62 + 12345])
def test_hash_params_match_java_random(seed):
    a, b = orc.hash_params(seed, 6)
    longs = _java_random_longs(seed, 12)
    for i in range(6):
        assert int(a[i]) == abs(longs[2 * i]) or longs[2 * i] == -2 ** 63
        assert int(b[i]) == abs(longs[2 * i + 1]) or longs[2 * i + 1] == -2 ** 63


def test_hash_params_survey_anchors():
    a, b = orc.hash_params(42, 4)
    assert [int(x) for x in a] == [5025562857975149833, 5694868678511409995,
                                   6169532649852302182, 6802844026563419272]
    assert [int(x) for x in b] == [5843495416241995736, 5111195811822994797,
                                   1782466964123969572, 5086654115216342560]
    a0, b0 = orc.hash_params(0, 1)
    assert (int(a0[0]), int(b0[0])) == (4962768465676381896, 4437113781045784766)


def test_hash_survey_anchors():
    keys = [0, 1, 2, 1682, 10 ** 12, -1]
    a, b = orc.hash_params(42, 4)
    want = [[671704, 767226, 862723, 888455, 1047010, 576207],
            [820589, 540881, 261148, 97706, 19899, 51721],
            [682020, 613258, 544521, 395821, 39671, 750757],
            [1031712, 203457, 423778, 404705, 700740, 811391]]
    for i in range(4):
        assert [orc.hash_one(a[i], b[i], 1 << 20, k) for k in keys] == want[i]
    a0, b0 = orc.hash_params(0, 1)
    assert [orc.hash_one(a0[0], b0[0], 1 << 20, k) for k in keys] == \
        [618686, 8095, 446055, 141103, 715017, 180701]


def test_hash_matches_bigint_semantics_random_and_edges():
    rng = np.random.default_rng(1)
    a, b = orc.hash_params(42, 4)
    edge = [0, 1, -1, 2 ** 63 - 1, -2 ** 63, P, P - 1, P + 1, -P, -P - 1, 2 ** 62, -2 ** 62, 25, -25]
    keys = edge + [int(x) for x in rng.integers(-2 ** 63, 2 ** 63 - 1, 2000, dtype=np.int64)]
    for w in (1 << 20, 4096, 1000003, 7, 1, 2 ** 31 - 1):
        for i in range(4):
            got = orc.hash_many(a[i], b[i], w, np.array(keys, np.int64))
            want = [_py_hash(int(a[i]), int(b[i]), w, k) for k in keys]
            assert got.tolist() == want
    # degenerate parameters Math.abs(Long.MIN_VALUE) < 0 must still follow BigInteger.mod
    for aa, bb in [(-2 ** 63, 5), (7, -2 ** 63), (-2 ** 63, -2 ** 63)]:
        for k in edge:
            assert orc.hash_one(aa, bb, 4096, k) == _py_hash(aa, bb, 4096, k)


def test_cm_dims():
    # AbstractCountMinSketch.java:69-83
    assert orc.cm_dims(0.01, 0.01) == (math.ceil(math.e / 0.01), math.ceil(math.log(100)))
    with pytest.raises(ValueError):
        orc.cm_dims(0.5, 0.1)      # delta > 1/e
    with pytest.raises(ValueError):
        orc.cm_dims(0.0, 0.1)
    with pytest.raises(ValueError):
        orc.cm_dims(0.1, 3.0)      # epsilon > e
    assert orc.cm_dims(math.exp(-1), math.e) == (1, 1)


def test_reference_cosine_known_answer():
    # VectorSimilarityMeasuresTest.testCosineSimilarity (:107-114) -> 0.769846046 +- 1e-6
    x = [0, 2, 0, 0, 8, 3, 0, 6, 0, 1, 2, 2, 0]
    y = [3, 0, 0, 0, 7, 0, 2, 2, 1, 3, 2, 1, 1]
    assert abs(orc.exact_cosine(x, y) - 0.769846046) < 1e-6


def test_sketch_cosine_equals_exact_when_collision_free():
    # SURVEY 8c cross-check: wide sketch, no collisions => sketch cosine == exact cosine
    x = [0, 2, 0, 0, 8, 3, 0, 6, 0, 1, 2, 2, 0]
    y = [3, 0, 0, 0, 7, 0, 2, 2, 1, 3, 2, 1, 1]
    w, d = 1 << 16, 4
    a, b = orc.hash_params(42, d)
    for i in range(d):  # keys 0..12 must not collide in any row for this check to be valid
        assert len(set(orc.hash_many(a[i], b[i], w, np.arange(13)).tolist())) == 13
    ca = np.zeros((d, w))
    cb = np.zeros((d, w))
    orc.cm_update(ca, w, d, a, b, np.arange(13), np.array(x, float))
    orc.cm_update(cb, w, d, a, b, np.arange(13), np.array(y, float))
    assert abs(orc.cm_cosine(ca, cb, w, d) - 0.769846046) < 1e-6
    # point queries return the exact values when collision-free
    for k in range(13):
        assert orc.cm_get(ca, w, d, a, b, k) == x[k]


def test_cm_cosine_nan_and_skip_rows():
    w, d = 8, 3
    z = np.zeros((d, w))
    o = np.zeros((d, w))
    o[:, 1] = 2.0
    assert math.isnan(orc.cm_cosine(z, o, w, d))          # every row has zero denominator
    p = o.copy()
    p[1, :] = 0.0                                          # row 1 skipped, others cos == 1
    assert orc.cm_cosine(o, p, w, d) == pytest.approx(1.0)
    q = np.zeros((d, w))
    q[:, 2] = 1.0
    assert orc.cm_cosine(o, q, w, d) == 0.0                # orthogonal


def test_clamp():
    assert orc.clamp_similarity(1.0000000002) == 1.0
    assert orc.clamp_similarity(-1.5) == -1.0
    assert orc.clamp_similarity(0.25) == 0.25
    assert math.isnan(orc.clamp_similarity(float("nan")))


def _csr_from_prefs(prefs):
    """prefs: list of (user, item, value) -> item x user CSR with dense indices."""
    users = sorted({u for u, _, _ in prefs})
    items = sorted({i for _, i, _ in prefs})
    uidx = {u: n for n, u in enumerate(users)}
    rows = {i: {} for i in items}
    for u, i, v in prefs:
        rows[i][uidx[u]] = v
    rowptr, colidx, vals = [0], [], []
    for i in items:
        for c in sorted(rows[i]):
            colidx.append(c)
            vals.append(rows[i][c])
        rowptr.append(len(colidx))
    return items, len(users), np.array(rowptr), np.array(colidx, np.int32), np.array(vals, np.float32)


def test_item_similarity_job_complete_job():
    # ItemSimilarityJobTest.testCompleteJob (:113-173): exactly two lines,
    # 1\t3\t~0.45 and 2\t3\t~0.89 (+-0.01)
    prefs = [(2, 1, 1), (1, 2, 1), (3, 4, 1), (1, 3, 2), (2, 3, 1)]
    items, ncols, rowptr, colidx, vals = _csr_from_prefs(prefs)
    idx, sim, cnt = orc.rowsim_cosine_topk(len(items), ncols, rowptr, colidx, vals, 100)
    out = orc.most_similar_item_pairs(idx, sim, cnt, items)
    assert len(out) == 2
    assert out[0][:2] == (1, 3) and abs(out[0][2] - 0.45) < 0.01
    assert out[1][:2] == (2, 3) and abs(out[1][2] - 0.89) < 0.01
    assert out[0][2] == pytest.approx(1 / math.sqrt(5), abs=1e-12)
    assert out[1][2] == pytest.approx(2 / math.sqrt(5), abs=1e-12)


def test_most_similar_pairs_mapper_and_reducer():
    # testMostSimilarItemsPairsMapper (:55-81): row 34 has {12:0.2, 56:0.9}, k=1 -> (34,56,0.9)
    idx = np.array([[2]])
    sim = np.array([[0.9]])
    cnt = np.array([1])
    ids = {0: 34, 2: 56, 1: 12}
    assert orc.most_similar_item_pairs(idx, sim, cnt, ids) == [(34, 56, 0.9)]
    # testMostSimilarItemPairsReducer (:86-99): duplicates collapse to one line
    idx = np.array([[1], [0]])
    sim = np.array([[0.5], [0.5]])
    cnt = np.array([1, 1])
    assert orc.most_similar_item_pairs(idx, sim, cnt, {0: 123, 1: 456}) == [(123, 456, 0.5)]


def test_topk_total_order_and_positive_only():
    # three identical rows + one orthogonal + one empty: ties broken by lower index,
    # zero / NaN similarities never reported (TopElementsQueue sentinel Double.MIN_VALUE)
    w, d = 64, 2
    a, b = orc.hash_params(7, d)
    bank = np.zeros((5, d, w))
    for e in (0, 1, 2):
        orc.cm_update(bank[e], w, d, a, b, [10, 11], [1.0, 2.0])
    orc.cm_update(bank[3], w, d, a, b, [500], [1.0])
    idx, sim, cnt = orc.bank_cosine_topk(bank, 1)
    assert idx[:, 0].tolist()[:3] == [1, 0, 0]
    assert cnt.tolist()[:3] == [1, 1, 1]
    assert cnt[4] == 0 and idx[4, 0] == -1          # empty sketch: all NaN
    idx2, sim2, cnt2 = orc.bank_cosine_topk(bank, 4)
    assert idx2[0, :2].tolist() == [1, 2] and cnt2[0] in (2, 3)
    assert np.all(sim2[cnt2[:, None] > np.arange(4)[None, :]] > 0)


def test_id_to_index_range():
    # TasteHadoopUtilsTest.java:26-39
    for v in (0, 1, -1, 2 ** 31, 2 ** 63 - 1, -2 ** 63, 123456789012345):
        i = orc.id_to_index(v)
        assert 0 <= i < 2 ** 31 - 1
    assert orc.id_to_index(5) == 5
    h = lambda v: ((v ^ ((v % 2 ** 64) >> 32)) + 2 ** 31) % 2 ** 32 - 2 ** 31   # Longs.hashCode
    for v in (12345678912, -99, 2 ** 40 + 17):
        hv = h(v % 2 ** 64)
        m = int(math.fmod(hv, 0x7FFFFFFE))
        assert orc.id_to_index(v) == (0x7FFFFFFF & (m % 2 ** 32))


def test_bank_update_mt_equals_sequential():
    rng = np.random.default_rng(3)
    n, E, d, w = 20000, 7, 3, 256
    a, b = orc.hash_params(42, d)
    ent = rng.integers(0, E, n)
    key = rng.integers(-10 ** 6, 10 ** 6, n)
    inc = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
    b1 = np.zeros((E, d, w))
    b2 = np.zeros((E, d, w))
    orc.bank_update(b1, d, w, a, b, ent, key, inc)
    orc.bank_update(b2, d, w, a, b, ent, key, inc, nthreads=4)
    assert np.array_equal(b1, b2)
    assert b1.sum() == pytest.approx(inc.astype(np.float64).sum() * d)


def test_golden_fixture():
    """Committed fixture (tests/golden/make_golden.py): freezes the oracle's outputs so a
    later edit of the oracle cannot silently drift."""
    path = os.path.join(GOLD, "sketch_small.json")
    g = json.load(open(path))
    a, b = orc.hash_params(g["seed"], g["d"])
    assert [int(x) for x in a] == g["a"] and [int(x) for x in b] == g["b"]
    bank = np.zeros((g["E"], g["d"], g["w"]))
    orc.bank_update(bank, g["d"], g["w"], a, b, np.array(g["entity"]), np.array(g["key"]),
                    np.array(g["inc"], np.float32))
    nz = np.flatnonzero(bank)
    assert nz.tolist() == g["nonzero_cells"]
    assert bank.ravel()[nz].tolist() == g["nonzero_values"]
    idx, sim, cnt = orc.bank_cosine_topk(bank, g["k"])
    assert idx.tolist() == g["topk_idx"]
    assert cnt.tolist() == g["topk_cnt"]
    assert np.allclose(sim, np.array(g["topk_sim"]), rtol=0, atol=1e-15)


@pytest.mark.parametrize("E,d,w,k,neg", [(150, 4, 256, 10, False), (90, 1, 64, 100, False), (120, 3, 100, 7, True)])
def test_fast_oracle_equals_loop_oracle_bit_for_bit(E, d, w, k, neg):
    """oracle/fast.py (blocked FP64 products over integer quanta; the checker of bench.py's config-4 / config-5
    parity legs) against the loop-for-loop C oracle: indices, counts and similarities bit-equal -- including tie
    groups (boolean-style data), empty rows (NaN) and column blocks smaller than k."""
    from oracle import fast
    rng = np.random.Generator(np.random.PCG64(E + k))
    n = 30 * E
    ent = rng.integers(0, E - 2, n).astype(np.int64)            # the last two entities stay empty
    key = rng.integers(0, 40 if k == 100 else 5000, n).astype(np.int64)   # few keys -> many exact ties
    inc = (rng.integers(-6 if neg else 1, 11, n) * 0.5).astype(np.float32)
    if k == 100:
        inc[:] = 1.0
    a, b = orc.hash_params(42, d)
    bank = np.zeros((E, d, w))
    orc.bank_update(bank, d, w, a, b, ent, key, inc)
    want_i, want_s, want_c = orc.bank_cosine_topk(bank, k)
    q = np.rint(bank * 2).astype(np.int32)                       # quanta at frac_bits = 1
    rows = np.arange(E)
    for block in (8192, 37):
        gi, gs, gc = fast.bank_rows_topk(q, rows, k, block=block)
        assert (gc == want_c).all()
        assert (gi == want_i).all()
        assert gs.tobytes() == want_s.tobytes()
    # columns split over "ranks", partial top-k lists merged: the distributed form bench.py uses
    parts = []
    for g in range(3):
        cols = np.arange(g, E, 3)
        parts.append(fast.rows_vs_columns_topk(q[rows], rows, q[cols], cols, k, block=50))
    gi, gs, gc = fast.merge_partials(parts, k)
    assert (gi == want_i).all() and gs.tobytes() == want_s.tobytes() and (gc == want_c).all()
    thr = float(np.median(want_s[want_s > 0]))
    ti, ts, tc = orc.bank_cosine_topk(bank, k, threshold=thr)
    gi, gs, gc = fast.bank_rows_topk(q, rows, k, threshold=thr, block=64)
    assert (gi == ti).all() and gs.tobytes() == ts.tobytes() and (gc == tc).all()
