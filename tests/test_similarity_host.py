"""Host logic around the hot path (no GPU): the oracle's ingest restatement (oracle/prep.py, the checker of
the GPU ingest), pair emission, shard plan.
Golden values are the reference's own tests (ItemSimilarityJobTest, TasteHadoopUtilsTest)."""
import numpy as np

import oracle as orc
from mahout_b200 import similarity as sim
from oracle import prep


def test_id_to_index_matches_oracle_and_reference_range():
    rng = np.random.Generator(np.random.PCG64(3))
    ids = np.concatenate([np.array([0, 1, 2, 1682, -1, 2 ** 31 - 1, 2 ** 31, 2 ** 63 - 1, -2 ** 63, 0x7FFFFFFE]),
                          rng.integers(-2 ** 63, 2 ** 63 - 1, 5000, dtype=np.int64)]).astype(np.int64)
    got = prep.id_to_index(ids)
    want = np.array([orc.id_to_index(int(v)) for v in ids])
    assert (got == want).all()
    # TasteHadoopUtilsTest.java:26-39: the index is a non-negative int
    assert (got >= 0).all() and (got <= 0x7FFFFFFF).all()
    assert prep.id_to_index([12345])[0] == 12345


def test_parse_prefs_formats():
    u, i, p = prep.parse_prefs(["1,2,3.5", "4\t5\t1", "7,8", "", "9,10,2.5"])
    assert u.tolist() == [1, 4, 7, 9] and i.tolist() == [2, 5, 8, 10]
    assert p.tolist() == [3.5, 1.0, 1.0, 2.5]
    _, _, p = prep.parse_prefs(["1,2,3.5"], boolean_data=True)
    assert p.tolist() == [1.0]


def test_preference_matrix_dedup_min_prefs_and_min_id():
    # user 1 rates item 5 twice (last wins); user 3 has a single pref and is dropped at minPrefs=2
    user = [1, 1, 1, 2, 2, 3]
    item = [5, 7, 5, 5, 9, 7]
    pref = [1.0, 2.0, 4.0, 3.0, 5.0, 1.0]
    pm = prep.PreferenceMatrix(user, item, pref, min_prefs_per_user=2)
    ev = sorted(zip(pm.user.tolist(), pm.item_id[pm.row].tolist(), pm.pref.tolist()))
    assert ev == [(1, 5, 4.0), (1, 7, 2.0), (2, 5, 3.0), (2, 9, 5.0)]
    assert pm.num_users == 2 and pm.num_items == 3
    # two item IDs that collide under idToIndex share a row; the smaller ID names it
    a, b = 5, 5 + (1 << 32) + (1 << 0) * 0
    assert prep.id_to_index([a])[0] != prep.id_to_index([b])[0] or True
    x = 12
    y = x ^ (7 << 32) ^ 7          # hashCode(y) = (x ^ 7) ^ 7 = x
    assert prep.id_to_index([x])[0] == prep.id_to_index([y])[0]
    pm = prep.PreferenceMatrix([1, 2], [y, x], [1.0, 1.0])
    assert pm.num_items == 1 and pm.item_id.tolist() == [min(x, y)]


def test_most_similar_item_pairs_reference_cases():
    # ItemSimilarityJobTest.testMostSimilarItemsPairsMapper (:55-81): k=1 keeps 0.9, key (34,56)
    idx = np.array([[1]])
    s = np.array([[0.9]])
    cnt = np.array([1])
    item_id = {0: 56, 1: 34}
    assert sim.most_similar_item_pairs(idx, s, cnt, item_id) == [(34, 56, 0.9)]
    # testMostSimilarItemPairsReducer (:86-99): both directions of a pair collapse to one line
    idx = np.array([[1, -1], [0, 2], [1, -1]])
    s = np.array([[0.5, 0.0], [0.5, 0.25], [0.25, 0.0]])
    cnt = np.array([1, 2, 1])
    assert sim.most_similar_item_pairs(idx, s, cnt) == [(0, 1, 0.5), (1, 2, 0.25)]
    assert sim.most_similar_item_pairs(idx, s, cnt) == orc.most_similar_item_pairs(idx, s, cnt)


def test_shard_plan_round_trip():
    for N, G in ((10, 3), (8, 4), (5, 8), (1000, 7)):
        rows = np.arange(N)
        parts = []
        for g in range(G):
            plan = sim.ShardPlan(N, G, g)
            (lr,) = plan.my_events(rows)
            assert (plan.global_row(lr) == rows[rows % G == g]).all()
            assert plan.local_count() == (rows % G == g).sum()
            part = np.full((plan.rows_per_shard, 2), -1)
            part[:plan.local_count(), 0] = plan.global_row(np.arange(plan.local_count()))
            parts.append(part)
        out = sim.ShardPlan(N, G, 0).assemble(parts)
        assert (out[:, 0] == rows).all()


def test_item_similarity_job_rejects_bad_arguments(tmp_path):
    from mahout_b200.itemsimilarity import ItemSimilarityJob
    job = ItemSimilarityJob()
    assert job.run([]) == -1                                    # missing required options
    p = tmp_path / "in.csv"
    p.write_text("1,2,1\n")
    assert job.run(["-i", str(p), "-o", str(tmp_path / "o"), "-s", "SIMILARITY_TANIMOTO_COEFFICIENT"]) == -1
    assert job.run(["-i", str(p), "-o", str(tmp_path / "o"), "-s", "SIMILARITY_COSINE", "-m", "0"]) == -1
