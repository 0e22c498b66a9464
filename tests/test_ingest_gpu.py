"""GPU ingest (csrc/ingest.cu through the C ABI) against the oracle's restatement (oracle/prep.py) of
ToEntityPrefsMapper / idToIndex / ItemIDIndexReducer / ToUserVectorsReducer.  Bit-exact: integers, and
floats parsed with Float.parseFloat's correct rounding."""
import numpy as np
import pytest

import oracle as orc
from oracle import prep as oprep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ing():
    from mahout_b200 import ingest
    return ingest


def _same(ev, want):
    u, i, p = ev.read()
    assert u.tolist() == want[0].tolist()
    assert i.tolist() == want[1].tolist()
    assert p.tobytes() == want[2].tobytes()


def test_parse_formats_match_reference_semantics(ing):
    lines = ["1,2,3.5", "4\t5\t1", "7,8", "", "9,10,2.5f", "1,2,", "3,4,0.1,999", "-5,+6,4", "11,12,5\r",
             "9223372036854775807,-9223372036854775808,0.5", "13,14,1e-3", "15,16,0.30000001192092896",
             "17,18,16777217", "19,20, 4.5 ", "21,22,.5", "23,24,5.", "25,26,0x1.8p1", "27,28,NaN", "29,30,-Infinity",
             "31,32,3.4028235e38", "33,34,1.17549435E-38d", "35,36,,,", "37,38\t4\t881250949"]
    text = "\n".join(lines) + "\n"
    want = oprep.parse_prefs(lines)
    ev = ing.Events.parse(text)
    u, i, p = ev.read()
    assert u.tolist() == want[0].tolist() and i.tolist() == want[1].tolist()
    assert p.view(np.uint32).tolist() == want[2].view(np.uint32).tolist()      # NaN-safe bit compare
    # no trailing newline, CRLF line ends, booleanData, ratingShift, transpose
    ev2 = ing.Events.parse("\r\n".join(l.rstrip("\r") for l in lines if l))
    assert ev2.read()[2].view(np.uint32).tolist() == want[2].view(np.uint32).tolist()
    _same(ing.Events.parse(text, boolean_data=True), oprep.parse_prefs(lines, boolean_data=True))
    plain = [l for l in lines if "NaN" not in l and "Inf" not in l]
    _same(ing.Events.parse("\n".join(plain), rating_shift=-2.5), oprep.parse_prefs(plain, rating_shift=-2.5))
    t = ing.Events.parse(text, transpose=True).read()
    assert t[0].tolist() == want[1].tolist() and t[1].tolist() == want[0].tolist()
    assert len(ing.Events.parse("")) == 0 and len(ing.Events.parse("\n\n\r\n")) == 0


def test_parse_large_random_text_and_tile_boundaries(ing):
    import torch
    rng = np.random.Generator(np.random.PCG64(11))
    n = 300_000
    user = rng.integers(-10 ** 12, 10 ** 12, n)
    item = rng.integers(1, 10 ** 6, n)
    kinds = rng.integers(0, 6, n)
    prefs = rng.integers(1, 11, n) * 0.5
    odd = rng.random(n)
    lines = []
    for t in range(n):
        k = kinds[t]
        sep = "," if t % 3 else "\t"
        if k == 0:
            lines.append(f"{user[t]}{sep}{item[t]}")
        elif k == 1:
            lines.append(f"{user[t]}{sep}{item[t]}{sep}{prefs[t]}")
        elif k == 2:
            lines.append(f"{user[t]}{sep}{item[t]}{sep}{float(odd[t])!r}")                # 16-17 digit decimals: host fix-up path
        elif k == 3:
            lines.append(f"{user[t]}{sep}{item[t]}{sep}{odd[t]:.6f}{sep}{88125094 + t}")
        elif k == 4:
            lines.append(f"{user[t]}{sep}{item[t]}{sep}{int(prefs[t])}{sep}" + "x" * int(odd[t] * 1500))  # long lines
        else:
            lines.append(f"{user[t]}{sep}{item[t]}{sep}{odd[t] * 1e-30:.8e}")
    text = "\n".join(lines) + "\n"
    want = oprep.parse_prefs(lines)
    _same(ing.Events.parse(text), want)
    # device-resident text at an odd address (byte-load path)
    raw = torch.frombuffer(bytearray(b"\n" + text.encode()), dtype=torch.uint8).cuda()
    _same(ing.Events.parse(raw[1:]), want)
    _same(ing.Events.parse(raw), want)


@pytest.mark.parametrize("bad,what", [("abc,1,2", "NumberFormatException"), ("1", "ArrayIndexOutOfBounds"),
                                      ("1,2,,5", "NumberFormatException"), ("1,2,x.5", "NumberFormatException"),
                                      ("9223372036854775808,1", "NumberFormatException"), ("1, 2,3", "NumberFormatException"),
                                      ("1,2x,3", "NumberFormatException")])
def test_parse_malformed_lines_fail_like_the_mapper(ing, bad, what):
    text = "1,2,3\n4,5,6\n" + bad + "\n7,8,9\n"
    with pytest.raises(ValueError) as e:
        ing.Events.parse(text)
    assert what in str(e.value)
    if "x.5" not in bad:
        assert "byte offset 12" in str(e.value)


def test_id_to_index_matches_oracle(ing):
    rng = np.random.Generator(np.random.PCG64(3))
    ids = np.concatenate([np.array([0, 1, 2, 1682, -1, 2 ** 31 - 1, 2 ** 31, 2 ** 63 - 1, -2 ** 63, 0x7FFFFFFE]),
                          rng.integers(-2 ** 63, 2 ** 63 - 1, 5000, dtype=np.int64)]).astype(np.int64)
    got = ing.id_to_index(ids)
    assert (got == np.array([orc.id_to_index(int(v)) for v in ids])).all()
    assert (got == oprep.id_to_index(ids)).all()


def _check_prepare(ing, user, item, pref, min_prefs):
    want = oprep.PreferenceMatrix(user, item, pref, min_prefs)
    ev = ing.Events.from_arrays(np.asarray(user, np.int64), np.asarray(item, np.int64), np.asarray(pref, np.float32))
    pm = ev.prepare(min_prefs)
    assert (pm.num_items, pm.num_users, pm.n) == (want.num_items, want.num_users, want.user.shape[0])
    assert pm.index_values.tolist() == want.index_values.tolist()
    assert pm.item_id.tolist() == want.item_id.tolist()
    got = sorted(zip(pm.user.cpu().tolist(), pm.row.cpu().tolist(), pm.pref.cpu().tolist()))
    exp = sorted(zip(want.user.tolist(), want.row.tolist(), want.pref.tolist()))
    assert got == exp
    # survivors keep the input order
    u, i, p = ev.read()
    pos = {}
    for t in range(len(u)):
        pos[(int(u[t]), int(oprep.id_to_index([i[t]])[0]))] = t
    order = [pos[(uu, int(pm.index_values[rr]))] for uu, rr in zip(pm.user.cpu().tolist(), pm.row.cpu().tolist())]
    assert order == sorted(order)
    pm.close()
    ev.close()


def test_prepare_small_reference_cases(ing):
    # user 1 rates item 5 twice (last wins); user 3 has a single pref and is dropped at minPrefs=2
    _check_prepare(ing, [1, 1, 1, 2, 2, 3], [5, 7, 5, 5, 9, 7], [1.0, 2.0, 4.0, 3.0, 5.0, 1.0], 2)
    _check_prepare(ing, [1, 1, 1, 2, 2, 3], [5, 7, 5, 5, 9, 7], [1.0, 2.0, 4.0, 3.0, 5.0, 1.0], 1)
    # two item IDs that collide under idToIndex share a row; the smaller ID names it
    x = 12
    y = x ^ (7 << 32) ^ 7
    _check_prepare(ing, [1, 2, 2], [y, x, y], [1.0, 1.0, 3.0], 1)
    # the user whose bit pattern is the hash tables' empty marker, and Long.MIN_VALUE
    _check_prepare(ing, [-1, -1, -2 ** 63, 5, -1], [3, 4, 3, 3, 3], [1.0, 2.0, 3.0, 4.0, 5.0], 2)


def test_prepare_zero_preferences_are_removed(ing):
    """userVector.set(index, 0.0) removes the element (RandomAccessSparseVector.setQuick): a pair whose LAST
    preference is 0.0 (rating + ratingShift == 0) neither counts toward minPrefsPerUser nor survives; an earlier
    0.0 overwritten by a later value does not matter."""
    user = [1, 1, 1, 2, 2, 2, 3, 3]
    item = [5, 7, 5, 5, 9, 7, 7, 7]
    pref = [1.0, 2.0, 0.0, 3.0, 0.0, 1.0, 0.0, 4.0]     # user 1 ends with 1 element, user 2 with 2, user 3 with 1
    for mp in (1, 2):
        _check_prepare(ing, user, item, pref, mp)
    rng = np.random.Generator(np.random.PCG64(15))
    n = 50_000
    u = rng.integers(1, 2000, n)
    i = rng.integers(1, 500, n).astype(np.int64)
    p = (rng.integers(-2, 9, n) * 0.5).astype(np.float32)      # ~9 % zeros, some negatives
    _check_prepare(ing, u, i, p, 1)
    _check_prepare(ing, u, i, p, 20)


def test_prepare_random_with_duplicates_and_collisions(ing):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 200_000
    user = rng.integers(1, 3000, n)
    base = rng.integers(1, 800, n)
    # a third of the items are replaced by an ID that collides with another item's index
    item = np.where(rng.random(n) < 0.33, base ^ (9 << 32) ^ 9, base).astype(np.int64)
    pref = (rng.integers(1, 11, n) * 0.5).astype(np.float32)
    _check_prepare(ing, user, item, pref, 1)
    _check_prepare(ing, user, item, pref, 60)


def test_item_similarity_job_from_text_equals_oracle_pipeline(ing, tmp_path):
    """text file -> GPU parse -> GPU prepare -> K1 -> cosine top-k -> pairs, against the same pipeline
    restated on the host (oracle/)."""
    from mahout_b200.itemsimilarity import ItemSimilarityJob
    rng = np.random.Generator(np.random.PCG64(8))
    n, U, I, k, d, w = 30000, 500, 300, 10, 4, 1024
    user = rng.integers(1, U, n)
    item = rng.integers(1, I, n) * 7
    pref = rng.integers(1, 11, n) * 0.5
    lines = [f"{u},{i},{p}" for u, i, p in zip(user, item, pref)]
    inp = tmp_path / "prefs.csv"
    inp.write_text("\n".join(lines) + "\n")
    out = tmp_path / "out.tsv"
    rc = ItemSimilarityJob().run(["-i", str(inp), "-o", str(out), "-s", "SIMILARITY_COSINE", "-m", str(k), "-mp", "3",
                                  "--sketchWidth", str(w), "--sketchDepth", str(d)])
    assert rc == 0
    got = [(int(a), int(b), float(s)) for a, b, s in (ln.split("\t") for ln in out.read_text().splitlines())]
    pm = oprep.PreferenceMatrix(*oprep.parse_prefs(lines), min_prefs_per_user=3)
    a, b = orc.hash_params(42, d)
    ref = np.zeros((pm.num_items, d, w))
    orc.bank_update(ref, d, w, a, b, pm.row, pm.user, pm.pref)
    oidx, osim, ocnt = orc.bank_cosine_topk(ref, k)
    assert got == orc.most_similar_item_pairs(oidx, osim, ocnt, pm.item_id)


@pytest.mark.parametrize("measure", ["SIMILARITY_COSINE", "SIMILARITY_SKETCH_COSINE"])
def test_native_cli_equals_python_job(tmp_path, measure):
    """the C++ host driver (mahout_b200_itemsimilarity) and the Python mirror write the same file, byte for byte"""
    import subprocess
    from mahout_b200 import build
    from mahout_b200.itemsimilarity import ItemSimilarityJob
    rng = np.random.Generator(np.random.PCG64(18))
    n = 20000
    lines = [f"{u}\t{i * 3}\t{p}" for u, i, p in zip(rng.integers(1, 400, n), rng.integers(1, 250, n),
                                                     rng.integers(1, 11, n) * 0.5)]
    inp = tmp_path / "prefs.tsv"
    inp.write_text("\n".join(lines) + "\n")
    flags = ["-s", measure, "-m", "7", "-mp", "2", "--sketchWidth", "512", "--sketchDepth", "3", "--threshold", "0.05"]
    py_out, cc_out = tmp_path / "py.tsv", tmp_path / "cc.tsv"
    assert ItemSimilarityJob().run(["-i", str(inp), "-o", str(py_out), *flags]) == 0
    r = subprocess.run([build.CLI_BIN, "-i", str(inp), "-o", str(cc_out), *flags], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert py_out.read_text() == cc_out.read_text()
    assert len(py_out.read_text().splitlines()) > 100


@pytest.mark.parametrize("measure", ["SIMILARITY_COSINE", "SIMILARITY_SKETCH_COSINE"])
def test_native_cli_num_gpus_writes_the_same_file(tmp_path, measure):
    """--numGpus N: phase 1 as ONE mb200_job_item_similarity call over N GPUs of the process (csrc/job.cu).  With the
    re-scored precision the similarities are bit-equal to DoubleCountMinSketch.cosine however the items are sharded,
    so the output file must be byte-identical to the single-GPU run -- for 1, 2 and all visible GPUs."""
    import subprocess
    import torch
    from mahout_b200 import build
    rng = np.random.Generator(np.random.PCG64(19))
    n = 30000
    lines = [f"{u},{i * 5},{p}" for u, i, p in zip(rng.integers(1, 500, n), rng.integers(1, 300, n),
                                                   rng.integers(1, 11, n) * 0.5)]
    inp = tmp_path / "prefs.csv"
    inp.write_text("\n".join(lines) + "\n")
    flags = ["-s", measure, "-m", "9", "-mp", "2", "--sketchWidth", "1024", "--sketchDepth", "4"]
    base = tmp_path / "single.tsv"
    r = subprocess.run([build.CLI_BIN, "-i", str(inp), "-o", str(base), *flags], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert len(base.read_text().splitlines()) > 100
    for g in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        out = tmp_path / f"gpus{g}.tsv"
        r = subprocess.run([build.CLI_BIN, "-i", str(inp), "-o", str(out), *flags, "--numGpus", str(g)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert f"gpus={g}" in r.stderr and "USED_OBSERVATIONS=" in r.stderr
        assert out.read_text() == base.read_text(), g
