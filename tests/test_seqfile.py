"""SequenceFile<IntWritable, VectorWritable> wire format (mahout_b200/seqfile.py): Varint properties pinned by
the reference's VarintTest (hdfs/src/test/java/org/apache/mahout/math/VarintTest.java:160-187), the
VectorWritable layout by hand-assembled bytes following VectorWritable.java:86-200, the container by its
reader.  Host-side formatting: no GPU."""
import struct

import numpy as np
import pytest

from mahout_b200 import seqfile as sf


def test_varint_sizes_match_reference_test():
    # VarintTest.testUnsignedSize: 1 << e takes 1 + e / 7 bytes; testSignedSize: 1 + (e + 1) / 7
    for e in range(63):
        assert len(sf.write_unsigned_varint(1 << e)) == 1 + e // 7
    for e in range(62):
        assert len(sf.write_signed_varint(1 << e)) == 1 + (e + 1) // 7
        assert len(sf.write_signed_varint(-(1 << e) - 1)) == 1 + (e + 1) // 7
    for v in (0, 1, 127, 128, 300, 2 ** 31 - 1, 2 ** 63 - 1):
        assert sf.read_unsigned_varint(sf.write_unsigned_varint(v), 0) == (v, len(sf.write_unsigned_varint(v)))
    assert sf.write_unsigned_varint(300) == bytes([0xAC, 0x02])           # LSB-first 7-bit groups


def test_vector_writable_bytes():
    # RandomAccessSparseVector(size 1682) with elements {5: 0.5, 300: 0.25}: flags 0, size, nnz, (index, double BE)
    b = sf.vector_writable(1682, [5, 300], [0.5, 0.25])
    want = bytes([0x00]) + bytes([0x92, 0x0D]) + bytes([0x02]) + bytes([0x05]) + struct.pack(">d", 0.5) + \
        bytes([0xAC, 0x02]) + struct.pack(">d", 0.25)
    assert b == want
    size, i, v, end = sf.parse_vector_writable(b)
    assert (size, i.tolist(), v.tolist(), end) == (1682, [5, 300], [0.5, 0.25], len(b))
    # sequential + lax precision: delta-coded indices, floats (what ToUserVectorsReducer writes, :78-79)
    b = sf.vector_writable(10, [7, 2], [1.5, 3.0], sequential=True, lax=True)
    assert b == bytes([0x0A, 0x0A, 0x02, 0x02]) + struct.pack(">f", 3.0) + bytes([0x05]) + struct.pack(">f", 1.5)
    assert sf.parse_vector_writable(b)[1].tolist() == [2, 7]
    with pytest.raises(ValueError, match="Unknown flags"):
        sf.parse_vector_writable(bytes([0x10, 0x00]))


def test_sequence_file_round_trip_with_sync_escapes(tmp_path):
    rng = np.random.Generator(np.random.PCG64(4))
    N, k = 400, 9
    idx = np.full((N, k), -1, np.int64)
    sim = np.zeros((N, k))
    cnt = rng.integers(0, k + 1, N).astype(np.int32)
    for r in range(N):
        idx[r, :cnt[r]] = rng.choice(N, cnt[r], replace=False)
        sim[r, :cnt[r]] = np.sort(rng.random(cnt[r]))[::-1]
    path = str(tmp_path / "part-r-00000")
    sf.write_similarity_matrix(path, idx, sim, cnt)
    raw = open(path, "rb").read()
    assert raw.startswith(b"SEQ\x06\x20org.apache.hadoop.io.IntWritable\x25org.apache.mahout.math.VectorWritable\x00\x00")
    assert raw.count(struct.pack(">i", -1) + raw[81:97]) >= 5            # sync escapes every ~2000 bytes
    back = sf.read_similarity_matrix(path)
    assert sorted(back) == [r for r in range(N) if cnt[r]]
    for r, (i, v) in back.items():
        assert i.tolist() == idx[r, :cnt[r]].tolist() and v.tolist() == sim[r, :cnt[r]].tolist()


def test_similarity_matrix_is_keyed_by_id_to_index(tmp_path):
    """What phase 2 / RecommenderJob read: keys and vector indices are TasteHadoopUtils.idToIndex(itemID), the
    cardinality is Integer.MAX_VALUE; decoding through the oracle's idToIndex map gives back the item IDs."""
    import oracle as orc
    item_id = np.array([10, 2 ** 31 + 5, 77, 2 ** 40 + 3, 123456789012], np.int64)     # dense row -> itemID
    index_values = np.array([orc.id_to_index(int(t)) for t in item_id], np.int64)
    order = np.argsort(index_values)                                                   # rows ascend by index
    item_id, index_values = item_id[order], index_values[order]
    idx = np.array([[1, 2], [0, -1], [4, 3], [2, -1], [-1, -1]], np.int64)
    sim = np.array([[0.9, 0.5], [0.9, 0.0], [0.7, 0.6], [0.6, 0.0], [0.0, 0.0]])
    cnt = np.array([2, 1, 2, 1, 0], np.int32)
    path = str(tmp_path / "part-r-00000")
    sf.write_similarity_matrix(path, idx, sim, cnt, index_values)
    index_to_id = {int(i): int(t) for i, t in zip(index_values, item_id)}
    rows = sf.read_sequence_file(path)
    assert [k for k, _ in rows] == [int(index_values[r]) for r in range(4)]
    for r, (key, value) in enumerate(rows):
        size, i, v, _ = sf.parse_vector_writable(value)
        assert size == 2147483647
        assert index_to_id[key] == int(item_id[r])
        assert [index_to_id[int(c)] for c in i] == [int(item_id[c]) for c in idx[r, :cnt[r]]]
        assert v.tolist() == sim[r, :cnt[r]].tolist()
