"""Wire format of phase 1's output: the similarity matrix as a Hadoop
`SequenceFile<IntWritable, VectorWritable>`, one row per item -- what `RowSimilarityJob`'s last step
writes (`RowSimilarityJob.java:208-212`, `MergeToTopKSimilaritiesReducer` :542-559) and what
`ItemSimilarityJob`'s phase 2 / `RecommenderJob` read (`ItemSimilarityJob.java:164-171`,
`RecommenderJob.java:189-201`).  Host-side byte formatting of N x k results; not a GPU path.

  VectorWritable      hdfs/src/main/java/org/apache/mahout/math/VectorWritable.java:30-34,86-200
                      flags byte (DENSE 1, SEQUENTIAL 2, NAMED 4, LAX_PRECISION 8), Varint size, then for a
                      sparse vector Varint nnz and per element Varint index (delta-coded when sequential)
                      + big-endian double (float when lax).  `Vectors.topKElements` builds a
                      RandomAccessSparseVector: flags 0, plain indices.
  Varint              hdfs/src/main/java/org/apache/mahout/math/Varint.java:87-93 (7 bits per byte, LSB first)
  SequenceFile        Hadoop 2.4.1 (pom.xml:128), version 6, uncompressed records: "SEQ\\x06", the two
                      class names as Text strings, two false flags, empty metadata, a 16-byte sync
                      marker; records `int recordLen, int keyLen, key, value`, with the escape
                      `int -1 + sync` before a record once 2000 bytes have passed since the last sync.
The container layout is restated from the published Hadoop format, which is not in /root/reference:
it is pinned here only by its own reader and by hand-assembled bytes (tests/test_seqfile.py).
"""
from __future__ import annotations

import hashlib
import struct

import numpy as np

FLAG_DENSE, FLAG_SEQUENTIAL, FLAG_NAMED, FLAG_LAX_PRECISION = 1, 2, 4, 8
KEY_CLASS = "org.apache.hadoop.io.IntWritable"
VALUE_CLASS = "org.apache.mahout.math.VectorWritable"
SYNC_INTERVAL = 2000          # 100 * (4 + 16)


def write_unsigned_varint(value: int) -> bytes:
    """Varint.writeUnsignedVarInt / writeUnsignedVarLong"""
    if value < 0:
        raise ValueError("unsigned varint of a negative value")
    out = bytearray()
    while value & ~0x7F:
        out.append((value & 0x7F) | 0x80)
        value >>= 7
    out.append(value & 0x7F)
    return bytes(out)


def read_unsigned_varint(buf, pos: int):
    value = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7
        if shift > 63:
            raise ValueError("Variable length quantity is too long")


def write_signed_varint(value: int, bits: int = 64) -> bytes:
    """Varint.writeSignedVarLong / writeSignedVarInt: zig-zag, then unsigned"""
    return write_unsigned_varint(((value << 1) ^ (value >> (bits - 1))) & ((1 << bits) - 1))


def vector_writable(size: int, indices, values, sequential: bool = False, lax: bool = False) -> bytes:
    """VectorWritable.writeVector of a sparse vector with the given non-zero elements (zeros are skipped,
    :177-180, but still counted by getNumNonZeroElements only if stored -- callers pass no zeros)."""
    flags = (FLAG_SEQUENTIAL if sequential else 0) | (FLAG_LAX_PRECISION if lax else 0)
    idx = np.asarray(indices, np.int64)
    val = np.asarray(values, np.float64)
    if sequential:
        order = np.argsort(idx, kind="stable")
        idx, val = idx[order], val[order]
    out = bytearray([flags]) + write_unsigned_varint(int(size)) + write_unsigned_varint(int(idx.shape[0]))
    last = 0
    for i, v in zip(idx.tolist(), val.tolist()):
        out += write_unsigned_varint(i - last if sequential else i)
        if sequential:
            last = i
        out += struct.pack(">f", v) if lax else struct.pack(">d", v)
    return bytes(out)


def parse_vector_writable(buf, pos: int = 0):
    """VectorWritable.readFields -> (size, indices, values, end position); dense vectors come back with
    indices 0..size-1"""
    flags = buf[pos]
    pos += 1
    if flags >> 4:
        raise ValueError(f"Unknown flags set: {flags:b}")
    size, pos = read_unsigned_varint(buf, pos)
    lax = bool(flags & FLAG_LAX_PRECISION)
    w, fmt = (4, ">f") if lax else (8, ">d")
    idx, val = [], []
    if flags & FLAG_DENSE:
        for i in range(size):
            idx.append(i)
            val.append(struct.unpack_from(fmt, buf, pos)[0])
            pos += w
    else:
        nnz, pos = read_unsigned_varint(buf, pos)
        last = 0
        for _ in range(nnz):
            d, pos = read_unsigned_varint(buf, pos)
            i = last + d if flags & FLAG_SEQUENTIAL else d
            last = i
            idx.append(i)
            val.append(struct.unpack_from(fmt, buf, pos)[0])
            pos += w
    if flags & FLAG_NAMED:
        n = struct.unpack_from(">H", buf, pos)[0]
        pos += 2 + n
    return size, np.array(idx, np.int64), np.array(val, np.float64), pos


def _text_string(s: str) -> bytes:
    b = s.encode("utf-8")
    assert len(b) < 128                      # one-byte Hadoop vint
    return bytes([len(b)]) + b


class SequenceFileWriter:
    """Uncompressed version-6 SequenceFile<IntWritable, VectorWritable>."""

    def __init__(self, path: str, sync: bytes | None = None):
        self.f = open(path, "wb")
        self.sync = sync if sync is not None else hashlib.md5(path.encode()).digest()
        assert len(self.sync) == 16
        self.f.write(b"SEQ\x06" + _text_string(KEY_CLASS) + _text_string(VALUE_CLASS) + b"\x00\x00" +
                     struct.pack(">i", 0) + self.sync)
        self.last_sync = self.f.tell()

    def append(self, key: int, value: bytes):
        if self.f.tell() >= self.last_sync + SYNC_INTERVAL:
            self.f.write(struct.pack(">i", -1) + self.sync)
            self.last_sync = self.f.tell()
        k = struct.pack(">i", int(key))
        self.f.write(struct.pack(">ii", len(k) + len(value), len(k)) + k + value)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_sequence_file(path: str):
    """-> list of (int key, value bytes); checks the header and the sync escapes"""
    buf = open(path, "rb").read()
    if buf[:3] != b"SEQ" or buf[3] != 6:
        raise ValueError("not a version-6 SequenceFile")
    pos = 4
    names = []
    for _ in range(2):
        n = buf[pos]
        names.append(buf[pos + 1:pos + 1 + n].decode())
        pos += 1 + n
    if names != [KEY_CLASS, VALUE_CLASS]:
        raise ValueError(f"unexpected key/value classes {names}")
    if buf[pos:pos + 2] != b"\x00\x00":
        raise ValueError("compressed SequenceFiles are not supported")
    pos += 2
    if struct.unpack_from(">i", buf, pos)[0] != 0:
        raise ValueError("metadata is not supported")
    pos += 4
    sync = buf[pos:pos + 16]
    pos += 16
    out = []
    while pos < len(buf):
        rec_len = struct.unpack_from(">i", buf, pos)[0]
        pos += 4
        if rec_len == -1:
            if buf[pos:pos + 16] != sync:
                raise ValueError("File is corrupt!")
            pos += 16
            continue
        key_len = struct.unpack_from(">i", buf, pos)[0]
        pos += 4
        key = struct.unpack_from(">i", buf, pos)[0]
        out.append((key, buf[pos + key_len:pos + rec_len]))
        pos += rec_len
    return out


INT_MAX = 2147483647


def write_similarity_matrix(path: str, idx, sim, cnt, index_values=None, num_columns: int | None = None):
    """rows of the top-k similarity matrix (idx [N,k], sim [N,k], cnt [N]) as RowSimilarityJob writes them.

    `index_values[r]` is the idToIndex value of dense row r (Prefs.index_values): the reference keys the rows
    of the item-item matrix by TasteHadoopUtils.idToIndex(itemID) (ToItemVectorsMapper.java:41-53 builds the
    item vectors under that index), the vector indices are the same index space, and every vector has
    cardinality Integer.MAX_VALUE (RandomAccessSparseVector(Integer.MAX_VALUE, ...), kept by
    Vectors.topKElements: `new RandomAccessSparseVector(original.size(), k)`, Vectors.java:74).  Phase 2
    (MostSimilarItemPairsMapper) and RecommenderJob resolve these indexes through the itemIDIndex map.
    Without `index_values` the dense row numbers are written with cardinality `num_columns` (a plain
    RowSimilarityJob over an already dense matrix, --numberOfColumns).  Rows without similarities are not
    written (the reducer never sees them)."""
    idx, sim, cnt = np.asarray(idx), np.asarray(sim), np.asarray(cnt)
    if index_values is not None:
        index_values = np.asarray(index_values, np.int64)
        n_cols = INT_MAX
    else:
        n_cols = int(num_columns if num_columns is not None else idx.shape[0])
    with SequenceFileWriter(path) as w:
        for r in range(idx.shape[0]):
            c = int(cnt[r])
            if c:
                if index_values is not None:
                    w.append(int(index_values[r]), vector_writable(n_cols, index_values[idx[r, :c]], sim[r, :c]))
                else:
                    w.append(r, vector_writable(n_cols, idx[r, :c], sim[r, :c]))


def read_similarity_matrix(path: str):
    """-> {row: (indices, values)}"""
    out = {}
    for key, value in read_sequence_file(path):
        _, i, v, _ = parse_vector_writable(value)
        out[key] = (i, v)
    return out
