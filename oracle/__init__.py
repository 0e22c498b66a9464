"""ctypes binding of the CPU oracle (oracle/mahout_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from mahout_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmahout_oracle.so")

PRIME = 9223372036854775783  # HashFunctionBuilder.java:24


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mahout_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, i32, f64 = C.c_int64, C.c_int32, C.c_double
        P = C.c_void_p
        L.orc_java_random_next_int.restype = i32
        L.orc_java_random_next_int.argtypes = [i64]
        L.orc_java_random_next_long.restype = i64
        L.orc_java_random_next_long.argtypes = [i64]
        L.orc_hash_params.argtypes = [i64, C.c_int, P, P]
        L.orc_hash.restype = i32
        L.orc_hash.argtypes = [i64, i64, i32, i64]
        L.orc_hash_many.argtypes = [i64, i64, i32, P, i64, P]
        L.orc_cm_dims.restype = C.c_int
        L.orc_cm_dims.argtypes = [f64, f64, P, P]
        L.orc_cm_update.argtypes = [P, i32, i32, P, P, P, P, i64]
        L.orc_cm_get.restype = f64
        L.orc_cm_get.argtypes = [P, i32, i32, P, P, i64]
        L.orc_cm_cosine.restype = f64
        L.orc_cm_cosine.argtypes = [P, P, i32, i32]
        L.orc_clamp_similarity.restype = f64
        L.orc_clamp_similarity.argtypes = [f64]
        L.orc_bank_update.argtypes = [P, i64, i32, i32, P, P, P, P, P, i64]
        L.orc_bank_update_mt.restype = C.c_int
        L.orc_bank_update_mt.argtypes = [P, i64, i32, i32, P, P, P, P, P, i64, C.c_int]
        L.orc_bank_query.argtypes = [P, i32, i32, P, P, P, P, i64, P]
        L.orc_bank_cosine_topk.argtypes = [P, i64, i32, i32, i64, i64, C.c_int, f64, C.c_int,
                                           C.c_int, P, P, P]
        L.orc_bank_cosine_dense.argtypes = [P, i64, i32, i32, i64, i64, C.c_int, P]
        L.orc_rowsim_cosine_topk.argtypes = [i64, i64, P, P, P, C.c_int, f64, C.c_int, C.c_int,
                                             P, P, P]
        L.orc_exact_cosine.restype = f64
        L.orc_exact_cosine.argtypes = [P, P, i64]
        L.orc_id_to_index.restype = i32
        L.orc_id_to_index.argtypes = [i64]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


NO_THRESHOLD = 4.9e-324  # Double.MIN_VALUE, RowSimilarityJob.java:56


def max_threads() -> int:
    return int(lib().orc_max_threads())


def hash_params(seed: int, depth: int):
    a = np.zeros(depth, np.int64)
    b = np.zeros(depth, np.int64)
    lib().orc_hash_params(seed, depth, _p(a), _p(b))
    return a, b


def hash_one(a: int, b: int, w: int, key: int) -> int:
    return int(lib().orc_hash(int(a), int(b), int(w), int(key)))


def hash_many(a: int, b: int, w: int, keys) -> np.ndarray:
    keys = _c(keys, np.int64)
    out = np.zeros(keys.shape[0], np.int32)
    lib().orc_hash_many(int(a), int(b), int(w), _p(keys), keys.shape[0], _p(out))
    return out


def cm_dims(delta: float, epsilon: float):
    w = C.c_int32()
    d = C.c_int32()
    rc = lib().orc_cm_dims(delta, epsilon, C.byref(w), C.byref(d))
    if rc == 1:
        raise ValueError("CountMinSketch: delta must be between 0 and 1, exclusive")
    if rc == 2:
        raise ValueError("CountMinSketch: epsilon must be between 0 and 1, exclusive")
    return w.value, d.value


def cm_update(count, w, d, a, b, keys, incs):
    keys = _c(keys, np.int64)
    incs = _c(incs, np.float64)
    assert count.dtype == np.float64 and count.flags.c_contiguous
    lib().orc_cm_update(_p(count), w, d, _p(a), _p(b), _p(keys), _p(incs), keys.shape[0])


def cm_get(count, w, d, a, b, key) -> float:
    return float(lib().orc_cm_get(_p(count), w, d, _p(a), _p(b), int(key)))


def cm_cosine(ca, cb, w, d) -> float:
    ca = _c(ca, np.float64)
    cb = _c(cb, np.float64)
    return float(lib().orc_cm_cosine(_p(ca), _p(cb), w, d))


def clamp_similarity(r: float) -> float:
    return float(lib().orc_clamp_similarity(r))


def bank_update(bank, d, w, a, b, entity, key, inc, nthreads: int = 1):
    """bank: float64 [E, d, w] updated in place."""
    E = bank.shape[0]
    key = _c(key, np.int64)
    inc = _c(inc, np.float32)
    ent = _c(entity, np.int64) if entity is not None else None
    assert bank.dtype == np.float64 and bank.flags.c_contiguous
    if nthreads > 1:
        rc = lib().orc_bank_update_mt(_p(bank), E, d, w, _p(a), _p(b), _p(ent), _p(key), _p(inc),
                                      key.shape[0], nthreads)
        if rc < 0:
            raise MemoryError("oracle: private banks")
    else:
        lib().orc_bank_update(_p(bank), E, d, w, _p(a), _p(b), _p(ent), _p(key), _p(inc),
                              key.shape[0])


def bank_query(bank, d, w, a, b, entity, key):
    key = _c(key, np.int64)
    ent = _c(entity, np.int64) if entity is not None else None
    out = np.zeros(key.shape[0], np.float64)
    lib().orc_bank_query(_p(bank), d, w, _p(a), _p(b), _p(ent), _p(key), key.shape[0], _p(out))
    return out


def bank_cosine_topk(bank, k, threshold=NO_THRESHOLD, exclude_self=True, r0=0, r1=None,
                     nthreads=None):
    E, d, w = bank.shape
    r1 = E if r1 is None else r1
    nthreads = max_threads() if nthreads is None else nthreads
    idx = np.zeros((r1 - r0, k), np.int64)
    sim = np.zeros((r1 - r0, k), np.float64)
    cnt = np.zeros(r1 - r0, np.int32)
    lib().orc_bank_cosine_topk(_p(bank), E, d, w, r0, r1, k, threshold, int(exclude_self),
                               nthreads, _p(idx), _p(sim), _p(cnt))
    return idx, sim, cnt


def bank_cosine_dense(bank, r0=0, r1=None, nthreads=None):
    E, d, w = bank.shape
    r1 = E if r1 is None else r1
    nthreads = max_threads() if nthreads is None else nthreads
    out = np.zeros((r1 - r0, E), np.float64)
    lib().orc_bank_cosine_dense(_p(bank), E, d, w, r0, r1, nthreads, _p(out))
    return out


def rowsim_cosine_topk(nrows, ncols, rowptr, colidx, vals, k, threshold=NO_THRESHOLD,
                       exclude_self=True, nthreads=None):
    rowptr = _c(rowptr, np.int64)
    colidx = _c(colidx, np.int32)
    vals = _c(vals, np.float32)
    nthreads = max_threads() if nthreads is None else nthreads
    idx = np.zeros((nrows, k), np.int64)
    sim = np.zeros((nrows, k), np.float64)
    cnt = np.zeros(nrows, np.int32)
    lib().orc_rowsim_cosine_topk(nrows, ncols, _p(rowptr), _p(colidx), _p(vals), k, threshold,
                                 int(exclude_self), nthreads, _p(idx), _p(sim), _p(cnt))
    return idx, sim, cnt


def exact_cosine(x, y) -> float:
    x = _c(x, np.float64)
    y = _c(y, np.float64)
    return float(lib().orc_exact_cosine(_p(x), _p(y), x.shape[0]))


def id_to_index(v: int) -> int:
    return int(lib().orc_id_to_index(int(v)))


def most_similar_item_pairs(idx, sim, cnt, index_to_id=None):
    """ItemSimilarityJob.MostSimilarItemPairsMapper/Reducer
    (ItemSimilarityJob.java:197-232): per-row top-k entries -> (minID, maxID) keys,
    duplicates collapse to one value, output ordered by (a, b)
    (EntityEntityWritable.java:64-71)."""
    pairs = {}
    for r in range(idx.shape[0]):
        rid = int(index_to_id[r]) if index_to_id is not None else r
        for t in range(int(cnt[r])):
            c = int(idx[r, t])
            cid = int(index_to_id[c]) if index_to_id is not None else c
            key = (rid, cid) if rid < cid else (cid, rid)
            pairs.setdefault(key, float(sim[r, t]))
    return sorted((a, b, s) for (a, b), s in pairs.items())
