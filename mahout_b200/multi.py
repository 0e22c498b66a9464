"""One process, all GPUs of the box: the item-similarity phase behind one C call (csrc/job.cu).

Mirror of what a JVM does through JNI for `--numGpus`: `MultiGpu(n)` is `mb200_create_multi`, `item_similarity`
is `mb200_job_item_similarity` = phase 1 of ItemSimilarityJob.run (ItemSimilarityJob.java:146-162).  The
one-process-per-GPU path over torch.distributed (similarity.sharded_item_similarity) computes the same thing for
callers that already run one rank per GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .sketch import _DTYPES, _PRECISIONS


class MultiGpu:
    def __init__(self, n_gpus: int = 0, devices=None):
        h = C.c_void_p()
        dv = None
        if devices is not None:
            dv = np.ascontiguousarray(devices, np.int32)
            n_gpus = int(dv.shape[0])
        N.check(N.lib().mb200_create_multi(int(n_gpus), dv.ctypes.data_as(C.c_void_p) if dv is not None else None,
                                           C.byref(h)))
        self._h = h
        n = C.c_int32()
        N.lib().mb200_multi_gpus(h, C.byref(n))
        self.n_gpus = n.value

    def close(self):
        if getattr(self, "_h", None) is not None:
            try:
                N.lib().mb200_multi_destroy(self._h)
            except Exception:
                pass
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc == N.OK:
            return
        s = N.lib().mb200_multi_last_error(self._h)
        msg = s.decode("utf-8", "replace") if s else ""
        if rc == N.ERR_BAD_ARG:
            raise ValueError(msg)
        if rc == N.ERR_OOM:
            raise MemoryError(msg)
        if rc in (N.ERR_INEXACT, N.ERR_RANGE):
            raise N.InexactError(rc, msg)
        raise N.NativeError(rc, msg)

    def item_similarity(self, row, key, pref, num_items: int, k: int = 100, threshold: float | None = None,
                        width: int = 4096, depth: int = 4, seed: int = 42, frac_bits: int = 1, dtype: str = "f16",
                        precision: str = "rescored", hash_params=None):
        """(idx [N,k] int64, sim [N,k] float64, cnt [N] int32, stats dict) for host arrays of prepared events
        (dense item rows, user keys, preferences).  hash_params = (a, b) replaces the seeded hash family
        (a = 1, b = 0, width = number of users: the exact measure)."""
        row = np.ascontiguousarray(row, np.int64)
        key = np.ascontiguousarray(key, np.int64)
        pref = np.ascontiguousarray(pref, np.float32)
        if not (row.shape == key.shape == pref.shape):
            raise ValueError("row, key and pref must have the same length")
        p = N.JobParams()
        p.k, p.threshold, p.width, p.depth, p.seed = int(k), float(threshold or 0.0), int(width), int(depth), int(seed)
        keep = None
        if hash_params is not None:
            keep = (np.ascontiguousarray(hash_params[0], np.int64), np.ascontiguousarray(hash_params[1], np.int64))
            p.hash_a, p.hash_b = keep[0].ctypes.data, keep[1].ctypes.data
        p.frac_bits, p.dtype, p.precision = int(frac_bits), _DTYPES[dtype], _PRECISIONS[precision]
        idx = np.empty((num_items, k), np.int64)
        sim = np.empty((num_items, k), np.float64)
        cnt = np.empty(num_items, np.int32)
        st = N.JobStats()
        self._check(N.lib().mb200_job_item_similarity(
            self._h, row.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p), pref.ctypes.data_as(C.c_void_p),
            row.shape[0], int(num_items), C.byref(p), idx.ctypes.data_as(C.c_void_p), sim.ctypes.data_as(C.c_void_p),
            cnt.ctypes.data_as(C.c_void_p), C.byref(st)))
        del keep
        return idx, sim, cnt, {f: getattr(st, f) for f, _ in N.JobStats._fields_}
