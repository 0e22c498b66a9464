"""TEST INFRASTRUCTURE -- blocked restatement of the oracle's sketch cosine + top-k for the parity legs of
bench.py at config-4 / config-5 scale, where the loop-for-loop C oracle (oracle/mahout_oracle.c,
orc_bank_cosine_topk) would need 32 GB of FP64 counters per depth row in one address space.

Same arithmetic as DoubleCountMinSketch.cosine (DoubleCountMinSketch.java:114-149), evaluated in blocks:

  * counters are fixed-point quanta (integers).  AA = sum xa^2, BB = sum xb^2, AB = sum xa*xb over integer-valued
    doubles are computed with FP64 matrix products (numpy / BLAS); while every partial sum stays below 2^53 each
    of them is an exactly representable integer whatever the summation order, so the result is the reference's
    sequential sum bit for bit.  The precondition (max|a| * max|b| * W < 2^53) is CHECKED per block; a block that
    fails it raises -- there is no silent approximation.
  * the reference works in preference units x = q * 2^-f, not quanta: AA_ref = AA_q * 4^-f exactly, sqrt(AA_ref) =
    sqrt(AA_q) * 2^-f exactly, and AB_ref / (sqrt(AA_ref) * sqrt(BB_ref)) = AB_q / (sqrt(AA_q) * sqrt(BB_q)) with the
    same roundings (power-of-two scalings commute with IEEE sqrt, * and /).  So the quanta give the same bits.
  * min over depth rows with a non-zero denominator, NaN when none (":138-146"); admission sim >= threshold and
    sim > Double.MIN_VALUE, diagonal excluded (RowSimilarityJob.java:489-499); top-k under (sim desc, index asc).

Pinned against the C oracle bit for bit on small banks by tests/test_oracle.py::test_fast_oracle_*.
Only tests/, bench.py's parity legs and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import numpy as np

NO_THRESHOLD = 4.9e-324  # Double.MIN_VALUE (RowSimilarityJob.java:56)
TWO53 = float(1 << 53)


def _blas_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=int(n))
    except Exception:  # pragma: no cover
        import contextlib
        return contextlib.nullcontext()


def cosine_block(sample_q, cols_q):
    """sample_q [m, d, W], cols_q [n, d, W] integer quanta -> sims [m, n] float64 (NaN = no comparable row)."""
    m, d, W = sample_q.shape
    n = cols_q.shape[0]
    out = np.full((m, n), np.inf)
    for i in range(d):
        a = np.ascontiguousarray(sample_q[:, i, :], dtype=np.float64)
        b = np.ascontiguousarray(cols_q[:, i, :], dtype=np.float64)
        amax = float(np.abs(a).max()) if a.size else 0.0
        bmax = float(np.abs(b).max()) if b.size else 0.0
        if max(amax, bmax) ** 2 * W >= TWO53:
            raise OverflowError("oracle.fast: counters too large for exact FP64 block products")
        aa = np.einsum("ij,ij->i", a, a)
        bb = np.einsum("ij,ij->i", b, b)
        ab = a @ b.T
        den = np.sqrt(aa)[:, None] * np.sqrt(bb)[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            c = ab / den
        ok = den != 0.0
        out = np.where(ok & (c < out), c, out)
    out[np.isinf(out)] = np.nan
    return out


def _topk_merge(sim, idx, k):
    """rows of candidate (sim, idx) pairs -> the k best under (sim desc, idx asc); -inf marks empty slots"""
    order = np.lexsort((idx, -sim), axis=1)[:, :k]
    return np.take_along_axis(sim, order, 1), np.take_along_axis(idx, order, 1)


def rows_vs_columns_topk(sample_q, sample_ids, cols_q, col_ids, k, threshold=NO_THRESHOLD, exclude_self=True,
                         block=8192, threads=1, chunk_loader=None):
    """Top-k of every sample row over the given columns.

    sample_q [m, d, W] quanta, sample_ids [m] global indices; cols_q [n, d, W] quanta (or None with
    chunk_loader(c0, c1) -> quanta [c1-c0, d, W]), col_ids [n] global indices.
    Returns (sim [m, k] with -inf in unused slots, idx [m, k] with -1)."""
    sample_ids = np.asarray(sample_ids, np.int64)
    col_ids = np.asarray(col_ids, np.int64)
    m, n = sample_q.shape[0], col_ids.shape[0]
    big = np.iinfo(np.int64).max
    best_s = np.full((m, k), -np.inf)
    best_i = np.full((m, k), big, np.int64)
    thr = NO_THRESHOLD if threshold is None or threshold <= 0 else float(threshold)
    with _blas_threads(threads):
        for c0 in range(0, n, block):
            c1 = min(n, c0 + block)
            cq = chunk_loader(c0, c1) if chunk_loader is not None else cols_q[c0:c1]
            s = cosine_block(sample_q, cq)
            ids = col_ids[c0:c1]
            keep = ~np.isnan(s) & (s >= thr) & (s > NO_THRESHOLD)
            if exclude_self:
                keep &= ids[None, :] != sample_ids[:, None]
            # only columns that reach the row's current k-th best can enter (>=: ties are settled by index below)
            keep &= s >= best_s[:, k - 1][:, None]
            maxc = int(keep.sum(axis=1).max()) if keep.size else 0
            if maxc == 0:
                continue
            s = np.where(keep, s, -np.inf)
            if maxc < s.shape[1]:
                part = np.argpartition(-s, maxc - 1, axis=1)[:, :maxc]      # every finite entry of a row is in its top maxc
                ps = np.take_along_axis(s, part, 1)
                pi = ids[part]
            else:
                ps, pi = s, np.broadcast_to(ids, s.shape)
            cand_s = np.concatenate([best_s, ps], axis=1)
            cand_i = np.concatenate([best_i, np.where(np.isfinite(ps), pi, big)], axis=1)
            best_s, best_i = _topk_merge(cand_s, cand_i, k)
    return best_s, np.where(np.isfinite(best_s), best_i, -1)


def merge_partials(parts, k):
    """parts: list of (sim [m, k], idx [m, k]) from disjoint column sets -> (idx [m, k], sim [m, k], cnt [m]) in
    the oracle's output convention (unused slots: idx -1, sim 0)."""
    s = np.concatenate([p[0] for p in parts], axis=1)
    i = np.concatenate([p[1] for p in parts], axis=1)
    i = np.where(np.isfinite(s), i, np.iinfo(np.int64).max)
    bs, bi = _topk_merge(s, i, k)
    ok = np.isfinite(bs)
    return np.where(ok, bi, -1), np.where(ok, bs, 0.0), ok.sum(axis=1).astype(np.int32)


def bank_rows_topk(bank_q, rows, k, threshold=NO_THRESHOLD, exclude_self=True, block=8192, threads=1):
    """convenience: top-k of the given rows of one resident bank of quanta [N, d, W] over all N columns"""
    rows = np.asarray(rows, np.int64)
    s, i = rows_vs_columns_topk(bank_q[rows], rows, bank_q, np.arange(bank_q.shape[0]), k, threshold, exclude_self,
                                block, threads)
    return merge_partials([(s, i)], k)
