"""Synthetic Zipf-distributed (user, item, pref) event streams.

`events_device` fills torch CUDA tensors with the kernel in csrc/synth.cu; `events_numpy`
restates the same arithmetic with numpy so the CPU oracle can regenerate any slice of the
stream bit for bit.  Workload generation only -- not part of the reference's path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def zipf_cdf(items: int, s: float) -> np.ndarray:
    """float64 CDF over item ranks 1..items of the truncated Zipf(s) distribution."""
    w = np.arange(1, items + 1, dtype=np.float64) ** (-float(s))
    c = np.cumsum(w)
    c /= c[-1]
    c[-1] = 1.0
    return c


def rank_permutation(items: int, seed: int) -> np.ndarray:
    """rank -> itemID (1-based) by a fixed random permutation."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.permutation(items) + 1).astype(np.int64)


def _fin(z: np.ndarray) -> np.ndarray:
    z = z ^ (z >> np.uint64(30))
    z = z * _M1
    z = z ^ (z >> np.uint64(27))
    z = z * _M2
    return z ^ (z >> np.uint64(31))


def events_numpy(seed: int, first: int, n: int, users: int, cdf: np.ndarray, perm=None):
    """(user, item, pref) of events [first, first+n) -- same values as events_device."""
    with np.errstate(over="ignore"):
        base = np.uint64(seed & (2 ** 64 - 1)) * _GOLD
        t4 = (np.arange(first, first + n, dtype=np.uint64)) * np.uint64(4)
        r0 = _fin(base + t4)
        r1 = _fin(base + t4 + np.uint64(1))
        r2 = _fin(base + t4 + np.uint64(2))
    u = (r1 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    rank = np.searchsorted(cdf, u, side="left")
    rank = np.minimum(rank, cdf.shape[0] - 1)
    user = (1 + (r0 % np.uint64(users))).astype(np.int64)
    item = perm[rank] if perm is not None else (rank + 1).astype(np.int64)
    pref = (0.5 * (1 + (r2 % np.uint64(10)).astype(np.int64))).astype(np.float32)
    return user, item.astype(np.int64), pref


def events_device(ctx, seed: int, first: int, n: int, users: int, cdf_dev, perm_dev=None,
                  want_user: bool = True, want_pref: bool = True):
    """Same stream, generated into torch CUDA tensors on ctx's device."""
    import torch
    dev = f"cuda:{ctx.device}"
    item = torch.empty(n, dtype=torch.int64, device=dev)
    user = torch.empty(n, dtype=torch.int64, device=dev) if want_user else None
    pref = torch.empty(n, dtype=torch.float32, device=dev) if want_pref else None
    torch.cuda.synchronize(ctx.device)
    N.check(N.lib().mb200_synth_events(
        ctx.handle, C.c_uint64(seed & (2 ** 64 - 1)), first, n, users,
        C.c_void_p(cdf_dev.data_ptr()), cdf_dev.numel(),
        C.c_void_p(perm_dev.data_ptr()) if perm_dev is not None else None,
        C.c_void_p(user.data_ptr()) if user is not None else None,
        C.c_void_p(item.data_ptr()),
        C.c_void_p(pref.data_ptr()) if pref is not None else None), ctx.handle)
    ctx.sync()
    return user, item, pref


def red64_peak(ctx, cells_log2: int = 22, updates: int = 1 << 31) -> float:
    """measured RED.ADD.64 rate (reductions / s) into a 2^cells_log2-word array (default 32 MiB: L2-resident, the
    shape of config 2's sketch) -- the denominator for K1's reductions per second"""
    ms, done = C.c_double(), C.c_double()
    N.check(N.lib().mb200_bench_red64(ctx.handle, cells_log2, updates, C.byref(ms), C.byref(done)), ctx.handle)
    return done.value / (ms.value * 1e-3)
