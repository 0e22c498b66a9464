"""Ingest in front of the sketch path, on the GPU (csrc/ingest.cu over the C ABI).

Mirrors PreparePreferenceMatrixJob (cf/taste/hadoop/preparation/PreparePreferenceMatrixJob.java:54-114):

  Events.parse(text)            ToEntityPrefsMapper.map over the whole file   ToEntityPrefsMapper.java:56-76
  Events.prepare(min_prefs)     idToIndex, index -> minimum itemID, last pref of a (user, index) wins,
                                minPrefsPerUser   TasteHadoopUtils.java:56-58, ItemIDIndexReducer.java:31-46,
                                                  ToUserVectorsReducer.java:66-82

The events stay device-resident from the text buffer to K1 (`PreparedPrefs.row/user/pref` are torch
CUDA views of the library's buffers).  No CPU path: every method is a C-ABI call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from . import sketch as sk


def _view(ptr, n: int, device: int, typestr: str):
    """torch view of n elements of device memory owned by the library (plumbing only)"""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=f"cuda:{device}")


def id_to_index(ids, ctx: sk.Context | None = None) -> np.ndarray:
    """TasteHadoopUtils.idToIndex for an array of ids (GPU)."""
    ctx = ctx or sk.default_context()
    a = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.empty(a.shape, np.int32)
    N.check(N.lib().mb200_id_to_index(ctx.handle, a.ctypes.data_as(C.c_void_p), a.size,
                                      out.ctypes.data_as(C.c_void_p), N.MEM_HOST), ctx.handle)
    return out.astype(np.int64)


class Events:
    """Device-resident (user, item, pref) columns."""

    def __init__(self, handle, ctx: sk.Context):
        self._h, self.ctx = handle, ctx

    @classmethod
    def parse(cls, text, boolean_data: bool = False, rating_shift: float = 0.0, transpose: bool = False,
              ctx: sk.Context | None = None) -> "Events":
        """text: bytes / bytearray / str (host) or a torch.uint8 CUDA tensor (device)."""
        ctx = ctx or sk.default_context()
        h = C.c_void_p()
        if sk._is_torch(text):
            if not text.is_cuda:
                text = bytes(text.numpy().tobytes())
            else:
                text = text.contiguous()
                N.check(N.lib().mb200_events_parse(ctx.handle, C.c_void_p(text.data_ptr()), text.numel(), N.MEM_DEVICE,
                                                   int(boolean_data), float(rating_shift), int(transpose),
                                                   C.byref(h)), ctx.handle)
                return cls(h, ctx)
        if isinstance(text, str):
            text = text.encode("utf-8")
        buf = (C.c_char * len(text)).from_buffer_copy(text) if not isinstance(text, bytearray) else \
            (C.c_char * len(text)).from_buffer(text)
        N.check(N.lib().mb200_events_parse(ctx.handle, C.cast(buf, C.c_void_p), len(text), N.MEM_HOST,
                                           int(boolean_data), float(rating_shift), int(transpose), C.byref(h)),
                ctx.handle)
        return cls(h, ctx)

    @classmethod
    def parse_file(cls, path: str, boolean_data: bool = False, rating_shift: float = 0.0, transpose: bool = False,
                   ctx: sk.Context | None = None) -> "Events":
        """Read the file straight into page-locked memory (mb200_host_alloc) and parse it from there: the
        H2D copy runs at link speed instead of through a pageable staging copy."""
        import os
        ctx = ctx or sk.default_context()
        size = os.path.getsize(path)
        if size == 0:
            return cls.parse(b"", boolean_data, rating_shift, transpose, ctx)
        p = C.c_void_p()
        N.check(N.lib().mb200_host_alloc(size, C.byref(p)), ctx.handle)
        try:
            buf = (C.c_char * size).from_address(p.value)
            with open(path, "rb", buffering=0) as f:
                got = 0
                mv = memoryview(buf).cast("B")
                while got < size:
                    r = f.readinto(mv[got:])
                    if not r:
                        break
                    got += r
            h = C.c_void_p()
            N.check(N.lib().mb200_events_parse(ctx.handle, p, got, N.MEM_HOST, int(boolean_data), float(rating_shift),
                                               int(transpose), C.byref(h)), ctx.handle)
            return cls(h, ctx)
        finally:
            N.lib().mb200_host_free(p)

    @classmethod
    def from_arrays(cls, user, item, pref, ctx: sk.Context | None = None) -> "Events":
        ctx = ctx or sk.default_context()
        u = sk._Arg(user, np.int64, "int64")
        i = sk._Arg(item, np.int64, "int64")
        p = sk._Arg(pref, np.float32, "float32")
        if not (u.n == i.n == p.n):
            raise ValueError("user, item and pref must have the same length")
        mem = sk._same_mem(u, i, p)
        h = C.c_void_p()
        N.check(N.lib().mb200_events_create(ctx.handle, u.ptr, i.ptr, p.ptr, u.n, mem, C.byref(h)), ctx.handle)
        return cls(h, ctx)

    @property
    def handle(self):
        if self._h is None:
            raise ValueError("events are closed")
        return self._h

    def __len__(self) -> int:
        n = C.c_int64()
        N.check(N.lib().mb200_events_count(self.handle, C.byref(n)), self.ctx.handle)
        return n.value

    def columns(self):
        """(user, item, pref) as torch CUDA views, valid until close()"""
        n = len(self)
        u, i, p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(N.lib().mb200_events_columns(self.handle, C.byref(u), C.byref(i), C.byref(p)), self.ctx.handle)
        if n == 0:
            import torch
            dev = f"cuda:{self.ctx.device}"
            return (torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                    torch.empty(0, dtype=torch.float32, device=dev))
        d = self.ctx.device
        return _view(u.value, n, d, "<i8"), _view(i.value, n, d, "<i8"), _view(p.value, n, d, "<f4")

    def read(self):
        """(user, item, pref) as numpy arrays"""
        n = len(self)
        u, i, p = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float32)
        N.check(N.lib().mb200_events_read(self.handle, u.ctypes.data_as(C.c_void_p), i.ctypes.data_as(C.c_void_p),
                                          p.ctypes.data_as(C.c_void_p)), self.ctx.handle)
        return u, i, p

    def prepare(self, min_prefs_per_user: int = 1) -> "PreparedPrefs":
        h = C.c_void_p()
        N.check(N.lib().mb200_events_prepare(self.handle, int(min_prefs_per_user), C.byref(h)), self.ctx.handle)
        return PreparedPrefs(h, self.ctx)

    def close(self):
        if getattr(self, "_h", None) is not None and self.ctx._h is not None:
            try:
                N.lib().mb200_events_destroy(self._h)
            except Exception:
                pass
        self._h = None

    __del__ = close


class PreparedPrefs:
    """Output of the preparation phase: surviving events over dense row numbers (device) + the
    row <-> itemID / index tables (host)."""

    def __init__(self, handle, ctx: sk.Context):
        self._h, self.ctx = handle, ctx
        n, ni, nu = C.c_int64(), C.c_int64(), C.c_int64()
        N.check(N.lib().mb200_prefs_info(handle, C.byref(n), C.byref(ni), C.byref(nu)), ctx.handle)
        self.n, self.num_items, self.num_users = n.value, ni.value, nu.value
        self.item_id = np.empty(self.num_items, np.int64)          # row r -> itemID written to the output
        self.index_values = np.empty(self.num_items, np.int32)     # row r -> idToIndex value (ascending)
        N.check(N.lib().mb200_prefs_tables(handle, self.item_id.ctypes.data_as(C.c_void_p),
                                           self.index_values.ctypes.data_as(C.c_void_p)), ctx.handle)
        r, u, p = C.c_void_p(), C.c_void_p(), C.c_void_p()
        N.check(N.lib().mb200_prefs_columns(handle, C.byref(r), C.byref(u), C.byref(p)), ctx.handle)
        uc = C.c_void_p()
        N.check(N.lib().mb200_prefs_user_columns(handle, C.byref(uc)), ctx.handle)
        if self.n:
            d = ctx.device
            self.row, self.user, self.pref = (_view(r.value, self.n, d, "<i8"), _view(u.value, self.n, d, "<i8"),
                                              _view(p.value, self.n, d, "<f4"))
            self.ucol = _view(uc.value, self.n, d, "<i8")       # dense user number (exact measure's column)
        else:
            import torch
            dev = f"cuda:{ctx.device}"
            self.row = torch.empty(0, dtype=torch.int64, device=dev)
            self.user = torch.empty(0, dtype=torch.int64, device=dev)
            self.pref = torch.empty(0, dtype=torch.float32, device=dev)
            self.ucol = torch.empty(0, dtype=torch.int64, device=dev)

    def close(self):
        if getattr(self, "_h", None) is not None and self.ctx._h is not None:
            self.row = self.user = self.pref = self.ucol = None
            try:
                N.lib().mb200_prefs_destroy(self._h)
            except Exception:
                pass
        self._h = None

    __del__ = close
