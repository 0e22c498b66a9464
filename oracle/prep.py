"""Host (numpy) restatement of the ingest steps in front of the sketch path -- the checker for
mahout_b200/ingest.py (csrc/ingest.cu).

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: imported from tests/ (and bench.py's reference
arm), never from mahout_b200/.

Reference:
  ToEntityPrefsMapper.map              cf/taste/hadoop/ToEntityPrefsMapper.java:56-76
  TasteHadoopUtils.idToIndex           cf/taste/hadoop/TasteHadoopUtils.java:56-58
  ItemIDIndexMapper / Reducer          cf/taste/hadoop/item/ItemIDIndex*.java (minimum itemID per index)
  ToUserVectorsReducer.reduce          cf/taste/hadoop/item/ToUserVectorsReducer.java:66-82
"""
from __future__ import annotations

import numpy as np

DEFAULT_MIN_PREFS_PER_USER = 1


def id_to_index(ids) -> np.ndarray:
    """TasteHadoopUtils.idToIndex (TasteHadoopUtils.java:56-58):
    0x7FFFFFFF & Longs.hashCode(id) % 0x7FFFFFFE, with Java's precedence and truncating '%'."""
    v = np.asarray(ids, dtype=np.int64).view(np.uint64)
    h = (v ^ (v >> np.uint64(32))).astype(np.uint32).view(np.int32).astype(np.int64)
    m = np.sign(h) * (np.abs(h) % 0x7FFFFFFE)      # Java remainder truncates toward zero
    return (m.astype(np.int32).view(np.uint32) & np.uint32(0x7FFFFFFF)).astype(np.int64)


def parse_prefs(lines, boolean_data: bool = False, rating_shift: float = 0.0):
    """ToEntityPrefsMapper.map: split on [\\t,]; user, item as long; pref as float (1.0 if absent or
    booleanData).  Returns (user int64, item int64, pref float32) in input order."""
    users, items, prefs = [], [], []
    for line in lines:
        line = line.strip()
        if not line:
            continue
        tok = line.replace("\t", ",").split(",")
        users.append(int(tok[0]))
        items.append(int(tok[1]))
        while len(tok) > 2 and tok[-1] == "":          # String.split drops trailing empty strings
            tok.pop()
        if boolean_data or len(tok) < 3:
            prefs.append(np.float32(1.0))
        else:
            prefs.append(np.float32(java_parse_float(tok[2])) + np.float32(rating_shift))   # float + float
    return (np.array(users, np.int64), np.array(items, np.int64), np.array(prefs, np.float32))


_libc = None


def java_parse_float(tok: str) -> np.float32:
    """Float.parseFloat: trims, accepts a trailing f/F/d/D, is correctly rounded (libc strtof is too;
    going through a Python float would round twice)."""
    import ctypes
    global _libc
    if _libc is None:
        _libc = ctypes.CDLL(None)
        _libc.strtof.restype = ctypes.c_float
        _libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    t = tok.strip()
    if t and t[-1] in "fFdD" and not t.lower().startswith(("0x", "+0x", "-0x")):
        t = t[:-1]
    body = t.lstrip("+-")
    if body == "NaN":
        return np.float32("nan")
    if body == "Infinity":
        return np.float32("-inf" if t.startswith("-") else "inf")
    if not t or not (body[:1].isdigit() or body[:1] == "."):
        raise ValueError(f'NumberFormatException: For input string: "{tok}"')
    end = ctypes.c_char_p()
    raw = t.encode()
    v = _libc.strtof(raw, ctypes.byref(end))
    consumed = ctypes.cast(end, ctypes.c_void_p).value - ctypes.cast(ctypes.c_char_p(raw), ctypes.c_void_p).value
    return np.float32(v)


class PreferenceMatrix:
    """Output of the preparation phase: de-duplicated events over dense row numbers + the
    index <-> itemID tables (ItemIDIndexReducer keeps the minimum itemID per index)."""

    def __init__(self, user, item, pref, min_prefs_per_user: int = DEFAULT_MIN_PREFS_PER_USER):
        user = np.asarray(user, np.int64)
        item = np.asarray(item, np.int64)
        pref = np.asarray(pref, np.float32)
        idx = id_to_index(item)
        # userVector.set(index, pref): the last preference of a (user, index) pair wins
        order = np.lexsort((np.arange(user.shape[0]), idx, user))
        u, i, p, it = user[order], idx[order], pref[order], item[order]
        last = np.ones(u.shape[0], bool)
        last[:-1] = (u[1:] != u[:-1]) | (i[1:] != i[:-1])
        # ItemIDIndexReducer: min itemID per index (over ALL input lines, before user filtering)
        uniq_idx, inv = np.unique(idx, return_inverse=True)
        min_id = np.full(uniq_idx.shape[0], np.iinfo(np.int64).max, np.int64)
        np.minimum.at(min_id, inv, item)
        # set(index, 0.0) removes the element (RandomAccessSparseVector.setQuick, math/.../RandomAccessSparseVector.java:125-132):
        # a pair whose last preference is 0.0 does not count toward minPrefsPerUser and is not in the vector
        last &= p != 0.0
        u, i, p = u[last], i[last], p[last]
        # ToUserVectorsReducer: users with fewer than minPrefsPerUser preferences are dropped
        uu, cnt = np.unique(u, return_counts=True)
        keep_users = uu[cnt >= min_prefs_per_user]
        keep = np.isin(u, keep_users)
        self.user, self.index, self.pref = u[keep], i[keep], p[keep]
        self.num_users = int(keep_users.shape[0])
        self.index_values = uniq_idx                 # row r of the matrix <-> index_values[r]
        self.item_id = min_id                        # row r -> itemID written to the output
        self.row = np.searchsorted(uniq_idx, self.index)   # dense row number of every event
        self.num_items = int(uniq_idx.shape[0])
