"""CosineCM -- the fork's sketch-based UserSimilarity, over the GPU sketch bank.

Mirrors cf/taste/impl/similarity/CosineCM.java: one count-min sketch per user built from the
user's preferences (key = itemID, increment = preference value, CosineCM.java:41-58), similarity
= DoubleCountMinSketch.cosine clamped to [-1, 1] (CosineCM.java:84-96 ->
AbstractSimilarity.normalizeWeightResult, AbstractSimilarity.java:313-330, unweighted), plus the
point-query consumer of GenericUserBasedRecommender.doEstimatePreference (:139-159).

The reference sizes u1's sketch by u2's (delta, epsilon) on every call and rebuilds it each time;
here all users share one (width, depth) -- what BASELINE.json's configurations fix -- and every
profile is built once, in one K1 launch.  Per-user sizing (CountMinSketchConfig) is SURVEY.md 8f-4.
"""
from __future__ import annotations

import math

import numpy as np

from . import sketch as sk


class CosineCM:
    def __init__(self, user, item, pref, width_or_delta=4096, depth_or_epsilon=4,
                 hfBuilder: sk.HashFunctionBuilder | int = 42, frac_bits: int = 1, ctx=None):
        user = np.asarray(user, np.int64)
        if isinstance(width_or_delta, float) or isinstance(depth_or_epsilon, float):
            try:
                w, d = sk.cm_dims(float(width_or_delta), float(depth_or_epsilon))
            except sk.N.CMException as e:        # CosineCM.java:45-47 wraps it in a TasteException
                raise RuntimeError(f"CountMinSketch error:{e}") from e
        else:
            w, d = int(width_or_delta), int(depth_or_epsilon)
        self.user_ids = np.unique(user)
        self._row = np.searchsorted(self.user_ids, user)
        self.bank = sk.SketchBank(self.user_ids.shape[0], w, d, hfBuilder, frac_bits, ctx)
        self.bank.update(self._row, np.asarray(item, np.int64), np.asarray(pref, np.float32))
        self.bank.check()

    def _rows(self, ids):
        ids = np.atleast_1d(np.asarray(ids, np.int64))
        r = np.searchsorted(self.user_ids, ids)
        ok = (r < self.user_ids.shape[0])
        ok[ok] = self.user_ids[r[ok]] == ids[ok]
        if not ok.all():
            raise KeyError(f"NoSuchUserException: {ids[~ok][:5].tolist()}")
        return r.astype(np.int64)

    def userSimilarity(self, userID1: int, userID2: int) -> float:
        return float(self.userSimilarities([userID1], [userID2])[0])

    def userSimilarities(self, ids1, ids2) -> np.ndarray:
        r = self.bank.pair_cosine(self._rows(ids1), self._rows(ids2))
        # normalizeWeightResult(result, 1, 0) for the non-NaN results: clamp to [-1, 1]
        return np.where(np.isnan(r), r, np.clip(r, -1.0, 1.0))

    def getExportedCMProfile(self, userID: int) -> np.ndarray:
        r = int(self._rows(userID)[0])
        return self.bank.read(r, r + 1)[0]

    def estimatePreference(self, userID: int, itemID: int) -> float:
        """GenericUserBasedRecommender.doEstimatePreference's sketch read: (float) cm.get(itemID),
        where 0.0 means "no preference" (NaN)."""
        v = float(np.float32(self.bank.query(self._rows(userID), np.array([itemID], np.int64))[0]))
        return math.nan if v == 0.0 else v

    def mostSimilarUserIDs(self, userID: int, howMany: int):
        """nearest users under the sketch cosine (ties: lower user index first)."""
        idx, sim, cnt = self.bank.cosine_topk(howMany)
        r = int(self._rows(userID)[0])
        return self.user_ids[idx[r, :cnt[r]]], sim[r, :cnt[r]]

    def close(self):
        self.bank.close()
