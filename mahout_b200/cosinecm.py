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
                 hfBuilder: sk.HashFunctionBuilder | int = 42, frac_bits: int = 1, ctx=None, config=None):
        """config: a configured CountMinSketchConfig -> `userSimilarityPerUserConfig` follows the reference's
        per-pair sizing (u1's sketch rebuilt with u2's (delta, epsilon), CosineCM.java:84-96)."""
        user = np.asarray(user, np.int64)
        self.config = config
        self._hfb = hfBuilder if isinstance(hfBuilder, sk.HashFunctionBuilder) else sk.HashFunctionBuilder(int(hfBuilder))
        self._frac_bits, self._ctx = frac_bits, ctx
        order = np.argsort(user, kind="stable")
        self._u_sorted = user[order]
        self._i_sorted = np.asarray(item, np.int64)[order]
        self._p_sorted = np.asarray(pref, np.float64)[order]
        if isinstance(width_or_delta, float) or isinstance(depth_or_epsilon, float):
            try:
                w, d = sk.cm_dims(float(width_or_delta), float(depth_or_epsilon))
            except sk.N.CMException as e:        # CosineCM.java:45-47 wraps it in a TasteException
                raise RuntimeError(f"CountMinSketch error:{e}") from e
        else:
            w, d = int(width_or_delta), int(depth_or_epsilon)
        self.user_ids = np.unique(user)
        self._row = np.searchsorted(self.user_ids, user)
        self.bank = sk.SketchBank(self.user_ids.shape[0], w, d, self._hfb, frac_bits, ctx)
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

    def _prefs_of(self, userID: int):
        lo, hi = np.searchsorted(self._u_sorted, [userID, userID + 1])
        if lo == hi:
            raise KeyError(f"NoSuchUserException: {userID}")
        return self._i_sorted[lo:hi], self._p_sorted[lo:hi]

    def exportProfile(self, userID: int, delta: float, epsilon: float) -> sk.DoubleCountMinSketch:
        """CosineCM.exportProfile (:41-58): a fresh sketch of the user's preferences sized by (delta, epsilon)"""
        try:
            cm = sk.DoubleCountMinSketch(float(delta), float(epsilon), self._hfb, self._frac_bits, self._ctx)
        except sk.N.CMException as e:
            raise RuntimeError(f"CountMinSketch error:{e}") from e
        items, prefs = self._prefs_of(userID)
        cm.update(items, prefs)
        return cm

    def userSimilarityPerUserConfig(self, userID1: int, userID2: int) -> float:
        """CosineCM.userSimilarity exactly as written (:84-96): both sketches sized by user 2's parameters."""
        if self.config is None:
            raise ValueError("no CountMinSketchConfig was given")
        d2, e2 = self.config.getDelta(userID2), self.config.getEpsilon(userID2)
        cm1, cm2 = self.exportProfile(userID1, d2, e2), self.exportProfile(userID2, d2, e2)
        r = sk.DoubleCountMinSketch.cosine(cm1, cm2)
        return r if math.isnan(r) else float(np.clip(r, -1.0, 1.0))

    def doEstimatePreference(self, theUserID: int, theNeighborhood, itemID: int) -> float:
        """GenericUserBasedRecommender.doEstimatePreference with the sketch point query (:134-184): one
        batched mb200_bank_query over the neighbourhood, one batched pair cosine; NaN unless at least two
        neighbours contribute."""
        nb = np.asarray([u for u in np.atleast_1d(theNeighborhood) if u != theUserID], np.int64)
        if np.atleast_1d(theNeighborhood).shape[0] == 0 or nb.shape[0] == 0:
            return math.nan
        rows = self._rows(nb)
        prefs = self.bank.query(rows, np.full(nb.shape[0], itemID, np.int64)).astype(np.float32)   # (float) cm.get
        sims = self.userSimilarities(np.full(nb.shape[0], theUserID, np.int64), nb)
        preference = totalSimilarity = 0.0
        count = 0
        for p, s in zip(prefs.tolist(), sims.tolist()):             # same accumulation order as the Java loop
            if p != 0.0 and not math.isnan(s):
                preference += s * p
                totalSimilarity += s
                count += 1
        if count <= 1:
            return math.nan
        return float(np.float32(preference / totalSimilarity))

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
