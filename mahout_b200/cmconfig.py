"""CountMinSketchConfig -- the fork's per-user (delta, epsilon) choice.

Mirrors cf/taste/impl/common/CountMinSketchConfig.java:120-219: for a user with n preferences out of
u items, maximise Fmeasure(w, d, n, u, q) over d in [1, 25), w in [d, n]; the LAST maximum in
(d-major, w ascending) order wins (`x >= bestMax`, :139-145); epsilon = e / w, delta = exp(-d)
(:153-154).  The result depends on the user only through n, so one grid per distinct n is evaluated
(vectorised; host-side scalar set-up like the reference's -- not a GPU path).  `Math.pow` is allowed
1 ulp of error by the Java SE spec, so argmax ties at the last ulp are not pinned by the reference.
The `.ser` cache of the reference (Java object serialisation, :74-110) is replaced by an `.npz`.
"""
from __future__ import annotations

import math
import os

import numpy as np

MAX_DEPTH, MIN_DEPTH = 25, 1      # CountMinSketchConfig.java:28-29


def probaNotExactRetrieve(w, d, n):
    """(1 - (1 - 1/W)^N)^D   (CountMinSketchConfig.java:186-192)"""
    w, d, n = np.asarray(w, np.float64), np.asarray(d, np.float64), np.asarray(n, np.float64)
    return np.power(1 - np.power(1 - 1 / w, n), d)


def probaInserted(w, d, n, u):
    """N / (N + falseP (U - N))   (CountMinSketchConfig.java:167-175)"""
    n_, u_ = np.asarray(n, np.float64), np.asarray(u, np.float64)
    return n_ / (n_ + probaNotExactRetrieve(w, d, n) * (u_ - n_))


def Fmeasure(w, d, n, u, q):
    """(1 + 2) beta p / (q^2 beta + p), 0 when beta or p is 0   (CountMinSketchConfig.java:206-215)"""
    beta = 1 - probaNotExactRetrieve(w, d, n)
    p = 1 - probaInserted(w, d, n, u)
    q2 = math.pow(q, 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        f = (1 + 2) * beta * p / (q2 * beta + p)
    return np.where((beta == 0) | (p == 0), 0.0, f)


class CountMinSketchConfig:
    def __init__(self, q: float):
        self.q = float(q)
        self._delta = None
        self._epsilon = None

    def best_dims(self, n: int, u: int):
        """(width, depth) chosen for a user with n preferences; raises like the reference when the grid is empty"""
        n = int(n)
        d = np.arange(MIN_DEPTH, MAX_DEPTH)[:, None]
        w = np.arange(1, max(n, 0) + 1)[None, :]
        ok = w >= d
        if n <= 0 or not ok.any():
            raise RuntimeError("No solution found (this should not happen) (w=0 and d=0")      # TasteException :149
        f = np.where(ok, Fmeasure(np.maximum(w, 1), d, n, u, self.q), -np.inf)
        best = f.max()
        if not best >= 0:
            raise RuntimeError("No solution found (this should not happen) (w=0 and d=0")
        last = np.flatnonzero(f.ravel() == best)[-1]            # `>=`: the last maximum in loop order wins
        di, wi = divmod(int(last), w.shape[1])
        return int(w[0, wi]), int(d[di, 0])

    def configure(self, user_ids, num_prefs, num_items: int, datasetName: str | None = None):
        """computeConfig for every user (user_ids[i] has num_prefs[i] preferences, u = num_items)."""
        user_ids = np.asarray(user_ids, np.int64)
        num_prefs = np.asarray(num_prefs, np.int64)
        path = f"ser/{datasetName}_q_{self.q}.npz" if datasetName else None
        if path and os.path.exists(path):
            z = np.load(path)
            self._delta = dict(zip(z["user"].tolist(), z["delta"].tolist()))
            self._epsilon = dict(zip(z["user"].tolist(), z["epsilon"].tolist()))
            return
        dims = {int(n): self.best_dims(int(n), num_items) for n in np.unique(num_prefs)}
        self._delta = {int(uid): math.exp(-float(dims[int(n)][1])) for uid, n in zip(user_ids, num_prefs)}
        self._epsilon = {int(uid): math.exp(1) / float(dims[int(n)][0]) for uid, n in zip(user_ids, num_prefs)}
        if path:
            os.makedirs("ser", exist_ok=True)
            np.savez(path, user=user_ids, delta=np.array([self._delta[int(x)] for x in user_ids]),
                     epsilon=np.array([self._epsilon[int(x)] for x in user_ids]))

    def getDelta(self, userID: int) -> float:
        if self._delta is None:
            raise RuntimeError("delta is null, call configure method first")
        return self._delta[int(userID)]

    def getEpsilon(self, userID: int) -> float:
        if self._epsilon is None:
            raise RuntimeError("epsilon is null, call configure method first")
        return self._epsilon[int(userID)]
