"""Loop-for-loop restatement of CountMinSketchConfig.computeConfig (checker of mahout_b200/cmconfig.py).
TEST INFRASTRUCTURE ONLY.  Reference: cf/taste/impl/common/CountMinSketchConfig.java:120-219."""
import math


def proba_not_exact(w, d, n):
    return math.pow(1 - math.pow(1 - 1 / float(w), float(n)), float(d))


def proba_inserted(w, d, n, u):
    return float(n) / (float(n) + proba_not_exact(w, d, n) * (float(u) - float(n)))


def fmeasure(w, d, n, u, q):
    beta = 1 - proba_not_exact(w, d, n)
    p = 1 - proba_inserted(w, d, n, u)
    if beta == 0 or p == 0:
        return 0.0
    return (1 + 2) * beta * p / (math.pow(q, 2) * beta + p)


def best_dims(n, u, q):
    best_w = best_d = 0
    best = 0.0
    for d in range(1, 25):
        for w in range(d, n + 1):
            x = fmeasure(w, d, n, u, q)
            if x >= best:
                best_w, best_d, best = w, d, x
    if best_w == 0 and best_d == 0:
        raise RuntimeError("No solution found")
    return best_w, best_d
