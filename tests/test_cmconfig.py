"""CountMinSketchConfig mirror (mahout_b200/cmconfig.py) against the loop-for-loop restatement of
CountMinSketchConfig.computeConfig (oracle/cmconfig.py).  Host-side set-up arithmetic: no GPU."""
import math

import numpy as np
import pytest

from mahout_b200.cmconfig import CountMinSketchConfig, Fmeasure, probaInserted, probaNotExactRetrieve
from oracle import cmconfig as oc


def test_formulas_match_the_restatement():
    for w, d, n, u, q in [(5, 2, 10, 100, 1.0), (57, 12, 120, 1682, 0.5), (1, 1, 1, 50, 2.0), (300, 24, 300, 26744, 1.0)]:
        assert float(probaNotExactRetrieve(w, d, n)) == oc.proba_not_exact(w, d, n)
        assert float(probaInserted(w, d, n, u)) == oc.proba_inserted(w, d, n, u)
        assert float(Fmeasure(w, d, n, u, q)) == oc.fmeasure(w, d, n, u, q)


@pytest.mark.parametrize("q", [0.5, 1.0, 2.0])
def test_best_dims_equal_loop_restatement(q):
    cfg = CountMinSketchConfig(q)
    for u in (100, 1682):
        for n in list(range(1, 60)) + [97, 150]:
            assert cfg.best_dims(n, u) == oc.best_dims(n, u, q), (n, u, q)


def test_configure_and_accessors():
    cfg = CountMinSketchConfig(1.0)
    with pytest.raises(RuntimeError, match="call configure method first"):
        cfg.getDelta(1)
    users = np.array([11, 22, 33])
    nprefs = np.array([5, 37, 5])
    cfg.configure(users, nprefs, 1682)
    w5, d5 = oc.best_dims(5, 1682, 1.0)
    w37, d37 = oc.best_dims(37, 1682, 1.0)
    assert cfg.getEpsilon(11) == math.exp(1) / w5 and cfg.getDelta(33) == math.exp(-float(d5))
    assert cfg.getEpsilon(22) == math.exp(1) / w37 and cfg.getDelta(22) == math.exp(-float(d37))
    with pytest.raises(RuntimeError, match="No solution found"):
        CountMinSketchConfig(1.0).best_dims(0, 10)
