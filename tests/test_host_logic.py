"""Host-side logic of the measurement drivers that can be exercised without a GPU."""


def test_big_stage_ladder_climbs_down_and_reports():
    """bench_big.climb_down: the forms of a big stage are tried in order, what was abandoned is reported with its error,
    the last form's error is raised"""
    import bench_big as bb
    ladder = [("fused", "certified"), ("pipelined", "certified"), ("pipelined", "tensor")]
    calls = []

    def attempt(m):
        calls.append(m)
        if m[0] == "fused":
            return None, RuntimeError("a peer block never arrived")
        return ("result of", m), None

    seen = []
    done, result, abandoned = bb.climb_down(ladder, attempt, lambda m, e: seen.append((m, str(e))))
    assert done == ("pipelined", "certified") and result == ("result of", done)
    assert calls == ladder[:2] and seen == [(("fused", "certified"), "a peer block never arrived")]
    assert abandoned == [{"form": "fused", "precision": "certified", "error": "RuntimeError('a peer block never arrived')"}]
    done, result, abandoned = bb.climb_down(ladder[:1], lambda m: (1, None))
    assert done == ladder[0] and result == 1 and abandoned == []
    import pytest
    with pytest.raises(KeyError):
        bb.climb_down(ladder, lambda m: ((None, None, None), KeyError(m[1])))
