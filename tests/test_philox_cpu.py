"""Philox4x32-10 oracle: Random123 known-answer vectors + draw-layout sanity (CPU only)."""
import numpy as np

from oracle import philox as ph


def _hex(r):
    return ["%08x" % x for x in r[0]]


def test_philox_known_answers():
    assert _hex(ph.philox4x32_10((0, 0), [[0, 0, 0, 0]])) == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert _hex(ph.philox4x32_10((0xFFFFFFFF, 0xFFFFFFFF), [[0xFFFFFFFF] * 4])) == \
        ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert _hex(ph.philox4x32_10((0xA4093822, 0x299F31D0), [[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]])) == \
        ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_uniform_layout_and_range():
    u = ph.uniform(0x5050_4E45_54, ph.STREAM_UNIFORM, 10, 4, 7)
    assert u.shape == (4, 7) and (u >= 0).all() and (u < 1).all()
    # unit-keyed: rows do not depend on how many units are requested or where the range starts
    assert np.array_equal(ph.uniform(0x5050_4E45_54, ph.STREAM_UNIFORM, 12, 1, 7)[0], u[2])
    big = ph.uniform(1, ph.STREAM_UNIFORM, 0, 64, 256).reshape(-1)
    assert abs(big.mean() - 0.5) < 0.01 and abs(big.var() - 1 / 12) < 0.005


def test_placement_and_candidate_draws():
    a, t0, t1 = ph.placement_draw(7, 123, 0, 224)
    assert -180 <= a < 180 and -112 <= t0 <= 112 and -112 <= t1 <= 112
    c = ph.candidates(7, 123, 50, 50.0, 5.0)
    assert c.shape == (50, 3) and (c[:, :2] < 50).all() and (c[:, 2] < 5).all()
    assert np.array_equal(ph.candidates(7, 123, 50, 50.0, 5.0), c)
    assert not np.array_equal(ph.candidates(7, 124, 50, 50.0, 5.0), c)
