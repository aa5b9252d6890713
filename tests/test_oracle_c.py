"""The plain-C oracle (oracle/merge_oracle.c) against the NumPy oracle, plus the arithmetic
identities the CUDA kernel relies on."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import merge_oracle as mo


@pytest.mark.parametrize("pvp", [True, False])
@pytest.mark.parametrize("auto_reset", [True, False])
def test_c_equals_numpy(pvp, auto_reset):
    N, T = 1000, 500
    rng = np.random.default_rng(5)
    v = mo.RefVecEnv(N, pvp=pvp, auto_reset=auto_reset)
    c = co.CVecEnv(N, pvp=pvp, auto_reset=auto_reset, nthreads=3)
    for t in range(T):
        a = rng.integers(0, 5, (N, 2)).astype(np.uint8)
        o, r, d, i = v.step(a[:, 0], a[:, 1] if pvp else None)
        o2, r2, d2, i2 = c.step(a[:, 0], a[:, 1])
        assert np.array_equal(d, d2) and np.array_equal(i, i2)
        assert np.array_equal(r, r2) and np.array_equal(o, o2)
    for k in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "steps", "winner"):
        assert np.array_equal(getattr(v, k), getattr(c, k)), k
    sv, sc = v.stats, c.stats
    assert sv["episodes"] > 0
    for k in sv:
        assert sv[k] == pytest.approx(sc[k], rel=1e-12), k


def test_bad_actions_are_flagged_and_clamped():
    c = co.CVecEnv(4); v = mo.RefVecEnv(4)
    a1 = np.array([0, 7, 4, 255], np.uint8); a2 = np.array([9, 1, 2, 3], np.uint8)
    _, _, _, i = c.step(a1, a2)
    _, _, _, i2 = v.step(a1, a2)
    assert np.array_equal(i, i2)
    assert list((i & mo.INFO_BAD_ACTION) != 0) == [True, True, False, True]
    assert np.array_equal(c.vel1, v.vel1)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(c, k):
        return [int(x) for x in mo.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))]
    assert run([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_actions_c_equals_numpy():
    for seed, base, step in [(0x5EED, 0, 0), (2**40 + 17, 2**33 + 5, 2**35 + 1), (1, 123456, 99)]:
        a1, a2 = mo.philox_actions(5000, seed, base, step)
        b1, b2 = co.philox_actions(5000, seed, base, step)
        assert np.array_equal(a1, b1) and np.array_equal(a2, b2)
        assert a1.max() == 4 and a1.min() == 0
    a1, _ = mo.philox_actions(200000, 7, 0, 3)
    assert np.abs(np.bincount(a1, minlength=5) / 200000 - 0.2).max() < 0.01


def test_division_by_constant_identity():
    """q' = fma(fma(-d,q,x), 1/d, q), q = x*(1/d)  ==  x/d bit for bit (kernel's div_const)."""
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.uniform(-40, 40, 2_000_000), np.arange(-4000, 4001) / 100.0,
                        np.nextafter(np.arange(0, 41.0), 100), np.nextafter(np.arange(0, 41.0), -100)])
    assert co.check_div(v, 3.0) == 0
    p = np.concatenate([rng.uniform(0, 25000, 2_000_000), 50 + 0.2 * np.arange(0, 125000),
                        rng.uniform(0, 1e7, 200_000), np.nextafter(30000.0 * np.arange(0, 100), 1e9)])
    assert co.check_div(p, 30000.0) == 0
    # states actually reached by random play
    env = mo.RefVecEnv(2000)
    for t in range(300):
        a = rng.integers(0, 5, (2000, 2))
        env.step(a[:, 0], a[:, 1])
        assert co.check_div(np.concatenate([env.pos1, env.pos2]), 30000.0) == 0
        assert co.check_div(np.concatenate([40 - env.vel1, env.vel2 - 10]), 3.0) == 0


def test_angle_constant():
    assert co.lib().mgo_atan2_h_r() == float(np.arctan2(1000, 30000)) == float.fromhex("0x1.10f7317226afdp-5")


@pytest.mark.parametrize("pvp", [True, False])
def test_c_equals_numpy_from_injected_states(pvp):
    """Same comparison from states no trajectory reaches (anywhere on [-50, 1300), speeds in [0, 60), close pairs,
    photo finishes, a winner already past the line), 25 steps with auto-reset: bit for bit."""
    n = 20000
    rng = np.random.default_rng(123 + pvp)
    p1 = rng.uniform(-50.0, 1300.0, n); p2 = rng.uniform(-50.0, 1300.0, n)
    k = n // 2
    p1[:k] = rng.uniform(900.0, 1010.0, k); p2[:k] = p1[:k] + rng.normal(0.0, 6.0, k)
    p1[k:k + 2000] = 950.0 - rng.uniform(0.0, 9.0, 2000); p2[k:k + 2000] = 950.0 - rng.uniform(0.0, 9.0, 2000)
    v1 = rng.uniform(0.0, 60.0, n); v2 = rng.uniform(0.0, 60.0, n)
    both = (p1 > 950.0) & (p2 >= 950.0)
    p2[both] = 900.0
    w0 = np.where(p1 > 950.0, 1, np.where(p2 >= 950.0, 2, 0)).astype(np.uint8)
    v = mo.RefVecEnv(n, pvp=pvp, auto_reset=True); c = co.CVecEnv(n, pvp=pvp, auto_reset=True, nthreads=4)
    for env in (v, c):
        env.pos1[:] = p1; env.vel1[:] = v1; env.pos2[:] = p2; env.vel2[:] = v2; env.winner[:] = w0
    for t in range(25):
        a = rng.integers(0, 5, (n, 2)).astype(np.uint8)
        o, r, d, i = v.step(a[:, 0], a[:, 1] if pvp else None)
        o2, r2, d2, i2 = c.step(a[:, 0], a[:, 1])
        assert np.array_equal(d, d2) and np.array_equal(i, i2), t
        assert np.array_equal(r, r2) and np.array_equal(o, o2), t
    for k_ in ("pos1", "vel1", "pos2", "vel2", "ret1", "ret2", "steps", "winner"):
        assert np.array_equal(getattr(v, k_), getattr(c, k_)), k_
    assert v.stats["collisions"] > 1000 and v.stats["episodes"] > 5000
