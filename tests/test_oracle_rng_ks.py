"""Oracle self-checks (CPU): R's RNG known answers and the two-sample KS test vs scipy."""
import numpy as np
import pytest
from scipy import stats

import np_restatement as npr
from oracle import pyoracle as po


def test_mersenne_twister_known_answers():
    # R: set.seed(1); runif(3) -> 0.2655087 0.3721239 0.5728534  (R documentation examples / folklore values)
    r = po.RRng(seed=1, kind="mt")
    u = [r.unif() for _ in range(3)]
    np.testing.assert_allclose(u, [0.2655086631421, 0.3721238996368, 0.5728533633519], rtol=0, atol=5e-8)
    # R: set.seed(42); runif(2) -> 0.914806 0.9370754
    r = po.RRng(seed=42, kind="mt")
    u = [r.unif() for _ in range(2)]
    np.testing.assert_allclose(u, [0.914806043496, 0.937075413530], rtol=0, atol=5e-7)


def test_sample_known_answers():
    # R >= 3.6: set.seed(1); sample(1:10) -> 9 4 7 1 2 5 3 10 6 8
    r = po.RRng(seed=1, kind="mt")
    assert list(r.sample(10)) == [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]
    # R >= 3.6: set.seed(42); sample(1:10) -> 1 5 10 8 2 4 6 9 7 3
    r = po.RRng(seed=42, kind="mt")
    assert list(r.sample(10)) == [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]


def test_sample_is_permutation():
    r = po.RRng(seed=123, kind="lecuyer")
    for n in (1, 2, 17, 200, 1999):
        s = r.sample(n)
        assert sorted(s) == list(range(1, n + 1))


def test_lecuyer_stream_jumps_are_consistent():
    r = po.RRng(seed=1, kind="lecuyer")
    s0 = r.lecuyer_state()
    a = po.lecuyer_next_stream(s0)
    b = po.lecuyer_next_substream(s0)
    assert not np.array_equal(a, b) and not np.array_equal(a, s0)
    # 2^127 = (2^76)^(2^51): cannot iterate; check the group law on small compositions instead:
    # substream(substream(s)) differs and stays in range
    c = po.lecuyer_next_substream(b)
    assert (c[:3] < 4294967087).all() and (c[3:] < 4294944443).all()


@pytest.mark.parametrize("seed", range(6))
def test_ks_matches_scipy_asymptotic(seed):
    rng = np.random.default_rng(seed)
    x = rng.gamma(2.0, 2.0, size=int(rng.integers(5, 300)))
    y = rng.gamma(2.2, 2.0, size=int(rng.integers(5, 300)))
    if seed % 2:  # ties
        x = np.round(x, 1)
        y = np.round(y, 1)
    D, p = po.ks_test(x, y)
    ref = stats.ks_2samp(x, y, method="asymp")
    assert abs(D - ref.statistic) < 1e-12
    # scipy uses kstwo/kolmogorov sf on sqrt(en)*D; R's series stops at 1e-6 increments
    assert abs(p - stats.distributions.kstwobign.sf(np.sqrt(x.size * y.size / (x.size + y.size)) * D)) < 2e-6
    assert abs(D - npr.ks_D(x, y)) == 0.0
