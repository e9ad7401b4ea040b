"""The feDP restatement (oracle/fedp_oracle.py, R/classify.dd.R:314-383) pinned piece by piece against scipy.stats
and brute force, and the library's host-side Student tail against scipy.special.  PARITY UNPINNED with respect to R
itself (no R in the image, no reference test vectors for feDP)."""
import itertools
import math

import numpy as np
import pytest
from scipy import special, stats

from oracle import fedp_oracle as fo
from scdd_b200 import _lib


def test_welch_matches_scipy():
    rng = np.random.default_rng(0)
    for n in (6, 40, 900):
        y = rng.normal(size=n) * rng.uniform(0.5, 3) + rng.uniform(-1, 1)
        c = np.arange(n) % 2
        y[c == 1] += rng.uniform(-1, 1)
        ref = stats.ttest_ind(y[c == 0], y[c == 1], equal_var=False).pvalue
        assert fo.welch_p(y, c) == pytest.approx(ref, rel=1e-12)


def test_lm_condition_coefficient_matches_normal_equations():
    rng = np.random.default_rng(1)
    for n in (12, 300):
        u = rng.uniform(0.2, 0.9, n)
        c = rng.integers(0, 2, n)
        y = 1.0 + 2.0 * u + 0.3 * c + rng.normal(size=n)
        X = np.column_stack([np.ones(n), u, c.astype(float)])
        beta, res, _, _ = np.linalg.lstsq(X, y, rcond=None)
        cov = float(res[0]) / (n - 3) * np.linalg.inv(X.T @ X)
        ref = 2 * stats.t.sf(abs(beta[2] / math.sqrt(cov[2, 2])), n - 3)
        assert fo.lm_cond_p(y, u, c) == pytest.approx(ref, rel=1e-10)


def _fisher_brute(tab, tol):
    tab = np.asarray(tab)
    rows, c0 = tab.sum(1), int(tab[:, 0].sum())
    N = int(tab.sum())

    def logp(a):
        return (sum(math.lgamma(int(r) + 1) - math.lgamma(x + 1) - math.lgamma(int(r) - x + 1) for r, x in zip(rows, a))
                - (math.lgamma(N + 1) - math.lgamma(c0 + 1) - math.lgamma(N - c0 + 1)))
    lobs = logp(tab[:, 0])
    tot = 0.0
    for a in itertools.product(*[range(int(r) + 1) for r in rows]):
        if sum(a) == c0:
            lp = logp(a)
            if lp <= lobs + tol:
                tot += math.exp(lp)
    return tot


def test_fisher_2x2_matches_scipy_and_rx2_matches_brute_force():
    rng = np.random.default_rng(2)
    for _ in range(20):
        t = rng.integers(0, 25, (2, 2)) + np.array([[1, 0], [0, 1]])
        assert fo.fisher_rx2(t) == pytest.approx(stats.fisher_exact(t)[1], rel=1e-9)
    for r in (3, 4, 5):
        for _ in range(3):
            t = rng.integers(0, 7, (r, 2)) + 1
            assert fo.fisher_rx2(t) == pytest.approx(_fisher_brute(t, fo.FEXACT_TOL), rel=1e-11)
    # all tables are at most as likely as the most likely one: p = 1
    assert fo.fisher_rx2([[5, 5], [5, 5], [5, 5]]) == pytest.approx(1.0, rel=1e-12)


def test_host_student_tail_matches_scipy():
    rng = np.random.default_rng(3)
    t = np.concatenate([rng.normal(size=4000) * 3, rng.normal(size=400) * 30, [0.0, 1e-8, 1e-3, np.nan, 2.0, np.inf]])
    df = np.concatenate([rng.uniform(1.5, 3000, 2200), rng.uniform(3000, 1e5, 2200), [5.0, 5.0, 5.0, 5.0, 1e6, 7.0]])
    out = np.zeros(t.size)
    assert _lib.lib().scdd_selftest_student_p(t.ctypes.data, df.ctypes.data, t.size, out.ctypes.data) == 0
    ref = 2 * special.stdtr(df, -np.abs(t))
    ok = np.isfinite(t) & (df <= 4e5) & (ref > 1e-290)
    assert np.max(np.abs(out[ok] - ref[ok]) / ref[ok]) < 1e-10
    assert out[-6] == 1.0 and math.isnan(out[-3]) and out[-1] == 0.0
    # df > 4e5: R's normal approximation (nmath/pt.c)
    val = 1 / 4e6
    assert out[-2] == pytest.approx(math.erfc(2.0 * (1 - val) / math.sqrt(1 + 4.0 * 2 * val) / math.sqrt(2)), rel=1e-14)


def test_fedp_labels_and_quirks():
    rng = np.random.default_rng(4)
    ng, nc = 12, 60
    cond = np.repeat([1, 2], nc // 2)
    X = np.exp(rng.normal(2.0, 0.5, (ng, nc)))
    X[rng.random(X.shape) < 0.2] = 0
    X[0, cond == 2] *= 4.0                     # strong mean shift, one cluster everywhere -> "NC"
    oa, c1, c2 = [], [], []
    for g in range(ng):
        nz = X[g] > 0
        lab = np.ones(int(nz.sum()), dtype=np.int64)
        if g == 1:                             # two clusters, tied to the condition -> Fisher "DP"
            lab = np.where(cond[nz] == 1, 1, 2)
            lab[:3] = [2, 2, 2]
            lab[-3:] = [1, 1, 1]
        oa.append(lab)
        c1.append(lab[cond[nz] == 1])
        c2.append(lab[cond[nz] == 2])
    pv, cat, ns = fo.feDP(X, cond, [5], oa, c1, c2)
    assert ns.tolist() == [g for g in range(ng) if g != 5] and pv.size == ng - 1
    assert cat[0] == "NC" and cat[1] == "DP" and set(cat[2:]) <= {"NS", "NC"}
    pv0, cat0, ns0 = fo.feDP(X, cond, [], oa, c1, c2)      # x[-integer(0)] is empty in R
    assert ns0.size == 0 and pv0.size == 0
