"""testZeroes (R/classify.dd.R:262-286): the oracle's restatement of arm::bayesglm against the published model, and
the device kernel against the oracle.

arm is not in /root/reference and R is on neither machine, so the oracle (oracle/zeroes_oracle.py) is PARITY UNPINNED
against arm itself; the CPU tests here pin what can be pinned without it: the link functions against closed forms, the
prior-scale rule, and -- run to tight convergence -- the two fixed-point equations of Gelman et al. (2008) section 3,
checked with arithmetic that shares nothing with the IRLS loop.
"""
import os

import numpy as np
import pytest

from oracle import zeroes_oracle as zo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    d = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=True)
    return d["normcounts"], d["condition"]


def _design(X, cond):
    det = (X > 0).sum(axis=0) / float(X.shape[0])
    return np.column_stack([np.ones(X.shape[1]), det, zo.factor_level2(cond).astype(np.float64)])


def test_link_functions_follow_family_c():
    eta = np.array([-800.0, -30.000001, -30.0, -1.0, 0.0, 2.5, 30.0, 30.000001, 800.0])
    mu = zo.linkinv(eta)
    eps = np.finfo(float).eps
    assert mu[0] == eps / (1 + eps) and mu[1] == eps / (1 + eps)
    assert mu[-1] == (1 / eps) / (1 + 1 / eps) and mu[-1] < 1.0
    np.testing.assert_allclose(mu[2:7], 1 / (1 + np.exp(-eta[2:7])), rtol=1e-15)
    me = zo.mu_eta(eta)
    assert me[0] == eps and me[1] == eps and me[-1] == eps and me[-2] == eps
    np.testing.assert_allclose(me[3:6], mu[3:6] * (1 - mu[3:6]), rtol=1e-14)
    y = np.array([0.0, 1.0, 1.0, 0.0])
    m = np.array([0.2, 0.2, 0.9, 0.9])
    np.testing.assert_allclose(zo.dev_resids(y, m), -2 * np.log([0.8, 0.2, 0.9, 0.1]), rtol=1e-15)


def test_prior_scale_rule():
    x = np.array([0.1, 0.4, 0.35, 0.8, 0.25])
    X = np.column_stack([np.ones(5), x, [0, 1, 1, 0, 1], [3.0, 3.0, 7.0, 7.0, 3.0], np.full(5, 2.0)])
    sc = zo.prior_scales(X)
    np.testing.assert_allclose(sc, [10.0, 2.5 / (2 * np.std(x, ddof=1)), 2.5, 2.5 / 4.0, 2.5], rtol=1e-15)
    assert zo.factor_level2(np.array([2, 1, 2, 1])).tolist() == [1, 0, 1, 0]          # levels sorted, not first-seen
    assert zo.factor_level2(np.array(["b", "a", "b"])).tolist() == [1, 0, 1]


@pytest.mark.parametrize("name", ["scDatEx", "scDatExSim"])
def test_fixed_point_of_the_published_algorithm(name):
    """At convergence (Gelman et al. 2008, eq. for the approximate EM): beta is the mode of the logistic likelihood
    under independent N(0, sigma_j^2) priors, and sigma_j^2 = (beta_j^2 + V_jj + nu s_j^2) / (1 + nu) with V the
    inverse of X'WX + diag(1 / sigma^2).  Both are checked from the definition, not through the IRLS code."""
    X, cond = _fixture(name)
    M = _design(X, cond)
    sc = zo.prior_scales(M)
    done = 0
    for g in range(X.shape[0]):
        y = (X[g] > 0).astype(np.float64)
        if (X[g] == 0).sum() == 0:
            continue
        r = zo.bayesglm_logit(y, M, scale=sc, epsilon=1e-15, maxit=20000)
        assert r["converged"]
        b, sd = r["coef"], r["prior_sd"]
        mu = 1.0 / (1.0 + np.exp(-(M @ b)))
        grad = M.T @ (y - mu) - b / sd ** 2                 # score of the penalised likelihood
        assert np.abs(grad).max() < 1e-6 * max(1.0, np.abs(M.T @ y).max())
        V = np.linalg.inv(M.T @ (M * (mu * (1 - mu))[:, None]) + np.diag(1.0 / sd ** 2))
        np.testing.assert_allclose(sd ** 2, (b ** 2 + np.diag(V) + sc ** 2) / 2.0, rtol=1e-6)
        np.testing.assert_allclose(r["se"], np.sqrt(np.diag(V)), rtol=1e-6)
        # the default stopping rule (epsilon 1e-8 on the deviance) stops within 1e-5 of that fixed point
        q = zo.bayesglm_logit(y, M, scale=sc)
        np.testing.assert_allclose(q["coef"], b, rtol=1e-4, atol=1e-6)
        assert 2 <= q["iter"] < 100
        done += 1
        if done == 12:
            break
    assert done >= 8


def test_oracle_test_zeroes_rules():
    X, cond = _fixture("scDatEx")
    p = zo.test_zeroes(X[:40], cond)
    nz = (X[:40] == 0).sum(axis=1)
    assert np.array_equal(np.isnan(p), nz == 0)             # R/classify.dd.R:267
    ok = ~np.isnan(p)
    assert ((p[ok] > 0) & (p[ok] <= 1)).all()
    # `these` selects rows but detection still counts every row of dat (:263)
    sel = np.array([3, 17, 5])
    np.testing.assert_array_equal(zo.test_zeroes(X[:40], cond, these=sel), p[sel])
    # swapping the condition labels changes the baseline level, hence the posterior mode -- but only slightly
    flipped = np.where(cond == cond[0], "b", "a")
    straight = np.where(cond == cond[0], "a", "b")
    pa, pb = zo.test_zeroes(X[:40], straight), zo.test_zeroes(X[:40], flipped)
    assert not np.allclose(pa[ok], pb[ok], rtol=1e-12) and np.allclose(pa[ok], pb[ok], rtol=0.2, atol=0.02)


def test_irls_core_against_an_independent_logistic_regression():
    """With an (almost) flat normal prior -- huge scale, df -> infinity switches the EM update off -- bayesglm's loop
    is plain IRLS: coefficients must equal scikit-learn's unpenalised maximum-likelihood fit and the standard errors
    the inverse Fisher information."""
    from sklearn.linear_model import LogisticRegression
    X, cond = _fixture("scDatEx")
    M = _design(X, cond)
    done = 0
    for g in range(200):
        y = (X[g] > 0).astype(np.float64)
        lvl = M[:, 2] == 1
        if min(y[lvl].sum(), (1 - y[lvl]).sum(), y[~lvl].sum(), (1 - y[~lvl]).sum()) < 6:
            continue                                     # (quasi-)separation: the maximum-likelihood fit runs away
        r = zo.bayesglm_logit(y, M, scale=np.full(3, 1e8), df=1e30, epsilon=1e-14, maxit=200)
        ref = LogisticRegression(C=np.inf, fit_intercept=False, tol=1e-12, max_iter=5000).fit(M, y)
        np.testing.assert_allclose(r["coef"], ref.coef_[0], rtol=2e-4, atol=2e-4)
        mu = 1.0 / (1.0 + np.exp(-(M @ r["coef"])))
        cov = np.linalg.inv(M.T @ (M * (mu * (1 - mu))[:, None]))
        np.testing.assert_allclose(r["se"], np.sqrt(np.diag(cov)), rtol=1e-6)
        done += 1
    assert done >= 5
