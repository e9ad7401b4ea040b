"""testZeroes on the device (scdd_test_zeroes, csrc/zeroes.cuh) against oracle/zeroes_oracle.py -- R/classify.dd.R:262-286.
Tolerance: 1e-8 relative on coefficients, standard errors and p-values (north_star's floating-point bar); iteration
counts and the NA pattern must be equal."""
import os

import numpy as np
import pytest

from oracle import zeroes_oracle as zo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    d = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=True)
    return d["normcounts"], d["condition"]


def _compare(X, cond, these=None, order="C"):
    from scdd_b200 import hotpath as hp
    want, det = zo.test_zeroes(X, cond, these=these, detail=True)
    got, gdet, st = hp.test_zeroes(np.array(X, order=order), cond, these=these, detail=True)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-8, atol=1e-300)
    for j in np.nonzero(ok)[0]:
        r = det[j]
        assert int(gdet[j, 6]) == r["iter"], (j, gdet[j, 6], r["iter"])
        np.testing.assert_allclose(gdet[j, 0:3], r["coef"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(gdet[j, 3:6], r["se"], rtol=1e-8)
        np.testing.assert_allclose(gdet[j, 7], r["dev"], rtol=1e-10)
    assert st["total_ms"] > 0
    return got


@pytest.mark.parametrize("name", ["scDatEx", "scDatExSim"])
def test_fixtures_all_genes_both_layouts(name):
    X, cond = _fixture(name)
    a = _compare(X, cond, order="C")
    b = _compare(X, cond, order="F")                       # an R matrix is column-major
    assert np.array_equal(a, b, equal_nan=True)


def test_row_subset_keeps_detection_over_all_rows():
    X, cond = _fixture("scDatEx")
    sel = np.array([7, 3, 250, 499, 3])
    got = _compare(X, cond, these=sel)
    full = _compare(X, cond)
    assert np.array_equal(got, full[sel], equal_nan=True)


def test_edge_cases():
    rng = np.random.default_rng(5)
    nc = 77                                                # not a multiple of 32
    cond = np.array([2] * 40 + [1] * 37)                   # first-seen level is NOT the baseline of factor(cond)
    X = np.exp(rng.normal(1.0, 1.0, size=(64, nc))) * (rng.random((64, nc)) < rng.uniform(0.2, 0.95, size=(64, 1)))
    X[0] = 0.0                                             # no nonzero cell at all
    X[1] = 1.0; X[1, 5] = 0.0                              # a single zero
    X[2] = 3.0                                             # no zero: NA
    X[3] = 0.0; X[3, :40] = 2.0                            # zeros in one condition only (separation; the prior bounds it)
    X[4] = 0.0; X[4, 40:] = 2.0
    X[5] = 0.0; X[5, 11] = 9.0                             # a single nonzero
    got = _compare(X, cond)
    assert np.isnan(got[2]) and not np.isnan(got[[0, 1, 3, 4, 5]]).any()
    assert got[3] < 1e-3 and got[4] < 1e-3


def test_long_genes_and_many_genes():
    rng = np.random.default_rng(11)
    nc = 5000
    cond = np.array(["ctl", "trt"])[(rng.random(nc) < 0.45).astype(int)]
    depth = rng.uniform(0.2, 1.0, size=nc)                 # per-cell capture efficiency -> detection varies
    rate = rng.uniform(0.05, 0.9, size=(300, 1))
    shift = np.where(rng.random((300, 1)) < 0.3, 0.6, 1.0) ** (cond == "trt")[None, :]
    X = (rng.random((300, nc)) < rate * depth[None, :] * shift) * np.exp(rng.normal(size=(300, nc)))
    _compare(X, cond, these=np.arange(0, 300, 7))


def test_bad_arguments_fail_loudly():
    from scdd_b200 import hotpath as hp
    X, cond = _fixture("scDatExSim")
    with pytest.raises(ValueError):
        hp.test_zeroes(X, np.zeros(X.shape[1]))
    with pytest.raises(hp.ScddError):
        hp.test_zeroes(X, cond, these=np.array([X.shape[0]]))


def test_scdd_with_test_zeroes_through_the_api():
    """R/scDD.R:517-526, :568-588: DZ / NS relabelling, zero.pvalue columns, Fisher's combined p-value"""
    from scdd_b200 import api
    X, cond = _fixture("scDatExSim")
    X = X.copy()
    X[:6, :] *= (np.arange(X.shape[1]) % 3 != 0)           # more zeros in a few genes
    X[6:9, cond == cond[0]] *= (np.arange((cond == cond[0]).sum()) % 2 == 0)   # condition-specific dropout: DZ candidates
    first = cond == cond[0]
    X[9, :] = 0.0; X[9, :4] = 1.5                          # too few nonzero cells: only tested for DZ, and not DZ
    X[10, first] = 0.0; X[10, np.nonzero(first)[0][:2]] = 2.0   # expressed in one condition only, < 3 cells in the other: DZ
    sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
    prior = dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01)
    out = api.scDD(sce, prior_param=prior, permutations=0, testZeroes=True, verbose=False)
    G = api.results(out)
    pz = zo.test_zeroes(X, cond)
    np.testing.assert_allclose(G["zero.pvalue"], pz, rtol=1e-8, atol=1e-300)
    adj = api.p_adjust_BH(pz)
    np.testing.assert_allclose(G["zero.pvalue.adj"], adj, rtol=1e-8, atol=1e-300)
    cats = np.array(G["DDcategory"], dtype=object)
    dd = np.isin(cats, ["DE", "DP", "DM", "DB"])
    with np.errstate(invalid="ignore"):
        assert (cats[~dd & (adj < 0.025)] == "DZ").all()
        assert (cats[~dd & (adj >= 0.025)] == "NS").all()
    assert cats[10] == "DZ" and cats[9] == "NS" and np.isnan(G["nonzero.pvalue"][[9, 10]]).all()
    # Fisher's method against scipy's chi-square tail
    from scipy import stats
    pn = G["nonzero.pvalue"]
    both = ~np.isnan(pn) & ~np.isnan(pz)
    np.testing.assert_allclose(G["combined.pvalue"][both], stats.chi2.sf(-2 * (np.log(pn[both]) + np.log(pz[both])), 4),
                               rtol=1e-10)
    one = np.isnan(pn) & ~np.isnan(pz)
    np.testing.assert_array_equal(G["combined.pvalue"][one], G["zero.pvalue"][one])
    # the KS branch used level / 2 for significance (R/scDD.R:362-363)
    assert out.metadata["_scdd_b200"]["stats"]["testZeroes"]["total_ms"] > 0
