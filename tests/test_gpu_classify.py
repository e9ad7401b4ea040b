"""classifyDD (R/classify.dd.R:86-238): the oracle's restatement on unambiguous clusterings (CPU), and the device
implementation (scdd_classify: host bookkeeping + Monte-Carlo kernel) against the oracle (GPU)."""
import os

import numpy as np
import pytest

from oracle import classify_oracle as co
from classify_cases import clear_cases

PRI = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_decision_tree_on_clear_cases():
    want, y, c, off, oa, cc = clear_cases()
    got = co.classify_dd(y, c, off, np.arange(len(want)), oa, cc, PRI, rng=np.random.default_rng(1),
                         p_shift=np.full(len(want), 1e-6))
    assert got.tolist() == want
    # the multimodal-DE rule needs the overall mean shift as well
    got2 = co.classify_dd(y, c, off, np.array([5]), oa, cc, PRI, rng=np.random.default_rng(1), p_shift=np.full(len(want), 0.5))
    assert got2.tolist() == ["NC"]
    # a gap in the label set is closed before the posterior parameters are taken (R/classify.dd.R:124-158)
    assert co.compress_labels(np.array([1, 3, 3, 1])).tolist() == [1, 2, 2, 1]
    assert co.lu_outlier(np.array([1, 1, 1, 2, 2, 3, 3, 3]), 3) == 2


@pytest.mark.gpu
def test_device_student_t_has_the_right_distribution():
    from scipy import stats
    from scdd_b200 import hotpath as hp
    for df in (1, 2, 3, 7, 30, 400):
        x = hp.student_t_draws(df, 200000, seed=df)
        assert np.isfinite(x).all()
        assert stats.kstest(x, stats.t(df).cdf).pvalue > 1e-3, df
    a, b = hp.student_t_draws(5, 1000, seed=9), hp.student_t_draws(5, 1000, seed=9)
    assert np.array_equal(a, b) and not np.array_equal(a, hp.student_t_draws(5, 1000, seed=10))


@pytest.mark.gpu
def test_device_classify_clear_cases_and_details():
    from scdd_b200 import hotpath as hp
    want, y, c, off, oa, cc = clear_cases()
    cfg = hp.make_cfg(**PRI)
    names, comp_off, comp, st = hp.classify(y, c, off, np.arange(len(want)), oa, cc, p_shift=np.full(len(want), 1e-6),
                                            cfg=cfg, seed=4, detail=True)
    assert names.tolist() == want
    onames, ocomp = co.classify_dd(y, c, off, np.arange(len(want)), oa, cc, PRI, rng=np.random.default_rng(1),
                                   p_shift=np.full(len(want), 1e-6), detail=True)
    assert comp.shape == ocomp.shape and comp_off[-1] == ocomp.shape[0]
    np.testing.assert_allclose(comp[:, :6], ocomp[:, :6], rtol=1e-12)        # df, scale, mean of both clusters
    assert np.array_equal(comp[:, 6], ocomp[:, 6])                            # far from the threshold: same flags
    assert hp.classify(y, c, off, np.array([5]), oa, cc, p_shift=np.full(len(want), 0.5), cfg=cfg)[0].tolist() == ["NC"]
    # same seed, same answer; a subset is classified as inside the full list
    again = hp.classify(y, c, off, np.arange(len(want)), oa, cc, p_shift=np.full(len(want), 1e-6), cfg=cfg, seed=4)[0]
    assert again.tolist() == want
    assert hp.classify(y, c, off, np.zeros(0, dtype=np.int32), oa, cc, cfg=cfg)[0].size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["scDatExSim", "scDatEx"])
def test_device_classify_on_fixture_clusterings(name):
    """every gene of the bundled data sets, clusterings from the device fits: the comparison lists equal the oracle's,
    and the flags / categories agree wherever three oracle seeds agree among themselves"""
    from scdd_b200 import hotpath as hp
    d = np.load(os.path.join(GOLD, f"{name}.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    logy, c01, off = hp.gene_csr(X, cond)
    cfg = hp.make_cfg(**PRI)
    fits, cls_oa, cls_c, bf, den, _ = hp.fit_observed(logy, c01, off, cfg=cfg)
    p_shift = hp.fedp(logy, c01, off, cls_oa, cls_c)[0]
    genes = np.arange(X.shape[0])
    names, comp_off, comp, st = hp.classify(logy, c01, off, genes, cls_oa, cls_c, p_shift=p_shift, cfg=cfg, seed=11, detail=True)
    runs = [co.classify_dd(logy, c01, off, genes, cls_oa, cls_c, PRI, rng=np.random.default_rng(s), p_shift=p_shift, detail=True)
            for s in (1, 2, 3)]
    ocomp = runs[0][1]
    assert comp.shape == ocomp.shape
    np.testing.assert_allclose(comp[:, :6], ocomp[:, :6], rtol=1e-10)
    # A comparison is decided by 10,000 draws: "different" iff no sampled difference crosses zero.  With z the
    # standardised distance of the two posterior means, z > 7 (and enough degrees of freedom for light tails) cannot
    # cross, z < 2.5 crosses with certainty; in between the outcome is a coin whose bias depends on z.
    df1, s1, m1, df2, s2, m2 = comp[:, :6].T
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.abs(m1 - m2) / np.sqrt(s1 ** 2 * df1 / np.maximum(df1 - 2, 0.5) + s2 ** 2 * df2 / np.maximum(df2 - 2, 0.5))
    sure1 = (z > 7) & (np.minimum(df1, df2) >= 10)
    sure0 = z < 2.5
    assert (sure1 | sure0).mean() > 0.5
    assert (comp[sure1, 6] == 1).all() and (comp[sure0, 6] == 0).all()
    for r in runs:
        assert (r[1][sure1, 6] == 1).all() and (r[1][sure0, 6] == 0).all()
        assert (comp[:, 6] == r[1][:, 6]).mean() > 0.95            # coin flips in the band between
    # a gene all of whose comparisons are decided has ONE possible category
    decided = np.array([(sure1 | sure0)[comp_off[g]:comp_off[g + 1]].all() for g in genes])
    assert decided.mean() > 0.3
    assert np.array_equal(names[decided], runs[0][0][decided])
    # the coin flips have the oracle's bias: fraction of "different" verdicts in the band, device vs three oracle runs
    band = ~(sure1 | sure0)
    if band.sum() > 50:
        o = np.mean([r[1][band, 6].mean() for r in runs])
        assert abs(comp[band, 6].mean() - o) < 4 * np.sqrt(0.25 / band.sum()) + 0.02
    assert st["total_ms"] > 0
