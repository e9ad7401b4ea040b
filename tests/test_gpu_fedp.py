"""GPU parity of scdd_fedp (feDP, R/classify.dd.R:314-383) against oracle/fedp_oracle.py.

Tolerances: cluster counts and NS/DP/NC labels identical; p-values within 1e-9 relative (Welch / lm statistics are
sums in a different order than R's long-double loops; Fisher sums the same table probabilities in a different order)."""
import os

import numpy as np
import pytest

from oracle import fedp_oracle as fo
from scdd_b200 import api, hotpath as hp

pytestmark = pytest.mark.gpu
PRIORS = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)


def _split(labels_oa, labels_c, c01, off):
    oa, c1, c2 = [], [], []
    for g in range(off.size - 1):
        s = slice(off[g], off[g + 1])
        oa.append(labels_oa[s])
        c1.append(labels_c[s][c01[s] == 0])
        c2.append(labels_c[s][c01[s] == 1])
    return oa, c1, c2


def _compare(X, cond, cls_oa, cls_c, adjust, sig):
    logy, c01, off = hp.gene_csr(X, cond)
    ng = X.shape[0]
    C = (X > 0).sum(axis=0) / float(ng)
    cdr = np.broadcast_to(C, X.shape)[X > 0] if adjust else None
    p_shift, p_fisher, cnt, st = hp.fedp(logy, c01, off, cls_oa, cls_c, cdr=cdr, min_size=3)
    oa, c1, c2 = _split(cls_oa, cls_c, c01, off)
    n_fisher = 0
    for g in range(ng):
        ps, pf, c = fo.fedp_gene(X[g], cond, C, oa[g], c1[g], c2[g], adjust_perms=adjust, min_size=3)
        assert tuple(cnt[g]) == c, g
        np.testing.assert_allclose(p_shift[g], ps, rtol=1e-9, atol=1e-300, err_msg=f"gene {g}")
        if np.isnan(pf):
            assert np.isnan(p_fisher[g]), g
        else:
            n_fisher += 1
            np.testing.assert_allclose(p_fisher[g], pf, rtol=1e-9, atol=1e-300, err_msg=f"gene {g}")
    # the labelling on top: same categories as the gene-by-gene restatement
    ns, pv, cat = api.feDP(logy, c01, off, sig, cls_oa, cls_c, cdr=cdr, min_size=3)
    opv, ocat, ons = fo.feDP(X, cond, sig, oa, c1, c2, adjust_perms=adjust, min_size=3)
    assert np.array_equal(ns, ons) and list(cat) == list(ocat)
    np.testing.assert_allclose(pv, opv, rtol=1e-9)
    return n_fisher, st


@pytest.mark.parametrize("adjust", [False, True])
def test_fedp_on_observed_fits_of_the_example_data(golden_dir, adjust):
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    logy, c01, off = hp.gene_csr(X, cond)
    _, cls_oa, cls_c, _, _, _ = hp.fit_observed(logy, c01, off, cfg=hp.make_cfg(**PRIORS))
    n_fisher, _ = _compare(X, cond, cls_oa, cls_c, adjust, sig=[0, 3, 11])
    assert n_fisher >= 1          # the EP / DP rows of the simulated example are bimodal in both conditions


def test_fedp_rx2_tables_and_empty_sig():
    """synthetic labels that make every gene pass the consistency test with r = 2..5 clusters"""
    rng = np.random.default_rng(9)
    ng, nc = 40, 160
    cond = np.array(["ctl", "trt"])[(np.arange(nc) >= nc // 2).astype(int)]
    X = np.exp(rng.normal(2.0, 0.7, (ng, nc)))
    X[rng.random(X.shape) < 0.25] = 0
    logy, c01, off = hp.gene_csr(X, cond)
    cls_oa = np.zeros(logy.size, dtype=np.int32)
    for g in range(ng):
        r = 2 + g % 4
        n = off[g + 1] - off[g]
        lab = rng.integers(1, r + 1, n)
        if g % 3 == 0:                         # association between cluster and condition
            sl = c01[off[g]:off[g + 1]]
            lab = np.where((sl == 1) & (rng.random(n) < 0.5), 1, lab)
        if g == 8:
            lab[:2] = 3                        # r = 2 plus an outlier cluster of 2 cells: dropped by findOutliers
        cls_oa[off[g]:off[g + 1]] = lab
    n_fisher, st = _compare(X, cond, cls_oa, cls_oa.copy(), False, sig=[1])
    assert n_fisher >= 30 and st["failed_sets"] == 0
    ns, pv, cat = api.feDP(logy, c01, off, [], cls_oa, cls_oa)
    assert ns.size == 0 and pv.size == 0      # R: (1:n)[-integer(0)] is empty


def test_fedp_larger_tables_are_split_and_deterministic():
    """2 400-cell genes: 2 x 2 tables are a single job, a 4-cluster table is enumerated by several CTAs; results are
    bitwise reproducible and match the restatement"""
    rng = np.random.default_rng(10)
    nc = 2400
    cond = np.repeat([0, 1], nc // 2)
    X = np.exp(rng.normal(2.0, 0.7, (3, nc)))
    logy, c01, off = hp.gene_csr(X, cond)
    cls = np.zeros(logy.size, dtype=np.int32)
    cls[off[0]:off[1]] = 1 + (rng.random(nc) < 0.4 + 0.05 * cond)
    cls[off[1]:off[2]] = rng.integers(1, 4, nc)
    cls[off[2]:off[3]] = rng.integers(1, 5, nc)          # 4 clusters of ~600: ~6.9e6 work items, more than one job
    a = hp.fedp(logy, c01, off, cls, cls)
    b = hp.fedp(logy, c01, off, cls, cls)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    assert a[3]["fit_sets"] > 3               # the 4-row table was split
    for g in range(3):
        s = slice(off[g], off[g + 1])
        tab = np.array([[int(((cls[s] == l) & (c01[s] == 0)).sum()), int(((cls[s] == l) & (c01[s] == 1)).sum())]
                        for l in np.unique(cls[s])])
        np.testing.assert_allclose(a[1][g], fo.fisher_rx2(tab), rtol=1e-9)


def test_scdd_categories_through_the_api(golden_dir):
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    pri = dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01)
    out = api.scDD(api.SingleCellExperiment({"normcounts": X}, {"condition": cond}), prior_param=pri, permutations=20,
                   testZeroes=False, param={"RNGseed": 7}, verbose=False)
    G = api.results(out, "Genes")
    cats = G["DDcategory"]
    padj = G["nonzero.pvalue.adj"]
    # significant genes get classifyDD's pattern; the others come out of feDP as NS / DP, or as classifyDD's verdict
    # on feDP's "NC" candidates
    assert all(c in ("DE", "DP", "DM", "DB", "NC") for c, p in zip(cats, padj) if p < 0.05)
    assert set(cats) <= {"NS", "DE", "DP", "DM", "DB", "NC"}
    # classify=False: exactly the genes the reference hands to classifyDD stay unlabelled
    out0 = api.scDD(api.SingleCellExperiment({"normcounts": X}, {"condition": cond}), prior_param=pri, permutations=20,
                    testZeroes=False, param={"RNGseed": 7}, verbose=False, classify=False)
    cats0 = api.results(out0, "Genes")["DDcategory"]
    assert all(c is None for c, p in zip(cats0, padj) if p < 0.05)
    assert set(c for c in cats0 if c is not None) <= {"NS", "DP"}


def test_scdd_with_filtered_and_constant_genes():
    """genes below min.nonzero and genes with constant nonzero values are skipped (R/scDD.R:270-310): NA rows in the
    Genes table, default 0/1 entries in the Zhat matrices, and the fitted genes unaffected by their presence"""
    rng = np.random.default_rng(21)
    ng, nc = 24, 80
    cond = np.repeat(["A", "B"], nc // 2)
    X = np.exp(rng.normal(2.0, 0.6, (ng, nc)))
    X[rng.random(X.shape) < 0.2] = 0
    X[3, :] = 0
    X[3, :2] = 5.0                                   # two nonzero cells in one condition only: not fitted
    X[10, :] = np.where(X[10] > 0, 7.0, 0.0)         # constant nonzero values: skipped
    X = np.asfortranarray(X)                         # an R-style (column-major) matrix goes through unchanged
    pri = dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01)
    out = api.scDD(api.SingleCellExperiment({"normcounts": X}, {"condition": cond}), prior_param=pri, permutations=8,
                   testZeroes=False, param={"RNGseed": 3}, verbose=False)
    G = api.results(out, "Genes")
    tofit = out.metadata["_scdd_b200"]["tofit"]
    assert 3 not in tofit and 10 not in tofit and tofit.size == ng - 2
    for g in (3, 10):
        assert np.isnan(G["nonzero.pvalue"][g]) and np.isnan(G["Clusters.combined"][g]) and G["DDcategory"][g] is None
    Z = api.results(out, "Zhat.combined")
    assert np.array_equal(Z[3], (X[3] != 0).astype(float)) and np.array_equal(Z[10], (X[10] != 0).astype(float))
    assert np.array_equal(Z == 0, X == 0)
    # the same genes alone give the same p-values: streams are assigned by position among the fitted genes
    keep = np.setdiff1d(np.arange(ng), [3, 10])
    out2 = api.scDD(api.SingleCellExperiment({"normcounts": np.ascontiguousarray(X[keep])}, {"condition": cond}),
                    prior_param=pri, permutations=8, testZeroes=False, param={"RNGseed": 3}, verbose=False)
    assert np.array_equal(api.results(out2, "Genes")["nonzero.pvalue"], G["nonzero.pvalue"][keep])
    assert np.array_equal(api.results(out2, "Zhat.c1"), api.results(out, "Zhat.c1")[keep])
