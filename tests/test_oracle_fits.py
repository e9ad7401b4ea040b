"""Oracle self-checks (CPU): the C oracle against the independent numpy restatement,
closed forms and invariants, on the reference's bundled example data.

PARITY UNPINNED: the reference ships no tests/golden vectors, so the
oracle is pinned by (i) closed forms, (ii) a second implementation from the same spec
(tests/np_restatement.py), (iii) invariants.
"""
import math
import os

import numpy as np
import pytest

import np_restatement as npr
from oracle import pyoracle as po

PRIORS = dict(alpha=0.01, m0=0.0, s0=0.01, a0=0.01, b0=0.01)  # R/scDD.R:203 example priors


def _halves(golden_dir, name, genes):
    d = np.load(os.path.join(golden_dir, name), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    ref = cond[0]
    for g in genes:
        y = X[g]
        nz = y > 0
        ly = np.log(y[nz])
        c = cond[nz]
        yield g, ly, ly[c == ref], ly[c != ref]


def test_g1_closed_form():
    rng = np.random.default_rng(0)
    y = rng.normal(2.0, 1.5, size=137)
    rows = po.fit_table(y)
    r = rows[0]
    mu = y.mean()
    ss = ((y - mu) ** 2).mean()
    assert r["valid"]
    assert math.isclose(r["mean"][0], mu, rel_tol=1e-14)
    assert math.isclose(r["sigsq"][0], ss, rel_tol=1e-13)
    ll = -0.5 * y.size * (math.log(2 * math.pi) + 1 + math.log(ss))
    assert math.isclose(r["loglik"], ll, rel_tol=1e-13)
    assert math.isclose(r["bic"], 2 * ll - 2 * math.log(y.size), rel_tol=1e-13)


def test_loglik_matches_final_parameters():
    """me1v returns z/loglik computed from the returned parameters (App. A.3)."""
    rng = np.random.default_rng(1)
    y = np.concatenate([rng.normal(0, 1, 80), rng.normal(5, 0.7, 60)])
    for r in po.fit_table(y)[1:]:
        if not r["valid"]:
            continue
        mu, ss, pr = r["mean_em"], r["sigsq_em"], r["pro_em"]
        dens = pr[None, :] * np.exp(-0.5 * (y[:, None] - mu[None, :]) ** 2 / ss[None, :]) / np.sqrt(2 * np.pi * ss[None, :])
        assert math.isclose(np.log(dens.sum(1)).sum(), r["loglik"], rel_tol=1e-10)
        assert math.isclose(r["bic"], 2 * r["loglik"] - (3 * r["G"] - 1) * math.log(y.size), rel_tol=1e-13)


@pytest.mark.parametrize("name,genes", [("scDatExSim.npz", range(0, 30, 3)), ("scDatEx.npz", range(0, 500, 41))])
def test_c_oracle_matches_numpy_restatement(golden_dir, name, genes):
    for g, ly, y1, y2 in _halves(golden_dir, name, genes):
        for y in (ly, y1, y2):
            rows_c = po.fit_table(y)
            rows_n = npr.fit_table(y)
            for rc in rows_c:
                assert rc["valid"] == (rc["G"] in rows_n), (g, rc["G"])
                if rc["valid"]:
                    rn = rows_n[rc["G"]]
                    assert math.isclose(rc["bic"], rn["bic"], rel_tol=1e-8), (g, rc["G"])
                    assert np.array_equal(rc["labels"], rn["classification"]), (g, rc["G"])
            fc = po.mclust_restricted(y)
            fn = npr.mclustRestricted(y)
            assert fc["G"] == len(fn["mean"])
            assert np.array_equal(fc["class"], fn["class"])
            np.testing.assert_allclose(fc["mean"], fn["mean"], rtol=1e-8)
            np.testing.assert_allclose(fc["var"], fn["var"], rtol=1e-8)
            jc = po.joint_posterior(y, fc, **PRIORS)
            jn = npr.jointPosterior(y, fn, PRIORS["alpha"], PRIORS["m0"], PRIORS["s0"], PRIORS["a0"], PRIORS["b0"])
            assert math.isclose(jc, jn, rel_tol=1e-10)


def test_joint_posterior_label_permutation_invariance():
    rng = np.random.default_rng(3)
    y = rng.normal(size=50)
    cls = rng.integers(1, 4, size=50).astype(np.int32)
    a = po.joint_posterior(y, cls, **PRIORS)
    perm = {1: 3, 2: 1, 3: 2}
    b = po.joint_posterior(y, np.array([perm[c] for c in cls], dtype=np.int32), **PRIORS)
    assert math.isclose(a, b, rel_tol=1e-13)


def test_joint_posterior_noncontiguous_labels_is_inf():
    """quirk: labels {1,3} -> r=2 -> nk[2]=0 -> lgamma(0)=Inf."""
    y = np.arange(10, dtype=np.float64)
    cls = np.array([1] * 5 + [3] * 5, dtype=np.int32)
    assert math.isinf(po.joint_posterior(y, cls, **PRIORS))


def test_restricted_merges_unimodal_to_one_component():
    rng = np.random.default_rng(5)
    f = po.mclust_restricted(rng.normal(3, 1, 400))
    assert f["G"] == 1 and f["model"] == "X" and f["df"] == 2
    f = po.mclust_restricted(np.concatenate([rng.normal(0, 0.5, 200), rng.normal(6, 0.5, 200)]))
    assert f["G"] == 2 and f["model"] == "V" and f["df"] == 5
    assert f["mean"][0] < f["mean"][1]
    assert set(np.unique(f["class"])) == {1, 2}


def test_perm_batch_equals_manual_loop(golden_dir):
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    X, cond = d["normcounts"][:3], d["condition"]
    logy, c01, off = [], [], [0]
    for y in X:
        nz = y > 0
        logy.append(np.log(y[nz]))
        c01.append((cond[nz] != cond[0]).astype(np.int32))
        off.append(off[-1] + int(nz.sum()))
    logy = np.concatenate(logy)
    c01 = np.concatenate(c01)
    off = np.array(off, dtype=np.int64)
    B = 4
    rng = po.RRng(seed=7, kind="mt")
    perms = np.concatenate([np.concatenate([rng.sample(off[g + 1] - off[g]) for _ in range(B)]) for g in range(3)])
    cfg = po.make_cfg(alpha=0.01)
    out = po.perm_batch(logy, c01, off, B, perm_idx=perms, cfg=cfg, nthreads=2)
    for g in range(3):
        n = off[g + 1] - off[g]
        y = logy[off[g]:off[g + 1]]
        c = c01[off[g]:off[g + 1]]
        for b in range(B):
            p = perms[off[g] * B + b * n: off[g] * B + (b + 1) * n]
            ny = y[p - 1]
            y1, y2 = ny[c == 0], ny[c == 1]
            v = po.joint_posterior(y1, po.mclust_restricted(y1), 0.01, 0, 0.01, 0.01, 0.01) + \
                po.joint_posterior(y2, po.mclust_restricted(y2), 0.01, 0, 0.01, 0.01, 0.01)
            assert v == out[g, b]
