"""GPU parity tests added in round 2 (all through the C ABI, against the CPU oracle on the same inputs):

  * BASELINE config 2 at its stated size: scDatEx (500 genes) and scDatExSim (30 genes), permutations = 100,
    device-drawn, plus testKS
  * the constant-subset jitter of mclustRestricted (R/mclust.restricted.R:48-50), drawn mid-stream on the device
  * parallelBy = "Permutations" stream shape (R/scDD.R:414-455, R/permutations.R:99,177)
  * the three jointPosterior terms of the adjusted path one by one at 1e-8 (R/permutations.R:261-267)
  * config-5 shaped genes (50,000 cells, adjust.perms) against the oracle
  * the sharded (multi-device) code path, the interrupt hook, the permutation kernel's group sizes
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from scdd_b200 import hotpath as hp

pytestmark = pytest.mark.gpu
RTOL = 1e-8
PRIORS = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)   # R/scDD.R:203


def _load(golden_dir, name, genes=None):
    d = np.load(os.path.join(golden_dir, name), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    if genes is not None:
        X = X[genes]
    return X, cond


# ------------------------------------------------------------------------------------------------ config 2
@pytest.mark.parametrize("name,ngenes", [("scDatExSim.npz", 30), ("scDatEx.npz", 500)])
def test_config2_at_stated_size(golden_dir, name, ngenes):
    """scDD(permutations = 100) + testKS on the bundled data sets, every gene, every permutation"""
    X, cond = _load(golden_dir, name)
    assert X.shape[0] == ngenes
    B = 100
    logy, c01, off = hp.gene_csr(X, cond)
    seeds = hp.gene_seeds(816, ngenes)
    cfg, ocfg = hp.make_cfg(**PRIORS), po.make_cfg(**PRIORS)
    out, st, terms = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=cfg, return_terms=True)
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cfg=ocfg)
    assert ref["rc"] == 0 and st["failed_sets"] == 0 and st["iter_cap_hits"] == 0
    assert st["fit_sets"] == 2 * ngenes * B
    np.testing.assert_allclose(out, ref["bf_perm"], rtol=RTOL)
    np.testing.assert_allclose(terms, ref["jp_terms"], rtol=RTOL)
    assert st["const_sets"] == ref["n_jittered"]
    # observed statistic and p-value counts (R/scDD.R:467-473)
    fits, cls_oa, cls_c, bf, den, _ = hp.fit_observed(logy, c01, off, cfg=cfg)
    ofits, ocls_oa, ocls_c, obf, oden = po.genefit_batch(logy, c01, off, cfg=ocfg)
    assert np.array_equal(cls_oa, ocls_oa) and np.array_equal(cls_c, ocls_c)
    assert [fits[g, w]["G"] for g in range(ngenes) for w in range(3)] == [ofits[g][w]["G"] for g in range(ngenes) for w in range(3)]
    np.testing.assert_allclose(bf, obf, rtol=RTOL)
    cnt, ocnt = (out > bf[:, None]).sum(1), (ref["bf_perm"] > obf[:, None]).sum(1)
    # "identical given the same permutations, apart from Bayes-factor ties within tolerance": count the ties
    near = (np.abs(ref["bf_perm"] - obf[:, None]) <= RTOL * np.abs(obf[:, None])).sum()
    assert (cnt != ocnt).sum() <= near
    if near == 0:
        assert np.array_equal(cnt, ocnt)
    # testKS (R/Test_KS.R:55-83): bit-exact
    vals, ck, offk = hp.gene_csr(X, cond, log=False)
    D, p = hp.ks(vals, ck, offk)
    oD, op = po.ks_batch(vals, ck, offk)
    assert np.array_equal(D, oD) and np.array_equal(p, op)


# ------------------------------------------------------------------------------------------------ jitter
def _tied_genes(ngenes, seed=5):
    rng = np.random.default_rng(seed)
    ys, cs = [], []
    for g in range(ngenes):
        n1, n2 = 3 + g % 4, 3 + (g // 4) % 5
        n = n1 + n2
        y = np.full(n, np.log(2.0))
        k = max(2, n // 4)
        y[rng.choice(n, k, replace=False)] = np.log(rng.integers(3, 9, size=k).astype(float))
        ys.append(y)
        cs.append(np.r_[np.zeros(n1, dtype=np.int32), np.ones(n2, dtype=np.int32)][rng.permutation(n)])
    off = np.r_[0, np.cumsum([y.size for y in ys])].astype(np.int64)
    return np.concatenate(ys), np.concatenate(cs), off


@pytest.mark.parametrize("gs", ["4", "8", "32"])
def test_constant_subsets_are_jittered_from_the_gene_stream(gs, monkeypatch):
    """tie-heavy counts: many permuted halves are constant.  The device must draw the reference's runif jitter from the
    gene's own stream at the position R draws it, so that neither the call aborts nor later permutations shift."""
    monkeypatch.setenv("SCDD_PERMGEN_GS", gs)
    logy, c01, off = _tied_genes(40)
    ng = off.size - 1
    B = 24
    seeds = hp.gene_seeds(11, ng)
    cfg, ocfg = hp.make_cfg(**PRIORS), po.make_cfg(**PRIORS)
    out, st, seeds_out = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=cfg, return_seeds=True)   # must not raise
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cfg=ocfg)
    assert ref["n_jittered"] > 50
    assert st["const_sets"] == ref["n_jittered"] and st["failed_sets"] == 0
    assert np.isfinite(out).all()
    np.testing.assert_allclose(out, ref["bf_perm"], rtol=RTOL)
    # the streams stayed in step: continuing from the returned states = one call of B + 3 permutations
    more, _ = hp.perm_bf(logy, c01, off, 3, seed6=seeds_out, cfg=cfg)
    both, _ = hp.perm_bf(logy, c01, off, B + 3, seed6=seeds, cfg=cfg)
    assert np.array_equal(np.hstack([out, more]), both)
    # without a stream (host-drawn permutations) such a fit-set is reported, not fitted
    pi = np.concatenate([hp.sample_perms(seeds[g], int(off[g + 1] - off[g]), 2)[0].ravel() for g in range(ng)])
    with pytest.raises(hp.ScddError) as ei:
        hp.perm_bf(logy, c01, off, 2, perm_idx=pi, cfg=cfg)
    assert ei.value.code == -5


def _tied_genes_balanced_covariate(ngenes, seed):
    """tie groups of even size whose covariate alternates 0.25 / 0.75 inside every group: cov(y, C) = 0 exactly, so
    lm(y ~ C) has slope 0, the fitted values are constant and residual ties survive the exp/log round trip"""
    rng = np.random.default_rng(seed)
    ys, cs, Cs = [], [], []
    for g in range(ngenes):
        sizes = [2 * rng.integers(2, 5), 2 * rng.integers(1, 3)] + ([2] if g % 3 == 0 else [])
        y = np.concatenate([np.full(k, np.log(v)) for k, v in zip(sizes, (2.0, 5.0, 9.0))])
        C = np.concatenate([np.tile([0.25, 0.75], k // 2) for k in sizes])
        n = y.size
        c = np.zeros(n, dtype=np.int32)
        c[rng.choice(n, size=n - 3 - rng.integers(0, max(1, n - 6)), replace=False)] = 1
        if c.sum() < 3:
            c[:] = 0
            c[rng.choice(n, size=3, replace=False)] = 1
        p = rng.permutation(n)
        ys.append(y[p]); Cs.append(C[p]); cs.append(c)
    off = np.r_[0, np.cumsum([y.size for y in ys])].astype(np.int64)
    return np.concatenate(ys), np.concatenate(cs), off, np.concatenate(Cs)


def test_constant_subsets_adjusted_path():
    """adjust.perms: the jitter of constant subsets is drawn in the order the reference fits them -- oa, c1, c2
    (R/permutations.R:261-263)"""
    logy, c01, off, cdr = _tied_genes_balanced_covariate(30, seed=9)
    ng = off.size - 1
    seeds = hp.gene_seeds(3, ng)
    B = 16
    out, st, terms = hp.perm_bf(logy, c01, off, B, seed6=seeds, cdr=cdr, cfg=hp.make_cfg(**PRIORS), return_terms=True,
                                allow_failed=True)
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cdr=cdr, cfg=po.make_cfg(**PRIORS))
    assert ref["n_jittered"] > 20 and st["const_sets"] == ref["n_jittered"]
    assert np.array_equal(np.isnan(terms), np.isnan(ref["jp_terms"]))
    ok = ~np.isnan(ref["jp_terms"])
    np.testing.assert_allclose(terms[ok], ref["jp_terms"][ok], rtol=RTOL)


def test_scdd_api_on_tied_counts_does_not_abort():
    """one constant permuted half used to abort the whole scDD() call"""
    from scdd_b200 import api
    rng = np.random.default_rng(2)
    ng, nc = 60, 24
    X = np.zeros((ng, nc))
    for g in range(ng):
        nz = rng.choice(nc, size=rng.integers(8, 16), replace=False)
        X[g, nz] = rng.choice([1.0, 1.0, 1.0, 2.0, 5.0], size=nz.size)
    cond = np.repeat([1, 2], nc // 2)
    sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
    out = api.scDD(sce, prior_param=dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01), permutations=30,
                   testZeroes=False, param={"RNGseed": 4}, verbose=False)
    meta = out.metadata["_scdd_b200"]
    st = meta["stats"]["permutations"]
    assert st["const_sets"] > 0 and st["failed_sets"] == 0
    p = np.asarray(api.results(out, "Genes")["nonzero.pvalue"])[meta["tofit"]]
    assert np.isfinite(p).all() and ((p >= 0) & (p <= 1)).all()
    # the oracle on the same streams gives the same p-values
    logy, c01, off = hp.gene_csr(X[meta["tofit"]], cond)
    ref = po.perm_batch_ex(logy, c01, off, 30, seed6=hp.gene_seeds(4, meta["tofit"].size), cfg=po.make_cfg(**PRIORS))
    _, _, _, obf, _ = po.genefit_batch(logy, c01, off, cfg=po.make_cfg(**PRIORS))
    assert np.array_equal(p, (ref["bf_perm"] > obf[:, None]).sum(1) / 30.0)


# ------------------------------------------------------------------------------------------------ stream shapes
def test_parallel_by_permutations_streams(golden_dir):
    """seed_mode 1: permutation b of every gene from the start of stream b (bplapply(1:nperms), R/permutations.R:177)"""
    X, cond = _load(golden_dir, "scDatEx.npz", genes=slice(0, 40))
    logy, c01, off = hp.gene_csr(X, cond)
    B = 9
    seeds = hp.gene_seeds(77, B)
    cfg = hp.make_cfg(**PRIORS)
    dev, st = hp.perm_bf(logy, c01, off, B, seed6=seeds, seed_mode=1, cfg=cfg)
    pi = np.concatenate([np.stack([hp.sample_perms(seeds[b], int(off[g + 1] - off[g]), 1)[0][0] for b in range(B)]).ravel()
                         for g in range(40)])
    host, _ = hp.perm_bf(logy, c01, off, B, perm_idx=pi, cfg=cfg)
    assert np.array_equal(dev, host)                         # device and host drawings agree in this mode too
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, seed_mode=1, cfg=po.make_cfg(**PRIORS))
    np.testing.assert_allclose(dev, ref["bf_perm"], rtol=RTOL)
    per_gene, _ = hp.perm_bf(logy, c01, off, B, seed6=hp.gene_seeds(77, 40), cfg=cfg)
    assert not np.array_equal(per_gene, dev)
    # the host mirror of scDD(parallelBy = "Permutations") uses it
    from scdd_b200 import api
    sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
    out = api.scDD(sce, prior_param=dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01), permutations=B,
                   testZeroes=False, param={"RNGseed": 77}, parallelBy="Permutations", verbose=False)
    _, _, _, bf, _, _ = hp.fit_observed(logy, c01, off, cfg=cfg)
    assert np.array_equal(api.results(out, "Genes")["nonzero.pvalue"], (dev > bf[:, None]).sum(1) / float(B))


# ------------------------------------------------------------------------------------------------ adjusted path
def _check_adjusted(logy, c01, off, cdr, B, seed):
    ng = off.size - 1
    seeds = hp.gene_seeds(seed, ng)
    out, st, terms = hp.perm_bf(logy, c01, off, B, seed6=seeds, cdr=cdr, cfg=hp.make_cfg(**PRIORS), return_terms=True)
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cdr=cdr, cfg=po.make_cfg(**PRIORS))
    assert ref["rc"] == 0 and st["fit_sets"] == 3 * ng * B
    # every jointPosterior term on its own at the north-star tolerance ...
    np.testing.assert_allclose(terms, ref["jp_terms"], rtol=RTOL, atol=0)
    # ... and their combination JP(c1) + JP(c2) - JP(oa), which cancels: 1e-8 of the terms' magnitude
    scale = np.abs(ref["jp_terms"]).sum(axis=2)
    assert (np.abs(out - ref["bf_perm"]) <= RTOL * scale).all()
    assert np.array_equal(out, (terms[:, :, 0] + terms[:, :, 1]) - terms[:, :, 2])
    return out, ref


def test_adjusted_terms_one_by_one(golden_dir):
    X, cond = _load(golden_dir, "scDatExSim.npz")
    logy, c01, off = hp.gene_csr(X, cond)
    C = (X > 0).sum(axis=0) / X.shape[0]                              # R/scDD.R:457
    cdr = np.broadcast_to(C, X.shape)[X > 0]
    _check_adjusted(logy, c01, off, cdr, 12, seed=5)


def test_config5_shape_against_oracle():
    """BASELINE config 5 scaled down in genes only: 20 genes x 50,000 cells, adjust.perms with the CDR covariate,
    3 permutations (R/permutations.R:44-104).  oa fit-sets have ~44,000 points (CTA team, global slab)."""
    from scdd_b200.simulate import simulate_set_style
    X, cond = simulate_set_style(20, 25000, seed=55)
    logy, c01, off = hp.gene_csr(X, cond)
    C = (X > 0).sum(axis=0) / X.shape[0]
    cdr = np.broadcast_to(C, X.shape)[X > 0]
    assert int(np.diff(off).max()) > 30000
    _check_adjusted(logy, c01, off, cdr, 3, seed=21)


# ------------------------------------------------------------------------------------------------ sharded path
def test_sharded_path_two_shards_on_one_device(golden_dir):
    """The multi-device path of scdd_perm_bf (what the R shim uses on an 8-GPU box) with the same device listed twice:
    gather into pinned staging, first phase on cell-count shards, re-deal by measured cost, per-gene streams carried
    between the phases.  Must equal the single-device call bit for bit."""
    X, cond = _load(golden_dir, "scDatEx.npz", genes=slice(0, 96))
    logy, c01, off = hp.gene_csr(X, cond)
    seeds = hp.gene_seeds(3, 96)
    cfg = hp.make_cfg(**PRIORS)
    B = 40                                                     # > 32: two phases
    hp.set_devices([0])
    one, st1, s1, t1 = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=cfg, return_seeds=True, return_terms=True)
    try:
        hp.set_devices([0, 0])
        two, st2, s2, t2 = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=cfg, return_seeds=True, return_terms=True)
        pstream, _ = hp.perm_bf(logy, c01, off, B, seed6=hp.gene_seeds(3, B), seed_mode=1, cfg=cfg)
        C = (X > 0).sum(axis=0) / X.shape[0]
        cdr = np.broadcast_to(C, X.shape)[X > 0]
        adj2, _ = hp.perm_bf(logy[:off[20]], c01[:off[20]], off[:21], 34, seed6=seeds[:20], cdr=cdr[:off[20]], cfg=cfg)
    finally:
        hp.set_devices([0])
    assert np.array_equal(one, two) and np.array_equal(s1, s2) and np.array_equal(t1, t2)
    assert st2["fit_sets"] == st1["fit_sets"] == 2 * 96 * B
    pstream1, _ = hp.perm_bf(logy, c01, off, B, seed6=hp.gene_seeds(3, B), seed_mode=1, cfg=cfg)
    assert np.array_equal(pstream, pstream1)
    adj1, _ = hp.perm_bf(logy[:off[20]], c01[:off[20]], off[:21], 34, seed6=seeds[:20], cdr=cdr[:off[20]], cfg=cfg)
    assert np.array_equal(adj1, adj2)
    # observed fits: contiguous gene ranges per device, same results
    Xo, _ = _load(golden_dir, "scDatEx.npz", genes=slice(0, 300))
    lo, co, oo = hp.gene_csr(Xo, cond)
    f1 = hp.fit_observed(lo, co, oo, cfg=cfg)
    try:
        hp.set_devices([0, 0])
        f2 = hp.fit_observed(lo, co, oo, cfg=cfg)
    finally:
        hp.set_devices([0])
    assert f1[0].tobytes() == f2[0].tobytes() and np.array_equal(f1[1], f2[1]) and np.array_equal(f1[2], f2[2])
    assert np.array_equal(f1[3], f2[3]) and np.array_equal(f1[4], f2[4]) and f2[5]["fit_sets"] == 900
    hp.release_cache()                                          # and the library still works afterwards
    again, _ = hp.perm_bf(logy, c01, off, 3, seed6=seeds, cfg=cfg)
    assert np.array_equal(again, one[:, :3])


def test_interrupt_hook_stops_between_chunks(golden_dir):
    X, cond = _load(golden_dir, "scDatEx.npz", genes=slice(0, 50))
    logy, c01, off = hp.gene_csr(X, cond)
    seeds = hp.gene_seeds(1, 50)
    calls = []

    def hook():
        calls.append(1)
        return len(calls) >= 2
    with pytest.raises(hp.ScddError) as ei:
        hp.perm_bf(logy, c01, off, 100, seed6=seeds, cfg=hp.make_cfg(**PRIORS), interrupt=hook)
    assert ei.value.code == -6 and len(calls) == 2              # polled once per chunk of 16 permutations
    ok, _ = hp.perm_bf(logy, c01, off, 20, seed6=seeds, cfg=hp.make_cfg(**PRIORS), interrupt=lambda: False)
    plain, _ = hp.perm_bf(logy, c01, off, 20, seed6=seeds, cfg=hp.make_cfg(**PRIORS))
    assert np.array_equal(ok, plain)                            # chunking by 16 is invisible


@pytest.mark.parametrize("ncells", [40, 700, 5000, 40000, 70000])
def test_permutation_kernel_group_sizes_draw_identical_permutations(ncells, monkeypatch):
    """4, 8 or 32 lanes per gene, candidates consumed by the leader one by one or by the whole group in batches: the
    same R permutations (checked through the Bayes-factor numerators and the final stream states, against the host
    sampler)"""
    rng = np.random.default_rng(ncells)
    ng = 7
    rows = [rng.normal(3, 1, size=max(6, int(ncells * f))) for f in (1.0, 0.9, 0.5, 0.31, 0.77, 0.13, 0.62)]
    off = np.r_[0, np.cumsum([r.size for r in rows])].astype(np.int64)
    logy = np.concatenate(rows)
    c01 = (rng.random(logy.size) < 0.5).astype(np.int32)
    for g in range(ng):
        c01[off[g]:off[g] + 3] = 0
        c01[off[g + 1] - 3:off[g + 1]] = 1
    seeds = hp.gene_seeds(9, ng)
    B = 3
    states = np.array([hp.sample_perms(seeds[g], int(off[g + 1] - off[g]), B)[1] for g in range(ng)])
    outs = []
    # lanes per gene x (leader-serial | whole-group batched) consumption of the candidate stream
    for gs, batch in (("4", "0"), ("8", "0"), ("32", "0"), ("32", "1"), ("8", "1"), ("4", "1")):
        monkeypatch.setenv("SCDD_PERMGEN_GS", gs)
        monkeypatch.setenv("SCDD_PERMGEN_BATCH", batch)
        out, st, so = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=hp.make_cfg(**PRIORS), return_seeds=True)
        assert np.array_equal(so, states), (gs, batch)
        outs.append(out)
    monkeypatch.delenv("SCDD_PERMGEN_BATCH")
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    if ncells <= 5000:
        pi = np.concatenate([hp.sample_perms(seeds[g], int(off[g + 1] - off[g]), B)[0].ravel() for g in range(ng)])
        host, _ = hp.perm_bf(logy, c01, off, B, perm_idx=pi, cfg=hp.make_cfg(**PRIORS))
        assert np.array_equal(outs[0], host)


def test_scdd_categories_on_the_simulated_example(golden_dir):
    """scDD(permutations = 100) on scDatExSim through the host mirror, classifyDD included: the row names carry the
    simulated pattern (DE1..5, DP6..10, DM11..15, DB16..20, EP21..25, EE26..30; R/ex-data.R:76-91), so most DD genes
    must come out with a DD pattern and most EE / EP genes as not significant"""
    from scdd_b200 import api
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    sce = api.SingleCellExperiment({"normcounts": X}, {"condition": cond}, rownames=[str(g) for g in d["genes"]])
    out = api.scDD(sce, prior_param=dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01), permutations=100,
                   testZeroes=False, param={"RNGseed": 816}, verbose=False)
    cats = api.results(out, "Genes")["DDcategory"]
    assert all(c in ("DE", "DP", "DM", "DB", "NC", "NS") for c in cats)
    assert sum(c in ("DE", "DP", "DM", "DB") for c in cats[:20]) >= 12
    assert sum(c == "NS" for c in cats[20:]) >= 7
    # classify=False leaves the genes the reference hands to classifyDD unlabelled
    sce2 = api.SingleCellExperiment({"normcounts": X}, {"condition": cond})
    out2 = api.scDD(sce2, prior_param=dict(alpha=0.01, mu0=0, s0=0.01, a0=0.01, b0=0.01), permutations=100,
                    testZeroes=False, param={"RNGseed": 816}, verbose=False, classify=False)
    cats2 = api.results(out2, "Genes")["DDcategory"]
    assert [c for c in cats2 if c is not None] == [c for c, c2 in zip(cats, cats2) if c2 is not None]
    assert np.array_equal(api.results(out, "Genes")["nonzero.pvalue"], api.results(out2, "Genes")["nonzero.pvalue"])


def test_mid_size_bulk_fit_sets_on_four_warp_teams():
    """bulk work on 2,049..4,096-point fit-sets runs on 4-warp teams, four to a CTA on their own named barriers
    (Team<4>): 60 genes x 10 permutations x 2 halves of ~2,600 points = 1,200 fit-sets (> 8 per SM)"""
    rng = np.random.default_rng(77)
    rows, conds = [], []
    for g in range(60):
        n = 5000 + int(rng.integers(0, 600))
        a = rng.normal(2.0, 0.6, size=int(n * (0.3 + 0.4 * rng.random())))
        b = rng.normal(4.0 + rng.random(), 0.5, size=n - a.size)
        v = np.concatenate([a, b])
        rng.shuffle(v)
        rows.append(v)
        conds.append((rng.random(n) < 0.5).astype(np.int32))
    off = np.r_[0, np.cumsum([r.size for r in rows])].astype(np.int64)
    logy, c01 = np.concatenate(rows), np.concatenate(conds)
    seeds = hp.gene_seeds(4, 60)
    out, st, terms = hp.perm_bf(logy, c01, off, 10, seed6=seeds, cfg=hp.make_cfg(**PRIORS), return_terms=True)
    ref = po.perm_batch_ex(logy, c01, off, 10, seed6=seeds, cfg=po.make_cfg(**PRIORS))
    assert st["fit_sets"] == 1200 and st["failed_sets"] == 0
    np.testing.assert_allclose(terms, ref["jp_terms"], rtol=RTOL)
    np.testing.assert_allclose(out, ref["bf_perm"], rtol=RTOL)


def test_almost_coincident_values_keep_tiny_variances():
    """A cluster of values that almost, but not exactly, coincide (spread 3e-7) far from the rest: mclust keeps valid
    components with variances ~1e-13 (> eps), whose log-terms for the other points reach -5e13 -- beyond the range in
    which the kernel's integer extraction is meaningful; such terms must count as exact zeros (mclust.f: below -708)."""
    rng = np.random.default_rng(3)
    rows, conds = [], []
    for trial in range(6):
        n = 60 + 10 * trial
        v = rng.normal(1.0, 0.8, size=n)
        k = 12 + trial
        v[:k] = 4.0 + rng.normal(0, 3e-7, size=k)
        rows.append(v)
        conds.append((rng.random(n) < 0.5).astype(np.int32))
    off = np.concatenate([[0], np.cumsum([r.size for r in rows])]).astype(np.int64)
    logy, c01 = np.concatenate(rows), np.concatenate(conds)
    fits, cls_oa, cls_c, bf, den, st = hp.fit_observed(logy, c01, off, cfg=hp.make_cfg(**PRIORS), allow_failed=True)
    ofits, ocls_oa, ocls_c, obf, oden = po.genefit_batch(logy, c01, off, cfg=po.make_cfg(**PRIORS))
    for g in range(off.size - 1):
        for w in range(3):
            assert fits[g, w]["G"] == ofits[g][w]["G"], (g, w)
            np.testing.assert_allclose(fits[g, w]["bic"], ofits[g][w]["bic"], rtol=1e-7)
    # the regime is really exercised: most G >= 2 rows of the tables hold a variance below 1e-11
    tiny = sum(int(r["valid"] and np.min(r["sigsq"]) < 1e-11) for v in rows for r in po.fit_table(v, po.make_cfg(**PRIORS))[1:])
    assert tiny >= 12
    assert np.array_equal(cls_oa, ocls_oa) and np.array_equal(cls_c, ocls_c)
    np.testing.assert_allclose(bf, obf, rtol=1e-7)
    np.testing.assert_allclose(den, oden, rtol=1e-7)
    B = 8
    seeds = hp.gene_seeds(11, off.size - 1)
    out, st2 = hp.perm_bf(logy, c01, off, B, seed6=seeds, cfg=hp.make_cfg(**PRIORS), allow_failed=True)
    ref = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(**PRIORS))
    np.testing.assert_allclose(out, ref["bf_perm"], rtol=1e-7, equal_nan=True)
