"""GPU property tests at the benchmark's full size (BASELINE config 3: 20,000 genes x 2,000 cells), where the
CPU oracle would take hours: size-independent identities of the domain, checked bit for bit, plus an
oracle comparison on a stratified sample of the same synthetic workload."""
import numpy as np
import pytest

from oracle import pyoracle as po
from scdd_b200 import hotpath as hp
from scdd_b200.simulate import simulate_set_style

pytestmark = pytest.mark.gpu
PRIORS = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)


@pytest.fixture(scope="module")
def workload():
    X, cond = simulate_set_style(20000, 1000, seed=284)
    logy, c01, off = hp.gene_csr(X, cond)
    return X, cond, logy, c01, off


def test_identity_permutation_reproduces_observed_bf(workload):
    """permMclustGene with the identity permutation is genefit's numerator: same fits -- and the same bits wherever the
    observed fit-set ran on a warp like the permuted one; the ~1 % of observed fit-sets that exceed the EM iteration
    budget are re-run by a CTA team (another summation order) and agree to rounding"""
    X, cond, logy, c01, off = workload
    ng = 4000
    o = off[:ng + 1]
    ident = np.concatenate([np.arange(1, int(o[g + 1] - o[g]) + 1, dtype=np.int32) for g in range(ng)])
    out, st = hp.perm_bf(logy, c01, o, 1, perm_idx=ident, cfg=hp.make_cfg(**PRIORS))
    _, _, _, bf, den, _ = hp.fit_observed(logy, c01, o, cfg=hp.make_cfg(**PRIORS))
    np.testing.assert_allclose(out[:, 0], bf, rtol=1e-11)
    assert (out[:, 0] == bf).mean() > 0.9
    assert np.isfinite(bf).all() and np.isfinite(den).all()


def test_full_size_permutations_deterministic_and_label_symmetric(workload):
    """20,000 genes x 2 device-drawn permutations: identical bits on a re-run; swapping the condition labels only
    swaps c1 and c2, and JP(c1) + JP(c2) is a commutative sum"""
    X, cond, logy, c01, off = workload
    seeds = hp.gene_seeds(284, 20000)
    a, st, sa = hp.perm_bf(logy, c01, off, 2, seed6=seeds, cfg=hp.make_cfg(**PRIORS), return_seeds=True)
    b, _, sb = hp.perm_bf(logy, c01, off, 2, seed6=seeds, cfg=hp.make_cfg(**PRIORS), return_seeds=True)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    assert np.isfinite(a).all() and st["failed_sets"] == 0 and st["iter_cap_hits"] == 0
    assert st["fit_sets"] == 2 * 20000 * 2
    c, _ = hp.perm_bf(logy, 1 - c01, off, 2, seed6=seeds, cfg=hp.make_cfg(**PRIORS))
    assert np.array_equal(a, c)
    # continuing the streams = drawing 4 permutations in one go (chunking is invisible)
    d, _ = hp.perm_bf(logy, c01, off, 2, seed6=sa, cfg=hp.make_cfg(**PRIORS))
    e, _ = hp.perm_bf(logy[:off[500]], c01[:off[500]], off[:501], 4, seed6=seeds[:500], cfg=hp.make_cfg(**PRIORS))
    assert np.array_equal(np.hstack([a[:500], d[:500]]), e)


def test_full_size_ks_properties(workload):
    """KS D depends on the order of the pooled sample only: bit-identical for counts and log counts, and
    symmetric in the two conditions; p in [0, 1]"""
    X, cond, logy, c01, off = workload
    vals = np.exp(logy)
    D1, p1 = hp.ks(vals, c01, off)
    D2, p2 = hp.ks(logy, c01, off)
    D3, p3 = hp.ks(vals, 1 - c01, off)
    assert np.array_equal(D1, D2) and np.array_equal(p1, p2)
    assert np.array_equal(p1, p3) and np.allclose(D1, D3, rtol=0, atol=1e-15)
    assert ((p1 >= 0) & (p1 <= 1)).all() and ((D1 > 0) & (D1 <= 1)).all()
    rows = np.linspace(0, 19999, 300).astype(int)
    for g in rows[:50]:
        s = slice(off[g], off[g + 1])
        d, p = po.ks_test(vals[s][c01[s] == 0], vals[s][c01[s] == 1])
        assert d == D1[g] and p == p1[g]


def test_oracle_sample_of_the_benchmark_workload(workload):
    """stratified sample of the synthetic workload (all six simulated categories) against the oracle:
    identical selected G and labels; BIC / log BF within 1e-8; near ties counted"""
    X, cond, logy, c01, off = workload
    rows = np.linspace(0, 19999, 120).astype(int)
    ly, cc, oo = hp.gene_csr(X[rows], cond)
    fits, cls_oa, cls_c, bf, den, st = hp.fit_observed(ly, cc, oo, cfg=hp.make_cfg(**PRIORS))
    ofits, ocls_oa, ocls_c, obf, oden = po.genefit_batch(ly, cc, oo, cfg=po.make_cfg(**PRIORS))
    mism = sum(int(fits[g, w]["G"] != ofits[g][w]["G"]) for g in range(rows.size) for w in range(3))
    assert mism == 0, f"{mism} fit-sets select a different G (near ties reported by the kernel: {st['bic_near']})"
    assert np.array_equal(cls_oa, ocls_oa) and np.array_equal(cls_c, ocls_c)
    np.testing.assert_allclose(bf, obf, rtol=1e-8)
    np.testing.assert_allclose(den, oden, rtol=1e-8)
    seeds = hp.gene_seeds(5, rows.size)
    out, _ = hp.perm_bf(ly, cc, oo, 2, seed6=seeds, cfg=hp.make_cfg(**PRIORS))
    ref = po.perm_batch(ly, cc, oo, 2, seed6=seeds, cfg=po.make_cfg(**PRIORS))
    np.testing.assert_allclose(out, ref, rtol=1e-8)


def test_plan_gene_costs_account_for_the_measured_chunk():
    """scdd_plan_gene_costs: the per-gene EM work the library orders its launches by sums to the chunk's triples"""
    from scdd_b200.simulate import simulate_set_style
    X, cond = simulate_set_style(400, 100, seed=5)
    logy, c01, off = hp.gene_csr(X, cond)
    plan = hp.Plan(0, logy, c01, off, 4, hp.gene_seeds(3, 400), cfg=hp.make_cfg(alpha=0.01))
    with pytest.raises(hp.ScddError, match="no chunk has been measured"):
        plan.gene_costs()
    plan.run_chunk(0, 4)
    st = plan.stats()
    cost = plan.gene_costs()
    assert cost.shape == (400,) and (cost > 0).all()
    np.testing.assert_allclose(cost.sum(), st["triples"], rtol=1e-12)
    plan.close()
