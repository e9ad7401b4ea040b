"""CPU checks of the host-side mirror (no GPU): CSR marshalling, BH adjustment, argument validation,
workload generator, gene sharding; plus a world_size-2 gloo run of the sharded path against the oracle."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import bench
from scdd_b200 import api, hotpath as hp
from scdd_b200.simulate import simulate_set_style

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gene_csr_matches_reference_preprocessing(golden_dir):
    d = np.load(os.path.join(golden_dir, "scDatExSim.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    logy, c01, off = hp.gene_csr(X, cond)
    for g in (0, 7, 29):
        y = X[g]
        assert logy[off[g]:off[g + 1]].tolist() == [math.log(v) for v in y[y > 0]]    # R/scDD.R:374-375, libm log
        assert np.array_equal(c01[off[g]:off[g + 1]], (cond[y > 0] != cond[0]).astype(np.int32))
    v, c, o = hp.gene_csr(X, cond, log=False, incl_zero=True)
    assert v.size == X.size and o[-1] == X.size


def _zhat_by_loops(X, cond, ref, sel, off, lab, which):
    """R/scDD.R:528-552 spelled out gene by gene."""
    M = np.ones(X.shape)
    M[X == 0] = 0
    for j, g in enumerate(sel):
        M[g][X[g] != 0] = lab[off[j]:off[j + 1]]
    return M if which < 0 else M[:, (cond != ref) == bool(which)]


@pytest.mark.parametrize("order", ["C", "F"])
def test_dense_helpers_match_the_reference_loops(order):
    """scdd_dense_scan / scdd_dense_to_csr / scdd_labels_to_dense / scdd_cluster_counts against the per-gene R
    expressions they replace (R/scDD.R:262, :270-274, :292-298, :374-375, :528-552, R/Cluster_actions.R:25),
    for a numpy (row-major) and an R-style (column-major) matrix."""
    rng = np.random.default_rng(1)
    X = rng.gamma(2.0, 3.0, (300, 57))
    X[rng.random(X.shape) < 0.4] = 0
    X[5, :] = 0                                  # all-zero gene
    X[7, :] = np.where(X[7] > 0, 3.5, 0)         # constant nonzero values in both conditions
    cond = np.array(["b", "a"])[rng.integers(0, 2, 57)]
    ref = cond[0]
    isref = cond == ref
    XX = np.asarray(X, order=order)
    n0, n1, k0, k1, neg = hp.dense_scan(XX, cond, ref=ref)
    assert neg == 0
    assert np.array_equal(n0, (X[:, isref] > 0).sum(1)) and np.array_equal(n1, (X[:, ~isref] > 0).sum(1))

    def const(rows):
        return np.array([len(np.unique(r[r > 0])) == 1 for r in rows])
    assert np.array_equal(k0, const(X[:, isref])) and np.array_equal(k1, const(X[:, ~isref]))
    assert k0[7] and k1[7] and not k0[5]
    Xn = XX.copy(order=order)
    Xn[3, 4] = -1.0
    Xn[200, 0] = -2.0
    assert hp.dense_scan(Xn, cond, ref=ref)[4] == 2

    sel = np.array([1, 2, 3, 10, 200, 299], dtype=np.int32)
    for genes in (None, sel):
        rows = np.arange(300) if genes is None else genes
        v, c, o = hp.gene_csr(XX, cond, ref=ref, genes=genes)
        assert np.array_equal(np.diff(o), (X[rows] > 0).sum(1))
        for j, g in enumerate(rows):
            y = X[g]
            assert v[o[j]:o[j + 1]].tolist() == [math.log(t) for t in y[y > 0]]
            assert np.array_equal(c[o[j]:o[j + 1]], (cond[y > 0] != ref).astype(np.int32))
        vz, cz, oz = hp.gene_csr(XX, cond, ref=ref, genes=genes, log=False, incl_zero=True)
        assert np.array_equal(vz.reshape(len(rows), 57), X[rows]) and np.array_equal(np.diff(oz), np.full(len(rows), 57))
        assert np.array_equal(cz.reshape(len(rows), 57), np.broadcast_to((cond != ref).astype(np.int32), (len(rows), 57)))
        lab = rng.integers(1, 6, v.size).astype(np.int32)
        for which in (-1, 0, 1):
            Z = hp.labels_to_dense(XX, cond, lab, o, genes=genes, which=which, ref=ref)
            assert Z.flags.f_contiguous if order == "F" else Z.flags.c_contiguous
            assert np.array_equal(Z, _zhat_by_loops(X, cond, ref, rows, o, lab, which))
            cnt = hp.cluster_counts(lab, c, o, min_size=3, which=which)
            for j in range(len(rows)):
                l, cc = lab[o[j]:o[j + 1]], c[o[j]:o[j + 1]]
                assert cnt[j] == api.luOutlier(l if which < 0 else l[cc == which], 3)
    with pytest.raises(hp.ScddError, match="gene_off does not match"):
        hp.gene_csr(XX, cond, ref=ref, nnz=np.full(300, 2))
    with pytest.raises(hp.ScddError, match="gene_off does not match"):
        hp.labels_to_dense(XX, cond, np.ones(600, dtype=np.int32), np.arange(0, 602, 2), ref=ref)
    with pytest.raises(ValueError, match="one entry per cell"):
        hp.dense_scan(XX, cond[:-1])
    v, c, o = hp.gene_csr(np.zeros((0, 57)), cond)
    assert v.size == 0 and o.tolist() == [0]


def test_p_adjust_bh():
    p = np.array([0.01, 0.04, 0.03, np.nan, 0.5, 0.02])
    np.testing.assert_allclose(api.p_adjust_BH(p), [0.05, 0.05, 0.05, np.nan, 0.5, 0.05])
    rng = np.random.default_rng(0)
    p = rng.random(200)
    adj = api.p_adjust_BH(p)
    o = np.argsort(p)
    assert np.all(np.diff(adj[o]) >= -1e-15) and np.all(adj >= p) and adj.max() <= 1.0


def test_scdd_argument_validation_mirrors_reference_messages():
    X = np.abs(np.random.default_rng(0).normal(size=(5, 10))) + 1
    good = api.SingleCellExperiment({"normcounts": X}, {"condition": np.repeat([1, 2], 5)})
    with pytest.raises(ValueError, match="Please provide a valid 'SingleCellExperiment' object."):
        api.scDD("x")
    with pytest.raises(ValueError, match="an assays slot named 'normcounts'"):
        api.scDD(api.SingleCellExperiment({"counts": X}, {"condition": np.repeat([1, 2], 5)}), testZeroes=False)
    with pytest.raises(ValueError, match="Categorization will be carried out if permuations are run."):
        api.scDD(good, permutations=10, categorize=False, testZeroes=False)
    with pytest.raises(ValueError, match="Please specify valid condition labels"):
        api.scDD(api.SingleCellExperiment({"normcounts": X}, {"condition": np.repeat([1, 2, 3, 4, 5], 2)}), testZeroes=False)
    neg = api.SingleCellExperiment({"normcounts": -X}, {"condition": np.repeat([1, 2], 5)})
    with pytest.raises(ValueError, match="Negative values"):
        api.scDD(neg, testZeroes=False)
    with pytest.raises(ValueError):
        api.results(good, "nope")


def test_simulated_workload_shape_and_mix():
    X, cond, cats = simulate_set_style(300, 100, seed=284, return_categories=True)
    assert X.shape == (300, 200) and set(cond) == {1, 2}
    assert (X >= 0).all() and 0.02 < (X == 0).mean() < 0.5
    u, c = np.unique(cats, return_counts=True)
    assert abs(dict(zip(u, c))["EE"] - 150) <= 4 and dict(zip(u, c))["EP"] == 120 and c.sum() == 300
    nz = X[X > 0]
    assert np.mean(nz == np.round(nz)) < 0.01          # continuity jitter, R/Simulate_set.R:257-260
    X2, _ = simulate_set_style(300, 100, seed=284)
    assert np.array_equal(X, X2)


def test_shards_partition_the_genes():
    for world in (1, 2, 4, 8):
        rows = [bench.shard_rows(20000, r, world) for r in range(world)]
        allr = np.sort(np.concatenate(rows))
        assert np.array_equal(allr, np.arange(20000))
        assert max(len(r) for r in rows) - min(len(r) for r in rows) <= 1


def test_cost_balanced_shards_partition_and_balance():
    rng = np.random.default_rng(3)
    cost = rng.lognormal(0.0, 1.2, 20000)           # heavy-tailed, like EM work per gene
    for world in (2, 8):
        parts = bench.balance_rows(cost, world)
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(20000))
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.001     # interleaving the same costs is off by several percent
        inter = np.array([cost[bench.shard_rows(20000, r, world)].sum() for r in range(world)])
        assert loads.max() <= inter.max()
        again = bench.balance_rows(cost, world)
        assert all(np.array_equal(a, b) for a, b in zip(parts, again))   # deterministic: every rank computes the same split


_GLOO = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SCDD_ROOT"])
import bench
from oracle import pyoracle as po
from scdd_b200 import hotpath as hp
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = np.load(os.path.join(os.environ["SCDD_ROOT"], "tests", "golden", "scDatExSim.npz"), allow_pickle=True)
X, cond = d["normcounts"][:10], d["condition"]
rows = bench.shard_rows(X.shape[0], rank, world)
logy, c01, off = hp.gene_csr(X[rows], cond)
seeds = hp.gene_seeds(3, X.shape[0])[rows]
B = 2
mine = po.perm_batch(logy, c01, off, B, seed6=seeds, cfg=po.make_cfg(alpha=0.01), nthreads=1)
full = torch.zeros(X.shape[0], B, dtype=torch.float64)
full[torch.from_numpy(rows)] = torch.from_numpy(mine)
dist.all_reduce(full)            # gather on the host: shards are disjoint, so a sum is a gather
units = torch.tensor([float(rows.size * B)], dtype=torch.float64)
dist.all_reduce(units)
if rank == 0:
    la, ca, oa = hp.gene_csr(X, cond)
    ref = po.perm_batch(la, ca, oa, B, seed6=hp.gene_seeds(3, X.shape[0]), cfg=po.make_cfg(alpha=0.01), nthreads=1)
    assert np.array_equal(full.numpy(), ref), "sharded result differs from the single-process result"
    assert units.item() == X.shape[0] * B
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_gene_sharding_gloo(tmp_path):
    """N > 1 path on CPU: two gloo ranks shard the genes exactly as bench.py does (rows + per-gene RNG streams);
    the gathered result must equal the single-process result bit for bit."""
    script = tmp_path / "gloo_shard.py"
    script.write_text(_GLOO)
    env = dict(os.environ, SCDD_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "GLOO_OK" in r.stdout


def test_fisher_combination_and_factor_levels():
    """R/scDD.R:570-578 (Fisher's method with R's NA rules) and the level order of factor(cond) (R/classify.dd.R:268)"""
    from scipy import stats
    for a, b in [(0.3, 0.04), (1e-5, 0.9), (1.0, 1.0), (1e-300, 1e-300), (0.5, 0.5)]:
        want = stats.chi2.sf(-2 * (np.log(a) + np.log(b)), 4)
        np.testing.assert_allclose(api.fishersCombinedPval((a, b)), want, rtol=1e-12, atol=1e-320)
    assert api.fishersCombinedPval((0.0, 0.5)) == 0.0
    assert api.fishersCombinedPval((np.nan, 0.2)) == 0.2 and api.fishersCombinedPval((0.7, np.nan)) == 0.7
    assert np.isnan(api.fishersCombinedPval((np.nan, np.nan)))
    from scdd_b200 import hotpath as hp
    assert hp.factor_level2(np.array([2, 1, 2, 1])).tolist() == [1, 0, 1, 0]            # sorted levels, not first-seen
    assert hp.factor_level2(np.array(["trt", "ctl", "ctl"])).tolist() == [1, 0, 0]
    with pytest.raises(ValueError):
        hp.factor_level2(np.array([1, 1, 1]))
