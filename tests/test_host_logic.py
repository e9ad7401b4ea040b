"""CPU checks of the host-side mirror (no GPU): CSR marshalling, BH adjustment, argument validation,
workload generator, gene sharding; plus a world_size-2 gloo run of the sharded path against the oracle."""
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
        assert np.array_equal(logy[off[g]:off[g + 1]], np.log(y[y > 0]))          # R/scDD.R:374-375
        assert np.array_equal(c01[off[g]:off[g + 1]], (cond[y > 0] != cond[0]).astype(np.int32))
    v, c, o = hp.gene_csr(X, cond, log=False, incl_zero=True)
    assert v.size == X.size and o[-1] == X.size


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
