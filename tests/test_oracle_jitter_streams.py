"""CPU checks of the oracle's RNG-coupled pieces added in round 2:

  * the constant-vector jitter of mclustRestricted (R/mclust.restricted.R:48-50) drawn mid-stream,
  * the two stream shapes of scDD()'s permutation dispatch (parallelBy = "Genes" / "Permutations",
    R/scDD.R:414-463, R/permutations.R:99,177),
  * a known answer for set.seed(kind = "L'Ecuyer-CMRG").
"""
import numpy as np

from oracle import pyoracle as po

PRI = dict(alpha=0.01, mu0=0.0, s0=0.01, a0=0.01, b0=0.01)


def _seeds(seed, n):
    """state of bplapply element i for RNGseed = seed: set.seed(seed, "L'Ecuyer-CMRG") + i substream jumps"""
    s = po.RRng(seed, kind="lecuyer").lecuyer_state()
    out = []
    for _ in range(n):
        out.append(s.copy())
        s = po.lecuyer_next_substream(s)
    return np.array(out, dtype=np.uint32)


def tied_gene(rng, n1=4, n2=5, heavy=1.0):
    """log-counts with massive ties: most cells share one value, so that permuted halves are often constant"""
    n = n1 + n2
    y = np.full(n, np.log(heavy + 1.0))
    k = max(2, n // 4)
    y[rng.choice(n, k, replace=False)] = np.log(rng.integers(3, 9, size=k).astype(float))
    c = np.r_[np.zeros(n1, dtype=np.int32), np.ones(n2, dtype=np.int32)]
    return y, c


def test_lecuyer_set_seed_known_answer():
    # R: RNGkind("L'Ecuyer-CMRG"); set.seed(123); .Random.seed[2:6]  (the state printed in the `parallel` package's
    # RNG-streams documentation; recalled from memory [MEM] -- the first five words are reproduced exactly, the sixth
    # is not relied on)
    s = po.RRng(123, kind="lecuyer").lecuyer_state().view(np.int32)
    assert s[:5].tolist() == [1806547166, -983674937, 643431772, 1162448557, -959247990]


def test_jitter_is_runif_from_the_live_stream():
    seeds = _seeds(7, 1)
    y = np.full(6, 1.25)
    r1 = po.RRng(state6=seeds[0])
    fit, jit = po.mclust_restricted_rng(y, r1, cfg=po.make_cfg(**PRI))
    assert jit and fit is not None
    # the same six uniforms, by hand: y + (a + (b - a) * u), a = -0.1, b = 0.1 (nmath/runif.c)
    r2 = po.RRng(state6=seeds[0])
    u = np.array([r2.unif() for _ in range(6)])
    yj = y + (-0.1 + (0.1 - (-0.1)) * u)
    ref = po.mclust_restricted(yj, cfg=po.make_cfg(**PRI))
    assert ref["G"] == fit["G"] and np.array_equal(ref["class"], fit["class"])
    assert ref["bic"] == fit["bic"]
    assert np.array_equal(r1.lecuyer_state(), r2.lecuyer_state())        # exactly n draws were consumed
    # a non-constant vector consumes nothing
    r3 = po.RRng(state6=seeds[0])
    _, jit3 = po.mclust_restricted_rng(np.array([1.0, 2.0, 2.5, 4.0, 4.5, 7.0]), r3, cfg=po.make_cfg(**PRI))
    assert not jit3 and np.array_equal(r3.lecuyer_state(), seeds[0])


def _replay_gene(y, c, state6, B, cfg):
    """permMclustGene by hand on the Python side of the oracle: sample(), then c1, c2 each jittering from the SAME
    stream when constant; returns (bf_perm, number of jittered sets)"""
    r = po.RRng(state6=state6)
    out, nj = [], 0
    for _ in range(B):
        new = r.sample(y.size)
        ny = y[new - 1]
        y1, y2 = ny[c == 0], ny[c == 1]
        f1, j1 = po.mclust_restricted_rng(y1, r, cfg=cfg)
        f2, j2 = po.mclust_restricted_rng(y2, r, cfg=cfg)
        nj += int(j1) + int(j2)
        out.append(po.joint_posterior(y1, f1, PRI["alpha"], PRI["mu0"], PRI["s0"], PRI["a0"], PRI["b0"]) +
                   po.joint_posterior(y2, f2, PRI["alpha"], PRI["mu0"], PRI["s0"], PRI["a0"], PRI["b0"]))
    return np.array(out), nj


def test_perm_batch_jitters_mid_stream_and_stays_in_step():
    rng = np.random.default_rng(5)
    ys, cs = zip(*[tied_gene(rng, 3 + g % 3, 4 + g % 2) for g in range(12)])
    off = np.r_[0, np.cumsum([y.size for y in ys])].astype(np.int64)
    logy, c01 = np.concatenate(ys), np.concatenate(cs)
    B = 12
    seeds = _seeds(11, len(ys))
    cfg = po.make_cfg(**PRI)
    res = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, cfg=cfg)
    assert res["rc"] == 0 and np.isfinite(res["bf_perm"]).all()
    assert res["n_jittered"] > 10                                        # the case is really exercised
    tot = 0
    for g in range(len(ys)):
        bf, nj = _replay_gene(ys[g], cs[g], seeds[g], B, cfg)
        assert np.array_equal(bf, res["bf_perm"][g])                     # later permutations did not desynchronise
        tot += nj
    assert tot == res["n_jittered"]
    # the terms add up to the statistic
    assert np.array_equal(res["jp_terms"][:, :, 0] + res["jp_terms"][:, :, 1], res["bf_perm"])
    # without a stream (host-drawn permutations) a constant subset has no valid model
    pi = np.concatenate([np.concatenate([po.RRng(state6=seeds[g]).sample(ys[g].size) for _ in range(1)]) for g in range(len(ys))])
    r0 = po.perm_batch_ex(logy, c01, off, 1, perm_idx=pi, cfg=cfg)
    first_const = [(np.unique(ys[g][po.RRng(state6=seeds[g]).sample(ys[g].size) - 1][cs[g] == 0]).size == 1) or
                   (np.unique(ys[g][po.RRng(state6=seeds[g]).sample(ys[g].size) - 1][cs[g] == 1]).size == 1)
                   for g in range(len(ys))]
    assert np.array_equal(np.isnan(r0["bf_perm"][:, 0]), np.array(first_const))


def test_permutation_streams_mode():
    """parallelBy = "Permutations": permutation b of every gene is sample() from the START of stream b"""
    d = np.load(__file__.rsplit("/", 1)[0] + "/golden/scDatExSim.npz", allow_pickle=True)
    from scdd_b200 import hotpath as hp            # host-side marshalling only (no GPU needed)
    X, cond = d["normcounts"][:5], d["condition"]
    logy, c01, off = hp.gene_csr(X, cond)
    B = 4
    seeds = _seeds(3, B)
    cfg = po.make_cfg(**PRI)
    res = po.perm_batch_ex(logy, c01, off, B, seed6=seeds, seed_mode=1, cfg=cfg)
    pi = []
    for g in range(5):
        n = int(off[g + 1] - off[g])
        for b in range(B):
            pi.append(po.RRng(state6=seeds[b]).sample(n))
    ref = po.perm_batch(logy, c01, off, B, perm_idx=np.concatenate(pi), cfg=cfg)
    assert np.array_equal(res["bf_perm"], ref)
    # and it differs from the per-gene stream shape
    res0 = po.perm_batch_ex(logy, c01, off, B, seed6=_seeds(3, 5), seed_mode=0, cfg=cfg)
    assert not np.array_equal(res0["bf_perm"], res["bf_perm"])
