"""simulateSet-STYLE synthetic input generator for the throughput harness.

This is NOT a port of the reference simulator (R/Simulate_set.R is out of scope): it draws data with the same
distribution families and category mix so that the benchmark exercises the fits on realistic shapes, without R:
  * a base gene is drawn from the bundled scDatEx condition-1 matrix among genes with < 75 % zeros
    (R/findIndex.R:54); its nonzero mean / variance give NB parameters by the method of moments
    (calcRP, R/Single_Cell_Simu.R:21-32); nonzero values are zero-truncated NB draws;
  * per-condition zero fraction = the base gene's empirical zero fraction, nonzeros placed at random
    cells (R/SimuDE.R:66-69);
  * a second mode at FC2 = exp(f * sqrt(log(1 + var/mean^2))), f in modeFC = {2,3,4}
    (R/Single_Cell_Simu.R:177-182);
  * categories: DE (mean shift by 2^(1..3 sd of logFC)), DP (mixing 0.33/0.66 flipped between conditions),
    DM (unimodal vs 50/50 bimodal), DB (50/50 bimodal vs shifted unimodal in the middle),
    EP (50/50 bimodal in both), EE (unimodal in both) -- default mix 250/250/250/250/5000/4000
    (R/Simulate_set.R:119-120), scaled to the requested number of genes;
  * + U(0,1) continuity jitter on nonzeros (R/Simulate_set.R:257-260).
"""
import os

import numpy as np

CATS = ("DE", "DP", "DM", "DB", "EE", "EP")
DEFAULT_MIX = dict(DE=250, DP=250, DM=250, DB=250, EE=5000, EP=4000)


def _calc_rp(m, v):
    r = max(0.01, m * m / (v - m)) if v > m else 0.01
    p = max(0.01, 1.0 - m / v) if v > 0 else 0.01
    if r == 0.01 and p == 0.01:
        r, p = 1.0, 0.5
    return r, min(p, 0.99999999)


def _ztnb(rng, r, p, size):
    """zero-truncated negative binomial: rnbinom(size=r, prob=1-p) redrawn while 0"""
    out = rng.negative_binomial(r, 1.0 - p, size=size)
    z = out == 0
    guard = 0
    while z.any() and guard < 200:
        out[z] = rng.negative_binomial(r, 1.0 - p, size=int(z.sum()))
        z = out == 0
        guard += 1
    out[z] = 1
    return out.astype(np.float64)


def _place(rng, ncell, vals):
    row = np.zeros(ncell)
    pos = rng.choice(ncell, size=vals.size, replace=False)
    row[pos] = vals
    return row


def _base_moments(base):
    zf = (base == 0).mean(axis=1)
    keep = np.nonzero(zf < 0.75)[0]
    mom = []
    for g in keep:
        a = base[g][base[g] != 0]
        if a.size < 5:
            continue
        m, v = a.mean(), a.var(ddof=1)
        if not (v > 0 and m > 0):
            continue
        mom.append((m, v, zf[g]))
    return np.array(mom)


def load_base_matrix():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = np.load(os.path.join(here, "tests", "golden", "scDatEx.npz"), allow_pickle=True)
    X, cond = d["normcounts"], d["condition"]
    return X[:, cond == cond[0]]


LOGFC_SD = 0.7   # spread of the binned observed log fold changes (findFC): DE genes move 1-3 of these


def _gene_row(rng, cat, mom, n_per_cond, mode_fc):
    """one gene of category `cat`: 2 * n_per_cond normalized counts (no continuity jitter yet)"""
    m, v, zf = mom[rng.integers(mom.shape[0])]
    f = mode_fc[rng.integers(len(mode_fc))]
    fc2 = np.exp(f * np.sqrt(np.log(1.0 + v / (m * m))))
    r1, p1 = _calc_rp(m, v)
    r2, p2 = _calc_rp(m * fc2, v * fc2 * fc2)            # second mode: same CV, mean x FC2
    nnz = max(3, int(round(n_per_cond * (1.0 - zf))))

    def draw(w2, shift=1.0):
        """nnz values: fraction w2 from the second mode; `shift` multiplies the first mode's mean"""
        n2 = int(round(nnz * w2))
        ra, pa = (r1, p1) if shift == 1.0 else _calc_rp(m * shift, v * shift * shift)
        a = _ztnb(rng, ra, pa, nnz - n2)
        b = _ztnb(rng, r2, p2, n2) if n2 else np.zeros(0)
        return np.concatenate([a, b])

    if cat == "EE":
        v1, v2 = draw(0.0), draw(0.0)
    elif cat == "DE":
        fc = 2.0 ** (rng.uniform(1.0, 3.0) * LOGFC_SD * (1 if rng.random() < 0.5 else -1))
        v1, v2 = draw(0.0), draw(0.0, shift=fc)
    elif cat == "DP":
        w = (0.33, 0.66) if rng.random() < 0.5 else (0.66, 0.33)
        v1, v2 = draw(w[0]), draw(w[1])
    elif cat == "DM":
        v1, v2 = draw(0.0), draw(0.5)
        if rng.random() < 0.5:
            v1, v2 = v2, v1
    elif cat == "DB":
        v1, v2 = draw(0.5), draw(0.0, shift=np.sqrt(fc2))
        if rng.random() < 0.5:
            v1, v2 = v2, v1
    else:  # EP
        v1, v2 = draw(0.5), draw(0.5)
    row = np.empty(2 * n_per_cond)
    row[:n_per_cond] = _place(rng, n_per_cond, v1)
    row[n_per_cond:] = _place(rng, n_per_cond, v2)
    return row


def _categories(n_genes, mix=None):
    mix = dict(DEFAULT_MIX if mix is None else mix)
    tot = float(sum(mix.values()))
    counts = {c: int(round(n_genes * mix.get(c, 0) / tot)) for c in CATS}
    counts["EE"] += n_genes - sum(counts.values())
    return np.concatenate([[c] * counts[c] for c in CATS])


def simulate_gene_block(genes, n_genes_total=20000, n_per_cond=25000, seed=284, mode_fc=(2, 3, 4), mix=None, base=None):
    """Rows `genes` (indices into an n_genes_total-gene data set) of a simulateSet-style matrix in which EVERY GENE HAS
    ITS OWN RNG STREAM (seed, gene index): a rank of a sharded job can draw just its shard -- BASELINE config 5 is
    20,000 x 50,000 = 8 GB dense, which no rank needs to hold.  Returns (normcounts [len(genes), 2*n_per_cond], condition)."""
    base = load_base_matrix() if base is None else base
    mom = _base_moments(base)
    cats = _categories(n_genes_total, mix)
    genes = np.asarray(genes, dtype=np.int64)
    X = np.empty((genes.size, 2 * n_per_cond))
    for j, g in enumerate(genes):
        rng = np.random.default_rng([int(seed), int(g)])
        row = _gene_row(rng, cats[g], mom, n_per_cond, mode_fc)
        nz = row != 0
        row[nz] += rng.random(int(nz.sum()))
        X[j] = row
    return X, np.repeat([1, 2], n_per_cond)


def simulate_set_style(n_genes=20000, n_per_cond=1000, seed=284, mode_fc=(2, 3, 4), mix=None, base=None,
                       return_categories=False):
    """Returns (normcounts [n_genes, 2*n_per_cond] float64, condition [2*n_per_cond] in {1,2})."""
    rng = np.random.default_rng(seed)
    base = load_base_matrix() if base is None else base
    mom = _base_moments(base)
    cats = _categories(n_genes, mix)
    X = np.zeros((n_genes, 2 * n_per_cond))
    for i, cat in enumerate(cats):
        X[i] = _gene_row(rng, cat, mom, n_per_cond, mode_fc)
    nz = X != 0
    X[nz] += rng.random(int(nz.sum()))
    cond = np.repeat([1, 2], n_per_cond)
    if return_categories:
        return X, cond, cats
    return X, cond
