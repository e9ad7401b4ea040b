"""CPU restatement of classifyDD (R/classify.dd.R:86-238) -- TEST INFRASTRUCTURE, never on the product path.

Per gene: label compression (:124-158), luOutlier / findOutliers (R/Cluster_actions.R:25, :78), getPosteriorParams
(:19-37), the Monte-Carlo comparison of cluster means (:165-187) and the decision tree (:189-234).  The reference draws
its 10,000 Student-t variates per cluster from R's own RNG stream (`rt`); this restatement draws them with numpy's
generator -- the same distribution, another stream -- so only comparisons that are not at the Monte-Carlo threshold
are comparable between two implementations (or between two R sessions).  tests/test_gpu_classify.py therefore checks
the deterministic parts exactly (posterior parameters, comparison lists) and the flags wherever several oracle seeds
agree with each other.
"""
import numpy as np


def lu_outlier(x, min_size=3):
    """luOutlier, R/Cluster_actions.R:25"""
    _, cnt = np.unique(np.asarray(x), return_counts=True)
    return int((cnt >= min_size).sum())


def compress_labels(cls):
    """R/classify.dd.R:124-158: when the label set has a gap (length(unique) != max) the labels above the FIRST missing
    one move down by one (the reference handles one gap; so does this)."""
    cls = np.asarray(cls).copy()
    mx = int(cls.max())
    present = np.unique(cls)
    if present.size != mx:
        missing = int(np.setdiff1d(np.arange(1, mx + 1), present)[0])
        cls[cls > missing] -= 1
    return cls


def posterior_params(y, cls, m0, s0, a0, b0):
    """getPosteriorParams, R/classify.dd.R:19-37: per cluster k = 1..r the normal-gamma posterior (mk, sk, ak, bk)"""
    r = np.unique(cls).size
    nk = np.array([(cls == k).sum() for k in range(1, r + 1)], dtype=np.float64)
    sy = np.array([y[cls == k].sum() for k in range(1, r + 1)])
    syy = np.array([(y[cls == k] ** 2).sum() for k in range(1, r + 1)])
    sk = s0 + nk
    mk = (s0 * m0 + sy) / sk
    ak = a0 + nk
    bk = b0 + syy + s0 * m0 ** 2 - sk * mk ** 2
    return mk, sk, ak, bk


def find_outliers(cls, min_size):
    """findOutliers, R/Cluster_actions.R:78: labels of the clusters with at least min.size members"""
    lab, cnt = np.unique(cls, return_counts=True)
    return lab[cnt >= min_size].astype(int)


def classify_dd(logy, cond01, gene_off, sig_genes, cls_oa, cls_c, prior_param, min_size=3, rng=None, p_shift=None,
                ndraws=10000, detail=False):
    """classifyDD over the CSR gene block; sig_genes 0-based; p_shift = unadjusted mean-shift p-values per gene
    (t.test / lm coefficient, :206-214).  Returns the category names, or (names, comparisons) with one row
    (df1, scale1, m1, df2, scale2, m2, flag, min, max) per comparison, genes in order."""
    rng = np.random.default_rng(0) if rng is None else rng
    m0, s0, a0, b0 = prior_param["mu0"], prior_param["s0"], prior_param["a0"], prior_param["b0"]
    out, comps = [], []
    for g in np.asarray(sig_genes, dtype=np.int64):
        o, e = int(gene_off[g]), int(gene_off[g + 1])
        y, c = logy[o:e], cond01[o:e]
        oa = compress_labels(cls_oa[o:e])
        c1 = compress_labels(cls_c[o:e][c == 0])
        c2 = compress_labels(cls_c[o:e][c == 1])
        n_oa, n_c1, n_c2 = lu_outlier(oa, min_size), lu_outlier(c1, min_size), lu_outlier(c2, min_size)
        mk1, sk1, ak1, bk1 = posterior_params(y[c == 0], c1, m0, s0, a0, b0)
        mk2, sk2, ak2, bk2 = posterior_params(y[c == 1], c2, m0, s0, a0, b0)
        comparisons = []
        for a in find_outliers(c1, min_size):
            for b in find_outliers(c2, min_size):
                sc1 = np.sqrt(bk1[a - 1] / (ak1[a - 1] * sk1[a - 1]))
                sc2 = np.sqrt(bk2[b - 1] / (ak2[b - 1] * sk2[b - 1]))
                d = (rng.standard_t(round(ak1[a - 1]), ndraws) * sc1 + mk1[a - 1]
                     - rng.standard_t(round(ak2[b - 1]), ndraws) * sc2 - mk2[b - 1])
                flag = 1 if (d.min() > 0 or d.max() < 0) else 0
                comparisons.append(flag)
                comps.append([round(ak1[a - 1]), sc1, mk1[a - 1], round(ak2[b - 1]), sc2, mk2[b - 1], flag, d.min(), d.max()])
        tot = sum(comparisons)
        if n_c1 == n_c2:
            if n_c1 == 1:
                cat = "DE" if tot > 0 else "NC"
            elif n_oa == n_c1:
                cat = "DP" if tot <= n_c1 else "NC"
            elif n_oa > n_c1:
                cat = ("DE" if p_shift[g] < 0.01 else "NC") if tot > n_c1 * (n_c1 - 1) else "NC"
            else:
                cat = "NC"
        else:
            cat = "DB" if tot == len(comparisons) else "DM"
        out.append(cat)
    names = np.array(out, dtype=object)
    return (names, np.array(comps, dtype=np.float64).reshape(-1, 9)) if detail else names
