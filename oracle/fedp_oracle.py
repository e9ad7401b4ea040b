"""CPU restatement of scDD's feDP (R/classify.dd.R:314-383) in numpy/scipy, gene by gene as the R loop runs.

TEST INFRASTRUCTURE -- PARITY UNPINNED: R, stats::t.test / lm / fisher.test are not available in the build image and
the reference ships no test vectors for feDP.  The restatement follows the R sources line by line; the pieces taken
from base R are restated from their documented algorithms:
  * t.test(y ~ factor(cond))          Welch statistic and Satterthwaite df (stats:::t.test.default, var.equal = FALSE)
  * summary(lm(y ~ cdr0 + factor(cond)))$coef[3, 4]   least squares by QR, t = est / se, p on n - 3 df
  * fisher.test(table)                2 x 2: sum of dhyper <= d_obs * (1 + 1e-7); r x 2: FEXACT sums the probabilities
                                      of all tables with the observed margins and log P <= log P_obs + 3.45254e-7
  * p.adjust(p, "BH")
tests/test_oracle_fedp.py pins these pieces against scipy.stats (ttest_ind, fisher_exact, linregress-style OLS).
Only tests/ may import this module.
"""
import math

import numpy as np
from scipy import special

FEXACT_TOL = 3.45254e-7


def lu_outlier(x, min_size=3):
    """R/Cluster_actions.R:25"""
    _, cnt = np.unique(np.asarray(x), return_counts=True)
    return int((cnt >= min_size).sum())


def find_outliers(x, min_size=3):
    """R/Cluster_actions.R:78-81: labels of the clusters with at least min.size members"""
    lab, cnt = np.unique(np.asarray(x), return_counts=True)
    return lab[cnt >= min_size]


def two_sided_t(t, df):
    """2 * pt(-abs(t), df)"""
    return 2.0 * float(special.stdtr(df, -abs(t)))


def _mean_var(x):
    """R's mean() (long double, second pass) and var() (two-pass, long double)"""
    x = np.asarray(x, dtype=np.longdouble)
    n = x.size
    m = x.sum() / n
    m = m + (x - m).sum() / n
    v = ((x - m) ** 2).sum() / (n - 1)
    return float(m), float(v)


def welch_p(y, cond):
    """t.test(y ~ factor(cond))$p.value"""
    lev = np.unique(cond)
    x, z = y[cond == lev[0]], y[cond == lev[1]]
    nx, nz = x.size, z.size
    mx, vx = _mean_var(x)
    mz, vz = _mean_var(z)
    sx, sz = math.sqrt(vx / nx), math.sqrt(vz / nz)
    se = math.sqrt(sx * sx + sz * sz)
    df = se ** 4 / (sx ** 4 / (nx - 1) + sz ** 4 / (nz - 1))
    return two_sided_t((mx - mz) / se, df)


def lm_cond_p(y, cdr0, cond):
    """summary(lm(y ~ cdr0 + factor(cond)))$coef[3, 4]"""
    lev = np.unique(cond)
    X = np.column_stack([np.ones(y.size), cdr0, (cond == lev[1]).astype(float)])
    Q, R = np.linalg.qr(X)
    beta = np.linalg.solve(R, Q.T @ y)
    res = y - X @ beta
    rdf = y.size - 3
    Rinv = np.linalg.inv(R)
    se = math.sqrt(float(res @ res) / rdf * float((Rinv[2] ** 2).sum()))
    return two_sided_t(beta[2] / se, rdf)


def fisher_rx2(tab):
    """fisher.test(tab)$p.value for an r x 2 table of counts (two-sided)"""
    tab = np.asarray(tab, dtype=np.int64)
    r = tab.shape[0]
    rows = tab.sum(axis=1)
    c0, N = int(tab[:, 0].sum()), int(tab.sum())
    lf = special.gammaln(np.arange(N + 2) + 1.0)
    K = lf[rows].sum() + lf[c0] + lf[N - c0] - lf[N]

    def g(i, a):
        return -(lf[a] + lf[rows[i] - a])
    lobs = sum(g(i, tab[i, 0]) for i in range(r))
    tol = math.log1p(1e-7) if r == 2 else FEXACT_TOL
    total = 0.0

    def rec(i, rem, part):
        nonlocal total
        if i == r - 2:
            a = np.arange(max(0, rem - rows[r - 1]), min(rows[r - 2], rem) + 1)
            if a.size == 0:
                return
            lp = part + g(r - 2, a) + g(r - 1, rem - a)
            total += float(np.exp(lp[lp <= lobs + tol] + K).sum())
            return
        for a in range(0, min(rows[i], rem) + 1):
            rec(i + 1, rem - a, part + g(i, a))
    rec(0, c0, 0.0)
    return total


def p_adjust_bh(p):
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pv = p[ok]
    n = pv.size
    if n:
        o = np.argsort(-pv, kind="stable")
        ro = np.argsort(o, kind="stable")
        out[ok] = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * pv[o]))[ro]
    return out


def fedp_gene(row, condition, cdr, oa, c1, c2, adjust_perms=False, min_size=3):
    """One pass of the loop body R/classify.dd.R:337-374 for gene `row` (raw counts, all cells).
    Returns (p_shift, p_fisher or nan, (c.oa, c.c1, c.c2))."""
    condition = np.asarray(condition)
    nz = row > 0
    cond = condition[nz]
    cdr0 = cdr[nz]
    y = np.log(row[nz])
    ref = list(dict.fromkeys(condition.tolist()))[0]      # unique(condition)[1]
    if len(np.unique(y[cond == ref])) == 1 or len(np.unique(y[cond != ref])) == 1:
        return float("nan"), float("nan"), None      # the reference jitters with runif here (RNG dependent)
    p_shift = lm_cond_p(y, cdr0, cond) if adjust_perms else welch_p(y, cond)
    coa, cc1, cc2 = lu_outlier(oa, min_size), lu_outlier(c1, min_size), lu_outlier(c2, min_size)
    p_f = float("nan")
    if cc1 == cc2 and cc1 == coa and coa > 1:
        keep = np.isin(oa, find_outliers(oa, min_size))
        labs = np.unique(oa[keep])
        cols = np.unique(cond[keep])
        if cols.size == 2:
            tab = np.array([[int(((oa == l) & keep & (cond == ref)).sum()), int(((oa == l) & keep & (cond != ref)).sum())]
                            for l in labs])
            p_f = fisher_rx2(tab)
    return p_shift, p_f, (coa, cc1, cc2)


def feDP(pe_mat, condition, sig_genes, oa, c1, c2, testZeroes=False, adjust_perms=False, min_size=3):
    """R/classify.dd.R:314-383.  sig_genes: 0-based indices.  Returns (BH-adjusted p, category names) over the
    non-significant genes in row order -- `names(pval.ns) <- cat`.  R quirk kept: an EMPTY sig_genes makes
    `(1:nrow)[-sig_genes]` empty, so nothing is tested."""
    thresh = 0.025 if testZeroes else 0.05
    n = pe_mat.shape[0]
    sig = set(int(i) for i in sig_genes)
    ns = [] if not sig else [g for g in range(n) if g not in sig]
    cdr = (pe_mat > 0).sum(axis=0) / float(n)
    pv, cat, cnt = [], [], []
    for g in ns:
        ps, pf, c = fedp_gene(pe_mat[g], condition, cdr, oa[g], c1[g], c2[g], adjust_perms, min_size)
        pv.append(ps)
        cnt.append(c if c is not None else (0, 0, 0))
        cat.append("DP" if (not math.isnan(pf) and pf < 0.05) else "NS")
    pv = p_adjust_bh(np.array(pv, dtype=np.float64))
    cat = np.array(cat, dtype=object)
    cnt = np.array(cnt, dtype=np.int64).reshape(-1, 3)
    with np.errstate(invalid="ignore"):
        lowp = pv < thresh
    cat[lowp & (cat != "DP")] = "NC"
    if len(ns):
        cat[lowp & (cnt[:, 1] == cnt[:, 2]) & (cnt[:, 1] == cnt[:, 0]) & (cnt[:, 1] > 1)] = "DP"
    return pv, cat, np.array(ns, dtype=np.int64)
