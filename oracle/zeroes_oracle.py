"""CPU restatement of testZeroes (R/classify.dd.R:262-286) -- TEST INFRASTRUCTURE, never on the product path.

testZeroes fits, per gene with at least one zero,

    arm::bayesglm(y > 0 ~ detection + factor(cond), family = binomial(link = "logit"))

and returns the Wald p-value of the condition coefficient, summary(M1)$coefficients[3, 4].  `arm` is a third-party
dependency that is NOT in /root/reference (DESCRIPTION:17 Imports: arm, no version pin; NAMESPACE:19
importFrom(arm, bayesglm)), and neither this image nor the GPU box has R.  What follows restates the published
algorithm -- Gelman, Jakulin, Pittau & Su (2008), "A weakly informative default prior distribution for logistic and
other regression models", section 3: IRLS on the data augmented with one pseudo-observation per coefficient, the
prior scale of each coefficient re-estimated by an EM step for the t prior -- with the defaults of bayesglm() as the
reference calls it [MEM: arm 1.10 - 1.13, bayesglm.R]:

  prior.mean = 0, prior.df = 1 (Cauchy), prior.scale = 2.5 for the predictors and 10 for the intercept;
  scaled = TRUE: the scale of a predictor is divided by max - min if the predictor takes two values, by 2 sd if more;
  start: mustart = (y + 0.5) / 2 (binomial()$initialize), eta = logit(mustart); dispersion = 1;
  per iteration: z, w of glm.fit; lm.fit on [X; I] with weights [w; 1 / prior.sd];
                 V = chol2inv(R); prior.sd = sqrt(((coef - mean)^2 + diag(V) + df scale^2) / (1 + df));
  stop when iter > 1 and |dev - devold| / (0.1 + |dev|) < epsilon (1e-8); at most maxit (100) iterations;
  summary.glm: se = sqrt(diag(chol2inv(qr of the LAST lm.fit))), z = coef / se, p = 2 pnorm(-|z|).

PARITY UNPINNED against arm itself (no R anywhere in this build): tests/test_zeroes.py pins the restatement to the
published model instead -- run to tight convergence, the result satisfies the paper's two fixed-point equations (the
score of the normal-penalised likelihood vanishes; sigma_j^2 = (beta_j^2 + V_jj + nu s_j^2) / (1 + nu)), both checked
from their definitions -- and tools/dump_golden.R writes the real bayesglm numbers for tests/test_r_golden.py on a
machine that has R.

binomial()'s link functions follow R's C code (src/library/stats/src/family.c): eta is clamped at +-30 in linkinv
and mu.eta, dev.resids = 2 wt (y log(y / mu) + (1 - y) log((1 - y) / (1 - mu))).
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps
THRESH = 30.0


def linkinv(eta):
    """logit_linkinv, family.c"""
    tmp = np.where(eta < -THRESH, EPS, np.where(eta > THRESH, 1.0 / EPS, np.exp(np.clip(eta, -THRESH, THRESH))))
    return tmp / (1.0 + tmp)


def mu_eta(eta):
    """logit_mu_eta, family.c"""
    e = np.exp(np.clip(eta, -THRESH, THRESH))
    opexp = 1.0 + e
    return np.where(np.abs(eta) > THRESH, EPS, e / (opexp * opexp))


def dev_resids(y, mu):
    """binomial_dev_resids, family.c, weights 1, y in {0, 1}"""
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(y != 0, y * np.log(np.where(y != 0, y, 1.0) / mu), 0.0)
        b = np.where(y != 1, (1 - y) * np.log(np.where(y != 1, 1 - y, 1.0) / (1 - mu)), 0.0)
    return 2.0 * (a + b)


def r_sd(x):
    return float(np.std(x, ddof=1)) if len(x) > 1 else float("nan")


def prior_scales(X, prior_scale=2.5, prior_scale_for_intercept=10.0, min_prior_scale=1e-12):
    """bayesglm's scaled = TRUE rule for a design matrix whose first column is the intercept"""
    sc = [prior_scale_for_intercept]
    for j in range(1, X.shape[1]):
        x = X[:, j]
        ncat = len(np.unique(x))
        xs = 1.0
        if ncat == 2:
            xs = float(x.max() - x.min())
        elif ncat > 2:
            xs = 2.0 * r_sd(x)
        sc.append(max(prior_scale / xs, min_prior_scale))
    return np.array(sc)


def bayesglm_logit(y, X, scale=None, df=1.0, mean=0.0, epsilon=1e-8, maxit=100):
    """the loop of arm::bayesglm.fit for family = binomial(logit), weights 1, no offset.
    Returns dict(coef, se, z, p, iter, dev, converged, prior_sd, V)."""
    y = np.asarray(y, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    scale = prior_scales(X) if scale is None else np.asarray(scale, dtype=np.float64)
    mean = np.full(p, float(mean))
    mustart = (y + 0.5) / 2.0
    eta = np.log(mustart / (1.0 - mustart))
    mu = linkinv(eta)
    devold = float(dev_resids(y, mu).sum())
    prior_sd = scale.copy()
    coef = np.zeros(p)
    V = np.eye(p)
    conv = False
    it = 0
    dev = devold
    for it in range(1, maxit + 1):
        me = mu_eta(eta)
        z = eta + (y - mu) / me
        w = np.sqrt(me * me / (mu * (1.0 - mu)))
        xs = np.vstack([X, np.eye(p)])
        zs = np.concatenate([z, mean])
        ws = np.concatenate([w, 1.0 / prior_sd])
        Q, R = np.linalg.qr(xs * ws[:, None])
        coef = np.linalg.solve(R, Q.T @ (zs * ws))
        eta = X @ coef
        mu = linkinv(eta)
        dev = float(dev_resids(y, mu).sum())
        Rinv = np.linalg.inv(R)
        V = Rinv @ Rinv.T                                  # chol2inv(R)
        prior_sd = np.sqrt(((coef - mean) ** 2 + np.diag(V) + df * scale ** 2) / (1.0 + df))
        if it > 1 and abs(dev - devold) / (0.1 + abs(dev)) < epsilon:
            conv = True
            break
        devold = dev
    se = np.sqrt(np.diag(V))
    zval = coef / se
    pval = np.array([math.erfc(abs(t) / math.sqrt(2.0)) for t in zval])     # 2 pnorm(-|z|)
    return dict(coef=coef, se=se, z=zval, p=pval, iter=it, dev=dev, converged=conv, prior_sd=prior_sd, V=V)


def factor_level2(cond):
    """1 where the cell is in the SECOND level of factor(cond) (levels = sort(unique(cond))): the dummy column
    model.matrix builds under treatment contrasts"""
    cond = np.asarray(cond)
    lev = sorted(set(cond.tolist()))
    if len(lev) != 2:
        raise ValueError("two condition levels expected")
    return (cond == lev[1]).astype(np.int32)


def test_zeroes(dat, cond, these=None, epsilon=1e-8, maxit=100, detail=False):
    """R/classify.dd.R:262-286.  dat: genes x cells; returns p-values (NaN = NA) for rows `these` (0-based)."""
    dat = np.asarray(dat, dtype=np.float64)
    detection = (dat > 0).sum(axis=0) / float(dat.shape[0])            # :263
    d2 = factor_level2(cond).astype(np.float64)
    X = np.column_stack([np.ones(dat.shape[1]), detection, d2])
    scale = prior_scales(X)
    these = np.arange(dat.shape[0]) if these is None else np.asarray(these)
    out = np.full(len(these), np.nan)
    det = []
    for j, g in enumerate(these):
        yrow = dat[g]
        if (yrow == 0).sum() > 0:                                      # :267
            r = bayesglm_logit((yrow > 0).astype(np.float64), X, scale=scale, epsilon=epsilon, maxit=maxit)
            out[j] = r["p"][2]                                         # :273 coefficients[3, 4]
            det.append(r)
        else:
            det.append(None)
    return (out, det) if detail else out


test_zeroes.__test__ = False   # not a pytest test
