"""Independent numpy restatement of the reference path, written R-call-by-R-call.

Purpose: a SECOND implementation from the same specification (the reference R files and the
published mclust / base-R algorithms) with a different code structure from
oracle/scdd_oracle.c: vectorised EM, and `mclustRestricted` transcribed literally as a sequence
of `Mclust(y, G=...)` calls.  tests/test_oracle_*.py require both to agree.  It is test
infrastructure only (slow, pure numpy).
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps
FLMAX = 1.7976931348623157e308
PI2LOG = 1.837877066409345


# ----------------------------------------------------------------------------- base R bits
def r_quantile7(xs, probs):
    n = xs.size
    out = []
    for p in probs:
        index = 1.0 + (n - 1) * p
        lo = math.floor(index)
        hi = math.ceil(index)
        qs = xs[lo - 1]
        if index > lo and xs[hi - 1] != qs:
            h = index - lo
            qs = (1.0 - h) * qs + h * xs[hi - 1]
        out.append(qs)
    return np.array(out)


def r_seq01(n):
    by = 1.0 / (n - 1)
    return np.array([0.0] + [j * by for j in range(1, n - 1)] + [1.0])


def r_unique(v):
    seen, out = set(), []
    for a in v:
        if a not in seen:
            seen.add(a)
            out.append(a)
    return np.array(out)


def r_table(cl):
    labs = sorted(set(int(c) for c in cl))
    return labs, [int(np.sum(cl == l)) for l in labs]


# ----------------------------------------------------------------------------- mclust
def qclass(x, k):
    xs = np.sort(x)
    xl = x.astype(np.longdouble)
    m = xl.sum() / x.size
    m = m + (xl - m).sum() / x.size
    sd = math.sqrt(float(((xl - m) ** 2).sum() / (x.size - 1)))
    eps = sd * math.sqrt(EPS)
    q = np.array([np.nan])
    n = k
    while q.size < k + 1:
        n += 1
        if n > 64:  # same bound as the C oracle / the kernels (R itself would keep going)
            break
        q = r_unique(r_quantile7(xs, r_seq01(n)))
    if q.size > k + 1:
        dq = np.diff(q)
        nr = q.size - k - 1
        drop = np.argsort(dq, kind="stable")[:nr]
        q = np.delete(q, drop)
    q[0] = xs[0] - eps
    q[-1] = xs[-1] + eps
    cl = np.zeros(x.size, dtype=int)
    for i in range(min(k, q.size - 1)):
        cl[(x >= q[i]) & (x < q[i + 1])] = i + 1
    return cl


def em_V(x, G, z, tol=1e-5):
    """me1v restated with numpy row operations (sums are numpy pairwise sums, NOT sequential:
    results agree with the C oracle to rounding, not bit-for-bit)."""
    n = x.size
    rteps = math.sqrt(EPS)
    hold = FLMAX / 2
    it = 0
    while True:
        it += 1
        s = z.sum(axis=0)
        pro = s / n
        if s.min() <= rteps:
            return None, -3, it
        mu = (z * x[:, None]).sum(axis=0) / s
        d2 = (x[:, None] - mu[None, :]) ** 2
        sig = (z * d2).sum(axis=0) / s
        if sig.min() <= EPS:
            return None, -1, it
        t = np.log(pro)[None, :] + (-(PI2LOG + np.log(sig)[None, :] + d2 / sig[None, :]) / 2.0)
        tmax = t.max(axis=1)
        a = t - tmax[:, None]
        e = np.where(a >= -708.0, np.exp(a), 0.0)
        ssum = e.sum(axis=1)
        hood = float((np.log(ssum) + tmax).sum())
        z = e * (1.0 / ssum)[:, None]
        err = abs(hold - hood) / (1.0 + abs(hood))
        hold = hood
        if not (err > tol):
            break
    return dict(z=z, mu=mu, sig=sig, pro=pro, loglik=hood), 0, it


_table_cache = {}


def fit_table(y, extra_mstep=True):
    key = (y.tobytes(), extra_mstep)
    if key in _table_cache:
        return _table_cache[key]
    n = y.size
    rows = {}
    for G in range(1, 6):
        if G > n:
            continue
        if G == 1:
            mu = float(np.sum(y * (1.0 / n)))
            ss = float(np.sum((y - mu) ** 2) / n)
            if ss == 0:
                continue
            ll = -n * (PI2LOG + (1.0 + math.log(ss))) / 2.0
            rows[1] = dict(G=1, loglik=ll, bic=2 * ll - 2 * math.log(n), mean=np.array([mu]), sigsq=np.array([ss]),
                           classification=np.ones(n, dtype=int), iters=0)
            continue
        cl = qclass(y, G)
        z = np.zeros((n, G))
        for k in range(G):
            z[cl == k + 1, k] = 1.0
        fit, code, it = em_V(y, G, z)
        if fit is None:
            continue
        mean, sig, pro = fit["mu"], fit["sig"], fit["pro"]
        if extra_mstep and np.sum((pro - fit["z"].mean(axis=0)) ** 2) > math.sqrt(EPS):
            zz = fit["z"]
            s = zz.sum(axis=0)
            mean = (zz * y[:, None]).sum(axis=0) / s
            sig = (zz * (y[:, None] - mean[None, :]) ** 2).sum(axis=0) / s
            pro = s / n
        rows[G] = dict(G=G, loglik=fit["loglik"], bic=2 * fit["loglik"] - (3 * G - 1) * math.log(n),
                       mean=mean, sigsq=sig, classification=np.argmax(fit["z"], axis=1) + 1, iters=it)
    _table_cache[key] = rows
    return rows


def Mclust(y, Gset):
    """Mclust(y, modelNames='V', G=Gset): best BIC, ties -> smaller G; None if all NA."""
    rows = fit_table(y)
    best = None
    for G in sorted(Gset):
        if G in rows and (best is None or rows[G]["bic"] > best["bic"]):
            best = rows[G]
    return best


# ----------------------------------------------------------------------------- scDD R code
def mclustRestricted(y, restrict=True, min_size=3):
    """Literal transcription of R/mclust.restricted.R:46-253 (NA is float nan; vectors are lists)."""
    nan = float("nan")
    inf = float("inf")

    def isna(v):
        return isinstance(v, float) and math.isnan(v)

    def tab_at(cnts, idx1):  # table(cl)[idx] positional, 1-based; NA when out of range
        return float(cnts[idx1 - 1]) if idx1 - 1 < len(cnts) else nan

    mc = Mclust(y, range(1, 6))
    cl = mc["classification"]
    comps = mc["G"]
    if comps > len(set(cl.tolist())):
        mc = Mclust(y, range(1, comps))
        cl = mc["classification"]
        comps = mc["G"]
    n = y.size
    remThresh = 1 / (1 + math.exp(-n / 60 + 0.4)) + 1.2
    addThresh = 1 / (1 + math.exp(n / 9 - 4.5)) + 3.4

    def bal(nmin, nmax):
        with np.errstate(all="ignore"):
            v = np.float64(nmin) * np.float64(nmax) / (np.float64(nmin) + np.float64(nmax)) ** 2
            return float(2 * np.sqrt(v))

    if restrict:
        if comps > 1:
            sig = mc["sigsq"]
            meandiff = [float(v) for v in np.diff(mc["mean"]) / np.max(np.sqrt(sig))]
            max_cl = int(np.argmax(sig)) + 1
            min_cl = int(np.argmin(sig)) + 1
            vardiff = float(sig[max_cl - 1] / sig[min_cl - 1])
            _, cnts = r_table(cl)
            nmin = tab_at(cnts, min_cl)
            mincat = float(min(cnts))
            tries = 0
            nmax = tab_at(cnts, max_cl)
            balance = bal(nmin, nmax)
            meandiff = [m * balance for m in meandiff]
            if isna(vardiff):
                vardiff = 1.0
            if isna(nmin):
                nmin = inf
            if len(meandiff) == 0 or any(isna(m) for m in meandiff):
                meandiff = [inf]
            err = 0
            while (min(meandiff) < remThresh or vardiff > 20) and mincat > 2 and tries <= 4 and err == 0:
                comps_old = comps
                comps = comps - 1
                mc = Mclust(y, range(1, comps + 1))
                cl = mc["classification"]
                comps = mc["G"]
                labs, cnts = r_table(cl)
                mincat = float(min(cnts))
                tries += 1
                if comps > 1 and comps_old - comps == 1:
                    xl = y.astype(np.longdouble)
                    compmeans = []
                    for l in labs:
                        v = xl[cl == l]
                        m = v.sum() / v.size
                        m = m + (v - m).sum() / v.size
                        compmeans.append(float(m))
                    sig = mc["sigsq"]
                    meandiff = [float(v) for v in np.diff(np.array(compmeans)) / np.max(np.sqrt(sig))]
                    balance = bal(nmin, nmax)
                    meandiff = [m * balance for m in meandiff]
                    max_cl = int(np.argmax(sig)) + 1
                    min_cl = int(np.argmin(sig)) + 1
                    vardiff = float(sig[max_cl - 1] / sig[min_cl - 1])
                    nmin = tab_at(cnts, min_cl)
                else:
                    meandiff = [10.0]
                    vardiff = -inf
                    nmin = 100.0
                if isna(vardiff):
                    vardiff = 1.0
                    err = 1
                if isna(nmin):
                    nmin = inf
                    err = 1
                if len(meandiff) == 0 or any(isna(m) for m in meandiff):
                    meandiff = [inf]
                    err = 1
                if err == 1:
                    comps = comps + 1
                    mc = Mclust(y, range(1, comps + 1))
                    cl = mc["classification"]
                    comps = mc["G"]
        else:
            comps_old = comps
            comps = comps + 1
            mc2 = Mclust(y, [comps])
            cl2 = mc2["classification"] if mc2 is not None else None
            if mc2 is not None and len(set(cl2.tolist())) == comps:
                mc, cl = mc2, cl2
                _, cnts = r_table(cl)
                sig = mc["sigsq"]
                meandiff = [float(v) for v in np.diff(mc["mean"]) / np.max(np.sqrt(sig))]
                max_cl = int(np.argmax(sig)) + 1
                min_cl = int(np.argmin(sig)) + 1
                vardiff = float(sig[max_cl - 1] / sig[min_cl - 1])
                nmin = tab_at(cnts, min_cl)
                mincat = float(min(cnts))
                nmax = tab_at(cnts, max_cl)
                balance = bal(nmin, nmax)
                meandiff = [m * balance for m in meandiff]
                if not ((min(meandiff) > addThresh and vardiff < 10) and mincat > 2):
                    mc = Mclust(y, [comps_old])
                    cl = mc["classification"]
                    comps = mc["G"]
            else:
                mc = Mclust(y, [comps_old])
                cl = mc["classification"]
                comps = mc["G"]
    _, cnts = r_table(cl)
    if comps > 1 and max(cnts) < min_size:
        comps = comps - 1
        mc = Mclust(y, range(1, comps + 1))
        cl = mc["classification"]
        comps = mc["G"]
    cl2 = cl.copy()
    order = np.array([0])
    if comps > 1:
        order = np.argsort(mc["mean"], kind="stable")
        for i in range(1, comps + 1):
            cl2[cl == i] = order[i - 1] + 1
    return {"class": cl2, "mean": mc["mean"][order], "var": mc["sigsq"][order], "df": 3 * comps - 1 if comps > 1 else 2,
            "n": n, "bic": mc["bic"], "loglik": mc["loglik"], "model": "V" if comps > 1 else "X", "G": comps}


def jointPosterior(y, mcobj, alpha, m0, s0, a0, b0):
    """R/joint.posterior.R:35-64"""
    cls = np.asarray(mcobj["class"])
    r = len(set(cls.tolist()))
    yl = y.astype(np.longdouble)
    nk = np.array([np.sum(cls == k) for k in range(1, r + 1)], dtype=float)
    sumyk = np.array([float(yl[cls == k].sum()) for k in range(1, r + 1)])
    sumyk2 = np.array([float((yl[cls == k] ** 2).sum()) for k in range(1, r + 1)])
    sk = s0 + nk
    mk = (s0 * m0 + sumyk) / sk
    ak = a0 + nk
    bk = b0 + sumyk2 + s0 * m0 ** 2 - sk * mk ** 2
    J = nk.sum()
    lg = math.lgamma
    logpost = 0.0
    for k in range(r):
        lgn = lg(nk[k]) if nk[k] > 0 else float("inf")
        logpost = logpost + lgn + lg(ak[k] / 2) - (ak[k] / 2) * math.log(bk[k] / 2) - 0.5 * math.log(sk[k])
    logpost = logpost + r * math.log(alpha) + lg(alpha) - J * lg(a0 / 2) + (J * a0 / 2) * math.log(b0 / 2) + \
        J * 0.5 * math.log(s0) - (J / 2) * math.log(2 * math.pi) - lg(alpha + J)
    return logpost


def ks_D(x, y):
    """D of stats::ks.test two-sample: long-double cumsum of +-1/n in stable order of c(x, y)."""
    nx, ny = x.size, y.size
    w = np.concatenate([x, y])
    o = np.argsort(w, kind="stable")
    steps = np.where(o < nx, np.float64(1.0) / nx, -np.float64(1.0) / ny).astype(np.longdouble)
    z = np.cumsum(steps).astype(np.float64)
    ws = w[o]
    keep = np.concatenate([np.nonzero(np.diff(ws) != 0)[0], [nx + ny - 1]])
    return float(np.max(np.abs(z[keep])))
