"""ctypes bindings of the CPU oracle (oracle/libscdd_oracle.so).

TEST INFRASTRUCTURE -- PARITY UNPINNED (see oracle/scdd_oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
GMAX = 5


class Cfg(C.Structure):
    _fields_ = [("alpha", C.c_double), ("m0", C.c_double), ("s0", C.c_double), ("a0", C.c_double),
                ("b0", C.c_double), ("min_size", C.c_int32), ("restrict_", C.c_int32),
                ("qclass_old", C.c_int32), ("extra_mstep", C.c_int32), ("ks_old_pvalue", C.c_int32),
                ("pad_", C.c_int32)]


class Row(C.Structure):
    _fields_ = [("valid", C.c_int32), ("G", C.c_int32), ("iters", C.c_int32), ("status", C.c_int32),
                ("extra_applied", C.c_int32), ("pad_", C.c_int32),
                ("loglik", C.c_double), ("bic", C.c_double),
                ("mean", C.c_double * GMAX), ("sigsq", C.c_double * GMAX), ("pro", C.c_double * GMAX),
                ("mean_em", C.c_double * GMAX), ("sigsq_em", C.c_double * GMAX), ("pro_em", C.c_double * GMAX)]


class Fit(C.Structure):
    _fields_ = [("G", C.c_int32), ("model_v", C.c_int32), ("n", C.c_int32), ("df", C.c_int32),
                ("bic", C.c_double), ("loglik", C.c_double),
                ("mean", C.c_double * GMAX), ("var", C.c_double * GMAX),
                ("n_mclust_calls", C.c_int32), ("pad_", C.c_int32)]


class Rng(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mti", C.c_int32), ("mt", C.c_uint32 * 624), ("lec", C.c_uint32 * 6)]


def build(force=False):
    so = os.path.join(_HERE, "libscdd_oracle.so")
    src = os.path.join(_HERE, "scdd_oracle.c")
    hdr = os.path.join(_HERE, "scdd_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libscdd_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        dp, ip, lp, up = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint32))
        L.orc_default_cfg.argtypes = [C.POINTER(Cfg)]
        L.orc_fit_table.argtypes = [dp, C.c_int32, C.POINTER(Cfg), C.POINTER(Row), ip]
        L.orc_mclust_restricted.argtypes = [dp, C.c_int32, C.POINTER(Cfg), C.POINTER(Fit), ip]
        L.orc_joint_posterior.argtypes = [dp, ip, C.c_int32] + [C.c_double] * 5
        L.orc_joint_posterior.restype = C.c_double
        L.orc_ks.argtypes = [dp, C.c_int32, dp, C.c_int32, C.c_int32, dp, dp]
        L.orc_rng_set_seed.argtypes = [C.POINTER(Rng), C.c_int32, C.c_uint32]
        L.orc_rng_set_lecuyer.argtypes = [C.POINTER(Rng), up]
        L.orc_unif_rand.argtypes = [C.POINTER(Rng)]
        L.orc_unif_rand.restype = C.c_double
        L.orc_unif_index.argtypes = [C.POINTER(Rng), C.c_double]
        L.orc_unif_index.restype = C.c_double
        L.orc_sample_perm.argtypes = [C.POINTER(Rng), C.c_int32, ip]
        L.orc_lecuyer_next_stream.argtypes = [up, up]
        L.orc_lecuyer_next_substream.argtypes = [up, up]
        L.orc_genefit_batch.argtypes = [dp, ip, lp, C.c_int32, C.POINTER(Cfg), C.POINTER(Fit), ip, ip, dp, dp, C.c_int32]
        L.orc_perm_batch.argtypes = [dp, ip, lp, C.c_int32, C.c_int32, ip, up, C.POINTER(Cfg), dp, C.c_int32]
        L.orc_perm_batch_cov.argtypes = [dp, ip, dp, lp, C.c_int32, C.c_int32, ip, up, C.POINTER(Cfg), dp, C.c_int32]
        L.orc_perm_batch_ex.argtypes = [dp, ip, dp, lp, C.c_int32, C.c_int32, ip, up, C.c_int32, C.POINTER(Cfg), dp,
                                        C.c_int32, lp, dp]
        L.orc_mclust_restricted_rng.argtypes = [dp, C.c_int32, C.POINTER(Cfg), C.POINTER(Fit), ip, C.POINTER(Rng), ip]
        L.orc_ks_batch.argtypes = [dp, ip, lp, C.c_int32, C.c_int32, dp, dp, C.c_int32]
        L.orc_last_work.argtypes = [dp, dp]
    return _LIB


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _l(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def make_cfg(alpha=0.10, mu0=0.0, s0=0.01, a0=0.01, b0=0.01, min_size=3, restrict=True,
             qclass_old=False, extra_mstep=True, ks_old_pvalue=False):
    c = Cfg()
    lib().orc_default_cfg(C.byref(c))
    c.alpha, c.m0, c.s0, c.a0, c.b0 = alpha, mu0, s0, a0, b0
    c.min_size = min_size
    c.restrict_ = int(restrict)
    c.qclass_old = int(qclass_old)
    c.extra_mstep = int(extra_mstep)
    c.ks_old_pvalue = int(ks_old_pvalue)
    return c


def _row_dict(r, labels):
    G = r.G
    return dict(valid=bool(r.valid), G=G, iters=r.iters, status=r.status, extra_applied=bool(r.extra_applied),
                loglik=r.loglik, bic=r.bic,
                mean=np.array(r.mean[:G]), sigsq=np.array(r.sigsq[:G]), pro=np.array(r.pro[:G]),
                mean_em=np.array(r.mean_em[:G]), sigsq_em=np.array(r.sigsq_em[:G]), pro_em=np.array(r.pro_em[:G]),
                labels=labels)


def fit_table(y, cfg=None):
    """mclust's BIC table for G = 1..5 on one vector: list of 5 dicts."""
    cfg = cfg or make_cfg()
    y = np.ascontiguousarray(y, dtype=np.float64)
    n = y.size
    rows = (Row * GMAX)()
    labels = np.zeros((GMAX, n), dtype=np.int32)
    lib().orc_fit_table(_d(y), n, C.byref(cfg), rows, _i(labels))
    return [_row_dict(rows[g], labels[g].copy()) for g in range(GMAX)]


def last_work():
    a, b = C.c_double(), C.c_double()
    lib().orc_last_work(C.byref(a), C.byref(b))
    return a.value, b.value


def _fit_dict(f, cls):
    G = f.G
    return {"class": cls, "mean": np.array(f.mean[:G]), "var": np.array(f.var[:G]), "df": f.df, "n": f.n,
            "bic": f.bic, "loglik": f.loglik, "model": "V" if f.model_v else "X", "G": G,
            "n_mclust_calls": f.n_mclust_calls}


def mclust_restricted(y, restrict=True, min_size=3, cfg=None):
    """R/mclust.restricted.R:46 -- returns the same list fields as the reference."""
    cfg = cfg or make_cfg(min_size=min_size, restrict=restrict)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = Fit()
    cls = np.zeros(y.size, dtype=np.int32)
    rc = lib().orc_mclust_restricted(_d(y), y.size, C.byref(cfg), C.byref(f), _i(cls))
    if rc != 0:
        raise RuntimeError("oracle mclustRestricted failed (all BIC NA)")
    return _fit_dict(f, cls)


def joint_posterior(y, mcobj, alpha, m0, s0, a0, b0):
    """R/joint.posterior.R:35"""
    y = np.ascontiguousarray(y, dtype=np.float64)
    cls = np.ascontiguousarray(mcobj["class"] if isinstance(mcobj, dict) else mcobj, dtype=np.int32)
    return lib().orc_joint_posterior(_d(y), _i(cls), y.size, alpha, m0, s0, a0, b0)


def ks_test(x, y, old_pvalue=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    D, p = C.c_double(), C.c_double()
    lib().orc_ks(_d(x), x.size, _d(y), y.size, int(old_pvalue), C.byref(D), C.byref(p))
    return D.value, p.value


class RRng:
    """R's RNG: kind 'mt' (Mersenne-Twister) or 'lecuyer' (L'Ecuyer-CMRG)."""

    def __init__(self, seed=None, kind="mt", state6=None):
        self.r = Rng()
        if state6 is not None:
            s = np.ascontiguousarray(state6, dtype=np.uint32)
            lib().orc_rng_set_lecuyer(C.byref(self.r), _u(s))
        else:
            lib().orc_rng_set_seed(C.byref(self.r), 0 if kind == "mt" else 1, C.c_uint32(seed & 0xFFFFFFFF))

    def unif(self):
        return lib().orc_unif_rand(C.byref(self.r))

    def unif_index(self, n):
        return int(lib().orc_unif_index(C.byref(self.r), float(n)))

    def sample(self, n):
        out = np.zeros(n, dtype=np.int32)
        lib().orc_sample_perm(C.byref(self.r), n, _i(out))
        return out

    def lecuyer_state(self):
        return np.array(self.r.lec[:], dtype=np.uint32)


def lecuyer_next_stream(s6):
    s = np.ascontiguousarray(s6, dtype=np.uint32)
    o = np.zeros(6, dtype=np.uint32)
    lib().orc_lecuyer_next_stream(_u(s), _u(o))
    return o


def lecuyer_next_substream(s6):
    s = np.ascontiguousarray(s6, dtype=np.uint32)
    o = np.zeros(6, dtype=np.uint32)
    lib().orc_lecuyer_next_substream(_u(s), _u(o))
    return o


def genefit_batch(logy, cond, gene_off, cfg=None, nthreads=0, with_bf=True):
    """R/scDD.R:373-391 over a CSR batch. Returns (fits[ngenes][3], cls_oa, cls_c, bf, den)."""
    cfg = cfg or make_cfg()
    logy = np.ascontiguousarray(logy, dtype=np.float64)
    cond = np.ascontiguousarray(cond, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    ng = gene_off.size - 1
    fits = (Fit * (3 * ng))()
    cls_oa = np.zeros(logy.size, dtype=np.int32)
    cls_c = np.zeros(logy.size, dtype=np.int32)
    bf = np.zeros(ng)
    den = np.zeros(ng)
    rc = lib().orc_genefit_batch(_d(logy), _i(cond), _l(gene_off), ng, C.byref(cfg), fits, _i(cls_oa), _i(cls_c),
                                 _d(bf) if with_bf else None, _d(den) if with_bf else None, nthreads)
    if rc != 0:
        raise RuntimeError("oracle genefit failed")
    out = []
    for g in range(ng):
        o, e = gene_off[g], gene_off[g + 1]
        c = cond[o:e]
        coa = cls_oa[o:e]
        cc = cls_c[o:e]
        out.append((_fit_dict(fits[3 * g], coa.copy()), _fit_dict(fits[3 * g + 1], cc[c == 0].copy()),
                    _fit_dict(fits[3 * g + 2], cc[c == 1].copy())))
    return out, cls_oa, cls_c, bf, den


def perm_batch(logy, cond, gene_off, B, perm_idx=None, seed6=None, cfg=None, nthreads=0, cdr=None):
    """R/permutations.R:211-286. perm_idx: 1-based [sum n_g * B]; seed6: [ngenes, 6] L'Ecuyer states.
    cdr given => adjusted branch (:243-267)."""
    cfg = cfg or make_cfg()
    logy = np.ascontiguousarray(logy, dtype=np.float64)
    cond = np.ascontiguousarray(cond, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    ng = gene_off.size - 1
    out = np.zeros((ng, B))
    pi = _i(np.ascontiguousarray(perm_idx, dtype=np.int32)) if perm_idx is not None else None
    sd = None
    if seed6 is not None:
        seed6 = np.ascontiguousarray(seed6, dtype=np.uint32)
        sd = _u(seed6)
    if perm_idx is not None:
        perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int32)
        pi = _i(perm_idx)
    if cdr is None:
        rc = lib().orc_perm_batch(_d(logy), _i(cond), _l(gene_off), ng, B, pi, sd, C.byref(cfg), _d(out), nthreads)
    else:
        cdr = np.ascontiguousarray(cdr, dtype=np.float64)
        rc = lib().orc_perm_batch_cov(_d(logy), _i(cond), _d(cdr), _l(gene_off), ng, B, pi, sd, C.byref(cfg),
                                      _d(out), nthreads)
    if rc != 0:
        raise RuntimeError("oracle permutation fit failed")
    return out


def perm_batch_ex(logy, cond, gene_off, B, perm_idx=None, seed6=None, seed_mode=0, cfg=None, nthreads=0, cdr=None,
                  allow_failed=True):
    """Both permutation drivers with every option (orc_perm_batch_ex): returns dict(bf_perm[ngenes, B],
    jp_terms[ngenes, B, sets] = JP(c1), JP(c2)[, JP(oa)], n_jittered, rc).
    seed_mode 1 = parallelBy "Permutations" streams (seed6 is [B, 6])."""
    cfg = cfg or make_cfg()
    logy = np.ascontiguousarray(logy, dtype=np.float64)
    cond = np.ascontiguousarray(cond, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    ng = gene_off.size - 1
    sets = 3 if cdr is not None else 2
    out = np.zeros((ng, B))
    terms = np.zeros((ng, B, sets))
    pi = sd = cd = None
    if perm_idx is not None:
        perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int32)
        pi = _i(perm_idx)
    if seed6 is not None:
        seed6 = np.ascontiguousarray(seed6, dtype=np.uint32)
        assert seed6.size == 6 * (B if seed_mode else ng)
        sd = _u(seed6)
    if cdr is not None:
        cdr = np.ascontiguousarray(cdr, dtype=np.float64)
        cd = _d(cdr)
    nj = C.c_int64(0)
    rc = lib().orc_perm_batch_ex(_d(logy), _i(cond), cd, _l(gene_off), ng, B, pi, sd, int(seed_mode), C.byref(cfg),
                                 _d(out), nthreads, C.cast(C.byref(nj), C.POINTER(C.c_int64)), _d(terms))
    if rc != 0 and not allow_failed:
        raise RuntimeError("oracle permutation fit failed")
    return dict(bf_perm=out, jp_terms=terms, n_jittered=int(nj.value), rc=rc)


def mclust_restricted_rng(y, rng, min_size=3, cfg=None):
    """mclustRestricted under a live stream (jitter of R/mclust.restricted.R:48-50): returns (fit dict or None, jittered)."""
    cfg = cfg or make_cfg(min_size=min_size)
    y = np.ascontiguousarray(y, dtype=np.float64)
    f = Fit()
    cls = np.zeros(y.size, dtype=np.int32)
    j = C.c_int32(0)
    rc = lib().orc_mclust_restricted_rng(_d(y), y.size, C.byref(cfg), C.byref(f), _i(cls), C.byref(rng.r),
                                         C.cast(C.byref(j), C.POINTER(C.c_int32)))
    return (None if rc != 0 else _fit_dict(f, cls)), bool(j.value)


def ks_batch(vals, cond, gene_off, old_pvalue=False, nthreads=0):
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    cond = np.ascontiguousarray(cond, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    ng = gene_off.size - 1
    D = np.zeros(ng)
    p = np.zeros(ng)
    lib().orc_ks_batch(_d(vals), _i(cond), _l(gene_off), ng, int(old_pvalue), _d(D), _d(p), nthreads)
    return D, p
