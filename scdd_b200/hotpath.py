"""numpy-level wrappers of the C ABI (include/scdd_b200.h): marshalling only, no arithmetic of the hot path.

Function names mirror the reference's internal interface for this path:
  fit_observed  <-> genefit (R/scDD.R:327-340, :373-391): mclustRestricted x3 + jointPosterior
  perm_bf       <-> permMclustGene / permMclust / permMclustCov (R/permutations.R)
  ks            <-> testKS::onegene (R/Test_KS.R:56-65)
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FIT_DTYPE, Stats, make_cfg, ScddError  # noqa: F401


def _dense(normcounts, condition, ref):
    """(X, col_major, cell_cond01): the matrix as the C ABI takes it -- C order, or Fortran order (an R matrix) as is."""
    X = np.asarray(normcounts, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("normcounts must be a genes x cells matrix")
    col_major = X.flags.f_contiguous and not X.flags.c_contiguous
    if not col_major:
        X = np.ascontiguousarray(X)
    condition = np.asarray(condition)
    if condition.shape[0] != X.shape[1]:
        raise ValueError("condition must have one entry per cell")
    if ref is None:
        ref = condition[0]
    return X, int(col_major), np.ascontiguousarray(condition != ref, dtype=np.int32)


def dense_scan(normcounts, condition, ref=None):
    """Per-gene filters of scDD() (R/scDD.R:262, :270-274, :292-298) in one threaded pass over the matrix:
    returns (nnz0, nnz1, const0, const1, n_negative)."""
    X, cm, cc = _dense(normcounts, condition, ref)
    ng, nc = X.shape
    nnz0, nnz1 = np.zeros(ng, dtype=np.int32), np.zeros(ng, dtype=np.int32)
    k0, k1 = np.zeros(ng, dtype=np.uint8), np.zeros(ng, dtype=np.uint8)
    neg = C.c_int64(0)
    _lib.check(_lib.lib().scdd_dense_scan(X.ctypes.data, ng, nc, cm, cc.ctypes.data, nnz0.ctypes.data, nnz1.ctypes.data,
                                          k0.ctypes.data, k1.ctypes.data, C.addressof(neg)))
    return nnz0, nnz1, k0.astype(bool), k1.astype(bool), int(neg.value)


def gene_csr(normcounts, condition, ref=None, log=True, incl_zero=False, genes=None, nnz=None):
    """Dense genes x cells matrix -> the CSR-over-genes layout of the C ABI (scdd_dense_to_csr).

    Returns (vals, cond01, gene_off): per gene the (log) nonzero values in cell order
    (`y <- log(y[y>0])`, R/scDD.R:374-375) and 0/1 condition labels (0 = reference = first level,
    R/scDD.R:258).  incl_zero keeps zeros (testKS(inclZero=TRUE), values not logged).
    genes: optional row subset (0-based); nnz: their nonzero counts if already known from dense_scan."""
    X, cm, cc = _dense(normcounts, condition, ref)
    ng, nc = X.shape
    if genes is not None:
        genes = np.ascontiguousarray(genes, dtype=np.int32)
    nsel = ng if genes is None else genes.size
    if incl_zero:
        counts = np.full(nsel, nc, dtype=np.int64)
    elif nnz is not None:
        counts = np.asarray(nnz, dtype=np.int64)
    else:
        n0, n1 = dense_scan(X, cc, ref=0)[:2]
        counts = n0.astype(np.int64) + n1
        if genes is not None:
            counts = counts[genes]
    gene_off = np.zeros(nsel + 1, dtype=np.int64)
    np.cumsum(counts, out=gene_off[1:])
    vals = np.empty(int(gene_off[-1]), dtype=np.float64)
    cond01 = np.empty(int(gene_off[-1]), dtype=np.int32)
    _lib.check(_lib.lib().scdd_dense_to_csr(X.ctypes.data, ng, nc, cm, cc.ctypes.data, _lib.ptr(genes), nsel,
                                            int(bool(log)), int(bool(incl_zero)), gene_off.ctypes.data,
                                            vals.ctypes.data, cond01.ctypes.data))
    return vals, cond01, gene_off


def cell_covariate(normcounts, cell_values, gene_off, genes=None):
    """A per-cell covariate (the detection rate C of R/scDD.R:457) on the nonzero cells of the selected genes, in CSR
    order: the `cdr` argument of perm_bf / fedp (`C[y>0]` per gene, R/permutations.R:216)."""
    X = np.asarray(normcounts, dtype=np.float64)
    col_major = X.flags.f_contiguous and not X.flags.c_contiguous
    if not col_major:
        X = np.ascontiguousarray(X)
    ng, nc = X.shape
    cv = np.ascontiguousarray(cell_values, dtype=np.float64)
    assert cv.size == nc
    if genes is not None:
        genes = np.ascontiguousarray(genes, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    out = np.empty(int(gene_off[-1]), dtype=np.float64)
    _lib.check(_lib.lib().scdd_dense_cell_covariate(X.ctypes.data, ng, nc, int(col_major), cv.ctypes.data, _lib.ptr(genes),
                                                    gene_off.size - 1, gene_off.ctypes.data, out.ctypes.data))
    return out


def factor_level2(condition):
    """1 where the cell is in the SECOND level of factor(cond) (levels = sort(unique(cond))): the dummy column of
    model.matrix(~ factor(cond)) under treatment contrasts, R/classify.dd.R:268"""
    cond = np.asarray(condition)
    lev = sorted(set(cond.tolist()))
    if len(lev) != 2:
        raise ValueError("Error: Please specify valid condition labels.")
    return np.ascontiguousarray(cond == lev[1], dtype=np.int32)


def test_zeroes(normcounts, condition, these=None, epsilon=1e-8, maxit=100, detail=False):
    """testZeroes, R/classify.dd.R:262-286: the Wald p-value of the condition coefficient of
    arm::bayesglm(y > 0 ~ detection + factor(cond), binomial) for the rows `these` (0-based; default all), NaN for a
    gene without zeros.  Returns (p, stats) or (p, detail[n, 8], stats); detail = coef[3], se[3], iterations, deviance."""
    X = np.asarray(normcounts, dtype=np.float64)
    col_major = X.flags.f_contiguous and not X.flags.c_contiguous
    if not col_major:
        X = np.ascontiguousarray(X)
    ng, nc = X.shape
    l2 = factor_level2(condition)
    assert l2.size == nc
    if these is not None:
        these = np.ascontiguousarray(these, dtype=np.int32)
    n = ng if these is None else these.size
    o = _lib.GlmOpts()
    _lib.lib().scdd_default_glm_opts(C.byref(o))
    o.epsilon, o.maxit = float(epsilon), int(maxit)
    p = np.empty(n, dtype=np.float64)
    det = np.empty((n, 8), dtype=np.float64) if detail else None
    st = Stats()
    _lib.check(_lib.lib().scdd_test_zeroes(X.ctypes.data, ng, nc, int(col_major), l2.ctypes.data, _lib.ptr(these), n,
                                           C.byref(o), p.ctypes.data, _lib.ptr(det), C.byref(st)))
    return (p, det, st.as_dict()) if detail else (p, st.as_dict())


test_zeroes.__test__ = False


def labels_to_dense(normcounts, condition, labels, gene_off, genes=None, which=-1, ref=None):
    """Zhat matrix (R/scDD.R:528-552): 0 where the count is zero, the class label on the nonzero cells of the fitted
    genes, 1 elsewhere.  which = 0 / 1 keeps only that condition's columns."""
    X, cm, cc = _dense(normcounts, condition, ref)
    ng, nc = X.shape
    ncols = nc if which < 0 else int((cc != 0).sum() if which else (cc == 0).sum())
    Z = np.empty((ng, ncols), dtype=np.float64, order="F" if cm else "C")
    if genes is not None:
        genes = np.ascontiguousarray(genes, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    nsel = gene_off.size - 1
    assert genes is None or genes.size == nsel
    _lib.check(_lib.lib().scdd_labels_to_dense(X.ctypes.data, ng, nc, cm, cc.ctypes.data, int(which), _lib.ptr(genes),
                                               nsel, gene_off.ctypes.data, labels.ctypes.data, Z.ctypes.data))
    return Z


def cluster_counts(labels, cond01, gene_off, min_size=3, which=-1):
    """luOutlier per gene (R/Cluster_actions.R:25): number of classes with at least min_size members."""
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    cond01 = np.ascontiguousarray(cond01, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    out = np.zeros(gene_off.size - 1, dtype=np.int32)
    _lib.check(_lib.lib().scdd_cluster_counts(labels.ctypes.data, cond01.ctypes.data, gene_off.ctypes.data,
                                              gene_off.size - 1, int(min_size), int(which), out.ctypes.data))
    return out


def _prep(vals, cond01, gene_off):
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    cond01 = np.ascontiguousarray(cond01, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    assert gene_off.ndim == 1 and gene_off.size >= 1
    assert vals.size >= gene_off[-1] and cond01.size >= gene_off[-1]
    return vals, cond01, gene_off


def set_devices(devices):
    d = np.ascontiguousarray(devices, dtype=np.int32)
    _lib.check(_lib.lib().scdd_set_devices(d.ctypes.data_as(C.POINTER(C.c_int32)), d.size))


def device_count():
    return _lib.lib().scdd_device_count()


def fit_observed(logy, cond01, gene_off, cfg=None, with_bf=True, allow_failed=False):
    """Returns (fits[ngenes, 3] structured (oa, c1, c2), cls_oa[nnz], cls_c[nnz], bf, den, stats)."""
    cfg = cfg or make_cfg()
    logy, cond01, gene_off = _prep(logy, cond01, gene_off)
    ng = gene_off.size - 1
    fits = np.zeros((ng, 3), dtype=FIT_DTYPE)
    cls_oa = np.zeros(logy.size, dtype=np.int32)
    cls_c = np.zeros(logy.size, dtype=np.int32)
    bf = np.zeros(ng) if with_bf else None
    den = np.zeros(ng) if with_bf else None
    st = Stats()
    rc = _lib.lib().scdd_fit_observed(logy.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data, ng, C.byref(cfg),
                                      fits.ctypes.data, cls_oa.ctypes.data, cls_c.ctypes.data,
                                      _lib.ptr(bf), _lib.ptr(den), C.byref(st))
    _lib.check(rc, allow=(_lib.ERR_FIT_FAILED,) if allow_failed else ())
    return fits, cls_oa, cls_c, bf, den, st.as_dict()


def perm_bf(logy, cond01, gene_off, B, perm_idx=None, seed6=None, cdr=None, cfg=None, allow_failed=False,
            return_seeds=False, seed_mode=0, return_terms=False, interrupt=None):
    """Bayes-factor numerators of B permutations per gene: array [ngenes, B] (+ stats dict).

    seed_mode 0: seed6 [ngenes, 6], one stream per gene (parallelBy "Genes"); 1: seed6 [B, 6], permutation b of every
    gene from stream b (parallelBy "Permutations").  return_terms adds the jointPosterior terms [ngenes, B, sets]
    (JP(c1), JP(c2)[, JP(oa)]) as the last return value.  interrupt: a callable polled between chunks (True = stop)."""
    cfg = cfg or make_cfg()
    logy, cond01, gene_off = _prep(logy, cond01, gene_off)
    ng = gene_off.size - 1
    out = np.zeros((ng, B))
    if perm_idx is not None:
        perm_idx = np.ascontiguousarray(perm_idx, dtype=np.int32)
        assert perm_idx.size == gene_off[-1] * B
    if seed6 is not None:
        seed6 = np.ascontiguousarray(seed6, dtype=np.uint32).reshape(B if seed_mode else ng, 6)
    if cdr is not None:
        cdr = np.ascontiguousarray(cdr, dtype=np.float64)
        assert cdr.size == logy.size
    seeds_out = np.zeros((ng, 6), dtype=np.uint32) if (return_seeds and seed6 is not None) else None
    terms = np.zeros((ng, B, 3 if cdr is not None else 2)) if return_terms else None
    st = Stats()
    opts = _lib.PermOpts()
    opts.seed_mode = int(seed_mode)
    opts.jp_terms = _lib.ptr(terms)
    hook = None
    if interrupt is not None:
        hook = _lib.INTERRUPT_FN(lambda _arg: 1 if interrupt() else 0)
        opts.interrupt = hook
    rc = _lib.lib().scdd_perm_bf_ex(logy.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data, ng, int(B),
                                    _lib.ptr(perm_idx), _lib.ptr(seed6), _lib.ptr(cdr), C.byref(cfg), C.byref(opts),
                                    out.ctypes.data, _lib.ptr(seeds_out), C.byref(st))
    _lib.check(rc, allow=(_lib.ERR_FIT_FAILED,) if allow_failed else ())
    ret = [out, st.as_dict()]
    if return_seeds:
        ret.append(seeds_out)
    if return_terms:
        ret.append(terms)
    return tuple(ret)


def release_cache():
    """frees the per-device buffers the library keeps between perm_bf calls (scdd_release_cache)"""
    _lib.lib().scdd_release_cache()


def ks(vals, cond01, gene_off, old_pvalue=False):
    """Two-sample KS per gene: (D, p_unadjusted)."""
    vals, cond01, gene_off = _prep(vals, cond01, gene_off)
    ng = gene_off.size - 1
    D = np.zeros(ng)
    p = np.zeros(ng)
    st = Stats()
    _lib.check(_lib.lib().scdd_ks(vals.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data, ng, int(old_pvalue),
                                  D.ctypes.data, p.ctypes.data, C.byref(st)))
    return D, p


def fedp(logy, cond01, gene_off, cls_oa, cls_c, cdr=None, min_size=3):
    """Per-gene pieces of feDP (R/classify.dd.R:337-374) for every gene of the block: returns
    (p_shift, p_fisher, counts[ngenes, 3] = luOutlier of oa / c1 / c2, stats)."""
    logy, cond01, gene_off = _prep(logy, cond01, gene_off)
    cls_oa = np.ascontiguousarray(cls_oa, dtype=np.int32)
    cls_c = np.ascontiguousarray(cls_c, dtype=np.int32)
    assert cls_oa.size == logy.size and cls_c.size == logy.size
    if cdr is not None:
        cdr = np.ascontiguousarray(cdr, dtype=np.float64)
        assert cdr.size == logy.size
    ng = gene_off.size - 1
    p_shift, p_fisher = np.zeros(ng), np.zeros(ng)
    counts = np.zeros((ng, 3), dtype=np.int32)
    st = Stats()
    _lib.check(_lib.lib().scdd_fedp(logy.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data, ng, cls_oa.ctypes.data,
                                    cls_c.ctypes.data, _lib.ptr(cdr), int(min_size), p_shift.ctypes.data,
                                    p_fisher.ctypes.data, counts.ctypes.data, C.byref(st)))
    return p_shift, p_fisher, counts, st.as_dict()


def classify(logy, cond01, gene_off, genes, cls_oa, cls_c, p_shift=None, cfg=None, seed=1, ndraws=10000, detail=False):
    """classifyDD (R/classify.dd.R:86-238) for the genes `genes` (0-based) of the block: returns (categories, stats) with
    categories an object array of "DE" / "DP" / "DM" / "DB" / "NC", or (categories, comp_off, comp_detail, stats) where
    comp_detail[t] = df1, scale1, m1, df2, scale2, m2, flag, min, max of comparison t (see scdd_classify)."""
    cfg = cfg or make_cfg()
    logy = np.ascontiguousarray(logy, dtype=np.float64)
    cond01 = np.ascontiguousarray(cond01, dtype=np.int32)
    gene_off = np.ascontiguousarray(gene_off, dtype=np.int64)
    cls_oa = np.ascontiguousarray(cls_oa, dtype=np.int32)
    cls_c = np.ascontiguousarray(cls_c, dtype=np.int32)
    genes = np.ascontiguousarray(genes, dtype=np.int32)
    if p_shift is not None:
        p_shift = np.ascontiguousarray(p_shift, dtype=np.float64)
    ng, nsel = gene_off.size - 1, genes.size
    o = _lib.ClassifyOpts()
    _lib.lib().scdd_default_classify_opts(C.byref(o))
    o.seed, o.ndraws = int(seed), int(ndraws)
    comp_off = comp = None
    if detail:
        comp_off = np.zeros(nsel + 1, dtype=np.int64)
        comp = np.zeros((max(1, nsel) * 25, 9), dtype=np.float64)
        o.comp_off, o.comp_detail, o.comp_cap = comp_off.ctypes.data, comp.ctypes.data, comp.shape[0]
    cat = np.zeros(nsel, dtype=np.int32)
    st = Stats()
    _lib.check(_lib.lib().scdd_classify(logy.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data, ng, cls_oa.ctypes.data,
                                        cls_c.ctypes.data, genes.ctypes.data, nsel, _lib.ptr(p_shift), C.byref(cfg),
                                        C.byref(o), cat.ctypes.data, C.byref(st)))
    names = np.array(_lib.CATEGORIES, dtype=object)[cat]
    if detail:
        return names, comp_off, comp[:int(comp_off[-1])], st.as_dict()
    return names, st.as_dict()


def student_t_draws(df, n, seed=1):
    """n variates of the device's Student-t generator (the one behind classify): for distribution tests"""
    out = np.zeros(int(n), dtype=np.float64)
    _lib.check(_lib.lib().scdd_selftest_student_t(float(df), int(n), int(seed), out.ctypes.data))
    return out


def gene_seeds(seed, ngenes):
    """L'Ecuyer-CMRG state of every bplapply element for RNGseed = seed (see scdd_rng_gene_seeds)."""
    out = np.zeros((ngenes, 6), dtype=np.uint32)
    _lib.check(_lib.lib().scdd_rng_gene_seeds(C.c_uint32(int(seed) & 0xFFFFFFFF), ngenes, out.ctypes.data))
    return out


def sample_perms(state6, n, B):
    """B draws of sample(1:n) from a L'Ecuyer state (host). Returns (perms[B, n] 1-based, new_state)."""
    st = np.ascontiguousarray(state6, dtype=np.uint32).copy()
    out = np.zeros((B, n), dtype=np.int32)
    _lib.check(_lib.lib().scdd_rng_sample_perms(st.ctypes.data, n, B, out.ctypes.data))
    return out, st


def mt_sample_perms(seed, n, count):
    """set.seed(seed); replicate(count, sample(1:n)) with R's default Mersenne-Twister (host)."""
    out = np.zeros((count, n), dtype=np.int32)
    _lib.check(_lib.lib().scdd_rng_mt_sample_perms(C.c_uint32(int(seed) & 0xFFFFFFFF), n, count, out.ctypes.data))
    return out


class Plan:
    """Gene block resident in HBM (throughput harness; scdd_plan_* of the C ABI)."""

    def __init__(self, device, logy, cond01, gene_off, chunk_perms, seed6, cfg=None, cdr=None):
        cfg = cfg or make_cfg()
        logy, cond01, gene_off = _prep(logy, cond01, gene_off)
        self.ngenes = gene_off.size - 1
        self.chunk_perms = int(chunk_perms)
        seed6 = np.ascontiguousarray(seed6, dtype=np.uint32).reshape(self.ngenes, 6)
        if cdr is not None:
            cdr = np.ascontiguousarray(cdr, dtype=np.float64)
        self._h = _lib.lib().scdd_plan_create(int(device), logy.ctypes.data, gene_off.ctypes.data, cond01.ctypes.data,
                                              self.ngenes, C.byref(cfg), self.chunk_perms, seed6.ctypes.data,
                                              _lib.ptr(cdr))
        if not self._h:
            raise ScddError(-3, _lib.last_error())

    def run_chunk(self, stream_ptr=0, nperms=None):
        _lib.check(_lib.lib().scdd_plan_run_chunk(self._h, C.c_void_p(stream_ptr), int(nperms or self.chunk_perms)))

    def bf_device_ptr(self):
        return _lib.lib().scdd_plan_bf_device_ptr(self._h)

    def stats(self):
        st = Stats()
        _lib.check(_lib.lib().scdd_plan_stats(self._h, C.byref(st)))
        return st.as_dict()

    def gene_costs(self):
        """EM work per gene measured on the first chunk (scdd_plan_gene_costs)."""
        out = np.zeros(self.ngenes)
        _lib.check(_lib.lib().scdd_plan_gene_costs(self._h, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            _lib.lib().scdd_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
