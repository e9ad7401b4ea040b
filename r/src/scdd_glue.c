/*
 * scdd_glue.c -- .Call glue between the reference's R code and libscdd_b200.so.
 *
 * Built by `R CMD INSTALL` where R exists:
 *   PKG_CPPFLAGS = -I<repo>/include     PKG_LIBS = -L<repo>/scdd_b200 -lscdd_b200
 * R is installed neither in the build image nor on the GPU box, so the tests compile this file with -Wall -Wextra
 * -Werror against a mock of R's C API (tests/mock_R) and EXECUTE every routine through it (tests/test_r_glue.py):
 * registered arities, PROTECT balance, argument type checks, error and interrupt paths.
 * Pure marshalling: R vectors in, R vectors out; errors are raised with Rf_error() only AFTER the library
 * call has returned (never a longjmp through C++/CUDA frames).  Must be called from the master R process,
 * not from inside bplapply workers (CUDA contexts do not survive fork()).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "scdd_b200.h"

/* R_CheckUserInterrupt() longjmps on a pending interrupt, which must not happen under C++/CUDA frames: the library
 * polls this hook between permutation chunks (always on the calling = main R thread) and unwinds by itself. */
static void check_interrupt_fn(void *dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int32_t interrupt_hook(void *arg) { (void)arg; return R_ToplevelExec(check_interrupt_fn, NULL) == FALSE; }

static void need_type(SEXP x, int type, const char *what) {
    if (!isNull(x) && TYPEOF(x) != type) Rf_error("scdd_b200: %s must be %s (use as.%s())", what,
                                                  type == INTSXP ? "integer" : "double", type == INTSXP ? "integer" : "double");
}

static scdd_cfg cfg_from(SEXP prior, SEXP min_size) {
    /* prior = c(alpha, mu0, s0, a0, b0)  (R/scDD.R:245-249) */
    scdd_cfg c;
    scdd_default_cfg(&c);
    if (TYPEOF(prior) != REALSXP || LENGTH(prior) != 5) Rf_error("scdd_b200: prior must be c(alpha, mu0, s0, a0, b0)");
    const double *p = REAL(prior);
    c.alpha = p[0]; c.m0 = p[1]; c.s0 = p[2]; c.a0 = p[3]; c.b0 = p[4];
    c.min_size = asInteger(min_size);
    return c;
}

/* gene_off is passed from R as a double vector (R has no int64): converted here */
static int64_t *offsets(SEXP off, int *ngenes) {
    if (TYPEOF(off) != REALSXP || LENGTH(off) < 1) Rf_error("scdd_b200: off must be a double vector of length ngenes + 1");
    int n = LENGTH(off);
    int64_t *o = (int64_t *)R_alloc(n, sizeof(int64_t));
    for (int i = 0; i < n; i++) o[i] = (int64_t)REAL(off)[i];
    *ngenes = n - 1;
    return o;
}

/* genefit over all genes: replaces the bplapply at R/scDD.R:343 / :393 */
SEXP C_scdd_fit_observed(SEXP logy, SEXP off, SEXP cond01, SEXP prior, SEXP min_size, SEXP want_bf) {
    int ng;
    need_type(logy, REALSXP, "logy"); need_type(cond01, INTSXP, "cond01");
    int64_t *o = offsets(off, &ng);
    scdd_cfg cfg = cfg_from(prior, min_size);
    R_xlen_t nnz = XLENGTH(logy);
    SEXP fits = PROTECT(allocVector(RAWSXP, (R_xlen_t)3 * ng * sizeof(scdd_fit)));
    SEXP cls_oa = PROTECT(allocVector(INTSXP, nnz));
    SEXP cls_c = PROTECT(allocVector(INTSXP, nnz));
    SEXP bf = PROTECT(allocVector(REALSXP, ng));
    SEXP den = PROTECT(allocVector(REALSXP, ng));
    int wb = asLogical(want_bf);
    scdd_stats st;
    int rc = scdd_fit_observed(REAL(logy), o, INTEGER(cond01), ng, &cfg, (scdd_fit *)RAW(fits),
                               INTEGER(cls_oa), INTEGER(cls_c), wb ? REAL(bf) : NULL, wb ? REAL(den) : NULL, &st);
    if (rc != SCDD_OK && rc != SCDD_ERR_FIT_FAILED) { UNPROTECT(5); Rf_error("scdd_b200: %s", scdd_last_error()); }
    /* unpack scdd_fit into numeric matrices so the R side can rebuild the mclustRestricted lists */
    SEXP G = PROTECT(allocVector(INTSXP, 3 * ng)), bic = PROTECT(allocVector(REALSXP, 3 * ng));
    SEXP ll = PROTECT(allocVector(REALSXP, 3 * ng)), n = PROTECT(allocVector(INTSXP, 3 * ng));
    SEXP mean = PROTECT(allocMatrix(REALSXP, SCDD_GMAX, 3 * ng)), var = PROTECT(allocMatrix(REALSXP, SCDD_GMAX, 3 * ng));
    const scdd_fit *f = (const scdd_fit *)RAW(fits);
    for (int i = 0; i < 3 * ng; i++) {
        INTEGER(G)[i] = f[i].G; INTEGER(n)[i] = f[i].n; REAL(bic)[i] = f[i].bic; REAL(ll)[i] = f[i].loglik;
        memcpy(REAL(mean) + (size_t)SCDD_GMAX * i, f[i].mean, sizeof(double) * SCDD_GMAX);
        memcpy(REAL(var) + (size_t)SCDD_GMAX * i, f[i].var, sizeof(double) * SCDD_GMAX);
    }
    const char *nm[] = {"G", "n", "bic", "loglik", "mean", "var", "class.oa", "class.c", "bf", "den", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, G); SET_VECTOR_ELT(out, 1, n); SET_VECTOR_ELT(out, 2, bic); SET_VECTOR_ELT(out, 3, ll);
    SET_VECTOR_ELT(out, 4, mean); SET_VECTOR_ELT(out, 5, var); SET_VECTOR_ELT(out, 6, cls_oa);
    SET_VECTOR_ELT(out, 7, cls_c); SET_VECTOR_ELT(out, 8, bf); SET_VECTOR_ELT(out, 9, den);
    UNPROTECT(12);
    return out;
}

/* permMclustGene / permMclust / permMclustCov over all genes: replaces R/scDD.R:420-465.
 * perm_idx: integer vector from R's own sample() calls (or NULL); seed6: integer matrix of .Random.seed[2:7] --
 * 6 x ngenes, one stream per gene (by_perm = FALSE, parallelBy = "Genes") or 6 x B, one stream per permutation
 * (by_perm = TRUE, parallelBy = "Permutations"); cdr: numeric (adjust.perms) or NULL.
 * Returns list(bf.perm = B x ngenes matrix, const.sets, failed.sets).  Interruptible: Ctrl-C between chunks of
 * 16 permutations ends the call with an R error after the library has unwound. */
SEXP C_scdd_perm_bf(SEXP logy, SEXP off, SEXP cond01, SEXP B, SEXP perm_idx, SEXP seed6, SEXP by_perm, SEXP cdr,
                    SEXP prior, SEXP min_size) {
    int ng;
    need_type(logy, REALSXP, "logy"); need_type(cond01, INTSXP, "cond01"); need_type(perm_idx, INTSXP, "perm_idx");
    need_type(seed6, INTSXP, "seed6"); need_type(cdr, REALSXP, "cdr");
    int64_t *o = offsets(off, &ng);
    scdd_cfg cfg = cfg_from(prior, min_size);
    int b = asInteger(B);
    if (b < 0 || b == NA_INTEGER) Rf_error("scdd_b200: permutations must be a non-negative integer");
    const int bp = asLogical(by_perm) == TRUE;
    if (!isNull(seed6) && XLENGTH(seed6) != (R_xlen_t)6 * (bp ? b : ng))
        Rf_error("scdd_b200: seed6 must be a 6 x %d integer matrix", bp ? b : ng);
    if (!isNull(perm_idx) && XLENGTH(perm_idx) != (R_xlen_t)o[ng] * b)
        Rf_error("scdd_b200: perm_idx must hold permutations x (nonzero cells) indices");
    SEXP out = PROTECT(allocMatrix(REALSXP, b, ng));   /* column g = bf.perm[[g]] */
    scdd_stats st;
    scdd_perm_opts opts;
    memset(&opts, 0, sizeof(opts));
    opts.seed_mode = bp;
    opts.interrupt = interrupt_hook;
    int rc = scdd_perm_bf_ex(REAL(logy), o, INTEGER(cond01), ng, b,
                             isNull(perm_idx) ? NULL : INTEGER(perm_idx),
                             isNull(seed6) ? NULL : (const uint32_t *)INTEGER(seed6),
                             isNull(cdr) ? NULL : REAL(cdr), &cfg, &opts, REAL(out), NULL, &st);
    if (rc == SCDD_ERR_INTERRUPTED) { UNPROTECT(1); Rf_error("scdd_b200: interrupted"); }
    if (rc != SCDD_OK && rc != SCDD_ERR_FIT_FAILED) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    const char *nm[] = {"bf.perm", "const.sets", "failed.sets", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(res, 0, out);
    SET_VECTOR_ELT(res, 1, ScalarReal(st.const_sets));
    SET_VECTOR_ELT(res, 2, ScalarReal(st.failed_sets));
    UNPROTECT(2);
    return res;
}

/* BiocParallel::register(param) of R/scDD.R:314 becomes a device list: later calls shard the genes over them */
SEXP C_scdd_set_devices(SEXP devices) {
    need_type(devices, INTSXP, "devices");
    int rc = scdd_set_devices(isNull(devices) ? NULL : INTEGER(devices), isNull(devices) ? 0 : LENGTH(devices));
    if (rc != SCDD_OK) Rf_error("scdd_b200: %s", scdd_last_error());
    return R_NilValue;
}
SEXP C_scdd_release_cache(void) { scdd_release_cache(); return R_NilValue; }

/* testKS::onegene over all genes: replaces the bplapply at R/Test_KS.R:67 */
SEXP C_scdd_ks(SEXP vals, SEXP off, SEXP cond01) {
    int ng;
    need_type(vals, REALSXP, "vals"); need_type(cond01, INTSXP, "cond01");
    int64_t *o = offsets(off, &ng);
    SEXP D = PROTECT(allocVector(REALSXP, ng)), p = PROTECT(allocVector(REALSXP, ng));
    int rc = scdd_ks(REAL(vals), o, INTEGER(cond01), ng, 0, REAL(D), REAL(p), NULL);
    if (rc != SCDD_OK) { UNPROTECT(2); Rf_error("scdd_b200: %s", scdd_last_error()); }
    const char *nm[] = {"D", "p", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, D); SET_VECTOR_ELT(out, 1, p);
    UNPROTECT(3);
    return out;
}

/* dense matrix -> CSR-over-genes in one threaded pass each (replaces t(dat)[t(dat > 0)] style R code).
 * dat: numeric genes x cells matrix (column-major, as R stores it); cond01: integer per cell. */
SEXP C_scdd_gene_csr(SEXP dat, SEXP nrow, SEXP ncol, SEXP cond01, SEXP take_log, SEXP incl_zero) {
    int ng = asInteger(nrow), nc = asInteger(ncol), iz = asLogical(incl_zero);
    need_type(dat, REALSXP, "dat"); need_type(cond01, INTSXP, "cond01");
    if (ng < 0 || nc < 0 || XLENGTH(dat) != (R_xlen_t)ng * nc || LENGTH(cond01) != nc)
        Rf_error("scdd_b200: dat must be a genes x cells double matrix with one condition entry per cell");
    int32_t *n0 = (int32_t *)R_alloc(ng > 0 ? ng : 1, sizeof(int32_t)), *n1 = (int32_t *)R_alloc(ng > 0 ? ng : 1, sizeof(int32_t));
    int64_t *o = (int64_t *)R_alloc((size_t)ng + 1, sizeof(int64_t));
    int rc = scdd_dense_scan(REAL(dat), ng, nc, 1, INTEGER(cond01), n0, n1, NULL, NULL, NULL);
    if (rc != SCDD_OK) Rf_error("scdd_b200: %s", scdd_last_error());
    SEXP off = PROTECT(allocVector(REALSXP, (R_xlen_t)ng + 1));
    o[0] = 0; REAL(off)[0] = 0;
    for (int g = 0; g < ng; g++) { o[g + 1] = o[g] + (iz ? nc : n0[g] + n1[g]); REAL(off)[g + 1] = (double)o[g + 1]; }
    SEXP vals = PROTECT(allocVector(REALSXP, (R_xlen_t)o[ng])), c = PROTECT(allocVector(INTSXP, (R_xlen_t)o[ng]));
    rc = scdd_dense_to_csr(REAL(dat), ng, nc, 1, INTEGER(cond01), NULL, ng, asLogical(take_log), iz, o, REAL(vals), INTEGER(c));
    if (rc != SCDD_OK) { UNPROTECT(3); Rf_error("scdd_b200: %s", scdd_last_error()); }
    const char *nm[] = {"vals", "cond01", "off", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, vals); SET_VECTOR_ELT(out, 1, c); SET_VECTOR_ELT(out, 2, off);
    UNPROTECT(4);
    return out;
}

/* C[y>0] of every gene in CSR order (R/permutations.R:216): the cdr argument of C_scdd_perm_bf / C_scdd_fedp */
SEXP C_scdd_cell_covariate(SEXP dat, SEXP nrow, SEXP ncol, SEXP values, SEXP off) {
    int ng = asInteger(nrow), nc = asInteger(ncol), nsel;
    need_type(dat, REALSXP, "dat"); need_type(values, REALSXP, "values");
    if (ng < 0 || nc < 0 || XLENGTH(dat) != (R_xlen_t)ng * nc || LENGTH(values) != nc)
        Rf_error("scdd_b200: dat must be a genes x cells double matrix with one covariate value per cell");
    int64_t *o = offsets(off, &nsel);
    if (nsel != ng) Rf_error("scdd_b200: off must have one entry per gene plus one");
    SEXP out = PROTECT(allocVector(REALSXP, (R_xlen_t)o[ng]));
    int rc = scdd_dense_cell_covariate(REAL(dat), ng, nc, 1, REAL(values), NULL, ng, o, REAL(out));
    if (rc != SCDD_OK) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    UNPROTECT(1);
    return out;
}

/* classifyDD, R/classify.dd.R:86-238, for the genes `genes` (1-based) of the block: integer codes 0..4 =
 * NC, DE, DP, DM, DB.  p_shift: the unadjusted mean-shift p-values of C_scdd_fedp (all genes) or NULL. */
SEXP C_scdd_classify(SEXP vals, SEXP off, SEXP cond01, SEXP cls_oa, SEXP cls_c, SEXP genes, SEXP p_shift, SEXP prior,
                     SEXP min_size, SEXP seed) {
    int ng;
    need_type(vals, REALSXP, "vals"); need_type(cond01, INTSXP, "cond01"); need_type(cls_oa, INTSXP, "class.oa");
    need_type(cls_c, INTSXP, "class.c"); need_type(genes, INTSXP, "genes");
    if (p_shift != R_NilValue) need_type(p_shift, REALSXP, "p.shift");
    int64_t *o = offsets(off, &ng);
    if (XLENGTH(vals) != (R_xlen_t)o[ng] || XLENGTH(cond01) != XLENGTH(vals) || XLENGTH(cls_oa) != XLENGTH(vals) ||
        XLENGTH(cls_c) != XLENGTH(vals) || (p_shift != R_NilValue && LENGTH(p_shift) != ng))
        Rf_error("scdd_b200: vals, cond01, class.oa, class.c must have off[ngenes+1] entries, p.shift one per gene");
    scdd_cfg cfg = cfg_from(prior, min_size);
    scdd_classify_opts op;
    scdd_default_classify_opts(&op);
    op.seed = (uint64_t)(unsigned int)asInteger(seed);
    int n = LENGTH(genes);
    int32_t *rows = (int32_t *)R_alloc(n > 0 ? n : 1, sizeof(int32_t));
    for (int j = 0; j < n; j++) rows[j] = INTEGER(genes)[j] - 1;
    SEXP out = PROTECT(allocVector(INTSXP, n));
    int rc = scdd_classify(REAL(vals), o, INTEGER(cond01), ng, INTEGER(cls_oa), INTEGER(cls_c), rows, n,
                           p_shift == R_NilValue ? NULL : REAL(p_shift), &cfg, &op, INTEGER(out), NULL);
    if (rc != SCDD_OK) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    UNPROTECT(1);
    return out;
}

/* testZeroes, R/classify.dd.R:262-286: per row of `these` (1-based) the Wald p-value of the condition coefficient of
 * arm::bayesglm(y > 0 ~ detection + factor(cond), binomial); level2 = as.integer(factor(cond)) - 1L; NA for a gene
 * without zeros.  One kernel for all genes instead of a bplapply of bayesglm calls. */
SEXP C_scdd_test_zeroes(SEXP dat, SEXP nrow, SEXP ncol, SEXP level2, SEXP these) {
    int ng = asInteger(nrow), nc = asInteger(ncol);
    need_type(dat, REALSXP, "dat"); need_type(level2, INTSXP, "level2"); need_type(these, INTSXP, "these");
    if (ng < 0 || nc < 0 || XLENGTH(dat) != (R_xlen_t)ng * nc || LENGTH(level2) != nc)
        Rf_error("scdd_b200: dat must be a genes x cells double matrix with one condition entry per cell");
    int n = LENGTH(these);
    int32_t *rows = (int32_t *)R_alloc(n > 0 ? n : 1, sizeof(int32_t));
    for (int j = 0; j < n; j++) rows[j] = INTEGER(these)[j] - 1;
    SEXP out = PROTECT(allocVector(REALSXP, n));
    int rc = scdd_test_zeroes(REAL(dat), ng, nc, 1, INTEGER(level2), rows, n, NULL, REAL(out), NULL, NULL);
    if (rc != SCDD_OK) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    for (int j = 0; j < n; j++) if (ISNAN(REAL(out)[j])) REAL(out)[j] = NA_REAL;
    UNPROTECT(1);
    return out;
}

/* the per-gene filters of scDD() in one threaded pass over the matrix: nonzero cells per condition (tofit,
 * R/scDD.R:270-274), "all nonzero values of that condition identical" flags (skipConstant, :292-298) and the number of
 * negative entries (:262) */
SEXP C_scdd_dense_scan(SEXP dat, SEXP nrow, SEXP ncol, SEXP cond01) {
    int ng = asInteger(nrow), nc = asInteger(ncol);
    need_type(dat, REALSXP, "dat"); need_type(cond01, INTSXP, "cond01");
    if (ng < 0 || nc < 0 || XLENGTH(dat) != (R_xlen_t)ng * nc || LENGTH(cond01) != nc)
        Rf_error("scdd_b200: dat must be a genes x cells double matrix with one condition entry per cell");
    SEXP n0 = PROTECT(allocVector(INTSXP, ng)), n1 = PROTECT(allocVector(INTSXP, ng));
    SEXP k0 = PROTECT(allocVector(LGLSXP, ng)), k1 = PROTECT(allocVector(LGLSXP, ng));
    uint8_t *b0 = (uint8_t *)R_alloc(ng > 0 ? ng : 1, 1), *b1 = (uint8_t *)R_alloc(ng > 0 ? ng : 1, 1);
    int64_t neg = 0;
    int rc = scdd_dense_scan(REAL(dat), ng, nc, 1, INTEGER(cond01), INTEGER(n0), INTEGER(n1), b0, b1, &neg);
    if (rc != SCDD_OK) { UNPROTECT(4); Rf_error("scdd_b200: %s", scdd_last_error()); }
    for (int g = 0; g < ng; g++) { LOGICAL(k0)[g] = b0[g]; LOGICAL(k1)[g] = b1[g]; }
    const char *nm[] = {"nnz0", "nnz1", "const0", "const1", "n.negative", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, n0); SET_VECTOR_ELT(out, 1, n1); SET_VECTOR_ELT(out, 2, k0); SET_VECTOR_ELT(out, 3, k1);
    SET_VECTOR_ELT(out, 4, ScalarReal((double)neg));
    UNPROTECT(5);
    return out;
}

/* Zhat matrix of R/scDD.R:528-552 (which = -1 combined, 0 = c1, 1 = c2); tofit: 1-based rows of the fitted genes,
 * labels: their class labels in CSR order.  Replaces the `for (g in 1:nrow(...))` loop. */
SEXP C_scdd_zhat(SEXP dat, SEXP nrow, SEXP ncol, SEXP cond01, SEXP which, SEXP tofit, SEXP off, SEXP labels) {
    int ng = asInteger(nrow), nc = asInteger(ncol), w = asInteger(which), nsel;
    need_type(dat, REALSXP, "dat"); need_type(cond01, INTSXP, "cond01"); need_type(tofit, INTSXP, "tofit");
    need_type(labels, INTSXP, "labels");
    int64_t *o = offsets(off, &nsel);
    int32_t *rows = (int32_t *)R_alloc(nsel > 0 ? nsel : 1, sizeof(int32_t));
    for (int j = 0; j < nsel; j++) rows[j] = INTEGER(tofit)[j] - 1;
    int cols = 0;
    for (int c = 0; c < nc; c++) cols += w < 0 || (INTEGER(cond01)[c] != 0) == (w != 0);
    SEXP Z = PROTECT(allocMatrix(REALSXP, ng, cols));
    int rc = scdd_labels_to_dense(REAL(dat), ng, nc, 1, INTEGER(cond01), w, rows, nsel, o, INTEGER(labels), REAL(Z));
    if (rc != SCDD_OK) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    UNPROTECT(1);
    return Z;
}

/* luOutlier over all genes (R/Cluster_actions.R:25): comps.all (which = -1), comps.c1 (0), comps.c2 (1) */
SEXP C_scdd_cluster_counts(SEXP labels, SEXP cond01, SEXP off, SEXP min_size, SEXP which) {
    int ng;
    need_type(labels, INTSXP, "labels"); need_type(cond01, INTSXP, "cond01");
    int64_t *o = offsets(off, &ng);
    SEXP out = PROTECT(allocVector(INTSXP, ng));
    int rc = scdd_cluster_counts(INTEGER(labels), INTEGER(cond01), o, ng, asInteger(min_size), asInteger(which), INTEGER(out));
    if (rc != SCDD_OK) { UNPROTECT(1); Rf_error("scdd_b200: %s", scdd_last_error()); }
    UNPROTECT(1);
    return out;
}

/* feDP's per-gene loop (R/classify.dd.R:337-374) for all genes of the block: mean-shift p-value (Welch, or the
 * lm condition coefficient when cdr is not NULL), Fisher p-value where the clusterings agree (NA otherwise) and the
 * three luOutlier counts.  labels as returned by C_scdd_fit_observed ("class.oa", "class.c"). */
SEXP C_scdd_fedp(SEXP logy, SEXP off, SEXP cond01, SEXP cls_oa, SEXP cls_c, SEXP cdr, SEXP min_size) {
    int ng;
    need_type(logy, REALSXP, "logy"); need_type(cond01, INTSXP, "cond01"); need_type(cls_oa, INTSXP, "cls_oa");
    need_type(cls_c, INTSXP, "cls_c"); need_type(cdr, REALSXP, "cdr");
    int64_t *o = offsets(off, &ng);
    SEXP ps = PROTECT(allocVector(REALSXP, ng)), pf = PROTECT(allocVector(REALSXP, ng));
    SEXP cnt = PROTECT(allocMatrix(INTSXP, 3, ng));
    int rc = scdd_fedp(REAL(logy), o, INTEGER(cond01), ng, INTEGER(cls_oa), INTEGER(cls_c),
                       isNull(cdr) ? NULL : REAL(cdr), asInteger(min_size), REAL(ps), REAL(pf), INTEGER(cnt), NULL);
    if (rc != SCDD_OK) { UNPROTECT(3); Rf_error("scdd_b200: %s", scdd_last_error()); }
    const char *nm[] = {"p.shift", "p.fisher", "counts", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, nm));
    SET_VECTOR_ELT(out, 0, ps); SET_VECTOR_ELT(out, 1, pf); SET_VECTOR_ELT(out, 2, cnt);
    UNPROTECT(4);
    return out;
}

SEXP C_scdd_device_count(void) { return ScalarInteger(scdd_device_count()); }

static const R_CallMethodDef calls[] = {
    {"C_scdd_fit_observed", (DL_FUNC)&C_scdd_fit_observed, 6},
    {"C_scdd_perm_bf", (DL_FUNC)&C_scdd_perm_bf, 10},
    {"C_scdd_set_devices", (DL_FUNC)&C_scdd_set_devices, 1},
    {"C_scdd_release_cache", (DL_FUNC)&C_scdd_release_cache, 0},
    {"C_scdd_ks", (DL_FUNC)&C_scdd_ks, 3},
    {"C_scdd_gene_csr", (DL_FUNC)&C_scdd_gene_csr, 6},
    {"C_scdd_dense_scan", (DL_FUNC)&C_scdd_dense_scan, 4},
    {"C_scdd_cell_covariate", (DL_FUNC)&C_scdd_cell_covariate, 5},
    {"C_scdd_zhat", (DL_FUNC)&C_scdd_zhat, 8},
    {"C_scdd_cluster_counts", (DL_FUNC)&C_scdd_cluster_counts, 5},
    {"C_scdd_fedp", (DL_FUNC)&C_scdd_fedp, 7},
    {"C_scdd_test_zeroes", (DL_FUNC)&C_scdd_test_zeroes, 5},
    {"C_scdd_classify", (DL_FUNC)&C_scdd_classify, 10},
    {"C_scdd_device_count", (DL_FUNC)&C_scdd_device_count, 0},
    {NULL, NULL, 0}};

void R_init_scDD(DllInfo *dll) {
    R_registerRoutines(dll, NULL, calls, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
