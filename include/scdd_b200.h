/*
 * scdd_b200.h -- C ABI of libscdd_b200.so: the B200 (sm_100a) implementation of scDD's
 * per-gene mixture-fitting / permutation-testing / KS hot path.
 *
 * The reference (kdkorthauer/scDD 1.19.1) is pure R and has NO FFI today; this is the
 * boundary its R shim binds through .Call (see INTEGRATION.md and r/src/scdd_glue.c).
 * Each entry point names the reference interface it replaces (file:line relative to the
 * reference tree).  Plain C: pointers + sizes, caller-allocated buffers, integer status
 * (0 = ok, <0 = error; message via scdd_last_error()).  No torch / R types.
 *
 * There is NO CPU fallback: every compute entry point fails with SCDD_ERR_NO_DEVICE when
 * no CUDA device is usable.
 *
 * Gene data layout ("CSR over genes"), shared by all entry points:
 *   gene_off[ngenes+1]  int64   offsets into the per-gene value arrays
 *   logy[nnz]           double  log of the NONZERO normalized counts of each gene, in cell order
 *                               (R/scDD.R:374-375  `y <- log(y[y>0])`)
 *   cond01[nnz]         int32   0 = reference condition (first level, R/scDD.R:258), 1 = other,
 *                               for the same cells (`cond0 <- condition[y>0]`)
 */
#ifndef SCDD_B200_H
#define SCDD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCDD_GMAX 5          /* Mclust(G = 1:5), R/mclust.restricted.R:52 */
#define SCDD_ABI_VERSION 2

enum {
    SCDD_OK = 0,
    SCDD_ERR_NO_DEVICE = -1,   /* no usable CUDA device (product path never falls back to CPU) */
    SCDD_ERR_BAD_ARG = -2,
    SCDD_ERR_CUDA = -3,
    SCDD_ERR_TOO_LARGE = -4,   /* a fit-set has more points than the kernels support */
    SCDD_ERR_FIT_FAILED = -5,  /* some fit-set had no valid model (all BIC NA) -- outputs hold NaN there */
    SCDD_ERR_INTERRUPTED = -6  /* the caller's interrupt hook asked to stop between two permutation chunks */
};

/* prior_param + the arguments of scDD() that reach the hot path (R/scDD.R:217-225, :245-249) */
typedef struct {
    double alpha, m0, s0, a0, b0;
    int32_t min_size;      /* min.size */
    int32_t restrict_;     /* mclustRestricted(restrict=) */
    int32_t extra_mstep;   /* 1: mclust summaryMclustBIC's extra M-step on $parameters */
    int32_t ks_old_pvalue; /* 1: R < 4.2 form `1 - pKS2` of the asymptotic KS p-value */
} scdd_cfg;

/* what mclustRestricted returns (R/mclust.restricted.R:249-252) minus `class` (separate array) */
typedef struct {
    int32_t G;        /* number of components == length(mean) */
    int32_t model_v;  /* 0: modelName "X" (G = 1), 1: "V" */
    int32_t n;
    int32_t df;
    double bic, loglik;
    double mean[SCDD_GMAX], var[SCDD_GMAX];   /* ordered by increasing mean */
} scdd_fit;

/* work / diagnostics counters, accumulated over one call */
typedef struct {
    double triples;        /* sum over executed EM iterations of (points x components)  [SURVEY 8d: T] */
    double point_iters;    /* sum over executed EM iterations of points                 [SURVEY 8d: P] */
    double fit_sets;       /* mclustRestricted-equivalents evaluated */
    double em_iters;       /* EM iterations summed over all G = 2..5 fits */
    double iter_cap_hits;  /* EM fits stopped by the safety cap instead of the tolerance (expected 0) */
    double tol_near;       /* EM fits whose stopping test was within 1e-7 (relative) of the tolerance */
    double bic_near;       /* fit-sets whose two best BICs differ by < 1e-9 relative (near ties) */
    double const_sets;     /* fit-sets with all values identical: jittered from the gene's stream as the reference does
                              (device-drawn permutations), or failed (no stream: observed data, perm_idx) */
    double failed_sets;    /* fit-sets with no valid model */
    double fit_kernel_ms;  /* device time of the mixture-fit kernel(s), CUDA events */
    double fit_kernel_launches;
    double rng_kernel_ms;  /* device time of the permutation-generation kernel(s) */
    double total_ms;       /* device time of the whole call, first H2D to last D2H */
    double h2d_bytes, d2h_bytes;
    double debug_violations; /* index/bounds checks that failed; only counted in -DSCDD_DEBUG_BOUNDS builds (else 0) */
} scdd_stats;

/* ---- library / device management ---- */
int32_t scdd_abi_version(void);
const char* scdd_last_error(void);              /* thread-local message of the last failing call */
int32_t scdd_device_count(void);                /* number of usable CUDA devices, 0 if none */
int32_t scdd_set_devices(const int32_t* devices, int32_t n); /* devices later calls shard genes over (default: current device only) */
void scdd_default_cfg(scdd_cfg* cfg);           /* scDD() defaults, R/scDD.R:217-225 */

/* ---- observed-data fits:  genefit, R/scDD.R:327-340 and :373-391 ----
 * fits[3*ngenes]: (oa, c1, c2) per gene.  cls_oa[nnz]: oa$class per nonzero cell.
 * cls_c[nnz]: c1$class / c2$class scattered back to cell order.
 * bf[ngenes] = jointPosterior(c1)+jointPosterior(c2), den[ngenes] = jointPosterior(oa)
 * (R/scDD.R:381-383, R/joint.posterior.R:35-64); either may be NULL (permutations == 0 path). */
int32_t scdd_fit_observed(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                          const scdd_cfg* cfg, scdd_fit* fits, int32_t* cls_oa, int32_t* cls_c,
                          double* bf, double* den, scdd_stats* stats);

/* ---- permutation Bayes-factor numerators:  permMclustGene R/permutations.R:211-286,
 *      permMclust :147-181, permMclustCov :44-104 ----
 * bf_perm[ngenes*B], gene-major.  Permutations come from exactly one of
 *   perm_idx : the output of R's `sample(1:n_g)` (1-based), [sum_g n_g*B], gene-major then
 *              permutation-major (host-drawn, R/permutations.R:273), or
 *   seed6    : [ngenes*6] L'Ecuyer-CMRG states (.Random.seed[2:7]) -- one stream per gene as
 *              BiocParallel gives each bplapply element (R/scDD.R:458); permutations are then
 *              drawn ON THE DEVICE with a bit-exact re-implementation of R's sampler
 *              (R_unif_index rejection sampling); seed6_out (nullable) receives the advanced states.
 * cdr != NULL selects adjust.perms=TRUE (R/permutations.R:230-267): cdr[nnz] is the cellular
 * detection rate C of each nonzero cell; residuals of lm(y ~ C) are permuted and oa is refitted,
 * bf_perm = JP(c1)+JP(c2)-JP(oa).
 * Constant subsets (R/mclust.restricted.R:48-50): with seed6 a permuted subset whose values are all identical gets
 * the reference's runif(n, -0.1, 0.1) jitter, drawn from the gene's stream at the position R draws it (between this
 * permutation's sample() and the next); stats->const_sets counts them.  With perm_idx no stream exists: such a
 * fit-set yields NaN and SCDD_ERR_FIT_FAILED.
 * With several devices (scdd_set_devices) the genes are sharded: 8 permutations on shards balanced by cell counts
 * measure the EM work per gene, the rest runs on shards re-dealt by that measured cost. */
int32_t scdd_perm_bf(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                     int32_t B, const int32_t* perm_idx, const uint32_t* seed6, const double* cdr,
                     const scdd_cfg* cfg, double* bf_perm, uint32_t* seed6_out, scdd_stats* stats);

/* Options of scdd_perm_bf_ex (all-zero = scdd_perm_bf).
 *   seed_mode 0: seed6 is [ngenes*6], one stream per gene = parallelBy "Genes" (bplapply over genes, R/scDD.R:458).
 *   seed_mode 1: seed6 is [B*6], permutation b of EVERY gene is drawn from the start of stream b = parallelBy
 *                "Permutations": scDD() loops over the genes in the master and permMclust / permMclustCov run
 *                bplapply(1:nperms, ...) under the same RNGseed for each of them (R/scDD.R:414-455,
 *                R/permutations.R:177, :99).  seed6_out is not written.
 *   jp_terms   : (nullable) [ngenes*B*sets], sets = 2 (3 with cdr): the jointPosterior terms JP(c1), JP(c2)[, JP(oa)]
 *                of every gene-permutation, whose combination is bf_perm (R/permutations.R:265-267, :281-282).
 *   interrupt  : (nullable) polled on the calling thread between chunks of <= 16 permutations; a non-zero return stops
 *                the call with SCDD_ERR_INTERRUPTED.  The R glue passes a hook that runs R_CheckUserInterrupt() under
 *                R_ToplevelExec(), so that a 60 s permutation job stays interruptible. */
typedef struct {
    int32_t seed_mode;
    int32_t reserved;
    double* jp_terms;
    int32_t (*interrupt)(void* arg);
    void* interrupt_arg;
} scdd_perm_opts;
int32_t scdd_perm_bf_ex(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                        int32_t B, const int32_t* perm_idx, const uint32_t* seed6, const double* cdr,
                        const scdd_cfg* cfg, const scdd_perm_opts* opts, double* bf_perm, uint32_t* seed6_out,
                        scdd_stats* stats);
/* frees the per-device buffers (HBM blocks, pinned staging, streams) that scdd_perm_bf keeps between calls */
void scdd_release_cache(void);

/* ---- KS test:  testKS / onegene, R/Test_KS.R:55-65  (stats::ks.test(x1, x2, exact=FALSE)) ----
 * vals[nnz]: the values to compare per gene (raw counts; zeros already dropped unless inclZero),
 * D[ngenes], p[ngenes] (unadjusted).  D reproduces R's long-double cumsum bit for bit. */
int32_t scdd_ks(const double* vals, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                int32_t old_pvalue, double* D, double* p, scdd_stats* stats);

/* ---- feDP:  R/classify.dd.R:314-383, the follow-up test of the genes the permutation test did not flag ----
 * Inputs are the gene block plus the class vectors scdd_fit_observed returned (cls_oa, cls_c).  Per gene:
 *   p_shift : unadjusted p-value of a mean shift of log-nonzero expression between the conditions --
 *             t.test(y ~ factor(cond)) (Welch, :353), or with cdr != NULL (adjust.perms) the condition coefficient
 *             of lm(y ~ cdr0 + factor(cond)) (:351).  NaN where the reference would jitter with runif (:345-347).
 *   p_fisher: fisher.test(table(oa class, cond)) over the non-outlier clusters (findOutliers, :360-365) when
 *             luOutlier(c1) == luOutlier(c2) == luOutlier(oa) > 1 (:358), NaN otherwise.  Exact enumeration of
 *             the r x 2 tables (r <= 5) on the device; stats->failed_sets counts tables too large to enumerate.
 *   counts3 : (nullable) [3*ngenes] luOutlier of oa, c1, c2 (:327-329).
 * BH adjustment and the NS / DP / NC labelling (:377-382) are a few vector operations left to the caller. */
int32_t scdd_fedp(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                  const int32_t* cls_oa, const int32_t* cls_c, const double* cdr, int32_t min_size,
                  double* p_shift, double* p_fisher, int32_t* counts3, scdd_stats* stats);

/* ---- R's RNG on the host (what the R shim gets for free from R itself) ----
 * scdd_rng_gene_seeds: the L'Ecuyer-CMRG state BiocParallel hands to element i of
 * bplapply(seq_len(ngenes), ..., BPPARAM = param) for RNGseed = seed:
 * set.seed(seed, kind = "L'Ecuyer-CMRG") then one nextRNGSubStream per element.  out6: [ngenes*6]. */
int32_t scdd_rng_gene_seeds(uint32_t seed, int32_t ngenes, uint32_t* out6);
/* sample(1:n) x B from one L'Ecuyer state (host; 1-based output [B*n]); state advanced in place */
int32_t scdd_rng_sample_perms(uint32_t* state6, int32_t n, int32_t B, int32_t* out);
/* Mersenne-Twister path: set.seed(seed); then `count` calls of sample(1:n); out [count*n] */
int32_t scdd_rng_mt_sample_perms(uint32_t seed, int32_t n, int32_t count, int32_t* out);

/* ---- dense genes x cells matrix <-> the layout above (host, multi-threaded marshalling; no GPU needed) ----
 * X: normalized counts, ngenes x ncells; col_major = 1 for an R matrix (X[g + c*ngenes]), 0 for C order.
 * cell_cond01[ncells]: 0 = reference condition.  These replace the apply()/for loops over genes that surround
 * the hot path in scDD().
 *
 * scdd_dense_scan: the per-gene filters -- nnz0/nnz1 = nonzero cells per condition (tofit, R/scDD.R:270-274),
 * const0/const1 = 1 when all nonzero values of that condition are identical (skipConstant, R/scDD.R:292-298),
 * n_negative = sum(X < 0) (R/scDD.R:262).  Any output may be NULL. */
int32_t scdd_dense_scan(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                        int32_t* nnz0, int32_t* nnz1, uint8_t* const0, uint8_t* const1, int64_t* n_negative);
/* scdd_dense_to_csr: gathers rows genes[nsel] (0-based; NULL = rows 0..nsel-1) into (vals, cond01).
 * gene_off[nsel+1] is INPUT: prefix sums of nnz0+nnz1 (of ncells when incl_zero).  vals = log(x) when take_log
 * (`y <- log(y[y>0])`, R/scDD.R:374-375; libm log, as R calls), zeros kept only when incl_zero (R/Test_KS.R:57-64). */
int32_t scdd_dense_to_csr(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                          const int32_t* genes, int32_t nsel, int32_t take_log, int32_t incl_zero,
                          const int64_t* gene_off, double* vals, int32_t* cond01);
/* scdd_dense_cell_covariate: a per-cell covariate -- the cellular detection rate C of R/scDD.R:457 -- spread over the
 * nonzero cells of the selected genes in CSR order: `C[y>0]` of every gene (R/permutations.R:216, R/classify.dd.R:341),
 * i.e. the `cdr` argument of scdd_perm_bf / scdd_fedp.  gene_off as for scdd_dense_to_csr (zeros dropped). */
int32_t scdd_dense_cell_covariate(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major,
                                  const double* cell_values, const int32_t* genes, int32_t nsel, const int64_t* gene_off,
                                  double* out);
/* scdd_labels_to_dense: the Zhat matrices, R/scDD.R:528-552.  Z (same layout as X) = 0 where X == 0, the class
 * label on the nonzero cells of the selected genes (labels in CSR order over ALL their nonzero cells), 1 elsewhere.
 * which = -1: all cells (Zhat.combined); 0 / 1: only that condition's cells (Zhat.c1 / Zhat.c2 -- Z then has
 * that many columns). */
int32_t scdd_labels_to_dense(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_cond01,
                             int32_t which, const int32_t* genes, int32_t nsel, const int64_t* gene_off,
                             const int32_t* labels, double* Z);
/* scdd_cluster_counts: luOutlier, R/Cluster_actions.R:25 -- per gene sum(table(class) >= min_size), over all its
 * cells (which = -1, comps.all) or those of one condition (which = 0 / 1, comps.c1 / comps.c2). */
int32_t scdd_cluster_counts(const int32_t* labels, const int32_t* cond01, const int64_t* gene_off, int32_t ngenes,
                            int32_t min_size, int32_t which, int32_t* counts);

/* ---- classifyDD, R/classify.dd.R:86-238 ----
 * The DE / DP / DM / DB / NC pattern of the genes `genes[nsel]` (0-based indices into the block) from the three
 * clusterings of scdd_fit_observed (cls_oa, cls_c): label compression (:124-158), luOutlier / findOutliers
 * (R/Cluster_actions.R:25,78), the normal-gamma posterior of every cluster (getPosteriorParams, :19-37), the
 * Monte-Carlo comparison of every (cluster of condition 1, cluster of condition 2) pair (:165-187: 10,000 Student-t
 * draws each, "different" iff all sampled differences share a sign) and the decision tree (:189-234).  p_shift[ngenes]
 * (nullable) = the unadjusted mean-shift p-values of scdd_fedp, used by the multimodal-DE rule (:206-214).
 * The draws run on the device, one warp per comparison, from Philox streams keyed by (seed, comparison index): the
 * reference's distribution, NOT R's stream (classifyDD consumes the user's RNG; see csrc/classify.cuh).
 * opts->comp_off [nsel+1] / comp_detail [comp_cap*9] (nullable) return the comparisons: df1, scale1, m1, df2, scale2,
 * m2, flag, min, max of the sampled differences (at most 25 comparisons per gene). */
enum { SCDD_CAT_NC = 0, SCDD_CAT_DE = 1, SCDD_CAT_DP = 2, SCDD_CAT_DM = 3, SCDD_CAT_DB = 4 };
typedef struct {
    int32_t ndraws;          /* 10000 */
    int32_t reserved;
    uint64_t seed;
    int64_t* comp_off;
    double* comp_detail;
    int64_t comp_cap;
} scdd_classify_opts;
void scdd_default_classify_opts(scdd_classify_opts* opts);
int32_t scdd_classify(const double* logy, const int64_t* gene_off, const int32_t* cond01, int32_t ngenes,
                      const int32_t* cls_oa, const int32_t* cls_c, const int32_t* genes, int32_t nsel,
                      const double* p_shift, const scdd_cfg* cfg, const scdd_classify_opts* opts,
                      int32_t* category, scdd_stats* stats);

/* ---- testZeroes, R/classify.dd.R:262-286 ----
 * Per gene with at least one zero: arm::bayesglm(y > 0 ~ detection + factor(cond), binomial(logit)) and the Wald
 * p-value of the condition coefficient, summary(M1)$coefficients[3, 4]; NaN (R's NA) for a gene without zeros (:267).
 * detection = colSums(X > 0) / nrow(X) over ALL ngenes rows (:263).  cell_level2[ncells]: 1 where the cell belongs to
 * the SECOND level of factor(cond) (levels are sort(unique(cond)) -- not scDD()'s `ref`); the posterior mode depends
 * on which level is the baseline because the intercept has its own prior.  these[nthese]: 0-based rows to test
 * (NULL = rows 0..nthese-1).  The host packs one bit per cell and gene in a threaded pass over X; the regressions run
 * on the first device of scdd_set_devices (they cost milliseconds, see scdd_b200/csrc/zeroes.cuh).
 * detail (nullable) [nthese*8]: coef[3], se[3], iterations, deviance.  Options follow bayesglm()'s defaults. */
typedef struct {
    double prior_scale;                 /* 2.5  (predictors; divided by max-min for a two-valued, 2 sd for another predictor) */
    double prior_scale_for_intercept;   /* 10   */
    double prior_df;                    /* 1    (Cauchy; also for the intercept) */
    double epsilon;                     /* 1e-8 glm.control */
    int32_t maxit;                      /* 100  */
    int32_t reserved;
} scdd_glm_opts;
void scdd_default_glm_opts(scdd_glm_opts* opts);
int32_t scdd_test_zeroes(const double* X, int64_t ngenes, int64_t ncells, int32_t col_major, const int32_t* cell_level2,
                         const int32_t* these, int64_t nthese, const scdd_glm_opts* opts, double* pvalue,
                         double* detail, scdd_stats* stats);

/* ---- device-resident plan (throughput harness; inputs stay in HBM between steps) ---- */
typedef struct scdd_plan scdd_plan;
/* uploads the gene block to `device`, runs the per-gene preparation kernels.
 * chunk_perms: permutations processed per scdd_plan_run_chunk call. */
scdd_plan* scdd_plan_create(int32_t device, const double* logy, const int64_t* gene_off, const int32_t* cond01,
                            int32_t ngenes, const scdd_cfg* cfg, int32_t chunk_perms, const uint32_t* seed6,
                            const double* cdr);
/* enqueues permutation generation + mixture fits + BF combine for the next `nperms` (<= chunk_perms)
 * permutations of every resident gene on `cuda_stream` (a cudaStream_t, 0 = default stream).
 * Asynchronous.  Results land in the plan's device buffer bf_chunk[ngenes*chunk_perms]. */
int32_t scdd_plan_run_chunk(scdd_plan* plan, void* cuda_stream, int32_t nperms);
/* device pointer of bf_chunk (for a D2H copy by the caller) */
void* scdd_plan_bf_device_ptr(scdd_plan* plan);
/* synchronises the device and reads the counters accumulated since plan creation */
int32_t scdd_plan_stats(scdd_plan* plan, scdd_stats* stats);
/* EM work per resident gene (point x component x iteration triples) measured on the plan's first chunk -- what the
 * library orders its own work by; exposed so that a multi-GPU host can balance its gene shards by measured cost. */
int32_t scdd_plan_gene_costs(scdd_plan* plan, double* cost /* [ngenes] */);
void scdd_plan_destroy(scdd_plan* plan);

/* ---- host-side self checks (no GPU needed) ---- */
/* software x87 extended-precision accumulate used by the KS kernel, run on the host:
 * out[i] = (double)(long-double running sum of steps[0..i]) */
int32_t scdd_selftest_x87_cumsum(const double* steps, int32_t n, double* out);
/* the kernels' exp arithmetic (host build of the same source, shift 0): out[i] = exp(x[i]) for x in [-708, 0] */
int32_t scdd_selftest_exp(const double* x, int32_t n, double* out);
/* softmax of G <= 5 log-terms exactly as the E-step evaluates it; logsumexp may be NULL */
int32_t scdd_selftest_softmax(const double* a, int32_t G, double* z, double* logsumexp);
/* the kernels' table-driven logarithm (EM loop: log of the running product of softmax sums, log(pro^2 / sigsq)):
 * out[i] = log(x[i]) for positive normal x */
int32_t scdd_selftest_log(const double* x, int32_t n, double* out);
/* the host-side tail of scdd_fedp: out[i] = 2 * pt(-|t[i]|, df[i]) */
int32_t scdd_selftest_student_p(const double* t, const double* df, int32_t n, double* out);
/* the device's Student-t generator behind scdd_classify (needs a GPU): out[i] = one t_df variate of stream i */
int32_t scdd_selftest_student_t(double df, int32_t n, uint64_t seed, double* out);
/* measured fp64 FMA throughput of the device (TFLOP/s, DFMA = 2 flop), used as the roofline denominator */
double scdd_measure_fp64_peak(int32_t device, double* sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* SCDD_B200_H */
