/*
 * scdd_oracle.h -- CPU ORACLE for the scDD mixture-fit / permutation hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (scdd_b200/) never links, imports or executes anything here.
 *
 * PARITY UNPINNED: the reference (scDD 1.19.1, pure R) ships no tests, no
 * golden vectors and cannot run in the build container (no R, no mclust, no
 * gfortran).  This oracle is a plain-C restatement of
 *   - the reference's R code, cited file:line below, and
 *   - the published algorithm of CRAN `mclust` (unversioned dependency,
 *     DESCRIPTION:17; 5.4.x/6.x semantics: qclass -> me1v -> BIC -> pickBIC ->
 *     summaryMclustBIC), of base R `stats::ks.test`, `sample`, `set.seed`
 *     and the Mersenne-Twister / L'Ecuyer-CMRG generators,
 * pinned only by closed forms, an independent numpy restatement
 * (tests/np_restatement.py), scipy cross-checks and R RNG known answers
 * recalled from R's documentation/behaviour (tests/test_oracle_rng.py).
 *
 * x86-64 `long double` (80-bit) is used where R accumulates in LDOUBLE
 * (sum, mean, cumsum, colMeans, var).
 */
#ifndef SCDD_ORACLE_H
#define SCDD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_GMAX 5

typedef struct {
    double alpha, m0, s0, a0, b0; /* R/scDD.R:245-249 */
    int32_t min_size;             /* R/scDD.R:223 */
    int32_t restrict_;            /* mclustRestricted(restrict=) */
    int32_t qclass_old;           /* 1: mclust <= 5.2 qclass */
    int32_t extra_mstep;          /* 1: summaryMclustBIC extra M-step (App. A.5) */
    int32_t ks_old_pvalue;        /* 1: R < 4.2 `1 - pKS2` p-value form */
    int32_t pad_;
} orc_cfg;

/* one row of the G = 1..5 table that every Mclust(y, G = ...) call re-reads */
typedef struct {
    int32_t valid;   /* BIC not NA */
    int32_t G;
    int32_t iters;
    int32_t status;  /* 0 ok, -1 sigma-squared below eps, -3 mixing proportion below sqrt(eps), -9 G > n */
    int32_t extra_applied;
    int32_t pad_;
    double loglik, bic;
    double mean[ORC_GMAX], sigsq[ORC_GMAX], pro[ORC_GMAX];          /* $parameters (after extra M-step) */
    double mean_em[ORC_GMAX], sigsq_em[ORC_GMAX], pro_em[ORC_GMAX]; /* EM-final */
} orc_row;

/* what mclustRestricted returns (R/mclust.restricted.R:249-252) + diagnostics */
typedef struct {
    int32_t G;          /* number of components of the returned model */
    int32_t model_v;    /* 0 = "X" (G=1), 1 = "V" */
    int32_t n;
    int32_t df;
    double bic, loglik;
    double mean[ORC_GMAX], var[ORC_GMAX]; /* ordered by mean */
    int32_t n_mclust_calls; /* how many Mclust() calls the reference would have made */
    int32_t pad_;
} orc_fit;

void orc_default_cfg(orc_cfg* c);

/* mclust: table of fits G=1..5 for one vector.  labels: [5*n] MAP labels (1-based), row-major by G */
int orc_fit_table(const double* y, int32_t n, const orc_cfg* cfg, orc_row* rows, int32_t* labels);

/* R/mclust.restricted.R:46-253.  cls: [n] labels (1-based, relabelled) */
int orc_mclust_restricted(const double* y, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls);

/* R/joint.posterior.R:35-64 */
double orc_joint_posterior(const double* y, const int32_t* cls, int32_t n,
                           double alpha, double m0, double s0, double a0, double b0);

/* stats::ks.test(x, y, exact = FALSE), two-sided */
int orc_ks(const double* x, int32_t nx, const double* y, int32_t ny, int32_t old_pvalue,
           double* D, double* p);

/* ---- R RNG ---- */
typedef struct {
    int32_t kind;          /* 0 = Mersenne-Twister, 1 = L'Ecuyer-CMRG */
    int32_t mti;
    uint32_t mt[624];
    uint32_t lec[6];
} orc_rng;

void orc_rng_set_seed(orc_rng* r, int32_t kind, uint32_t seed);   /* set.seed(seed, kind=) */
void orc_rng_set_lecuyer(orc_rng* r, const uint32_t* seed6);      /* .Random.seed[2:7] */
double orc_unif_rand(orc_rng* r);
double orc_unif_index(orc_rng* r, double dn);                     /* R_unif_index, sample.kind="Rejection" */
void orc_sample_perm(orc_rng* r, int32_t n, int32_t* out);        /* sample(1:n), 1-based */
void orc_lecuyer_next_stream(const uint32_t* in6, uint32_t* out6);    /* parallel::nextRNGStream */
void orc_lecuyer_next_substream(const uint32_t* in6, uint32_t* out6); /* parallel::nextRNGSubStream */

/* mclustRestricted under a live RNG stream: applies the `runif(n, -0.1, 0.1)` jitter of R/mclust.restricted.R:48-50
 * to a local copy when all values are identical (rng == NULL: no jitter, such a vector has no valid model).
 * *jittered (nullable) reports whether the jitter fired. */
int orc_mclust_restricted_rng(const double* y, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls,
                              orc_rng* rng, int32_t* jittered);
int orc_jitter_if_constant(double* y, int32_t n, orc_rng* rng);

/* ---- batch drivers (OpenMP over genes; double as the CPU baseline) ---- */
/* genes in CSR form: logy/cond hold, per gene, the log nonzero values in cell order and
 * cond (0 = reference condition, 1 = other) of those cells. */

/* R/scDD.R:373-391 genefit.  fit arrays are [3*ngenes] (oa, c1, c2 per gene);
 * cls_oa: [total nnz] labels of oa; cls_c: [total nnz] labels of c1/c2 scattered back to cell order */
int orc_genefit_batch(const double* logy, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                      const orc_cfg* cfg, orc_fit* fits, int32_t* cls_oa, int32_t* cls_c,
                      double* bf, double* den, int32_t nthreads);

/* R/permutations.R:211-286 (unadjusted).  Either perm_idx (1-based sample() output,
 * [sum_g n_g*B], gene-major then perm-major) or per-gene L'Ecuyer seeds (seed6: [ngenes*6]) */
int orc_perm_batch(const double* logy, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                   int32_t B, const int32_t* perm_idx, const uint32_t* seed6,
                   const orc_cfg* cfg, double* bf_perm, int32_t nthreads);

/* R/permutations.R:230-267 (adjust.perms=TRUE; residual permutation).  cdr: per nonzero cell */
int orc_perm_batch_cov(const double* logy, const int32_t* cond, const double* cdr, const int64_t* gene_off,
                       int32_t ngenes, int32_t B, const int32_t* perm_idx, const uint32_t* seed6,
                       const orc_cfg* cfg, double* bf_perm, int32_t nthreads);

/* Both permutation drivers with the options the plain entry points fix:
 *   cdr        NULL = unadjusted (permMclustGene / permMclust), else the adjusted branch (permMclustCov)
 *   seed_mode  0: seed6 = [ngenes*6], one stream per gene (parallelBy = "Genes", R/scDD.R:458)
 *              1: seed6 = [B*6], permutation b of every gene starts stream b (parallelBy = "Permutations",
 *                 bplapply(1:nperms) inside permMclust / permMclustCov, R/permutations.R:177, :99)
 *   n_jittered (nullable) number of fit-sets that drew the constant-vector jitter from their stream
 *   jp_terms   (nullable) [ngenes*B*sets] the jointPosterior terms of every gene-permutation: JP(c1), JP(c2)[, JP(oa)] */
int orc_perm_batch_ex(const double* logy, const int32_t* cond, const double* cdr, const int64_t* gene_off,
                      int32_t ngenes, int32_t B, const int32_t* perm_idx, const uint32_t* seed6, int32_t seed_mode,
                      const orc_cfg* cfg, double* bf_perm, int32_t nthreads, int64_t* n_jittered, double* jp_terms);

/* R/Test_KS.R:56-65 on raw nonzero values */
int orc_ks_batch(const double* vals, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                 int32_t old_pvalue, double* D, double* p, int32_t nthreads);

/* work counters of the last orc_fit_table call on this thread: sum over EM iterations of n*G, and of n */
void orc_last_work(double* triples, double* point_iters);

#ifdef __cplusplus
}
#endif
#endif
