/*
 * scdd_oracle.c -- CPU ORACLE (test infrastructure; see scdd_oracle.h header note:
 * PARITY UNPINNED, never on the product path).
 *
 * Every function cites the reference lines (relative to /root/reference) or the
 * third-party routine (mclust / base R) whose published algorithm it restates.
 */
#include "scdd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FLMAX 1.7976931348623157e308
#define PI2LOG 1.837877066409345 /* log(2*pi), mclust.f parameter pi2log */
#define SMALOG (-708.0)

static __thread double g_triples = 0.0, g_point_iters = 0.0;

void orc_last_work(double* triples, double* point_iters) {
    *triples = g_triples;
    *point_iters = g_point_iters;
}

void orc_default_cfg(orc_cfg* c) {
    /* scDD() defaults, R/scDD.R:217-225 */
    c->alpha = 0.10; c->m0 = 0.0; c->s0 = 0.01; c->a0 = 0.01; c->b0 = 0.01;
    c->min_size = 3; c->restrict_ = 1; c->qclass_old = 0; c->extra_mstep = 1;
    c->ks_old_pvalue = 0; c->pad_ = 0;
}

/* ------------------------------------------------------------------------- */
/* base R helpers                                                            */
/* ------------------------------------------------------------------------- */

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* base::mean.default on doubles: long-double sum, then one refinement pass (summary.c real_mean) */
static double r_mean(const double* x, int n) {
    long double s = 0.0L;
    for (int i = 0; i < n; i++) s += x[i];
    s /= n;
    long double t = 0.0L;
    for (int i = 0; i < n; i++) t += (x[i] - s);
    s += t / n;
    return (double)s;
}

/* stats::sd = sqrt(var); cov.c cov_complete1-style: refined long-double mean, sum of squares / (n-1) */
static double r_sd(const double* x, int n) {
    long double s = 0.0L;
    for (int i = 0; i < n; i++) s += x[i];
    long double m = s / n;
    long double t = 0.0L;
    for (int i = 0; i < n; i++) t += (x[i] - m);
    m += t / n;
    long double ss = 0.0L;
    for (int i = 0; i < n; i++) ss += (x[i] - m) * (x[i] - m);
    return sqrt((double)(ss / (n - 1)));
}

/* stats::quantile(type = 7) on a pre-sorted vector, one prob */
static double r_quantile7(const double* xs, int n, double p) {
    double index = 1.0 + (double)(n - 1 > 0 ? n - 1 : 0) * p;
    double lo = floor(index), hi = ceil(index);
    double qs = xs[(int)lo - 1];
    double xh = xs[(int)hi - 1];
    if (index > lo && xh != qs) {
        double h = index - lo;
        qs = (1.0 - h) * qs + h * xh;
    }
    return qs;
}

/* ------------------------------------------------------------------------- */
/* mclust: qclass                                            */
/* ------------------------------------------------------------------------- */

/* current mclust (>= 5.3):  quantile breaks with unique()/gap-pruning, [q_i, q_{i+1}) classes */
static void qclass_new(const double* x, const double* xs, int n, int k, int* cl) {
    double eps = r_sd(x, n) * sqrt(DBL_EPSILON);
    int nq = k;
    int cap = 16, L = 0;
    double* q = (double*)malloc(sizeof(double) * cap);
    while (L < k + 1) {
        nq += 1;
        /* R would keep growing the grid (forever when the data has too few distinct values); the
           restatement and the CUDA kernels both stop at 64 grid points and keep the last grid */
        if (nq > 64) break;
        if (nq > cap) { cap = nq * 2; q = (double*)realloc(q, sizeof(double) * cap); }
        /* seq(0, 1, length = nq): c(0, (1:(nq-2)) * (1/(nq-1)), 1) */
        double by = 1.0 / (double)(nq - 1);
        L = 0;
        for (int j = 0; j < nq; j++) {
            double p = (j == 0) ? 0.0 : (j == nq - 1 ? 1.0 : (double)j * by);
            double v = r_quantile7(xs, n, p);
            int dup = 0;
            for (int m = 0; m < L; m++) if (q[m] == v) { dup = 1; break; }   /* unique() */
            if (!dup) q[L++] = v;
        }
    }
    if (L > k + 1) {
        /* q <- q[-order(diff(q))[1:nr]] ; stable ascending order of the gaps */
        int nd = L - 1, nr = L - k - 1;
        int* ord = (int*)malloc(sizeof(int) * nd);
        double* dq = (double*)malloc(sizeof(double) * nd);
        for (int i = 0; i < nd; i++) { dq[i] = q[i + 1] - q[i]; ord[i] = i; }
        for (int i = 1; i < nd; i++) { /* stable insertion sort */
            int o = ord[i]; int j = i - 1;
            while (j >= 0 && dq[ord[j]] > dq[o]) { ord[j + 1] = ord[j]; j--; }
            ord[j + 1] = o;
        }
        char* drop = (char*)calloc(L, 1);
        for (int i = 0; i < nr; i++) drop[ord[i]] = 1;
        int w = 0;
        for (int i = 0; i < L; i++) if (!drop[i]) q[w++] = q[i];
        L = w;
        free(ord); free(dq); free(drop);
    }
    q[0] = xs[0] - eps;
    q[L - 1] = xs[n - 1] + eps;
    for (int i = 0; i < n; i++) {
        cl[i] = 0;
        for (int c = 0; c < k && c + 1 < L; c++)
            if (x[i] >= q[c] && x[i] < q[c + 1]) cl[i] = c + 1;
    }
    free(q);
}

/* mclust <= 5.2:  q <- quantile(x, seq(0,1,by=1/k)); q[1] <- q[1]-1; (q_i, q_{i+1}] classes */
static void qclass_old(const double* x, const double* xs, int n, int k, int* cl) {
    double q[ORC_GMAX + 2];
    double by = 1.0 / (double)k;
    for (int j = 0; j <= k; j++) {
        double p = (double)j * by;
        if (p > 1.0) p = 1.0;
        q[j] = r_quantile7(xs, n, p);
    }
    q[0] -= 1.0;
    for (int i = 0; i < n; i++) {
        cl[i] = 0;
        for (int c = 0; c < k; c++)
            if (x[i] > q[c] && x[i] <= q[c + 1]) cl[i] = c + 1;
    }
}

/* ------------------------------------------------------------------------- */
/* mclust Fortran: mvn1d, me1v, ms1v               */
/* ------------------------------------------------------------------------- */

/* mvn1d: G = 1 closed form.  returns loglik (FLMAX if sigsq == 0) */
static double mvn1d(const double* x, int n, double* mu, double* sigsq) {
    double dn = (double)n, scl = 1.0 / dn, m = 0.0;
    for (int i = 0; i < n; i++) m += scl * x[i]; /* ddot(n, 1/dn, 0, x, 1) */
    double ss = 0.0;
    for (int i = 0; i < n; i++) { double t = x[i] - m; ss += t * t; }
    ss /= dn;
    *mu = m; *sigsq = ss;
    if (ss == 0.0) return FLMAX;
    return -dn * (PI2LOG + (1.0 + log(ss))) / 2.0;
}

/* me1v: EM for the 1-D unequal-variance mixture, M-step first, z is n x G column-major.
 * returns loglik (FLMAX: sigma-squared below eps; -FLMAX: mixing proportion below sqrt(eps)) */
static double me1v(const double* x, int n, int G, double* z, double tol, double eps, int maxi,
                   double* mu, double* sigsq, double* pro, int* iters) {
    double rteps = sqrt(eps);
    double hold = FLMAX / 2.0, hood = FLMAX, err = FLMAX;
    int iter = 0;
    for (int k = 0; k < G; k++) { mu[k] = 0.0; sigsq[k] = 0.0; pro[k] = 0.0; }
    for (;;) {
        iter++;
        double zsum = 1.0;
        for (int k = 0; k < G; k++) {
            double* zk = z + (size_t)k * n;
            double sum = 0.0, smu = 0.0;
            for (int i = 0; i < n; i++) { double t = zk[i]; sum += t; smu += t * x[i]; }
            pro[k] = sum / (double)n;
            if (sum < zsum) zsum = sum;
            if (sum > rteps) {
                smu /= sum;
                mu[k] = smu;
                double sk = 0.0;
                for (int i = 0; i < n; i++) {
                    double t = x[i] - smu;
                    t = t * t;
                    sk += zk[i] * t;
                    zk[i] = t;
                }
                sigsq[k] = sk / sum;
            }
        }
        if (zsum <= rteps) { *iters = iter; return -FLMAX; }
        double sigmin = FLMAX;
        for (int k = 0; k < G; k++) if (sigsq[k] < sigmin) sigmin = sigsq[k];
        if (sigmin <= eps) { *iters = iter; return FLMAX; }
        for (int k = 0; k < G; k++) {
            double* zk = z + (size_t)k * n;
            double sk = sigsq[k], cst = PI2LOG + log(sk);
            for (int i = 0; i < n; i++) zk[i] = -(cst + (zk[i] / sk)) / 2.0;
        }
        hood = 0.0;
        for (int i = 0; i < n; i++) {
            double tmax = -FLMAX;
            for (int k = 0; k < G; k++) {
                double* zk = z + (size_t)k * n;
                if (pro[k] == 0.0) zk[i] = 0.0;
                else {
                    double t = log(pro[k]) + zk[i];
                    if (t > tmax) tmax = t;
                    zk[i] = t;
                }
            }
            double sum = 0.0;
            for (int k = 0; k < G; k++) {
                double* zk = z + (size_t)k * n;
                if (pro[k] != 0.0) {
                    double t = zk[i] - tmax;
                    if (t >= SMALOG) { zk[i] = exp(t); sum += zk[i]; }
                    else zk[i] = 0.0;
                }
            }
            hood += (log(sum) + tmax);
            double rs = 1.0 / sum; /* dscal(nz, one/sum, z(i,1), n) */
            for (int k = 0; k < G; k++) z[(size_t)k * n + i] *= rs;
        }
        g_triples += (double)n * G;
        g_point_iters += (double)n;
        err = fabs(hold - hood) / (1.0 + fabs(hood));
        hold = hood;
        if (!(err > tol && iter < maxi)) break;
    }
    *iters = iter;
    return hood;
}

/* ms1v: one M-step from z (used by summaryMclustBIC's extra M-step) */
static int ms1v(const double* x, const double* z, int n, int G, double* mu, double* sigsq, double* pro) {
    int ok = 1;
    for (int k = 0; k < G; k++) {
        const double* zk = z + (size_t)k * n;
        double sum = 0.0, smu = 0.0;
        for (int i = 0; i < n; i++) { sum += zk[i]; smu += zk[i] * x[i]; }
        pro[k] = sum / (double)n;
        if (sum > 1.0 || smu <= sum * FLMAX) {
            smu /= sum;
            mu[k] = smu;
            double sk = 0.0;
            for (int i = 0; i < n; i++) { double t = fabs(x[i] - smu); sk += zk[i] * (t * t); }
            sigsq[k] = sk / sum;
        } else { mu[k] = FLMAX; sigsq[k] = FLMAX; ok = 0; }
    }
    return ok;
}

/* ------------------------------------------------------------------------- */
/* mclust R level: the G = 1..5 table (mclustBIC + summaryMclustBIC per row)   */
/* ------------------------------------------------------------------------- */

int orc_fit_table(const double* y, int32_t n, const orc_cfg* cfg, orc_row* rows, int32_t* labels) {
    const double eps = DBL_EPSILON, tol = 1e-5; /* emControl() defaults */
    g_triples = 0.0; g_point_iters = 0.0;
    double* xs = (double*)malloc(sizeof(double) * n);
    memcpy(xs, y, sizeof(double) * n);
    qsort(xs, n, sizeof(double), cmp_double);
    double* z = (double*)malloc(sizeof(double) * (size_t)n * ORC_GMAX);
    int* cl = (int*)malloc(sizeof(int) * n);
    const double logn = log((double)n);

    for (int G = 1; G <= ORC_GMAX; G++) {
        orc_row* r = &rows[G - 1];
        memset(r, 0, sizeof(*r));
        r->G = G;
        int32_t* lab = labels ? labels + (size_t)(G - 1) * n : NULL;
        if (G > n) { r->status = -9; continue; } /* mclustBIC: G <- G[G <= n] */
        if (G == 1) {
            double mu, ss;
            double ll = mvn1d(y, n, &mu, &ss);
            r->iters = 0;
            r->mean[0] = r->mean_em[0] = mu;
            r->sigsq[0] = r->sigsq_em[0] = ss;
            r->pro[0] = r->pro_em[0] = 1.0;
            if (ll > 1.79769e308) { r->status = -1; continue; } /* mvnX: sigma-squared vanishes -> NA */
            r->valid = 1; r->loglik = ll;
            r->bic = 2.0 * ll - 2.0 * logn;
            if (lab) for (int i = 0; i < n; i++) lab[i] = 1;
            continue;
        }
        if (cfg->qclass_old) qclass_old(y, xs, n, G, cl); else qclass_new(y, xs, n, G, cl);
        /* unmap(cl, groups = 1:G) */
        for (int k = 0; k < G; k++)
            for (int i = 0; i < n; i++) z[(size_t)k * n + i] = (cl[i] == k + 1) ? 1.0 : 0.0;
        int iters = 0;
        double ll = me1v(y, n, G, z, tol, eps, 2147483647, r->mean_em, r->sigsq_em, r->pro_em, &iters);
        r->iters = iters;
        for (int k = 0; k < G; k++) { r->mean[k] = r->mean_em[k]; r->sigsq[k] = r->sigsq_em[k]; r->pro[k] = r->pro_em[k]; }
        int bad_sig = 0;
        for (int k = 0; k < G; k++) if (r->sigsq_em[k] <= eps) bad_sig = 1;
        if (ll > 1.79769e308 || bad_sig) { r->status = -1; continue; }      /* meV: sigma-squared falls below threshold */
        if (ll < -1.79769e308) { r->status = -3; continue; }                 /* meV: mixing proportion fell below threshold */
        r->valid = 1; r->loglik = ll;
        r->bic = 2.0 * ll - (double)(3 * G - 1) * logn;
        /* summaryMclustBIC: if sum((pro - colMeans(z))^2) > sqrt(eps): extra M-step replaces $parameters */
        if (cfg->extra_mstep) {
            double dev = 0.0;
            for (int k = 0; k < G; k++) {
                long double s = 0.0L;
                for (int i = 0; i < n; i++) s += z[(size_t)k * n + i];
                double cm = (double)(s / n);
                dev += (r->pro_em[k] - cm) * (r->pro_em[k] - cm);
            }
            if (dev > sqrt(DBL_EPSILON)) {
                double mu2[ORC_GMAX], ss2[ORC_GMAX], pr2[ORC_GMAX];
                if (ms1v(y, z, n, G, mu2, ss2, pr2)) {
                    for (int k = 0; k < G; k++) { r->mean[k] = mu2[k]; r->sigsq[k] = ss2[k]; r->pro[k] = pr2[k]; }
                    r->extra_applied = 1;
                }
            }
        }
        /* map(z): first index of the row maximum */
        if (lab) {
            for (int i = 0; i < n; i++) {
                double best = z[i]; int bk = 0;
                for (int k = 1; k < G; k++) { double v = z[(size_t)k * n + i]; if (v > best) { best = v; bk = k; } }
                lab[i] = bk + 1;
            }
        }
    }
    free(xs); free(z); free(cl);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* mclustRestricted   R/mclust.restricted.R:46-253                            */
/* ------------------------------------------------------------------------- */

typedef struct {
    const orc_row* rows;
    const int32_t* labels; /* [5*n] */
    int n;
} table_t;

/* Mclust(y, modelNames="V", G=1:k): pickBIC over valid rows 1..k, ties -> smaller G.  returns G or 0 (NULL) */
static int mclust_upto(const table_t* T, int k) {
    int best = 0;
    for (int g = 1; g <= k && g <= ORC_GMAX; g++) {
        if (!T->rows[g - 1].valid) continue;
        if (best == 0 || T->rows[g - 1].bic > T->rows[best - 1].bic) best = g;
    }
    return best;
}

/* table(cl): counts of present labels in increasing label order. returns number of present labels */
static int r_table(const int32_t* cl, int n, int* cnt, int* lab) {
    int c[ORC_GMAX + 1] = {0};
    for (int i = 0; i < n; i++) c[cl[i]]++;
    int m = 0;
    for (int g = 1; g <= ORC_GMAX; g++) if (c[g] > 0) { cnt[m] = c[g]; lab[m] = g; m++; }
    return m;
}

static int which_max(const double* v, int n) { int b = 0; for (int i = 1; i < n; i++) if (v[i] > v[b]) b = i; return b; }
static int which_min(const double* v, int n) { int b = 0; for (int i = 1; i < n; i++) if (v[i] < v[b]) b = i; return b; }

static double vmin(const double* v, int n) {
    /* R min(): NaN propagates */
    double m = v[0];
    for (int i = 0; i < n; i++) { if (isnan(v[i])) return NAN; if (v[i] < m) m = v[i]; }
    return m;
}

/* R/mclust.restricted.R:48-50: `if (length(unique(y))==1) y <- y + runif(length(y), -0.1, 0.1)`.
   runif(n, a, b) is `a + (b - a) * unif_rand()` per element (nmath/runif.c), drawn in element order from the
   stream that is current when mclustRestricted runs -- inside bplapply that is the element's own L'Ecuyer stream,
   so the draw sits mid-stream between two sample() calls of permMclustGene.  Returns 1 if the jitter was applied. */
int orc_jitter_if_constant(double* y, int32_t n, orc_rng* rng) {
    if (n < 1) return 0;                       /* length(unique(numeric(0))) == 0 */
    for (int i = 1; i < n; i++) if (y[i] != y[0]) return 0;
    const double a = -0.1, b = 0.1;
    for (int i = 0; i < n; i++) {
        double u = orc_unif_rand(rng);
        double w = (b - a) * u;
        y[i] = y[i] + (a + w);
    }
    return 1;
}

static int mclust_restricted_core(const double* y_in, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls);

/* mclustRestricted with the caller's RNG: the jittered copy is local to the function, as in R (the caller's y, which
   jointPosterior receives afterwards, keeps the identical values -- R/permutations.R:278-282). */
int orc_mclust_restricted_rng(const double* y_in, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls,
                              orc_rng* rng, int32_t* jittered) {
    if (jittered) *jittered = 0;
    if (!rng || n < 1) return mclust_restricted_core(y_in, n, cfg, out, cls);
    int constant = 1;
    for (int i = 1; i < n; i++) if (y_in[i] != y_in[0]) { constant = 0; break; }
    if (!constant) return mclust_restricted_core(y_in, n, cfg, out, cls);
    double* yj = (double*)malloc(sizeof(double) * (size_t)n);
    memcpy(yj, y_in, sizeof(double) * (size_t)n);
    orc_jitter_if_constant(yj, n, rng);
    if (jittered) *jittered = 1;
    int rc = mclust_restricted_core(yj, n, cfg, out, cls);
    free(yj);
    return rc;
}

/* without an RNG the jitter cannot be drawn: a constant vector then has no valid model (returns -1), which the batch
   drivers report; scDD() itself never gets here on observed data (skipConstant, R/scDD.R:292-310). */
int orc_mclust_restricted(const double* y_in, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls) {
    return mclust_restricted_core(y_in, n, cfg, out, cls);
}

static int mclust_restricted_core(const double* y_in, int32_t n, const orc_cfg* cfg, orc_fit* out, int32_t* cls) {
    const double* y = y_in;
    orc_row rows[ORC_GMAX];
    int32_t* labels = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * ORC_GMAX);
    orc_fit_table(y, n, cfg, rows, labels);
    table_t T = { rows, labels, n };
    int ncalls = 1;
    int ret = 0;

    int comps = mclust_upto(&T, 5);                          /* :52-55 */
    if (comps == 0) { free(labels); memset(out, 0, sizeof(*out)); return -1; }
    const int32_t* cl = labels + (size_t)(comps - 1) * n;
    int cnt[ORC_GMAX], lab[ORC_GMAX];
    int ntab = r_table(cl, n, cnt, lab);

    if (comps > ntab) {                                      /* :57-63 */
        comps = mclust_upto(&T, comps - 1); ncalls++;
        cl = labels + (size_t)(comps - 1) * n;
        ntab = r_table(cl, n, cnt, lab);
    }
    const orc_row* mc = &rows[comps - 1];

    double remThresh = 1.0 / (1.0 + exp(-(double)n / 60.0 + 0.4)) + 1.2;   /* :67 */
    double addThresh = 1.0 / (1.0 + exp((double)n / 9.0 - 4.5)) + 3.4;     /* :72 */

    if (cfg->restrict_) {
        if (comps > 1) {                                     /* :75-185 */
            double meandiff[ORC_GMAX]; int nmd = comps - 1;
            double sdmax = 0.0;
            for (int k = 0; k < comps; k++) { double s = sqrt(mc->sigsq[k]); if (s > sdmax) sdmax = s; }
            for (int j = 0; j < nmd; j++) meandiff[j] = (mc->mean[j + 1] - mc->mean[j]) / sdmax;   /* :77 */
            int maxcl = which_max(mc->sigsq, comps), mincl = which_min(mc->sigsq, comps);          /* :79-80 */
            double vardiff = mc->sigsq[maxcl] / mc->sigsq[mincl];                                  /* :81 */
            double nmin = (mincl < ntab) ? (double)cnt[mincl] : NAN;                               /* :83 positional */
            double mincat = cnt[0];
            for (int i = 1; i < ntab; i++) if (cnt[i] < mincat) mincat = cnt[i];                   /* :84 */
            int tries = 0;
            double nmax = (maxcl < ntab) ? (double)cnt[maxcl] : NAN;                               /* :86 */
            double balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));            /* :87 */
            for (int j = 0; j < nmd; j++) meandiff[j] *= balance;                                  /* :88 */
            /* :91-110 */
            if (isnan(vardiff)) vardiff = 1.0;
            if (isnan(nmin)) nmin = INFINITY;
            { int anyna = 0; for (int j = 0; j < nmd; j++) if (isnan(meandiff[j])) anyna = 1;
              if (nmd == 0 || anyna) { meandiff[0] = INFINITY; nmd = 1; } }
            int err = 0;
            while ((vmin(meandiff, nmd) < remThresh || vardiff > 20.0) && mincat > 2 && tries <= 4 && err == 0) { /* :115 */
                int comps_old = comps;
                comps = comps - 1;
                comps = mclust_upto(&T, comps); ncalls++;                                          /* :120-124 */
                mc = &rows[comps - 1];
                cl = labels + (size_t)(comps - 1) * n;
                ntab = r_table(cl, n, cnt, lab);
                mincat = cnt[0];
                for (int i = 1; i < ntab; i++) if (cnt[i] < mincat) mincat = cnt[i];               /* :125 */
                tries++;
                if (comps > 1 && comps_old - comps == 1) {                                         /* :128 */
                    /* :129 compmeans <- by(y, cl, mean): empirical means of the PRESENT labels */
                    double cm[ORC_GMAX];
                    double* tmp = (double*)malloc(sizeof(double) * n);
                    for (int t = 0; t < ntab; t++) {
                        int m = 0;
                        for (int i = 0; i < n; i++) if (cl[i] == lab[t]) tmp[m++] = y[i];
                        cm[t] = r_mean(tmp, m);
                    }
                    free(tmp);
                    sdmax = 0.0;
                    for (int k = 0; k < comps; k++) { double s = sqrt(mc->sigsq[k]); if (s > sdmax) sdmax = s; }
                    nmd = ntab - 1;
                    for (int j = 0; j < nmd; j++) meandiff[j] = (cm[j + 1] - cm[j]) / sdmax;       /* :130 */
                    balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));           /* :131 stale nmin/nmax */
                    for (int j = 0; j < nmd; j++) meandiff[j] *= balance;
                    maxcl = which_max(mc->sigsq, comps); mincl = which_min(mc->sigsq, comps);
                    vardiff = mc->sigsq[maxcl] / mc->sigsq[mincl];
                    nmin = (mincl < ntab) ? (double)cnt[mincl] : NAN;                              /* :138 */
                } else {
                    meandiff[0] = 10.0; nmd = 1; vardiff = -INFINITY; nmin = 100.0;                /* :140-142 */
                }
                /* :146-173 */
                if (isnan(vardiff)) { vardiff = 1.0; err = 1; }
                if (isnan(nmin)) { nmin = INFINITY; err = 1; }
                { int anyna = 0; for (int j = 0; j < nmd; j++) if (isnan(meandiff[j])) anyna = 1;
                  if (nmd == 0 || anyna) { meandiff[0] = INFINITY; nmd = 1; err = 1; } }
                if (err == 1) {                                                                    /* :176-183 */
                    comps = comps + 1;
                    comps = mclust_upto(&T, comps); ncalls++;
                    mc = &rows[comps - 1];
                    cl = labels + (size_t)(comps - 1) * n;
                    ntab = r_table(cl, n, cnt, lab);
                }
            }
        } else {                                             /* :186-225 try to add a component */
            ncalls++;
            int keep2 = 0;
            if (n >= 2 && rows[1].valid) {
                const int32_t* cl2 = labels + (size_t)1 * n;
                int c2[ORC_GMAX], l2[ORC_GMAX];
                int nt2 = r_table(cl2, n, c2, l2);
                if (nt2 == 2) {                                                                    /* :195 */
                    const orc_row* m2 = &rows[1];
                    double mincat = c2[0] < c2[1] ? c2[0] : c2[1];
                    double sdmax = fmax(sqrt(m2->sigsq[0]), sqrt(m2->sigsq[1]));
                    double md = (m2->mean[1] - m2->mean[0]) / sdmax;
                    int maxcl = which_max(m2->sigsq, 2), mincl = which_min(m2->sigsq, 2);
                    double vardiff = m2->sigsq[maxcl] / m2->sigsq[mincl];
                    double nmin = c2[mincl], nmax = c2[maxcl];
                    double balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));
                    md *= balance;
                    if ((md > addThresh && vardiff < 10.0) && mincat > 2) keep2 = 1;               /* :211 */
                }
            }
            if (keep2) comps = 2;
            else { comps = 1; ncalls++; }   /* G=comps_old refit at :212 / :219 */
            mc = &rows[comps - 1];
            cl = labels + (size_t)(comps - 1) * n;
            ntab = r_table(cl, n, cnt, lab);
        }
    }

    /* :230-237 all clusters smaller than min.size -> merge */
    {
        int mx = 0;
        for (int i = 0; i < ntab; i++) if (cnt[i] > mx) mx = cnt[i];
        if (comps > 1 && mx < cfg->min_size) {
            comps = mclust_upto(&T, comps - 1); ncalls++;
            mc = &rows[comps - 1];
            cl = labels + (size_t)(comps - 1) * n;
            ntab = r_table(cl, n, cnt, lab);
        }
    }

    /* :240-252 relabel with order() (quirk: order, not rank) and return */
    int ord[ORC_GMAX];
    for (int k = 0; k < comps; k++) ord[k] = k;
    if (comps > 1) {
        for (int i = 1; i < comps; i++) { /* stable order(mean) */
            int o = ord[i]; int j = i - 1;
            while (j >= 0 && mc->mean[ord[j]] > mc->mean[o]) { ord[j + 1] = ord[j]; j--; }
            ord[j + 1] = o;
        }
    }
    for (int i = 0; i < n; i++) cls[i] = (comps > 1) ? ord[cl[i] - 1] + 1 : cl[i];
    memset(out, 0, sizeof(*out));
    out->G = comps;
    out->model_v = comps > 1;
    out->n = n;
    out->df = comps > 1 ? 3 * comps - 1 : 2;
    out->bic = mc->bic;
    out->loglik = mc->loglik;
    for (int k = 0; k < comps; k++) { out->mean[k] = mc->mean[ord[k]]; out->var[k] = mc->sigsq[ord[k]]; }
    out->n_mclust_calls = ncalls;
    free(labels);
    return ret;
}

/* ------------------------------------------------------------------------- */
/* jointPosterior   R/joint.posterior.R:35-64                                 */
/* ------------------------------------------------------------------------- */

double orc_joint_posterior(const double* y, const int32_t* cls, int32_t n,
                           double alpha, double m0, double s0, double a0, double b0) {
    int present[ORC_GMAX + 2] = {0};
    for (int i = 0; i < n; i++) if (cls[i] >= 0 && cls[i] <= ORC_GMAX) present[cls[i]] = 1;
    int r = 0;
    for (int g = 0; g <= ORC_GMAX; g++) r += present[g];     /* :37 length(unique(class)) */
    double logpost = 0.0;
    double J = 0.0;
    double nk[ORC_GMAX + 1], sk[ORC_GMAX + 1], ak[ORC_GMAX + 1], bk[ORC_GMAX + 1];
    for (int k = 1; k <= r; k++) {                            /* :41-45; R sum() accumulates in long double */
        long double s1 = 0.0L, s2 = 0.0L; int c = 0;
        for (int i = 0; i < n; i++) if (cls[i] == k) { c++; s1 += y[i]; s2 += y[i] * y[i]; }
        double sumyk = (double)s1, sumyk2 = (double)s2;
        nk[k] = c;
        sk[k] = s0 + c;                                       /* :47 */
        double mk = (s0 * m0 + sumyk) / sk[k];                /* :48 */
        ak[k] = a0 + c;                                       /* :49 */
        bk[k] = b0 + sumyk2 + s0 * m0 * m0 - sk[k] * mk * mk; /* :50 */
        J += c;
    }
    for (int k = 1; k <= r; k++) {                            /* :55-58 */
        double lg_nk = (nk[k] == 0.0) ? INFINITY : lgamma(nk[k]);
        logpost = logpost + lg_nk + lgamma(ak[k] / 2.0) - (ak[k] / 2.0) * log(bk[k] / 2.0) - 0.5 * log(sk[k]);
    }
    logpost = logpost + r * log(alpha) + lgamma(alpha) - J * lgamma(a0 / 2.0) +
              (J * a0 / 2.0) * log(b0 / 2.0) + J * 0.5 * log(s0) - (J / 2.0) * log(2.0 * M_PI) - lgamma(alpha + J); /* :59-60 */
    return logpost;
}

/* ------------------------------------------------------------------------- */
/* stats::ks.test(x, y, exact = FALSE)                         */
/* ------------------------------------------------------------------------- */

typedef struct { double v; int idx; } kv_t;
static int cmp_kv(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a; const kv_t* y = (const kv_t*)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* stable order() */
}

static double pkolmogorov_two_limit_upper(double x, double tol) {
    /* R >= 4.2 ks.c pkolmogorov_two_limit(lower.tail = FALSE) */
    if (x <= 0) return 1.0;
    int k_max = (int)sqrt(2.0 - log(tol));
    if (x < 1.0) {
        double z = -(M_PI_2 * M_PI_4) / (x * x), w = log(x), s = 0.0;
        for (int k = 1; k < k_max; k += 2) s += exp(k * k * z - w);
        return 1.0 - s / 0.398942280401432677939946059934; /* M_1_SQRT_2PI */
    }
    double z = -2.0 * x * x, s = -1.0, old = 0.0, nw = 2.0 * exp(z);
    int k = 2;
    while (fabs(old - nw) > tol) { old = nw; nw += 2.0 * s * exp(z * k * k); s *= -1.0; k++; }
    return nw;
}

static double pks2_lower(double x, double tol) {
    /* R < 4.2 ks.c pKS2: limiting cdf; the caller forms 1 - p */
    if (x <= 0) return 0.0;
    int k_max = (int)sqrt(2.0 - log(tol));
    if (x < 1.0) {
        double z = -(M_PI_2 * M_PI_4) / (x * x), w = log(x), s = 0.0;
        for (int k = 1; k < k_max; k += 2) s += exp(k * k * z - w);
        return s / 0.398942280401432677939946059934;
    }
    double z = -2.0 * x * x, s = -1.0, old = 0.0, nw = 1.0;
    int k = 1;
    while (fabs(old - nw) > tol) { old = nw; nw += 2.0 * s * exp(z * k * k); s *= -1.0; k++; }
    return nw;
}

int orc_ks(const double* x, int32_t nx, const double* y, int32_t ny, int32_t old_pvalue, double* D, double* p) {
    int N = nx + ny;
    if (nx < 1 || ny < 1) { *D = NAN; *p = NAN; return -1; } /* ks.test: "not enough 'x'/'y' data" */
    kv_t* w = (kv_t*)malloc(sizeof(kv_t) * N);
    for (int i = 0; i < nx; i++) { w[i].v = x[i]; w[i].idx = i; }
    for (int i = 0; i < ny; i++) { w[nx + i].v = y[i]; w[nx + i].idx = nx + i; }
    qsort(w, N, sizeof(kv_t), cmp_kv);
    double dnx = (double)nx, dny = (double)ny;
    double cx = 1.0 / dnx, cy = -1.0 / dny;
    long double run = 0.0L; /* cumsum accumulates in LDOUBLE, stores double */
    double dmax = 0.0;
    for (int i = 0; i < N; i++) {
        run += (w[i].idx < nx) ? cx : cy;
        double z = (double)run;
        /* with ties keep only the last position of each tie run */
        if (i == N - 1 || w[i + 1].v != w[i].v) { double a = fabs(z); if (a > dmax) dmax = a; }
    }
    free(w);
    *D = dmax;
    double n = dnx * dny / (dnx + dny);
    double pv;
    if (old_pvalue) pv = 1.0 - pks2_lower(sqrt(n) * dmax, 1e-6);
    else pv = pkolmogorov_two_limit_upper(sqrt(n) * dmax, 1e-6);
    if (pv > 1.0) pv = 1.0;
    if (pv < 0.0) pv = 0.0;
    *p = pv;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* R RNG: set.seed, Mersenne-Twister, L'Ecuyer-CMRG, R_unif_index, sample      */
/* (R src/main/RNG.c, src/main/random.c)                        */
/* ------------------------------------------------------------------------- */

#define MT_N 624
#define MT_M 397
#define LEC_M1 4294967087LL
#define LEC_M2 4294944443LL

void orc_rng_set_seed(orc_rng* r, int32_t kind, uint32_t seed) {
    /* RNG_Init: 50 LCG scrambles, then fill i_seed[]; FixupSeeds */
    r->kind = kind;
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    if (kind == 0) {
        /* i_seed[0] is dummy[0] = mti, i_seed[1..624] = mt[] ; FixupSeeds: dummy[0] = 624 */
        seed = 69069u * seed + 1u; /* i_seed[0], overwritten */
        for (int j = 0; j < MT_N; j++) { seed = 69069u * seed + 1u; r->mt[j] = seed; }
        r->mti = MT_N;
    } else {
        for (int j = 0; j < 6; j++) { seed = 69069u * seed + 1u; r->lec[j] = seed; }
        /* FixupSeeds: seeds must be < m1 (first three) / m2 (last three), not all zero */
        for (int j = 0; j < 3; j++) if (r->lec[j] >= (uint32_t)LEC_M1) r->lec[j] = 0; /* R: checks & resets via Randomize; kept simple */
        for (int j = 3; j < 6; j++) if (r->lec[j] >= (uint32_t)LEC_M2) r->lec[j] = 0;
    }
}

void orc_rng_set_lecuyer(orc_rng* r, const uint32_t* seed6) {
    r->kind = 1;
    for (int j = 0; j < 6; j++) r->lec[j] = seed6[j];
}

static double fixup(double x) {
    const double i2_32m1 = 2.328306437080797e-10;
    if (x <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
}

static double mt_genrand(orc_rng* r) {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t y;
    uint32_t* mt = r->mt;
    if (r->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 0x1];
        }
        for (; kk < MT_N - 1; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 0x1];
        }
        y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 0x1];
        r->mti = 0;
    }
    y = mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return (double)y * 2.3283064365386963e-10;
}

static double lecuyer_rand(orc_rng* r) {
    uint32_t* s = r->lec;
    int64_t p1 = 1403580LL * (int64_t)s[1] - 810728LL * (int64_t)s[0];
    int64_t k = p1 / LEC_M1;
    p1 -= k * LEC_M1;
    if (p1 < 0) p1 += LEC_M1;
    s[0] = s[1]; s[1] = s[2]; s[2] = (uint32_t)p1;
    int64_t p2 = 527612LL * (int64_t)s[5] - 1370589LL * (int64_t)s[3];
    k = p2 / LEC_M2;
    p2 -= k * LEC_M2;
    if (p2 < 0) p2 += LEC_M2;
    s[3] = s[4]; s[4] = s[5]; s[5] = (uint32_t)p2;
    return (double)((p1 > p2) ? (p1 - p2) : (p1 - p2 + LEC_M1)) * 2.328306549295727688e-10;
}

double orc_unif_rand(orc_rng* r) {
    return fixup(r->kind == 0 ? mt_genrand(r) : lecuyer_rand(r));
}

static double rbits(orc_rng* r, int bits) {
    int64_t v = 0;
    for (int n = 0; n <= bits; n += 16) {
        int v1 = (int)floor(orc_unif_rand(r) * 65536.0);
        v = 65536 * v + v1;
    }
    if (bits < 64) v &= (((int64_t)1 << bits) - 1);
    return (double)v;
}

double orc_unif_index(orc_rng* r, double dn) {
    if (dn <= 0) return 0.0;
    int bits = (int)ceil(log2(dn));
    double dv;
    do { dv = rbits(r, bits); } while (dn <= dv);
    return dv;
}

void orc_sample_perm(orc_rng* r, int32_t n, int32_t* out) {
    /* do_sample, replace = FALSE, k = n (sample(1:n) -> x[sample.int(n)]) */
    int32_t* x = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) x[i] = i;
    int m = n;
    for (int i = 0; i < n; i++) {
        int j = (int)orc_unif_index(r, (double)m);
        out[i] = x[j] + 1;
        x[j] = x[--m];
    }
    free(x);
}

/* MRG32k3a stream jumps: parallel::nextRNGStream (2^127) / nextRNGSubStream (2^76) */
static void mat3_mulmod(const uint64_t A[3][3], const uint64_t B[3][3], uint64_t C[3][3], uint64_t m) {
    uint64_t T[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            unsigned __int128 s = 0;
            for (int k = 0; k < 3; k++) s += (unsigned __int128)A[i][k] * B[k][j];
            T[i][j] = (uint64_t)(s % m);
        }
    memcpy(C, T, sizeof(T));
}

static void lecuyer_jump(const uint32_t* in6, uint32_t* out6, int log2jump) {
    uint64_t A1[3][3] = {{0, 1, 0}, {0, 0, 1}, {(uint64_t)(LEC_M1 - 810728LL), 1403580ULL, 0}};
    uint64_t A2[3][3] = {{0, 1, 0}, {0, 0, 1}, {(uint64_t)(LEC_M2 - 1370589LL), 0, 527612ULL}};
    for (int i = 0; i < log2jump; i++) { mat3_mulmod(A1, A1, A1, LEC_M1); mat3_mulmod(A2, A2, A2, LEC_M2); }
    for (int i = 0; i < 3; i++) {
        unsigned __int128 s1 = 0, s2 = 0;
        for (int k = 0; k < 3; k++) { s1 += (unsigned __int128)A1[i][k] * in6[k]; s2 += (unsigned __int128)A2[i][k] * in6[3 + k]; }
        out6[i] = (uint32_t)(s1 % LEC_M1);
        out6[3 + i] = (uint32_t)(s2 % LEC_M2);
    }
}

void orc_lecuyer_next_stream(const uint32_t* in6, uint32_t* out6) { uint32_t t[6]; lecuyer_jump(in6, t, 127); memcpy(out6, t, sizeof(t)); }
void orc_lecuyer_next_substream(const uint32_t* in6, uint32_t* out6) { uint32_t t[6]; lecuyer_jump(in6, t, 76); memcpy(out6, t, sizeof(t)); }

/* ------------------------------------------------------------------------- */
/* batch drivers                                                             */
/* ------------------------------------------------------------------------- */

static int set_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    return nthreads;
#else
    (void)nthreads; return 1;
#endif
}

/* genefit, R/scDD.R:373-391 */
int orc_genefit_batch(const double* logy, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                      const orc_cfg* cfg, orc_fit* fits, int32_t* cls_oa, int32_t* cls_c,
                      double* bf, double* den, int32_t nthreads) {
    int nt = set_threads(nthreads);
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int g = 0; g < ngenes; g++) {
        int64_t o = gene_off[g];
        int n = (int)(gene_off[g + 1] - o);
        const double* y = logy + o;
        const int32_t* c = cond + o;
        double* y1 = (double*)malloc(sizeof(double) * (n + 1));
        double* y2 = (double*)malloc(sizeof(double) * (n + 1));
        int32_t* k1 = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int32_t* k2 = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int n1 = 0, n2 = 0;
        for (int i = 0; i < n; i++) { if (c[i] == 0) y1[n1++] = y[i]; else y2[n2++] = y[i]; }
        int r0 = orc_mclust_restricted(y, n, cfg, &fits[3 * g + 0], cls_oa + o);      /* :377 */
        int r1 = orc_mclust_restricted(y1, n1, cfg, &fits[3 * g + 1], k1);            /* :378 */
        int r2 = orc_mclust_restricted(y2, n2, cfg, &fits[3 * g + 2], k2);            /* :379 */
        if (r0 || r1 || r2) {
#pragma omp atomic write
            rc = -1;
        }
        if (bf) bf[g] = orc_joint_posterior(y1, k1, n1, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0) +
                        orc_joint_posterior(y2, k2, n2, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0); /* :381-382 */
        if (den) den[g] = orc_joint_posterior(y, cls_oa + o, n, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0); /* :383 */
        int a = 0, b = 0;
        for (int i = 0; i < n; i++) cls_c[o + i] = (c[i] == 0) ? k1[a++] : k2[b++];
        free(y1); free(y2); free(k1); free(k2);
    }
    return rc;
}

/* permMclustGene (R/permutations.R:211-286), permMclust (:147-181), permMclustCov (:44-104).
   seed_mode 0: seed6 is [ngenes*6], gene g draws all its B permutations from its own stream (parallelBy = "Genes":
                bplapply over genes, R/scDD.R:458).
   seed_mode 1: seed6 is [B*6], permutation b of EVERY gene is drawn from the start of stream b (parallelBy =
                "Permutations": scDD loops over the genes in the master and each permMclust / permMclustCov call runs
                bplapply(1:nperms, ...) under the same RNGseed, R/permutations.R:99,177).
   In both modes a permuted subset with identical values draws its runif jitter from the stream it runs under
   (orc_mclust_restricted_rng); with host-drawn permutations (perm_idx) no stream exists and such a set fails.
   n_jittered (nullable) counts the jittered fit-sets. */
static int perm_batch_any(const double* logy, const int32_t* cond, const double* cdr, const int64_t* gene_off,
                          int32_t ngenes, int32_t B, const int32_t* perm_idx, const uint32_t* seed6,
                          int32_t seed_mode, const orc_cfg* cfg, double* bf_perm, int32_t nthreads,
                          int64_t* n_jittered, double* jp_terms) {
    int nt = set_threads(nthreads);
    int rc = 0;
    long long njit = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(+ : njit)
    for (int g = 0; g < ngenes; g++) {
        int64_t o = gene_off[g];
        int n = (int)(gene_off[g + 1] - o);
        const double* y = logy + o;
        const int32_t* c = cond + o;
        double* fit = NULL; double* res = NULL;
        if (cdr) {
            /* lm(y ~ C): least squares with intercept (R uses Householder QR; same estimate up to rounding) */
            const double* C = cdr + o;
            long double sx = 0, sy = 0;
            for (int i = 0; i < n; i++) { sx += C[i]; sy += y[i]; }
            double mx = (double)(sx / n), my = (double)(sy / n);
            long double sxx = 0, sxy = 0;
            for (int i = 0; i < n; i++) { sxx += (C[i] - mx) * (C[i] - mx); sxy += (C[i] - mx) * (y[i] - my); }
            double b1 = (double)(sxy / sxx), b0 = my - b1 * mx;
            fit = (double*)malloc(sizeof(double) * (n + 1));
            res = (double*)malloc(sizeof(double) * (n + 1));
            for (int i = 0; i < n; i++) { fit[i] = b0 + b1 * C[i]; res[i] = y[i] - fit[i]; }
        }
        double* ya = (double*)malloc(sizeof(double) * (n + 1));
        double* y1 = (double*)malloc(sizeof(double) * (n + 1));
        double* y2 = (double*)malloc(sizeof(double) * (n + 1));
        int32_t* k0 = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int32_t* k1 = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int32_t* k2 = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int32_t* nw = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        orc_rng rng;
        if (seed6 && seed_mode == 0) orc_rng_set_lecuyer(&rng, seed6 + 6 * (size_t)g);
        orc_fit f0, f1, f2;
        for (int b = 0; b < B; b++) {
            const int32_t* pi;
            if (seed6 && seed_mode == 1) orc_rng_set_lecuyer(&rng, seed6 + 6 * (size_t)b);
            if (perm_idx) pi = perm_idx + ((size_t)o * B + (size_t)b * n);
            else { orc_sample_perm(&rng, n, nw); pi = nw; }             /* :273 sample(1:length(y)) / :243 sample(residuals) */
            int n1 = 0, n2 = 0;
            for (int i = 0; i < n; i++) {                                /* :274-276 / :244-253 */
                double v = cdr ? log(exp(fit[i] + res[pi[i] - 1])) : y[pi[i] - 1];
                if (c[i] == 0) y1[n1++] = v; else y2[n2++] = v;
            }
            orc_rng* rp = perm_idx ? NULL : &rng;
            int32_t j0 = 0, j1 = 0, j2 = 0;
            int r0 = 0;
            double jp0 = 0.0;
            if (cdr) {
                memcpy(ya, y1, sizeof(double) * n1);                     /* c(y1, y2) :261 */
                memcpy(ya + n1, y2, sizeof(double) * n2);
                r0 = orc_mclust_restricted_rng(ya, n, cfg, &f0, k0, rp, &j0);                 /* :261 (oa first) */
            }
            int r1 = orc_mclust_restricted_rng(y1, n1, cfg, &f1, k1, rp, &j1);                /* :278 / :262 */
            int r2 = orc_mclust_restricted_rng(y2, n2, cfg, &f2, k2, rp, &j2);                /* :279 / :263 */
            njit += j0 + j1 + j2;
            if (r0 || r1 || r2) {
#pragma omp atomic write
                rc = -1;
            }
            if (cdr) jp0 = orc_joint_posterior(ya, k0, n, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0);
            double jp1 = orc_joint_posterior(y1, k1, n1, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0);
            double jp2 = orc_joint_posterior(y2, k2, n2, cfg->alpha, cfg->m0, cfg->s0, cfg->a0, cfg->b0);
            double v = jp1 + jp2;                                                                 /* :281-282 */
            if (cdr) v = v - jp0;                                                                 /* :265-267 */
            if (r0 || r1 || r2) v = NAN;
            if (jp_terms) {
                const int sets = cdr ? 3 : 2;
                double* t = jp_terms + ((size_t)g * B + b) * sets;
                t[0] = r1 ? NAN : jp1; t[1] = r2 ? NAN : jp2;
                if (cdr) t[2] = r0 ? NAN : jp0;
            }
            bf_perm[(size_t)g * B + b] = v;
        }
        free(fit); free(res); free(ya); free(y1); free(y2); free(k0); free(k1); free(k2); free(nw);
    }
    if (n_jittered) *n_jittered = njit;
    return rc;
}

int orc_perm_batch(const double* logy, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                   int32_t B, const int32_t* perm_idx, const uint32_t* seed6,
                   const orc_cfg* cfg, double* bf_perm, int32_t nthreads) {
    return perm_batch_any(logy, cond, NULL, gene_off, ngenes, B, perm_idx, seed6, 0, cfg, bf_perm, nthreads, NULL, NULL);
}

int orc_perm_batch_cov(const double* logy, const int32_t* cond, const double* cdr, const int64_t* gene_off,
                       int32_t ngenes, int32_t B, const int32_t* perm_idx, const uint32_t* seed6,
                       const orc_cfg* cfg, double* bf_perm, int32_t nthreads) {
    return perm_batch_any(logy, cond, cdr, gene_off, ngenes, B, perm_idx, seed6, 0, cfg, bf_perm, nthreads, NULL, NULL);
}

int orc_perm_batch_ex(const double* logy, const int32_t* cond, const double* cdr, const int64_t* gene_off,
                      int32_t ngenes, int32_t B, const int32_t* perm_idx, const uint32_t* seed6, int32_t seed_mode,
                      const orc_cfg* cfg, double* bf_perm, int32_t nthreads, int64_t* n_jittered, double* jp_terms) {
    return perm_batch_any(logy, cond, cdr, gene_off, ngenes, B, perm_idx, seed6, seed_mode, cfg, bf_perm, nthreads,
                          n_jittered, jp_terms);
}

/* testKS onegene, R/Test_KS.R:56-65 (inclZero = FALSE form: callers pass nonzero raw values) */
int orc_ks_batch(const double* vals, const int32_t* cond, const int64_t* gene_off, int32_t ngenes,
                 int32_t old_pvalue, double* D, double* p, int32_t nthreads) {
    int nt = set_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int g = 0; g < ngenes; g++) {
        int64_t o = gene_off[g];
        int n = (int)(gene_off[g + 1] - o);
        double* x1 = (double*)malloc(sizeof(double) * (n + 1));
        double* x2 = (double*)malloc(sizeof(double) * (n + 1));
        int n1 = 0, n2 = 0;
        for (int i = 0; i < n; i++) { if (cond[o + i] == 0) x1[n1++] = vals[o + i]; else x2[n2++] = vals[o + i]; }
        orc_ks(x1, n1, x2, n2, old_pvalue, &D[g], &p[g]);
        free(x1); free(x2);
    }
    return 0;
}
