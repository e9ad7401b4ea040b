// fitset.cuh -- one "fit-set" (= one mclustRestricted call of the reference) executed by one warp.
//
// A fit-set is: the G = 1..5 table of mclust's 1-D unequal-variance ("V") Gaussian-mixture fits of one
// vector (qclass quantile initialisation -> EM with M-step first -> BIC), followed by
// scDD's restricted selection (R/mclust.restricted.R:46-253) and, for the Bayes factor, the
// normal-gamma / product-partition score of the selected partition (R/joint.posterior.R:35-64).
// Every Mclust(y, G = ...) call the reference makes inside mclustRestricted re-reads this one table
// (the fits are deterministic in y), so the warp computes the five rows once and then runs the
// reference's scalar decision procedure on them.
//
// Data placement: the vector lives in the warp's shared-memory slab, SORTED (the EM sums and the
// selection are order independent; sorting makes mclust's quantile initialisation a lookup).
// Lanes own points i = lane, lane+32, ...; sufficient statistics are reduced with warp shuffles.
// All arithmetic is fp64.
#pragma once
#include <cfloat>
#include <cstdint>

#include "device_math.cuh"

namespace scdd {

constexpr int GMAX = 5;
constexpr unsigned FULLMASK = 0xffffffffu;
constexpr double PI2LOG = 1.837877066409345;   // log(2*pi), mclust.f
constexpr double SMALOG = -708.0;              // mclust.f: exp() arguments below this give z = 0
constexpr int NQCAP = 64;                      // qclass: give up growing the quantile grid beyond this
constexpr int EM_ITER_CAP = 100000;            // safety cap (mclust: itmax = .Machine$integer.max)

struct DevCfg {
    double alpha, m0, s0, a0, b0;
    int min_size, restrict_, extra_mstep, pad;
};

// one row of the G = 1..5 table (per warp, shared memory)
struct Row {
    double bic, loglik;
    double mean[GMAX], sigsq[GMAX];                  // $parameters (after summaryMclustBIC's extra M-step)
    double mu_em[GMAX], nh_em[GMAX], lc_em[GMAX];    // EM-final parameters in E-step form (label passes)
    double cnt[GMAX], sy[GMAX], syy[GMAX];           // per MAP label: count, sum y, sum y^2
    int valid, iters;
};

struct WarpStats {
    double triples, point_iters, em_iters;
    double cap_hits, tol_near, bic_near, const_sets, failed_sets, fit_sets;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void warp_sum_vec(double (&v)[NV]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int j = 0; j < NV; j++) v[j] += __shfl_xor_sync(FULLMASK, v[j], o);
    }
}

// ---------------------------------------------------------------------------------------------
// in-place ascending bitonic sort of np (power of two) doubles in shared memory by one warp
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_sort(double* xs, int np, int lane) {
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (np >> 1); t += 32) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = ((i & k) == 0);
                double a = xs[i], b = xs[l];
                if ((a > b) == up) { xs[i] = b; xs[l] = a; }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// stats::quantile(type = 7) on the sorted slab -- the arithmetic order of R's quantile.default
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double quantile7(const double* xs, int n, double p) {
    double index = __dadd_rn(1.0, __dmul_rn((double)(n - 1), p));
    double lo = floor(index), hi = ceil(index);
    double qs = xs[(int)lo - 1];
    double xh = xs[(int)hi - 1];
    if (index > lo && xh != qs) {
        double h = __dsub_rn(index, lo);
        qs = __dadd_rn(__dmul_rn(__dsub_rn(1.0, h), qs), __dmul_rn(h, xh));
    }
    return qs;
}

// mclust::qclass(x, k): interior class breaks.  Class of x = 1 + #{c : x >= br[c]}.
// Returns the number of interior breaks (k-1 when k classes could be formed; fewer => some class is
// empty and the EM start fails exactly as in mclust).
__device__ __noinline__ int qclass_breaks(const double* xs, int n, int k, double* br) {
    // common case: the k+1 quantiles of seq(0, 1, length = k+1) are distinct
    {
        double q[GMAX + 1];
        double by = 1.0 / (double)k;
        bool distinct = true;
        for (int j = 0; j <= k; j++) {
            double p = (j == 0) ? 0.0 : (j == k ? 1.0 : __dmul_rn((double)j, by));
            q[j] = quantile7(xs, n, p);
            if (j > 0 && !(q[j] > q[j - 1])) distinct = false;
        }
        if (distinct) {
            for (int c = 1; c < k; c++) br[c - 1] = q[c];
            return k - 1;
        }
    }
    // ties among the quantiles: grow the grid until k+1 distinct values exist (unique()), then drop the
    // breaks with the smallest gaps (q[-order(diff(q))[1:nr]])
    double u[NQCAP + 1];
    int L = 0;
    int nq = k;
    while (L < k + 1) {
        nq += 1;
        if (nq > NQCAP) break;
        double by = 1.0 / (double)(nq - 1);
        L = 0;
        for (int j = 0; j < nq; j++) {
            double p = (j == 0) ? 0.0 : (j == nq - 1 ? 1.0 : __dmul_rn((double)j, by));
            double v = quantile7(xs, n, p);
            bool dup = false;
            for (int m = 0; m < L; m++) if (u[m] == v) { dup = true; break; }
            if (!dup) u[L++] = v;
        }
    }
    if (L > k + 1) {
        int nr = L - k - 1;
        // remove, nr times, the (stable) smallest gap of the ORIGINAL diff vector
        bool drop[NQCAP + 1];
        for (int i = 0; i < L; i++) drop[i] = false;
        bool used[NQCAP];
        for (int i = 0; i < L - 1; i++) used[i] = false;
        for (int r = 0; r < nr; r++) {
            int best = -1;
            double bd = 0.0;
            for (int i = 0; i < L - 1; i++) {
                if (used[i]) continue;
                double d = __dsub_rn(u[i + 1], u[i]);
                if (best < 0 || d < bd) { best = i; bd = d; }
            }
            used[best] = true;
            drop[best] = true;  // order(dq)[r] indexes q itself: q[-idx] drops the LEFT end of that gap
        }
        int w = 0;
        for (int i = 0; i < L; i++) if (!drop[i]) u[w++] = u[i];
        L = w;
    }
    int nbr = L - 2;
    if (nbr > k - 1) nbr = k - 1;
    if (nbr < 0) nbr = 0;
    for (int c = 0; c < nbr; c++) br[c] = u[c + 1];
    return nbr;
}

// ---------------------------------------------------------------------------------------------
// E-step for one point: unnormalised responsibilities e_k = exp(t_k - max t), their sum S and the max.
// Written with explicit rounding intrinsics so every inlined copy performs identical arithmetic.
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void estep_point(double x, const double (&mu)[G], const double (&nh)[G],
                                            const double (&lc)[G], double (&e)[G], double (&d)[G],
                                            double (&q)[G], double& S, double& tmax) {
    double t[G];
#pragma unroll
    for (int k = 0; k < G; k++) {
        d[k] = __dsub_rn(x, mu[k]);
        q[k] = __dmul_rn(d[k], d[k]);
        t[k] = fma(q[k], nh[k], lc[k]);
    }
    tmax = t[0];
#pragma unroll
    for (int k = 1; k < G; k++) tmax = fmax(tmax, t[k]);
    S = 0.0;
#pragma unroll
    for (int k = 0; k < G; k++) {
        double a = __dsub_rn(t[k], tmax);
        double ek = exp_neg(a);
        ek = (a >= SMALOG) ? ek : 0.0;
        e[k] = ek;
        S = __dadd_rn(S, ek);
    }
}

// map(z): first index of the row maximum of z = e / S
template <int G>
__device__ __forceinline__ int map_label(double x, const double (&mu)[G], const double (&nh)[G],
                                         const double (&lc)[G]) {
    double e[G], d[G], q[G], S, tmax;
    estep_point<G>(x, mu, nh, lc, e, d, q, S, tmax);
    double rs = __drcp_rn(S);
    double best = __dmul_rn(e[0], rs);
    int bk = 0;
#pragma unroll
    for (int k = 1; k < G; k++) {
        double z = __dmul_rn(e[k], rs);
        if (z > best) { best = z; bk = k; }
    }
    return bk;
}

// ---------------------------------------------------------------------------------------------
// G >= 2 row: qclass start, EM (mclust me1v semantics: M-step first, stop when the relative change of
// the log-likelihood is <= 1e-5), BIC, extra M-step, MAP label statistics.
//
// M-step form: with d = x - mu_old and q = d^2 already needed by the E-step, the new parameters are
//   mu = mu_old + S(z d)/S(z),  sigsq = S(z q)/S(z) - (S(z d)/S(z))^2
// which equals mclust's two-pass update in exact arithmetic and needs no stored z matrix.
// ---------------------------------------------------------------------------------------------
template <int G>
__device__ __noinline__ void fit_row(const double* xs, int n, int lane, Row* row, WarpStats& st,
                                     int extra_mstep) {
    const double eps = DBL_EPSILON, rteps = 1.4901161193847656e-08, tol = 1e-5;
    const double dn = (double)n;
    if (lane == 0) { row->valid = 0; row->iters = 0; }
    if (G > n) return;   // mclustBIC: G <- G[G <= n]

    double br[GMAX];
    int nbr = qclass_breaks(xs, n, G, br);

    // ---- first M-step from the hard quantile classes (z = unmap(qclass)) ----
    double mu[G], sig[G], pro[G], nh[G], lc[G];
    {
        double acc[2 * G];
#pragma unroll
        for (int k = 0; k < 2 * G; k++) acc[k] = 0.0;
        for (int i = lane; i < n; i += 32) {
            double x = xs[i];
            int cl = 0;
            for (int c = 0; c < nbr; c++) cl += (x >= br[c]) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < G; k++) {
                if (k == cl) { acc[k] += 1.0; acc[G + k] += x; }
            }
        }
        warp_sum_vec<2 * G>(acc);
        double cmin = acc[0];
#pragma unroll
        for (int k = 1; k < G; k++) cmin = fmin(cmin, acc[k]);
        if (cmin <= rteps) return;   // an empty class: "mixing proportion fell below threshold"
        double ss[G];
#pragma unroll
        for (int k = 0; k < G; k++) { mu[k] = acc[G + k] / acc[k]; pro[k] = acc[k] / dn; ss[k] = 0.0; }
        for (int i = lane; i < n; i += 32) {
            double x = xs[i];
            int cl = 0;
            for (int c = 0; c < nbr; c++) cl += (x >= br[c]) ? 1 : 0;
#pragma unroll
            for (int k = 0; k < G; k++) {
                if (k == cl) { double t = x - mu[k]; ss[k] = fma(t, t, ss[k]); }
            }
        }
        warp_sum_vec<G>(ss);
#pragma unroll
        for (int k = 0; k < G; k++) sig[k] = ss[k] / acc[k];
    }

    double hold = 8.988465674311579e307, hood = 0.0;   // "previous" log-likelihood sentinel (FLMAX/2 in mclust)
    int iter = 0;
    double A[G], B[G], C[G];
    bool failed = false;
    for (;;) {
        double smin = sig[0];
#pragma unroll
        for (int k = 1; k < G; k++) smin = fmin(smin, sig[k]);
        if (smin <= eps) { failed = true; break; }   // "sigma-squared falls below threshold"
#pragma unroll
        for (int k = 0; k < G; k++) {
            nh[k] = -0.5 / sig[k];
            lc[k] = log(pro[k]) - 0.5 * (PI2LOG + log(sig[k]));
        }
        iter++;
        // ---- E-step fused with the sufficient statistics of the next M-step ----
#pragma unroll
        for (int k = 0; k < G; k++) { A[k] = 0.0; B[k] = 0.0; C[k] = 0.0; }
        double hl = 0.0, pl = 1.0;
        int since = 0;
#pragma unroll 2
        for (int i = lane; i < n; i += 32) {
            double x = xs[i];
            double e[G], d[G], q[G], S, tmax;
            estep_point<G>(x, mu, nh, lc, e, d, q, S, tmax);
            double rs = __drcp_rn(S);
#pragma unroll
            for (int k = 0; k < G; k++) {
                double z = __dmul_rn(e[k], rs);
                A[k] += z;
                B[k] = fma(z, d[k], B[k]);
                C[k] = fma(z, q[k], C[k]);
            }
            // sum_i log S_i as the log of a running product (S in [1, G]); flushed before it can overflow
            pl *= S;
            hl += tmax;
            if (++since == 64) { hl += log(pl); pl = 1.0; since = 0; }
        }
        hl += log(pl);
        {
            double red[3 * G + 1];
#pragma unroll
            for (int k = 0; k < G; k++) { red[k] = A[k]; red[G + k] = B[k]; red[2 * G + k] = C[k]; }
            red[3 * G] = hl;
            warp_sum_vec<3 * G + 1>(red);
#pragma unroll
            for (int k = 0; k < G; k++) { A[k] = red[k]; B[k] = red[G + k]; C[k] = red[2 * G + k]; }
            hood = red[3 * G];
        }
        st.triples += dn * G;
        st.point_iters += dn;
        double err = fabs(hold - hood) / (1.0 + fabs(hood));
        hold = hood;
        if (fabs(err - tol) <= 1e-7 * tol) st.tol_near += 1.0;
        if (!(err > tol)) break;
        if (iter >= EM_ITER_CAP) { st.cap_hits += 1.0; break; }
        // ---- M-step ----
        double amin = A[0];
#pragma unroll
        for (int k = 1; k < G; k++) amin = fmin(amin, A[k]);
        if (amin <= rteps) { failed = true; break; }
#pragma unroll
        for (int k = 0; k < G; k++) {
            double ra = 1.0 / A[k];
            double dm = B[k] * ra;
            mu[k] = mu[k] + dm;
            sig[k] = fma(-dm, dm, C[k] * ra);
            pro[k] = A[k] / dn;
        }
    }
    st.em_iters += (double)iter;
    if (lane == 0) row->iters = iter;
    if (failed) return;

    // ---- converged: BIC, parameters, labels ----
    row->loglik = hood;
    row->bic = 2.0 * hood - (double)(3 * G - 1) * log(dn);
    bool use_extra = false;
    double mu2[G], sig2[G];
    if (extra_mstep) {
        // summaryMclustBIC: if sum((pro - colMeans(z))^2) > sqrt(eps) replace $parameters by one more M-step
        double dev = 0.0;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < G; k++) {
            double cm = A[k] / dn;
            dev += (pro[k] - cm) * (pro[k] - cm);
            if (!(A[k] > 0.0)) ok = false;
        }
        if (dev > rteps && ok) {
            use_extra = true;
#pragma unroll
            for (int k = 0; k < G; k++) {
                double ra = 1.0 / A[k];
                double dm = B[k] * ra;
                mu2[k] = mu[k] + dm;
                sig2[k] = fma(-dm, dm, C[k] * ra);
            }
        }
    }
    // MAP labels from the final z (same parameters as the last E-step) and their statistics
    double lab[3 * G];
#pragma unroll
    for (int k = 0; k < 3 * G; k++) lab[k] = 0.0;
    for (int i = lane; i < n; i += 32) {
        double x = xs[i];
        int bk = map_label<G>(x, mu, nh, lc);
#pragma unroll
        for (int k = 0; k < G; k++) {
            if (k == bk) { lab[k] += 1.0; lab[G + k] += x; lab[2 * G + k] = fma(x, x, lab[2 * G + k]); }
        }
    }
    warp_sum_vec<3 * G>(lab);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < G; k++) {
            row->mean[k] = use_extra ? mu2[k] : mu[k];
            row->sigsq[k] = use_extra ? sig2[k] : sig[k];
            row->mu_em[k] = mu[k];
            row->nh_em[k] = nh[k];
            row->lc_em[k] = lc[k];
            row->cnt[k] = lab[k];
            row->sy[k] = lab[G + k];
            row->syy[k] = lab[2 * G + k];
        }
        row->valid = 1;
    }
}

// G = 1 row: mvn1d closed form
__device__ __forceinline__ void fit_row1(const double* xs, int n, int lane, Row* row) {
    double a[2] = {0.0, 0.0};
    for (int i = lane; i < n; i += 32) { double x = xs[i]; a[0] += x; a[1] = fma(x, x, a[1]); }
    warp_sum_vec<2>(a);
    double dn = (double)n;
    double mu = a[0] / dn;
    double ss = 0.0;
    for (int i = lane; i < n; i += 32) { double t = xs[i] - mu; ss = fma(t, t, ss); }
    ss = warp_sum(ss) / dn;
    if (lane == 0) {
        row->iters = 0;
        row->mean[0] = mu; row->sigsq[0] = ss;
        row->cnt[0] = dn; row->sy[0] = a[0]; row->syy[0] = a[1];
        row->mu_em[0] = mu; row->nh_em[0] = 0.0; row->lc_em[0] = 0.0;
        if (ss == 0.0) { row->valid = 0; }
        else {
            double ll = -dn * (PI2LOG + (1.0 + log(ss))) / 2.0;
            row->loglik = ll;
            row->bic = 2.0 * ll - 2.0 * log(dn);
            row->valid = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The reference's scalar decision procedure on the table.  All lanes execute it redundantly on
// identical shared-memory data.
// ---------------------------------------------------------------------------------------------
struct Tab {   // table(cl): counts of the PRESENT labels in increasing label order
    int nt;
    double cnt[GMAX];
    int lab[GMAX];
};

__device__ __forceinline__ int pick_upto(const Row* rows, int k) {   // Mclust(y, "V", G = 1:k): pickBIC, ties -> smaller G
    int best = 0;
    for (int g = 1; g <= k && g <= GMAX; g++) {
        if (!rows[g - 1].valid) continue;
        if (best == 0 || rows[g - 1].bic > rows[best - 1].bic) best = g;
    }
    return best;
}

__device__ __forceinline__ void table_of(const Row* rows, int comps, Tab& t) {
    t.nt = 0;
    for (int k = 0; k < comps; k++)
        if (rows[comps - 1].cnt[k] > 0.0) { t.cnt[t.nt] = rows[comps - 1].cnt[k]; t.lab[t.nt] = k; t.nt++; }
}

__device__ __forceinline__ double tab_min(const Tab& t) {
    double m = t.cnt[0];
    for (int i = 1; i < t.nt; i++) m = fmin(m, t.cnt[i]);
    return m;
}

__device__ __forceinline__ double r_min_nan(const double* v, int n) {   // R's min(): NaN propagates
    double m = v[0];
    for (int i = 0; i < n; i++) { if (isnan(v[i])) return nan(""); if (v[i] < m) m = v[i]; }
    return m;
}

// diagnostics of a >= 2 component model: variance ratio and the min/max-variance component indices
__device__ __forceinline__ void var_diag(const Row& mc, int comps, int& maxcl, int& mincl, double& vardiff, double& sdmax) {
    maxcl = 0; mincl = 0;
    for (int k = 1; k < comps; k++) {
        if (mc.sigsq[k] > mc.sigsq[maxcl]) maxcl = k;   // which.max / which.min: first extremum
        if (mc.sigsq[k] < mc.sigsq[mincl]) mincl = k;
    }
    vardiff = mc.sigsq[maxcl] / mc.sigsq[mincl];
    sdmax = 0.0;
    for (int k = 0; k < comps; k++) sdmax = fmax(sdmax, sqrt(mc.sigsq[k]));
}

// R/mclust.restricted.R:52-237.  Returns the number of components of the selected model (0: no valid model).
__device__ __noinline__ int restricted_select(const Row* rows, int n, const DevCfg& cfg) {
    const double NaN = nan("");
    int comps = pick_upto(rows, 5);                                   // :52-55
    if (comps == 0) return 0;
    Tab tb;
    table_of(rows, comps, tb);
    if (comps > tb.nt) {                                              // :57-63
        comps = pick_upto(rows, comps - 1);
        if (comps == 0) return 0;
        table_of(rows, comps, tb);
    }
    const double dn = (double)n;
    const double remThresh = 1.0 / (1.0 + exp(-dn / 60.0 + 0.4)) + 1.2;   // :67
    const double addThresh = 1.0 / (1.0 + exp(dn / 9.0 - 4.5)) + 3.4;     // :72

    if (cfg.restrict_) {
        if (comps > 1) {                                              // :75-185
            const Row* mc = &rows[comps - 1];
            double meandiff[GMAX];
            int nmd = comps - 1;
            int maxcl, mincl;
            double vardiff, sdmax;
            var_diag(*mc, comps, maxcl, mincl, vardiff, sdmax);        // :79-82
            for (int j = 0; j < nmd; j++) meandiff[j] = (mc->mean[j + 1] - mc->mean[j]) / sdmax;   // :77
            double nmin = (mincl < tb.nt) ? tb.cnt[mincl] : NaN;       // :83  table(cl)[min.cl] is positional
            double mincat = tab_min(tb);                               // :84
            int tries = 0;
            double nmax = (maxcl < tb.nt) ? tb.cnt[maxcl] : NaN;       // :86
            double balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));   // :87
            for (int j = 0; j < nmd; j++) meandiff[j] *= balance;      // :88
            if (isnan(vardiff)) vardiff = 1.0;                         // :91-110
            if (isnan(nmin)) nmin = INFINITY;
            {
                bool anyna = false;
                for (int j = 0; j < nmd; j++) anyna |= isnan(meandiff[j]);
                if (nmd == 0 || anyna) { meandiff[0] = INFINITY; nmd = 1; }
            }
            int err = 0;
            while ((r_min_nan(meandiff, nmd) < remThresh || vardiff > 20.0) && mincat > 2.0 && tries <= 4 && err == 0) {   // :115
                int comps_old = comps;
                comps = pick_upto(rows, comps - 1);                    // :119-124
                if (comps == 0) return 0;
                mc = &rows[comps - 1];
                table_of(rows, comps, tb);
                mincat = tab_min(tb);                                  // :125
                tries++;
                if (comps > 1 && comps_old - comps == 1) {             // :128
                    // :129 by(y, cl, mean): empirical means of the present labels
                    var_diag(*mc, comps, maxcl, mincl, vardiff, sdmax);
                    nmd = tb.nt - 1;
                    for (int j = 0; j < nmd; j++) {
                        double m1 = mc->sy[tb.lab[j + 1]] / tb.cnt[j + 1];
                        double m0 = mc->sy[tb.lab[j]] / tb.cnt[j];
                        meandiff[j] = (m1 - m0) / sdmax;               // :130
                    }
                    balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));   // :131 (stale nmin / nmax)
                    for (int j = 0; j < nmd; j++) meandiff[j] *= balance;
                    nmin = (mincl < tb.nt) ? tb.cnt[mincl] : NaN;      // :138
                } else {
                    meandiff[0] = 10.0; nmd = 1; vardiff = -INFINITY; nmin = 100.0;   // :140-142
                }
                if (isnan(vardiff)) { vardiff = 1.0; err = 1; }        // :146-173
                if (isnan(nmin)) { nmin = INFINITY; err = 1; }
                {
                    bool anyna = false;
                    for (int j = 0; j < nmd; j++) anyna |= isnan(meandiff[j]);
                    if (nmd == 0 || anyna) { meandiff[0] = INFINITY; nmd = 1; err = 1; }
                }
                if (err == 1) {                                        // :176-183 revert
                    comps = pick_upto(rows, comps + 1);
                    if (comps == 0) return 0;
                    table_of(rows, comps, tb);
                }
            }
        } else {                                                       // :186-225 try to add a component
            bool keep2 = false;
            if (n >= 2 && rows[1].valid) {
                Tab t2;
                table_of(rows, 2, t2);
                if (t2.nt == 2) {                                      // :195
                    const Row* m2 = &rows[1];
                    double mincat = fmin(t2.cnt[0], t2.cnt[1]);
                    int maxcl, mincl;
                    double vardiff, sdmax;
                    var_diag(*m2, 2, maxcl, mincl, vardiff, sdmax);
                    double md = (m2->mean[1] - m2->mean[0]) / sdmax;
                    double nmin = t2.cnt[mincl], nmax = t2.cnt[maxcl];
                    double balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));
                    md *= balance;
                    if ((md > addThresh && vardiff < 10.0) && mincat > 2.0) keep2 = true;   // :211
                }
            }
            comps = keep2 ? 2 : 1;
            table_of(rows, comps, tb);
        }
    }
    // :230-237 every cluster smaller than min.size -> drop one component
    {
        double mx = 0.0;
        for (int i = 0; i < tb.nt; i++) mx = fmax(mx, tb.cnt[i]);
        if (comps > 1 && mx < (double)cfg.min_size) {
            comps = pick_upto(rows, comps - 1);
        }
    }
    return comps;
}

// order(mc$parameters$mean), stable (R/mclust.restricted.R:243)
__device__ __forceinline__ void mean_order(const Row& mc, int comps, int* ord) {
    for (int k = 0; k < comps; k++) ord[k] = k;
    for (int i = 1; i < comps; i++) {
        int o = ord[i];
        int j = i - 1;
        while (j >= 0 && mc.mean[ord[j]] > mc.mean[o]) { ord[j + 1] = ord[j]; j--; }
        ord[j + 1] = o;
    }
}

// R/joint.posterior.R:35-64 on the relabelled classes of the selected model.
// Quirk kept: the loop runs over labels 1..r with r = number of DISTINCT labels, so a gap in the label
// set gives nk = 0 -> lgamma(0) = Inf (and cells of labels > r are not counted).
__device__ __noinline__ double joint_posterior(const Row& mc, int comps, const DevCfg& cfg) {
    int ord[GMAX];
    mean_order(mc, comps, ord);
    double nk[GMAX], sy[GMAX], syy[GMAX];
    for (int k = 0; k < GMAX; k++) { nk[k] = 0.0; sy[k] = 0.0; syy[k] = 0.0; }
    int r = 0;
    for (int l = 0; l < comps; l++) {
        // cl2[cl == l] <- as.numeric(names(mean[order(mean)])[l])  (order(), not rank(): quirk viii)
        int nl = (comps > 1) ? ord[l] : l;
        nk[nl] = mc.cnt[l]; sy[nl] = mc.sy[l]; syy[nl] = mc.syy[l];
        if (mc.cnt[l] > 0.0) r++;
    }
    const double alpha = cfg.alpha, m0 = cfg.m0, s0 = cfg.s0, a0 = cfg.a0, b0 = cfg.b0;
    double logpost = 0.0, J = 0.0;
    for (int k = 0; k < r; k++) J += nk[k];
    for (int k = 0; k < r; k++) {
        double sk = s0 + nk[k];
        double mk = (s0 * m0 + sy[k]) / sk;
        double ak = a0 + nk[k];
        double bk = __dsub_rn(__dadd_rn(__dadd_rn(b0, syy[k]), __dmul_rn(s0, __dmul_rn(m0, m0))), __dmul_rn(sk, __dmul_rn(mk, mk)));
        double lgn = (nk[k] == 0.0) ? INFINITY : lgamma(nk[k]);
        logpost = logpost + lgn + lgamma(ak / 2.0) - (ak / 2.0) * log(bk / 2.0) - 0.5 * log(sk);
    }
    logpost = logpost + (double)r * log(alpha) + lgamma(alpha) - J * lgamma(a0 / 2.0) +
              (J * a0 / 2.0) * log(b0 / 2.0) + J * 0.5 * log(s0) - (J / 2.0) * log(2.0 * 3.14159265358979323846) -
              lgamma(alpha + J);
    return logpost;
}

// ---------------------------------------------------------------------------------------------
// The whole fit-set on a sorted slab.  Returns the selected number of components (0 = failed).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fit_set(const double* xs, int n, int lane, Row* rows, const DevCfg& cfg, WarpStats& st) {
    fit_row1(xs, n, lane, &rows[0]);
    fit_row<2>(xs, n, lane, &rows[1], st, cfg.extra_mstep);
    fit_row<3>(xs, n, lane, &rows[2], st, cfg.extra_mstep);
    fit_row<4>(xs, n, lane, &rows[3], st, cfg.extra_mstep);
    fit_row<5>(xs, n, lane, &rows[4], st, cfg.extra_mstep);
    __syncwarp();
    st.fit_sets += 1.0;
    // near-tie diagnostic: two best BICs within 1e-9 relative
    {
        double b1 = -INFINITY, b2 = -INFINITY;
        for (int g = 0; g < GMAX; g++) if (rows[g].valid) {
            double b = rows[g].bic;
            if (b > b1) { b2 = b1; b1 = b; } else if (b > b2) b2 = b;
        }
        if (b2 > -INFINITY && fabs(b1 - b2) <= 1e-9 * fabs(b1)) st.bic_near += 1.0;
    }
    int comps = restricted_select(rows, n, cfg);
    if (comps == 0) st.failed_sets += 1.0;
    return comps;
}

}  // namespace scdd
