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
// Data placement: the vector lives in the team's slab (a warp's shared-memory slab, or for large vectors a
// CTA's shared / L2-resident slab), SORTED (the EM sums and the selection are order independent; sorting makes
// mclust's quantile initialisation a lookup).  Threads own points i = tid, tid + team size, ...; sufficient
// statistics are reduced with warp shuffles (plus one shared-memory hop for CTA teams).  All arithmetic is fp64.
#pragma once
#include <utility>
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
constexpr int FIT_DEFERRED = -2;               // fit_set: stopped by the caller's iteration budget

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

__device__ double EXP_TABLE_GLOBAL[SCDD_EXP_TABLE_SIZE];   // source of the per-CTA shared copy, filled by init_exp_table()
// The CTA's copy: a statically allocated shared array, so that every function (including the non-inlined E-step
// dispatchers) addresses it as [index * 8 + link-time constant] with no base register and no address arithmetic.
__shared__ double EXP_TABLE_SH[SCDD_EXP_TABLE_SIZE];
// fast_log's table (device_math.cuh): 4 KB of static shared memory per CTA next to the exp table
__device__ LogTab LOG_TABLE_GLOBAL[SCDD_LOG_TABLE_SIZE];
__shared__ __align__(16) LogTab LOG_TABLE_SH[SCDD_LOG_TABLE_SIZE];
__device__ __forceinline__ double fast_log(double x) {
    const unsigned base = (unsigned)__cvta_generic_to_shared(LOG_TABLE_SH);
    unsigned addr;
    double rc, l;
    asm("{ .reg .b32 j; bfe.u32 j, %1, 12, 8; mad.lo.u32 %0, j, 16, %2; }" : "=r"(addr) : "r"(__double2hiint(x)), "r"(base));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rc), "=d"(l) : "r"(addr));
    return fast_log_core(x, rc, l);
}
// constant-bank copies of device_math.cuh's expc:: constants (fetched with LDCU, no per-iteration immediates)
__constant__ double EXP_TAB_CONST[3] = {2954.639443740597, -0.00033845077175778578, 1.0 / 6.0};

// 1/S for S in (0.97, 1.03 GMAX]: MUFU.RCP64H seed + one cubic Newton step (3 DFMA, ~1 ulp).  No special-case
// branch: the argument range is known, unlike for the general-purpose __drcp_rn.
__device__ __forceinline__ double rcp_sum(double x) {
#ifndef SCDD_FAST_RCP
#define SCDD_FAST_RCP 1
#endif
#if SCDD_FAST_RCP
    // the seed is good to ~2^-23, e + e^2 makes the step cubic (2^-69): one step reaches rounding level
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
#else
    return __drcp_rn(x);
#endif
}

// 1/x for positive normal x (M-step scalars): same seed + two Newton steps; within 1 ulp.  Used instead of the
// IEEE division of the reference's Fortran (a rounding-level difference) because nvcc's correctly rounded
// fp64 division is a ~60-instruction subroutine call on the per-iteration critical path of every warp.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// Component parameters in E-step form: mu_k, nh_k = -1/(2 sigsq_k), lc_k = log pro_k - (log 2 pi + log sigsq_k)/2.
// ParamsReg keeps them in registers (3G doubles per lane); ParamsSmem re-reads them from the warp's
// shared-memory block for every point (broadcast loads on the LSU pipe), which frees 6G registers; used for
// G = 5, where the register file is the scarcer resource (load_params).
// 2^((D mod 2048)/2048) from the CTA's table: mask, then one multiply-add forms the shared address
// (the compiler's own sequence for tab[D & 2047] is shift, and, add).
__device__ __forceinline__ double exp_table_sh(int N) {
    const unsigned base = (unsigned)__cvta_generic_to_shared(EXP_TABLE_SH);
    unsigned addr;
    double T;
    asm("{ .reg .b32 j; and.b32 j, %1, %2; mad.lo.u32 %0, j, 8, %3; }" : "=r"(addr) : "r"(N), "n"(expc::TMASK), "r"(base));
    asm("ld.shared.f64 %0, [%1];" : "=d"(T) : "r"(addr));
    return T;
}

template <int G, bool GTAB = false>
struct ParamsReg {
    double mu[G], nh[G], lc[G];
    __device__ __forceinline__ static double c(int i) { return EXP_TAB_CONST[i]; }
    // 2^(j/2048): the CTA's shared copy, or (GTAB, label_of outside the fit loop) the global source
    __device__ __forceinline__ static double lookup(int D) { return GTAB ? EXP_TABLE_GLOBAL[D & expc::TMASK] : exp_table_sh(D); }
    __device__ __forceinline__ void get(int k, double& m, double& h, double& l) const { m = mu[k]; h = nh[k]; l = lc[k]; }
};
struct ParamsSmem {
    __device__ __forceinline__ static double lookup(int D) { return exp_table_sh(D); }
    __device__ __forceinline__ static double c(int i) { return EXP_TAB_CONST[i]; }
    unsigned addr;   // shared-window address of the warp's {mu, nh, lc, pad}[GMAX] block (32 B per component)
    __device__ __forceinline__ void get(int k, double& m, double& h, double& l) const {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m), "=d"(h) : "r"(addr + 32u * k));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(l) : "r"(addr + 32u * k + 16u));
    }
};

// sum of G values as a balanced tree (dependency depth ceil(log2 G) instead of G - 1)
template <int G>
__device__ __forceinline__ double tree_sum(const double* e) {
    if (G == 1) return e[0];
    if (G == 2) return __dadd_rn(e[0], e[1]);
    if (G == 3) return __dadd_rn(__dadd_rn(e[0], e[1]), e[2]);
    double s = __dadd_rn(__dadd_rn(e[0], e[1]), __dadd_rn(e[2], e[3]));
    if (G == 5) s = __dadd_rn(s, e[4]);
    return s;
}
template <int G>
__device__ __forceinline__ int tree_imax(const int* v) {
    if (G == 1) return v[0];
    if (G == 2) return max(v[0], v[1]);
    if (G == 3) return max(max(v[0], v[1]), v[2]);
    int m = max(max(v[0], v[1]), max(v[2], v[3]));
    if (G == 5) m = max(m, v[4]);
    return m;
}

// U points x G components: softmax terms (device_math.cuh, "The E-step's softmax terms"), written column-wise:
// every step is emitted for all U*G terms before the next one (U*G independent fp64 instructions between an
// operation and its consumer; ptxas is free to re-order, see DESIGN.md "Issue model").
//   Outputs: e (U*G), d = x - mu, q = d^2, S = sum_k e_k and the integer shift M per point.
// SAT = false drops the saturation test of term_n: legal when the caller has proved |t_k| < 2^19 for every point
// (all variances above (data range)^2 / 2^20 -- see fit_row), which is the case for all but degenerate fits.
// SAT: additionally, a point whose EVERY term is saturated (more than 2^18 log-units from all components) is redone
// on a_k - max_k a_k, and that fp64 shift is returned in `far` for the log-likelihood (0 otherwise): the softmax is
// exact in every regime, as mclust's own exp(t_k - max t) is.
template <int G, int U, bool SAT, class P>
__device__ __forceinline__ void estep_points(const double (&x)[U], const P& prm, double (&e)[U * G], double (&d)[U * G],
                                             double (&q)[U * G], double (&S)[U], int (&M)[U], double* far = nullptr) {
    using namespace expc;
    constexpr int N = U * G;
    double a[N], t[N], r[N], p[N], h[G], l[G], T[N];
    int n[N], nh[N];
#pragma unroll
    for (int k = 0; k < G; k++) {
        double m;
        prm.get(k, m, h[k], l[k]);
#pragma unroll
        for (int u = 0; u < U; u++) d[u * G + k] = __dsub_rn(x[u], m);
    }
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = __dmul_rn(d[i], d[i]);
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
        for (int k = 0; k < G; k++) a[u * G + k] = fma(q[u * G + k], h[k], l[k]);
    }
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = fma(a[i], prm.c(0), MAGIC);                 // term_t
#pragma unroll
    for (int i = 0; i < N; i++) n[i] = SAT ? term_n(a[i], t[i]) : __double2loint(t[i]);
#pragma unroll
    for (int i = 0; i < N; i++) T[i] = P::lookup(n[i]);          // 2^(j/2048), j = N & 2047: independent of the shift
#pragma unroll
    for (int i = 0; i < N; i++) nh[i] = n[i] >> TB;
#pragma unroll
    for (int u = 0; u < U; u++) M[u] = tree_imax<G>(&nh[u * G]);
    if (SAT) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            double amax = 0.0;
            if (M[u] == HSAT) {            // cold: every term saturated
                amax = a[u * G];
#pragma unroll
                for (int k = 1; k < G; k++) amax = fmax(amax, a[u * G + k]);
#pragma unroll
                for (int k = 0; k < G; k++) {
                    const int i = u * G + k;
                    a[i] = __dsub_rn(a[i], amax);
                    t[i] = fma(a[i], prm.c(0), MAGIC);
                    n[i] = term_n(a[i], t[i]);
                    T[i] = P::lookup(n[i]);
                    nh[i] = n[i] >> TB;
                }
                M[u] = tree_imax<G>(&nh[u * G]);
            }
            if (far) far[u] = amax;
        }
    }
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = __dsub_rn(t[i], MAGIC);                             // term_r
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = fma(p[i], prm.c(1), a[i]);
    if (SAT) {
        // a saturated term is below every finite one by construction (k = KMIN), but beyond |a| ~ 2^52 ln2 / 2^11 =
        // 1.5e12 the sum a c + MAGIC is no longer an integer and the "reduced argument" is garbage of any size: with
        // 1 + s <= 1/4 the exponent insert below would borrow from the sign bit.  Reachable when a variance sits
        // between eps and ~1e-11 (range^2 / 2 sigsq > 1.5e12): values that almost, but not exactly, coincide.
#pragma unroll
        for (int i = 0; i < N; i++) r[i] = (n[i] == NSAT) ? 0.0 : r[i];
    }
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fma(prm.c(2), r[i], 0.5);                   // term_s
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = fma(p[i], r[i], 1.0);
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = p[i] * r[i];
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
        for (int k = 0; k < G; k++) {                                                      // term_e
            const int i = u * G + k;
            const double v = fma(T[i], p[i], T[i]);
            int hi;   // hi(v) + (max(H - Mh, KMIN) << 20): add-max, then multiply-add (two integer instructions)
            asm("{ .reg .s32 k; sub.s32 k, %1, %2; max.s32 k, k, %3; mad.lo.s32 %0, k, 1048576, %4; }"
                : "=r"(hi) : "r"(nh[i]), "r"(M[u]), "n"(KMIN), "r"(__double2hiint(v)));
            e[i] = __hiloint2double(hi, __double2loint(v));
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) S[u] = tree_sum<G>(&e[u * G]);
}

// map(z): first index of the row maximum of z = e / S
template <int G, class P>
__device__ __forceinline__ int map_label(double x, const P& prm) {
    double xx[1] = {x};
    double e[G], d[G], q[G], S[1];
    int M[1];
    estep_points<G, 1, true, P>(xx, prm, e, d, q, S, M);
    double rs = rcp_sum(S[0]);
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
// Transposing warp reduction: every lane contributes 16 partial sums; afterwards lane L holds the
// warp total of slot (L >> 1).  16 + 8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5 butterflies.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double reduce16(double (&v)[16], int lane) {
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int h = 8 >> s, off = 16 >> s;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < h; j++) {
            double send = up ? v[j] : v[j + h];
            double keep = up ? v[j + h] : v[j];
            v[j] = keep + __shfl_xor_sync(FULLMASK, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(FULLMASK, v[0], 1);
}
__device__ __forceinline__ double slot(double t, int e) { return __shfl_sync(FULLMASK, t, 2 * e); }

// A fit-set is executed by a "team": one warp (TW = 1; the vector lives in the warp's shared-memory slab) or TW warps
// of a CTA (large vectors; shared or global memory).  A CTA of 16 warps holds 16 / TW teams; teams of several warps
// synchronise on their own named barrier (id 1 + team index), so the teams of a CTA run independent fit-sets.
// Every thread of a team follows the same control flow on identical reduced values, so team-wide barriers inside the
// fit are legal.
template <int TW>
struct Team {
    static constexpr int NT = 32 * TW;
    __device__ __forceinline__ static int index() { return TW == 1 ? (int)(threadIdx.x >> 5) : (int)threadIdx.x / NT; }
    __device__ __forceinline__ static int tid(int lane) { return TW == 1 ? lane : (int)threadIdx.x % NT; }
    __device__ __forceinline__ static int warp() { return TW == 1 ? 0 : ((int)(threadIdx.x >> 5)) % TW; }   // warp within the team
    __device__ __forceinline__ static void sync() {
        if (TW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + index()), "n"(NT) : "memory");
    }
    __device__ __forceinline__ static bool lead() { return TW == 1 || warp() == 0; }
    // all-reduce of 16 slots: afterwards lane L of EVERY warp holds the team total of slot (L >> 1)
    __device__ __forceinline__ static double reduce(double (&v)[16], int lane, double* red) {
        double t = reduce16(v, lane);
        if (TW > 1) {
            const int w = warp();
            sync();
            if ((lane & 1) == 0) red[w * 16 + (lane >> 1)] = t;
            sync();
            double s = 0.0;
#pragma unroll 1
            for (int k = 0; k < TW; k++) s += red[k * 16 + (lane >> 1)];
            t = s;
        }
        return t;
    }
};

// ---------------------------------------------------------------------------------------------
// E-step over the warp's points fused with the sufficient statistics of the next M-step.
// Component k's parameters live in lane k ("owner lane"); they are broadcast into registers here.
// Per-lane partial sums go to v: [0,G) sum z, [G,2G) sum z d, [2G,3G) sum z d^2, [3G] log-likelihood.
//
// M-step form: with d = x - mu_old and q = d^2 already needed by the E-step, the new parameters are
//   mu = mu_old + S(z d)/S(z),  sigsq = S(z q)/S(z) - (S(z d)/S(z))^2
// which equals mclust's two-pass update in exact arithmetic and needs no stored z matrix.
// ---------------------------------------------------------------------------------------------
// back half of a point: normalise the softmax terms and add the point to the M-step sums.
// sum_k z_k = 1 for every point (to rounding), so the weight of the LAST component is not accumulated: estep_pass
// derives it as (points of this thread) - sum_{k < G-1} A_k, one DADD per point less.
template <int G, int U>
__device__ __forceinline__ void accumulate_back(const double (&e)[U * G], const double (&d)[U * G], const double (&q)[U * G],
                                                const double (&S)[U], const int (&M)[U], double (&A)[G], double (&B)[G],
                                                double (&C)[G], double& pl, int& hs) {
#pragma unroll
    for (int u = 0; u < U; u++) {
        double rs = rcp_sum(S[u]);
#pragma unroll
        for (int k = 0; k < G; k++) {
            double z = __dmul_rn(e[u * G + k], rs);
            if (k < G - 1) A[k] += z;
            B[k] = fma(z, d[u * G + k], B[k]);
            C[k] = fma(z, q[u * G + k], C[k]);
        }
        pl *= S[u];      // sum_i log S_i is taken as the log of a running product (S in (0.97, 1.03 G])
        hs += M[u];      // the shifts are whole powers of two, |Mh| <= 2^19: summed exactly in 32 bits over a 256-point chunk
    }
}

// G = 2 on the fast path: the softmax of two terms is a logistic function of their difference, so a point needs ONE
// exponential: with dl = a0 - a1 and e = exp(-|dl|) in (0, 1], the larger term has weight 1 / (1 + e), the smaller
// e / (1 + e), and log(e^a0 + e^a1) = max(a0, a1) + log(1 + e).  The maxima are summed in fp64 (hmax), log(1 + e)
// through the running product like the other loops.  27 fp64 instructions per point instead of 32, one table lookup
// instead of two, no integer max / shift bookkeeping.  Same reduction, table and polynomial as term_e; differences
// |dl| >= 2^18 saturate to the smallest representable weight (term_n).
template <class P>
__device__ __forceinline__ void accumulate_point2(double x, const P& prm, double (&A)[2], double (&B)[2], double (&C)[2],
                                                  double& pl, double& hmax) {
    using namespace expc;
    double m0, h0, l0, m1, h1, l1;
    prm.get(0, m0, h0, l0);
    prm.get(1, m1, h1, l1);
    const double d0 = __dsub_rn(x, m0), d1 = __dsub_rn(x, m1);
    const double q0 = __dmul_rn(d0, d0), q1 = __dmul_rn(d1, d1);
    const double a0 = fma(q0, h0, l0), a1 = fma(q1, h1, l1);
    const double dl = __dsub_rn(a0, a1);
    const bool first = dl >= 0.0;                        // component 0 carries the larger term (ties: component 0)
    const double u = -fabs(dl);
    const double t = fma(u, prm.c(0), MAGIC);
    const int n = term_n(u, t);
    const double T = P::lookup(n);
    int k = n >> TB;
    k = k < KMIN ? KMIN : k;
    const double r = fma(__dsub_rn(t, MAGIC), prm.c(1), u);
    double p = fma(prm.c(2), r, 0.5);
    p = fma(p, r, 1.0);
    p = p * r;
    const double v = fma(T, p, T);
    const double e = __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
    const double S = __dadd_rn(1.0, e);
    const double rs = rcp_sum(S);
    const double zs = __dmul_rn(e, rs);
    const double z0 = first ? rs : zs, z1 = first ? zs : rs;
    A[0] += z0;
    B[0] = fma(z0, d0, B[0]);
    C[0] = fma(z0, q0, C[0]);
    B[1] = fma(z1, d1, B[1]);
    C[1] = fma(z1, q1, C[1]);
    pl *= S;
    hmax += first ? a0 : a1;
}

template <int G, int U, bool SAT, class P>
__device__ __forceinline__ void accumulate_points(const double (&x)[U], const P& prm, double (&A)[G], double (&B)[G],
                                                  double (&C)[G], double& pl, int& hs, double& hfar) {
    if constexpr (G == 2 && U == 1 && !SAT) {
        accumulate_point2(x[0], prm, A, B, C, pl, hfar);   // hfar doubles as the sum of the maxima (unused when !SAT)
        return;
    }
    double e[U * G], d[U * G], q[U * G], S[U];
    int M[U];
    if (SAT) {
        double far[U];
        estep_points<G, U, SAT, P>(x, prm, e, d, q, S, M, far);
#pragma unroll
        for (int u = 0; u < U; u++) hfar += far[u];
    } else {
        estep_points<G, U, SAT, P>(x, prm, e, d, q, S, M);
    }
    accumulate_back<G, U>(e, d, q, S, M, A, B, C, pl, hs);
}

#ifndef SCDD_PARAMS_SMEM
#define SCDD_PARAMS_SMEM 1
#endif

#ifndef SCDD_PARAMS_REG_MAXG
#define SCDD_PARAMS_REG_MAXG 5      // all G: 6G registers save 2 shared loads per term (A/B at 8192 genes: G <= 4 only is 2.3 % slower)
#endif
// Loads the three exp constants once and hides their origin from the compiler, which otherwise re-loads them from
// the constant bank inside every loop iteration (three extra issue slots per point).
template <int G>
__device__ __forceinline__ auto load_params(const double* pm) {
#if SCDD_PARAMS_SMEM
    if constexpr (G > SCDD_PARAMS_REG_MAXG) {
        ParamsSmem p;
        p.addr = (unsigned)__cvta_generic_to_shared(pm);
        return p;
    } else {
        ParamsReg<G> p;
#pragma unroll
        for (int k = 0; k < G; k++) { p.mu[k] = pm[4 * k]; p.nh[k] = pm[4 * k + 1]; p.lc[k] = pm[4 * k + 2]; }
        return p;
    }
#else
    ParamsReg<G> p;
#pragma unroll
    for (int k = 0; k < G; k++) { p.mu[k] = pm[4 * k]; p.nh[k] = pm[4 * k + 1]; p.lc[k] = pm[4 * k + 2]; }
    return p;
#endif
}

// one full staging batch: STAGE_K 8-byte cp.async with immediate offsets off one shared and one global base address
constexpr int STAGE_K = 16;
template <int NT, int J>
__device__ __forceinline__ void stage_copy_one(unsigned dst, size_t src) {
    asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%2], 8;" ::"r"(dst), "l"(src), "n"(J * NT * 8) : "memory");
}
template <int NT, int... J>
__device__ __forceinline__ void stage_copy_full(unsigned dst, size_t src, std::integer_sequence<int, J...>) {
    (stage_copy_one<NT, J>(dst, src), ...);
}

template <int G, int U, int TW, bool SAT>
__device__ __forceinline__ void estep_pass(const double* __restrict__ xs, int n, int lane, const double* pm,
                                           double (&v)[16], double* stage = nullptr, int stage_k = 0) {
    constexpr int NT = Team<TW>::NT;
    const auto prm = load_params<G>(pm);
    double A[G], B[G], C[G];
#pragma unroll
    for (int k = 0; k < G; k++) { A[k] = 0.0; B[k] = 0.0; C[k] = 0.0; }
    double hl = 0.0, hfar = 0.0;
    long long hs = 0;
    constexpr int CHUNK = 256 * NT;   // the running product is flushed every 256 points per thread (5^256 < DBL_MAX)
#pragma unroll 1
    for (int base = 0; base < n; base += CHUNK) {
        const int stop = (n - base < CHUNK) ? n : base + CHUNK;
        double pl = 1.0;
        int hc = 0;
        // pointer-bumped loops: the load is [register + immediate] and the bound test compares addresses
        const double* px = xs + base + Team<TW>::tid(lane);
        const double* const pend = xs + stop;
        if (U == 1) {
            if (TW == 1 || stage == nullptr) {
                // the slab is shared memory (always for a warp; for a multi-warp team unless it did not fit): 32-bit
                // shared-window addresses for the load and the bound test
                unsigned sx = (unsigned)__cvta_generic_to_shared(px);
                const unsigned send = (unsigned)__cvta_generic_to_shared(pend);
#pragma unroll 1
                for (; sx < send; sx += NT * (unsigned)sizeof(double)) {
                    double x[1];
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[0]) : "r"(sx));
                    accumulate_points<G, 1, SAT>(x, prm, A, B, C, pl, hc, hfar);
                }
            } else {
                // CTA teams whose slab lives in global memory (L2) -- fit-sets beyond 16k points, all of config 5.
                // Every thread stages ITS OWN points through shared memory in batches of stage_k with cp.async, two
                // batches in flight (the tile the sort used is free by now): no registers hold loads in flight, no
                // barrier is needed (a thread only reads what its own copies wrote), and the arithmetic runs the
                // same shared-memory loop as above instead of a loop that opens every iteration with an L2 round trip.
                const int mine = (px < pend) ? (int)((pend - px + NT - 1) / NT) : 0;      // this thread's points in the chunk
                const int K = stage_k;
                const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)Team<TW>::tid(lane) * 8u;
                const unsigned bstride = (unsigned)K * NT * 8u;                           // one buffer
                const size_t gpx = __cvta_generic_to_global(px);
                auto issue = [&](int b) {
                    size_t src = gpx + (size_t)b * K * NT * 8u;
                    unsigned dst = sbase + (unsigned)(b & 1) * bstride;
                    const int c = min(K, mine - b * K);
                    if (K == STAGE_K && c == STAGE_K) {
                        stage_copy_full<NT>(dst, src, std::make_integer_sequence<int, STAGE_K>());   // immediate offsets
                    } else {
#pragma unroll 1
                        for (int j = 0; j < c; j++) {
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
                            dst += NT * 8u;
                            src += NT * 8u;
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                const int nbatch = (mine + K - 1) / K;
                if (nbatch > 0) issue(0);
#pragma unroll 1
                for (int b = 0; b < nbatch; b++) {
                    if (b + 1 < nbatch) {
                        issue(b + 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    unsigned sx = sbase + (unsigned)(b & 1) * bstride;
                    const unsigned send = sx + (unsigned)min(K, mine - b * K) * NT * 8u;
#pragma unroll 1
                    for (; sx < send; sx += NT * (unsigned)sizeof(double)) {
                        double x[1];
                        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x[0]) : "r"(sx));
                        accumulate_points<G, 1, SAT>(x, prm, A, B, C, pl, hc, hfar);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (; px + NT * (U - 1) < pend; px += NT * U) {
                double x[U];
#pragma unroll
                for (int u = 0; u < U; u++) x[u] = px[NT * u];
                accumulate_points<G, U, SAT>(x, prm, A, B, C, pl, hc, hfar);
            }
#pragma unroll 1
            for (; px < pend; px += NT) {
                double x[1] = {*px};
                accumulate_points<G, 1, SAT>(x, prm, A, B, C, pl, hc, hfar);
            }
        }
        hs += hc;
        hl += fast_log(pl);
    }
    hl = fma((double)hs, expc::LN2_STEP, hl);   // log-likelihood: sum_i log S_i + 2^TB c sum_i Mh_i, the SAME c as in term_r
    if (SAT || G == 2) hl += hfar;          // G = 2 fast path: hfar = sum of the larger log-terms (accumulate_point2)
    // weight of the last component from sum_k z_k = 1: this thread's point count minus the other weights
    const int tid = Team<TW>::tid(lane);
    double rest = (double)(tid < n ? (n - tid + NT - 1) / NT : 0);
#pragma unroll
    for (int k = 0; k < G - 1; k++) rest -= A[k];
    A[G - 1] = rest;
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
#pragma unroll
    for (int k = 0; k < G; k++) { v[k] = A[k]; v[G + k] = B[k]; v[2 * G + k] = C[k]; }
    v[3 * G] = hl;
}

// MAP labels of the final z and their statistics: v[0,G) count, [G,2G) sum y, [2G,3G) sum y^2
template <int G, int TW>
__device__ __forceinline__ void label_pass(const double* __restrict__ xs, int n, int lane, const double* pm,
                                           double (&v)[16]) {
    const auto prm = load_params<G>(pm);
    double cnt[G], sy[G], syy[G];
#pragma unroll
    for (int k = 0; k < G; k++) { cnt[k] = 0.0; sy[k] = 0.0; syy[k] = 0.0; }
#pragma unroll 1
    for (int i = Team<TW>::tid(lane); i < n; i += Team<TW>::NT) {
        double x = xs[i];
        int bk = map_label<G>(x, prm);
#pragma unroll
        for (int k = 0; k < G; k++) {
            if (k == bk) { cnt[k] += 1.0; sy[k] += x; syy[k] = fma(x, x, syy[k]); }
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
#pragma unroll
    for (int k = 0; k < G; k++) { v[k] = cnt[k]; v[G + k] = sy[k]; v[2 * G + k] = syy[k]; }
}

// unroll factor of the point loop: small G has little instruction-level parallelism per point
#ifndef SCDD_UNROLL_MAXG
#define SCDD_UNROLL_MAXG 1          // two points per thread per step pay for their registers at no G (A/B: G <= 2 is 0.2 % slower)
#endif
template <int G> struct Unroll { static constexpr int value = (G <= SCDD_UNROLL_MAXG) ? 2 : 1; };

// same pass with the saturation test (degenerate fits: some variance below range^2 / 2^20); cold code, kept out of line
// so that it does not weigh on the register allocation of the EM loop
// The EM loop's state, resident in the team's shared memory across the E-step call.  ptxas keeps
// nothing in callee-saved registers around that call: in the round-1 shape every iteration reloaded ~45 spilled
// values from local memory in ~9 dependent clusters (ncu: the caller's loop was 4.8 % of the executed instructions
// but 10.6 % of the warp time, 60 % of it long-scoreboard stalls).  With the state in shared memory the loop holds
// two registers across the call (context pointer, lane) and reads what it needs in one batch.
struct EmCtx {
    const double* xs;
    double* pm;
    double* red;
    int n, G, iter, flags;          // flags bit 0: a stop test came within 1e-7 of the tolerance
    double* stage;                  // staging tile for a slab that lives in global memory (set once by fit_kernel), else null
    int stage_k, pad;               // points per thread and batch
    double hold, range2, inv_n;
    double sig[GMAX], pro[GMAX];    // mu lives in pm[4 k]
};
static_assert(sizeof(EmCtx) % 16 == 0, "keeps the slab 16-byte aligned");

// same pass with the saturation test (degenerate fits: some variance below range^2 / 2^20); cold code
template <int TW>
__device__ __noinline__ double estep_dispatch_sat(int G, const double* xs, int n, int lane, const double* pm,
                                                  double* red, double* stage, int stage_k) {
    double v[16];
    switch (G) {
        case 2: estep_pass<2, 1, TW, true>(xs, n, lane, pm, v, stage, stage_k); break;
        case 3: estep_pass<3, 1, TW, true>(xs, n, lane, pm, v, stage, stage_k); break;
        case 4: estep_pass<4, 1, TW, true>(xs, n, lane, pm, v, stage, stage_k); break;
        default: estep_pass<5, 1, TW, true>(xs, n, lane, pm, v, stage, stage_k); break;
    }
    return Team<TW>::reduce(v, lane, red);
}

// returns the team total of slot (lane >> 1): the 16 partial sums never leave registers.
// ONE non-inlined function holds the point loops of all four G: ptxas then allocates for the widest (G = 5, 128
// registers, callee-saved registers spilled once in the prologue) and every loop keeps its 64-bit accumulators in
// aligned pairs.  One function per G (tried in round 2) is allocated tightly enough to avoid the prologue spill, and
// pays for it with ~20 register moves per point inside the loop.
// Arguments come from the team's context (EmCtx).
template <int TW>
__device__ __noinline__ double estep_ctx(const EmCtx* cx, int lane) {
    const int G = cx->G, n = cx->n;
    const double* xs = cx->xs;
    const double* pm = cx->pm;
    // pointers that went through memory have lost their address space: say it again (LDS instead of generic loads)
    __builtin_assume(__isShared(pm));
    if (TW == 1) __builtin_assume(__isShared(xs));
    double* stage = TW == 1 ? nullptr : cx->stage;
    const int stage_k = TW == 1 ? 0 : cx->stage_k;
    double v[16];
    switch (G) {
        case 2: estep_pass<2, Unroll<2>::value, TW, false>(xs, n, lane, pm, v, stage, stage_k); break;
        case 3: estep_pass<3, Unroll<3>::value, TW, false>(xs, n, lane, pm, v, stage, stage_k); break;
        case 4: estep_pass<4, Unroll<4>::value, TW, false>(xs, n, lane, pm, v, stage, stage_k); break;
        default: estep_pass<5, Unroll<5>::value, TW, false>(xs, n, lane, pm, v, stage, stage_k); break;
    }
    return Team<TW>::reduce(v, lane, cx->red);
}
template <int TW>
__device__ __noinline__ double estep_ctx_sat(const EmCtx* cx, int lane) {
    return estep_dispatch_sat<TW>(cx->G, cx->xs, cx->n, lane, cx->pm, cx->red, TW == 1 ? nullptr : cx->stage,
                                  TW == 1 ? 0 : cx->stage_k);
}

template <int TW>
__device__ __noinline__ double label_dispatch(int G, const double* xs, int n, int lane, const double* pm,
                                              double* red) {
    double v[16];
    switch (G) {
        case 2: label_pass<2, TW>(xs, n, lane, pm, v); break;
        case 3: label_pass<3, TW>(xs, n, lane, pm, v); break;
        case 4: label_pass<4, TW>(xs, n, lane, pm, v); break;
        default: label_pass<5, TW>(xs, n, lane, pm, v); break;
    }
    return Team<TW>::reduce(v, lane, red);
}

// ---------------------------------------------------------------------------------------------
// Start of a G >= 2 row: mclust's qclass quantile classes and the first M-step from them (z = unmap(qclass)).
// One copy for every G.  Lane k (k < G) returns the parameters of component k; false = a class is empty
// ("mixing proportion fell below threshold": the row has no valid model).
// ---------------------------------------------------------------------------------------------
template <int TW>
__device__ __noinline__ bool em_start(int G, const double* xs, int n, int lane, double* red, double& mu_own,
                                      double& sig_own, double& pro_own) {
    typedef Team<TW> T;
    const double rteps = 1.4901161193847656e-08;
    const double dn = (double)n;
    const int tid = T::tid(lane);
    const int kown = lane < G ? lane : 0;
    const bool owner = lane < G;
    double br[GMAX];
    const int nbr = qclass_breaks(xs, n, G, br);
    double b0 = nbr > 0 ? br[0] : INFINITY, b1 = nbr > 1 ? br[1] : INFINITY;
    double b2 = nbr > 2 ? br[2] : INFINITY, b3 = nbr > 3 ? br[3] : INFINITY;
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
    for (int i = tid; i < n; i += T::NT) {
        double x = xs[i];
        int cl = (x >= b0) + (x >= b1) + (x >= b2) + (x >= b3);
#pragma unroll
        for (int k = 0; k < GMAX; k++) {
            if (k == cl) { v[k] += 1.0; v[GMAX + k] += x; }
        }
    }
    double t = T::reduce(v, lane, red);
    double cnt = slot(t, kown), sx = slot(t, GMAX + kown);
    if (__any_sync(FULLMASK, owner && cnt <= rteps)) return false;
    mu_own = sx / cnt;
    pro_own = cnt / dn;
    double m0 = __shfl_sync(FULLMASK, mu_own, 0), m1 = __shfl_sync(FULLMASK, mu_own, 1);
    double m2 = __shfl_sync(FULLMASK, mu_own, 2), m3 = __shfl_sync(FULLMASK, mu_own, 3);
    double m4 = __shfl_sync(FULLMASK, mu_own, 4);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
    for (int i = tid; i < n; i += T::NT) {
        double x = xs[i];
        int cl = (x >= b0) + (x >= b1) + (x >= b2) + (x >= b3);
        double m = cl == 0 ? m0 : cl == 1 ? m1 : cl == 2 ? m2 : cl == 3 ? m3 : m4;
        double d = x - m;
        double dd = d * d;
#pragma unroll
        for (int k = 0; k < GMAX; k++) {
            if (k == cl) v[k] += dd;
        }
    }
    t = T::reduce(v, lane, red);
    sig_own = slot(t, kown) / cnt;
    return true;
}

// ---------------------------------------------------------------------------------------------
// G >= 2 row: qclass start, EM (mclust me1v semantics: M-step first, stop when the relative change of
// the log-likelihood is <= 1e-5), BIC, extra M-step, MAP label statistics.
// One copy of this code serves every G: the scalar M-step runs once per warp with lane k owning component k
// (lanes >= G shadow component 0), only the point loops are specialised on G.
// ---------------------------------------------------------------------------------------------
template <int TW, bool BUDGET>
__device__ __noinline__ bool fit_row(int G, const double* xs, int n, int lane, Row* row, WarpStats* st,
                                     double* pm, double* red, int extra_mstep, int iter_cap) {
    typedef Team<TW> T;
    const double eps = DBL_EPSILON, rteps = 1.4901161193847656e-08, tol = 1e-5;
    const double dn = (double)n;
    const bool lead0 = T::lead() && lane == 0;
    if (lead0) { row->valid = 0; row->iters = 0; }
    if (G > n) return false;   // mclustBIC: G <- G[G <= n]
    const int kown = lane < G ? lane : 0;
    const bool owner = lane < G;
    const double range2 = (xs[n - 1] - xs[0]) * (xs[n - 1] - xs[0]);   // xs is sorted: every |x - mu_k| <= range

    double mu_own, sig_own, pro_own;
    if (!em_start<TW>(G, xs, n, lane, red, mu_own, sig_own, pro_own)) return false;

    double hold = 8.988465674311579e307, hood = 0.0;   // "previous" log-likelihood sentinel (FLMAX/2 in mclust)
    int iter = 0;
    double A = 0.0, B = 0.0, C = 0.0, nh_own = 0.0, lc_own = 0.0;
    const double inv_n = 1.0 / dn;
    bool failed = false, near_tol = false, capped = false;
    EmCtx* const cx = reinterpret_cast<EmCtx*>(pm + 4 * GMAX);
    T::sync();
    if (lead0) {
        cx->xs = xs; cx->pm = pm; cx->red = red; cx->n = n; cx->G = G; cx->iter = 0; cx->flags = 0;
        cx->hold = hold; cx->range2 = range2; cx->inv_n = inv_n;
    }
    if (owner && T::lead()) { pm[4 * lane] = mu_own; cx->sig[lane] = sig_own; cx->pro[lane] = pro_own; }
    T::sync();
    for (;;) {
        // ---- state of the owned component from shared memory, E-step constants ----
        mu_own = cx->pm[4 * kown]; sig_own = cx->sig[kown]; pro_own = cx->pro[kown];
        if (__any_sync(FULLMASK, owner && sig_own <= eps)) { failed = true; break; }   // "sigma-squared falls below threshold"
        const double rsig = fast_rcp(sig_own);
        nh_own = -0.5 * rsig;
        lc_own = 0.5 * fast_log(pro_own * pro_own * rsig) - 0.5 * PI2LOG;
        if (owner && T::lead()) { cx->pm[4 * lane + 1] = nh_own; cx->pm[4 * lane + 2] = lc_own; }
        T::sync();
        // |t_k| <= |lc_k| + range^2 / (2 sigsq_k) with |lc_k| < 100: below 2^19 for every point iff the smallest
        // variance exceeds range^2 / 2^20; then the integer extraction needs no saturation test (fast pass)
        const bool wide = !__any_sync(FULLMASK, owner && !(sig_own * 1.0e6 > cx->range2));
        const double t = wide ? estep_ctx<TW>(cx, lane) : estep_ctx_sat<TW>(cx, lane);
        {
            const int g = cx->G, k = lane < g ? lane : 0;
            A = slot(t, k); B = slot(t, g + k); C = slot(t, 2 * g + k);
            hood = slot(t, 3 * g);
        }
        hold = cx->hold;
        iter = cx->iter + 1;
        // stop when |hold - hood| / (1 + |hood|) <= tol  (tested in product form: no division on the critical path)
        const double diff = fabs(hold - hood), bound = tol * (1.0 + fabs(hood));
        const bool close = fabs(diff - bound) <= 1e-7 * bound;
        T::sync();                                  // every thread has read hold / iter
        if (T::lead() && lane == 0) { cx->hold = hood; cx->iter = iter; if (close) cx->flags |= 1; }
        if (!(diff > bound)) break;
        if (iter >= (BUDGET ? iter_cap : EM_ITER_CAP)) { capped = true; break; }
        // ---- M-step (lane k: component k) ----
        const bool own = lane < cx->G;
        if (__any_sync(FULLMASK, own && A <= rteps)) { failed = true; break; }
        const double ra = fast_rcp(A);
        const double dm = B * ra;
        if (own && T::lead()) {
            cx->pm[4 * lane] = cx->pm[4 * lane] + dm;
            cx->sig[lane] = fma(-dm, dm, C * ra);
            cx->pro[lane] = A * cx->inv_n;
        }
        T::sync();
    }
    T::sync();
    // the loop's registers are gone by now: the state that the epilogue needs comes back from shared memory
    mu_own = pm[4 * kown]; sig_own = cx->sig[kown]; pro_own = cx->pro[kown];
    nh_own = pm[4 * kown + 1]; lc_own = pm[4 * kown + 2];
    iter = cx->iter;
    near_tol = (cx->flags & 1) != 0;
    // stopped by the caller's iteration budget (not by the safety cap): nothing is recorded, the caller re-runs the
    // whole fit-set elsewhere (fit_kernel, "deferred" items)
    if (BUDGET && capped && iter_cap < EM_ITER_CAP) return true;
    if (lead0) {
        // work counters (per team, shared memory): iterations executed, points x components, points
        st->em_iters += (double)iter;
        st->triples += dn * (double)G * (double)iter;
        st->point_iters += dn * (double)iter;
        if (near_tol) st->tol_near += 1.0;
        if (capped) st->cap_hits += 1.0;
        row->iters = iter;
    }
    if (failed) return false;

    // ---- converged: BIC, parameters, labels ----
    double mean_out = mu_own, sig_out = sig_own;
    if (extra_mstep) {
        // summaryMclustBIC: if sum((pro - colMeans(z))^2) > sqrt(eps) replace $parameters by one more M-step
        double cm = A * inv_n;
        double term = owner ? (pro_own - cm) * (pro_own - cm) : 0.0;
        double dev = 0.0;
#pragma unroll
        for (int k = 0; k < GMAX; k++) dev += __shfl_sync(FULLMASK, term, k);
        bool ok = !__any_sync(FULLMASK, owner && !(A > 0.0));
        if (dev > rteps && ok) {
            double ra = fast_rcp(A);
            double dm = B * ra;
            mean_out = mu_own + dm;
            sig_out = fma(-dm, dm, C * ra);
        }
    }
    // MAP labels from the final z (same parameters as the last E-step) and their statistics
    const double t = label_dispatch<TW>(G, xs, n, lane, pm, red);   // pm still holds the last E-step's parameters
    double cnt = slot(t, kown), sy = slot(t, G + kown), syy = slot(t, 2 * G + kown);
    if (owner && T::lead()) {
        row->mean[lane] = mean_out;
        row->sigsq[lane] = sig_out;
        row->mu_em[lane] = mu_own;
        row->nh_em[lane] = nh_own;
        row->lc_em[lane] = lc_own;
        row->cnt[lane] = cnt;
        row->sy[lane] = sy;
        row->syy[lane] = syy;
    }
    if (lead0) {
        row->loglik = hood;
        row->bic = 2.0 * hood - (double)(3 * G - 1) * log(dn);
        row->valid = 1;
    }
    return false;
}

// G = 1 row: mvn1d closed form
template <int TW>
__device__ __forceinline__ void fit_row1(const double* xs, int n, int lane, Row* row, double* red) {
    typedef Team<TW> T;
    const int tid = T::tid(lane);
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
    for (int i = tid; i < n; i += T::NT) { double x = xs[i]; v[0] += x; v[1] = fma(x, x, v[1]); }
    double t = T::reduce(v, lane, red);
    const double sx = slot(t, 0), sxx = slot(t, 1);
    double dn = (double)n;
    double mu = sx / dn;
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
    for (int i = tid; i < n; i += T::NT) { double d = xs[i] - mu; v[0] = fma(d, d, v[0]); }
    t = T::reduce(v, lane, red);
    const double ss = slot(t, 0) / dn;
    if (T::lead() && lane == 0) {
        row->iters = 0;
        row->mean[0] = mu; row->sigsq[0] = ss;
        row->cnt[0] = dn; row->sy[0] = sx; row->syy[0] = sxx;
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
    #pragma unroll 1
    for (int g = 1; g <= k && g <= GMAX; g++) {
        if (!rows[g - 1].valid) continue;
        if (best == 0 || rows[g - 1].bic > rows[best - 1].bic) best = g;
    }
    return best;
}

__device__ __forceinline__ void table_of(const Row* rows, int comps, Tab& t) {
    t.nt = 0;
    #pragma unroll 1
    for (int k = 0; k < comps; k++)
        if (rows[comps - 1].cnt[k] > 0.0) { t.cnt[t.nt] = rows[comps - 1].cnt[k]; t.lab[t.nt] = k; t.nt++; }
}

__device__ __forceinline__ double tab_min(const Tab& t) {
    double m = t.cnt[0];
    #pragma unroll 1
    for (int i = 1; i < t.nt; i++) m = fmin(m, t.cnt[i]);
    return m;
}

__device__ __forceinline__ double r_min_nan(const double* v, int n) {   // R's min(): NaN propagates
    double m = v[0];
    #pragma unroll 1
    for (int i = 0; i < n; i++) { if (isnan(v[i])) return nan(""); if (v[i] < m) m = v[i]; }
    return m;
}

// diagnostics of a >= 2 component model: variance ratio and the min/max-variance component indices
__device__ __forceinline__ void var_diag(const Row& mc, int comps, int& maxcl, int& mincl, double& vardiff, double& sdmax) {
    maxcl = 0; mincl = 0;
    #pragma unroll 1
    for (int k = 1; k < comps; k++) {
        if (mc.sigsq[k] > mc.sigsq[maxcl]) maxcl = k;   // which.max / which.min: first extremum
        if (mc.sigsq[k] < mc.sigsq[mincl]) mincl = k;
    }
    vardiff = mc.sigsq[maxcl] / mc.sigsq[mincl];
    sdmax = 0.0;
    #pragma unroll 1
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
            #pragma unroll 1
            for (int j = 0; j < nmd; j++) meandiff[j] = (mc->mean[j + 1] - mc->mean[j]) / sdmax;   // :77
            double nmin = (mincl < tb.nt) ? tb.cnt[mincl] : NaN;       // :83  table(cl)[min.cl] is positional
            double mincat = tab_min(tb);                               // :84
            int tries = 0;
            double nmax = (maxcl < tb.nt) ? tb.cnt[maxcl] : NaN;       // :86
            double balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));   // :87
            #pragma unroll 1
            for (int j = 0; j < nmd; j++) meandiff[j] *= balance;      // :88
            if (isnan(vardiff)) vardiff = 1.0;                         // :91-110
            if (isnan(nmin)) nmin = INFINITY;
            {
                bool anyna = false;
                #pragma unroll 1
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
                    #pragma unroll 1
                    for (int j = 0; j < nmd; j++) {
                        double m1 = mc->sy[tb.lab[j + 1]] / tb.cnt[j + 1];
                        double m0 = mc->sy[tb.lab[j]] / tb.cnt[j];
                        meandiff[j] = (m1 - m0) / sdmax;               // :130
                    }
                    balance = 2.0 * sqrt(nmin * nmax / ((nmin + nmax) * (nmin + nmax)));   // :131 (stale nmin / nmax)
                    #pragma unroll 1
                    for (int j = 0; j < nmd; j++) meandiff[j] *= balance;
                    nmin = (mincl < tb.nt) ? tb.cnt[mincl] : NaN;      // :138
                } else {
                    meandiff[0] = 10.0; nmd = 1; vardiff = -INFINITY; nmin = 100.0;   // :140-142
                }
                if (isnan(vardiff)) { vardiff = 1.0; err = 1; }        // :146-173
                if (isnan(nmin)) { nmin = INFINITY; err = 1; }
                {
                    bool anyna = false;
                    #pragma unroll 1
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
        #pragma unroll 1
        for (int i = 0; i < tb.nt; i++) mx = fmax(mx, tb.cnt[i]);
        if (comps > 1 && mx < (double)cfg.min_size) {
            comps = pick_upto(rows, comps - 1);
        }
    }
    return comps;
}

// order(mc$parameters$mean), stable (R/mclust.restricted.R:243)
__device__ __forceinline__ void mean_order(const Row& mc, int comps, int* ord) {
    #pragma unroll 1
    for (int k = 0; k < comps; k++) ord[k] = k;
    #pragma unroll 1
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
// jittered: the fit ran on y + runif jitter (all y identical = yconst, R/mclust.restricted.R:48-50) but the score is
// taken on the original values: sum(y[class == k]) = nk * yconst, sum(y[class == k]^2) = nk * yconst^2.
__device__ __noinline__ double joint_posterior(const Row& mc, int comps, const DevCfg& cfg, bool jittered = false,
                                               double yconst = 0.0) {
    int ord[GMAX];
    mean_order(mc, comps, ord);
    double nk[GMAX], sy[GMAX], syy[GMAX];
    #pragma unroll 1
    for (int k = 0; k < GMAX; k++) { nk[k] = 0.0; sy[k] = 0.0; syy[k] = 0.0; }
    int r = 0;
    #pragma unroll 1
    for (int l = 0; l < comps; l++) {
        // cl2[cl == l] <- as.numeric(names(mean[order(mean)])[l])  (order(), not rank(): quirk viii)
        int nl = (comps > 1) ? ord[l] : l;
        nk[nl] = mc.cnt[l]; sy[nl] = mc.sy[l]; syy[nl] = mc.syy[l];
        if (jittered) { sy[nl] = __dmul_rn(mc.cnt[l], yconst); syy[nl] = __dmul_rn(mc.cnt[l], __dmul_rn(yconst, yconst)); }
        if (mc.cnt[l] > 0.0) r++;
    }
    const double alpha = cfg.alpha, m0 = cfg.m0, s0 = cfg.s0, a0 = cfg.a0, b0 = cfg.b0;
    double logpost = 0.0, J = 0.0;
    #pragma unroll 1
    for (int k = 0; k < r; k++) J += nk[k];
    #pragma unroll 1
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
// The whole fit-set on a sorted slab.  Returns the selected number of components (0 = failed), or FIT_DEFERRED when
// an EM needed more than iter_cap (< EM_ITER_CAP) iterations: nothing has been recorded and the caller re-runs the
// fit-set with a wider team.
// ---------------------------------------------------------------------------------------------
template <int TW, bool BUDGET = false>
__device__ __forceinline__ int fit_set(const double* xs, int n, int lane, Row* rows, const DevCfg& cfg, WarpStats* st,
                                       double* pm, double* red, int iter_cap = EM_ITER_CAP) {
    typedef Team<TW> T;
    fit_row1<TW>(xs, n, lane, &rows[0], red);
    for (int G = 2; G <= GMAX; G++)
        if (fit_row<TW, BUDGET>(G, xs, n, lane, &rows[G - 1], st, pm, red, cfg.extra_mstep, iter_cap)) return FIT_DEFERRED;
    T::sync();
    // near-tie diagnostic: two best BICs within 1e-9 relative
    bool bic_near = false;
    {
        double b1 = -INFINITY, b2 = -INFINITY;
        for (int g = 0; g < GMAX; g++) if (rows[g].valid) {
            double b = rows[g].bic;
            if (b > b1) { b2 = b1; b1 = b; } else if (b > b2) b2 = b;
        }
        if (b2 > -INFINITY && fabs(b1 - b2) <= 1e-9 * fabs(b1)) bic_near = true;
    }
    int comps = restricted_select(rows, n, cfg);
    if (T::lead() && lane == 0) {
        st->fit_sets += 1.0;
        if (bic_near) st->bic_near += 1.0;
        if (comps == 0) st->failed_sets += 1.0;
    }
    return comps;
}

// in-place ascending bitonic sort of np (power of two) doubles by a team
template <int TW>
__device__ __forceinline__ void team_sort(double* xs, int np, int lane) {
    typedef Team<TW> T;
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = T::tid(lane); t < (np >> 1); t += T::NT) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                bool up = ((i & k) == 0);
                double a = xs[i], b = xs[l];
                if ((a > b) == up) { xs[i] = b; xs[l] = a; }
            }
            T::sync();
        }
    }
}

// The same network for a slab in global memory (CTA teams on vectors beyond shared memory: config 5), tile by tile
// through shared memory: compare-exchanges at distance j < TILE stay inside one TILE-aligned block, so every stage
// with j < TILE runs on a tile resident in shared memory and only the stages with j >= TILE touch global memory
// (np = 65,536, TILE = 16,384: 3 global stages and 3 residencies per tile instead of 136 global stages; ncu had the
// all-global sort at 11 % of the config-5 kernel, 82 % of it long-scoreboard stalls).
template <int TW>
__device__ __forceinline__ void team_sort_tiled(double* xs, int np, int lane, double* tile, int TILE) {
    typedef Team<TW> T;
    const int tid = T::tid(lane);
    auto local = [&](int base, int k_lo, int k_hi, int j_top) {
        // tile [base, base + TILE): stages (k, j) for k = k_lo..k_hi, j from min(k/2, j_top) down to 1
        for (int t = tid; t < TILE; t += T::NT) tile[t] = xs[base + t];
        T::sync();
        for (int k = k_lo; k <= k_hi; k <<= 1) {
            for (int j = min(k >> 1, j_top); j > 0; j >>= 1) {
                for (int t = tid; t < (TILE >> 1); t += T::NT) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    const bool up = (((base + i) & k) == 0);
                    const double a = tile[i], b = tile[l];
                    if ((a > b) == up) { tile[i] = b; tile[l] = a; }
                }
                T::sync();
            }
        }
        for (int t = tid; t < TILE; t += T::NT) xs[base + t] = tile[t];
        T::sync();
    };
    for (int base = 0; base < np; base += TILE) local(base, 2, TILE, TILE >> 1);          // every k <= TILE
    for (int k = TILE << 1; k <= np; k <<= 1) {
        for (int j = k >> 1; j >= TILE; j >>= 1) {                                         // global stages
            for (int t = tid; t < (np >> 1); t += T::NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = ((i & k) == 0);
                const double a = xs[i], b = xs[l];
                if ((a > b) == up) { xs[i] = b; xs[l] = a; }
            }
            T::sync();
        }
        for (int base = 0; base < np; base += TILE) local(base, k, k, TILE >> 1);          // the stages j < TILE of this k
    }
}

}  // namespace scdd
