// device_math.cuh -- fp64 building blocks shared by the kernels (and compiled for the host by
// the self-test entry points so they can be checked without a GPU).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define SCDD_HD __host__ __device__ __forceinline__
#else
#define SCDD_HD inline
#endif

namespace scdd {

// ---------------------------------------------------------------------------------------------
// The E-step's softmax terms.  For the log-terms t_k of one point (k = 1..G) the kernels need
// e_k = exp(t_k - shift) for ANY common shift that keeps the sum in range (log-sum-exp is shift invariant).
//   N_k = rint(t_k * 2^TB/ln2)  (TB = 11)                 (magic-number rounding; also the exp range reduction)
//   H_k = N_k >> TB,  j_k = N_k & (2^TB - 1)              (whole powers of two / table index: N_k = 2^TB H_k + j_k)
//   Mh  = max_k H_k                          (integer max: the shift is Mh ln2, the max rounded DOWN to a whole power of 2)
//   r_k = t_k - N_k * c,  c = fl(ln2/2^TB)   (ONE fma: the error N_k (ln2/2^TB - c) splits into 2^TB Mh (..), common to
//                                             all k and cancelled exactly when the log-likelihood adds the shift back as
//                                             Mh * 2^TB c with the same c, and a remainder <= 1e-16 for every term that matters)
//   e_k = 2^(H_k - Mh) * T[j_k] * (1 + s(r_k)),  the largest term lies in [1, 2), the sum in [1, 2G)
// s(r) = exp(r) - 1 by a degree-3 polynomial on |r| <= ln2/4096 (truncation 3.4e-17), T = 2048-entry table of
// 2^(j/2048) (16 KB of shared memory per CTA).  Because the shift is a multiple of 2^TB the table index j_k does not
// depend on the maximum: the lookup is issued as soon as N_k exists and its latency hides under the polynomial
// (round 1 shifted by max N_k itself and stalled on the lookup at the end of the integer chain).
// fp64 cost per term: 1 (t) + 1 (kf) + 1 (r) + 2 (polynomial) + 1 (s) + 1 (table fma) = 7; max, shift, table index
// and 2^k scaling are integer instructions.  These helpers are the single source of the arithmetic: the kernels
// issue them column-wise over the G (or 2G) terms of a point, the host self-test calls them term by term.
// ---------------------------------------------------------------------------------------------
#define SCDD_EXP_TABLE_BITS 11
#define SCDD_EXP_TABLE_SIZE (1 << SCDD_EXP_TABLE_BITS)

namespace expc {
constexpr int TB = SCDD_EXP_TABLE_BITS;
constexpr int TMASK = SCDD_EXP_TABLE_SIZE - 1;
constexpr double L2E_T = 2954.639443740597;            // 2^TB / ln 2
constexpr double MAGIC = 6755399441055744.0;           // 1.5 * 2^52: rint() in the low mantissa bits
constexpr double LN2_T = 0.00033845077175778578;       // fl(ln2 / 2^TB)
constexpr double C2 = 0.5, C3 = 1.0 / 6.0;
constexpr int NSAT = -(1 << 30);                       // N of a log-term below -2^18 ("minus infinity")
constexpr int HSAT = NSAT >> SCDD_EXP_TABLE_BITS;      // its H = N >> TB
constexpr int KMIN = -1021;                            // smallest power of two kept (the reference zeroes below e^-708)
constexpr double LN2_STEP = 0.00033845077175778578 * SCDD_EXP_TABLE_SIZE;   // 2^TB * LN2_T, exact: the log of one unit of Mh
}  // namespace expc

// 2^(j / 2^TB), j = 0 .. 2^TB - 1, correctly rounded from long double (host; uploaded to the device once)
inline const double* exp_table_host() {
    static double tab[SCDD_EXP_TABLE_SIZE];
    static bool done = false;
    if (!done) {
        for (int j = 0; j < SCDD_EXP_TABLE_SIZE; j++) tab[j] = (double)exp2l((long double)j / (long double)SCDD_EXP_TABLE_SIZE);
        done = true;
    }
    return tab;
}

SCDD_HD int hi_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
SCDD_HD int lo_word(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    int64_t b; std::memcpy(&b, &x, 8); return (int)(int32_t)(b & 0xffffffff);
#endif
}
SCDD_HD double from_words(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x;
#endif
}

SCDD_HD double term_t(double a) { return fma(a, expc::L2E_T, expc::MAGIC); }
// low word of t = N (two's complement) while |N| < 2^31.  Log-terms a <= -2^18 saturate to NSAT = -2^30: every
// unsaturated term has N = rint(a 2^11/ln2) > -2^18 * 2954.7 > -2^30, so a saturated ("minus infinity") term can
// never out-rank a finite one in the integer max (round 1 saturated from -2^19, where N already reaches -1.5e9).
SCDD_HD int term_n(double a, double t) { return ((unsigned)hi_word(a) >= 0xC1100000u) ? expc::NSAT : lo_word(t); }
SCDD_HD double term_r(double a, double t) { return fma(t - expc::MAGIC, -expc::LN2_T, a); }
SCDD_HD double term_s(double r) {
    using namespace expc;
    double p = fma(C3, r, C2);
    p = fma(p, r, 1.0);
    return p * r;
}
SCDD_HD int term_h(int n) { return n >> expc::TB; }
SCDD_HD double term_e(double s, int n, int Mh, const double* tab) {
    int k = (n >> expc::TB) - Mh;
    if (k < expc::KMIN) k = expc::KMIN;
    const double T = tab[n & expc::TMASK];
    const double v = fma(T, s, T);
    return from_words(hi_word(v) + (k << 20), lo_word(v));
}

// ---------------------------------------------------------------------------------------------
// log(x) for positive NORMAL x, table driven (Tang): x = 2^E m, m in [1, 2); the top 8 fraction bits of m pick
// c_j = 1 + (j + 1/2)/256; u = m rc_j - 1 (one fma, |u| <= 2^-9 + ulp) with rc_j = fl(1/c_j) and L_j = -log(rc_j)
// tabulated in double from long double; log(x) = E ln2 + L_j + log1p(u), log1p by its series through u^5
// (truncation u^6/6 < 1e-17).  ~1 ulp of the result; 9 fp64 + 8 other instructions, no branches.  The EM loop takes
// two logarithms per iteration (the running product of the softmax sums, and log(pro^2 / sigsq) of the M-step):
// CUDA's log() costs ~75 instructions per call with its range checks.
// ---------------------------------------------------------------------------------------------
#define SCDD_LOG_TABLE_SIZE 256
struct LogTab { double rc, l; };

inline const LogTab* log_table_host() {
    static LogTab tab[SCDD_LOG_TABLE_SIZE];
    static bool done = false;
    if (!done) {
        for (int j = 0; j < SCDD_LOG_TABLE_SIZE; j++) {
            const long double c = 1.0L + ((long double)j + 0.5L) / (long double)SCDD_LOG_TABLE_SIZE;
            tab[j].rc = (double)(1.0L / c);
            tab[j].l = (double)(-logl((long double)tab[j].rc));
        }
        done = true;
    }
    return tab;
}

SCDD_HD double fast_log_core(double x, double rc, double l) {
    const int hi = hi_word(x);
    const double m = from_words((hi & 0x000fffff) | 0x3ff00000, lo_word(x));
    const double u = fma(m, rc, -1.0);
    double p = fma(u, 0.2, -0.25);
    p = fma(p, u, 0.33333333333333331);
    p = fma(p, u, -0.5);
    p = p * u;
    const double r = fma(u, p, u);
    return fma((double)((hi >> 20) - 1023), 0.69314718055994529, l) + r;
}
SCDD_HD int fast_log_index(double x) { return (hi_word(x) >> 12) & (SCDD_LOG_TABLE_SIZE - 1); }
SCDD_HD double fast_log_tab(double x, const LogTab* tab) {
    const LogTab t = tab[fast_log_index(x)];
    return fast_log_core(x, t.rc, t.l);
}

// ---------------------------------------------------------------------------------------------
// x87 80-bit extended precision accumulator (64-bit significand, round-to-nearest-even), as R's
// cumsum() uses on x86-64 (LDOUBLE accumulator, cumsum.c), emulated with integers so that the KS
// statistic can be reproduced bit for bit on the device.
// Value = (-1)^neg * mant * 2^(exp - 63), mant normalised (top bit set) unless zero.
// ---------------------------------------------------------------------------------------------
struct Ext80 {
    uint64_t mant;
    int32_t exp;   // unbiased exponent of the leading bit
    int32_t neg;
};

SCDD_HD Ext80 ext_zero() {
    Ext80 z;
    z.mant = 0; z.exp = -100000; z.neg = 0;
    return z;
}

SCDD_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// finite, non-subnormal-or-subnormal double -> Ext80 (exact)
SCDD_HD Ext80 ext_from_double(double d) {
    uint64_t b;
#if defined(__CUDA_ARCH__)
    b = (uint64_t)__double_as_longlong(d);
#else
    std::memcpy(&b, &d, 8);
#endif
    Ext80 e;
    e.neg = (int32_t)(b >> 63);
    int bexp = (int)((b >> 52) & 0x7ff);
    uint64_t frac = b & 0xfffffffffffffULL;
    if (bexp == 0) {
        if (frac == 0) return ext_zero();
        int lz = clz64(frac);
        e.mant = frac << lz;
        e.exp = -1022 - 52 + (63 - lz);
        return e;
    }
    e.mant = (frac | (1ULL << 52)) << 11;
    e.exp = bexp - 1023;
    return e;
}

// a + b rounded to a 64-bit significand (round to nearest, ties to even) -- one x87 FADD
SCDD_HD Ext80 ext_add(Ext80 a, Ext80 b) {
    if (a.mant == 0) return b;
    if (b.mant == 0) return a;
    // make |a| >= |b|
    if (b.exp > a.exp || (b.exp == a.exp && b.mant > a.mant)) { Ext80 t = a; a = b; b = t; }
    int sh = a.exp - b.exp;
    // 128-bit working significands: value bits in the high word, 64 guard bits below
    uint64_t ah = a.mant, al = 0, bh, bl;
    if (sh == 0) { bh = b.mant; bl = 0; }
    else if (sh < 64) { bh = b.mant >> sh; bl = b.mant << (64 - sh); }
    else if (sh == 64) { bh = 0; bl = b.mant; }
    else if (sh < 128) { bh = 0; bl = b.mant >> (sh - 64); if (b.mant << (128 - sh)) bl |= 1; }  // sticky
    else { bh = 0; bl = 1; }  // pure sticky
    Ext80 r;
    r.neg = a.neg;
    int exp = a.exp;
    uint64_t rh, rl;
    if (a.neg == b.neg) {
        rl = al + bl;
        uint64_t c = rl < al;
        rh = ah + bh;
        uint64_t c2 = rh < ah;
        rh += c;
        c2 |= (rh < c);
        if (c2) {  // carry out: shift right by one, keep sticky
            uint64_t st = rl & 1;
            rl = (rl >> 1) | (rh << 63) | st;
            rh = (rh >> 1) | (1ULL << 63);
            exp += 1;
        }
    } else {
        uint64_t br = al < bl;
        rl = al - bl;
        rh = ah - bh - br;
        if (rh == 0 && rl == 0) return ext_zero();  // exact cancellation: +0 in round-to-nearest
        // normalise left
        if (rh == 0) { rh = rl; rl = 0; exp -= 64; }
        int lz = clz64(rh);
        if (lz) {
            rh = (rh << lz) | (rl >> (64 - lz));
            rl <<= lz;
            exp -= lz;
        }
    }
    // round to nearest even on the 64 guard bits
    const uint64_t HALF = 1ULL << 63;
    if (rl > HALF || (rl == HALF && (rh & 1))) {
        rh += 1;
        if (rh == 0) { rh = HALF; exp += 1; }
    }
    r.mant = rh;
    r.exp = exp;
    return r;
}

// (double) of an extended value: round the 64-bit significand to 53 bits, nearest even.
// Magnitudes met here are within the normal double range.
SCDD_HD double ext_to_double(Ext80 e) {
    if (e.mant == 0) return 0.0;
    uint64_t m = e.mant >> 11;
    uint64_t rem = e.mant & 0x7ff;
    int exp = e.exp;
    if (rem > 0x400 || (rem == 0x400 && (m & 1))) {
        m += 1;
        if (m == (1ULL << 53)) { m >>= 1; exp += 1; }
    }
    uint64_t b = ((uint64_t)e.neg << 63) | ((uint64_t)(exp + 1023) << 52) | (m & 0xfffffffffffffULL);
    double d;
#if defined(__CUDA_ARCH__)
    d = __longlong_as_double((long long)b);
#else
    std::memcpy(&d, &b, 8);
#endif
    return d;
}

}  // namespace scdd
