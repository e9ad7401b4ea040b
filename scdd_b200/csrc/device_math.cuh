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
// exp(a) for a in [-708, 0]  (the EM E-step only ever needs exp(t_k - max_k t_k); values below
// mclust's SMALOG = -708 clamp are replaced by 0 by the caller, so no subnormal handling is needed).
//
// exp(a) = 2^k * 2^(j/32) * exp(r),  32 k + j = rint(a * 32/ln2),  r = a - (32 k + j) ln2/32 (two-term
// Cody-Waite, |r| <= ln2/64), exp(r) - 1 by a degree-6 polynomial (truncation < 4e-18), 2^(j/32) from a
// 32-entry table (shared memory on the device).  11 fp64 operations instead of the 15 of a table-free
// degree-11 evaluation; the fp64 pipe is the bottleneck of the fit kernel.  Max error 0.9 ulp.
// ---------------------------------------------------------------------------------------------
#define SCDD_EXP_TABLE_INIT                                                                                   \
    {1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924, \
     1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,         \
     1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,       \
     1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,      \
     1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,         \
     1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,         \
     1.9571441241754002}

namespace expc {
constexpr double L2E32 = 46.16624130844683;            // 32 / ln 2
constexpr double MAGIC = 6755399441055744.0;           // 1.5 * 2^52: rint() in the low mantissa bits
constexpr double LN2_32_HI = 0.021660849392446835;     // ln2/32, low 17 mantissa bits zero
constexpr double LN2_32_LO = 5.145609244655338e-14;
constexpr double C2 = 0.5, C3 = 1.0 / 6.0, C4 = 1.0 / 24.0, C5 = 1.0 / 120.0, C6 = 1.0 / 720.0;
}  // namespace expc

SCDD_HD double exp_neg_tab(double a, const double* tab) {
    using namespace expc;
    double t = fma(a, L2E32, MAGIC);
    double kf = t - MAGIC;
    double r = fma(kf, -LN2_32_HI, a);
    r = fma(kf, -LN2_32_LO, r);
    double p = fma(C6, r, C5);
    p = fma(p, r, C4);
    p = fma(p, r, C3);
    p = fma(p, r, C2);
    p = fma(p, r, 1.0);
    double s = p * r;                       // exp(r) - 1
#if defined(__CUDA_ARCH__)
    int n = __double2loint(t);
    double T = tab[n & 31];
    double v = fma(T, s, T);
    return __hiloint2double(__double2hiint(v) + ((n >> 5) << 20), __double2loint(v));
#else
    int64_t tb;
    std::memcpy(&tb, &t, 8);
    int n = (int)(int32_t)(tb & 0xffffffff);
    double T = tab[n & 31];
    double v = std::fma(T, s, T);
    int64_t vb;
    std::memcpy(&vb, &v, 8);
    vb += (int64_t)(n >> 5) << 52;
    double out;
    std::memcpy(&out, &vb, 8);
    return out;
#endif
}

// Table-free variant: exp(a) = 2^k * exp(r), k = rint(a*log2(e)), r = a - k*ln2 in [-ln2/2, ln2/2],
// degree-11 polynomial from tools/gen_exp_poly.py (15 fp64 operations, no loads, max error 0.65 ulp).
namespace expp {
constexpr double L2E = 1.4426950408889634074;
constexpr double LN2_HI = 6.93147180369123816490e-01;  // fdlibm split: low 21 bits of hi are zero
constexpr double LN2_LO = 1.90821492927058770002e-10;
constexpr double C[10] = {2.510038549551032e-08, 2.7620088445409746e-07, 2.7557268459997064e-06,
                          2.4801521295954376e-05, 0.00019841269863053618, 0.0013888888917213717,
                          0.008333333333330062, 0.04166666666662413, 0.16666666666666669, 0.5000000000000001};
}  // namespace expp

SCDD_HD double exp_neg_poly(double a) {
    using namespace expp;
    const double MAGIC = expc::MAGIC;
    double t = fma(a, L2E, MAGIC);
    double kf = t - MAGIC;
    double r = fma(kf, -LN2_HI, a);
    r = fma(kf, -LN2_LO, r);
    double p = fma(2.510038549551032e-08, r, 2.7620088445409746e-07);
    p = fma(p, r, 2.7557268459997064e-06);
    p = fma(p, r, 2.4801521295954376e-05);
    p = fma(p, r, 0.00019841269863053618);
    p = fma(p, r, 0.0013888888917213717);
    p = fma(p, r, 0.008333333333330062);
    p = fma(p, r, 0.04166666666662413);
    p = fma(p, r, 0.16666666666666669);
    p = fma(p, r, 0.5000000000000001);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#if defined(__CUDA_ARCH__)
    int k = __double2loint(t);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
    int64_t tb, pb;
    std::memcpy(&tb, &t, 8);
    int k = (int)(int32_t)(tb & 0xffffffff);
    std::memcpy(&pb, &p, 8);
    pb += (int64_t)k << 52;
    double out;
    std::memcpy(&out, &pb, 8);
    return out;
#endif
}

#ifndef SCDD_EXP_TABLE
#define SCDD_EXP_TABLE 1
#endif
SCDD_HD double exp_neg(double a, const double* tab) {
#if SCDD_EXP_TABLE
    return exp_neg_tab(a, tab);
#else
    (void)tab;
    return exp_neg_poly(a);
#endif
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
