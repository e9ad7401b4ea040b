// rrng.h -- R's random number generation, re-implemented for host and device.
//
// What the reference consumes on the hot path is `sample(1:n)` (R/permutations.R:71,164,243,273),
// i.e. base R's do_sample without replacement driven by R_unif_index (sample.kind = "Rejection",
// the default since R 3.6.0) on top of unif_rand().  Inside BiocParallel::bplapply every element
// runs under its own L'Ecuyer-CMRG (MRG32k3a) stream, so that generator is the one needed on the
// device; Mersenne-Twister (R's default kind) is provided on the host for serial use.
// Algorithms follow R's documented RNG (src/main/RNG.c, random.c; L'Ecuyer 1999) -- SURVEY App. B.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SCDD_RNG_HD __host__ __device__ __forceinline__
#else
#define SCDD_RNG_HD inline
#endif

namespace scdd {

struct Lecuyer {
    uint32_t s[6];  // .Random.seed[2:7]
};

constexpr int64_t MRG_M1 = 4294967087LL;
constexpr int64_t MRG_M2 = 4294944443LL;

// keeps a uniform strictly inside (0, 1), as R's fixup()
SCDD_RNG_HD double r_fixup(double x) {
    const double i2_32m1 = 2.328306437080797e-10;
    if (x <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
}

// one MRG32k3a step -> unif_rand() of kind "L'Ecuyer-CMRG"
SCDD_RNG_HD double lecuyer_unif(Lecuyer& g) {
    int64_t p1 = 1403580LL * (int64_t)g.s[1] - 810728LL * (int64_t)g.s[0];
    p1 %= MRG_M1;
    if (p1 < 0) p1 += MRG_M1;
    g.s[0] = g.s[1]; g.s[1] = g.s[2]; g.s[2] = (uint32_t)p1;
    int64_t p2 = 527612LL * (int64_t)g.s[5] - 1370589LL * (int64_t)g.s[3];
    p2 %= MRG_M2;
    if (p2 < 0) p2 += MRG_M2;
    g.s[3] = g.s[4]; g.s[4] = g.s[5]; g.s[5] = (uint32_t)p2;
    int64_t d = (p1 > p2) ? (p1 - p2) : (p1 - p2 + MRG_M1);
    return r_fixup((double)d * 2.328306549295727688e-10);
}

// number of random bits R_unif_index asks for: ceil(log2(m))
SCDD_RNG_HD int index_bits(uint32_t m) {
    if (m <= 1) return 0;
#if defined(__CUDA_ARCH__)
    return 32 - __clz((int)(m - 1));
#else
    return 32 - __builtin_clz(m - 1);
#endif
}

// R_unif_index(m): rejection sampling on `bits` random bits assembled from 16-bit chunks
template <class Gen>
SCDD_RNG_HD uint32_t r_unif_index(Gen& g, uint32_t m) {
    int bits = index_bits(m);
    uint64_t mask = (bits >= 64) ? ~0ULL : ((1ULL << bits) - 1ULL);
    for (;;) {
        uint64_t v = 0;
        for (int nb = 0; nb <= bits; nb += 16) {
            int v1 = (int)floor(g.unif() * 65536.0);
            v = 65536ULL * v + (uint64_t)v1;
        }
        v &= mask;
        if (v < (uint64_t)m) return (uint32_t)v;
    }
}

struct LecuyerGen {
    Lecuyer st;
    SCDD_RNG_HD double unif() { return lecuyer_unif(st); }
};

// sample(1:n): partial Fisher-Yates of do_sample (replace = FALSE, size = n).
// scratch[n] is working storage; out receives 0-BASED indices (R's value minus 1).
template <class Gen, class IdxT>
SCDD_RNG_HD void r_sample_perm(Gen& g, uint32_t n, IdxT* scratch, IdxT* out) {
    for (uint32_t i = 0; i < n; i++) scratch[i] = (IdxT)i;
    uint32_t m = n;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t j = r_unif_index(g, m);
        out[i] = scratch[j];
        scratch[j] = scratch[--m];
    }
}

}  // namespace scdd
