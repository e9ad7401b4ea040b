// classify.cuh -- device side of classifyDD's Monte-Carlo comparison of cluster means (R/classify.dd.R:165-187).
//
// For every pair (cluster a of condition 1, cluster b of condition 2) the reference draws 10,000 variates from each
// cluster's posterior-predictive Student t,
//     rt(10000, df = round(ak)) * sqrt(bk / (ak sk)) + mk,
// and calls the means "different" iff all 10,000 sampled differences share a sign (min > 0 or max < 0).  That is 2e4
// t variates per comparison and ~1e8 per scDD() call on config 3 -- seconds in R or numpy, and after the permutation
// test moved to the device the largest remaining piece of the call.  Here one warp owns one comparison: every lane
// draws its share of the 10,000 pairs from its own Philox-4x32-10 stream (subsequence = 32 comparison + lane, so the
// result depends on the seed and on the comparison's position only -- not on the grid or the device count), and the
// warp stops as soon as both signs have been seen (the answer is then known).
//
// The draws have the reference's distribution, not R's stream: classifyDD consumes the USER's RNG (Mersenne-Twister
// + inversion by default), whose rt() = norm_rand() / sqrt(rchisq(df) / df) cannot be reproduced bit for bit without
// R's sources.  A comparison that sits at the threshold can therefore come out differently -- exactly as between two
// R sessions with different seeds.  t = Z / sqrt(Gamma(df/2) / (df/2)), the gamma variate by Marsaglia & Tsang (2000).
#pragma once
#include <cstdint>

#include <curand_kernel.h>

namespace scdd {

struct McPair {            // one comparison: d = (t1 scale1 + m1) - (t2 scale2 + m2)
    double df1, scale1, m1;
    double df2, scale2, m2;
};

// Gamma(alpha, 1) for alpha >= 1 (Marsaglia-Tsang squeeze); alpha < 1 is boosted by U^(1/alpha)
__device__ __forceinline__ double gamma_variate(curandStatePhilox4_32_10_t* st, double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    double v, x, u;
    for (;;) {
        x = curand_normal_double(st);
        v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        u = curand_uniform_double(st);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) break;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) break;
    }
    double g = d * v;
    if (alpha < 1.0) g *= pow(curand_uniform_double(st), 1.0 / alpha);
    return g;
}

__device__ __forceinline__ double student_t(curandStatePhilox4_32_10_t* st, double df) {
    const double z = curand_normal_double(st);
    const double h = 0.5 * df;
    return z / sqrt(gamma_variate(st, h) / h);
}

// flags[c] = 1 iff every sampled difference of comparison c has the same (strict) sign; lohi (nullable) [2c], [2c+1]
// = min and max of the differences actually drawn (the loop stops early once both signs occurred)
__global__ void classify_mc_kernel(const McPair* __restrict__ pairs, int ncomp, int ndraws, unsigned long long seed,
                                   int32_t* __restrict__ flags, double* __restrict__ lohi) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < ncomp; c += warps) {
        const McPair p = pairs[c];
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)c * 32ULL + (unsigned long long)lane, 0ULL, &st);
        double mn = INFINITY, mx = -INFINITY;
        const int rounds = (ndraws + 31) / 32;
        for (int r = 0; r < rounds; r++) {
            if (r * 32 + lane < ndraws) {
                const double d = (student_t(&st, p.df1) * p.scale1 + p.m1) - (student_t(&st, p.df2) * p.scale2 + p.m2);
                mn = fmin(mn, d);
                mx = fmax(mx, d);
            }
            if ((r & 7) == 7) {   // both signs (or a zero) seen somewhere in the warp: decided
                const bool neg = __any_sync(0xffffffffu, mn <= 0.0), pos = __any_sync(0xffffffffu, mx >= 0.0);
                if (neg && pos) break;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, s));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
        }
        if (lane == 0) {
            flags[c] = (mn > 0.0 || mx < 0.0) ? 1 : 0;     // R/classify.dd.R:182-183
            if (lohi) { lohi[2 * c] = mn; lohi[2 * c + 1] = mx; }
        }
    }
}

// raw variates for the distribution test: out[i] = t_df, i < n (thread i, subsequence i)
__global__ void student_t_kernel(double df, int n, unsigned long long seed, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)i, 0ULL, &st);
    out[i] = student_t(&st, df);
}

}  // namespace scdd
