// fedp.cuh -- device side of feDP (R/classify.dd.R:314-383), the test scDD() runs on every gene the permutation
// test did NOT call significant:
//   fedp_stats_kernel   warp per gene: the mean-shift statistic of log-nonzero expression between conditions
//                       (Welch t.test, :353, or the condition coefficient of lm(y ~ cdr + cond), :351) and the
//                       cluster-count / contingency-table bookkeeping of :327-329, :358-365
//   fisher_kernel       CTA per qualifying gene: Fisher's exact test of the r x 2 table (r = 2..5 non-outlier
//                       clusters x condition) by direct enumeration of all tables with the observed margins
// The t / F distribution tail (pt) is finished on the host, like the KS series.
#pragma once
#include <cstdint>

#include "../../include/scdd_b200.h"

namespace scdd {

struct FedpGene {
    double tstat, df;        // two-sided p = 2 pt(-|t|, df); NaN when the statistic is undefined
    int32_t c_oa, c_c1, c_c2;  // luOutlier of oa / c1 / c2 classes
    int32_t r;               // rows of the Fisher table (non-outlier oa clusters), 0 when feDP skips the test
    int32_t row[SCDD_GMAX];  // row totals
    int32_t a[SCDD_GMAX];    // observed counts in the reference-condition column
    int32_t c0;              // reference-condition column total
    int32_t flag;            // bit 0: a condition is constant (reference would jitter with runif); bit 1: table lost a column
};

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

__global__ void fedp_stats_kernel(const double* __restrict__ logy, const int64_t* __restrict__ off,
                                  const int32_t* __restrict__ cond, const int32_t* __restrict__ cls_oa,
                                  const int32_t* __restrict__ cls_c, const double* __restrict__ cdr, int ngenes,
                                  int min_size, FedpGene* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngenes; g += warps) {
        const int64_t o = off[g];
        const int n = (int)(off[g + 1] - o);
        // ---- pass 1: counts, sums, extreme values per condition, label counts
        double s0 = 0, s1 = 0, su = 0;
        double mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
        int n1 = 0;
        int cnt_oa[2][SCDD_GMAX], cnt_c[2][SCDD_GMAX];
#pragma unroll
        for (int k = 0; k < SCDD_GMAX; k++) cnt_oa[0][k] = cnt_oa[1][k] = cnt_c[0][k] = cnt_c[1][k] = 0;
        for (int i = lane; i < n; i += 32) {
            const double y = logy[o + i];
            const int c = cond[o + i] != 0;
            if (c) { s1 += y; n1++; mn1 = fmin(mn1, y); mx1 = fmax(mx1, y); }
            else { s0 += y; mn0 = fmin(mn0, y); mx0 = fmax(mx0, y); }
            if (cdr) su += cdr[o + i];
            const int lo = cls_oa[o + i] - 1, lc = cls_c[o + i] - 1;
#pragma unroll
            for (int k = 0; k < SCDD_GMAX; k++) {
                cnt_oa[0][k] += (lo == k) & (c == 0);
                cnt_oa[1][k] += (lo == k) & (c == 1);
                cnt_c[0][k] += (lc == k) & (c == 0);
                cnt_c[1][k] += (lc == k) & (c == 1);
            }
        }
        s0 = wsum(s0); s1 = wsum(s1); su = wsum(su);
        n1 = __reduce_add_sync(0xffffffffu, n1);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mn0 = fmin(mn0, __shfl_xor_sync(0xffffffffu, mn0, s)); mx0 = fmax(mx0, __shfl_xor_sync(0xffffffffu, mx0, s));
            mn1 = fmin(mn1, __shfl_xor_sync(0xffffffffu, mn1, s)); mx1 = fmax(mx1, __shfl_xor_sync(0xffffffffu, mx1, s));
        }
#pragma unroll
        for (int k = 0; k < SCDD_GMAX; k++) {
            cnt_oa[0][k] = __reduce_add_sync(0xffffffffu, cnt_oa[0][k]);
            cnt_oa[1][k] = __reduce_add_sync(0xffffffffu, cnt_oa[1][k]);
            cnt_c[0][k] = __reduce_add_sync(0xffffffffu, cnt_c[0][k]);
            cnt_c[1][k] = __reduce_add_sync(0xffffffffu, cnt_c[1][k]);
        }
        const int n0 = n - n1;
        const double m0 = s0 / n0, m1 = s1 / n1, mu = su / n, my = (s0 + s1) / n, mv = (double)n1 / n;
        // ---- pass 2: centred second moments (two-pass, with the first-moment correction term)
        double q0 = 0, q1 = 0, e0 = 0, e1 = 0;          // Welch
        double suu = 0, suv = 0, suy = 0, svy = 0, syy = 0, svv = 0;   // lm(y ~ u + v)
        for (int i = lane; i < n; i += 32) {
            const double y = logy[o + i];
            const int c = cond[o + i] != 0;
            const double d = y - (c ? m1 : m0);
            if (c) { q1 = fma(d, d, q1); e1 += d; } else { q0 = fma(d, d, q0); e0 += d; }
            if (cdr) {
                const double u = cdr[o + i] - mu, v = (double)c - mv, yy = y - my;
                suu = fma(u, u, suu); suv = fma(u, v, suv); suy = fma(u, yy, suy);
                svy = fma(v, yy, svy); syy = fma(yy, yy, syy); svv = fma(v, v, svv);
            }
        }
        q0 = wsum(q0); q1 = wsum(q1); e0 = wsum(e0); e1 = wsum(e1);
        if (cdr) { suu = wsum(suu); suv = wsum(suv); suy = wsum(suy); svy = wsum(svy); syy = wsum(syy); svv = wsum(svv); }
        if (lane == 0) {
            FedpGene r;
            r.flag = ((n0 > 0 && mn0 == mx0) || (n1 > 0 && mn1 == mx1)) ? 1 : 0;
            r.tstat = NAN; r.df = NAN;
            if (!cdr) {
                // stats::t.test(y ~ factor(cond)), var.equal = FALSE
                if (n0 >= 2 && n1 >= 2) {
                    const double v0 = (q0 - e0 * e0 / n0) / (n0 - 1), v1 = (q1 - e1 * e1 / n1) / (n1 - 1);
                    const double se0 = v0 / n0, se1 = v1 / n1, se = sqrt(se0 + se1);
                    r.df = (se0 + se1) * (se0 + se1) / (se0 * se0 / (n0 - 1) + se1 * se1 / (n1 - 1));
                    r.tstat = (m0 - m1) / se;
                }
            } else {
                // summary(lm(y ~ cdr0 + factor(cond)))$coef[3, ]: intercept removed by centring, 2 x 2 normal equations
                const double det = suu * svv - suv * suv;
                if (n > 3 && det > 0 && suu > 0) {
                    const double bu = (svv * suy - suv * svy) / det, bv = (suu * svy - suv * suy) / det;
                    const double rss = syy - bu * suy - bv * svy;
                    r.df = n - 3;
                    r.tstat = bv / sqrt(rss / (n - 3) * suu / det);
                }
            }
            int coa = 0, cc1 = 0, cc2 = 0;
#pragma unroll
            for (int k = 0; k < SCDD_GMAX; k++) {
                coa += (cnt_oa[0][k] + cnt_oa[1][k]) >= min_size;
                cc1 += cnt_c[0][k] >= min_size;
                cc2 += cnt_c[1][k] >= min_size;
            }
            r.c_oa = coa; r.c_c1 = cc1; r.c_c2 = cc2;
            r.r = 0; r.c0 = 0;
#pragma unroll
            for (int k = 0; k < SCDD_GMAX; k++) { r.row[k] = 0; r.a[k] = 0; }
            if (cc1 == cc2 && cc1 == coa && coa > 1) {          // R/classify.dd.R:358
                int rr = 0, c0 = 0, c1 = 0;
#pragma unroll
                for (int k = 0; k < SCDD_GMAX; k++) {
                    const int tot = cnt_oa[0][k] + cnt_oa[1][k];
                    if (tot >= min_size) {                      // findOutliers: non-outlying clusters
                        r.row[rr] = tot; r.a[rr] = cnt_oa[0][k];
                        c0 += cnt_oa[0][k]; c1 += cnt_oa[1][k];
                        rr++;
                    }
                }
                r.r = rr; r.c0 = c0;
                if (c0 == 0 || c1 == 0) { r.flag |= 2; r.r = 0; }   // table() would have one column: fisher.test stops
            }
            out[g] = r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fisher's exact test of an r x 2 table, r <= 5: p = sum of P(T) over all tables T with the observed margins and
// log P(T) <= log P(observed) + tol.  stats::fisher.test does the same sum -- through dhyper for 2 x 2 (relErr
// 1 + 1e-7) and through the FEXACT network algorithm otherwise (tolerance 3.45254e-7 on the log scale).
// Rows are ordered so that the largest is the dependent one and the second largest the inner loop; the remaining
// (smallest) rows span the outer tuples.  Work items = (outer tuple, 32-wide segment of the inner range), strided
// over the CTA's threads; lf[] = log-factorial table.
// ------------------------------------------------------------------------------------------------
constexpr int FISHER_SEG = 32;
constexpr int FISHER_THREADS = 256;

struct FisherJob {
    int32_t slot, r;          // slot: where this job's partial sum goes (large tables are split over several jobs)
    int32_t row[SCDD_GMAX];   // sorted ascending by total
    int32_t a[SCDD_GMAX];     // observed reference-column counts, same order
    int32_t c0;
    int32_t pad;
    double tol;
    long long item_begin, item_end;
};

__global__ void __launch_bounds__(FISHER_THREADS) fisher_kernel(const FisherJob* __restrict__ jobs, int njobs,
                                                                const double* __restrict__ lf,
                                                                double* __restrict__ p_out) {
    __shared__ double red[FISHER_THREADS];
    for (int j = blockIdx.x; j < njobs; j += gridDim.x) {
        const FisherJob jb = jobs[j];
        const int r = jb.r;
        int N = 0;
        double K = 0, lobs = 0;
        for (int i = 0; i < r; i++) {
            N += jb.row[i];
            K += lf[jb.row[i]];
            lobs -= lf[jb.a[i]] + lf[jb.row[i] - jb.a[i]];
        }
        K += lf[jb.c0] + lf[N - jb.c0] - lf[N];
        const double thr = lobs + jb.tol;     // compared without K on both sides
        const int Rin = jb.row[r - 2], Rdep = jb.row[r - 1];
        const long long nseg = (Rin + FISHER_SEG) / FISHER_SEG;
        double acc = 0;
        for (long long it = jb.item_begin + threadIdx.x; it < jb.item_end; it += FISHER_THREADS) {
            long long t = it / nseg;
            const int seg = (int)(it - t * nseg);
            int rem = jb.c0;
            double part = 0;
            for (int i = 0; i < r - 2; i++) {
                const int radix = jb.row[i] + 1;
                const int ai = (int)(t % radix);
                t /= radix;
                rem -= ai;
                part -= lf[ai] + lf[jb.row[i] - ai];
            }
            if (rem < 0) continue;
            int lo = max(max(0, rem - Rdep), seg * FISHER_SEG);
            int hi = min(min(Rin, rem), seg * FISHER_SEG + FISHER_SEG - 1);
            for (int a = lo; a <= hi; a++) {
                const int b = rem - a;
                const double lp = part - (lf[a] + lf[Rin - a]) - (lf[b] + lf[Rdep - b]);
                if (lp <= thr) acc += exp(lp + K);
            }
        }
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int s = FISHER_THREADS / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) p_out[jb.slot] = red[0];
        __syncthreads();
    }
}

}  // namespace scdd
