// The real EM code (fit_row<1>, all G) run back to back on a fixed sorted vector per warp: separates the cost
// of the EM iterations themselves from the per-fit-set work around them (gather, sort, selection, PPM score).
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#include "../../scdd_b200/csrc/kernels.cuh"
using namespace scdd;
#ifndef WARPS
#define WARPS 16
#endif
__global__ void __launch_bounds__(WARPS * 32, 1) k(const double* src, const double* unsorted, int n, int np, int reps, double* out, double* counters, int gonly, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < SCDD_EXP_TABLE_SIZE; i += blockDim.x) EXP_TABLE_SH[i] = EXP_TABLE_GLOBAL[i];
    __syncthreads();
    unsigned char* region = smem_raw + warp * fit_smem_per_warp(np);
    double* pm = reinterpret_cast<double*>(region);
    Row* rows = reinterpret_cast<Row*>(pm + 4 * GMAX);
    WarpStats* st = reinterpret_cast<WarpStats*>(rows + GMAX);
    double* red = reinterpret_cast<double*>(st + 1);
    double* xs = reinterpret_cast<double*>(region + fit_region_head(1));
    if (lane == 0) *st = WarpStats{0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = lane; j < n; j += 32) xs[j] = src[j];
    __syncwarp();
    DevCfg cfg; cfg.alpha = 0.01; cfg.m0 = 0; cfg.s0 = 0.01; cfg.a0 = 0.01; cfg.b0 = 0.01; cfg.min_size = 3; cfg.restrict_ = 1; cfg.extra_mstep = 1;
    double acc = 0;
    for (int r = 0; r < reps; r++) {
        if (mode >= 1) {   // gather + sort as the real kernel does
            for (int j = lane; j < np; j += 32) xs[j] = j < n ? unsorted[(j * 7 + r) % n] : INFINITY;
            __syncwarp();
            team_sort<1>(xs, np, lane);
        }
        if (mode >= 2) {
            int comps = fit_set<1>(xs, n, lane, rows, cfg, st, pm, red);
            if (comps > 0) acc += joint_posterior(rows[comps - 1], comps, cfg);
        } else {
            for (int G = 2; G <= GMAX; G++) {
                if (gonly && G != gonly) continue;
                fit_row<1, false>(G, xs, n, lane, &rows[G - 1], st, pm, red, 1, EM_ITER_CAP);
            }
        }
        __syncwarp();
    }
    if (lane == 1) out[blockIdx.x * WARPS + warp] += acc;
    if (lane == 0) { atomicAdd(&counters[0], st->triples); atomicAdd(&counters[1], st->em_iters); out[blockIdx.x * WARPS + warp] = rows[4].bic; }
}
int main() {
    init_exp_table();
    int n = 880, np = 1024;
    std::mt19937 g(1); std::normal_distribution<double> nd(3.0, 1.0);
    std::vector<double> h(n); for (auto& v : h) v = nd(g); std::sort(h.begin(), h.end());
    double *d, *du, *out, *cnt; cudaMalloc(&d, n * 8); cudaMalloc(&du, n * 8);
    { std::vector<double> hu = h; std::shuffle(hu.begin(), hu.end(), g); cudaMemcpy(du, hu.data(), n * 8, cudaMemcpyHostToDevice); } cudaMalloc(&out, 148 * WARPS * 8); cudaMalloc(&cnt, 16);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    size_t smem = WARPS * fit_smem_per_warp(np);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int cfgi = 0; cfgi < 7; cfgi++) {
        int gonly = cfgi < 5 ? (cfgi == 0 ? 0 : cfgi + 1) : 0, mode = cfgi < 5 ? 0 : cfgi - 4;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f; double tr = 0, it = 0;
        for (int r = 0; r < 3; r++) {
            cudaMemset(cnt, 0, 16);
            cudaEventRecord(a); k<<<148, WARPS * 32, smem>>>(d, du, n, np, 20, out, cnt, gonly, mode); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms;
            double c[2]; cudaMemcpy(c, cnt, 16, cudaMemcpyDeviceToHost); tr = c[0]; it = c[1];
        }
        printf("[W%d] mode %d G=%d%s: %.2f ms, %.3e triples/s, %.0f EM iterations (%.1f per fit), err=%s\n", WARPS, mode, gonly, gonly ? "" : " (all)", best, tr / (best * 1e-3), it,
               it / (148.0 * WARPS * 20 * (gonly ? 1 : 4)), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
