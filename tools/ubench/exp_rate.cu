// Throughput of the kernels' column-wise exp (exp_neg_vec<N>) and of a full E-step point block in isolation,
// 16 warps per SM: cycles per fp64 instruction achieved vs the 2.05-cycle pipe limit.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../scdd_b200/csrc/fitset.cuh"
using namespace scdd;
template <int N>
__global__ void k_exp(double* out, int iters, double seed) {
    __shared__ double tab[32];
    if (threadIdx.x < 32) tab[threadIdx.x] = EXP_TABLE_GLOBAL[threadIdx.x];
    __syncthreads();
    double a[N], e[N], acc[N];
    for (int j = 0; j < N; j++) { a[j] = -seed * (j + 1) - 1e-3 * threadIdx.x; acc[j] = 0; }
    for (int i = 0; i < iters; i++) {
        exp_neg_vec<N>(a, e, tab);
        for (int j = 0; j < N; j++) { acc[j] += e[j]; a[j] = a[j] * 0.999; }
    }
    double s = 0;
    for (int j = 0; j < N; j++) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int G, int U>
__global__ void k_point(double* out, int iters, double seed) {
    __shared__ double tab[32];
    __shared__ __align__(16) double pm[4 * GMAX];
    if (threadIdx.x < 32) tab[threadIdx.x] = EXP_TABLE_GLOBAL[threadIdx.x];
    if (threadIdx.x < GMAX) { pm[4 * threadIdx.x] = 1.0 + threadIdx.x; pm[4 * threadIdx.x + 1] = -0.5 / (0.3 + 0.1 * threadIdx.x); pm[4 * threadIdx.x + 2] = -1.0 - 0.1 * threadIdx.x; }
    __syncthreads();
    const auto prm = load_params<G>(pm, tab);
    double A[G], B[G], C[G], pl = 1.0, hl = 0.0;
    for (int k = 0; k < G; k++) { A[k] = B[k] = C[k] = 0; }
    double x[U];
    for (int u = 0; u < U; u++) x[u] = seed + 0.01 * threadIdx.x + u;
    for (int i = 0; i < iters; i++) {
        accumulate_points<G, U>(x, prm, A, B, C, pl, hl);
        for (int u = 0; u < U; u++) x[u] = x[u] * 0.9999 + 1e-4;
        if ((i & 63) == 63) { hl += log(pl); pl = 1.0; }
    }
    double s = hl + pl;
    for (int k = 0; k < G; k++) s += A[k] + B[k] + C[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) { cudaEventRecord(a); f(r); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double* d; cudaMalloc(&d, sizeof(double) * sms * 512);
    int iters = 4096;
    auto rep = [&](const char* name, float ms, double per_iter_units, const char* unit) {
        double cyc = ms * 1e-3 * 1.965e9 / iters;   // cycles per iteration per SMSP (4 warps)
        printf("%-18s %8.1f cycles/iter/SMSP  = %.1f cycles per %s per warp-slot\n", name, cyc, cyc / 4.0 / per_iter_units, unit);
    };
    rep("exp_vec<2>", timeit([&](int r) { k_exp<2><<<sms, 512>>>(d, iters, 1.0 + r); }), 2, "exp");
    rep("exp_vec<3>", timeit([&](int r) { k_exp<3><<<sms, 512>>>(d, iters, 1.0 + r); }), 3, "exp");
    rep("exp_vec<5>", timeit([&](int r) { k_exp<5><<<sms, 512>>>(d, iters, 1.0 + r); }), 5, "exp");
    rep("point<2,1>", timeit([&](int r) { k_point<2, 1><<<sms, 512>>>(d, iters, 1.0 + r); }), 2, "triple");
    rep("point<2,2>", timeit([&](int r) { k_point<2, 2><<<sms, 512>>>(d, iters, 1.0 + r); }), 4, "triple");
    rep("point<3,1>", timeit([&](int r) { k_point<3, 1><<<sms, 512>>>(d, iters, 1.0 + r); }), 3, "triple");
    rep("point<4,1>", timeit([&](int r) { k_point<4, 1><<<sms, 512>>>(d, iters, 1.0 + r); }), 4, "triple");
    rep("point<5,1>", timeit([&](int r) { k_point<5, 1><<<sms, 512>>>(d, iters, 1.0 + r); }), 5, "triple");
    return 0;
}
