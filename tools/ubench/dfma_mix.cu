// Does a DFMA block the SMSP issue port for one cycle or two?  8 independent DFMA chains per thread mixed with
// K independent integer (ALU-pipe) or shared-load (LSU) instructions per loop iteration; 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, int MODE>
__global__ void k(double* out, int iters, double seed, const double* gsrc) {
    __shared__ double sh[512];
    sh[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a[8];
    unsigned b[16];
    double ld = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = seed + j;
#pragma unroll
    for (int j = 0; j < 16; j++) b[j] = threadIdx.x + j;
    const double m = 0.999999, c = 1e-9 * (threadIdx.x + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma(a[j], m, c);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < K; j++) b[j & 15] = (b[j & 15] ^ (unsigned)i) + 0x9e3779b9u * (j + 1);   // LOP3 + IADD/IMAD
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh[(threadIdx.x + j) & 511]))); ld += v; }
        }
    }
    double s = ld;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) t ^= b[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <int K, int MODE>
void run(int sms, double* d) {
    int iters = 1 << 13;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        k<K, MODE><<<sms, 512>>>(d, iters, 1.0 + r, d);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    // cycles per loop iteration per SMSP (4 warps per SMSP): ideal fp64-bound = 8 DFMA * 2 cycles * 4 warps = 64
    double cyc = best * 1e-3 * 1.965e9 / iters;
    printf("mode %s K=%2d extra ops per 8 DFMA: %7.1f cycles/iter/SMSP (4 warps)  -> %.2f cycles per DFMA\n", MODE ? "LDS " : "INT ", K, cyc, cyc / 32.0);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 512);
    run<0, 0>(p.multiProcessorCount, d); run<2, 0>(p.multiProcessorCount, d); run<4, 0>(p.multiProcessorCount, d);
    run<8, 0>(p.multiProcessorCount, d); run<12, 0>(p.multiProcessorCount, d); run<16, 0>(p.multiProcessorCount, d);
    run<2, 1>(p.multiProcessorCount, d); run<4, 1>(p.multiProcessorCount, d); run<8, 1>(p.multiProcessorCount, d);
    return 0;
}
