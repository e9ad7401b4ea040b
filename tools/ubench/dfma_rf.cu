// Is the fp64 pipe fed at full rate when every DFMA reads three distinct 64-bit registers (no operand reuse),
// alone and mixed with integer / select instructions?  16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int K>
__global__ void k(double* out, int iters, double seed) {
    double a[8], b[8], c[8];
    unsigned u[8], w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { a[j] = seed + j; b[j] = 0.999 + 1e-6 * j + 1e-9 * threadIdx.x; c[j] = 1e-9 * (j + 1); u[j] = threadIdx.x + j; w[j] = threadIdx.x * 3 + j; }
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma(a[j], 0.999999, 1e-9);                 // immediates / uniform operands
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma(b[j], a[j], 1e-9);                      // two distinct registers + immediate
#pragma unroll
            for (int j = 0; j < 8; j++) { double t = b[j]; b[j] = c[j]; c[j] = t; }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma(b[j], c[j], a[j]);                      // three distinct registers
#pragma unroll
            for (int j = 0; j < 8; j++) { double t = b[j]; b[j] = c[j]; c[j] = t; }       // rotate so nothing is loop invariant (renaming only)
        }
#pragma unroll
        for (int j = 0; j < K; j++) { u[j & 7] = (u[j & 7] ^ w[j & 7]) + w[(j + 1) & 7]; }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { s += a[j] + b[j] + c[j]; t ^= u[j]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <int MODE, int K>
void run(int sms, double* d, const char* name) {
    int iters = 1 << 13;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) { cudaEventRecord(a); k<MODE, K><<<sms, 512>>>(d, iters, 1.0 + r); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms; }
    double cyc = best * 1e-3 * 1.965e9 / iters;
    printf("%-28s K=%2d int ops: %6.1f cycles/iter/SMSP (32 DFMA + %d INT per SMSP-iter) -> %.2f cycles per DFMA\n", name, K, cyc, 8 * K, cyc / 32.0);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, sizeof(double) * sms * 512);
    run<0, 0>(sms, d, "DFMA reg,imm,imm"); run<1, 0>(sms, d, "DFMA reg,reg,reg");
    run<1, 4>(sms, d, "DFMA reg,reg,reg"); run<1, 8>(sms, d, "DFMA reg,reg,reg"); run<1, 12>(sms, d, "DFMA reg,reg,reg");
    run<0, 8>(sms, d, "DFMA reg,imm,imm");
    run<2, 0>(sms, d, "DFMA reg,reg,imm"); run<2, 4>(sms, d, "DFMA reg,reg,imm"); run<2, 8>(sms, d, "DFMA reg,reg,imm");
    return 0;
}
