// DFMA throughput vs instruction-level parallelism and warps per SM (B200 microbenchmark).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, double seed) {
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = seed + j;
    const double m = 0.999999, c = 1e-9 * (threadIdx.x + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) a[j] = fma(a[j], m, c);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, int sms, double* d) {
    int iters = 1 << 14;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(a);
        k<ILP><<<sms, warps_per_sm * 32>>>(d, iters, 1.0 + r);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    double fl = 2.0 * ILP * (double)iters * sms * warps_per_sm * 32;
    // cycles per DFMA warp-instruction per SMSP at 1.965 GHz
    double cyc = best * 1e-3 * 1.965e9 / ((double)iters * ILP * (warps_per_sm / 4.0));
    printf("ILP %d warps/SM %2d : %7.2f TFLOP/s  (%.2f cycles per DFMA per SMSP-warp-slot)\n", ILP, warps_per_sm, fl / (best * 1e-3) / 1e12, cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* d; cudaMalloc(&d, sizeof(double) * sms * 1024);
    for (int w : {4, 8, 12, 16, 32}) {
        run<1>(w, sms, d); run<2>(w, sms, d); run<3>(w, sms, d); run<4>(w, sms, d); run<5>(w, sms, d); run<8>(w, sms, d);
    }
    return 0;
}
