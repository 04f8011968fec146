// How do non-FP64 instructions share issue bandwidth with FP64 ones on sm_100a?
// Each warp runs ITER iterations of (8 independent DFMA) + (M integer ops | M MUFU.RCP64H | M LDS).
// Prints cycles per iteration per SM sub-partition at full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KIND>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
    __shared__ double sh[256];
    sh[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    double r0 = a0, r1 = a1;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
#pragma unroll
        for (int k = 0; k < M; ++k) {
            if (KIND == 0) {          // integer ALU / IMAD
                if (k & 1) i0 = i0 * 3u + i1; else i2 = (i2 ^ i3) + 0x9e3779b9u;
                if ((k & 3) == 3) { i1 += i0; i3 ^= i2; }
            } else if (KIND == 1) {   // MUFU.RCP64H
                double y;
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(k & 1 ? r0 : r1));
                if (k & 1) r0 = y; else r1 = y;
            } else {                  // LDS.64
                r0 += sh[(i0 + k * 33) & 255];
            }
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + r0 + r1 + (double)(i0 + i1 + i2 + i3);
    if (s == 123.456) out[0] = s;
}

template <int M, int KIND>
static void run(const char* name, double* out, int sms) {
    const int iters = 20000, blocks = sms * 8;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    probe<M, KIND><<<blocks, 256>>>(out, 100, 1.0);
    cudaEventRecord(t0);
    probe<M, KIND><<<blocks, 256>>>(out, iters, 1.0);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    // warps per SM sub-partition = 8 blocks * 8 warps / 4 = 16; iterations per sub-partition = 16 * iters
    double cyc = ms * 1e-3 * khz * 1e3 / (16.0 * iters);
    printf("%-8s M=%2d : %.3f ms  %.2f cycles per (8 DFMA + %d other) per sub-partition [assuming %d MHz]\n", name, M, ms, cyc, M, khz / 1000);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    run<0, 0>("int", out, sms); run<2, 0>("int", out, sms); run<4, 0>("int", out, sms); run<8, 0>("int", out, sms);
    run<16, 0>("int", out, sms); run<32, 0>("int", out, sms);
    run<1, 1>("mufu", out, sms); run<2, 1>("mufu", out, sms); run<4, 1>("mufu", out, sms); run<8, 1>("mufu", out, sms);
    run<2, 2>("lds", out, sms); run<4, 2>("lds", out, sms); run<8, 2>("lds", out, sms);
    return 0;
}
