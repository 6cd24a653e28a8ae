// fp64_mix.cu -- how many non-FP64 instructions can ride along with DFMA on sm_100a?
// 8 independent DFMA chains; after every DFMA, K copies of one "other" instruction (inline PTX, volatile,
// independent chains).  Reports cycles per DFMA per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int K>
__global__ void __launch_bounds__(256) mix(double* out, int* iout, int iters, double a, double b, int seed) {
    double x[8]; int q[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; q[i] = threadIdx.x + i + seed; f[i] = q[i]; }
    int pz = seed & 1;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (MODE == 1) asm volatile("{.reg .pred p; setp.ne.s32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(pz));
                    if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(seed));
                    if (MODE == 3) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(q[i]) : "r"(seed), "r"(q[(i + 3) & 7]));
                    if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(f[(i + 1) & 7]));
                    if (MODE == 5) asm volatile("mov.b32 %0, %1;" : "=r"(q[i]) : "r"(q[(i + 1) & 7]));
                }
            }
        }
    }
    double s = 0; int qs = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; qs += q[i]; fs += f[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + fs;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = qs;
}
template <int MODE, int K> void run(const char* name, double* d, int* di, int sms, double ghz) {
    const int blocks = sms * 8, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); mix<MODE, K><<<blocks, 256>>>(d, di, iters, 1.0000001, 1e-9, r); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    double dfma_warp_per_smsp = 32.0 * iters * blocks * 8.0 / (sms * 4.0);   // warp-DFMAs per sub-partition
    printf("%-12s K=%d  %8.3f ms  DFMA %6.2f TFLOP/s  %5.2f cycles/DFMA/SMSP (@%.3f GHz)\n", name, K, best,
           2 * 32.0 * iters * blocks * 256.0 / best / 1e9, best * 1e-3 * ghz * 1e9 / dfma_warp_per_smsp, ghz);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double ghz = p.clockRate * 1e-6;
    double* d; int* di; cudaMalloc(&d, p.multiProcessorCount * 8 * 256 * 8); cudaMalloc(&di, p.multiProcessorCount * 8 * 256 * 4);
    int n = p.multiProcessorCount;
    run<0, 0>("pure DFMA", d, di, n, ghz);
    run<1, 1>("+SEL", d, di, n, ghz); run<1, 2>("+SEL", d, di, n, ghz); run<1, 3>("+SEL", d, di, n, ghz);
    run<2, 1>("+LOP3", d, di, n, ghz); run<2, 2>("+LOP3", d, di, n, ghz); run<2, 3>("+LOP3", d, di, n, ghz);
    run<3, 1>("+IMAD", d, di, n, ghz); run<3, 2>("+IMAD", d, di, n, ghz);
    run<4, 1>("+FFMA", d, di, n, ghz); run<4, 2>("+FFMA", d, di, n, ghz);
    run<5, 1>("+MOV", d, di, n, ghz); run<5, 2>("+MOV", d, di, n, ghz);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
