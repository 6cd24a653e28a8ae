// fp64_operands.cu -- DFMA / DMUL / DADD issue cost on sm_100a as a function of where the operands live.
// 8 independent chains per thread, 8 warps per SM sub-partition; cycles per warp instruction per sub-partition.
//   MODE 0  x = fma(x, U, U)      one vector register, two uniform (kernel parameters)
//   MODE 1  x = fma(x, y, U)      two distinct vector registers + uniform
//   MODE 2  x = fma(x, y, z)      three distinct vector registers
//   MODE 3  x = fma(y, z, x)      three distinct, accumulator last
//   MODE 4  x = fma(x, x, y)      repeated source
//   MODE 5  x = x * y             DMUL two vector
//   MODE 6  x = x + y             DADD two vector
//   MODE 7  x = fma(x, y, z) with y,z = neighbours in the chain array (more register banks in flight)
//   MODE 8  x = fma(x, Y, z)      three distinct, but Y is ONE register shared by consecutive instructions (operand reuse cache, slot b)
//   MODE 9  x = fma(x, y, Z)      the same with the shared register in slot c
//   MODE 10 x = fma(Y, x, z)      shared register in slot a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = 1.0 + 1e-9 * (threadIdx.x + i); y[i] = 1.0 - 1e-9 * (threadIdx.x + 2 * i); z[i] = 1e-12 * (i + threadIdx.x); }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
                if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(b));
                if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[i]) : "d"(y[i]), "d"(z[i]));
                if (MODE == 4) asm volatile("fma.rn.f64 %0, %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
                if (MODE == 5) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
                if (MODE == 6) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(y[i]));
                if (MODE == 7) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(x[(i + 3) & 7]), "d"(x[(i + 5) & 7]));
                if (MODE == 8) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[0]), "d"(z[i]));
                if (MODE == 9) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(y[i]), "d"(z[0]));
                if (MODE == 10) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(x[i]) : "d"(y[0]), "d"(z[i]));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double* d, int sms, double ghz) {
    const int blocks = sms * 8, iters = 1024;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    const double warp_inst_per_smsp = (double)iters * 64.0 * blocks * 8.0 / (sms * 4.0);
    printf("%-28s %8.3f ms  %6.2f T inst/s  %5.2f cycles/warp-inst/SMSP (@%.3f GHz)\n", name, best,
           32.0 * iters * 64.0 * blocks * 8.0 / best / 1e9, best * 1e-3 * ghz * 1e9 / warp_inst_per_smsp, ghz);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate * 1e-6;
    double* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 8);
    const int n = p.multiProcessorCount;
    run<0>("fma(x,U,U)", d, n, ghz); run<1>("fma(x,y,U)", d, n, ghz); run<2>("fma(x,y,z)", d, n, ghz);
    run<3>("fma(y,z,x)", d, n, ghz); run<4>("fma(x,x,y)", d, n, ghz); run<5>("mul(x,y)", d, n, ghz);
    run<6>("add(x,y)", d, n, ghz); run<7>("fma(x,x3,x5)", d, n, ghz);
    run<8>("fma(x,Yshared,z)", d, n, ghz); run<9>("fma(x,y,Zshared)", d, n, ghz); run<10>("fma(Yshared,x,z)", d, n, ghz);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
