// rcp_accuracy.cu -- relative error of MUFU.RCP64H (rcp.approx.ftz.f64) and of the Newton refinements built on it.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(double* out, int n) {
    double w0 = 0, w1 = 0, w2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // x sweeps the mantissa range [1,2) densely plus a spread of exponents
        const double x = ldexp(1.0 + (double)i / n, (i % 41) - 20);
        double y;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        const double exact = 1.0 / x;
        w0 = fmax(w0, fabs(y - exact) / exact);
        double e = fma(-x, y, 1.0);
        const double y1 = fma(y, fma(e, e, e), y);          // 3 FMA: cubic
        w1 = fmax(w1, fabs(y1 - exact) / exact);
        double y2 = fma(y, e, y); e = fma(-x, y2, 1.0); y2 = fma(y2, e, y2);   // 4 FMA: two quadratic steps
        w2 = fmax(w2, fabs(y2 - exact) / exact);
    }
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 0] = w0;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 1] = w1;
    out[(blockIdx.x * blockDim.x + threadIdx.x) * 3 + 2] = w2;
}
int main() {
    const int threads = 148 * 256;
    double* d; cudaMalloc(&d, threads * 3 * 8);
    k<<<148, 256>>>(d, 1 << 26);
    double* h = new double[threads * 3];
    cudaMemcpy(h, d, threads * 3 * 8, cudaMemcpyDeviceToHost);
    double w[3] = {0, 0, 0};
    for (int i = 0; i < threads; ++i) for (int j = 0; j < 3; ++j) w[j] = fmax(w[j], h[i * 3 + j]);
    printf("rcp.approx.ftz.f64 max rel err %.3e (2^%.1f); +3 FMA cubic %.3e; +4 FMA quadratic x2 %.3e; %s\n", w[0], log2(w[0]), w[1], w[2],
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
