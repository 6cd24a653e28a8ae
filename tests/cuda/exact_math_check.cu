// exact_math_check.cu -- are the branch-light IEEE division / sqrt sequences of csrc/ieee.cuh correctly rounded?
// Compares sb_div_rn / sb_div_by / sb_sqrt_rn with the compiler's IEEE `a / b` and `sqrt(a)` (nvcc -prec-div=true -prec-sqrt=true)
// bit for bit over random mantissas and exponents, exact zeros, near-1 quotients and perfect squares.
// Exit code 0 = identical on every sample.  Run on the GPU by tests/test_gpu_parity.py.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../sopot_b200/csrc/ieee.cuh"

__device__ uint64_t splitmix(uint64_t& s) { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }

__global__ void check(unsigned long long* bad, unsigned long long* first, int iters) {
    uint64_t s = 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long nbad = 0;
    for (int i = 0; i < iters; ++i) {
        const uint64_t ra = splitmix(s), rb = splitmix(s);
        // mantissas random; exponents in a window that depends on the sample class
        const int cls = i & 7;
        const int span = cls < 4 ? 40 : (cls < 6 ? 600 : 2046);       // +-20, +-300, everything (incl. subnormal/inf/nan)
        const uint64_t ea = (uint64_t)(1023 - span / 2 + (int)(ra >> 52) % (span + 1)) & 0x7ff;
        const uint64_t eb = (uint64_t)(1023 - span / 2 + (int)(rb >> 52) % (span + 1)) & 0x7ff;
        double a = __longlong_as_double((long long)((ra & 0x800fffffffffffffull) | (ea << 52)));
        double b = __longlong_as_double((long long)((rb & 0x800fffffffffffffull) | (eb << 52)));
        if (cls == 1) a = 0.0;                                         // exact zero numerator
        if (cls == 2) a = -0.0;
        if (cls == 3) a = b * (1.0 + 1e-15 * (double)(int)(ra & 15));  // quotient next to 1
        const double q0 = a / b, q1 = sb_div_rn(a, b);
        if (__double_as_longlong(q0) != __double_as_longlong(q1) && !(q0 != q0 && q1 != q1)) { if (!nbad) { first[0] = __double_as_longlong(a); first[1] = __double_as_longlong(b); } ++nbad; }
        // division by a launch-constant divisor with its correctly rounded reciprocal precomputed (sb_div_by)
        const double q2 = sb_div_by(a, b, 1.0 / b);
        if (__double_as_longlong(q0) != __double_as_longlong(q2) && !(q0 != q0 && q2 != q2)) { if (!nbad) { first[0] = __double_as_longlong(a); first[1] = __double_as_longlong(b) ^ 1ull << 63; } ++nbad; }
        const double q6 = sb_div_by(a, 6.0, 1.0 / 6.0);
        if (__double_as_longlong(a / 6.0) != __double_as_longlong(q6) && !(q6 != q6 && a != a)) { if (!nbad) { first[0] = __double_as_longlong(a); first[1] = 6; } ++nbad; }
        const double x = cls == 3 ? fabs(a) * fabs(a) : fabs(a);
        const double r0 = sqrt(x), r1 = sb_sqrt_rn(x);
        if (__double_as_longlong(r0) != __double_as_longlong(r1) && !(r0 != r0 && r1 != r1)) { if (!nbad) { first[0] = __double_as_longlong(x); first[1] = 0; } ++nbad; }
    }
    if (nbad) atomicAdd(bad, nbad);
}

int main() {
    unsigned long long *d_bad, *d_first, h_bad = 0, h_first[2] = {0, 0};
    cudaMalloc(&d_bad, 8); cudaMalloc(&d_first, 16);
    cudaMemset(d_bad, 0, 8); cudaMemset(d_first, 0, 16);
    check<<<148 * 4, 256>>>(d_bad, d_first, 4096);
    cudaMemcpy(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_first, d_first, 16, cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaGetLastError();
    printf("exact_math_check: %llu mismatches in %.3g samples (first: %016llx %016llx) %s\n", h_bad, 148.0 * 4 * 256 * 4096,
           h_first[0], h_first[1], cudaGetErrorString(e));
    return (h_bad == 0 && e == cudaSuccess) ? 0 : 1;
}
