// cartn_linearize.cu -- lane-parallel batched linearization of the N-link cart pendulum (cart_n_linearize.cuh).
//   cartn_linearize <in.bin> <out.bin>
// in.bin : int32 N, int32 P, then doubles mc[P], m[N][P], L[N][P], g[P], x[2(N+1)][P], u[P]
// out.bin: doubles A[(2N+2)^2][P], B[2N+2][P]            (compared with the reference's makeLinearizer by the pytest)
//   cartn_linearize bench          throughput: N = 1 lanes vs thread-per-point kernel, N = 6 lanes
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <sopot/physics/control/cart_n_linearize.cuh>

using namespace sopot;

template <size_t N>
static int run_file(FILE* in, int P, const char* out_path) {
    const size_t NS = 2 * (N + 1);
    std::vector<double> mc(P), m(N * P), L(N * P), g(P), x(NS * P), u(P);
    auto rd = [&](std::vector<double>& v) { return fread(v.data(), 8, v.size(), in) == v.size(); };
    if (!(rd(mc) && rd(m) && rd(L) && rd(g) && rd(x) && rd(u))) return 2;
    auto J = control::linearizeCartNPendulumBatch<N>(b200::Engine::shared(), mc, m, L, g, x, u);
    FILE* out = fopen(out_path, "wb");
    fwrite(J.A.data(), 8, J.A.size(), out); fwrite(J.B.data(), 8, J.B.size(), out);
    fclose(out);
    return 0;
}

template <size_t N>
static void bench(size_t P) {
    const size_t NS = 2 * (N + 1);
    std::vector<double> mc{1.0}, m(N, 0.1), L(N, 0.5), g{9.81}, x(NS * P), u(P, 0.3);
    for (size_t i = 0; i < x.size(); ++i) x[i] = 0.3 * std::sin(0.001 * double(i));
    auto& eng = b200::Engine::shared();
    cudaStream_t st = b200::generic::stream_of(eng);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        auto J = control::linearizeCartNPendulumBatch<N>(eng, mc, m, L, g, x, u);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = std::fmin(best, ms);
    }
    std::printf("lanes   N=%zu  P=%zu: %.3f ms incl. H2D/D2H (%zu B out per point)  %.3e points/s\n", N, P, best, NS * (NS + 1) * 8, P / (best * 1e-3));
}

int main(int argc, char** argv) {
    if (argc == 2 && !strcmp(argv[1], "bench")) {
        bench<1>(1 << 20);
        bench<2>(1 << 18);
        bench<6>(1 << 16);
        // the thread-per-point Dual<5> kernel on the same N = 1 problem, same host-buffer interface
        const size_t P = 1 << 20;
        std::vector<double> x(4 * P), u(P, 0.3);
        for (size_t i = 0; i < x.size(); ++i) x[i] = 0.3 * std::sin(0.001 * double(i));
        auto& eng = b200::Engine::shared();
        cudaStream_t st = b200::generic::stream_of(eng);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, st);
            auto J = control::linearizeCartPoleBatch(eng, 1.0, 0.1, 0.5, 9.81, x, u);
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = std::fmin(best, ms);
        }
        std::printf("thread  N=1  P=%zu: %.3f ms incl. H2D/D2H  %.3e points/s\n", P, best, P / (best * 1e-3));
        return 0;
    }
    if (argc != 3) { std::fprintf(stderr, "usage: cartn_linearize <in.bin> <out.bin> | bench\n"); return 64; }
    FILE* in = fopen(argv[1], "rb");
    int hdr[2];
    if (!in || fread(hdr, 4, 2, in) != 2) return 2;
    int rc = 3;
    if (hdr[0] == 1) rc = run_file<1>(in, hdr[1], argv[2]);
    if (hdr[0] == 2) rc = run_file<2>(in, hdr[1], argv[2]);
    if (hdr[0] == 6) rc = run_file<6>(in, hdr[1], argv[2]);
    fclose(in);
    if (!rc) std::puts("cartn_linearize: OK");
    return rc;
}
