// plugin_bench.cu -- throughput of the GENERIC (compiler-generated) kernel against the hand-tuned one on the same
// workload: double-pendulum ensemble, 1 Mi instances x 1000 steps.  The generic kernel evaluates the component's
// derivatives() as written (libdevice sin/cos, IEEE divisions); the tuned kernel is rk4_ensemble_kernel<DpendSys>.
#include <cstdio>
#include <vector>
#include <cmath>
#include <sopot/core/solver.hpp>
#include <sopot/physics/pendulum/double_pendulum.hpp>
using namespace sopot;
struct Factory {
    const double* th1; const double* th2;
    SOPOT_HD auto operator()(size_t i) const { return makeTypedODESystem<double>(pendulum::DoublePendulum<double>(1.0, 1.0, 1.0, 1.0, 9.81, th1[i], th2[i], 0.0, 0.0)); }
};
int main() {
    const size_t n = 1 << 20, steps = 1000;
    std::vector<double> a(n), b(n);
    for (size_t i = 0; i < n; ++i) { a[i] = 3.0 * std::sin(0.001 * i); b[i] = 3.0 * std::cos(0.0017 * i); }
    double *da, *db;
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8);
    cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    b200::Engine& eng = b200::Engine::shared();
    b200::EnsembleOptions opt; opt.want_status = false;
    auto solver = createRK4Solver();
    for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaStream_t st = b200::generic::stream_of(eng);
        cudaEventRecord(e0, st);
        auto r = solver.solveEnsemble(b200::ensembleOf(n, Factory{da, db}), 0.0, 1.0, 1e-3, opt);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) std::printf("generic kernel : %.2f ms (incl. D2H of the final state)  %.3e instance-steps/s\n", ms, n * steps / (ms * 1e-3));
        pendulum::DoublePendulumEnsemble ens{n, 1.0, 1.0, 1.0, 1.0, 9.81, a, b, 0.0, 0.0};
        cudaEventRecord(e0, st);
        auto r2 = solver.solveEnsemble(ens, 0.0, 1.0, 1e-3, opt);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep) {
            std::printf("tuned kernel   : %.2f ms (incl. H2D/D2H)                  %.3e instance-steps/s\n", ms, n * steps / (ms * 1e-3));
            double err = 0, scale = 0;
            for (size_t i = 0; i < 4 * n; ++i) { err = std::fmax(err, std::fabs(r.state[i] - r2.state[i])); scale = std::fmax(scale, std::fabs(r2.state[i])); }
            std::printf("max |generic - tuned| / max|state| after 1 s = %.2e\n", err / scale);
        }
    }
    return 0;
}
