// control::LQR<N,1> -- interface of physics/control/lqr.hpp:27-448 (solve / getGain / getRiccatiSolution / isSolved /
// computeControl) on top of the batched LQR kernels (csrc/lqr.cu, sopot_b200_lqr_gain).  The Riccati iteration of
// :260-396 runs on the device; one solve() is a P = 1 launch, control::solveLqrBatch runs a whole sweep in one.
// Kernels exist for N = 4, 6 and 14 (cart + 1, 2 and 6 links -- the sizes the reference's own tests and wrappers use).
// linearizeCartNPendulum / linearizeCartDoublePendulum (:462-681): the upright Jacobians come from the forward-mode
// kernels (Dual<double,2(N+1)+1> about x = 0, u = 0) instead of the reference's analytic M0^-1 algebra; the two agree to
// a few 1e-15 (tests/test_cpp_host_layer.py).
#pragma once
#include <array>
#include <stdexcept>
#include <utility>
#include <vector>
#include "../../b200/engine.hpp"
#include "cart_n_pendulum.hpp"
#if SOPOT_B200_DEVICE_COMPILER
#include "cart_n_linearize.cuh"
#endif

namespace sopot::control {

template <size_t N, size_t M>
class LQR {
    static_assert(M == 1 && (N == 4 || N == 6 || N == 14), "sopot-b200 ships the LQR<4,1>, LQR<6,1> and LQR<14,1> kernels");
public:
    using StateMatrix = std::array<std::array<double, N>, N>;
    using InputMatrix = std::array<std::array<double, M>, N>;
    using GainMatrix = std::array<std::array<double, N>, M>;
    using StateWeightMatrix = std::array<std::array<double, N>, N>;
    using InputWeightMatrix = std::array<std::array<double, M>, M>;

    LQR() : m_K{}, m_P{}, m_solved(false) {}

    bool solve(const StateMatrix& A, const InputMatrix& B, const StateWeightMatrix& Q, const InputWeightMatrix& R,
               double dt = 0.01, size_t max_iter = 1000, double tol = 1e-8) {
        double a[N * N], b[N], q[N * N], k[N], p[N * N];
        for (size_t i = 0; i < N; ++i) {
            for (size_t j = 0; j < N; ++j) { a[i * N + j] = A[i][j]; q[i * N + j] = Q[i][j]; }
            b[i] = B[i][0];
        }
        int32_t st = 0, iters = 0;
        b200::Engine& eng = b200::Engine::shared();
        b200::EnsembleOptions o; o.fp_mode = b200::FpMode::Strict;
        const auto opts = b200::make_opts(o, 0);
        eng.check(sopot_b200_lqr_gain(eng.ctx(), (int)N, a, b, 1, q, R[0][0], dt, max_iter, tol, &opts, k, p, &iters, &st));
        eng.sync();
        if (st == SOPOT_B200_ST_SINGULAR)       // the reference throws from solveForGain (:196-198)
            throw std::runtime_error("LQR solveForGain: (R + B'PB) is singular or ill-conditioned in scalar case");
        for (size_t i = 0; i < N; ++i) for (size_t j = 0; j < N; ++j) m_P[i][j] = p[i * N + j];
        m_solved = st == 0;
        if (m_solved) for (size_t j = 0; j < N; ++j) m_K[0][j] = k[j];
        m_iterations = static_cast<size_t>(iters);
        return m_solved;
    }
    const GainMatrix& getGain() const {
        if (!m_solved) throw std::runtime_error("LQR not solved yet");
        return m_K;
    }
    const StateMatrix& getRiccatiSolution() const { return m_P; }
    bool isSolved() const { return m_solved; }
    size_t iterations() const { return m_iterations; }      // additive: Riccati iterations the solve used
    std::array<double, M> computeControl(const std::array<double, N>& state, const std::array<double, N>& reference = {}) const {
        std::array<double, M> u{};
        for (size_t i = 0; i < M; ++i) {
            u[i] = 0.0;
            for (size_t j = 0; j < N; ++j) u[i] -= m_K[i][j] * (state[j] - reference[j]);
        }
        return u;
    }
private:
    GainMatrix m_K;
    StateMatrix m_P;
    bool m_solved;
    size_t m_iterations = 0;
};

// One launch for a whole sweep: A [16][P], B [4][P] as linearizeCartPoleBatch returns them.
struct LqrSweep {
    size_t P = 0;
    std::vector<double> K;          // [4][P], u = -K x
    std::vector<int32_t> iterations, status;
    double k(size_t j, size_t p) const { return K[j * P + p]; }
};
inline LqrSweep solveLqrBatch(b200::Engine& eng, const std::vector<double>& A, const std::vector<double>& B,
                              const std::array<std::array<double, 4>, 4>& Q, double R, double dt = 0.01, size_t max_iter = 1000,
                              double tol = 1e-8, b200::FpMode fp = b200::FpMode::Contract) {
    const size_t P = B.size() / 4;
    if (A.size() != 16 * P || B.size() != 4 * P) throw std::invalid_argument("A must hold 16*P and B 4*P values");
    double q[16];
    for (size_t i = 0; i < 4; ++i) for (size_t j = 0; j < 4; ++j) q[i * 4 + j] = Q[i][j];
    b200::EnsembleOptions o; o.fp_mode = fp;
    const auto opts = b200::make_opts(o, 0);
    LqrSweep r;
    r.P = P; r.K.resize(4 * P); r.iterations.resize(P); r.status.resize(P);
    eng.check(sopot_b200_lqr4_gain(eng.ctx(), A.data(), B.data(), P, q, R, dt, max_iter, tol, &opts, r.K.data(), nullptr,
                                   r.iterations.data(), r.status.data()));
    eng.sync();
    return r;
}

// Upright linearization of the cart + N-link pendulum, xdot = A x + B u (reference :578-681): (A, B) as a pair, so
// `auto [A, B] = linearizeCartNPendulum<N>(...)` works as in the reference's tests.  One link: the Dual<double,5> kernel
// behind the C ABI (any host compiler).  More links: the lane-parallel Dual kernel of cart_n_linearize.cuh, instantiated in
// the caller's translation unit, i.e. under nvcc -- like every other N-link path of this layer.
template <size_t NLinks>
std::pair<std::array<std::array<double, 2 * (NLinks + 1)>, 2 * (NLinks + 1)>, std::array<std::array<double, 1>, 2 * (NLinks + 1)>>
linearizeCartNPendulum(double mc, const std::array<double, NLinks>& masses, const std::array<double, NLinks>& lengths, double g) {
    constexpr size_t NS = 2 * (NLinks + 1);
    std::pair<std::array<std::array<double, NS>, NS>, std::array<std::array<double, 1>, NS>> out{};
    b200::Engine& eng = b200::Engine::shared();
    if constexpr (NLinks == 1) {
        const auto J = linearizeCartPoleBatch(eng, mc, masses[0], lengths[0], g, std::vector<double>(4, 0.0), std::vector<double>(1, 0.0),
                                              b200::FpMode::Strict);
        if (J.status[0] == SOPOT_B200_ST_SINGULAR) throw std::runtime_error("linearizeCartNPendulum: mass matrix is singular");
        for (size_t i = 0; i < NS; ++i) { for (size_t j = 0; j < NS; ++j) out.first[i][j] = J.a(i, j, 0); out.second[i][0] = J.b(i, 0); }
    } else {
#if SOPOT_B200_DEVICE_COMPILER
        const auto J = linearizeCartNPendulumBatch<NLinks>(eng, {mc}, std::vector<double>(masses.begin(), masses.end()),
                                                           std::vector<double>(lengths.begin(), lengths.end()), {g},
                                                           std::vector<double>(NS, 0.0), std::vector<double>(1, 0.0));
        for (size_t i = 0; i < NS; ++i) {
            for (size_t j = 0; j < NS; ++j) {
                if (J.a(i, j, 0) != J.a(i, j, 0)) throw std::runtime_error("linearizeCartNPendulum: mass matrix is singular");
                out.first[i][j] = J.a(i, j, 0);
            }
            out.second[i][0] = J.b(i, 0);
        }
#else
        static_assert(NLinks == 1, "linearizeCartNPendulum<N> for N > 1 instantiates a kernel in the caller's translation unit: compile with nvcc");
#endif
    }
    return out;
}
// :462-560: the generic plant with two links IS the analytic double pendulum (tests/inverted_pendulum_test.cpp:67-93)
inline std::pair<std::array<std::array<double, 6>, 6>, std::array<std::array<double, 1>, 6>>
linearizeCartDoublePendulum(double mc, double m1, double m2, double L1, double L2, double g) {
#if SOPOT_B200_DEVICE_COMPILER
    return linearizeCartNPendulum<2>(mc, {m1, m2}, {L1, L2}, g);
#else
    (void)mc; (void)m1; (void)m2; (void)L1; (void)L2; (void)g;
    throw std::runtime_error("linearizeCartDoublePendulum instantiates a kernel in the caller's translation unit: compile with nvcc");
#endif
}

}  // namespace sopot::control
