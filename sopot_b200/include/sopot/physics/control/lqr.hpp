// control::LQR<4,1> -- interface of physics/control/lqr.hpp:27-448 (solve / getGain / getRiccatiSolution / isSolved /
// computeControl) on top of the batched LQR kernel (csrc/lqr.cu, sopot_b200_lqr4_gain).  The Riccati iteration of
// :260-396 runs on the device; one solve() is a P = 1 launch, control::solveLqrBatch runs a whole sweep in one.
// Only the <4,1> instantiation (cart + one link, the BASELINE linearization config) has a kernel.
#pragma once
#include <array>
#include <stdexcept>
#include <vector>
#include "../../b200/engine.hpp"

namespace sopot::control {

template <size_t N, size_t M>
class LQR {
    static_assert(N == 4 && M == 1, "sopot-b200 ships the LQR<4,1> kernel only");
public:
    using StateMatrix = std::array<std::array<double, N>, N>;
    using InputMatrix = std::array<std::array<double, M>, N>;
    using GainMatrix = std::array<std::array<double, N>, M>;
    using StateWeightMatrix = std::array<std::array<double, N>, N>;
    using InputWeightMatrix = std::array<std::array<double, M>, M>;

    LQR() : m_K{}, m_P{}, m_solved(false) {}

    bool solve(const StateMatrix& A, const InputMatrix& B, const StateWeightMatrix& Q, const InputWeightMatrix& R,
               double dt = 0.01, size_t max_iter = 1000, double tol = 1e-8) {
        double a[16], b[4], q[16], k[4], p[16];
        for (size_t i = 0; i < N; ++i) {
            for (size_t j = 0; j < N; ++j) { a[i * N + j] = A[i][j]; q[i * N + j] = Q[i][j]; }
            b[i] = B[i][0];
        }
        int32_t st = 0, iters = 0;
        b200::Engine& eng = b200::Engine::shared();
        b200::EnsembleOptions o; o.fp_mode = b200::FpMode::Strict;
        const auto opts = b200::make_opts(o, 0);
        eng.check(sopot_b200_lqr4_gain(eng.ctx(), a, b, 1, q, R[0][0], dt, max_iter, tol, &opts, k, p, &iters, &st));
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

}  // namespace sopot::control
