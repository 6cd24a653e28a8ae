// control::StateFeedbackController<N,M> (physics/control/state_feedback_controller.hpp:21-118) and
// CartNPendulumController<1>::createWithLQR (:216-266).  The controller itself is a small host object (u = -K (x -
// x_ref), optional clamp) exactly as in the reference; what runs on the device is (a) createWithLQR's linearization
// and Riccati iteration and (b) closed-loop ensembles: CartPoleClosedLoopEnsemble integrates n plants under their
// own gains with the force sampled once per RK4 step (sopot_b200_cartpole_feedback_rk4).
#pragma once
#include <algorithm>
#include <array>
#include "cart_n_pendulum.hpp"
#include "lqr.hpp"
#include "../../core/linearization.hpp"

namespace sopot::control {

template <size_t N, size_t M>
class StateFeedbackController {
public:
    using StateVector = std::array<double, N>;
    using ControlVector = std::array<double, M>;
    using GainMatrix = std::array<std::array<double, N>, M>;
    explicit StateFeedbackController(const GainMatrix& K) : m_K(K) {}
    void setGain(const GainMatrix& K) { m_K = K; }
    void setReference(const StateVector& r) { m_reference = r; }
    const StateVector& getReference() const { return m_reference; }
    void setSaturationLimits(const ControlVector& u_min, const ControlVector& u_max) { m_u_min = u_min; m_u_max = u_max; m_saturate = true; }
    void disableSaturation() { m_saturate = false; }
    bool saturates() const { return m_saturate; }
    const ControlVector& lowerLimit() const { return m_u_min; }
    const ControlVector& upperLimit() const { return m_u_max; }
    ControlVector compute(const StateVector& state) const {                 // :88-104
        ControlVector u{};
        for (size_t i = 0; i < M; ++i) {
            u[i] = 0.0;
            for (size_t j = 0; j < N; ++j) u[i] -= m_K[i][j] * (state[j] - m_reference[j]);
            if (m_saturate) u[i] = std::clamp(u[i], m_u_min[i], m_u_max[i]);
        }
        return u;
    }
    const GainMatrix& getGain() const { return m_K; }
private:
    GainMatrix m_K;
    StateVector m_reference{};
    ControlVector m_u_min{}, m_u_max{};
    bool m_saturate = false;
};

// :124-212: the six-state controller of the cart + double pendulum
class InvertedDoublePendulumController : public StateFeedbackController<6, 1> {
public:
    using Base = StateFeedbackController<6, 1>;
    explicit InvertedDoublePendulumController(const GainMatrix& K) : Base(K) {}
    static InvertedDoublePendulumController createWithLQR(double mc, double m1, double m2, double L1, double L2, double g,
                                                          const std::array<double, 6>& q_diag = {10.0, 100.0, 100.0, 1.0, 10.0, 10.0},
                                                          double r = 0.01) {
        auto [A, B] = linearizeCartDoublePendulum(mc, m1, m2, L1, L2, g);
        LQR<6, 1>::StateWeightMatrix Q{};
        for (size_t i = 0; i < 6; ++i) Q[i][i] = q_diag[i];
        LQR<6, 1>::InputWeightMatrix R{};
        R[0][0] = r;
        LQR<6, 1> lqr;
        lqr.solve(A, B, Q, R);              // the reference uses the gain whether or not the iteration met its tolerance (:196-204)
        return InvertedDoublePendulumController(lqr.getGain());
    }
    double computeForce(double x, double theta1, double theta2, double xdot, double omega1, double omega2) const {
        return compute(StateVector{x, theta1, theta2, xdot, omega1, omega2})[0];
    }
};

template <size_t NLinks>
class CartNPendulumController : public StateFeedbackController<2 * (NLinks + 1), 1> {
public:
    static constexpr size_t state_dim = 2 * (NLinks + 1);
    using Base = StateFeedbackController<state_dim, 1>;
    using typename Base::GainMatrix;
    explicit CartNPendulumController(const GainMatrix& K) : Base(K) {}
    // :234-260.  The reference linearizes analytically (linearizeCartNPendulum); here the upright Jacobians come from
    // the forward-mode kernels (equal to the analytic ones to ~4e-15), the Riccati iteration
    // from the LQR kernel.  One link works under any host compiler, N > 1 under nvcc (see linearizeCartNPendulum).
    static CartNPendulumController createWithLQR(double mc, const std::array<double, NLinks>& masses,
                                                 const std::array<double, NLinks>& lengths, double g,
                                                 const std::array<double, state_dim>& q_diag, double r, double riccati_dt = 0.002) {
        auto [A, B] = linearizeCartNPendulum<NLinks>(mc, masses, lengths, g);
        typename LQR<state_dim, 1>::StateWeightMatrix Q{};
        for (size_t i = 0; i < state_dim; ++i) Q[i][i] = q_diag[i];
        typename LQR<state_dim, 1>::InputWeightMatrix R{};
        R[0][0] = r;
        LQR<state_dim, 1> lqr;
        lqr.solve(A, B, Q, R, riccati_dt);
        return CartNPendulumController(lqr.getGain());
    }
    double computeForce(const std::array<double, state_dim>& state) const { return this->compute(state)[0]; }
};

// closed-loop ensemble for RK4Solver::solveEnsemble: n cart-poles, each under its own gain row K[:, i]
struct CartPoleClosedLoopResult : b200::EnsembleResult {
    std::vector<double> stats;          // [2][n]: max |theta|, max |u|
    double maxAngle(size_t i) const { return stats[i]; }
    double maxForce(size_t i) const { return stats[n + i]; }
};
struct CartPoleClosedLoopEnsemble {
    size_t n;
    b200::Param cart_mass, link_mass, link_length, gravity, x, theta, x_dot, omega;
    std::vector<double> K;              // [4][n]
    std::vector<double> x_ref;          // [4][n] or empty (= 0)
    bool saturate = false;
    double u_min = 0.0, u_max = 0.0;
    CartPoleClosedLoopResult integrate(b200::Engine& eng, size_t n_steps, double t0, double dt, const b200::EnsembleOptions& o) const {
        if (K.size() != 4 * n || (!x_ref.empty() && x_ref.size() != 4 * n)) throw std::invalid_argument("K / x_ref must hold 4*n values");
        const b200::Param* f[8] = {&cart_mass, &link_mass, &link_length, &gravity, &x, &theta, &x_dot, &omega};
        uint32_t bc = 0;
        for (int j = 0; j < 8; ++j) bc |= (f[j]->broadcast(n) ? 1u : 0u) << j;
        sopot_b200_cartpole_params p{f[0]->data(), f[1]->data(), f[2]->data(), f[3]->data(), f[4]->data(), f[5]->data(),
                                     f[6]->data(), f[7]->data(), nullptr};
        sopot_b200_feedback fb{K.data(), x_ref.empty() ? nullptr : x_ref.data(), u_min, u_max, saturate ? 1 : 0};
        const auto opts = b200::make_opts(o, bc);
        CartPoleClosedLoopResult r;
        r.n = n; r.dim = 4; r.n_steps = n_steps; r.store_every = o.store_every;
        r.state.resize(4 * n);
        r.trajectory.resize(r.snapshots() * 4 * n);
        r.stats.resize(2 * n);
        if (o.want_status) r.status.resize(n);
        eng.check(sopot_b200_cartpole_feedback_rk4(eng.ctx(), &p, &fb, n, t0, dt, n_steps, &opts, r.state.data(),
                                                   r.trajectory.empty() ? nullptr : r.trajectory.data(), r.stats.data(),
                                                   o.want_status ? r.status.data() : nullptr));
        eng.sync();
        return r;
    }
};

}  // namespace sopot::control
