// LQR and closed loop in the shape of the reference's tests/inverted_pendulum_test.cpp (LQR converges :225-227,
// closed-loop stabilisation :241-292: max angle < 0.2 rad, final angle < 0.01 rad) for the shipped one-link plant,
// plus the sweep entry points.  Needs a B200.
#include <cstdio>
#include <sopot/core/solver.hpp>
#include <sopot/physics/control/state_feedback_controller.hpp>
#include "check.hpp"

using namespace sopot;
using namespace sopot::control;

int main() {
    const double mc = 1.0, m = 0.1, L = 0.5, g = 9.81;
    // LQR on the analytic upright linearization of the cart-pole
    LQR<4, 1>::StateMatrix A{}, Q{};
    LQR<4, 1>::InputMatrix B{};
    A[0][2] = 1.0; A[1][3] = 1.0; A[2][1] = -m * g / mc; A[3][1] = (mc + m) * g / (mc * L);
    B[2][0] = 1.0 / mc; B[3][0] = -1.0 / (mc * L);
    const double qd[4] = {10.0, 100.0, 1.0, 10.0};
    for (int i = 0; i < 4; ++i) Q[i][i] = qd[i];
    LQR<4, 1>::InputWeightMatrix R{};
    R[0][0] = 0.01;
    LQR<4, 1> lqr;
    ASSERT_THROWS(lqr.getGain(), std::runtime_error);             // "LQR not solved yet"
    ASSERT_TRUE(lqr.solve(A, B, Q, R));
    ASSERT_TRUE(lqr.isSolved() && lqr.iterations() > 10 && lqr.iterations() <= 1000);
    const auto K = lqr.getGain();
    // a stabilising gain for this sign convention pushes the cart under the falling pole: K_theta < 0, |K_theta| dominant
    ASSERT_TRUE(K[0][1] < 0.0 && std::abs(K[0][1]) > std::abs(K[0][0]));
    const auto Pric = lqr.getRiccatiSolution();
    for (int i = 0; i < 4; ++i) { ASSERT_TRUE(Pric[i][i] > 0.0); for (int j = 0; j < 4; ++j) ASSERT_TRUE(Pric[i][j] == Pric[j][i]); }
    const auto u = lqr.computeControl({0.0, 0.05, 0.0, 0.0});
    ASSERT_NEAR(u[0], -K[0][1] * 0.05, 1e-15);
    R[0][0] = 0.0;
    LQR<4, 1>::InputMatrix B0{};
    LQR<4, 1> singular;
    ASSERT_THROWS(singular.solve(A, B0, Q, R), std::runtime_error);      // singular scalar (R + B'PB)

    // controller factory: linearization + Riccati iteration on the device
    auto controller = CartNPendulumController<1>::createWithLQR(mc, {m}, {L}, g, {10.0, 100.0, 1.0, 10.0}, 0.01, 0.005);
    controller.setSaturationLimits({-50.0}, {50.0});
    ASSERT_TRUE(controller.getGain()[0][1] < 0.0);
    ASSERT_NEAR(controller.computeForce({0.0, 1e3, 0.0, 0.0}), 50.0, 0.0);         // clamped

    // closed loop, 512 plants with jittered link mass under the same gain; force sampled once per step
    const size_t n = 512;
    std::vector<double> link_mass(n), theta0(n), Kall(4 * n);
    for (size_t i = 0; i < n; ++i) {
        link_mass[i] = m * (1.0 + 0.2 * std::sin(double(i)));
        theta0[i] = 0.05 * std::cos(0.3 * double(i));
        for (size_t j = 0; j < 4; ++j) Kall[j * n + i] = controller.getGain()[0][j];
    }
    CartPoleClosedLoopEnsemble ens{n, mc, link_mass, L, g, 0.0, theta0, 0.0, 0.0, Kall, {}, true, -50.0, 50.0};
    b200::EnsembleOptions opt;
    opt.fp_mode = b200::FpMode::Strict;
    auto r = createRK4Solver().solveEnsemble(ens, 0.0, 5.0, 0.001, opt);
    for (size_t i = 0; i < n; ++i) {
        ASSERT_TRUE(r.status[i] == 0);
        ASSERT_TRUE(r.maxAngle(i) < 0.2);                       // never falls
        ASSERT_TRUE(std::abs(r.final(1, i)) < 0.01);            // converges back toward upright
        ASSERT_TRUE(r.maxForce(i) <= 50.0);
    }
    // the ensemble's instance 0 equals a host-driven loop: controller.compute -> setControlForce -> one solver step
    auto plant = makeTypedODESystem<double>(CartNPendulum<1, double>(mc, {link_mass[0]}, {L}, g));
    std::vector<double> y = {0.0, theta0[0], 0.0, 0.0};
    auto derivs = [&plant](double t, StateView s) { return plant.computeDerivatives(t, std::vector<double>(s.begin(), s.end())); };
    double t = 0.0;
    for (int step = 0; step < 25; ++step) {
        plant.getComponent<0>().setControlForce(controller.computeForce({y[0], y[1], y[2], y[3]}));
        y = createRK4Solver().solve(derivs, 4, t, t + 0.001, 0.001, y).states.back();
        t += 0.001;
    }
    auto r25 = createRK4Solver().solveEnsemble(ens, 0.0, 0.025, 0.001, opt);
    for (int j = 0; j < 4; ++j) ASSERT_TRUE(r25.final(j, 0) == y[j]);

    // sweep: Jacobians of 300 plants at upright -> 300 gains in one launch each
    const size_t P = 300;
    std::vector<double> masses(P), x0(4 * P, 0.0), u0(P, 0.0);
    for (size_t p = 0; p < P; ++p) masses[p] = 0.05 + 0.001 * double(p);
    auto J = linearizeCartPoleBatch(b200::Engine::shared(), mc, masses, L, g, x0, u0);
    auto sweep = solveLqrBatch(b200::Engine::shared(), J.A, J.B, Q, 0.01);
    for (size_t p = 0; p < P; ++p) { ASSERT_TRUE(sweep.status[p] == 0); ASSERT_TRUE(sweep.k(1, p) < 0.0); }
    ASSERT_TRUE(std::abs(sweep.k(1, 50) - K[0][1]) < 1e-6 * std::abs(K[0][1]));     // m = 0.1 is plant 50
    std::puts("lqr_test: OK");
    return 0;
}
