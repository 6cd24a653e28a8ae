// Mirror of the reference's tests/double_pendulum_test.cpp (state functions :136-162, energy drift < 1 %
// over 20 s :208) plus the reference's known answer for (pi/4, pi/6) after 10 s through RK4Solver::solve
// (SURVEY.md section 6).  Needs a B200.
#include <cstdio>
#include <numbers>
#include <sopot/core/solver.hpp>
#include <sopot/physics/pendulum/double_pendulum.hpp>
#include "check.hpp"

using namespace sopot;
using namespace sopot::pendulum;

int main() {
    const double pi = std::numbers::pi;
    auto system = makeTypedODESystem<double>(DoublePendulum<double>(1.0, 1.0, 1.0, 1.0, 9.81, pi / 4, pi / 6, 0.0, 0.0));
    static_assert(decltype(system)::template hasFunction<mass1::Angle>());
    static_assert(decltype(system)::template hasFunction<mass2::CartesianPosition>());
    static_assert(decltype(system)::template hasFunction<pendulum::system::TotalEnergy>());
    auto s0 = system.getInitialState();
    ASSERT_NEAR(system.computeStateFunction<mass1::Angle>(s0), pi / 4, 1e-15);
    ASSERT_NEAR(system.computeStateFunction<mass2::Angle>(s0), pi / 6, 1e-15);
    auto p1 = system.computeStateFunction<mass1::CartesianPosition>(s0);
    ASSERT_NEAR(p1[0], std::sin(pi / 4), 1e-12);
    ASSERT_NEAR(p1[1], -std::cos(pi / 4), 1e-12);
    auto p2 = system.computeStateFunction<mass2::CartesianPosition>(s0);
    ASSERT_NEAR(p2[0], std::sin(pi / 4) + std::sin(pi / 6), 1e-12);
    ASSERT_THROWS(DoublePendulum<double>(0.0, 1.0, 1.0, 1.0), std::invalid_argument);
    ASSERT_THROWS(DoublePendulum<double>(1.0, 1.0, -1.0, 1.0), std::invalid_argument);

    // device derivative at rest: omega' from the Cramer solve, theta' = omega = 0
    auto d = system.computeDerivatives(0.0, s0);
    ASSERT_TRUE(d[0] == 0.0 && d[1] == 0.0 && d[2] < 0.0);

    // energy drift over 20 s with the reference test's own small-angle case (0.3, 0.3), dt = 5e-4, bound 1 %
    // (tests/double_pendulum_test.cpp:170-213).  The bound is loose for a reason: the reference's b1/b2
    // (double_pendulum.hpp:168-169) carry the centrifugal terms with the opposite sign of the Lagrangian
    // equations, so energy is only approximately conserved; the kernel reproduces the reference, not the textbook.
    auto solver = createRK4Solver();
    auto small = makeTypedODESystem<double>(DoublePendulum<double>(1.0, 1.0, 1.0, 1.0, 9.81, 0.3, 0.3, 0.0, 0.0));
    auto derivs = [&small](double t, StateView s) { return small.computeDerivatives(t, std::vector<double>(s.begin(), s.end())); };
    auto r = solver.solve(derivs, 4, 0.0, 20.0, 5e-4, small.getInitialState());
    ASSERT_TRUE(r.size() == 40001);
    const double e0 = small.computeStateFunction<pendulum::system::TotalEnergy>(r.states.front());
    double worst = 0.0;
    for (const auto& y : r.states)
        worst = std::max(worst, std::abs(small.computeStateFunction<pendulum::system::TotalEnergy>(y) - e0) / std::abs(e0));
    ASSERT_TRUE(worst < 0.01);

    // known answer after 10 s (reference CPU run): <= 1e-9 relative on the final state
    auto r10 = solver.solve(system, 0.0, 10.0, 1e-3);
    const double ref[4] = {-2.0583664572549361, -9.8402182120618615, -1.5202026871720531, 0.27300742483572987};
    for (int j = 0; j < 4; ++j) ASSERT_NEAR(r10.states.back()[j], ref[j], 1e-9 * std::abs(ref[j]) + 1e-12);

    // ensemble, contract mode (the fast path) agrees with strict to the stated tolerance over 1 s
    const size_t n = 512;
    std::vector<double> th1(n), th2(n);
    for (size_t i = 0; i < n; ++i) { th1[i] = 3.0 * std::sin(1.0 + double(i)); th2[i] = 3.0 * std::cos(2.0 * double(i)); }
    DoublePendulumEnsemble ens{n, 1.0, 1.0, 1.0, 1.0, 9.81, th1, th2, 0.0, 0.0};
    b200::EnsembleOptions fast, strict;
    strict.fp_mode = b200::FpMode::Strict;
    auto a = solver.solveEnsemble(ens, 0.0, 1.0, 1e-3, fast);
    auto b = solver.solveEnsemble(ens, 0.0, 1.0, 1e-3, strict);
    double scale = 0.0, err = 0.0;
    for (size_t i = 0; i < 4 * n; ++i) { scale = std::max(scale, std::abs(b.state[i])); err = std::max(err, std::abs(a.state[i] - b.state[i])); }
    ASSERT_TRUE(err <= 1e-9 * scale);
    for (auto st : a.status) ASSERT_TRUE(st == 0);
    // Test 8 of the reference (tests/double_pendulum_test.cpp:383-412): autodiff Jacobian of DoublePendulum<Dual<double,4>>
    {
        using Dual4 = Dual<double, 4>;
        auto dsys = makeTypedODESystem<Dual4>(DoublePendulum<Dual4>(1.0, 1.0, 1.0, 1.0, 9.81, 0.5, 0.3, 0.1, -0.1));
        auto dstate = dsys.getInitialState();
        std::array<double, 4> values;
        for (size_t i = 0; i < 4; ++i) values[i] = value_of(dstate[i]);
        const auto jac = computeJacobian(dsys, 0.0, values);
        ASSERT_NEAR(jac[0][2], 1.0, 1e-10);
        ASSERT_NEAR(jac[1][3], 1.0, 1e-10);
        ASSERT_NEAR(jac[0][0], 0.0, 1e-10);
        ASSERT_NEAR(jac[0][1], 0.0, 1e-10);
        // the reference's own numbers for this point (oracle/_ref, tests/golden/dpend_jacobian.npz column 0)
        const double row2[4] = {-14.735500613870823, 7.007393204894467, 0.03746318109240295, -0.03822513892364516};
        const double row3[4] = {13.177970444475825, -14.975761621293028, -0.07645027784729032, 0.03746318109240295};
        for (int j = 0; j < 4; ++j) { ASSERT_NEAR(jac[2][j], row2[j], 1e-12 * 15.0); ASSERT_NEAR(jac[3][j], row3[j], 1e-12 * 15.0); }
    }
    std::puts("double_pendulum_test: OK");
    return 0;
}
