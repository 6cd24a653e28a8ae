// Mirror of the reference's linearization tests (tests/autodiff_test.cpp:516-523 makeLinearizer<2,1>,
// tests/inverted_pendulum_test.cpp:186-203 analytic A/B structure) for the shipped plant:
// makeLinearizer<4,1> around CartNPendulum<1, Dual<double,5>> with setControlForce(u[0]) in the dynamics
// lambda, the reference's own pattern.  Needs a B200.
#include <cstdio>
#include <sopot/core/linearization.hpp>
#include <sopot/core/solver.hpp>
#include <sopot/physics/control/cart_n_pendulum.hpp>
#include "check.hpp"

using namespace sopot;

int main() {
    using Lin = SystemLinearizer<4, 1>;
    using D = Lin::DualT;
    static_assert(std::is_same_v<D, Dual<double, 5>>);
    const double mc = 1.0, m = 0.1, L = 0.5, g = 9.81;
    auto sys = makeTypedODESystem<D>(control::CartNPendulum<1, D>(mc, {m}, {L}, g));
    auto lin = makeLinearizer<4, 1>([&sys](D t, const Lin::DualState& xs, const Lin::DualInput& us) {
        sys.template getComponent<0>().setControlForce(us[0]);
        std::vector<D> sv(xs.begin(), xs.end());
        auto d = sys.computeDerivatives(t, sv);
        Lin::DualDerivative out;
        for (size_t i = 0; i < 4; ++i) out[i] = d[i];
        return out;
    });
    // upright equilibrium: analytic Jacobians of the cart-pole
    auto up = lin.linearize(0.0, {0.0, 0.0, 0.0, 0.0}, {0.0});
    const double Aref[4][4] = {{0, 0, 1, 0}, {0, 0, 0, 1}, {0, -m * g / mc, 0, 0}, {0, (mc + m) * g / (mc * L), 0, 0}};
    const double Bref[4] = {0, 0, 1.0 / mc, -1.0 / (mc * L)};
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) ASSERT_NEAR(up.A[i][j], Aref[i][j], 1e-12);
        ASSERT_NEAR(up.B[i][0], Bref[i], 1e-12);
    }
    ASSERT_TRUE(up.t0 == 0.0 && up.x0[1] == 0.0 && up.u0[0] == 0.0);

    // away from equilibrium: central finite differences of the double-precision plant (device derivative)
    const Vector<4> x{0.3, 0.7, -0.4, 1.1};
    const Vector<1> u{2.5};
    auto off = lin.linearize(0.0, x, u);
    auto f = [&](Vector<4> xx, double uu) {
        auto s = makeTypedODESystem<double>(control::CartNPendulum<1, double>(mc, {m}, {L}, g));
        s.getComponent<0>().setControlForce(uu);
        return s.computeDerivatives(0.0, std::vector<double>(xx.begin(), xx.end()));
    };
    const double h = 1e-6;
    for (int j = 0; j < 4; ++j) {
        auto xp = x, xm = x;
        xp[j] += h; xm[j] -= h;
        auto fp = f(xp, u[0]), fm = f(xm, u[0]);
        for (int i = 0; i < 4; ++i) ASSERT_NEAR(off.A[i][j], (fp[i] - fm[i]) / (2 * h), 1e-7);
    }
    auto fp = f(x, u[0] + h), fm = f(x, u[0] - h);
    for (int i = 0; i < 4; ++i) ASSERT_NEAR(off.B[i][0], (fp[i] - fm[i]) / (2 * h), 1e-7);

    // batched: P points in one launch equal the single-point linearizer bit for bit (strict mode)
    const size_t P = 257;
    std::vector<double> x0(4 * P), u0(P);
    for (size_t p = 0; p < P; ++p) {
        x0[p] = std::sin(double(p)); x0[P + p] = 3.0 * std::cos(0.37 * double(p));
        x0[2 * P + p] = std::cos(double(p)); x0[3 * P + p] = 2.0 * std::sin(0.11 * double(p));
        u0[p] = 5.0 * std::sin(0.5 * double(p));
    }
    auto J = control::linearizeCartPoleBatch(b200::Engine::shared(), mc, m, L, g, x0, u0, b200::FpMode::Strict);
    for (size_t p : {size_t(0), size_t(100), P - 1}) {
        auto one = lin.linearize(0.0, {x0[p], x0[P + p], x0[2 * P + p], x0[3 * P + p]}, {u0[p]});
        for (size_t i = 0; i < 4; ++i) {
            for (size_t j = 0; j < 4; ++j) ASSERT_TRUE(one.A[i][j] == J.a(i, j, p));
            ASSERT_TRUE(one.B[i][0] == J.b(i, p));
        }
    }
    // plant validation keeps the reference's exceptions
    ASSERT_THROWS((control::CartNPendulum<1, double>(0.0, {0.1}, {0.5})), std::invalid_argument);
    ASSERT_THROWS((control::CartNPendulum<1, double>(1.0, {0.1}, {-0.5})), std::invalid_argument);

    // closed-loop callable (force depends on the state) cannot be a constant-force kernel: refused loudly
    auto plant = makeTypedODESystem<double>(control::CartNPendulum<1, double>(mc, {m}, {L}, g, {0.0, 0.1, 0.0, 0.0}));
    auto closed = [&plant](double t, StateView s) {
        plant.getComponent<0>().setControlForce(-10.0 * s[1]);
        return plant.computeDerivatives(t, std::vector<double>(s.begin(), s.end()));
    };
    ASSERT_THROWS(createRK4Solver().solve(closed, 4, 0.0, 0.5, 0.01, plant.getInitialState()), std::runtime_error);
    // open loop (constant force) is the compiled kernel
    plant.getComponent<0>().setControlForce(0.25);
    auto open = [&plant](double t, StateView s) { return plant.computeDerivatives(t, std::vector<double>(s.begin(), s.end())); };
    auto r = createRK4Solver().solve(open, 4, 0.0, 0.5, 0.01, plant.getInitialState());
    ASSERT_TRUE(r.size() == 51 && std::isfinite(r.states.back()[1]));
    std::puts("linearization_test: OK");
    return 0;
}
