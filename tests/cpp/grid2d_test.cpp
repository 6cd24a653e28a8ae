// Grid tests in the shape of the reference's tests/grid_2d_test.cpp (3x3 quad :13-106, 4x4 cloth with
// diagonals :111-190) and tests/verify_grid_physics.cpp (equilibrium force < 1e-10 :76, hand-computed spring
// force :151).  The reference tests only print; the known answers here come from the unmodified templated
// makeGrid2DSystem compiled by oracle/ and integrated with RK4Solver::solve.  The grid kernel in strict mode
// is bit-identical to it (no libm call on this path).  Needs a B200.
#include <cstdio>
#include <sopot/physics/connected_masses/grid_2d.hpp>
#include <sopot/physics/unified/grid_system.hpp>
#include "check.hpp"

using namespace sopot;
using namespace sopot::connected_masses;

int main() {
    auto system = makeGrid2DSystem<double, 3, 3, false>(1.0, 1.0, 10.0, 0.5);
    ASSERT_TRUE(system.getStateDimension() == 36);
    auto state = system.getInitialState();
    for (size_t r = 0; r < 3; ++r)
        for (size_t c = 0; c < 3; ++c) {
            ASSERT_TRUE(state[(r * 3 + c) * 4 + 0] == double(c) && state[(r * 3 + c) * 4 + 1] == double(r));
            ASSERT_TRUE(state[(r * 3 + c) * 4 + 2] == 0.0 && state[(r * 3 + c) * 4 + 3] == 0.0);
        }
    // rest grid: no force anywhere
    for (double v : system.computeDerivatives(0.0, state)) ASSERT_NEAR(v, 0.0, 1e-10);

    // displace the centre mass by +0.5 in y (tests/grid_2d_test.cpp:70)
    state[4 * 4 + 1] += 0.5;
    auto d = system.computeDerivatives(0.0, state);
    // vertical neighbours: k * extension along y; horizontal ones stretch to sqrt(1.25)
    ASSERT_TRUE(d[1 * 4 + 3] == 5.0);
    ASSERT_TRUE(d[4 * 4 + 3] == -11.05572809000084);
    double sum_fx = 0.0, sum_fy = 0.0;
    for (size_t i = 0; i < 9; ++i) { sum_fx += d[i * 4 + 2]; sum_fy += d[i * 4 + 3]; }
    ASSERT_NEAR(sum_fx, 0.0, 1e-10);       // Newton's third law (verify_grid_physics.cpp:116)
    ASSERT_NEAR(sum_fy, 0.0, 1e-10);

    // 2 s at dt = 1e-3 through RK4Solver::solve with the reference's wrapper lambda
    auto derivs = [&system](double t, StateView s) { return system.computeDerivatives(t, std::vector<double>(s.begin(), s.end())); };
    auto r = createRK4Solver().solve(derivs, system.getStateDimension(), 0.0, 2.0, 1e-3, state);
    ASSERT_TRUE(r.size() == 2001);
    const double m4[4] = {1.0, 1.0591398935343284, 9.242763473708605e-15, 0.34510243258588474};
    const double m0[4] = {-0.020994531627621757, 0.04803146894323672, 0.0060584375046768705, 0.02911491181820517};
    for (int j = 0; j < 4; ++j) {
        ASSERT_TRUE(r.states.back()[4 * 4 + j] == m4[j]);
        ASSERT_TRUE(r.states.back()[j] == m0[j]);
    }
    // bulk entry point == stored-trajectory path
    auto bulk = state;
    system.integrate(bulk, 1e-3, 2000, b200::FpMode::Strict);
    ASSERT_TRUE(bulk == r.states.back());

    // 4x4 cloth with diagonals and the triangular variant (different edge orders, grid_2d.hpp:111-139, :274-281)
    auto cloth = makeGrid2DSystem<double, 4, 4, true>(0.1, 0.5, 50.0, 1.0);
    auto tri = makeTriangularGridSystem<double, 4, 4>(0.1, 0.5, 50.0, 1.0);
    auto s = cloth.getInitialState();
    for (size_t i = 0; i < 16; ++i) s[i * 4 + 3] = -2.0;
    s[5 * 4] += 0.05;
    auto s_tri = s;
    cloth.integrate(s, 5e-4, 400, b200::FpMode::Strict);
    tri.integrate(s_tri, 5e-4, 400, b200::FpMode::Strict);
    const double c5[4] = {0.5022594504116984, 0.09959025766244589, -3.914618797161261e-06, -1.996861584618476};
    const double t5[4] = {0.5022594504116984, 0.09959025766244589, -3.914618797164784e-06, -1.996861584618476};
    for (int j = 0; j < 4; ++j) { ASSERT_TRUE(s[5 * 4 + j] == c5[j]); ASSERT_TRUE(s_tri[5 * 4 + j] == t5[j]); }

    // runtime-sized twin, far beyond what the templated reference can instantiate; contract mode stays within
    // 1e-12 per step of strict
    auto big = makeGrid2DSystem(256, 320, 1.0, 0.5, 50.0, 0.5);
    auto y = big.getInitialState();
    y[(128 * 320 + 160) * 4] += 0.1;
    auto y_fast = y;
    big.integrate(y, 1e-3, 1, b200::FpMode::Strict);
    big.integrate(y_fast, 1e-3, 1, b200::FpMode::Contract);
    double err = 0.0, scale = 0.0;
    for (size_t i = 0; i < y.size(); ++i) { err = std::max(err, std::abs(y[i] - y_fast[i])); scale = std::max(scale, std::abs(y[i])); }
    ASSERT_TRUE(err <= 1e-12 * scale);
    ASSERT_THROWS(makeGrid2DSystem(1, 8, 1.0, 0.5, 50.0), std::invalid_argument);
    // unified 6-state grid (the reference's Grid2DSimulator system): known answers of ValidatedSystem on makeGridTopology<3,3>
    // with the wasm layout (radius 0.05, attachment angles 0 / pi / pi/2 / -pi/2), centre mass displaced and turned
    {
        unified::UnifiedGridSystem ug(3, 3);
        ASSERT_TRUE(ug.getStateDimension() == 54 && ug.getNodeCount() == 21);
        auto s = ug.getInitialState();
        ASSERT_TRUE(s[4 * 6 + 0] == 0.5 && s[4 * 6 + 1] == 0.5 && s[8 * 6 + 1] == 1.0);
        s[4 * 6 + 0] += 0.1; s[4 * 6 + 4] = 0.2;
        auto d = ug.computeDerivatives(0.0, s);
        const double d4[6] = {0.0, 0.0, -7.9027668235041935, 0.047077641230074185, 0.0, 163.52234344663196};
        for (int j = 0; j < 6; ++j) ASSERT_NEAR(d[4 * 6 + j], d4[j], 1e-12 * 164.0);
        auto derivs_u = [&ug](double t, StateView sv) { return ug.computeDerivatives(t, sv); };      // tests/unified_scalability_test.cpp:105-111
        auto ru = createRK4Solver().solve(derivs_u, ug.getStateDimension(), 0.0, 0.1, 0.001, s);
        ASSERT_TRUE(ru.size() == 101);
        const double m4[6] = {0.5660625774724096, 0.49981725468061716, -0.596435952017727, -0.010561942338296723, 1.3521740056572316, 25.18932655850739};
        const double m0[6] = {-0.02402791697931773, -0.02349146844165254, -0.4623865235560934, -0.4440399378395715, 0.013163793079625924, 0.5247417179459242};
        for (int j = 0; j < 6; ++j) { ASSERT_NEAR(ru.states.back()[4 * 6 + j], m4[j], 1e-9 * 26.0); ASSERT_NEAR(ru.states.back()[j], m0[j], 1e-9 * 26.0); }
        // the simulator's other two integrators, 100 steps of 1e-3 from the same state (reference loops of wasm_grid2d.cpp:329-413)
        const double se4[6] = {0.5657439557543504, 0.49981696742827497, -0.5967200284371587, -0.01049069666494754, 1.3655253482151337, 25.20979387132567};
        const double vv4[6] = {0.566042265060754, 0.49982221093902496, -0.5967051876102133, -0.01049151429179392, 1.3529916086711211, 25.208957483209815};
        for (auto fp : {b200::FpMode::Strict, b200::FpMode::Contract}) {
            auto ye = s, yv = s;
            ug.integrate(ye, 1e-3, 100, unified::Integrator::SymplecticEuler, fp);
            ug.integrate(yv, 1e-3, 100, unified::Integrator::VelocityVerlet, fp);
            for (int j = 0; j < 6; ++j) { ASSERT_NEAR(ye[4 * 6 + j], se4[j], 1e-9 * 26.0); ASSERT_NEAR(yv[4 * 6 + j], vv4[j], 1e-9 * 26.0); }
        }
        // the simulator's "triangle" grid type: both diagonals of every cell (makeTriangleGridTopology, createTriangleSystem)
        {
            unified::GridConfig tc; tc.triangle = true;
            unified::UnifiedGridSystem tri(3, 3, tc);
            ASSERT_TRUE(tri.getNodeCount() == 9 + 6 + 6 + 8);
            auto dt4 = tri.computeDerivatives(0.0, s);
            const double e4[6] = {0.0, 0.0, -16.043417617551782, 0.034385918659547876, 0.0, 324.8571509671915};
            for (int j = 0; j < 6; ++j) ASSERT_NEAR(dt4[4 * 6 + j], e4[j], 1e-12 * 325.0);
            auto yt = s;
            tri.integrate(yt, 1e-3, 100, b200::FpMode::Strict);
            const double r4[6] = {0.5354162697126642, 0.49921997063637474, -1.0503719607409456, -0.018609478585536958, 2.1729978058446977, 23.806551307485407};
            for (int j = 0; j < 6; ++j) ASSERT_NEAR(yt[4 * 6 + j], r4[j], 1e-9 * 24.0);
        }
        unified::UnifiedGridSystem large(512, 384);
        auto y = large.getInitialState();
        y[(256 * 384 + 192) * 6] += 0.1;
        large.integrate(y, 1e-3, 20);
        double moved = 0.0;
        for (size_t i = 0; i < y.size(); i += 6) moved = std::max(moved, std::abs(y[i + 4]));
        ASSERT_TRUE(std::isfinite(moved) && moved > 0.0);                                       // the disturbance turns the neighbours
        ASSERT_THROWS(unified::UnifiedGridSystem(1, 1), std::invalid_argument);
    }
    std::puts("grid2d_test: OK");
    return 0;
}
