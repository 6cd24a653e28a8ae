// Mirror of the reference's tests/compile_time_test.cpp (state functions at t = 0 :59-62, batch evaluation
// :84-101, energy drift < 0.01 % :153, analytic error < 1e-6 :190) against the B200 headers, with the
// reference's own call shapes: the wrapper lambda is passed to RK4Solver::solve unchanged.  Needs a B200.
#include <cstdio>
#include <iomanip>
#include <sopot/core/solver.hpp>
#include <sopot/core/typed_component.hpp>
#include <sopot/physics/harmonic_oscillator.hpp>
#include "check.hpp"

using namespace sopot;
using namespace sopot::physics;

int main() {
    // 1. single component
    auto oscillator = HarmonicOscillator<double>(1.0, 4.0, 0.0, 1.0, 0.0, "test_osc");
    auto single_system = makeTypedODESystem<double>(std::move(oscillator));
    static_assert(decltype(single_system)::template hasFunction<kinematics::Position>());
    static_assert(decltype(single_system)::template hasFunction<kinematics::Velocity>());
    static_assert(decltype(single_system)::template hasFunction<energy::Total>());
    static_assert(decltype(single_system)::template hasFunction<dynamics::Mass>());
    static_assert(!decltype(single_system)::template hasFunction<dynamics::Force>());
    ASSERT_TRUE(single_system.getComponentCount() == 1 && single_system.getStateDimension() == 2);
    auto state = single_system.getInitialState();
    ASSERT_NEAR(single_system.template computeStateFunction<kinematics::Position>(state), 1.0, 1e-10);
    ASSERT_NEAR(single_system.template computeStateFunction<kinematics::Velocity>(state), 0.0, 1e-10);
    ASSERT_NEAR(single_system.template computeStateFunction<energy::Total>(state), 2.0, 1e-10);
    ASSERT_NEAR(single_system.template computeStateFunction<dynamics::Mass>(state), 1.0, 1e-10);
    ASSERT_THROWS(HarmonicOscillator<double>(-1.0, 4.0), std::invalid_argument);
    ASSERT_THROWS(HarmonicOscillator<double>(1.0, 0.0), std::invalid_argument);
    ASSERT_THROWS(HarmonicOscillator<double>(1.0, 1.0, -0.1), std::invalid_argument);

    // 2./3. damped oscillator, batch evaluation
    auto damped_system = makeTypedODESystem<double>(createDampedOscillator<double>(2.0, 9.0, 0.1, 1.5));
    auto damped_state = damped_system.getInitialState();
    auto [pos, vel, ke, pe, total] = damped_system.template computeStateFunctions<
        kinematics::Position, kinematics::Velocity, energy::Kinetic, energy::Potential, energy::Total>(damped_state);
    ASSERT_NEAR(pos, 1.5, 1e-15);
    ASSERT_NEAR(vel, 0.0, 1e-15);
    ASSERT_NEAR(total, ke + pe, 1e-10);

    // one derivative evaluation on the device: {v, (-k/m) x - (c/m) v}
    auto d = damped_system.computeDerivatives(0.0, damped_state);
    ASSERT_TRUE(d[0] == 0.0 && d[1] == (-9.0 / 2.0) * 1.5 - (0.1 / 2.0) * 0.0);

    // 5. integration through the reference's wrapper lambda
    auto solver = createRK4Solver();
    auto derivs_func = [&single_system](double t, StateView s) {
        std::vector<double> v(s.begin(), s.end());
        return single_system.computeDerivatives(t, v);
    };
    auto result = solver.solve(derivs_func, single_system.getStateDimension(), 0.0, 2.0, 0.01, single_system.getInitialState());
    ASSERT_TRUE(result.size() == 201);
    const double e0 = single_system.template computeStateFunction<energy::Total>(result.states[0]);
    const double e1 = single_system.template computeStateFunction<energy::Total>(result.states.back());
    ASSERT_TRUE(std::abs(e1 - e0) / e0 * 100.0 < 0.01);

    // 6. analytic validation
    auto simple_system = makeTypedODESystem<double>(createSimpleOscillator<double>(1.0, 1.0, 1.0));
    auto simple_derivs = [&simple_system](double t, StateView s) {
        std::vector<double> v(s.begin(), s.end());
        return simple_system.computeDerivatives(t, v);
    };
    auto simple_result = solver.solve(simple_derivs, simple_system.getStateDimension(), 0.0, 1.0, 0.01, simple_system.getInitialState());
    const auto& osc_ref = simple_system.template getComponent<0>();
    double max_error = 0.0;
    for (size_t i = 0; i < simple_result.size(); ++i) {
        auto [ap, av] = osc_ref.analyticalSolution(simple_result.times[i]);
        max_error = std::max(max_error, std::abs(simple_result.states[i][0] - ap));
    }
    ASSERT_NEAR(max_error, 0.0, 1e-6);

    // 7. known answer: README oscillator through RK4Solver::solve, reference run on the host:
    // bit-identical in the solver's default strict mode, including the accumulated time stamp
    auto readme = solver.solve(damped_system, 0.0, 10.0, 0.01);
    ASSERT_TRUE(readme.size() == 1001);
    ASSERT_TRUE(readme.states.back()[0] == -0.82129045056399141);
    ASSERT_TRUE(readme.states.back()[1] == -1.7419130577838189);
    ASSERT_TRUE(readme.times.back() == 9.9999999999998312);

    // 8. a callable that is not a plain system evaluation has nothing to run on: no host integrator
    auto not_a_system = [](double, StateView s) { return std::vector<double>(s.size(), 1.0); };
    ASSERT_THROWS(solver.solve(not_a_system, 2, 0.0, 1.0, 0.01, std::vector<double>{0.0, 0.0}), std::runtime_error);
    auto post_processed = [&single_system](double t, StateView s) {
        std::vector<double> v(s.begin(), s.end());
        auto r = single_system.computeDerivatives(t, v);
        r[1] += 1.0;
        return r;
    };
    ASSERT_THROWS(solver.solve(post_processed, 2, 0.0, 1.0, 0.01, single_system.getInitialState()), std::runtime_error);
    // ... also when the difference only shows up mid-run: a forcing term that is zero where the callable is traced (t = 0)
    // and zero again at the end (a hat on [0.2, 0.8]), and a callable with hidden state that starts to deviate after 30 calls
    auto time_dependent = [&single_system](double t, StateView s) {
        auto r = single_system.computeDerivatives(t, std::vector<double>(s.begin(), s.end()));
        if (t > 0.2 && t < 0.8) r[1] += 0.5;
        return r;
    };
    ASSERT_THROWS(solver.solve(time_dependent, 2, 0.0, 1.0, 0.01, single_system.getInitialState()), std::runtime_error);
    int calls = 0;
    auto hidden_state = [&single_system, &calls](double t, StateView s) {
        auto r = single_system.computeDerivatives(t, std::vector<double>(s.begin(), s.end()));
        if (++calls == 3) r[0] *= 2.0;                // the traced call is #1, the first interior re-check #2
        return r;
    };
    ASSERT_THROWS(solver.solve(hidden_state, 2, 0.0, 1.0, 0.01, single_system.getInitialState()), std::runtime_error);

    // 9. ensemble: 1000 jittered oscillators in one launch == 1000 single solves, bit for bit (strict mode)
    const size_t n = 1000;
    std::vector<double> k(n), c(n);
    for (size_t i = 0; i < n; ++i) { k[i] = 9.0 * (1.0 + 0.05 * std::sin(double(i))); c[i] = 0.1 * (1.0 + 0.2 * std::cos(double(i))); }
    HarmonicOscillatorEnsemble ens{n, 2.0, k, c, 1.5, 0.0};
    b200::EnsembleOptions opt;
    opt.fp_mode = b200::FpMode::Strict;
    auto er = solver.solveEnsemble(ens, 0.0, 10.0, 0.01, opt);
    for (size_t i : {size_t(0), size_t(17), n - 1}) {
        auto sys_i = makeTypedODESystem<double>(HarmonicOscillator<double>(2.0, k[i], c[i], 1.5, 0.0));
        auto r_i = solver.solve(sys_i, 0.0, 10.0, 0.01);
        ASSERT_TRUE(r_i.states.back()[0] == er.final(0, i) && r_i.states.back()[1] == er.final(1, i));
    }
    std::puts("compile_time_test: OK");
    return 0;
}
