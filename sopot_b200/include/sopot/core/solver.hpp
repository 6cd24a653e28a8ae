// sopot/core/solver.hpp -- RK4Solver with the reference's interface (core/solver.hpp:12-184), executed by
// rk4_ensemble_kernel on the B200.
//
//   createRK4Solver().solve(derivs, dim, t0, t1, dt, y0)       the reference's call shape; `derivs` is the usual
//                                                              wrapper lambda around system.computeDerivatives
//   createRK4Solver().solve(system, t0, t1, dt)                shorthand: names the system directly
//   createRK4Solver().solveEnsemble(ensemble, t0, t1, dt, opt) NEW: millions of instances per launch
//
// A host callable cannot be launched on the device and this library has no host integrator.  The callable is
// therefore traced once (b200/trace.hpp) to find the compiled system it evaluates, the RK4 kernel of that system
// integrates all steps on the GPU, and the callable is re-checked against the traced system at three interior stored
// states and at the final one.
// A callable that is not a plain evaluation of one compiled system throws std::runtime_error.  Step count,
// accumulated time stamps and the stored trajectory follow core/solver.hpp:74-132 exactly
// (while (t < t_end - dt*0.5); t += dt; every step stored, initial state first).
#pragma once
#include <chrono>
#include <concepts>
#include <functional>
#include <span>
#include <string_view>
#include <vector>

#include "../b200/engine.hpp"
#include "typed_component.hpp"

namespace sopot {

using StateVector = std::vector<double>;
using StateDerivative = std::vector<double>;
using StateView = std::span<const double>;
using MutableStateView = std::span<double>;

struct SolutionResult {
    std::vector<double> times;
    std::vector<StateVector> states;
    double total_time{0.0};          // wall-clock milliseconds, as in the reference (:134-137)
    size_t size() const noexcept { return times.size(); }
    bool empty() const noexcept { return times.empty(); }
    void reserve(size_t c) { times.reserve(c); states.reserve(c); }
};

template <typename System>
concept ODESystemConcept = requires(const System& s) {
    { s.getStateDimension() } -> std::convertible_to<size_t>;
    { s.getInitialState() } -> std::convertible_to<StateVector>;
};

namespace b200 {
// stands where the reference passes a derivative lambda: identifies the compiled system to launch
template <typename System>
struct SystemDerivs { const System* system; };
template <typename System>
SystemDerivs<System> derivs(const System& s) { return SystemDerivs<System>{&s}; }
}  // namespace b200

class RK4Solver {
public:
    // reference signature: solve(derivs, dim, t0, t1, dt, y0)   (core/solver.hpp:63-71)
    template <typename System>
    SolutionResult solve(b200::SystemDerivs<System> derivs, size_t state_dim, double t_start, double t_end, double dt,
                         const StateVector& initial_state) const {
        using Binding = typename System::Binding;
        static_assert(Binding::available || (System::device_integrable && SOPOT_B200_DEVICE_COMPILER),
                      "no B200 kernel for this system (see core/typed_component.hpp)");
        const auto wall0 = std::chrono::high_resolution_clock::now();
        if (state_dim != System::getStateDimension() || initial_state.size() != state_dim)
            throw std::invalid_argument("state dimension mismatch");
        const size_t n_steps = sopot_b200_rk4_num_steps(t_start, t_end, dt);
        b200::EnsembleOptions opt;
        opt.fp_mode = m_fp_mode;
        opt.store_every = 1;                                  // the reference keeps every step (:130-131)
        b200::EnsembleResult r;
        if constexpr (Binding::available)
            r = Binding::solve(b200::Engine::shared(), derivs.system->components(), initial_state, n_steps, t_start, dt, opt);
        else
            r = b200::generic::solve_one(b200::Engine::shared(), *derivs.system, initial_state, n_steps, t_start, dt, opt);
        r.throwOnFailure();
        SolutionResult out = unpack(r, state_dim, n_steps, t_start, dt, initial_state);
        out.total_time = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - wall0).count();
        return out;
    }
    // zero initial state overload (:143-153)
    template <typename System>
    SolutionResult solve(b200::SystemDerivs<System> d, size_t dim, double t0, double t1, double dt) const {
        return solve(d, dim, t0, t1, dt, StateVector(dim, 0.0));
    }
    // system overload (:156-170); the derivative argument names the same system
    template <ODESystemConcept System>
    SolutionResult solve(const System& system, b200::SystemDerivs<System> d, double t0, double t1, double dt) const {
        return solve(d, system.getStateDimension(), t0, t1, dt, system.getInitialState());
    }
    template <ODESystemConcept System>
    SolutionResult solve(const System& system, double t0, double t1, double dt) const {
        return solve(b200::derivs(system), system.getStateDimension(), t0, t1, dt, system.getInitialState());
    }
    // The reference's own call shape: an arbitrary host callable (core/solver.hpp:63-71).  It is traced once to
    // find the compiled system it evaluates (b200/trace.hpp); a callable that is anything else than a plain
    // evaluation of one compiled system throws -- there is no host integrator to fall back to.
    template <typename F>
        requires std::invocable<F&, double, StateView> &&
                 std::convertible_to<std::invoke_result_t<F&, double, StateView>, StateDerivative>
    SolutionResult solve(F&& derivs, size_t state_dim, double t_start, double t_end, double dt,
                         const StateVector& initial_state) const {
        const auto wall0 = std::chrono::high_resolution_clock::now();
        if (initial_state.size() != state_dim) throw std::invalid_argument("state dimension mismatch");
        b200::detail::TraceRecord rec;
        StateDerivative d0;
        {
            b200::detail::TraceScope scope(&rec);
            d0 = derivs(t_start, StateView(initial_state));
        }
        if (rec.calls != 1 || !b200::detail::same_bits(d0, rec.returned))
            throw std::runtime_error("sopot-b200: the derivative callable is not a plain evaluation of one compiled "
                                     "system (computeDerivatives); there is no host integrator to run it");
        const size_t n_steps = sopot_b200_rk4_num_steps(t_start, t_end, dt);
        b200::EnsembleOptions opt;
        opt.fp_mode = m_fp_mode;
        opt.store_every = 1;
        b200::Engine& eng = b200::Engine::shared();
        b200::EnsembleResult r = rec.solve(eng, initial_state, n_steps, t_start, dt, opt);
        r.throwOnFailure();
        SolutionResult out = unpack(r, state_dim, n_steps, t_start, dt, initial_state);
        // The callable must describe the traced system along the WHOLE run, not only where it was traced: it is evaluated
        // again at three interior stored states and at the final one and compared with the traced system's own derivative
        // there.  A callable that depends on t (a forcing term that vanishes at t_start), on hidden state, or that changes a
        // parameter on the way (a feedback force) would otherwise be integrated silently as the traced system.
        const size_t last = out.size() - 1;
        size_t probes[4] = {last / 4, last / 2, (3 * last) / 4, last};
        size_t prev = 0;                                      // index 0 is the traced call itself
        for (size_t i : probes) {
            if (i == prev) continue;
            prev = i;
            if (derivs(out.times[i], StateView(out.states[i])) != rec.rhs(eng, out.states[i]))
                throw std::runtime_error("sopot-b200: the derivative callable is not the traced system along the run (it "
                                         "depends on t or hidden state, or changed system parameters during the integration, "
                                         "e.g. a feedback force); use the batched closed-loop entry points");
        }
        out.total_time = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - wall0).count();
        return out;
    }
    template <typename F>
        requires std::invocable<F&, double, StateView>
    SolutionResult solve(F&& derivs, size_t dim, double t0, double t1, double dt) const {
        return solve(std::forward<F>(derivs), dim, t0, t1, dt, StateVector(dim, 0.0));
    }
    template <ODESystemConcept System, typename F>
        requires std::invocable<F&, double, StateView>
    SolutionResult solve(const System& system, F&& derivs, double t0, double t1, double dt) const {
        return solve(std::forward<F>(derivs), system.getStateDimension(), t0, t1, dt, system.getInitialState());
    }

    // NEW (additive): one launch for a whole ensemble.  `Ensemble` is the per-system SoA parameter block
    // (HarmonicOscillatorEnsemble, DoublePendulumEnsemble, ...), see the physics headers.
    template <typename Ensemble>
    auto solveEnsemble(const Ensemble& e, double t_start, double t_end, double dt,
                                       b200::EnsembleOptions opt = {}, b200::Engine* engine = nullptr) const {
        const size_t n_steps = sopot_b200_rk4_num_steps(t_start, t_end, dt);
        return e.integrate(engine ? *engine : b200::Engine::shared(), n_steps, t_start, dt, opt);
    }

    // times and states exactly as core/solver.hpp:84-131 stores them (t accumulated, initial state first)
    static SolutionResult unpack(const b200::EnsembleResult& r, size_t state_dim, size_t n_steps, double t_start,
                                 double dt, const StateVector& initial_state) {
        SolutionResult out;
        out.reserve(n_steps + 1);
        double t = t_start;
        out.times.push_back(t);
        out.states.push_back(initial_state);
        for (size_t s = 0; s < n_steps; ++s) {
            t += dt;                                          // accumulated, not t0 + n*dt (:127)
            StateVector y(state_dim);
            for (size_t d = 0; d < state_dim; ++d) y[d] = r.at(s, d, 0);
            out.times.push_back(t);
            out.states.push_back(std::move(y));
        }
        return out;
    }

    RK4Solver& setFpMode(b200::FpMode m) { m_fp_mode = m; return *this; }
    constexpr std::string_view getSolverName() const { return "RK4"; }
    constexpr size_t getOrder() const { return 4; }
private:
    b200::FpMode m_fp_mode = b200::FpMode::Strict;   // single-instance API: bit-faithful association by default
};

inline RK4Solver createRK4Solver() { return RK4Solver{}; }

}  // namespace sopot
