// sopot/b200/trace.hpp -- how a host derivative callable names the compiled system it evaluates.
//
// The reference hands RK4Solver::solve an arbitrary callable, in practice always the wrapper lambda
//     [&system](double t, StateView s) { return system.computeDerivatives(t, {s.begin(), s.end()}); }
// (tests/compile_time_test.cpp:126-129, rocket/rocket.hpp:282-292).  A host callable cannot run inside a
// kernel and this library has no host integrator, so the solver calls the callable ONCE at (t0, y0) inside
// a TraceScope: TypedODESystem::computeDerivatives (device evaluation, N = 1) notices the scope and records
// a snapshot of its components together with launchers for the compiled RK4 kernel of that system.  The
// solver then checks that the callable returned exactly what the system returned (it is a plain
// evaluation, nothing was added on the host), launches the kernel, and afterwards re-checks the callable
// against the recorded snapshot at the final state (the callable did not change parameters on the way,
// e.g. a state-feedback force).  Anything else throws std::runtime_error: there is no CPU fallback.
#pragma once
#include <cstring>
#include <functional>
#include <vector>
#include "engine.hpp"

namespace sopot::b200::detail {

struct TraceRecord {
    int calls = 0;
    std::vector<double> returned;      // derivative the traced system call produced
    std::function<EnsembleResult(Engine&, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                                 const EnsembleOptions&)> solve;
    std::function<std::vector<double>(Engine&, const std::vector<double>& y)> rhs;   // snapshot parameters
};

inline TraceRecord*& active_trace() {
    static thread_local TraceRecord* t = nullptr;
    return t;
}

struct TraceScope {
    TraceRecord* prev;
    explicit TraceScope(TraceRecord* r) : prev(active_trace()) { active_trace() = r; }
    ~TraceScope() { active_trace() = prev; }
    TraceScope(const TraceScope&) = delete;
    TraceScope& operator=(const TraceScope&) = delete;
};

inline bool same_bits(const std::vector<double>& a, const std::vector<double>& b) {
    return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
}

}  // namespace sopot::b200::detail
