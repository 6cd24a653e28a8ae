// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" driver around the UNMODIFIED reference headers, which are included
// from where they lie under /root/reference (-I/root/reference); no reference source is
// copied into this repository. It is compiled by oracle/Makefile into
// oracle/_ref/libsopot_ref.so with the reference's own arithmetic flags
// (g++ -std=c++20 -O3, no -march, no fast-math => no FMA contraction; CMakeLists.txt:7-9,33-35).
//
// Every entry point integrates through the reference's own public API exactly the way
// its tests do (tests/compile_time_test.cpp:126-136): makeTypedODESystem -> lambda that
// copies the StateView into a std::vector -> createRK4Solver().solve(...).
// Ensemble entry points fan instances out over std::thread (one system + one solver per
// thread: both are non re-entrant, core/solver.hpp:48-49).
#include "core/typed_component.hpp"
#include "core/solver.hpp"
#include "core/linearization.hpp"
#include "physics/harmonic_oscillator.hpp"
#include "physics/pendulum/double_pendulum.hpp"
#include "physics/control/cart_n_pendulum.hpp"
#include "physics/control/lqr.hpp"
#include "physics/coupled_oscillator/coupled_system.hpp"
#include "physics/unified/validated_system.hpp"
#include "physics/control/state_feedback_controller.hpp"
#include "physics/connected_masses/connectivity_matrix_2d.hpp"
#include "rocket/rocket.hpp"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

using namespace sopot;

namespace {

template <class F>
void parallel_for(size_t n, int nthreads, F&& body) {
    if (nthreads <= 1 || n <= 1) { for (size_t i = 0; i < n; ++i) body(i); return; }
    std::atomic<size_t> next{0};
    const size_t chunk = 16;
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                size_t e = b + chunk < n ? b + chunk : n;
                for (size_t i = b; i < e; ++i) body(i);
            }
        });
    for (auto& th : pool) th.join();
}

// Copy every `store_every`-th stored step (and always the last) of a SolutionResult.
// traj layout: [n_snap][dim][n_inst] ; final layout: [dim][n_inst]
template <class Sys>
SolutionResult solve_system(const Sys& sys, double t0, double t1, double dt) {
    auto solver = createRK4Solver();
    auto derivs = [&sys](double t, StateView sv) {
        std::vector<double> v(sv.begin(), sv.end());
        return sys.computeDerivatives(t, v);
    };
    return solver.solve(derivs, sys.getStateDimension(), t0, t1, dt, sys.getInitialState());
}

void scatter_result(const SolutionResult& r, size_t dim, size_t i, size_t n,
                    size_t store_every, double* final_out, double* traj_out, double* t_last) {
    const auto& last = r.states.back();
    for (size_t d = 0; d < dim; ++d) final_out[d * n + i] = last[d];
    if (t_last) t_last[i] = r.times.back();
    if (traj_out && store_every) {
        size_t snap = 0;
        for (size_t s = store_every; s < r.states.size(); s += store_every, ++snap)
            for (size_t d = 0; d < dim; ++d) traj_out[(snap * dim + d) * n + i] = r.states[s][d];
    }
}

} // namespace

// unified 6-state grid (physics/unified/): topologies must be objects with static storage to serve as template arguments
namespace ugrid_topo {
constexpr auto t3x3 = sopot::unified::makeGridTopology<3, 3>();
constexpr auto t4x5 = sopot::unified::makeGridTopology<4, 5>();
constexpr auto t2x7 = sopot::unified::makeGridTopology<2, 7>();
constexpr auto t6x6 = sopot::unified::makeGridTopology<6, 6>();
constexpr auto t10x10 = sopot::unified::makeGridTopology<10, 10>();
constexpr auto x3x3 = sopot::unified::makeTriangleGridTopology<3, 3>();
constexpr auto x4x5 = sopot::unified::makeTriangleGridTopology<4, 5>();
constexpr auto x2x7 = sopot::unified::makeTriangleGridTopology<2, 7>();
constexpr auto x6x6 = sopot::unified::makeTriangleGridTopology<6, 6>();
}  // namespace ugrid_topo

// populate exactly as Grid2DSimulator::createQuadSystem does (wasm/wasm_grid2d.cpp:417-462): masses row-major with radius,
// horizontal springs row-major, then vertical springs; attachment angles per orientation
template <auto& Topology>
static int ugrid_run(size_t rows, size_t cols, const double* prm /*18*/, double dt, size_t n_steps, const double* state_in, double* out,
                     int mode /*0 rhs, 1 rk4, 2 initial state*/) {
    using namespace sopot::unified;
    ValidatedSystem<double, Topology, PointMass2D<double>, Spring2D<double>> sys;
    const double mass = prm[0], radius = prm[1], spacing = prm[2], k = prm[3], L0 = prm[4], c = prm[5], mind = prm[6], krep = prm[7];
    auto& mb = sys.template getBatch<0>();
    for (size_t r = 0; r < rows; ++r)
        for (size_t cc = 0; cc < cols; ++cc) mb.add(PointMass2D<double>(mass, {double(cc) * spacing, double(r) * spacing}, {0.0, 0.0}, radius, 0.0, 0.0));
    auto& sb = sys.template getBatch<1>();
    for (size_t r = 0; r < rows; ++r)
        for (size_t cc = 0; cc + 1 < cols; ++cc) sb.add(Spring2D<double>(k, L0, c, mind, krep, prm[8], prm[9]));
    for (size_t r = 0; r + 1 < rows; ++r)
        for (size_t cc = 0; cc < cols; ++cc) sb.add(Spring2D<double>(k, L0, c, mind, krep, prm[10], prm[11]));
    if (prm[12] == 1.0)      // createTriangleSystem, wasm/wasm_grid2d.cpp:508-524: per cell the main diagonal, then the anti-diagonal
        for (size_t r = 0; r + 1 < rows; ++r)
            for (size_t cc = 0; cc + 1 < cols; ++cc) {
                sb.add(Spring2D<double>(k, prm[13], c, mind, krep, prm[14], prm[15]));
                sb.add(Spring2D<double>(k, prm[13], c, mind, krep, prm[16], prm[17]));
            }
    const size_t n = sys.getStateDimension();
    if (n != rows * cols * 6) return 3;
    if (mode == 2) { auto s0 = sys.getInitialState(); std::copy(s0.begin(), s0.end(), out); return 0; }
    std::vector<double> st(state_in, state_in + n);
    if (mode == 0) { auto d = sys.computeDerivatives(0.0, st); std::copy(d.begin(), d.end(), out); return 0; }
    if (mode == 3 || mode == 4) {
        // the two extra integrators of Grid2DSimulator (wasm/wasm_grid2d.cpp:329-413; that file needs emscripten, so its
        // three-line update loops are repeated here around the reference's own computeDerivatives)
        const size_t nm = rows * cols;
        double t = 0.0;
        for (size_t sidx = 0; sidx < n_steps; ++sidx) {
            auto d = sys.computeDerivatives(t, st);
            if (mode == 3) {                                                       // symplecticEulerStep :329-358
                for (size_t i = 0; i < nm; ++i) { st[i * 6 + 2] += dt * d[i * 6 + 2]; st[i * 6 + 3] += dt * d[i * 6 + 3]; st[i * 6 + 5] += dt * d[i * 6 + 5]; }
                for (size_t i = 0; i < nm; ++i) { st[i * 6 + 0] += dt * st[i * 6 + 2]; st[i * 6 + 1] += dt * st[i * 6 + 3]; st[i * 6 + 4] += dt * st[i * 6 + 5]; }
            } else {                                                               // velocityVerletStep :375-410
                const double half_dt_sq = 0.5 * dt * dt;
                for (size_t i = 0; i < nm; ++i) {
                    st[i * 6 + 0] += dt * st[i * 6 + 2] + half_dt_sq * d[i * 6 + 2];
                    st[i * 6 + 1] += dt * st[i * 6 + 3] + half_dt_sq * d[i * 6 + 3];
                    st[i * 6 + 4] += dt * st[i * 6 + 5] + half_dt_sq * d[i * 6 + 5];
                }
                auto dn = sys.computeDerivatives(t + dt, st);
                const double half_dt = 0.5 * dt;
                for (size_t i = 0; i < nm; ++i) {
                    st[i * 6 + 2] += half_dt * (d[i * 6 + 2] + dn[i * 6 + 2]);
                    st[i * 6 + 3] += half_dt * (d[i * 6 + 3] + dn[i * 6 + 3]);
                    st[i * 6 + 5] += half_dt * (d[i * 6 + 5] + dn[i * 6 + 5]);
                }
            }
            t += dt;
        }
        std::copy(st.begin(), st.end(), out);
        return 0;
    }
    auto solver = createRK4Solver();
    auto derivs = [&sys](double t, StateView sv) { return sys.computeDerivatives(t, sv); };
    double t = 0.0;
    // n_steps explicit steps through RK4Solver::solve (one step per call keeps the step count independent of rounding in t_end)
    for (size_t sidx = 0; sidx < n_steps; ++sidx) {
        auto r = solver.solve(derivs, n, t, t + dt, dt, st);
        st = r.states.back();
        t += dt;
    }
    std::copy(st.begin(), st.end(), out);
    return 0;
}

extern "C" {

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- config 1: damped harmonic oscillator (physics/harmonic_oscillator.hpp:44-85) ----
int ref_ho_rk4(size_t n, const double* m, const double* k, const double* c, const double* x0,
               const double* v0, double t0, double t1, double dt, size_t store_every,
               double* final_out, double* traj_out, double* t_last, int nthreads) {
    parallel_for(n, nthreads, [&](size_t i) {
        auto sys = makeTypedODESystem<double>(
            physics::HarmonicOscillator<double>(m[i], k[i], c[i], x0[i], v0[i]));
        auto r = solve_system(sys, t0, t1, dt);
        scatter_result(r, 2, i, n, store_every, final_out, traj_out, t_last);
    });
    return 0;
}

// ---- config 2: double pendulum (physics/pendulum/double_pendulum.hpp:68-179) ----
int ref_dpend_rk4(size_t n, const double* m1, const double* m2, const double* L1, const double* L2,
                  const double* g, const double* th1, const double* th2, const double* w1,
                  const double* w2, double t0, double t1, double dt, size_t store_every,
                  double* final_out, double* traj_out, double* t_last, int nthreads) {
    parallel_for(n, nthreads, [&](size_t i) {
        auto sys = makeTypedODESystem<double>(pendulum::DoublePendulum<double>(
            m1[i], m2[i], L1[i], L2[i], g[i], th1[i], th2[i], w1[i], w2[i]));
        auto r = solve_system(sys, t0, t1, dt);
        scatter_result(r, 4, i, n, store_every, final_out, traj_out, t_last);
    });
    return 0;
}

// derivative evaluation only (single RHS), for per-call parity
int ref_dpend_rhs(size_t n, const double* m1, const double* m2, const double* L1, const double* L2,
                  const double* g, const double* state /*[4][n]*/, double* out /*[4][n]*/) {
    for (size_t i = 0; i < n; ++i) {
        auto sys = makeTypedODESystem<double>(
            pendulum::DoublePendulum<double>(m1[i], m2[i], L1[i], L2[i], g[i]));
        std::vector<double> s{state[i], state[n + i], state[2 * n + i], state[3 * n + i]};
        auto d = sys.computeDerivatives(0.0, s);
        for (int j = 0; j < 4; ++j) out[j * n + i] = d[j];
    }
    return 0;
}

// ---- config 3: cart-pole linearization via Dual<double,5> ----
// makeLinearizer<4,1> (core/linearization.hpp:121-126) around CartNPendulum<1,Dual<double,5>>
// (physics/control/cart_n_pendulum.hpp:143-221) with setControlForce(u[0]) inside the lambda.
int ref_cartpole_linearize(size_t P, const double* x /*[4][P]*/, const double* u /*[P]*/,
                           const double* mc, const double* m, const double* L, const double* g,
                           double* A /*[16][P]*/, double* B /*[4][P]*/, int nthreads) {
    using Lin = SystemLinearizer<4, 1>;
    using D = Lin::DualT;
    parallel_for(P, nthreads, [&](size_t p) {
        control::CartNPendulum<1, D> plant(mc[p], {m[p]}, {L[p]}, g[p]);
        auto sys = makeTypedODESystem<D>(std::move(plant));
        auto lin = makeLinearizer<4, 1>(
            [&sys](D t, const Lin::DualState& xs, const Lin::DualInput& us) {
                sys.template getComponent<0>().setControlForce(us[0]);
                std::vector<D> sv(xs.begin(), xs.end());
                auto d = sys.computeDerivatives(t, sv);
                Lin::DualDerivative out;
                for (size_t i = 0; i < 4; ++i) out[i] = d[i];
                return out;
            });
        auto r = lin.linearize(0.0, {x[p], x[P + p], x[2 * P + p], x[3 * P + p]}, {u[p]});
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 4; ++j) A[(i * 4 + j) * P + p] = r.A[i][j];
            B[i * P + p] = r.B[i][0];
        }
    });
    return 0;
}

// plain-double cart-pole integration with a constant force (CartNPendulum<1,double>)
int ref_cartpole_rk4(size_t n, const double* mc, const double* m, const double* L, const double* g,
                     const double* s0 /*[4][n]*/, const double* force, double t0, double t1,
                     double dt, double* final_out, int nthreads) {
    parallel_for(n, nthreads, [&](size_t i) {
        control::CartNPendulum<1, double> plant(mc[i], {m[i]}, {L[i]}, g[i],
                                                {s0[i], s0[n + i], s0[2 * n + i], s0[3 * n + i]});
        plant.setControlForce(force[i]);
        auto sys = makeTypedODESystem<double>(std::move(plant));
        sys.template getComponent<0>().setControlForce(force[i]);
        auto r = solve_system(sys, t0, t1, dt);
        scatter_result(r, 4, i, n, 0, final_out, nullptr, nullptr);
    });
    return 0;
}

// ---- section 8f-1: LQR<4,1>::solve and the closed loop ----
// The reference's own LQR class on A, B as the linearizer produced them; status 0 = solve() true, 8 = false,
// 2 = solveForGain threw.
int ref_lqr4(size_t P, const double* A /*[16][P]*/, const double* B /*[4][P]*/, const double* Q16, double R, double dt,
             size_t max_iter, double tol, double* K /*[4][P]*/, double* Pric, int* iters_unused, int* status, int nthreads) {
    (void)iters_unused;
    parallel_for(P, nthreads, [&](size_t p) {
        control::LQR<4, 1>::StateMatrix Am, Qm;
        control::LQR<4, 1>::InputMatrix Bm;
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < 4; ++j) { Am[i][j] = A[(i * 4 + j) * P + p]; Qm[i][j] = Q16[i * 4 + j]; }
            Bm[i][0] = B[i * P + p];
        }
        control::LQR<4, 1>::InputWeightMatrix Rm{};
        Rm[0][0] = R;
        control::LQR<4, 1> lqr;
        int st = 0;
        try { st = lqr.solve(Am, Bm, Qm, Rm, dt, max_iter, tol) ? 0 : 8; } catch (const std::runtime_error&) { st = 2; }
        for (int j = 0; j < 4; ++j) K[j * P + p] = st ? NAN : lqr.getGain()[0][j];
        if (Pric) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Pric[(i * 4 + j) * P + p] = lqr.getRiccatiSolution()[i][j];
        if (status) status[p] = st;
    });
    return 0;
}

// StateFeedbackController<4,1>::compute -> setControlForce -> ONE RK4Solver::solve step with the force held,
// repeated n_steps times (the stepping pattern of wasm/wasm_inverted_pendulum.cpp:110-135 with core/solver.hpp's RK4).
int ref_cartpole_feedback_rk4(size_t n, const double* mc, const double* m, const double* L, const double* g,
                              const double* s0, const double* K, const double* xref, double umin, double umax, int saturate,
                              double t0, double dt, size_t n_steps, double* final_out, double* stats, int nthreads) {
    parallel_for(n, nthreads, [&](size_t i) {
        control::StateFeedbackController<4, 1>::GainMatrix Km{};
        std::array<double, 4> ref{};
        for (int q = 0; q < 4; ++q) { Km[0][q] = K[q * n + i]; ref[q] = xref ? xref[q * n + i] : 0.0; }
        control::StateFeedbackController<4, 1> ctl(Km);
        ctl.setReference(ref);
        if (saturate) ctl.setSaturationLimits({umin}, {umax});
        control::CartNPendulum<1, double> plant(mc[i], {m[i]}, {L[i]}, g[i]);
        auto sys = makeTypedODESystem<double>(std::move(plant));
        auto solver = createRK4Solver();
        std::vector<double> y = {s0[i], s0[n + i], s0[2 * n + i], s0[3 * n + i]};
        double t = t0, max_th = std::abs(y[1]), max_u = 0.0;
        auto derivs = [&sys](double tt, StateView sv) {
            std::vector<double> v(sv.begin(), sv.end());
            return sys.computeDerivatives(tt, v);
        };
        for (size_t step = 0; step < n_steps; ++step) {
            const double u = ctl.compute({y[0], y[1], y[2], y[3]})[0];
            max_u = std::max(max_u, std::abs(u));
            sys.template getComponent<0>().setControlForce(u);
            auto r = solver.solve(derivs, 4, t, t + dt, dt, y);
            y = r.states.back();
            t += dt;
            max_th = std::max(max_th, std::abs(y[1]));
        }
        for (int q = 0; q < 4; ++q) final_out[q * n + i] = y[q];
        if (stats) { stats[i] = max_th; stats[n + i] = max_u; }
    });
    return 0;
}

// ---- plugin-path check: the reference's own multi-component example (physics/coupled_oscillator/) ----
int ref_coupled_rk4(size_t n, const double* m1, const double* m2, const double* k, const double* L0, const double* c,
                    const double* x1, const double* x2, double t0, double t1, double dt, double* final_out /*[4][n]*/,
                    double* energy_out /*[n] total energy of the final state*/, int nthreads) {
    parallel_for(n, nthreads, [&](size_t i) {
        auto sys = physics::coupled::createCoupledOscillator<double>(m1[i], x1[i], 0.0, m2[i], x2[i], 0.0, k[i], L0[i], c[i]);
        auto r = solve_system(sys, t0, t1, dt);
        scatter_result(r, 4, i, n, 0, final_out, nullptr, nullptr);
        if (energy_out) energy_out[i] = sys.template computeStateFunction<physics::coupled::system::TotalEnergy>(r.states.back());
    });
    return 0;
}

// N-link cart pendulum, constant force (the reference's own tests use N = 2 and N = 6, tests/inverted_pendulum_test.cpp)
extern "C++" {
template <size_t NL>
static void cartn_one(const double* mc, const double* m, const double* L, const double* g, const double* s0, const double* force,
                      size_t n, size_t i, double t0, double t1, double dt, double* final_out) {
    std::array<double, NL> mm, LL;
    std::array<double, 2 * (NL + 1)> init;
    for (size_t q = 0; q < NL; ++q) { mm[q] = m[q * n + i]; LL[q] = L[q * n + i]; }
    for (size_t q = 0; q < 2 * (NL + 1); ++q) init[q] = s0[q * n + i];
    control::CartNPendulum<NL, double> plant(mc[i], mm, LL, g[i], init);
    auto sys = makeTypedODESystem<double>(std::move(plant));
    sys.template getComponent<0>().setControlForce(force[i]);
    auto r = solve_system(sys, t0, t1, dt);
    scatter_result(r, 2 * (NL + 1), i, n, 0, final_out, nullptr, nullptr);
}
}  // extern "C++"
int ref_cartn_rk4(int n_links, size_t n, const double* mc, const double* m /*[N][n]*/, const double* L /*[N][n]*/, const double* g,
                  const double* s0 /*[2(N+1)][n]*/, const double* force, double t0, double t1, double dt, double* final_out,
                  int nthreads) {
    if (n_links != 1 && n_links != 2 && n_links != 3 && n_links != 6) return -1;
    parallel_for(n, nthreads, [&](size_t i) {
        switch (n_links) {
            case 1: cartn_one<1>(mc, m, L, g, s0, force, n, i, t0, t1, dt, final_out); break;
            case 2: cartn_one<2>(mc, m, L, g, s0, force, n, i, t0, t1, dt, final_out); break;
            case 3: cartn_one<3>(mc, m, L, g, s0, force, n, i, t0, t1, dt, final_out); break;
            default: cartn_one<6>(mc, m, L, g, s0, force, n, i, t0, t1, dt, final_out); break;
        }
    });
    return 0;
}

// N-link linearization: makeLinearizer<2(N+1),1> around CartNPendulum<N, Dual<double, 2N+3>> (the pattern of
// ref_cartpole_linearize), A [(2N+2)^2][P], B [2N+2][P]
extern "C++" {
template <size_t NL>
static void cartn_lin_one(const double* mc, const double* m, const double* L, const double* g, const double* x, const double* u,
                          size_t P, size_t p, double* A, double* B) {
    constexpr size_t NS = 2 * (NL + 1);
    using Lin = SystemLinearizer<NS, 1>;
    using D = typename Lin::DualT;
    std::array<double, NL> mm, LL;
    for (size_t q = 0; q < NL; ++q) { mm[q] = m[q * P + p]; LL[q] = L[q * P + p]; }
    control::CartNPendulum<NL, D> plant(mc[p], mm, LL, g[p]);
    auto sys = makeTypedODESystem<D>(std::move(plant));
    auto lin = makeLinearizer<NS, 1>([&sys](D t, const typename Lin::DualState& xs, const typename Lin::DualInput& us) {
        sys.template getComponent<0>().setControlForce(us[0]);
        std::vector<D> sv(xs.begin(), xs.end());
        auto d = sys.computeDerivatives(t, sv);
        typename Lin::DualDerivative out;
        for (size_t i = 0; i < NS; ++i) out[i] = d[i];
        return out;
    });
    Vector<NS> x0;
    for (size_t i = 0; i < NS; ++i) x0[i] = x[i * P + p];
    auto r = lin.linearize(0.0, x0, {u[p]});
    for (size_t i = 0; i < NS; ++i) {
        for (size_t j = 0; j < NS; ++j) A[(i * NS + j) * P + p] = r.A[i][j];
        B[i * P + p] = r.B[i][0];
    }
}
}  // extern "C++"
int ref_cartn_linearize(int n_links, size_t P, const double* mc, const double* m, const double* L, const double* g,
                        const double* x, const double* u, double* A, double* B, int nthreads) {
    if (n_links != 1 && n_links != 2 && n_links != 6) return -1;
    parallel_for(P, nthreads, [&](size_t p) {
        switch (n_links) {
            case 1: cartn_lin_one<1>(mc, m, L, g, x, u, P, p, A, B); break;
            case 2: cartn_lin_one<2>(mc, m, L, g, x, u, P, p, A, B); break;
            default: cartn_lin_one<6>(mc, m, L, g, x, u, P, p, A, B); break;
        }
    });
    return 0;
}

// LQR<NS,1>::solve for NS = 4, 6, 14 and the N-link closed loop (CartNPendulumController-style state feedback, force held
// over one RK4Solver step)
extern "C++" {
template <size_t NS>
static void lqr_one(const double* A, const double* B, const double* Q, double R, double dt, size_t max_iter, double tol, size_t P,
                    size_t p, double* K, double* Pric, int* status) {
    typename control::LQR<NS, 1>::StateMatrix Am, Qm;
    typename control::LQR<NS, 1>::InputMatrix Bm;
    for (size_t i = 0; i < NS; ++i) {
        for (size_t j = 0; j < NS; ++j) { Am[i][j] = A[(i * NS + j) * P + p]; Qm[i][j] = Q[i * NS + j]; }
        Bm[i][0] = B[i * P + p];
    }
    typename control::LQR<NS, 1>::InputWeightMatrix Rm{};
    Rm[0][0] = R;
    control::LQR<NS, 1> lqr;
    int st = 0;
    try { st = lqr.solve(Am, Bm, Qm, Rm, dt, max_iter, tol) ? 0 : 8; } catch (const std::runtime_error&) { st = 2; }
    for (size_t j = 0; j < NS; ++j) K[j * P + p] = st ? NAN : lqr.getGain()[0][j];
    if (Pric) for (size_t i = 0; i < NS; ++i) for (size_t j = 0; j < NS; ++j) Pric[(i * NS + j) * P + p] = lqr.getRiccatiSolution()[i][j];
    if (status) status[p] = st;
}
template <size_t NL>
static void cartn_feedback_one(const double* mc, const double* m, const double* L, const double* g, const double* s0, const double* K,
                               const double* xref, double umin, double umax, int saturate, double t0, double dt, size_t n_steps,
                               size_t n, size_t i, double* final_out, double* stats) {
    constexpr size_t NS = 2 * (NL + 1);
    typename control::StateFeedbackController<NS, 1>::GainMatrix Km{};
    std::array<double, NS> ref{};
    for (size_t q = 0; q < NS; ++q) { Km[0][q] = K[q * n + i]; ref[q] = xref ? xref[q * n + i] : 0.0; }
    control::StateFeedbackController<NS, 1> ctl(Km);
    ctl.setReference(ref);
    if (saturate) ctl.setSaturationLimits({umin}, {umax});
    std::array<double, NL> mm, LL;
    for (size_t q = 0; q < NL; ++q) { mm[q] = m[q * n + i]; LL[q] = L[q * n + i]; }
    control::CartNPendulum<NL, double> plant(mc[i], mm, LL, g[i]);
    auto sys = makeTypedODESystem<double>(std::move(plant));
    auto solver = createRK4Solver();
    std::vector<double> y(NS);
    for (size_t q = 0; q < NS; ++q) y[q] = s0[q * n + i];
    double t = t0, max_th = 0.0, max_u = 0.0;
    for (size_t q = 1; q <= NL; ++q) max_th = std::max(max_th, std::abs(y[q]));
    auto derivs = [&sys](double tt, StateView sv) { std::vector<double> v(sv.begin(), sv.end()); return sys.computeDerivatives(tt, v); };
    for (size_t step = 0; step < n_steps; ++step) {
        std::array<double, NS> ya;
        for (size_t q = 0; q < NS; ++q) ya[q] = y[q];
        const double u = ctl.compute(ya)[0];
        max_u = std::max(max_u, std::abs(u));
        sys.template getComponent<0>().setControlForce(u);
        auto r = solver.solve(derivs, NS, t, t + dt, dt, y);
        y = r.states.back();
        t += dt;
        for (size_t q = 1; q <= NL; ++q) max_th = std::max(max_th, std::abs(y[q]));
    }
    for (size_t q = 0; q < NS; ++q) final_out[q * n + i] = y[q];
    if (stats) { stats[i] = max_th; stats[n + i] = max_u; }
}
}  // extern "C++"
int ref_lqr_n(int ns, size_t P, const double* A, const double* B, const double* Q, double R, double dt, size_t max_iter, double tol,
              double* K, double* Pric, int* iters_unused, int* status, int nthreads) {
    (void)iters_unused;
    if (ns != 4 && ns != 6 && ns != 14) return -1;
    parallel_for(P, nthreads, [&](size_t p) {
        if (ns == 4) lqr_one<4>(A, B, Q, R, dt, max_iter, tol, P, p, K, Pric, status);
        else if (ns == 6) lqr_one<6>(A, B, Q, R, dt, max_iter, tol, P, p, K, Pric, status);
        else lqr_one<14>(A, B, Q, R, dt, max_iter, tol, P, p, K, Pric, status);
    });
    return 0;
}
int ref_cartn_feedback_rk4(int n_links, size_t n, const double* mc, const double* m, const double* L, const double* g, const double* s0,
                           const double* K, const double* xref, double umin, double umax, int saturate, double t0, double dt,
                           size_t n_steps, double* final_out, double* stats, int nthreads) {
    if (n_links != 1 && n_links != 2 && n_links != 6) return -1;
    parallel_for(n, nthreads, [&](size_t i) {
        if (n_links == 1) cartn_feedback_one<1>(mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, n, i, final_out, stats);
        else if (n_links == 2) cartn_feedback_one<2>(mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, n, i, final_out, stats);
        else cartn_feedback_one<6>(mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, n, i, final_out, stats);
    });
    return 0;
}

// computeJacobian of the reference (core/typed_component.hpp:486-512) on DoublePendulum<Dual<double,4>>, as
// tests/double_pendulum_test.cpp:383-412 calls it; J[16][n] row-major 4x4 per point, SoA over the points
void ref_dpend_jacobian(size_t n, const double* m1, const double* m2, const double* L1, const double* L2, const double* g,
                        const double* state, double* J) {
    using D4 = sopot::Dual<double, 4>;
    for (size_t p = 0; p < n; ++p) {
        auto sys = sopot::makeTypedODESystem<D4>(sopot::pendulum::DoublePendulum<D4>(m1[p], m2[p], L1[p], L2[p], g[p], 0.0, 0.0, 0.0, 0.0));
        const std::array<double, 4> x{state[p], state[n + p], state[2 * n + p], state[3 * n + p]};
        const auto jac = sopot::computeJacobian(sys, 0.0, x);
        for (size_t i = 0; i < 4; ++i)
            for (size_t j = 0; j < 4; ++j) J[(i * 4 + j) * n + p] = jac[i][j];
    }
}

// prm[18]: mass, radius, spacing, stiffness, rest_length, damping, min_distance, repulsion_stiffness, h_attach0, h_attach1, v_attach0,
// v_attach1, topology (0 quad / 1 triangle), diagonal rest length, main-diagonal attach 0/1, anti-diagonal attach 0/1
int ref_ugrid(size_t rows, size_t cols, const double* prm, double dt, size_t n_steps, const double* state_in, double* out, int mode) {
#define UG_CASE(R, C, T) if (rows == R && cols == C) return ugrid_run<ugrid_topo::T>(rows, cols, prm, dt, n_steps, state_in, out, mode);
    if (prm[12] == 1.0) {
        UG_CASE(3, 3, x3x3) UG_CASE(4, 5, x4x5) UG_CASE(2, 7, x2x7) UG_CASE(6, 6, x6x6)
        return -1;
    }
    UG_CASE(3, 3, t3x3) UG_CASE(4, 5, t4x5) UG_CASE(2, 7, t2x7) UG_CASE(6, 6, t6x6) UG_CASE(10, 10, t10x10)
#undef UG_CASE
    return -1;
}

// ---- config 4: 6-DOF rocket ----
// The 11 components of Rocket<T>::SystemType (rocket/rocket.hpp:73-85) composed by hand so
// that gravity can be jittered (Rocket<T> has no gravity setter). Tables reach the reference
// only through CSV files (rocket/interpolated_engine.hpp:61-85, rocket_body.hpp:55-85), so
// they are written with %.17g (round-trips through std::stod exactly).
static std::string write_csv(const std::string& dir, const char* name, size_t rows, size_t cols,
                             const double* data /*row-major*/) {
    std::string path = dir + "/" + name;
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return path;
    for (size_t r = 0; r < rows; ++r) {
        for (size_t c = 0; c < cols; ++c)
            std::fprintf(f, c ? ",%.17g" : "%.17g", data[r * cols + c]);
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return path;
}

// aero blob (same layout as the port): engine_on, then 7 x (rows, cols, row_vals, col_vals, data) in the order cd_on, cd_off,
// cl_on, cl_off, cs_on, cs_off, cmq_yz.  Tables reach AxisymmetricAerodynamics only through its CSV loaders
// (rocket/axisymmetric_aero.hpp:73-111): first line = column headers behind a placeholder cell, then one line per row.
static void apply_aero_blob(sopot::rocket::AxisymmetricAerodynamics<double>& aero, const double* blob, const std::string& dir,
                            std::vector<std::string>& files) {
    if (!blob) return;
    aero.setEngineOn(blob[0] != 0.0);
    const double* q = blob + 1;
    for (int k = 0; k < 7; ++k) {
        const size_t rows = (size_t)q[0], cols = (size_t)q[1];
        q += 2;
        const double *rv = q, *cv = q + rows, *data = q + rows + cols;
        q += rows + cols + rows * cols;
        if (!rows) continue;
        const std::string path = dir + "/aero" + std::to_string(k) + ".csv";
        FILE* f = std::fopen(path.c_str(), "w");
        std::fprintf(f, "0");
        for (size_t c = 0; c < cols; ++c) std::fprintf(f, ",%.17g", cv[c]);
        std::fprintf(f, "\n");
        for (size_t r = 0; r < rows; ++r) {
            std::fprintf(f, "%.17g", rv[r]);
            for (size_t c = 0; c < cols; ++c) std::fprintf(f, ",%.17g", data[r * cols + c]);
            std::fprintf(f, "\n");
        }
        std::fclose(f);
        files.push_back(path);
        switch (k) {
            case 0: aero.loadCdPowerOn(path); break;
            case 1: aero.loadCdPowerOff(path); break;
            case 2: aero.loadClPowerOn(path); break;
            case 3: aero.loadClPowerOff(path); break;
            case 4: aero.loadCsPowerOn(path); break;
            case 5: aero.loadCsPowerOff(path); break;
            default: aero.loadDampingYZ(path); break;
        }
    }
}

struct RocketTables {
    std::string dir;
    std::string engine, mass, ix, iyz;
};

// stats layout [8][n]: apogee, apogee_t, max_speed, max_speed_t, burnout_alt, burnout_speed,
// downrange_E (final state[1]), downrange_N (final state[2])   (rocket/simulation_result.hpp:151-189)
int ref_rocket_rk4_aero(size_t n, const double* elev_deg, const double* az_deg, const double* diameter,
                   const double* g0, size_t engine_rows, const double* engine /*[rows][8]*/,
                   size_t mass_rows, const double* mass_tab /*[rows][2]*/,
                   size_t ix_rows, const double* ix_tab, size_t iyz_rows, const double* iyz_tab,
                   double t_end, double dt, size_t store_every, double* final_out /*[14][n]*/,
                   double* traj_out, double* stats_out /*[8][n] or null*/, int nthreads, const double* aero_blob) {
    using namespace sopot::rocket;
    char tmpl[] = "/tmp/sopot_ref_XXXXXX";
    char* d = mkdtemp(tmpl);
    if (!d) return 1;
    std::string dir(d);
    std::string f_eng = engine_rows ? write_csv(dir, "sim_engine.csv", engine_rows, 8, engine) : "";
    std::string f_mass = mass_rows ? write_csv(dir, "sim_mass.csv", mass_rows, 2, mass_tab) : "";
    std::string f_ix = ix_rows ? write_csv(dir, "sim_ix.csv", ix_rows, 2, ix_tab) : "";
    std::string f_iyz = iyz_rows ? write_csv(dir, "sim_iyz.csv", iyz_rows, 2, iyz_tab) : "";

    // load tables once; components are copied per instance
    InterpolatedEngine<double> engine0;
    if (engine_rows) engine0.loadFromFile(f_eng);
    RocketBody<double> body0;
    if (mass_rows) body0.loadMass(f_mass);
    if (ix_rows) body0.loadInertiaX(f_ix);
    if (iyz_rows) body0.loadInertiaYZ(f_iyz);
    const double burn = engine0.getBurnTime();
    AxisymmetricAerodynamics<double> aero0;
    std::vector<std::string> aero_files;
    apply_aero_blob(aero0, aero_blob, dir, aero_files);

    parallel_for(n, nthreads, [&](size_t i) {
        RotationKinematics<double> att;
        att.setInitialFromLauncherAngles(elev_deg[i], az_deg[i]);
        AxisymmetricAerodynamics<double> aero_i(aero0);
        aero_i.setReferenceDiameter(diameter[i]);
        auto sys = makeTypedODESystem<double>(
            SimulationTime<double>(), TranslationKinematics<double>(), TranslationDynamics<double>(),
            std::move(att), RotationDynamics<double>(), StandardAtmosphere<double>(),
            Gravity<double>(g0[i]), RocketBody<double>(body0), InterpolatedEngine<double>(engine0),
            std::move(aero_i), ForceAggregator<double>());
        auto r = solve_system(sys, 0.0, t_end, dt);
        scatter_result(r, 14, i, n, store_every, final_out, traj_out, nullptr);
        if (stats_out) {
            // same scan as SimulationResult::computeStatistics (rocket/simulation_result.hpp:151-189)
            double apogee = 0, apogee_t = 0, vmax = 0, vmax_t = 0, bo_alt = 0, bo_spd = 0;
            for (size_t s = 0; s < r.size(); ++s) {
                double t = r.times[s];
                const auto& st = r.states[s];
                double alt = st[3];
                double spd = std::sqrt(st[4] * st[4] + st[5] * st[5] + st[6] * st[6]);
                if (alt > apogee) { apogee = alt; apogee_t = t; }
                if (spd > vmax) { vmax = spd; vmax_t = t; }
                if (burn > 0 && std::abs(t - burn) < 0.1) { bo_alt = alt; bo_spd = spd; }
            }
            const auto& last = r.states.back();
            double v[8] = {apogee, apogee_t, vmax, vmax_t, bo_alt, bo_spd, last[1], last[2]};
            for (int j = 0; j < 8; ++j) stats_out[j * n + i] = v[j];
        }
    });
    for (auto* p : {&f_eng, &f_mass, &f_ix, &f_iyz}) if (!p->empty()) unlink(p->c_str());
    for (auto& f : aero_files) unlink(f.c_str());
    rmdir(dir.c_str());
    return 0;
}
int ref_rocket_rk4(size_t n, const double* elev_deg, const double* az_deg, const double* diameter, const double* g0,
                   size_t engine_rows, const double* engine, size_t mass_rows, const double* mass_tab, size_t ix_rows,
                   const double* ix_tab, size_t iyz_rows, const double* iyz_tab, double t_end, double dt, size_t store_every,
                   double* final_out, double* traj_out, double* stats_out, int nthreads) {
    return ref_rocket_rk4_aero(n, elev_deg, az_deg, diameter, g0, engine_rows, engine, mass_rows, mass_tab, ix_rows, ix_tab,
                               iyz_rows, iyz_tab, t_end, dt, store_every, final_out, traj_out, stats_out, nthreads, nullptr);
}

// single RHS evaluation of the full rocket at given states [14][n] (shared config)
int ref_rocket_rhs_aero(size_t n, double elev_deg, double az_deg, double diameter, double g0,
                   size_t engine_rows, const double* engine, size_t mass_rows, const double* mass_tab,
                   const double* state, double* out, const double* aero_blob) {
    using namespace sopot::rocket;
    char tmpl[] = "/tmp/sopot_ref_XXXXXX";
    char* d = mkdtemp(tmpl);
    if (!d) return 1;
    std::string dir(d);
    std::string f_eng = write_csv(dir, "sim_engine.csv", engine_rows, 8, engine);
    std::string f_mass = write_csv(dir, "sim_mass.csv", mass_rows, 2, mass_tab);
    InterpolatedEngine<double> engine0; engine0.loadFromFile(f_eng);
    RocketBody<double> body0; body0.loadMass(f_mass);
    RotationKinematics<double> att; att.setInitialFromLauncherAngles(elev_deg, az_deg);
    AxisymmetricAerodynamics<double> aero0(diameter);
    std::vector<std::string> aero_files;
    apply_aero_blob(aero0, aero_blob, dir, aero_files);
    auto sys = makeTypedODESystem<double>(
        SimulationTime<double>(), TranslationKinematics<double>(), TranslationDynamics<double>(),
        std::move(att), RotationDynamics<double>(), StandardAtmosphere<double>(),
        Gravity<double>(g0), std::move(body0), std::move(engine0),
        std::move(aero0), ForceAggregator<double>());
    for (size_t i = 0; i < n; ++i) {
        std::vector<double> s(14);
        for (int j = 0; j < 14; ++j) s[j] = state[j * n + i];
        auto dv = sys.computeDerivatives(0.0, s);
        for (int j = 0; j < 14; ++j) out[j * n + i] = dv[j];
    }
    unlink(f_eng.c_str()); unlink(f_mass.c_str());
    for (auto& f : aero_files) unlink(f.c_str());
    rmdir(dir.c_str());
    return 0;
}
// the state functions of rocket/rocket_tags.hpp that the registry of the full rocket provides, at given states [14][n]:
// out[28][n] in the row order of sopot_b200_rocket_state_functions (include/sopot_b200.h)
int ref_rocket_state_functions(size_t n, double elev_deg, double az_deg, double diameter, double g0,
                               size_t engine_rows, const double* engine, size_t mass_rows, const double* mass_tab,
                               const double* state, double* out, const double* aero_blob) {
    using namespace sopot::rocket;
    char tmpl[] = "/tmp/sopot_ref_XXXXXX";
    char* d = mkdtemp(tmpl);
    if (!d) return 1;
    std::string dir(d);
    std::string f_eng = write_csv(dir, "sim_engine.csv", engine_rows, 8, engine);
    std::string f_mass = write_csv(dir, "sim_mass.csv", mass_rows, 2, mass_tab);
    InterpolatedEngine<double> engine0; engine0.loadFromFile(f_eng);
    RocketBody<double> body0; body0.loadMass(f_mass);
    RotationKinematics<double> att; att.setInitialFromLauncherAngles(elev_deg, az_deg);
    AxisymmetricAerodynamics<double> aero0(diameter);
    std::vector<std::string> aero_files;
    apply_aero_blob(aero0, aero_blob, dir, aero_files);
    auto sys = makeTypedODESystem<double>(
        SimulationTime<double>(), TranslationKinematics<double>(), TranslationDynamics<double>(),
        std::move(att), RotationDynamics<double>(), StandardAtmosphere<double>(),
        Gravity<double>(g0), std::move(body0), std::move(engine0),
        std::move(aero0), ForceAggregator<double>());
    for (size_t i = 0; i < n; ++i) {
        std::vector<double> s(14);
        for (int j = 0; j < 14; ++j) s[j] = state[j * n + i];
        double r[28];
        r[0] = sys.computeStateFunction<sopot::rocket::dynamics::Mass>(s);
        const auto I = sys.computeStateFunction<sopot::rocket::dynamics::MomentOfInertia>(s);
        const auto F = sys.computeStateFunction<sopot::rocket::dynamics::TotalForceENU>(s);
        const auto Tq = sys.computeStateFunction<sopot::rocket::dynamics::TotalTorqueBody>(s);
        const auto Fab = sys.computeStateFunction<sopot::rocket::aero::AeroForceBody>(s);
        const auto Fae = sys.computeStateFunction<sopot::rocket::aero::AeroForceENU>(s);
        const auto Mab = sys.computeStateFunction<sopot::rocket::aero::AeroMomentBody>(s);
        const auto Th = sys.computeStateFunction<sopot::rocket::propulsion::ThrustForceBody>(s);
        for (int k = 0; k < 3; ++k) {
            r[1 + k] = I[k]; r[4 + k] = F[k]; r[7 + k] = Tq[k]; r[11 + k] = Fab[k]; r[14 + k] = Fae[k]; r[17 + k] = Mab[k]; r[25 + k] = Th[k];
        }
        r[10] = sys.computeStateFunction<sopot::rocket::dynamics::GravityAcceleration>(s);
        r[20] = sys.computeStateFunction<sopot::rocket::environment::AtmosphericDensity>(s);
        r[21] = sys.computeStateFunction<sopot::rocket::environment::AtmosphericPressure>(s);
        r[22] = sys.computeStateFunction<sopot::rocket::environment::AtmosphericTemperature>(s);
        r[23] = sys.computeStateFunction<sopot::rocket::environment::SpeedOfSound>(s);
        r[24] = sys.computeStateFunction<sopot::rocket::propulsion::MassFlowRate>(s);
        for (int k = 0; k < 28; ++k) out[k * n + i] = r[k];
    }
    unlink(f_eng.c_str()); unlink(f_mass.c_str());
    for (auto& f : aero_files) unlink(f.c_str());
    rmdir(dir.c_str());
    return 0;
}
int ref_rocket_rhs(size_t n, double elev_deg, double az_deg, double diameter, double g0, size_t engine_rows, const double* engine,
                   size_t mass_rows, const double* mass_tab, const double* state, double* out) {
    return ref_rocket_rhs_aero(n, elev_deg, az_deg, diameter, g0, engine_rows, engine, mass_rows, mass_tab, state, out, nullptr);
}

// engine precompute columns (pressure ratio, mass flow, engine force) as the reference derives
// them at load time (rocket/interpolated_engine.hpp:176-252), observed through the public API:
// not directly exposed, so we expose thrust magnitude at (t, P_atm) instead.
int ref_engine_thrust(size_t engine_rows, const double* engine, size_t nq, const double* tq,
                      const double* patm, double* thrust) {
    using namespace sopot::rocket;
    char tmpl[] = "/tmp/sopot_ref_XXXXXX";
    char* d = mkdtemp(tmpl);
    if (!d) return 1;
    std::string dir(d);
    std::string f_eng = write_csv(dir, "sim_engine.csv", engine_rows, 8, engine);
    InterpolatedEngine<double> e; e.loadFromFile(f_eng);
    for (size_t i = 0; i < nq; ++i) thrust[i] = e.computeThrustMagnitude(tq[i], patm[i]);
    unlink(f_eng.c_str()); rmdir(dir.c_str());
    return 0;
}

int ref_atmosphere(size_t n, const double* h, double* T, double* P, double* rho, double* a) {
    sopot::rocket::StandardAtmosphere<double> atm;
    for (size_t i = 0; i < n; ++i) {
        T[i] = atm.computeTemperature(h[i]); P[i] = atm.computePressure(h[i]);
        rho[i] = atm.computeDensity(h[i]); a[i] = atm.computeSpeedOfSound(h[i]);
    }
    return 0;
}

// ---- config 5: templated grid2d (physics/connected_masses/connectivity_matrix_2d.hpp:156-198) ----
// Only small sizes are instantiable (6x6 takes 101 s to compile); topology: 0 quad, 1 quad+diagonals
// (makeGrid2DSystem<...,true>), 2 triangular (makeTriangularGridSystem).
} // extern "C"
namespace {
template <size_t R, size_t C, int Topo>
auto make_grid(double mass, double spacing, double k, double c) {
    using namespace sopot::connected_masses;
    if constexpr (Topo == 0) return makeGrid2DSystem<double, R, C, false>(mass, spacing, k, c);
    else if constexpr (Topo == 1) return makeGrid2DSystem<double, R, C, true>(mass, spacing, k, c);
    else return makeTriangularGridSystem<double, R, C>(mass, spacing, k, c);
}
template <size_t R, size_t C, int Topo>
int grid_run(double mass, double spacing, double k, double c, const double* state, int mode,
             double dt, size_t n_steps, double* out) {
    auto sys = make_grid<R, C, Topo>(mass, spacing, k, c);
    const size_t dim = 4 * R * C;
    std::vector<double> s(state, state + dim);
    if (mode == 0) {                       // one RHS
        auto d = sys.computeDerivatives(0.0, s);
        std::memcpy(out, d.data(), dim * sizeof(double));
    } else if (mode == 1) {                // RK4Solver::solve for n_steps
        auto solver = createRK4Solver();
        auto derivs = [&sys](double t, StateView sv) {
            std::vector<double> v(sv.begin(), sv.end());
            return sys.computeDerivatives(t, v);
        };
        auto r = solver.solve(derivs, dim, 0.0, dt * (double)n_steps, dt, s);
        std::memcpy(out, r.states.back().data(), dim * sizeof(double));
        return (int)(r.states.size() - 1);
    } else {                               // initial state
        auto s0 = sys.getInitialState();
        std::memcpy(out, s0.data(), dim * sizeof(double));
    }
    return 0;
}
} // namespace

extern "C" {
#define GRID_CASE(R, C, T) if (rows == R && cols == C && topo == T) \
    return grid_run<R, C, T>(mass, spacing, k, c, state, mode, dt, n_steps, out);

// returns -1 when (rows, cols, topo) is not one of the instantiated sizes
int ref_grid2d(size_t rows, size_t cols, int topo, double mass, double spacing, double k, double c,
               const double* state, int mode, double dt, size_t n_steps, double* out) {
    GRID_CASE(2, 2, 0) GRID_CASE(3, 3, 0) GRID_CASE(4, 5, 0) GRID_CASE(2, 5, 0)
    GRID_CASE(3, 3, 1) GRID_CASE(4, 4, 1) GRID_CASE(3, 4, 2) GRID_CASE(4, 4, 2)
    return -1;
}

} // extern "C"
