// unified::UnifiedGridSystem -- the reference's 6-state mass-spring grid (physics/unified/: PointMass2D [x, y, vx, vy, theta,
// omega] + Spring2D with attachment points on the rim of each mass, torque coupling and optional repulsion), assembled as
// Grid2DSimulator::createQuadSystem does (wasm/wasm_grid2d.cpp:417-462) on makeGridTopology (constexpr_topology.hpp:368-414),
// with the TypedODESystem-like surface of ValidatedSystem (physics/unified/validated_system.hpp: getStateDimension,
// getInitialState, computeDerivatives).  Runtime sized; the derivative and the RK4 stages run in csrc/ugrid.cu.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>
#include "../../core/solver.hpp"

namespace sopot::unified {

struct GridConfig {
    double mass = 1.0, radius = 0.05, spacing = 0.5;
    double stiffness = 50.0, rest_length = 0.5, damping = 0.5, min_distance = 0.0, repulsion_stiffness = -1.0;
    // attachment angles: horizontal springs on their left / right mass, vertical springs on their upper / lower mass
    double h_attach0 = 0.0, h_attach1 = 3.14159265358979323846, v_attach0 = 3.14159265358979323846 / 2.0, v_attach1 = -3.14159265358979323846 / 2.0;
    // the simulator's "triangle" grid type (createTriangleSystem, wasm/wasm_grid2d.cpp:466-525, on makeTriangleGridTopology):
    // both diagonals of every cell; rest length < 0 means sqrt(2) * spacing; attachment angles of the main diagonal on its
    // upper-left / lower-right mass and of the anti-diagonal on its upper-right / lower-left mass
    bool triangle = false;
    double diag_rest_length = -1.0;
    double d_attach0 = 3.14159265358979323846 / 4.0, d_attach1 = -3.0 * 3.14159265358979323846 / 4.0;
    double a_attach0 = 3.0 * 3.14159265358979323846 / 4.0, a_attach1 = -3.14159265358979323846 / 4.0;
};

// the three schemes of the reference's Grid2DSimulator (wasm/wasm_grid2d.cpp:38-42, :286-413)
enum class Integrator : int { RK4 = SOPOT_B200_INTEGRATOR_RK4, SymplecticEuler = SOPOT_B200_INTEGRATOR_SYMPLECTIC_EULER,
                              VelocityVerlet = SOPOT_B200_INTEGRATOR_VELOCITY_VERLET };

class UnifiedGridSystem {
public:
    UnifiedGridSystem(size_t rows, size_t cols, const GridConfig& c = {})
        : m_p{rows, cols, c.mass, c.radius, c.spacing, c.stiffness, c.rest_length, c.damping, c.min_distance, c.repulsion_stiffness,
              c.h_attach0, c.h_attach1, c.v_attach0, c.v_attach1, 0, 0, c.triangle ? 1 : 0,
              c.diag_rest_length < 0.0 ? std::sqrt(2.0) * c.spacing : c.diag_rest_length, c.d_attach0, c.d_attach1, c.a_attach0, c.a_attach1} {
        if (rows * cols < 2) throw std::invalid_argument("the grid needs at least two masses");
        if (!(c.mass > 0.0) || !(c.radius > 0.0)) throw std::invalid_argument("mass and radius must be positive");
    }
    size_t getStateDimension() const noexcept { return 6 * m_p.rows * m_p.cols; }
    size_t getNodeCount() const noexcept {      // masses + springs (constexpr_topology.hpp:370-373, :437-442)
        return m_p.rows * m_p.cols + m_p.rows * (m_p.cols - 1) + (m_p.rows - 1) * m_p.cols + (m_p.topology ? 2 * (m_p.rows - 1) * (m_p.cols - 1) : 0);
    }
    std::vector<double> getInitialState() const {
        std::vector<double> s(getStateDimension());
        b200::Engine& eng = b200::Engine::shared();
        eng.check(sopot_b200_ugrid_initial_state(eng.ctx(), &m_p, SOPOT_B200_MEM_HOST, s.data()));
        eng.sync();
        return s;
    }
    std::vector<double> computeDerivatives(double /*t*/, std::span<const double> state) const {
        if (state.size() != getStateDimension()) throw std::invalid_argument("state has the wrong dimension");
        b200::Engine& eng = b200::Engine::shared();
        std::vector<double> out(state.size());
        b200::EnsembleOptions o; o.fp_mode = b200::FpMode::Strict;
        const auto opts = b200::make_opts(o, 0);
        eng.check(sopot_b200_ugrid_rhs(eng.ctx(), &m_p, &opts, state.data(), out.data()));
        eng.sync();
        if (auto* tr = b200::detail::active_trace()) {
            tr->calls++;
            tr->returned = out;
            const UnifiedGridSystem snap = *this;
            tr->solve = [snap](b200::Engine& en, const std::vector<double>& y0, size_t n_steps, double, double dt, const b200::EnsembleOptions& opt) {
                b200::EnsembleResult r;
                r.n = 1; r.dim = snap.getStateDimension(); r.n_steps = n_steps; r.store_every = 1;
                r.trajectory.resize(n_steps * r.dim);
                std::vector<double> y = y0;
                for (size_t s = 0; s < n_steps; ++s) {
                    snap.integrate(y, dt, 1, opt.fp_mode, &en);
                    std::copy(y.begin(), y.end(), r.trajectory.begin() + s * r.dim);
                }
                r.state = y;
                r.status.assign(1, 0);
                return r;
            };
            tr->rhs = [snap](b200::Engine&, const std::vector<double>& y) { return snap.computeDerivatives(0.0, y); };
        }
        return out;
    }
    std::vector<double> computeDerivatives(double t, const std::vector<double>& state) const { return computeDerivatives(t, std::span<const double>(state)); }
    // advance `state` in place by n_steps RK4 steps
    void integrate(std::vector<double>& state, double dt, size_t n_steps, b200::FpMode fp = b200::FpMode::Contract, b200::Engine* engine = nullptr) const {
        integrate(state, dt, n_steps, Integrator::RK4, fp, engine);
    }
    // ... or by n_steps of one of the simulator's three integrators (wasm/wasm_grid2d.cpp:38-42 IntegratorType)
    void integrate(std::vector<double>& state, double dt, size_t n_steps, Integrator integrator, b200::FpMode fp = b200::FpMode::Contract,
                   b200::Engine* engine = nullptr) const {
        if (state.size() != getStateDimension()) throw std::invalid_argument("state has the wrong dimension");
        b200::Engine& eng = engine ? *engine : b200::Engine::shared();
        b200::EnsembleOptions o; o.fp_mode = fp;
        const auto opts = b200::make_opts(o, 0);
        eng.check(sopot_b200_ugrid_integrate(eng.ctx(), &m_p, static_cast<int>(integrator), dt, n_steps, &opts, state.data()));
        eng.sync();
    }
private:
    sopot_b200_ugrid_params m_p;
};

}  // namespace sopot::unified
