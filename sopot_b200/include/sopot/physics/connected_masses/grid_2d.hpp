// connected_masses::makeGrid2DSystem<T, Rows, Cols, IncludeDiagonals>(mass, spacing, stiffness, damping) and
// makeTriangularGridSystem -- the factories of physics/connected_masses/connectivity_matrix_2d.hpp:156-266 --
// plus a runtime-sized twin (makeGrid2DSystem(rows, cols, ...)) because the reference's one-type-per-mass
// composition cannot be instantiated beyond ~6x6.
// The returned Grid2DSystem has the TypedODESystem surface the reference's grid tests use (getStateDimension,
// getInitialState, computeDerivatives; state AoS [row][col][x, y, vx, vy], masses first, row-major:
// grid_2d.hpp:97-99).  Springs, force aggregation in edge-list order and the point-mass derivative run in
// csrc/grid2d.cu (K3).
#pragma once
#include <stdexcept>
#include <vector>
#include "../../core/solver.hpp"

namespace sopot::connected_masses {

enum class GridTopology : int { Quad = 0, QuadWithDiagonals = 1, Triangular = 2 };

class Grid2DSystem {
public:
    Grid2DSystem(size_t rows, size_t cols, GridTopology topo, double mass, double spacing, double stiffness, double damping)
        : m_p{rows, cols, 0, rows, static_cast<int>(topo), mass, spacing, stiffness, damping} {
        if (rows < 2) throw std::invalid_argument("Grid must have at least 2 rows");
        if (cols < 2) throw std::invalid_argument("Grid must have at least 2 columns");
        if (!(mass > 0.0)) throw std::invalid_argument("Mass must be positive");
    }
    size_t getStateDimension() const noexcept { return 4 * m_p.rows * m_p.cols; }
    size_t rows() const noexcept { return m_p.rows; }
    size_t cols() const noexcept { return m_p.cols; }
    static constexpr size_t stateIndex(size_t mass_index, size_t component) { return mass_index * 4 + component; }
    const sopot_b200_grid2d_params& params() const noexcept { return m_p; }

    // rest grid: x = col*spacing, y = row*spacing, v = 0 (connectivity_matrix_2d.hpp:171-180)
    std::vector<double> getInitialState() const {
        std::vector<double> s(getStateDimension());
        b200::Engine& eng = b200::Engine::shared();
        eng.check(sopot_b200_grid2d_initial_state(eng.ctx(), &m_p, SOPOT_B200_MEM_HOST, s.data()));
        eng.sync();
        return s;
    }
    std::vector<double> computeDerivatives(double /*t*/, const std::vector<double>& state) const {
        if (state.size() != getStateDimension()) throw std::invalid_argument("state has the wrong dimension");
        b200::Engine& eng = b200::Engine::shared();
        std::vector<double> out(state.size());
        const auto opts = b200::make_opts(strictOpts(), 0);
        eng.check(sopot_b200_grid2d_rhs(eng.ctx(), &m_p, &opts, state.data(), out.data()));
        eng.sync();
        if (auto* tr = b200::detail::active_trace()) {
            tr->calls++;
            tr->returned = out;
            const Grid2DSystem snap = *this;
            tr->solve = [snap](b200::Engine& en, const std::vector<double>& y0, size_t n_steps, double, double dt,
                               const b200::EnsembleOptions& o) { return snap.trajectory(en, y0, n_steps, dt, o); };
            tr->rhs = [snap](b200::Engine&, const std::vector<double>& y) { return snap.computeDerivatives(0.0, y); };
        }
        return out;
    }
    // advance `state` in place by n_steps RK4 steps (the bulk entry point; no trajectory is kept)
    void integrate(std::vector<double>& state, double dt, size_t n_steps, b200::FpMode fp = b200::FpMode::Contract,
                   b200::Engine* engine = nullptr) const {
        if (state.size() != getStateDimension()) throw std::invalid_argument("state has the wrong dimension");
        b200::Engine& eng = engine ? *engine : b200::Engine::shared();
        b200::EnsembleOptions o; o.fp_mode = fp;
        const auto opts = b200::make_opts(o, 0);
        eng.check(sopot_b200_grid2d_rk4(eng.ctx(), &m_p, dt, n_steps, &opts, state.data()));
        eng.sync();
    }
private:
    static b200::EnsembleOptions strictOpts() { b200::EnsembleOptions o; o.fp_mode = b200::FpMode::Strict; return o; }
    // RK4Solver::solve keeps every step: one launch sequence per stored step, "instance" = the whole grid
    b200::EnsembleResult trajectory(b200::Engine& eng, const std::vector<double>& y0, size_t n_steps, double dt,
                                    const b200::EnsembleOptions& o) const {
        b200::EnsembleResult r;
        r.n = 1; r.dim = getStateDimension(); r.n_steps = n_steps; r.store_every = 1;
        r.trajectory.resize(n_steps * r.dim);
        std::vector<double> y = y0;
        for (size_t s = 0; s < n_steps; ++s) {
            integrate(y, dt, 1, o.fp_mode, &eng);
            std::copy(y.begin(), y.end(), r.trajectory.begin() + s * r.dim);
        }
        r.state = y;
        r.status.assign(1, 0);
        return r;
    }
    sopot_b200_grid2d_params m_p;
};

template <typename T, size_t Rows, size_t Cols, bool IncludeDiagonals = false>
Grid2DSystem makeGrid2DSystem(double mass, double spacing, double stiffness, double damping = 0.0) {
    static_assert(std::is_same_v<T, double>, "the grid kernels are FP64");
    static_assert(Rows >= 2, "Grid must have at least 2 rows");
    static_assert(Cols >= 2, "Grid must have at least 2 columns");
    return Grid2DSystem(Rows, Cols, IncludeDiagonals ? GridTopology::QuadWithDiagonals : GridTopology::Quad, mass, spacing, stiffness, damping);
}
template <typename T, size_t Rows, size_t Cols>
Grid2DSystem makeTriangularGridSystem(double mass, double spacing, double stiffness, double damping = 0.0) {
    static_assert(std::is_same_v<T, double>, "the grid kernels are FP64");
    return Grid2DSystem(Rows, Cols, GridTopology::Triangular, mass, spacing, stiffness, damping);
}
inline Grid2DSystem makeGrid2DSystem(size_t rows, size_t cols, double mass, double spacing, double stiffness, double damping = 0.0,
                                     GridTopology topo = GridTopology::Quad) {
    return Grid2DSystem(rows, cols, topo, mass, spacing, stiffness, damping);
}

}  // namespace sopot::connected_masses
