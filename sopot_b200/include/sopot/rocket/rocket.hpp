// rocket::Rocket<double> -- 6-DOF rocket with the reference's driver interface (rocket/rocket.hpp:54-303):
// setLauncher / setDiameter / load*Data / setupBeforeSimulation / getInitialState / computeDerivatives /
// getStateDimension / convenience accessors, and simulateRocket(rocket, t_end, dt).
// State layout (rocket/rocket.hpp:34-39): [t; pos ENU(3); vel ENU(3); q1..q4 scalar-last; omega body(3)].
// The 11-component right-hand side is RocketSys in csrc/rocket.cu; this
// header only carries configuration to sopot_b200_rocket_rk4 / _rhs.  Tables can be given as arrays
// (setEngineTable ...) or read from the reference's CSV files (numeric rows, '#' comments, header cells
// skipped: io/csv_parser.hpp:44-83).  Aerodynamic coefficient tables (loadAeroData / loadDampingData / aero()) feed the
// bilinear lookups of the kernel; without them the reference defaults apply (Cd 0.3, Cl = Cs = 0, cmq = -450).
#pragma once
#include <array>
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../core/solver.hpp"
#include "../core/typed_component.hpp"
#include "../io/csv_parser.hpp"
#include "rocket_tags.hpp"

namespace sopot::rocket {

namespace detail {
inline std::vector<std::vector<double>> parse_csv(const std::string& filename, char delimiter = ',') {
    return io::CsvParser::parseFile(filename, delimiter).data;
}
inline std::vector<double> flatten(const std::vector<std::vector<double>>& rows, size_t cols, const char* what) {
    std::vector<double> out;
    out.reserve(rows.size() * cols);
    for (const auto& r : rows) {
        if (r.size() < cols) throw std::invalid_argument(std::string(what) + ": every row needs " + std::to_string(cols) + " columns");
        out.insert(out.end(), r.begin(), r.begin() + cols);
    }
    return out;
}
}  // namespace detail

// io::BilinearInterpolator's table (io/interpolation.hpp:77-133): first CSV line = column headers behind a placeholder
// cell, then one line per row: row header, data...
struct CoefficientTable {
    std::vector<double> rows, cols, data;      // data[rows][cols]
    bool empty() const { return data.empty(); }
    void setup(std::vector<double> r, std::vector<double> c, std::vector<double> d) {
        if (d.size() != r.size() * c.size() || r.size() < 2 || c.size() < 2)
            throw std::invalid_argument("coefficient table needs >= 2 rows and columns and rows*cols values");
        rows = std::move(r); cols = std::move(c); data = std::move(d);
    }
    void setupFromTable(const std::vector<std::vector<double>>& table) {
        if (table.empty() || table[0].empty()) throw std::runtime_error("BilinearInterpolator: empty table");
        std::vector<double> r, c(table[0].begin() + 1, table[0].end()), d;
        for (size_t i = 1; i < table.size(); ++i) {
            if (table[i].empty()) continue;
            r.push_back(table[i][0]);
            if (table[i].size() != c.size() + 1) throw std::invalid_argument("coefficient table: ragged row");
            d.insert(d.end(), table[i].begin() + 1, table[i].end());
        }
        setup(std::move(r), std::move(c), std::move(d));
    }
    void loadFromFile(const std::string& filename) { setupFromTable(detail::parse_csv(filename)); }
    sopot_b200_table2d view() const {
        return empty() ? sopot_b200_table2d{0, 0, nullptr, nullptr, nullptr}
                       : sopot_b200_table2d{rows.size(), cols.size(), rows.data(), cols.data(), data.data()};
    }
};

// configuration surface of AxisymmetricAerodynamics (rocket/axisymmetric_aero.hpp:52-111): coefficient tables and the
// power-on switch; the force / moment model itself is RocketSysT<true> in csrc/rocket.cu
class AerodynamicsConfig {
public:
    void setEngineOn(bool on) { m_engine_on = on; }
    void loadCdPowerOff(const std::string& f) { cd_power_off.loadFromFile(f); }
    void loadCdPowerOn(const std::string& f) { cd_power_on.loadFromFile(f); }
    void loadClPowerOff(const std::string& f) { cl_power_off.loadFromFile(f); }
    void loadClPowerOn(const std::string& f) { cl_power_on.loadFromFile(f); }
    void loadCsPowerOff(const std::string& f) { cs_power_off.loadFromFile(f); }
    void loadCsPowerOn(const std::string& f) { cs_power_on.loadFromFile(f); }
    void loadDampingX(const std::string& f) { cmq_x.loadFromFile(f); }       // loaded and, as in the reference, never used
    void loadDampingYZ(const std::string& f) { cmq_yz.loadFromFile(f); }
    bool any() const {
        return !(cd_power_on.empty() && cd_power_off.empty() && cl_power_on.empty() && cl_power_off.empty() &&
                 cs_power_on.empty() && cs_power_off.empty() && cmq_yz.empty());
    }
    sopot_b200_aero_tables view() const {
        return {cd_power_on.view(), cd_power_off.view(), cl_power_on.view(), cl_power_off.view(), cs_power_on.view(),
                cs_power_off.view(), cmq_yz.view(), m_engine_on ? 1 : 0};
    }
    CoefficientTable cd_power_on, cd_power_off, cl_power_on, cl_power_off, cs_power_on, cs_power_off, cmq_x, cmq_yz;
private:
    bool m_engine_on = true;
};

// tables shared by every trajectory of a call (row-major, see include/sopot_b200.h)
struct RocketTables {
    std::vector<double> engine;   // [rows][8] time, throat_d, exit_d, p_comb, T_comb, gamma, mol_mass, efficiency
    std::vector<double> mass, ix, iyz;   // [rows][2] time, value
    AerodynamicsConfig aero;
    // `view` must outlive the call that uses p (it is what p.aero points to)
    void fill(sopot_b200_rocket_params& p, sopot_b200_aero_tables& view) const {
        p.engine_rows = engine.size() / 8; p.engine = engine.empty() ? nullptr : engine.data();
        p.mass_rows = mass.size() / 2;     p.mass = mass.empty() ? nullptr : mass.data();
        p.ix_rows = ix.size() / 2;         p.ix = ix.empty() ? nullptr : ix.data();
        p.iyz_rows = iyz.size() / 2;       p.iyz = iyz.empty() ? nullptr : iyz.data();
        p.aero = nullptr;
        if (aero.any()) { view = aero.view(); p.aero = &view; }
    }
};

// per-trajectory dispersion statistics rows of stats_out (SimulationResult::computeStatistics,
// rocket/simulation_result.hpp:151-189, plus the final downrange position)
enum RocketStat : size_t { Apogee = 0, ApogeeTime, MaxSpeed, MaxSpeedTime, BurnoutAltitude, BurnoutSpeed, FinalEast, FinalNorth, RocketStatCount };

struct RocketEnsembleResult : b200::EnsembleResult {
    std::vector<double> stats;    // [8][n]
    double stat(RocketStat s, size_t i) const { return stats[size_t(s) * n + i]; }
};

// Monte-Carlo ensemble: launcher angles, reference diameter and gravity per trajectory, shared tables
struct RocketEnsemble {
    size_t n;
    b200::Param elevation_deg, azimuth_deg, diameter, gravity;
    RocketTables tables;
    sopot_b200_rocket_params params(uint32_t& bc, sopot_b200_aero_tables& aero_view) const {
        const b200::Param* f[4] = {&elevation_deg, &azimuth_deg, &diameter, &gravity};
        bc = 0;
        for (int j = 0; j < 4; ++j) bc |= (f[j]->broadcast(n) ? 1u : 0u) << j;
        sopot_b200_rocket_params p{};
        p.elevation_deg = f[0]->data(); p.azimuth_deg = f[1]->data(); p.diameter = f[2]->data(); p.gravity = f[3]->data();
        tables.fill(p, aero_view);
        return p;
    }
    RocketEnsembleResult integrate(b200::Engine& eng, size_t n_steps, double /*t0: the rocket carries its own clock, state[0]*/,
                                   double dt, const b200::EnsembleOptions& o) const {
        uint32_t bc;
        sopot_b200_aero_tables aero_view;
        const auto p = params(bc, aero_view);
        const auto opts = b200::make_opts(o, bc);
        RocketEnsembleResult r;
        r.n = n; r.dim = 14; r.n_steps = n_steps; r.store_every = o.store_every;
        r.state.resize(14 * n);
        r.trajectory.resize(r.snapshots() * 14 * n);
        r.stats.resize(8 * n);
        if (o.want_status) r.status.resize(n);
        eng.check(sopot_b200_rocket_rk4(eng.ctx(), &p, n, dt, n_steps, &opts, r.state.data(),
                                        r.trajectory.empty() ? nullptr : r.trajectory.data(), r.stats.data(),
                                        o.want_status ? r.status.data() : nullptr));
        eng.sync();
        return r;
    }
    // dispersion statistics over all trajectories (and all ranks of the engine's NCCL communicator)
    static std::array<sopot_b200_dispersion, RocketStatCount> dispersion(b200::Engine& eng, const RocketEnsembleResult& r) {
        std::array<sopot_b200_dispersion, RocketStatCount> out{};
        eng.check(sopot_b200_dispersion_stats(eng.ctx(), r.stats.data(), RocketStatCount, r.n, SOPOT_B200_MEM_HOST, out.data()));
        return out;
    }
};

template <Scalar T = double>
class Rocket {
    static_assert(std::is_same_v<T, double>, "the compiled rocket kernels are FP64 (Rocket<double>)");
public:
    using scalar_type = T;
    Rocket() = default;

    void setLauncher(T elevation_deg, T azimuth_deg) { m_elevation_deg = elevation_deg; m_azimuth_deg = azimuth_deg; m_ready = false; }
    void setDiameter(double diameter) { m_diameter = diameter; m_ready = false; }
    void setGravity(double g) { m_gravity = g; m_ready = false; }            // Gravity<T>(g), rocket/gravity.hpp
    void loadMassData(const std::string& path) {                             // rocket/rocket.hpp:124-129
        m_tables.mass = detail::flatten(detail::parse_csv(path + "sim_mass.csv"), 2, "sim_mass.csv");
        m_tables.ix = detail::flatten(detail::parse_csv(path + "sim_ix.csv"), 2, "sim_ix.csv");
        m_tables.iyz = detail::flatten(detail::parse_csv(path + "sim_iyz.csv"), 2, "sim_iyz.csv");
        m_ready = false;
    }
    void loadEngineData(const std::string& path) {                           // :131-133
        m_tables.engine = detail::flatten(detail::parse_csv(path + "sim_engine.csv"), 8, "sim_engine.csv");
        m_ready = false;
    }
    void loadAeroData(const std::string& path) {                             // :135-142
        m_tables.aero.loadCdPowerOff(path + "cd_power_off_rocket.csv"); m_tables.aero.loadCdPowerOn(path + "cd_power_on_rocket.csv");
        m_tables.aero.loadClPowerOff(path + "cl_power_off_rocket.csv"); m_tables.aero.loadClPowerOn(path + "cl_power_on_rocket.csv");
        m_tables.aero.loadCsPowerOff(path + "cs_power_off_rocket.csv"); m_tables.aero.loadCsPowerOn(path + "cs_power_on_rocket.csv");
        m_ready = false;
    }
    void loadDampingData(const std::string& path) {                          // :144-147
        m_tables.aero.loadDampingX(path + "cmq_x_power_on_rocket.csv"); m_tables.aero.loadDampingYZ(path + "cmq_yz_power_on_rocket.csv");
        m_ready = false;
    }
    AerodynamicsConfig& aero() { m_ready = false; return m_tables.aero; }     // :152-153
    const AerodynamicsConfig& aero() const { return m_tables.aero; }
    void setEngineTable(std::vector<double> rows_by_8) { m_tables.engine = std::move(rows_by_8); m_ready = false; }
    void setMassTable(std::vector<double> rows_by_2) { m_tables.mass = std::move(rows_by_2); m_ready = false; }
    void setInertiaTables(std::vector<double> ix, std::vector<double> iyz) { m_tables.ix = std::move(ix); m_tables.iyz = std::move(iyz); m_ready = false; }
    double getBurnTime() const { return m_tables.engine.empty() ? 0.0 : m_tables.engine[m_tables.engine.size() - 8]; }
    const RocketTables& tables() const { return m_tables; }

    void setupBeforeSimulation() {                                           // :159-191
        // the initial attitude comes from the launcher angles (Quaternion::from_launcher_angles,
        // rocket/quaternion.hpp:164-190), evaluated by the kernel's own load(): a zero-step launch
        m_initial = ensemble().integrate(b200::Engine::shared(), 0, 0.0, 0.01, strictOpts()).state;
        m_ready = true;
    }
    std::vector<T> getInitialState() const { need(); return m_initial; }
    size_t getStateDimension() const { need(); return 14; }
    std::vector<T> computeDerivatives(T /*t: components read state[0]*/, const std::vector<T>& state) const {
        need();
        if (state.size() != 14) throw std::invalid_argument("state has the wrong dimension");
        b200::Engine& eng = b200::Engine::shared();
        const auto e = ensemble();
        uint32_t bc;
        sopot_b200_aero_tables aero_view;
        const auto p = e.params(bc, aero_view);
        const auto opts = b200::make_opts(strictOpts(), bc);
        std::vector<double> out(14);
        eng.check(sopot_b200_rocket_rhs(eng.ctx(), &p, 1, &opts, state.data(), out.data()));
        eng.sync();
        if (auto* tr = b200::detail::active_trace()) {
            tr->calls++;
            tr->returned = out;
            const Rocket snap = *this;
            tr->solve = [snap](b200::Engine& en, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                               const b200::EnsembleOptions& o) -> b200::EnsembleResult {
                if (y0 != snap.m_initial)
                    throw std::invalid_argument("sopot-b200: the rocket kernel integrates from the launcher state (getInitialState())");
                return snap.ensemble().integrate(en, n_steps, t0, dt, o);
            };
            tr->rhs = [snap](b200::Engine&, const std::vector<double>& y) { return snap.computeDerivatives(0.0, y); };
        }
        return out;
    }
    // Any state function of the rocket at `state` (:222-228).  Kinematic tags and the time are slices of the state vector;
    // every other tag is an intermediate of the derivative and comes from one strict-mode launch of the state-function
    // kernel (sopot_b200_rocket_state_functions), so a query sees exactly the value the integrator used.
    template <RocketTag Tag>
    auto queryStateFunction(const std::vector<T>& state) const {
        need();
        if (state.size() != 14) throw std::invalid_argument("state has the wrong dimension");
        if constexpr (Tag::source == TagSource::Slice) return rocketTagValue<Tag>(state.data());
        else return rocketTagValue<Tag>(stateFunctions(state).data());
    }
    // the whole record at once (28 values, rows as in include/sopot_b200.h): one launch for several queries
    std::array<double, SOPOT_B200_ROCKET_PROBE_VALUES> stateFunctions(const std::vector<T>& state) const {
        need();
        b200::Engine& eng = b200::Engine::shared();
        const auto e = ensemble();
        uint32_t bc;
        sopot_b200_aero_tables aero_view;
        const auto p = e.params(bc, aero_view);
        const auto opts = b200::make_opts(strictOpts(), bc);
        std::array<double, SOPOT_B200_ROCKET_PROBE_VALUES> out{};
        eng.check(sopot_b200_rocket_state_functions(eng.ctx(), &p, 1, &opts, state.data(), out.data()));
        eng.sync();
        return out;
    }
    // convenience accessors (:229-262)
    T getTime(const std::vector<T>& s) const { return queryStateFunction<propulsion::Time>(s); }
    Vector3<T> getPosition(const std::vector<T>& s) const { return queryStateFunction<kinematics::PositionENU>(s); }
    Vector3<T> getVelocity(const std::vector<T>& s) const { return queryStateFunction<kinematics::VelocityENU>(s); }
    Quaternion<T> getQuaternion(const std::vector<T>& s) const { return queryStateFunction<kinematics::AttitudeQuaternion>(s); }
    T getAltitude(const std::vector<T>& s) const { return queryStateFunction<kinematics::Altitude>(s); }
    T getMass(const std::vector<T>& s) const { return queryStateFunction<dynamics::Mass>(s); }
    T getSpeed(const std::vector<T>& s) const { return getVelocity(s).norm(); }

    RocketEnsemble ensemble() const { return {1, m_elevation_deg, m_azimuth_deg, m_diameter, m_gravity, m_tables}; }

private:
    static b200::EnsembleOptions strictOpts() { b200::EnsembleOptions o; o.fp_mode = b200::FpMode::Strict; return o; }
    void need() const { if (!m_ready) throw std::runtime_error("Call setupBeforeSimulation() first"); }
    RocketTables m_tables;
    T m_elevation_deg{90}, m_azimuth_deg{0};
    double m_diameter = 0.45;       // AxisymmetricAerodynamics default (rocket/axisymmetric_aero.hpp:53)
    double m_gravity = 9.80665;     // Gravity<T> default
    std::vector<double> m_initial;
    bool m_ready = false;
};

// rocket/rocket.hpp:268-303: setup, wrap computeDerivatives in a lambda, RK4Solver::solve from t = 0
template <Scalar T = double>
SolutionResult simulateRocket(Rocket<T>& rocket, double t_end, double dt = 0.01) {
    rocket.setupBeforeSimulation();
    auto initial_state = rocket.getInitialState();
    const size_t state_dim = rocket.getStateDimension();
    auto derivs_func = [&rocket](double t, StateView sv) -> StateDerivative {
        return rocket.computeDerivatives(T(t), std::vector<T>(sv.begin(), sv.end()));
    };
    return createRK4Solver().solve(derivs_func, state_dim, 0.0, t_end, dt, initial_state);
}

}  // namespace sopot::rocket
