// 6-DOF rocket through the reference's driver interface (rocket/rocket.hpp:115-303): configure, setup,
// simulateRocket, statistics.  The reference ships no rocket data files and no ctest touches Rocket<T>
// (parity unpinned by reference tests), so the known answer below comes from the
// unmodified reference compiled by oracle/ (Rocket components composed by hand, synthetic engine and mass
// tables of wasm/wasm_rocket.cpp:385-428, elevation 85 deg, azimuth 0, diameter 0.16 m, 30 s at dt = 0.01).
// Needs a B200.
#include <cstdio>
#include <fstream>
#include <sopot/rocket/simulation_result.hpp>
#include "check.hpp"

using namespace sopot;
using namespace sopot::rocket;

static std::vector<double> engine_table(int rows = 36, double thrust = 1500.0, double t_flat = 3.0, double t_burn = 3.5) {
    std::vector<double> tab;
    for (int i = 0; i < rows; ++i) {
        const double t = t_burn * double(i) / double(rows - 1);   // numpy.linspace: start + i*step
        double F = t <= t_flat ? thrust : thrust * (t_burn - t) / (t_burn - t_flat);
        F = std::max(F, 0.0);
        const double p_c = 5.0e6, CF = 1.6, eps = 6.0;
        const double A_throat = F > 0 ? F / (p_c * CF) : 1e-6;
        const double d_throat = 2.0 * std::sqrt(A_throat / 3.14159265359);
        const double row[8] = {t, d_throat, d_throat * std::sqrt(eps), p_c, 3000.0, 1.24, 0.024, 0.96};
        tab.insert(tab.end(), row, row + 8);
    }
    return tab;
}
static std::vector<double> mass_table(int rows = 36, double m_wet = 20.0, double m_dry = 14.75, double t_burn = 3.5) {
    std::vector<double> tab;
    for (int i = 0; i < rows; ++i) {
        const double t = t_burn * double(i) / double(rows - 1);
        tab.push_back(t);
        tab.push_back(m_wet + (m_dry - m_wet) * t / t_burn);
    }
    return tab;
}

int main() {
    Rocket<double> rocket;
    ASSERT_THROWS(rocket.getInitialState(), std::runtime_error);     // "Call setupBeforeSimulation() first"
    rocket.setLauncher(85.0, 0.0);
    rocket.setDiameter(0.16);
    rocket.setEngineTable(engine_table());
    rocket.setMassTable(mass_table());
    ASSERT_NEAR(rocket.getBurnTime(), 3.5, 1e-15);

    auto result = simulateRocket(rocket, 30.0, 0.01);
    ASSERT_TRUE(result.size() == 3001);
    ASSERT_TRUE(rocket.getStateDimension() == 14);
    const auto& y0 = result.states.front();
    ASSERT_TRUE(y0[0] == 0.0 && y0[3] == 0.0 && y0[11] == 0.0);
    ASSERT_NEAR(y0[7] * y0[7] + y0[8] * y0[8] + y0[9] * y0[9] + y0[10] * y0[10], 1.0, 1e-14);

    // final state of the reference run; libm-level differences only (<= 1e-9 relative on the final state)
    const double ref_final[14] = {30.00000000000189, 4.751621710833997e-13, 473.78546748858656, 1668.16272022101,
                                  1.3195469879623669e-14, 13.851338727967878, -85.25061288170275,
                                  0.47771441710826085, -0.4777144171082609, 0.5213338044735968, 0.5213338044735969,
                                  0.0, 0.0, 0.0};
    const auto& yN = result.states.back();
    double scale = 0.0;
    for (double v : ref_final) scale = std::max(scale, std::abs(v));
    for (int j = 0; j < 14; ++j) ASSERT_NEAR(yN[j], ref_final[j], 1e-9 * scale);
    ASSERT_TRUE(result.times.back() == yN[0] || std::abs(result.times.back() - yN[0]) < 1e-9);

    // apogee from the stored trajectory (SimulationResult::computeStatistics semantics: max of state[3])
    double apogee = 0.0, apogee_t = 0.0;
    for (size_t i = 0; i < result.size(); ++i)
        if (result.states[i][3] > apogee) { apogee = result.states[i][3]; apogee_t = result.times[i]; }
    ASSERT_NEAR(apogee, 2071.874676142213, 1e-9 * 2071.874676142213);
    ASSERT_NEAR(apogee_t, 20.79, 1e-9);
    ASSERT_NEAR(rocket.getAltitude(yN), yN[3], 0.0);

    // the same flight inside a Monte-Carlo ensemble: fused running statistics, no trajectory stored
    const size_t n = 64;
    std::vector<double> elev(n), az(n);
    for (size_t i = 0; i < n; ++i) { elev[i] = 85.0 + 2.0 * std::sin(double(i)); az[i] = 360.0 * double(i) / double(n); }
    elev[0] = 85.0; az[0] = 0.0;
    RocketEnsemble ens{n, elev, az, 0.16, 9.80665, rocket.tables()};
    b200::EnsembleOptions opt;
    opt.fp_mode = b200::FpMode::Strict;
    auto er = createRK4Solver().solveEnsemble(ens, 0.0, 30.0, 0.01, opt);
    ASSERT_TRUE(er.stat(Apogee, 0) == apogee);
    for (int j = 0; j < 14; ++j) ASSERT_TRUE(er.final(j, 0) == yN[j]);
    auto disp = RocketEnsemble::dispersion(b200::Engine::shared(), er);
    ASSERT_TRUE(disp[Apogee].count == double(n) && disp[Apogee].min <= apogee && disp[Apogee].max >= apogee);
    ASSERT_TRUE(disp[Apogee].stddev > 0.0 && disp[Apogee].nonfinite == 0.0);

    // SimulationResult (rocket/simulation_result.hpp): statistics of the stored trajectory == the kernel's running ones,
    // time-indexed queries interpolate between stored steps
    {
        auto sim = simulate(rocket, 30.0, 0.01);
        ASSERT_TRUE(sim.size() == 3001 && sim.startTime() == 0.0);
        ASSERT_TRUE(sim.apogee() == er.stat(Apogee, 0) && sim.apogeeTime() == er.stat(ApogeeTime, 0));
        ASSERT_TRUE(sim.maxSpeed() == er.stat(MaxSpeed, 0) && sim.maxSpeedTime() == er.stat(MaxSpeedTime, 0));
        ASSERT_TRUE(sim.burnoutAltitude() == er.stat(BurnoutAltitude, 0) && sim.burnoutSpeed() == er.stat(BurnoutSpeed, 0));
        const auto [i, alpha] = sim.findIndex(12.3456);
        ASSERT_TRUE(i == 1234 && alpha > 0.55 && alpha < 0.57);
        const double a0 = sim.stateAt(1234)[3], a1 = sim.stateAt(1235)[3];
        ASSERT_NEAR(sim.altitude(12.3456), a0 + alpha * (a1 - a0), 1e-9);
        ASSERT_TRUE(sim.altitude(-1.0) == 0.0 && sim.altitude(99.0) == sim.stateAt(3000)[3]);
        ASSERT_TRUE(sim.altitudeProfile().size() == 3001 && sim.speed(3.0) > 100.0);
        // query<Tag>(t) (rocket/simulation_result.hpp:111-115) for any tag of rocket_tags.hpp: slices of the interpolated state ...
        const auto st = sim.interpolateState(2.0);
        const auto pos = sim.query<rocket::kinematics::PositionENU>(2.0);
        ASSERT_TRUE(pos.x == st[1] && pos.y == st[2] && pos.z == st[3] && sim.position(2.0)[2] == st[3]);
        ASSERT_TRUE(sim.query<rocket::propulsion::Time>(2.0) == st[0] && sim.attitude(2.0).q4 == st[10]);
        ASSERT_NEAR(sim.velocity(2.0).norm(), sim.speed(2.0), 0.0);
        // ... and intermediates of the derivative from the state-function kernel: mass table (20 -> 14.75 kg over 3.5 s), thrust on
        // the plateau, the standard atmosphere at that altitude, Newton's second law against computeDerivatives
        ASSERT_NEAR(sim.mass(2.0), 20.0 + (14.75 - 20.0) * 2.0 / 3.5, 1e-12 * 20.0);
        ASSERT_NEAR(sim.query<rocket::dynamics::Mass>(10.0), 14.75, 0.0);
        const auto thrust = sim.query<rocket::propulsion::ThrustForceBody>(2.0);
        ASSERT_TRUE(thrust.x > 1400.0 && thrust.x < 1600.0 && thrust.y == 0.0 && thrust.z == 0.0);
        ASSERT_TRUE(sim.query<rocket::propulsion::ThrustForceBody>(10.0).x == 0.0 && sim.query<rocket::propulsion::MassFlowRate>(10.0) == 0.0);
        ASSERT_TRUE(sim.query<rocket::propulsion::MassFlowRate>(2.0) > 0.5);
        const double T_alt = sim.query<rocket::environment::AtmosphericTemperature>(2.0);
        ASSERT_NEAR(T_alt, 288.15 - 0.0065 * st[3], 1e-12 * 288.15);
        ASSERT_NEAR(sim.query<rocket::environment::AtmosphericDensity>(2.0),
                    sim.query<rocket::environment::AtmosphericPressure>(2.0) / ((8.3144598 / 0.0289644) * T_alt), 1e-15 * 1.3);
        ASSERT_NEAR(sim.query<rocket::dynamics::GravityAcceleration>(2.0), 9.80665, 0.0);
        const auto F = sim.query<rocket::dynamics::TotalForceENU>(2.0);
        const auto dv = rocket.computeDerivatives(0.0, st);
        for (int k = 0; k < 3; ++k) ASSERT_NEAR(F[k] / sim.mass(2.0), dv[4 + k], 1e-13 * std::abs(dv[6]));
        const auto Fa = sim.query<rocket::aero::AeroForceENU>(2.0);
        ASSERT_TRUE(Fa.norm() > 1.0 && sim.query<rocket::aero::AeroMomentBody>(2.0).norm() == sim.query<rocket::dynamics::TotalTorqueBody>(2.0).norm());
        ASSERT_NEAR(rocket.getMass(st), sim.mass(2.0), 0.0);
    }

    // CSV path of the reference (rocket/rocket.hpp:124-133): numeric rows, header cells skipped
    {
        std::ofstream e("/tmp/sopot_b200_sim_engine.csv");
        e << "time,throat_d,exit_d,p_comb,T_comb,gamma,mol_mass,efficiency\n";
        e.precision(17);
        const auto tab = engine_table();
        for (size_t r = 0; r < tab.size() / 8; ++r) { for (int c = 0; c < 8; ++c) e << (c ? "," : "") << tab[r * 8 + c]; e << "\n"; }
    }
    {   // io::CsvParser (io/csv_parser.hpp): comments, header cells, blank lines, both table shapes
        const std::string text = "# engine\ntime, thrust, note\n0.0, 10.5, a\n\n 1.0 ,\t20.5\n2.0,30.5,7\n";
        const auto t = io::CsvParser::parseString(text);
        ASSERT_TRUE(t.rows() == 3 && t.cols() == 2 && t(1, 1) == 20.5 && t.row(2).size() == 3);
        ASSERT_TRUE((t.column(0) == std::vector<double>{0.0, 1.0, 2.0}) && (t.column(2) == std::vector<double>{7.0}));
        const auto t2 = io::CsvParser::parseString("x, 0.5, 1.0, 2.0\n0, 1, 2, 3\n10, 4, 5, 6\n", ',', true);
        ASSERT_TRUE(t2.rows() == 3 && (t2.row(0) == std::vector<double>{0.5, 1.0, 2.0}) && (t2.row(2) == std::vector<double>{4.0, 5.0, 6.0}));
        ASSERT_THROWS(io::CsvParser::parseFile("/nonexistent/table.csv"), std::runtime_error);
        ASSERT_THROWS(t(5, 0), std::out_of_range);
    }
    Rocket<double> from_csv;
    from_csv.loadEngineData("/tmp/sopot_b200_");
    ASSERT_TRUE(from_csv.tables().engine == engine_table());
    ASSERT_THROWS(from_csv.loadEngineData("/nonexistent/dir/"), std::runtime_error);
    // coefficient tables through the reference's loaders (rocket/rocket.hpp:135-147): lift and a drag rise change the flight
    {
        auto write = [](const char* path, const std::vector<double>& cols, const std::vector<double>& rows, auto f) {
            std::ofstream o(path);
            o.precision(17);
            o << "0";                                            // placeholder cell in front of the column headers
            for (double c : cols) o << "," << c;
            o << "\n";
            for (double r : rows) { o << r; for (double c : cols) o << "," << f(r, c); o << "\n"; }
        };
        const std::vector<double> mach{0.0, 0.3, 0.6, 0.9, 1.2}, ang{0.0, 2.0, 5.0, 10.0, 20.0};
        write("/tmp/sopot_b200_cd_on.csv", ang, mach, [](double m, double a) { return 0.25 + 0.1 * m + 0.004 * a; });
        write("/tmp/sopot_b200_cl_on.csv", ang, mach, [](double m, double a) { return 0.02 * a * (1 + 0.2 * m); });
        Rocket<double> tabled;
        tabled.setLauncher(85.0, 0.0);
        tabled.setDiameter(0.16);
        tabled.setEngineTable(engine_table());
        tabled.setMassTable(mass_table());
        tabled.aero().loadCdPowerOn("/tmp/sopot_b200_cd_on.csv");
        tabled.aero().loadClPowerOn("/tmp/sopot_b200_cl_on.csv");
        ASSERT_TRUE(tabled.aero().cd_power_on.rows.size() == 5 && tabled.aero().cd_power_on.cols.size() == 5);
        auto rt = simulateRocket(tabled, 30.0, 0.01);
        double apogee_t = 0.0;
        for (const auto& st : rt.states) apogee_t = std::max(apogee_t, st[3]);
        ASSERT_TRUE(std::isfinite(apogee_t) && std::abs(apogee_t - apogee) > 1.0);     // Cd(M, alpha) != 0.3: a different flight
        ASSERT_THROWS(tabled.aero().loadCsPowerOn("/nonexistent.csv"), std::runtime_error);
    }
    std::puts("rocket_test: OK");
    return 0;
}
