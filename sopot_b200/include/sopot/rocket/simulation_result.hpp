// rocket::SimulationResult<RocketType> and rocket::simulate(rocket, t_end, dt) -- interface of
// rocket/simulation_result.hpp:20-245 and rocket/rocket.hpp:308-318: time-indexed access to a stored trajectory
// (findIndex / interpolateState / query<Tag>(t) / position(t), altitude(t), mass(t), ...), flight statistics and profiles.  Host-side
// post-processing of what the device integrated; the statistics scan is the reference's (:151-189) and equals, bit for
// bit, what the ensemble kernel accumulates on the fly (RocketEnsembleResult::stat), which is how Monte-Carlo runs get
// their apogee / burnout numbers without storing trajectories.
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include "rocket.hpp"

namespace sopot::rocket {

template <typename RocketType>
class SimulationResult {
public:
    using T = typename RocketType::scalar_type;
    SimulationResult() : m_rocket(nullptr) {}
    SimulationResult(const RocketType& rocket, SolutionResult&& solution) : m_rocket(&rocket), m_solution(std::move(solution)) {}

    size_t size() const { return m_solution.size(); }
    bool empty() const { return m_solution.empty(); }
    double startTime() const { return m_solution.times.empty() ? 0 : m_solution.times.front(); }
    double endTime() const { return m_solution.times.empty() ? 0 : m_solution.times.back(); }
    double simulationTimeMs() const { return m_solution.total_time; }
    const std::vector<double>& times() const { return m_solution.times; }
    const std::vector<StateVector>& states() const { return m_solution.states; }
    const StateVector& stateAt(size_t i) const { return m_solution.states[i]; }
    double timeAt(size_t i) const { return m_solution.times[i]; }

    std::pair<size_t, double> findIndex(double t) const {                       // :58-75
        if (t <= m_solution.times.front()) return {0, 0.0};
        if (t >= m_solution.times.back()) return {m_solution.size() - 1, 0.0};
        size_t i = std::distance(m_solution.times.begin(), std::lower_bound(m_solution.times.begin(), m_solution.times.end(), t));
        if (i > 0) --i;
        const double t0 = m_solution.times[i], t1 = m_solution.times[i + 1];
        return {i, (t - t0) / (t1 - t0)};
    }
    StateVector interpolateState(double t) const {                             // :78-95
        const auto [i, alpha] = findIndex(t);
        if (alpha == 0.0 || i >= m_solution.size() - 1) return m_solution.states[i];
        const auto &s0 = m_solution.states[i], &s1 = m_solution.states[i + 1];
        StateVector r(s0.size());
        for (size_t j = 0; j < r.size(); ++j) r[j] = s0[j] + alpha * (s1[j] - s0[j]);
        return r;
    }
    // any state function at time t (:111-115): interpolate the stored trajectory, then ask the rocket
    template <StateTagConcept Tag>
    auto query(double t) const { return m_rocket->template queryStateFunction<Tag>(interpolateState(t)); }
    Vector3<T> position(double t) const { return query<kinematics::PositionENU>(t); }
    Vector3<T> velocity(double t) const { return query<kinematics::VelocityENU>(t); }
    T altitude(double t) const { return query<kinematics::Altitude>(t); }
    T speed(double t) const { return velocity(t).norm(); }
    T mass(double t) const { return query<dynamics::Mass>(t); }
    Quaternion<T> attitude(double t) const { return query<kinematics::AttitudeQuaternion>(t); }

    void computeStatistics(double burn_time = 0) const {                       // :151-189
        if (m_stats_computed) return;
        m_apogee = 0; m_max_speed = 0;
        for (size_t i = 0; i < m_solution.size(); ++i) {
            const double t = m_solution.times[i];
            const auto& s = m_solution.states[i];
            const double alt = s[3], spd = std::sqrt(s[4] * s[4] + s[5] * s[5] + s[6] * s[6]);
            if (alt > m_apogee) { m_apogee = alt; m_apogee_time = t; }
            if (spd > m_max_speed) { m_max_speed = spd; m_max_speed_time = t; }
            if (burn_time > 0 && std::abs(t - burn_time) < 0.1) { m_burnout_altitude = alt; m_burnout_speed = spd; }
        }
        m_stats_computed = true;
    }
    double apogee() const { computeStatistics(); return m_apogee; }
    double apogeeTime() const { computeStatistics(); return m_apogee_time; }
    double maxSpeed() const { computeStatistics(); return m_max_speed; }
    double maxSpeedTime() const { computeStatistics(); return m_max_speed_time; }
    double burnoutAltitude() const { computeStatistics(); return m_burnout_altitude; }
    double burnoutSpeed() const { computeStatistics(); return m_burnout_speed; }

    template <typename Func>
    std::vector<double> sampleFunction(Func&& func, double t_start, double t_end, double dt) const {
        std::vector<double> v;
        v.reserve(static_cast<size_t>((t_end - t_start) / dt) + 1);
        for (double t = t_start; t <= t_end; t += dt) v.push_back(func(interpolateState(t), t));
        return v;
    }
    std::vector<double> altitudeProfile() const {
        std::vector<double> a;
        a.reserve(m_solution.size());
        for (const auto& s : m_solution.states) a.push_back(s[3]);
        return a;
    }
    std::vector<double> speedProfile() const {
        std::vector<double> a;
        a.reserve(m_solution.size());
        for (const auto& s : m_solution.states) a.push_back(std::sqrt(s[4] * s[4] + s[5] * s[5] + s[6] * s[6]));
        return a;
    }
private:
    const RocketType* m_rocket;
    SolutionResult m_solution;
    mutable double m_apogee{-1}, m_apogee_time{0}, m_max_speed{0}, m_max_speed_time{0}, m_burnout_altitude{0}, m_burnout_speed{0};
    mutable bool m_stats_computed{false};
};

// rocket/rocket.hpp:308-318
template <Scalar T = double>
SimulationResult<Rocket<T>> simulate(Rocket<T>& rocket, double t_end, double dt = 0.01) {
    auto solution = simulateRocket(rocket, t_end, dt);
    auto result = SimulationResult<Rocket<T>>(rocket, std::move(solution));
    result.computeStatistics(rocket.getBurnTime());
    return result;
}

}  // namespace sopot::rocket
