// physics::HarmonicOscillator<T> -- m x'' + c x' + k x = 0, state [x, v].
// Constructor, validation, state functions and factories follow physics/harmonic_oscillator.hpp:44-167.
// A system made of this component alone runs on the hand-tuned kernel (HoSys, csrc/ho_dpend.cu) through its
// SystemBinding; the SOPOT_HD derivatives() below ({v, (-k/m) x - (c/m) v}, :75-85) exists so that packs which MIX it
// with user-defined components compile into a generic kernel (b200/generic_ensemble.cuh).  Nothing calls it on the host.
#pragma once
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include "../core/typed_component.hpp"

namespace sopot::physics {

template <Scalar T = double>
class HarmonicOscillator final : public TypedComponent<2, T> {
public:
    using Base = TypedComponent<2, T>;
    using typename Base::LocalState;
    using typename Base::LocalDerivative;

    HarmonicOscillator(double mass, double spring_constant, double damping = 0.0, double initial_position = 1.0,
                       double initial_velocity = 0.0, std::string name = "harmonic_oscillator")
        : m_mass(mass), m_k(spring_constant), m_c(damping), m_x0(initial_position), m_v0(initial_velocity),
          m_omega(std::sqrt(spring_constant / mass)) {
        std::strncpy(m_name, name.c_str(), sizeof(m_name) - 1);      // fixed storage: the component stays trivially copyable
        if (mass <= 0.0) throw std::invalid_argument("Mass must be positive");
        if (spring_constant <= 0.0) throw std::invalid_argument("Spring constant must be positive");
        if (damping < 0.0) throw std::invalid_argument("Damping must be non-negative");
    }
    SOPOT_HD LocalState getInitialLocalState() const { return {T(m_x0), T(m_v0)}; }
    template <typename Registry>
    SOPOT_HD LocalDerivative derivatives(T, std::span<const T> local, std::span<const T>, const Registry&) const {
        return {local[1], T(-m_k / m_mass) * local[0] - T(m_c / m_mass) * local[1]};
    }
    std::string_view getComponentType() const { return "HarmonicOscillator"; }
    std::string_view getComponentName() const { return m_name; }

    SOPOT_HD T compute(kinematics::Position, std::span<const T> s) const { return this->getGlobalState(s, 0); }
    SOPOT_HD T compute(kinematics::Velocity, std::span<const T> s) const { return this->getGlobalState(s, 1); }
    SOPOT_HD T compute(kinematics::Acceleration, std::span<const T> s) const {
        return T(-m_k / m_mass) * this->getGlobalState(s, 0) - T(m_c / m_mass) * this->getGlobalState(s, 1);
    }
    SOPOT_HD T compute(energy::Kinetic, std::span<const T> s) const { const T v = this->getGlobalState(s, 1); return T(0.5 * m_mass) * v * v; }
    SOPOT_HD T compute(energy::Potential, std::span<const T> s) const { const T x = this->getGlobalState(s, 0); return T(0.5 * m_k) * x * x; }
    SOPOT_HD T compute(energy::Total, std::span<const T> s) const { return compute(energy::Kinetic{}, s) + compute(energy::Potential{}, s); }
    SOPOT_HD T compute(dynamics::Mass, std::span<const T>) const { return T(m_mass); }

    double getMass() const noexcept { return m_mass; }
    double getSpringConstant() const noexcept { return m_k; }
    double getDamping() const noexcept { return m_c; }
    double getNaturalFrequency() const noexcept { return m_omega; }
    double getInitialPosition() const noexcept { return m_x0; }
    double getInitialVelocity() const noexcept { return m_v0; }
    std::pair<double, double> analyticalSolution(double t) const {
        if (m_c != 0.0) throw std::runtime_error("Analytical solution only available for undamped case");
        const double A = m_x0, B = m_v0 / m_omega;
        return {A * std::cos(m_omega * t) + B * std::sin(m_omega * t),
                -A * m_omega * std::sin(m_omega * t) + B * m_omega * std::cos(m_omega * t)};
    }
private:
    double m_mass, m_k, m_c, m_x0, m_v0, m_omega;
    char m_name[24] = {};
};

template <Scalar T = double>
HarmonicOscillator<T> createSimpleOscillator(double mass = 1.0, double k = 1.0, double x0 = 1.0) {
    return HarmonicOscillator<T>(mass, k, 0.0, x0, 0.0);
}
template <Scalar T = double>
HarmonicOscillator<T> createDampedOscillator(double mass = 1.0, double k = 1.0, double c = 0.1, double x0 = 1.0) {
    return HarmonicOscillator<T>(mass, k, c, x0, 0.0);
}

// SoA parameter block for RK4Solver::solveEnsemble: every field is one value or one value per instance
struct HarmonicOscillatorEnsemble {
    size_t n;
    b200::Param mass, spring_constant, damping, initial_position, initial_velocity;
    b200::EnsembleResult integrate(b200::Engine& eng, size_t n_steps, double t0, double dt, const b200::EnsembleOptions& o) const {
        const b200::Param* f[5] = {&mass, &spring_constant, &damping, &initial_position, &initial_velocity};
        uint32_t bc = 0;
        for (int j = 0; j < 5; ++j) bc |= (f[j]->broadcast(n) ? 1u : 0u) << j;
        sopot_b200_ho_params p{f[0]->data(), f[1]->data(), f[2]->data(), f[3]->data(), f[4]->data()};
        const auto opts = b200::make_opts(o, bc);
        b200::EnsembleResult r;
        r.n = n; r.dim = 2; r.n_steps = n_steps; r.store_every = o.store_every;
        r.state.resize(2 * n);
        r.trajectory.resize(r.snapshots() * 2 * n);
        if (o.want_status) r.status.resize(n);
        eng.check(sopot_b200_ho_rk4(eng.ctx(), &p, n, t0, dt, n_steps, &opts, r.state.data(),
                                    r.trajectory.empty() ? nullptr : r.trajectory.data(),
                                    o.want_status ? r.status.data() : nullptr));
        eng.sync();
        return r;
    }
};

}  // namespace sopot::physics

namespace sopot::b200 {
template <>
struct SystemBinding<physics::HarmonicOscillator<double>> {
    static constexpr bool available = true;
    using Tuple = std::tuple<physics::HarmonicOscillator<double>>;
    static physics::HarmonicOscillatorEnsemble one(const Tuple& t, double x, double v) {
        const auto& c = std::get<0>(t);
        return {1, c.getMass(), c.getSpringConstant(), c.getDamping(), x, v};
    }
    static std::vector<double> rhs(Engine& eng, const Tuple& t, const std::vector<double>& s) {
        const auto e = one(t, s[0], s[1]);
        sopot_b200_ho_params p{e.mass.data(), e.spring_constant.data(), e.damping.data(), nullptr, nullptr};
        EnsembleOptions o; o.fp_mode = FpMode::Strict;
        const auto opts = make_opts(o, 0);
        std::vector<double> out(2);
        eng.check(sopot_b200_ho_rhs(eng.ctx(), &p, 1, &opts, s.data(), out.data()));
        eng.sync();
        return out;
    }
    static EnsembleResult solve(Engine& eng, const Tuple& t, const std::vector<double>& y0, size_t n_steps, double t0,
                                double dt, const EnsembleOptions& o) {
        return one(t, y0[0], y0[1]).integrate(eng, n_steps, t0, dt, o);
    }
};
}  // namespace sopot::b200
