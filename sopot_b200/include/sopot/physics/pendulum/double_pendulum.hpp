// pendulum::DoublePendulum<T> -- Lagrangian double pendulum, state [theta1, theta2, omega1, omega2].
// Interface of physics/pendulum/double_pendulum.hpp:34-337; the Cramer-rule derivative (:126-179) is
// DpendSys in csrc/ho_dpend.cu.
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <string>
#include "../../core/typed_component.hpp"
#include "tags.hpp"

namespace sopot::pendulum {

template <Scalar T = double>
class DoublePendulum final : public TypedComponent<4, T> {
public:
    using Base = TypedComponent<4, T>;
    using typename Base::LocalState;
    using typename Base::LocalDerivative;

    // SOPOT_HD so that a device factory can build per-instance pendulums; argument validation (std::invalid_argument,
    // reference :86-97) happens on the host side only -- device code cannot throw
    SOPOT_HD explicit DoublePendulum(double mass1, double mass2, double length1, double length2, double gravity = 9.81,
                                     double theta1_0 = 0.0, double theta2_0 = 0.0, double omega1_0 = 0.0, double omega2_0 = 0.0)
        : m_m1(mass1), m_m2(mass2), m_L1(length1), m_L2(length2), m_g(gravity), m_ic{theta1_0, theta2_0, omega1_0, omega2_0} {
#ifndef __CUDA_ARCH__
        auto need = [](double v, const char* what) {
            if (v <= 0.0) throw std::invalid_argument(std::string(what) + " must be positive (got " + std::to_string(v) + ")");
        };
        need(mass1, "Mass 1"); need(mass2, "Mass 2"); need(length1, "Length 1"); need(length2, "Length 2");
#endif
    }
    SOPOT_HD LocalState getInitialLocalState() const { return {T(m_ic[0]), T(m_ic[1]), T(m_ic[2]), T(m_ic[3])}; }
    // :126-179, for packs that mix this component with user-defined ones (generic kernel); the single-component
    // system runs on the tuned DpendSys kernel through its SystemBinding
    template <typename Registry>
    SOPOT_HD LocalDerivative derivatives(T, std::span<const T> local, std::span<const T>, const Registry&) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T th1 = local[0], th2 = local[1], w1 = local[2], w2 = local[3];
        const T delta = th1 - th2;
        const T sd = sin(delta), cd = cos(delta), s1 = sin(th1), s2 = sin(th2);
        const T m1(m_m1), m2(m_m2), L1(m_L1), L2(m_L2), g(m_g);
        const T a11 = (m1 + m2) * L1, a12 = m2 * L2 * cd, a21 = L1 * cd, a22 = L2;
        const T b1 = m2 * L2 * w2 * w2 * sd - (m1 + m2) * g * s1;
        const T b2 = -L1 * w1 * w1 * sd - g * s2;
        const T det = a11 * a22 - a12 * a21;
        return {w1, w2, (b1 * a22 - b2 * a12) / det, (a11 * b2 - a21 * b1) / det};
    }
    std::string_view getComponentType() const { return "DoublePendulum"; }
    std::string_view getComponentName() const { return "DoublePendulum"; }

    SOPOT_HD T compute(mass1::Angle, std::span<const T> s) const { return this->getGlobalState(s, 0); }
    SOPOT_HD T compute(mass1::AngularVelocity, std::span<const T> s) const { return this->getGlobalState(s, 2); }
    SOPOT_HD T compute(mass2::Angle, std::span<const T> s) const { return this->getGlobalState(s, 1); }
    SOPOT_HD T compute(mass2::AngularVelocity, std::span<const T> s) const { return this->getGlobalState(s, 3); }
    SOPOT_HD T compute(mass1::Mass, std::span<const T>) const { return T(m_m1); }
    SOPOT_HD T compute(mass2::Mass, std::span<const T>) const { return T(m_m2); }
    SOPOT_HD T compute(system::Length1, std::span<const T>) const { return T(m_L1); }
    SOPOT_HD T compute(system::Length2, std::span<const T>) const { return T(m_L2); }
    SOPOT_HD std::array<T, 2> compute(mass1::CartesianPosition, std::span<const T> s) const {
        using std::sin; using std::cos;
        const T th1 = this->getGlobalState(s, 0);
        return {T(m_L1) * sin(th1), -T(m_L1) * cos(th1)};
    }
    SOPOT_HD std::array<T, 2> compute(mass1::CartesianVelocity, std::span<const T> s) const {
        using std::sin; using std::cos;
        const T th1 = this->getGlobalState(s, 0), w1 = this->getGlobalState(s, 2);
        return {T(m_L1) * cos(th1) * w1, T(m_L1) * sin(th1) * w1};
    }
    SOPOT_HD std::array<T, 2> compute(mass2::CartesianPosition, std::span<const T> s) const {
        using std::sin; using std::cos;
        const T th1 = this->getGlobalState(s, 0), th2 = this->getGlobalState(s, 1);
        return {T(m_L1) * sin(th1) + T(m_L2) * sin(th2), -T(m_L1) * cos(th1) - T(m_L2) * cos(th2)};
    }
    SOPOT_HD std::array<T, 2> compute(mass2::CartesianVelocity, std::span<const T> s) const {
        using std::sin; using std::cos;
        const T th1 = this->getGlobalState(s, 0), th2 = this->getGlobalState(s, 1);
        const T w1 = this->getGlobalState(s, 2), w2 = this->getGlobalState(s, 3);
        return {T(m_L1) * cos(th1) * w1 + T(m_L2) * cos(th2) * w2, T(m_L1) * sin(th1) * w1 + T(m_L2) * sin(th2) * w2};
    }
    SOPOT_HD T compute(system::KineticEnergy, std::span<const T> s) const {
        using std::cos;
        const T th1 = this->getGlobalState(s, 0), th2 = this->getGlobalState(s, 1);
        const T w1 = this->getGlobalState(s, 2), w2 = this->getGlobalState(s, 3);
        const T L1(m_L1), L2(m_L2);
        const T v1 = L1 * L1 * w1 * w1;
        const T v2 = L1 * L1 * w1 * w1 + L2 * L2 * w2 * w2 + T(2.0) * L1 * L2 * w1 * w2 * cos(th1 - th2);
        return T(0.5) * T(m_m1) * v1 + T(0.5) * T(m_m2) * v2;
    }
    SOPOT_HD T compute(system::PotentialEnergy, std::span<const T> s) const {
        using std::cos;
        const T th1 = this->getGlobalState(s, 0), th2 = this->getGlobalState(s, 1);
        return -T(m_g) * ((T(m_m1) + T(m_m2)) * T(m_L1) * cos(th1) + T(m_m2) * T(m_L2) * cos(th2));
    }
    SOPOT_HD T compute(system::TotalEnergy, std::span<const T> s) const {
        return compute(system::KineticEnergy{}, s) + compute(system::PotentialEnergy{}, s);
    }
    double getMass1() const { return m_m1; }
    double getMass2() const { return m_m2; }
    double getLength1() const { return m_L1; }
    double getLength2() const { return m_L2; }
    double getGravity() const { return m_g; }
private:
    double m_m1, m_m2, m_L1, m_L2, m_g;
    double m_ic[4];
};

struct DoublePendulumEnsemble {
    size_t n;
    b200::Param mass1, mass2, length1, length2, gravity, theta1, theta2, omega1, omega2;
    sopot_b200_dpend_params params(uint32_t& bc) const {
        const b200::Param* f[9] = {&mass1, &mass2, &length1, &length2, &gravity, &theta1, &theta2, &omega1, &omega2};
        bc = 0;
        for (int j = 0; j < 9; ++j) bc |= (f[j]->broadcast(n) ? 1u : 0u) << j;
        return {f[0]->data(), f[1]->data(), f[2]->data(), f[3]->data(), f[4]->data(), f[5]->data(), f[6]->data(),
                f[7]->data(), f[8]->data()};
    }
    b200::EnsembleResult integrate(b200::Engine& eng, size_t n_steps, double t0, double dt, const b200::EnsembleOptions& o) const {
        uint32_t bc;
        const auto p = params(bc);
        const auto opts = b200::make_opts(o, bc);
        b200::EnsembleResult r;
        r.n = n; r.dim = 4; r.n_steps = n_steps; r.store_every = o.store_every;
        r.state.resize(4 * n);
        r.trajectory.resize(r.snapshots() * 4 * n);
        if (o.want_status) r.status.resize(n);
        eng.check(sopot_b200_dpend_rk4(eng.ctx(), &p, n, t0, dt, n_steps, &opts, r.state.data(),
                                       r.trajectory.empty() ? nullptr : r.trajectory.data(),
                                       o.want_status ? r.status.data() : nullptr));
        eng.sync();
        return r;
    }
};

}  // namespace sopot::pendulum

namespace sopot::b200 {
template <>
struct SystemBinding<pendulum::DoublePendulum<double>> {
    static constexpr bool available = true;
    using Tuple = std::tuple<pendulum::DoublePendulum<double>>;
    static pendulum::DoublePendulumEnsemble one(const Tuple& t, const std::vector<double>& y) {
        const auto& c = std::get<0>(t);
        return {1, c.getMass1(), c.getMass2(), c.getLength1(), c.getLength2(), c.getGravity(), y[0], y[1], y[2], y[3]};
    }
    static std::vector<double> rhs(Engine& eng, const Tuple& t, const std::vector<double>& s) {
        const auto e = one(t, s);
        uint32_t bc;
        const auto p = e.params(bc);
        EnsembleOptions o; o.fp_mode = FpMode::Strict;
        const auto opts = make_opts(o, 0);
        std::vector<double> out(4);
        eng.check(sopot_b200_dpend_rhs(eng.ctx(), &p, 1, &opts, s.data(), out.data()));
        eng.sync();
        return out;
    }
    static EnsembleResult solve(Engine& eng, const Tuple& t, const std::vector<double>& y0, size_t n_steps, double t0,
                                double dt, const EnsembleOptions& o) {
        return one(t, y0).integrate(eng, n_steps, t0, dt, o);
    }
};
// DoublePendulum<Dual<double,K>>: the derivative in dual arithmetic = value from the derivative kernel, partials by the
// chain rule through the 4x4 state Jacobian of sopot_b200_dpend_jacobian (forward-mode AD on the device) and the caller's
// seeds -- what computeJacobian (core/typed_component.hpp) and the linearizers evaluate.
template <size_t K>
struct SystemBinding<pendulum::DoublePendulum<Dual<double, K>>> {
    static constexpr bool available = true;
    using D = Dual<double, K>;
    using Tuple = std::tuple<pendulum::DoublePendulum<D>>;
    static std::vector<D> rhs_dual(Engine& eng, const Tuple& t, const std::vector<D>& s) {
        const auto& c = std::get<0>(t);
        const double m1 = c.getMass1(), m2 = c.getMass2(), L1 = c.getLength1(), L2 = c.getLength2(), g = c.getGravity();
        const double x0[4] = {s[0].value(), s[1].value(), s[2].value(), s[3].value()};
        EnsembleOptions o; o.fp_mode = FpMode::Strict;
        const auto opts = make_opts(o, 0);
        const sopot_b200_dpend_params p{&m1, &m2, &L1, &L2, &g, nullptr, nullptr, nullptr, nullptr};
        double J[16], f[4];
        eng.check(sopot_b200_dpend_jacobian(eng.ctx(), &p, 1, &opts, x0, J));
        eng.check(sopot_b200_dpend_rhs(eng.ctx(), &p, 1, &opts, x0, f));
        eng.sync();
        std::vector<D> out(4);
        for (size_t i = 0; i < 4; ++i) {
            std::array<double, K> d{};
            for (size_t k = 0; k < K; ++k) {
                double acc = 0.0;
                for (size_t j = 0; j < 4; ++j) acc += J[i * 4 + j] * s[j].derivative(k);
                d[k] = acc;
            }
            out[i] = D(f[i], d);
        }
        return out;
    }
};
}  // namespace sopot::b200
