// control::CartDoublePendulum<T> -- cart with an inverted double pendulum, state [x, th1, th2, xd, w1, w2]
// (physics/control/cart_double_pendulum.hpp:34-403): the closed-form plant the reference checks the generic
// CartNPendulum<2> against (tests/inverted_pendulum_test.cpp:67-93).
//
// Same constructor (validation messages of :89-94), setControlForce / getControlForce, state functions (cart::, link1::,
// link2::, system::, controller::ControlInput) and parameter accessors.  The derivative is the reference's
// M(q) qdd = tau + B u solved through the adjugate of the symmetric 3x3 mass matrix (:176-229) with the same
// association of every product, as a SOPOT_HD function: on the host it is an ordinary member (the reference's test calls it
// directly); compiled by nvcc the component is trivially copyable and integrates on the device through
// sopot/b200/generic_ensemble.cuh like any other user component pack (no name string is stored for that reason).
// Where the reference throws "mass matrix is singular" (:221-223) the host build throws the same error and device code
// yields NaN accelerations, which the ensemble status reports.
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string_view>
#include "../../core/typed_component.hpp"
#include "control_tags.hpp"

namespace sopot::control {

template <Scalar T = double>
class CartDoublePendulum final : public TypedComponent<6, T> {
public:
    using Base = TypedComponent<6, T>;
    using typename Base::LocalState;
    using typename Base::LocalDerivative;

    SOPOT_HD explicit CartDoublePendulum(double cart_mass, double mass1, double mass2, double length1, double length2,
                                         double gravity = 9.81, double x0 = 0.0, double theta1_0 = 0.0, double theta2_0 = 0.0,
                                         double xdot_0 = 0.0, double omega1_0 = 0.0, double omega2_0 = 0.0)
        : m_mc(cart_mass), m_m{mass1, mass2}, m_L{length1, length2}, m_g(gravity),
          m_y0{x0, theta1_0, theta2_0, xdot_0, omega1_0, omega2_0} {
#ifndef __CUDA_ARCH__
        if (cart_mass <= 0.0 || mass1 <= 0.0 || mass2 <= 0.0) throw std::invalid_argument("All masses must be positive");
        if (length1 <= 0.0 || length2 <= 0.0) throw std::invalid_argument("Lengths must be positive");
#endif
    }

    SOPOT_HD void setControlForce(T force) const { m_force = force; }     // const + mutable as in the reference (:97-99)
    SOPOT_HD T getControlForce() const { return m_force; }

    SOPOT_HD LocalState getInitialLocalState() const {
        LocalState s;
        for (size_t i = 0; i < 6; ++i) s[i] = T(m_y0[i]);
        return s;
    }
    std::string_view getComponentType() const { return "CartDoublePendulum"; }
    std::string_view getComponentName() const { return "CartDoublePendulum"; }

    template <typename Registry>
    SOPOT_HD LocalDerivative derivatives(T, std::span<const T> local, std::span<const T>, const Registry&) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T th1 = local[1], th2 = local[2], xd = local[3], w1 = local[4], w2 = local[5];
        const T mc(m_mc), m1(m_m[0]), m2(m_m[1]), L1(m_L[0]), L2(m_L[1]), g(m_g);
        const T s1 = sin(th1), c1 = cos(th1), s2 = sin(th2), c2 = cos(th2);
        const T s12 = sin(th1 - th2), c12 = cos(th1 - th2);
        // symmetric mass matrix, upper triangle: M[0] = M11, M[1] = M12, M[2] = M13, M[3] = M22, M[4] = M23, M[5] = M33 (:176-181)
        const T M[6] = {mc + m1 + m2, (m1 + m2) * L1 * c1, m2 * L2 * c2, (m1 + m2) * L1 * L1, m2 * L1 * L2 * c12, m2 * L2 * L2};
        // generalised forces: control on the cart, centrifugal and gravity terms (:187-196)
        const T r[3] = {m_force + (m1 + m2) * L1 * s1 * w1 * w1 + m2 * L2 * s2 * w2 * w2,
                        (m1 + m2) * g * L1 * s1 - m2 * L1 * L2 * s12 * w2 * w2,
                        m2 * g * L2 * s2 + m2 * L1 * L2 * s12 * w1 * w1};
        // adjugate, row by row (:207-219), and the determinant expanded along the first row (:202-204)
        const T adj[3][3] = {
            {M[3] * M[5] - M[4] * M[4], -(M[1] * M[5] - M[2] * M[4]), M[1] * M[4] - M[2] * M[3]},
            {-(M[1] * M[5] - M[2] * M[4]), M[0] * M[5] - M[2] * M[2], -(M[0] * M[4] - M[1] * M[2])},
            {M[1] * M[4] - M[2] * M[3], -(M[0] * M[4] - M[1] * M[2]), M[0] * M[3] - M[1] * M[1]}};
        const T det = M[0] * (M[3] * M[5] - M[4] * M[4]) - M[1] * (M[1] * M[5] - M[4] * M[2]) + M[2] * (M[1] * M[4] - M[3] * M[2]);
        const bool singular = value_of(det) == 0.0;
#ifndef __CUDA_ARCH__
        if (singular) throw std::runtime_error("CartDoublePendulum: mass matrix is singular; cannot compute accelerations.");
#endif
        const T inv_det = T(1.0) / det;
        LocalDerivative d;
        d[0] = xd; d[1] = w1; d[2] = w2;
        SOPOT_UNROLL for (size_t i = 0; i < 3; ++i) {
            const T acc = inv_det * (adj[i][0] * r[0] + adj[i][1] * r[1] + adj[i][2] * r[2]);
            d[3 + i] = singular ? T(std::numeric_limits<double>::quiet_NaN()) : acc;
        }
        return d;
    }

    // ---- state functions (:236-392) ----
    SOPOT_HD T compute(cart::Position, std::span<const T> s) const { return this->getGlobalState(s, 0); }
    SOPOT_HD T compute(cart::Velocity, std::span<const T> s) const { return this->getGlobalState(s, 3); }
    SOPOT_HD T compute(cart::Mass, std::span<const T>) const { return T(m_mc); }
    SOPOT_HD T compute(link1::Angle, std::span<const T> s) const { return this->getGlobalState(s, 1); }
    SOPOT_HD T compute(link1::AngularVelocity, std::span<const T> s) const { return this->getGlobalState(s, 4); }
    SOPOT_HD T compute(link1::Mass, std::span<const T>) const { return T(m_m[0]); }
    SOPOT_HD T compute(link1::Length, std::span<const T>) const { return T(m_L[0]); }
    SOPOT_HD T compute(link2::Angle, std::span<const T> s) const { return this->getGlobalState(s, 2); }
    SOPOT_HD T compute(link2::AngularVelocity, std::span<const T> s) const { return this->getGlobalState(s, 5); }
    SOPOT_HD T compute(link2::Mass, std::span<const T>) const { return T(m_m[1]); }
    SOPOT_HD T compute(link2::Length, std::span<const T>) const { return T(m_L[1]); }
    SOPOT_HD std::array<T, 2> compute(link1::TipPosition, std::span<const T> s) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T th1 = this->getGlobalState(s, 1);
        return {this->getGlobalState(s, 0) + T(m_L[0]) * sin(th1), T(m_L[0]) * cos(th1)};
    }
    SOPOT_HD std::array<T, 2> compute(link2::TipPosition, std::span<const T> s) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T th1 = this->getGlobalState(s, 1), th2 = this->getGlobalState(s, 2);
        return {this->getGlobalState(s, 0) + T(m_L[0]) * sin(th1) + T(m_L[1]) * sin(th2), T(m_L[0]) * cos(th1) + T(m_L[1]) * cos(th2)};
    }
    SOPOT_HD std::array<T, 6> compute(system::FullState, std::span<const T> s) const {
        std::array<T, 6> y;
        for (size_t i = 0; i < 6; ++i) y[i] = this->getGlobalState(s, i);
        return y;
    }
    SOPOT_HD T compute(system::KineticEnergy, std::span<const T> s) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T th1 = this->getGlobalState(s, 1), th2 = this->getGlobalState(s, 2), xd = this->getGlobalState(s, 3);
        const T w1 = this->getGlobalState(s, 4), w2 = this->getGlobalState(s, 5);
        const T L1(m_L[0]), L2(m_L[1]);
        const T c1 = cos(th1), c2 = cos(th2);
        const T ke_cart = T(0.5) * T(m_mc) * xd * xd;
        // tip of link 1 moves with (xd + L1 c1 w1, -L1 s1 w1): v1^2 expanded as in :354
        const T v1_sq = xd * xd + T(2.0) * L1 * c1 * xd * w1 + L1 * L1 * w1 * w1;
        const T ke1 = T(0.5) * T(m_m[0]) * v1_sq;
        const T s1 = sin(th1), s2 = sin(th2);
        const T vx2 = xd + L1 * c1 * w1 + L2 * c2 * w2, vy2 = -L1 * s1 * w1 - L2 * s2 * w2;
        const T ke2 = T(0.5) * T(m_m[1]) * (vx2 * vx2 + vy2 * vy2);
        return ke_cart + ke1 + ke2;
    }
    SOPOT_HD T compute(system::PotentialEnergy, std::span<const T> s) const {
        using std::cos; using sopot::cos;
        const T th1 = this->getGlobalState(s, 1), th2 = this->getGlobalState(s, 2);
        const T L1(m_L[0]), L2(m_L[1]), g(m_g);
        const T h1 = L1 * cos(th1), h2 = L1 * cos(th1) + L2 * cos(th2);      // heights above the pivot
        return T(m_m[0]) * g * h1 + T(m_m[1]) * g * h2;
    }
    SOPOT_HD T compute(system::TotalEnergy, std::span<const T> s) const {
        return compute(system::KineticEnergy{}, s) + compute(system::PotentialEnergy{}, s);
    }
    SOPOT_HD T compute(controller::ControlInput, std::span<const T>) const { return m_force; }

    double getCartMass() const { return m_mc; }
    double getMass1() const { return m_m[0]; }
    double getMass2() const { return m_m[1]; }
    double getLength1() const { return m_L[0]; }
    double getLength2() const { return m_L[1]; }
    double getGravity() const { return m_g; }

private:
    double m_mc;
    double m_m[2], m_L[2];
    double m_g;
    double m_y0[6];
    mutable T m_force = T(0.0);
};

}  // namespace sopot::control
