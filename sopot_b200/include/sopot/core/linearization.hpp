// sopot/core/linearization.hpp -- LinearizedSystem / SystemLinearizer / makeLinearizer with the reference's
// interface (core/linearization.hpp:24-126).  linearize() seeds Dual<double,N+M> variables exactly as the
// reference (:86-95), calls the user's dynamics once and reads A, B out of the partials (:102-111).  The
// dynamics of a shipped plant evaluate on the device: TypedODESystem<Dual<...>>::computeDerivatives launches
// the batched Dual linearization kernel (K2) for that one point.  For many operating points use the batched
// entry points next to the plant (control::linearizeCartPoleBatch) -- one launch for all of them.
#pragma once
#include <array>
#include <functional>
#include <iomanip>
#include <iostream>
#include "dual.hpp"
#include "typed_component.hpp"

namespace sopot {

template <size_t Rows, size_t Cols> using Matrix = std::array<std::array<double, Cols>, Rows>;
template <size_t N> using Vector = std::array<double, N>;

template <size_t StateSize, size_t InputSize>
struct LinearizedSystem {
    Matrix<StateSize, StateSize> A{};
    Matrix<StateSize, InputSize> B{};
    Vector<StateSize> x0{};
    Vector<InputSize> u0{};
    double t0{0.0};
    void print(std::ostream& os = std::cout) const {
        os << std::fixed << std::setprecision(6) << "A matrix (" << StateSize << "x" << StateSize << "):\n";
        for (const auto& row : A) { os << "  ["; for (size_t j = 0; j < StateSize; ++j) os << std::setw(12) << row[j] << (j + 1 < StateSize ? ", " : ""); os << "]\n"; }
        os << "B matrix (" << StateSize << "x" << InputSize << "):\n";
        for (const auto& row : B) { os << "  ["; for (size_t j = 0; j < InputSize; ++j) os << std::setw(12) << row[j] << (j + 1 < InputSize ? ", " : ""); os << "]\n"; }
    }
};

template <size_t StateSize, size_t InputSize>
class SystemLinearizer {
public:
    static constexpr size_t TotalDerivs = StateSize + InputSize;
    using DualT = Dual<double, TotalDerivs>;
    using DualState = std::array<DualT, StateSize>;
    using DualInput = std::array<DualT, InputSize>;
    using DualDerivative = std::array<DualT, StateSize>;
    using DynamicsFunc = std::function<DualDerivative(DualT, const DualState&, const DualInput&)>;

    explicit SystemLinearizer(DynamicsFunc dynamics) : m_dynamics(std::move(dynamics)) {}

    LinearizedSystem<StateSize, InputSize> linearize(double t, const Vector<StateSize>& x, const Vector<InputSize>& u) const {
        LinearizedSystem<StateSize, InputSize> r;
        r.x0 = x; r.u0 = u; r.t0 = t;
        DualState dx;
        for (size_t i = 0; i < StateSize; ++i) dx[i] = DualT::variable(x[i], i);
        DualInput du;
        for (size_t i = 0; i < InputSize; ++i) du[i] = DualT::variable(u[i], StateSize + i);
        const DualDerivative xdot = m_dynamics(DualT::constant(t), dx, du);
        for (size_t i = 0; i < StateSize; ++i) {
            for (size_t j = 0; j < StateSize; ++j) r.A[i][j] = xdot[i].derivative(j);
            for (size_t j = 0; j < InputSize; ++j) r.B[i][j] = xdot[i].derivative(StateSize + j);
        }
        return r;
    }
private:
    DynamicsFunc m_dynamics;
};

template <size_t StateSize, size_t InputSize>
auto makeLinearizer(typename SystemLinearizer<StateSize, InputSize>::DynamicsFunc dynamics) {
    return SystemLinearizer<StateSize, InputSize>(std::move(dynamics));
}

// autonomous form f(t, x) (core/linearization.hpp:130-175)
template <size_t StateSize>
class AutonomousLinearizer {
public:
    using DualT = Dual<double, StateSize>;
    using DualState = std::array<DualT, StateSize>;
    using DynamicsFunc = std::function<DualState(DualT, const DualState&)>;
    explicit AutonomousLinearizer(DynamicsFunc f) : m_dynamics(std::move(f)) {}
    Matrix<StateSize, StateSize> computeJacobian(double t, const Vector<StateSize>& x) const {
        DualState dx;
        for (size_t i = 0; i < StateSize; ++i) dx[i] = DualT::variable(x[i], i);
        const DualState xdot = m_dynamics(DualT::constant(t), dx);
        Matrix<StateSize, StateSize> A{};
        for (size_t i = 0; i < StateSize; ++i)
            for (size_t j = 0; j < StateSize; ++j) A[i][j] = xdot[i].derivative(j);
        return A;
    }
private:
    DynamicsFunc m_dynamics;
};
template <size_t StateSize>
auto makeAutonomousLinearizer(typename AutonomousLinearizer<StateSize>::DynamicsFunc f) {
    return AutonomousLinearizer<StateSize>(std::move(f));
}

}  // namespace sopot

namespace sopot {
// Jacobian of a whole TypedODESystem<Dual<double,N>, ...> (core/linearization.hpp:179-218)
template <typename... Components>
class ODESystemLinearizer {
public:
    static constexpr size_t StateSize = (Components::state_size + ...);
    using DualT = Dual<double, StateSize>;
    explicit ODESystemLinearizer(TypedODESystem<DualT, Components...>& system) : m_system(system) {}
    Matrix<StateSize, StateSize> computeJacobian(double t, const Vector<StateSize>& x) const {
        std::vector<DualT> s(StateSize);
        for (size_t i = 0; i < StateSize; ++i) s[i] = DualT::variable(x[i], i);
        const auto d = m_system.computeDerivatives(DualT::constant(t), s);
        Matrix<StateSize, StateSize> A{};
        for (size_t i = 0; i < StateSize; ++i)
            for (size_t j = 0; j < StateSize; ++j) A[i][j] = d[i].derivative(j);
        return A;
    }
private:
    TypedODESystem<DualT, Components...>& m_system;
};
template <typename... Components>
auto makeODESystemLinearizer(TypedODESystem<Dual<double, (Components::state_size + ...)>, Components...>& system) {
    return ODESystemLinearizer<Components...>(system);
}
}  // namespace sopot
