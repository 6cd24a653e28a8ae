// Dual<T,N>: forward-mode AD number, API of the reference (core/dual.hpp:15-157, 181-348): variable /
// constant / value / derivative / derivatives, arithmetic with the reference's formulas, math functions
// found by ADL.  Host-side numeric TYPE only: batched Jacobians run on the device (K2,
// csrc/dual.cuh restates the same formulas) -- see core/linearization.hpp.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <ostream>
#include <type_traits>
#include "macros.hpp"

namespace sopot {

template <typename T = double, size_t N = 1>
class Dual {
public:
    using value_type = T;
    static constexpr size_t num_derivatives = N;

    constexpr Dual() noexcept : m_v(T{0}), m_d{} {}
    constexpr explicit Dual(T v) noexcept : m_v(v), m_d{} {}
    constexpr Dual(T v, std::array<T, N> d) noexcept : m_v(v), m_d(d) {}
    static constexpr Dual variable(T v, size_t index = 0) noexcept {
        Dual r(v);
        if (index < N) r.m_d[index] = T{1};
        return r;
    }
    static constexpr Dual constant(T v) noexcept { return Dual(v); }

    constexpr T value() const noexcept { return m_v; }
    constexpr T derivative(size_t i = 0) const noexcept { return i < N ? m_d[i] : T{0}; }
    constexpr const std::array<T, N>& derivatives() const noexcept { return m_d; }
    constexpr Dual noDerivs() const noexcept { return Dual(m_v); }
    constexpr explicit operator T() const noexcept { return m_v; }

    // element-wise helper: r.d[i] = f(i)
    template <class F> static constexpr Dual make(T v, F&& f) {
        Dual r(v);
        for (size_t i = 0; i < N; ++i) r.m_d[i] = f(i);
        return r;
    }

    constexpr Dual operator+(const Dual& o) const noexcept { return make(m_v + o.m_v, [&](size_t i) { return m_d[i] + o.m_d[i]; }); }
    constexpr Dual operator-(const Dual& o) const noexcept { return make(m_v - o.m_v, [&](size_t i) { return m_d[i] - o.m_d[i]; }); }
    constexpr Dual operator*(const Dual& o) const noexcept {            // u*dv + du*v   (reference :82-89)
        return make(m_v * o.m_v, [&](size_t i) { return m_v * o.m_d[i] + m_d[i] * o.m_v; });
    }
    constexpr Dual operator/(const Dual& o) const noexcept {            // (v*du - u*dv) * 1/v^2   (:91-99)
        const T inv_v2 = T{1} / (o.m_v * o.m_v);
        return make(m_v / o.m_v, [&](size_t i) { return (o.m_v * m_d[i] - m_v * o.m_d[i]) * inv_v2; });
    }
    constexpr Dual operator-() const noexcept { return make(-m_v, [&](size_t i) { return -m_d[i]; }); }
    constexpr Dual operator+(T s) const noexcept { return Dual(m_v + s, m_d); }
    constexpr Dual operator-(T s) const noexcept { return Dual(m_v - s, m_d); }
    constexpr Dual operator*(T s) const noexcept { return make(m_v * s, [&](size_t i) { return m_d[i] * s; }); }
    constexpr Dual operator/(T s) const noexcept {                       // multiplies by the reciprocal (:126-132)
        const T inv = T{1} / s;
        return make(m_v * inv, [&](size_t i) { return m_d[i] * inv; });
    }
    constexpr Dual& operator+=(const Dual& o) noexcept { return *this = *this + o; }
    constexpr Dual& operator-=(const Dual& o) noexcept { return *this = *this - o; }
    constexpr Dual& operator*=(const Dual& o) noexcept { return *this = *this * o; }
    constexpr Dual& operator/=(const Dual& o) noexcept { return *this = *this / o; }
    constexpr Dual& operator+=(T s) noexcept { return *this = *this + s; }
    constexpr Dual& operator-=(T s) noexcept { return *this = *this - s; }
    constexpr Dual& operator*=(T s) noexcept { return *this = *this * s; }
    constexpr Dual& operator/=(T s) noexcept { return *this = *this / s; }

    // comparisons look at the value only (:146-156)
    constexpr bool operator==(const Dual& o) const noexcept { return m_v == o.m_v; }
    constexpr bool operator!=(const Dual& o) const noexcept { return m_v != o.m_v; }
    constexpr bool operator<(const Dual& o) const noexcept { return m_v < o.m_v; }
    constexpr bool operator<=(const Dual& o) const noexcept { return m_v <= o.m_v; }
    constexpr bool operator>(const Dual& o) const noexcept { return m_v > o.m_v; }
    constexpr bool operator>=(const Dual& o) const noexcept { return m_v >= o.m_v; }
    constexpr bool operator==(T s) const noexcept { return m_v == s; }
    constexpr bool operator<(T s) const noexcept { return m_v < s; }
    constexpr bool operator>(T s) const noexcept { return m_v > s; }

private:
    T m_v;
    std::array<T, N> m_d;
};

template <typename T, size_t N> constexpr Dual<T, N> operator+(T s, const Dual<T, N>& d) noexcept { return d + s; }
template <typename T, size_t N> constexpr Dual<T, N> operator-(T s, const Dual<T, N>& d) noexcept { return Dual<T, N>::constant(s) - d; }
template <typename T, size_t N> constexpr Dual<T, N> operator*(T s, const Dual<T, N>& d) noexcept { return d * s; }
template <typename T, size_t N> constexpr Dual<T, N> operator/(T s, const Dual<T, N>& d) noexcept { return Dual<T, N>::constant(s) / d; }

// chain rule: f(u) with scalar factor f'(u)
template <typename T, size_t N> SOPOT_HD Dual<T, N> dual_chain(const Dual<T, N>& x, T value, T factor) {
    return Dual<T, N>::make(value, [&](size_t i) { return factor * x.derivative(i); });
}
template <typename T, size_t N> SOPOT_HD Dual<T, N> sin(const Dual<T, N>& x) { return dual_chain(x, std::sin(x.value()), std::cos(x.value())); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> cos(const Dual<T, N>& x) { return dual_chain(x, std::cos(x.value()), -std::sin(x.value())); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> tan(const Dual<T, N>& x) {
    const T c = std::cos(x.value());
    return dual_chain(x, std::tan(x.value()), T{1} / (c * c));
}
template <typename T, size_t N> SOPOT_HD Dual<T, N> asin(const Dual<T, N>& x) { return dual_chain(x, std::asin(x.value()), T{1} / std::sqrt(T{1} - x.value() * x.value())); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> acos(const Dual<T, N>& x) { return dual_chain(x, std::acos(x.value()), T{-1} / std::sqrt(T{1} - x.value() * x.value())); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> atan(const Dual<T, N>& x) { return dual_chain(x, std::atan(x.value()), T{1} / (T{1} + x.value() * x.value())); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> atan2(const Dual<T, N>& y, const Dual<T, N>& x) {
    const T inv = T{1} / (x.value() * x.value() + y.value() * y.value());
    return Dual<T, N>::make(std::atan2(y.value(), x.value()),
                            [&](size_t i) { return (x.value() * y.derivative(i) - y.value() * x.derivative(i)) * inv; });
}
template <typename T, size_t N> SOPOT_HD Dual<T, N> exp(const Dual<T, N>& x) { const T e = std::exp(x.value()); return dual_chain(x, e, e); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> log(const Dual<T, N>& x) { return dual_chain(x, std::log(x.value()), T{1} / x.value()); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> sqrt(const Dual<T, N>& x) { const T s = std::sqrt(x.value()); return dual_chain(x, s, T{0.5} / s); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> abs(const Dual<T, N>& x) { return dual_chain(x, std::abs(x.value()), x.value() >= T{0} ? T{1} : T{-1}); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> tanh(const Dual<T, N>& x) { const T t = std::tanh(x.value()); return dual_chain(x, t, T{1} - t * t); }
template <typename T, size_t N> SOPOT_HD Dual<T, N> pow(const Dual<T, N>& b, T e) {
    return dual_chain(b, std::pow(b.value(), e), e * std::pow(b.value(), e - T{1}));
}
template <typename T, size_t N> SOPOT_HD Dual<T, N> pow(const Dual<T, N>& b, const Dual<T, N>& e) {
    const T p = std::pow(b.value(), e.value()), lb = std::log(b.value()), ib = T{1} / b.value();
    return Dual<T, N>::make(p, [&](size_t i) { return p * (e.derivative(i) * lb + e.value() * b.derivative(i) * ib); });
}
template <typename T, size_t N> SOPOT_HD Dual<T, N> sign(const Dual<T, N>& x) {
    return Dual<T, N>((x.value() > T{0}) ? T{1} : ((x.value() < T{0}) ? T{-1} : T{0}));
}
template <typename T, size_t N> std::ostream& operator<<(std::ostream& os, const Dual<T, N>& d) {
    os << "(" << d.value() << "; [";
    for (size_t i = 0; i < N; ++i) os << (i ? ", " : "") << d.derivative(i);
    return os << "])";
}

template <typename T> struct is_dual : std::false_type {};
template <typename T, size_t N> struct is_dual<Dual<T, N>> : std::true_type {};
template <typename T> inline constexpr bool is_dual_v = is_dual<T>::value;
template <typename T> constexpr auto get_value(const T& x) { if constexpr (is_dual_v<T>) return x.value(); else return x; }

using Dual1 = Dual<double, 1>;
using Dual3 = Dual<double, 3>;
using Dual6 = Dual<double, 6>;
using Dual13 = Dual<double, 13>;

}  // namespace sopot
