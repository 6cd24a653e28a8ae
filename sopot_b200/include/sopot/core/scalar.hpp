// Scalar concept, value_of and ScalarState<T,N> with the reference's interface (core/scalar.hpp:12-165).
// (units.hpp is out of scope: no hot-path component uses Quantity)
#pragma once
#include <array>
#include <concepts>
#include <cstddef>
#include "dual.hpp"

namespace sopot {

template <typename T>
concept Scalar = requires(T a, T b, double d) {
    { a + b } -> std::convertible_to<T>;
    { a - b } -> std::convertible_to<T>;
    { a * b } -> std::convertible_to<T>;
    { a / b } -> std::convertible_to<T>;
    { -a } -> std::convertible_to<T>;
    { a * d } -> std::convertible_to<T>;
    { d * a } -> std::convertible_to<T>;
};

template <typename T>
concept Differentiable = Scalar<T> && requires(T a) {
    { a.derivative(0) } -> std::convertible_to<double>;
    { a.value() } -> std::convertible_to<double>;
};

template <typename T> constexpr auto value_of(const T& x) { if constexpr (is_dual_v<T>) return x.value(); else return x; }
template <typename T> constexpr double derivative_of(const T& x, size_t i = 0) { if constexpr (is_dual_v<T>) return x.derivative(i); else return 0.0; }
template <typename T> constexpr T make_variable(double v, size_t i = 0) { if constexpr (is_dual_v<T>) return T::variable(v, i); else return T(v); }
template <typename T> constexpr T make_constant(double v) { if constexpr (is_dual_v<T>) return T::constant(v); else return T(v); }

template <typename T> struct base_scalar { using type = T; };
template <typename T, size_t N> struct base_scalar<Dual<T, N>> { using type = T; };
template <typename T> using base_scalar_t = typename base_scalar<T>::type;
template <typename T> struct num_derivatives { static constexpr size_t value = 0; };
template <typename T, size_t N> struct num_derivatives<Dual<T, N>> { static constexpr size_t value = N; };
template <typename T> inline constexpr size_t num_derivatives_v = num_derivatives<T>::value;

template <typename T, size_t N>
class ScalarState {
public:
    using scalar_type = T;
    static constexpr size_t size = N;
    constexpr ScalarState() : m{} {}
    template <typename... A>
        requires(sizeof...(A) == N) && (std::convertible_to<A, T> && ...)
    constexpr ScalarState(A... a) : m{T(a)...} {}
    constexpr T& operator[](size_t i) { return m[i]; }
    constexpr const T& operator[](size_t i) const { return m[i]; }
    constexpr T* data() { return m.data(); }
    constexpr const T* data() const { return m.data(); }
    constexpr auto begin() { return m.begin(); }
    constexpr auto end() { return m.end(); }
    constexpr auto begin() const { return m.begin(); }
    constexpr auto end() const { return m.end(); }
    std::array<double, N> values() const {
        std::array<double, N> r;
        for (size_t i = 0; i < N; ++i) r[i] = value_of(m[i]);
        return r;
    }
private:
    std::array<T, N> m;
};

}  // namespace sopot
