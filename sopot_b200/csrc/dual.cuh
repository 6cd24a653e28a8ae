// dual.cuh -- device forward-mode dual number, value + N partials in registers.
// Operation formulas follow the reference's Dual<T,N> one for one (core/dual.hpp:66-133, 181-201)
// so that a Strict instantiation reproduces its rounding:  *  : u*dv + du*v ;
//   /  : inv = 1/(v*v); (v*du - u*dv)*inv ; value u/v ;   sin: cos(u)*du ;  cos: (-sin(u))*du.
#pragma once
#include "common.cuh"

template <int N, bool S>
struct DualR {
    Real<S> v;
    Real<S> d[N];
    __device__ __forceinline__ DualR() : v(0.0) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = Real<S>(0.0);
    }
    __device__ __forceinline__ explicit DualR(double x) : v(x) {      // Dual(T value): constant
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = Real<S>(0.0);
    }
    __device__ __forceinline__ explicit DualR(Real<S> x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = Real<S>(0.0);
    }
    static __device__ __forceinline__ DualR variable(double x, int idx) {   // core/dual.hpp:35-41
        DualR r(x);
#pragma unroll
        for (int i = 0; i < N; ++i) if (i == idx) r.d[i] = Real<S>(1.0);
        return r;
    }
};

template <int N, bool S> __device__ __forceinline__ DualR<N, S> operator+(const DualR<N, S>& a, const DualR<N, S>& b) {
    DualR<N, S> r; r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int N, bool S> __device__ __forceinline__ DualR<N, S> operator-(const DualR<N, S>& a, const DualR<N, S>& b) {
    DualR<N, S> r; r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int N, bool S> __device__ __forceinline__ DualR<N, S> operator*(const DualR<N, S>& a, const DualR<N, S>& b) {
    DualR<N, S> r; r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v;
    return r;
}
template <int N, bool S> __device__ __forceinline__ DualR<N, S> operator/(const DualR<N, S>& a, const DualR<N, S>& b) {
    DualR<N, S> r;
    const Real<S> inv_v2 = Real<S>(1.0) / (b.v * b.v);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = (b.v * a.d[i] - a.v * b.d[i]) * inv_v2;
    r.v = a.v / b.v;
    return r;
}
template <int N, bool S> __device__ __forceinline__ DualR<N, S> r_sin(const DualR<N, S>& x) {
    DualR<N, S> r; Real<S> s, c; r_sincos(x.v, s, c);
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = c * x.d[i];
    r.v = s;
    return r;
}
template <int N, bool S> __device__ __forceinline__ DualR<N, S> r_cos(const DualR<N, S>& x) {
    DualR<N, S> r; Real<S> s, c; r_sincos(x.v, s, c);
    const Real<S> ns = -s;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = ns * x.d[i];
    r.v = c;
    return r;
}

template <int N, bool S> __device__ __forceinline__ void r_sincos(const DualR<N, S>& x, DualR<N, S>& sn, DualR<N, S>& cs) {
    Real<S> s, c; r_sincos(x.v, s, c);
    const Real<S> ns = -s;
#pragma unroll
    for (int i = 0; i < N; ++i) { sn.d[i] = c * x.d[i]; cs.d[i] = ns * x.d[i]; }
    sn.v = s; cs.v = c;
}

template <bool S> __device__ __forceinline__ double sb_value_of(const Real<S>& x) { return x.v; }
template <int N, bool S> __device__ __forceinline__ double sb_value_of(const DualR<N, S>& x) { return x.v.v; }

// scalar construction from a plain double for either scalar kind: T(x) in the reference
template <class T> struct ScalarOf;
template <bool S> struct ScalarOf<Real<S>> { static __device__ __forceinline__ Real<S> make(double x) { return Real<S>(x); } };
template <int N, bool S> struct ScalarOf<DualR<N, S>> { static __device__ __forceinline__ DualR<N, S> make(double x) { return DualR<N, S>(x); } };
