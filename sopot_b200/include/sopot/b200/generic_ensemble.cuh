// sopot/b200/generic_ensemble.cuh -- the plugin path: RK4 ensemble kernels for USER-DEFINED component systems.
//
// Included automatically by core/typed_component.hpp when the program is compiled with nvcc
//     nvcc -std=c++20 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a  user.cu  -lsopot_b200
// A TypedODESystem whose stateful components implement the reference's
//     derivatives(T t, std::span<const T> local, std::span<const T> global, const Registry&) const
// as SOPOT_HD functions (and are trivially copyable: no std::string members) gets its own instantiation of the kernel
// below, in the user's translation unit.  One instance per thread; the state vector, the RK4 accumulator and the stage
// vector are plain local arrays that ptxas keeps in registers because every offset (TypedODESystem::offset_array) and
// every registry lookup (TypedRegistry::providerIndex) is a compile-time constant after inlining -- the generic kernel
// of a two-mass/spring system compiles to 44 registers and no local memory.
//
// The step is RK4Solver::solve's, written in the reference's association (core/solver.hpp:95-125):
//     k_i = dt * f(.),   y += (k1 + 2.0*k2 + 2.0*k3 + k4) / 6.0   left to right,   t += dt.
// Floating-point mode is the translation unit's: nvcc's default -fmad=true contracts a*b+c into DFMA (differences at
// rounding level, <= 1e-12 relative per step); -fmad=false evaluates every operation separately and reproduces the
// reference's g++ -O3 result bit for bit wherever the components use + - * / sqrt only.
//
// Per-instance parameters come from a FACTORY: a functor passed to the kernel by value with
//     SOPOT_HD System operator()(size_t i) const      (__device__ alone works when the return type is spelled out:
//                                                      nvcc's host pass cannot deduce `auto` of a __device__ function)
// that builds the i-th system (typically from device arrays it holds pointers to).  The integration starts from the
// components' own initial states unless an explicit y0 array is given.
#pragma once
#ifndef __CUDACC__
#error "generic_ensemble.cuh needs nvcc: user-defined component systems are compiled into CUDA kernels"
#endif
#include <cuda_runtime.h>
#include <cmath>
#include <string>
#include <type_traits>
#include <vector>
#include "engine.hpp"
#include "../core/macros.hpp"

namespace sopot::b200::generic {

inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string("sopot-b200: ") + what + ": " + cudaGetErrorString(e));
}

template <typename Factory>
using system_of_t = decltype(std::declval<const Factory&>()(size_t{0}));

// every instance is a copy of one host-configured system (N = 1 launches, ensembles without parameter jitter)
template <typename System>
struct CopyFactory {
    System sys;
    __device__ System operator()(size_t) const { return sys; }
};

// A CONTROLLER is an optional functor  SOPOT_HD void operator()(size_t i, double t, const double (&y)[D], System& sys) const
// that runs once before every RK4 step and may reconfigure the instance's system from the state at the start of the step
// (a sampled-data control law: e.g. sys.getComponent<0>().setControlForce(-K x), held over the four stages).
struct NoController {
    template <typename System, size_t D>
    SOPOT_HD void operator()(size_t, double, const double (&)[D], System&) const {}
};

template <typename Factory, typename Controller>
__global__ void __launch_bounds__(128)
rk4_kernel(const Factory make, const Controller control, const size_t n, const double t0, const double dt, const size_t n_steps,
           const size_t store_every, const double* __restrict__ y0, double* __restrict__ state_out, double* __restrict__ traj_out,
           int32_t* __restrict__ status) {
    using System = system_of_t<Factory>;
    constexpr size_t D = System::getStateDimension();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    System sys = make(i);
    sys.initializeOffsets();              // offsets become compile-time constants of this local copy
    double y[D], k[D], acc[D], tmp[D];
    if (y0) {
#pragma unroll
        for (size_t d = 0; d < D; ++d) y[d] = y0[d * n + i];
    } else {
        sys.getInitialStateInto(y);
    }
    double t = t0;
    size_t until_snap = store_every, snap = 0;
    for (size_t step = 0; step < n_steps; ++step) {
        control(i, t, y, sys);
        sys.computeDerivativesInto(t, y, k);
#pragma unroll
        for (size_t d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = k[d]; tmp[d] = y[d] + 0.5 * k[d]; }
        sys.computeDerivativesInto(t + 0.5 * dt, tmp, k);
#pragma unroll
        for (size_t d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = acc[d] + 2.0 * k[d]; tmp[d] = y[d] + 0.5 * k[d]; }
        sys.computeDerivativesInto(t + 0.5 * dt, tmp, k);
#pragma unroll
        for (size_t d = 0; d < D; ++d) { k[d] = k[d] * dt; acc[d] = acc[d] + 2.0 * k[d]; tmp[d] = y[d] + k[d]; }
        sys.computeDerivativesInto(t + dt, tmp, k);
#pragma unroll
        for (size_t d = 0; d < D; ++d) { k[d] = k[d] * dt; y[d] = y[d] + (acc[d] + k[d]) / 6.0; }
        t += dt;
        if (store_every && --until_snap == 0) {
            until_snap = store_every;
            if (traj_out) {
#pragma unroll
                for (size_t d = 0; d < D; ++d) traj_out[(snap * D + d) * n + i] = y[d];
            }
            ++snap;
        }
    }
    bool finite = true;
#pragma unroll
    for (size_t d = 0; d < D; ++d) { state_out[d * n + i] = y[d]; finite &= isfinite(y[d]); }
    if (status) status[i] = finite ? SOPOT_B200_ST_OK : SOPOT_B200_ST_NONFINITE;
}

template <typename Factory>
__global__ void __launch_bounds__(128)
rhs_kernel(const Factory make, const size_t n, const double t, const double* __restrict__ state, double* __restrict__ out) {
    using System = system_of_t<Factory>;
    constexpr size_t D = System::getStateDimension();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    System sys = make(i);
    sys.initializeOffsets();
    double y[D], dy[D];
#pragma unroll
    for (size_t d = 0; d < D; ++d) y[d] = state[d * n + i];
    sys.computeDerivativesInto(t, y, dy);
#pragma unroll
    for (size_t d = 0; d < D; ++d) out[d * n + i] = dy[d];
}

// one derivative evaluation of a pack instantiated on Dual<double,K>: state and result travel as [D][K+1][n] doubles
// (value, then the K partials); the arithmetic is the components' own code in the reference's Dual formulas (core/dual.hpp)
template <typename Factory>
__global__ void __launch_bounds__(128)
rhs_dual_kernel(const Factory make, const size_t n, const double t, const double* __restrict__ state, double* __restrict__ out) {
    using System = system_of_t<Factory>;
    using T = typename System::scalar_type;
    constexpr size_t D = System::getStateDimension(), K = T::num_derivatives;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    System sys = make(i);
    sys.initializeOffsets();
    T y[D], dy[D];
#pragma unroll
    for (size_t d = 0; d < D; ++d)
        y[d] = T::make(state[(d * (K + 1)) * n + i], [&](size_t k) { return state[(d * (K + 1) + 1 + k) * n + i]; });
    sys.computeDerivativesInto(T::constant(t), y, dy);
#pragma unroll
    for (size_t d = 0; d < D; ++d) {
        out[(d * (K + 1)) * n + i] = dy[d].value();
#pragma unroll
        for (size_t k = 0; k < K; ++k) out[(d * (K + 1) + 1 + k) * n + i] = dy[d].derivative(k);
    }
}

// RAII device buffer on the engine's device
template <typename U>
struct DeviceBuffer {
    U* p = nullptr;
    explicit DeviceBuffer(size_t count) { if (count) cuda_check(cudaMalloc(&p, count * sizeof(U)), "cudaMalloc"); }
    ~DeviceBuffer() { if (p) cudaFree(p); }
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
};

inline cudaStream_t stream_of(Engine& eng) {
    void* s = nullptr;
    eng.check(sopot_b200_get_stream(eng.ctx(), &s));
    return static_cast<cudaStream_t>(s);
}

// Integrate n instances built by `make`; results come back as host SoA vectors like every other ensemble entry point.
template <typename Factory, typename Controller = NoController>
EnsembleResult integrate(Engine& eng, const Factory& make, size_t n, size_t n_steps, double t0, double dt,
                         const EnsembleOptions& o, const double* y0_host = nullptr, const Controller& control = Controller{}) {
    using System = system_of_t<Factory>;
    static_assert(System::device_integrable, "every stateful component needs SOPOT_HD derivatives() and must be trivially copyable");
    // (the factory travels to the kernel by value: std::tuple is not formally trivially copyable in libstdc++, its
    // trivially copyable elements -- checked by device_integrable -- are what makes the bitwise copy sound)
    static_assert(std::is_trivially_destructible_v<Factory>, "the factory is passed to the kernel by value");
    constexpr size_t D = System::getStateDimension();
    if (!(dt > 0.0)) throw std::invalid_argument("sopot-b200: dt must be positive");
    EnsembleResult r;
    r.n = n; r.dim = D; r.n_steps = n_steps; r.store_every = o.store_every;
    r.state.resize(D * n);
    r.trajectory.resize(r.snapshots() * D * n);
    if (o.want_status) r.status.resize(n);
    if (n == 0) return r;
    cudaStream_t st = stream_of(eng);
    DeviceBuffer<double> d_state(D * n), d_traj(r.trajectory.size()), d_y0(y0_host ? D * n : 0);
    DeviceBuffer<int32_t> d_status(o.want_status ? n : 0);
    if (y0_host) cuda_check(cudaMemcpyAsync(d_y0.p, y0_host, D * n * sizeof(double), cudaMemcpyHostToDevice, st), "H2D y0");
    rk4_kernel<Factory, Controller><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(make, control, n, t0, dt, n_steps, o.store_every,
                                                                                  d_y0.p, d_state.p, d_traj.p, d_status.p);
    cuda_check(cudaGetLastError(), "generic rk4 kernel launch");
    eng.check(sopot_b200_note_launches(eng.ctx(), 1));
    cuda_check(cudaMemcpyAsync(r.state.data(), d_state.p, D * n * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H state");
    if (!r.trajectory.empty())
        cuda_check(cudaMemcpyAsync(r.trajectory.data(), d_traj.p, r.trajectory.size() * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H trajectory");
    if (o.want_status) cuda_check(cudaMemcpyAsync(r.status.data(), d_status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "D2H status");
    cuda_check(cudaStreamSynchronize(st), "generic rk4 kernel");
    return r;
}

template <typename System>
std::vector<double> rhs_one(const System& sys, const std::vector<double>& state) {
    static_assert(System::device_integrable, "every stateful component needs SOPOT_HD derivatives() and must be trivially copyable");
    constexpr size_t D = System::getStateDimension();
    Engine& eng = Engine::shared();
    cudaStream_t st = stream_of(eng);
    DeviceBuffer<double> d_in(D), d_out(D);
    std::vector<double> out(D);
    cuda_check(cudaMemcpyAsync(d_in.p, state.data(), D * sizeof(double), cudaMemcpyHostToDevice, st), "H2D state");
    rhs_kernel<CopyFactory<System>><<<1, 128, 0, st>>>(CopyFactory<System>{sys}, 1, 0.0, d_in.p, d_out.p);
    cuda_check(cudaGetLastError(), "generic rhs kernel launch");
    eng.check(sopot_b200_note_launches(eng.ctx(), 1));
    cuda_check(cudaMemcpyAsync(out.data(), d_out.p, D * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H derivative");
    cuda_check(cudaStreamSynchronize(st), "generic rhs kernel");
    return out;
}

template <typename System>
std::vector<typename System::scalar_type> rhs_one_dual(const System& sys, const std::vector<typename System::scalar_type>& state) {
    using T = typename System::scalar_type;
    static_assert(System::device_differentiable, "every stateful component needs SOPOT_HD derivatives() and must be trivially copyable");
    constexpr size_t D = System::getStateDimension(), K = T::num_derivatives, W = D * (K + 1);
    Engine& eng = Engine::shared();
    cudaStream_t st = stream_of(eng);
    std::vector<double> flat(W), res(W);
    for (size_t d = 0; d < D; ++d) {
        flat[d * (K + 1)] = state[d].value();
        for (size_t k = 0; k < K; ++k) flat[d * (K + 1) + 1 + k] = state[d].derivative(k);
    }
    DeviceBuffer<double> d_in(W), d_out(W);
    cuda_check(cudaMemcpyAsync(d_in.p, flat.data(), W * sizeof(double), cudaMemcpyHostToDevice, st), "H2D dual state");
    rhs_dual_kernel<CopyFactory<System>><<<1, 128, 0, st>>>(CopyFactory<System>{sys}, 1, 0.0, d_in.p, d_out.p);
    cuda_check(cudaGetLastError(), "generic dual rhs kernel launch");
    eng.check(sopot_b200_note_launches(eng.ctx(), 1));
    cuda_check(cudaMemcpyAsync(res.data(), d_out.p, W * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H dual derivative");
    cuda_check(cudaStreamSynchronize(st), "generic dual rhs kernel");
    std::vector<T> out(D);
    for (size_t d = 0; d < D; ++d) {
        std::array<double, K> part{};
        for (size_t k = 0; k < K; ++k) part[k] = res[d * (K + 1) + 1 + k];
        out[d] = T(res[d * (K + 1)], part);
    }
    return out;
}

template <typename System>
EnsembleResult solve_one(Engine& eng, const System& sys, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                         const EnsembleOptions& o) {
    return integrate(eng, CopyFactory<System>{sys}, 1, n_steps, t0, dt, o, y0.data());
}

}  // namespace sopot::b200::generic

namespace sopot::b200 {

// what RK4Solver::solveEnsemble takes for a user-defined system: n instances, built on the device by `make(i)`
template <typename Factory, typename Controller = generic::NoController>
struct GenericEnsemble {
    size_t n;
    Factory make;
    Controller control{};
    std::vector<double> y0{};          // optional explicit initial states [D][n]; empty: the components' own
    EnsembleResult integrate(Engine& eng, size_t n_steps, double t0, double dt, const EnsembleOptions& o) const {
        return generic::integrate(eng, make, n, n_steps, t0, dt, o, y0.empty() ? nullptr : y0.data(), control);
    }
    // the same ensemble under a sampled-data controller (see generic::NoController for the functor contract)
    template <typename C>
    GenericEnsemble<Factory, C> withController(C c) const { return GenericEnsemble<Factory, C>{n, make, c, y0}; }
    GenericEnsemble withInitialStates(std::vector<double> states) const { GenericEnsemble e = *this; e.y0 = std::move(states); return e; }
};
template <typename Factory>
GenericEnsemble<Factory> ensembleOf(size_t n, Factory make) { return GenericEnsemble<Factory>{n, make}; }

}  // namespace sopot::b200
