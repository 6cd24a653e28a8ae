// control::linearizeCartNPendulumBatch<N> -- batched SystemLinearizer<2(N+1),1>::linearize (core/linearization.hpp:75-114)
// around CartNPendulum<N, Dual> for ANY number of links, with the partial derivatives spread over the lanes of a warp.
//
// Forward-mode AD evaluates the dynamics once in Dual<double, K> arithmetic, K = 2(N+1) + 1 seeds.  Every Dual
// operation treats its K partials independently and identically (core/dual.hpp:66-133: d_j of a product is
// u*dv_j + du_j*v, of a quotient (v*du_j - u*dv_j) * 1/v^2, ...), so partial j never meets partial k.  Here lane j of a
// group of K lanes therefore evaluates the SAME dynamics in Dual<double, 1> arithmetic with only seed j switched on:
// its value part repeats what every other lane computes, its single partial is column j of [A | B].  The plant code is
// the component's own SOPOT_HD derivatives() (cart_n_pendulum.hpp here, generic in the scalar), nothing is restated.
// No shuffles are needed for the arithmetic -- pivoting decisions depend on values only and agree across the lanes of a
// group -- the exchange happens at the end: the group's K columns go through shared memory so that the Jacobians leave
// as coalesced SoA rows A[(i*NS + j)*P + p], B[i*P + p] (the layout of sopot_b200_cartpole_linearize).
//   N = 1: K = 5 lanes per operating point, 6 points per warp;  N = 6: K = 15, 2 points per warp.
// Each lane carries 2 doubles per scalar instead of K + 1: the 7x7 elimination of the 6-link plant fits in registers,
// where a thread-per-point Dual<double,15> (16 doubles per scalar, 49 + 7 + ... scalars) would live in local memory.
// Bit-compatibility: with -fmad=false each partial undergoes exactly the reference's operations; libm sin/cos aside the
// result equals makeLinearizer<2(N+1),1> on the CPU.
#pragma once
#ifndef __CUDACC__
#error "cart_n_linearize.cuh needs nvcc"
#endif
#include <vector>
#include "../../b200/generic_ensemble.cuh"
#include "../../core/dual.hpp"
#include "cart_n_pendulum.hpp"

namespace sopot::control {

namespace detail {

template <size_t NL, int PB /*operating points per block*/>
__global__ void __launch_bounds__(PB * (2 * (NL + 1) + 1))
cartn_linearize_lanes_kernel(const double* __restrict__ mc, const double* __restrict__ m /*[NL][P] or [NL]*/,
                             const double* __restrict__ L, const double* __restrict__ g, const int per_point_params,
                             const double* __restrict__ x0 /*[NS][P]*/, const double* __restrict__ u0 /*[P]*/, const size_t P,
                             double* __restrict__ A /*[NS*NS][P]*/, double* __restrict__ B /*[NS][P]*/) {
    constexpr int NS = 2 * (NL + 1), K = NS + 1;
    using D = Dual<double, 1>;
    extern __shared__ double tile[];                    // [NS][K][PB]: row i, column j (j == NS: B), point q
    const int q = threadIdx.x / K, j = threadIdx.x % K; // point within the block, seed / column of this lane
    const size_t p = (size_t)blockIdx.x * PB + q;
    if (p < P) {
        const size_t pp = per_point_params ? p : 0, stride = per_point_params ? P : 1;
        std::array<double, NL> mm, LL;
        for (size_t a = 0; a < NL; ++a) { mm[a] = m[a * stride + pp]; LL[a] = L[a * stride + pp]; }
        const CartNPendulum<NL, D> plant(mc[pp], mm, LL, g[pp]);
        D xs[NS];
        for (int i = 0; i < NS; ++i) xs[i] = D(x0[(size_t)i * P + p], {i == j ? 1.0 : 0.0});      // linearization.hpp:86-89
        plant.setControlForce(D(u0[p], {j == NS ? 1.0 : 0.0}));                                    // :92-95
        const std::span<const D> local(xs, NS);
        const auto d = plant.derivatives(D(0.0), local, local, 0);
        for (int i = 0; i < NS; ++i) tile[(i * K + j) * PB + q] = d[i].derivative(0);
    }
    __syncthreads();
    // coalesced write-out: consecutive threads take consecutive points of one (i, j) row
    const size_t p0 = (size_t)blockIdx.x * PB;
    for (int idx = threadIdx.x; idx < NS * K * PB; idx += blockDim.x) {
        const int row = idx / PB, qq = idx % PB, i = row / K, jj = row % K;
        if (p0 + qq < P) {
            if (jj < NS) A[(size_t)(i * NS + jj) * P + p0 + qq] = tile[idx];
            else B[(size_t)i * P + p0 + qq] = tile[idx];
        }
    }
}

}  // namespace detail

template <size_t NLinks>
struct CartNJacobians {
    static constexpr size_t NS = 2 * (NLinks + 1);
    size_t P = 0;
    std::vector<double> A, B;                  // [NS*NS][P], [NS][P]
    double a(size_t i, size_t j, size_t p) const { return A[(i * NS + j) * P + p]; }
    double b(size_t i, size_t p) const { return B[i * P + p]; }
};

// Plant parameters: one value set for all points (mc, m[N], L[N], g given as N-vectors / scalars) or one per point
// (mc, g of size P; m, L of size N*P laid out [N][P]).
template <size_t NLinks>
CartNJacobians<NLinks> linearizeCartNPendulumBatch(b200::Engine& eng, const std::vector<double>& mc, const std::vector<double>& m,
                                                   const std::vector<double>& L, const std::vector<double>& g,
                                                   const std::vector<double>& x0 /*[NS][P]*/, const std::vector<double>& u0 /*[P]*/) {
    constexpr size_t NS = 2 * (NLinks + 1), K = NS + 1;
    constexpr int PB = NLinks <= 2 ? 32 : 16;          // threads per block: 160 (N=1), 224 (N=2), 240 (N=6)
    const size_t P = u0.size();
    if (x0.size() != NS * P) throw std::invalid_argument("x0 must hold 2(N+1)*P values");
    const bool per_point = mc.size() == P && P != 1;
    if (!per_point && !(mc.size() == 1 && g.size() == 1 && m.size() == NLinks && L.size() == NLinks))
        throw std::invalid_argument("plant parameters: one set (mc, g scalars; m, L of N values) or one per point");
    if (per_point && !(g.size() == P && m.size() == NLinks * P && L.size() == NLinks * P))
        throw std::invalid_argument("per-point plant parameters need mc, g of P and m, L of N*P values");
    CartNJacobians<NLinks> r;
    r.P = P; r.A.resize(NS * NS * P); r.B.resize(NS * P);
    if (P == 0) return r;
    using b200::generic::cuda_check;
    using b200::generic::DeviceBuffer;
    cudaStream_t st = b200::generic::stream_of(eng);
    DeviceBuffer<double> d_mc(mc.size()), d_m(m.size()), d_L(L.size()), d_g(g.size()), d_x(x0.size()), d_u(P), d_A(r.A.size()), d_B(r.B.size());
    auto up = [&](DeviceBuffer<double>& d, const std::vector<double>& h) {
        cuda_check(cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, st), "H2D");
    };
    up(d_mc, mc); up(d_m, m); up(d_L, L); up(d_g, g); up(d_x, x0); up(d_u, u0);
    auto kern = detail::cartn_linearize_lanes_kernel<NLinks, PB>;
    constexpr size_t smem = NS * K * PB * sizeof(double);
    if (smem > 48 * 1024) cuda_check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attribute");
    kern<<<(unsigned)((P + PB - 1) / PB), PB * K, smem, st>>>(d_mc.p, d_m.p, d_L.p, d_g.p, per_point ? 1 : 0, d_x.p, d_u.p, P, d_A.p, d_B.p);
    cuda_check(cudaGetLastError(), "lane-parallel linearization launch");
    eng.check(sopot_b200_note_launches(eng.ctx(), 1));
    cuda_check(cudaMemcpyAsync(r.A.data(), d_A.p, r.A.size() * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H A");
    cuda_check(cudaMemcpyAsync(r.B.data(), d_B.p, r.B.size() * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H B");
    cuda_check(cudaStreamSynchronize(st), "lane-parallel linearization");
    return r;
}

// ---- closed loop for the N-link plant ------------------------------------------------------------------------------------
// Device factory (per-instance plant parameters from arrays) and the state-feedback controller of
// state_feedback_controller.hpp:88-104 as a generic-kernel controller: u = -K (x - x_ref), clamped when saturate is set,
// evaluated from the state at the start of every RK4 step and held over its four stages.
template <size_t NL>
struct CartNFactory {
    const double* mc;      // [n] device
    const double* m;       // [NL][n]
    const double* L;       // [NL][n]
    const double* g;       // [n]
    size_t n;
    SOPOT_HD TypedODESystem<double, CartNPendulum<NL, double>> operator()(size_t i) const {
        std::array<double, NL> mm{}, LL{};
        for (size_t a = 0; a < NL; ++a) { mm[a] = m[a * n + i]; LL[a] = L[a * n + i]; }
        return makeTypedODESystem<double>(CartNPendulum<NL, double>(mc[i], mm, LL, g[i]));
    }
};
template <size_t NL>
struct CartNStateFeedback {
    static constexpr size_t NS = 2 * (NL + 1);
    const double* K;       // [NS][n] device (e.g. straight from sopot_b200_lqr_gain)
    const double* x_ref;   // [NS][n] or nullptr
    size_t n;
    double u_min, u_max;
    int saturate;
    template <typename System>
    SOPOT_HD void operator()(size_t i, double, const double (&y)[NS], System& sys) const {
        double u = 0.0;
        for (size_t j = 0; j < NS; ++j) u -= K[j * n + i] * (y[j] - (x_ref ? x_ref[j * n + i] : 0.0));
        if (saturate) u = u < u_min ? u_min : (u_max < u ? u_max : u);
        sys.template getComponent<0>().setControlForce(u);
    }
};

}  // namespace sopot::control
