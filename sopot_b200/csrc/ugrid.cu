// ugrid.cu -- the "unified" 6-state mass-spring grid: physics/unified/components.hpp
//   PointMass2D   state [x, y, vx, vy, theta, omega], I = 1/2 m r^2, derivative {vx, vy, Fx/m, Fy/m, omega, tau/I}   (:52-138)
//   Spring2D      attaches at radius r on each mass at (theta + attachment_angle); force along the line between the two
//                 attachment points, damping on their relative velocity (which includes omega x r), optional steep
//                 repulsion, torque r x F on both masses                                                              (:206-291)
// driven the way ValidatedSystem::computeDerivatives does (physics/unified/validated_system.hpp:204-301; this is the system
// behind the reference's Grid2DSimulator, wasm/wasm_grid2d.cpp:223): springs are visited in node order -- all horizontal
// springs row-major, then all vertical ones (constexpr_topology.hpp:368-414) -- and each adds its force / torque to both
// end masses, so a mass receives: left spring (as mass 1), right spring (as mass 0), upper spring (as mass 1), lower
// spring (as mass 0).  The kernels below are the gather form of that: one thread per mass evaluates its incident springs
// in exactly this order.  State is AoS [rows][cols][6] (48 B per mass), the reference's node order.
// One RK4 step = four stage kernels with the running accumulator of ensemble.cuh (1.5x the traffic of the 4-state grid).
// Multi-GPU: contiguous slabs of rows per rank (params row_begin / rows_local).  Default: one exchange of H boundary rows per
// step into H ghost rows at both ends of every buffer, the stages evaluated on shrinking row ranges (see
// sopot_b200_ugrid_integrate); for slabs thinner than H rows: one ghost row above and below each stencil input, refreshed
// from the neighbouring ranks with ncclSend/ncclRecv after the kernel that produced the buffer.
#include <cmath>
#include <cstdlib>
#include "common.cuh"
#include "nccl_api.h"
#include "trig.cuh"

namespace {

struct UGridDev {
    size_t rows, cols;              // the GLOBAL grid
    long row_begin;                 // global row of local row 0 (negative for a buffer that starts with ghost rows above row 0)
    size_t rows_local;              // rows [row_begin, row_begin + rows_local) are addressable on this rank
    double mass, radius, k, L0, c, min_distance, krep, inertia, inv_mass, inv_inertia;
    double ha0, ha1, va0, va1;      // attachment angles: horizontal spring on its left / right mass, vertical on upper / lower
    double cha0, sha0, cha1, sha1, cva0, sva0, cva1, sva1;      // their cos / sin (contract mode: addition theorem)
    // triangle topology (makeTriangleGridTopology): the two diagonals of every cell
    int topo;
    double Ld, da0, da1, aa0, aa1;                              // diagonal rest length; main- / anti-diagonal attachment angles
    double cda0, sda0, cda1, sda1, caa0, saa0, caa1, saa1;
};

struct Body { double x, y, vx, vy, th, om; double s, c; };      // s, c = sin / cos(th), filled in contract mode only
struct UIn { const double* slab; const double* top; const double* bot; };     // local rows + ghost row above / below

__device__ __forceinline__ Body load_body(const double* __restrict__ s, size_t idx) {
    const double2* p = reinterpret_cast<const double2*>(s + 6 * idx);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    return Body{a.x, a.y, b.x, b.y, c.x, c.y, 0.0, 0.0};
}
// contract mode: sin/cos of the body angle ONCE per body; the four attachment directions follow by the addition theorem
template <bool S>
__device__ __forceinline__ Body load_body_t(const double* __restrict__ s, size_t idx) {
    Body b = load_body(s, idx);
    if constexpr (!S) sbtrig::sincos(b.th, b.s, b.c);
    return b;
}
__device__ __forceinline__ void store6(double* p, const double (&v)[6]) {
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(v[0], v[1]); q[1] = make_double2(v[2], v[3]); q[2] = make_double2(v[4], v[5]);
}

// Spring2D::computeForcesAndTorques for the spring (m0 -> m1); returns force and torque on the requested end
// contract mode: the same with sin/cos(theta + attachment angle) from the bodies' own sin/cos and the constant
// cos/sin of the attachment angle, the spring length and its reciprocal from one rsqrt seed (no IEEE division)
__device__ __forceinline__ void spring_ft_fast(const UGridDev& g, const Body& m0, const Body& m1, const double ca0, const double sa0,
                                               const double ca1, const double sa1, const int end, double& fx, double& fy, double& tq,
                                               const double L0) {
    const double r = g.radius;
    const double c0 = fma(m0.c, ca0, -(m0.s * sa0)), s0 = fma(m0.s, ca0, m0.c * sa0);
    const double c1 = fma(m1.c, ca1, -(m1.s * sa1)), s1 = fma(m1.s, ca1, m1.c * sa1);
    const double o0x = r * c0, o0y = r * s0, o1x = r * c1, o1y = r * s1;
    const double dx = (m1.x + o1x) - (m0.x + o0x), dy = (m1.y + o1y) - (m0.y + o0y);
    const double dvx = fma(-m1.om, o1y, m1.vx) - fma(-m0.om, o0y, m0.vx), dvy = fma(m1.om, o1x, m1.vy) - fma(m0.om, o0x, m0.vy);
    const double len2 = fma(dx, dx, dy * dy);
    double len, inv;
    if (len2 >= 1e-20) { sbtrig::sqrt_rsqrt_fast(len2, len, inv); }            // components.hpp:246-247: len_safe = max(len, 1e-10)
    else { len = sqrt(len2); inv = 1e10; }
    const double ux = dx * inv, uy = dy * inv;
    double F = fma(g.k, len - L0, g.c * fma(dvx, ux, dvy * uy));
    if (g.min_distance > 0.0 && len < g.min_distance) F = fma(-g.krep, fma(g.min_distance, inv, -1.0), F);
    if (end == 0) { fx = F * ux; fy = F * uy; tq = fma(o0x, fy, -(o0y * fx)); }
    else { fx = -F * ux; fy = -F * uy; tq = fma(o1x, fy, -(o1y * fx)); }
}

template <bool S>
__device__ __forceinline__ void spring_ft(const UGridDev& g, const Body& m0, const Body& m1, const double a0, const double a1,
                                          const int end, Real<S>& fx, Real<S>& fy, Real<S>& tq, const double L0) {
    using R = Real<S>;
    const R r(g.radius);
    R s0, c0, s1, c1;
    r_sincos(R(m0.th) + R(a0), s0, c0); r_sincos(R(m1.th) + R(a1), s1, c1);
    const R o0x = r * c0, o0y = r * s0, o1x = r * c1, o1y = r * s1;                    // offsets to the attachment points
    const R a0x = R(m0.x) + o0x, a0y = R(m0.y) + o0y, a1x = R(m1.x) + o1x, a1y = R(m1.y) + o1y;
    const R v0x = R(m0.vx) - R(m0.om) * o0y, v0y = R(m0.vy) + R(m0.om) * o0x;          // v + omega x r
    const R v1x = R(m1.vx) - R(m1.om) * o1y, v1y = R(m1.vy) + R(m1.om) * o1x;
    const R dx = a1x - a0x, dy = a1y - a0y;
    const R len = r_sqrt(dx * dx + dy * dy);
    const R len_safe = len.v < 1e-10 ? R(1e-10) : len;
    const R ux = dx / len_safe, uy = dy / len_safe;
    const R ext = len - R(L0);
    const R dvx = v1x - v0x, dvy = v1y - v0y;
    const R rel = dvx * ux + dvy * uy;
    R F = R(g.k) * ext + R(g.c) * rel;
    if (g.min_distance > 0.0 && len.v < g.min_distance)
        F = F + (-R(g.krep)) * (R(g.min_distance) / len_safe - R(1.0));
    if (end == 0) { fx = F * ux; fy = F * uy; tq = o0x * fy - o0y * fx; }
    else { fx = (-F) * ux; fy = (-F) * uy; tq = o1x * fy - o1y * fx; }
}

// body at local row lrow in [-1, rows_local] (ghost rows at the two ends), column col
template <bool S>
__device__ __forceinline__ Body body_at(const UGridDev& g, const UIn& in, const long lrow, const size_t col) {
    if (lrow < 0) return load_body_t<S>(in.top, col);
    if ((size_t)lrow >= g.rows_local) return load_body_t<S>(in.bot, col);
    return load_body_t<S>(in.slab, (size_t)lrow * g.cols + col);
}

// PointMass2D derivative of the mass at GLOBAL (grow, col) from its own body P and a neighbour accessor nb(drow, dcol):
// springs in the reference's visiting order -- left (as mass 1), right (as mass 0), upper (as mass 1), lower (as mass 0)
template <bool S, class Nb>
__device__ __forceinline__ void derivative_from(const UGridDev& g, const Body& P, const size_t grow, const size_t col, Nb nb,
                                                Real<S> (&d)[6]) {
    using R = Real<S>;
    R Fx(0.0), Fy(0.0), Tq(0.0), fx, fy, tq;
    const bool left = col > 0, right = col + 1 < g.cols, up = grow > 0, down = grow + 1 < g.rows;
    if constexpr (S) {
        if (left) { spring_ft<S>(g, nb(0, -1), P, g.ha0, g.ha1, 1, fx, fy, tq, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (right) { spring_ft<S>(g, P, nb(0, 1), g.ha0, g.ha1, 0, fx, fy, tq, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (up) { spring_ft<S>(g, nb(-1, 0), P, g.va0, g.va1, 1, fx, fy, tq, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (down) { spring_ft<S>(g, P, nb(1, 0), g.va0, g.va1, 0, fx, fy, tq, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (g.topo == 1) {
            // diagonals in spring-node order (cells row-major, main before anti): cell (r-1,c-1) main, this mass is its lower
            // right end; cell (r-1,c) anti, lower left end; cell (r,c-1) anti, upper right end; cell (r,c) main, upper left end
            if (up && left) { spring_ft<S>(g, nb(-1, -1), P, g.da0, g.da1, 1, fx, fy, tq, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (up && right) { spring_ft<S>(g, nb(-1, 1), P, g.aa0, g.aa1, 1, fx, fy, tq, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (down && left) { spring_ft<S>(g, P, nb(1, -1), g.aa0, g.aa1, 0, fx, fy, tq, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (down && right) { spring_ft<S>(g, P, nb(1, 1), g.da0, g.da1, 0, fx, fy, tq, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
        }
    } else {
        if (left) { spring_ft_fast(g, nb(0, -1), P, g.cha0, g.sha0, g.cha1, g.sha1, 1, fx.v, fy.v, tq.v, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (right) { spring_ft_fast(g, P, nb(0, 1), g.cha0, g.sha0, g.cha1, g.sha1, 0, fx.v, fy.v, tq.v, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (up) { spring_ft_fast(g, nb(-1, 0), P, g.cva0, g.sva0, g.cva1, g.sva1, 1, fx.v, fy.v, tq.v, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (down) { spring_ft_fast(g, P, nb(1, 0), g.cva0, g.sva0, g.cva1, g.sva1, 0, fx.v, fy.v, tq.v, g.L0); Fx += fx; Fy += fy; Tq += tq; }
        if (g.topo == 1) {
            if (up && left) { spring_ft_fast(g, nb(-1, -1), P, g.cda0, g.sda0, g.cda1, g.sda1, 1, fx.v, fy.v, tq.v, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (up && right) { spring_ft_fast(g, nb(-1, 1), P, g.caa0, g.saa0, g.caa1, g.saa1, 1, fx.v, fy.v, tq.v, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (down && left) { spring_ft_fast(g, P, nb(1, -1), g.caa0, g.saa0, g.caa1, g.saa1, 0, fx.v, fy.v, tq.v, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
            if (down && right) { spring_ft_fast(g, P, nb(1, 1), g.cda0, g.sda0, g.cda1, g.sda1, 0, fx.v, fy.v, tq.v, g.Ld); Fx += fx; Fy += fy; Tq += tq; }
        }
    }
    d[0] = R(P.vx); d[1] = R(P.vy);
    d[4] = R(P.om);
    d[2] = r_div_by<S>(Fx, g.mass, g.inv_mass); d[3] = r_div_by<S>(Fy, g.mass, g.inv_mass);
    d[5] = r_div_by<S>(Tq, g.inertia, g.inv_inertia);
}

template <bool S>
__device__ __forceinline__ void body_derivative(const UGridDev& g, const UIn& in, const size_t row /*local*/, const size_t col,
                                                Real<S> (&d)[6], Body& self) {
    const long lr = (long)row;
    self = body_at<S>(g, in, lr, col);
    derivative_from<S>(g, self, (size_t)(g.row_begin + (long)row), col,
                       [&](int dr, int dc) { return body_at<S>(g, in, lr + dr, (size_t)((long)col + dc)); }, d);
}

// STAGE 0: out <- f(in);  1..4: the RK4 stages with S-accumulator (see ensemble.cuh / grid2d.cu)
// what a stage does with the derivative d of the body `self` at element offset off: shared by the gather and marching kernels
template <bool S, int STAGE>
__device__ __forceinline__ void ustage_apply(const Body& self, const Real<S> (&d)[6], const size_t off, const double dt_,
                                             double* __restrict__ y, double* __restrict__ acc, double* __restrict__ out) {
    using R = Real<S>;
    double o[6];
    if constexpr (STAGE == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) o[j] = d[j].v;
        store6(out + off, o);
        return;
    }
    const R dt(dt_);
    if constexpr (STAGE == 5) {
        // symplectic Euler, wasm/wasm_grid2d.cpp:329-358: v += dt*a(x_n, v_n), then x += dt*v_new (likewise omega, theta)
        const R vx = R(self.vx) + dt * d[2], vy = R(self.vy) + dt * d[3], om = R(self.om) + dt * d[5];
        o[0] = (R(self.x) + dt * vx).v; o[1] = (R(self.y) + dt * vy).v; o[2] = vx.v; o[3] = vy.v;
        o[4] = (R(self.th) + dt * om).v; o[5] = om.v;
        store6(out + off, o);
        return;
    }
    if constexpr (STAGE == 6) {
        // velocity Verlet, first half (:375-398): x += dt*v + (0.5*dt*dt)*a_n; velocities wait for a_{n+1}; a_n is kept in acc
        const R hdt2 = R(0.5) * dt * dt;
        o[0] = (R(self.x) + (dt * R(self.vx) + hdt2 * d[2])).v; o[1] = (R(self.y) + (dt * R(self.vy) + hdt2 * d[3])).v;
        o[2] = self.vx; o[3] = self.vy;
        o[4] = (R(self.th) + (dt * R(self.om) + hdt2 * d[5])).v; o[5] = self.om;
        store6(out + off, o);
        const double a[6] = {d[2].v, d[3].v, d[5].v, 0.0, 0.0, 0.0};
        store6(acc + off, a);
        return;
    }
    if constexpr (STAGE == 7) {
        // velocity Verlet, second half (:400-410): a_{n+1} = a(x_{n+1}, v_n); v += (0.5*dt)*(a_n + a_{n+1})
        const R hdt = R(0.5) * dt;
        o[0] = self.x; o[1] = self.y; o[4] = self.th;
        o[2] = (R(self.vx) + hdt * (R(acc[off]) + d[2])).v; o[3] = (R(self.vy) + hdt * (R(acc[off + 1]) + d[3])).v;
        o[5] = (R(self.om) + hdt * (R(acc[off + 2]) + d[5])).v;
        store6(out + off, o);
        return;
    }
    R k[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) k[j] = d[j] * dt;
    R yv[6], sv[6];
    if constexpr (STAGE == 1) {
        yv[0] = R(self.x); yv[1] = R(self.y); yv[2] = R(self.vx); yv[3] = R(self.vy); yv[4] = R(self.th); yv[5] = R(self.om);
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) { yv[j] = R(y[off + j]); sv[j] = R(acc[off + j]); }
    }
    if constexpr (STAGE == 4) {
#pragma unroll
        for (int j = 0; j < 6; ++j) o[j] = (yv[j] + r_div_const(sv[j] + k[j], 6.0)).v;
        store6(y + off, o);
    } else {
        double a[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            a[j] = (STAGE == 1 ? k[j] : sv[j] + 2.0 * k[j]).v;
            o[j] = (STAGE == 3 ? yv[j] + k[j] : yv[j] + 0.5 * k[j]).v;
        }
        store6(acc + off, a);
        store6(out + off, o);
    }
}

template <bool S, int STAGE>
__global__ void __launch_bounds__(256)
ugrid_stage(const UGridDev g, const UIn in, const double dt_, double* __restrict__ y, double* __restrict__ acc,
            double* __restrict__ out, const size_t r_lo, const size_t r_hi) {
    using R = Real<S>;
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x, row = r_lo + (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= g.cols || row >= r_hi) return;
    R d[6];
    Body self;
    body_derivative<S>(g, in, row, col, d, self);
    ustage_apply<S, STAGE>(self, d, (row * g.cols + col) * 6, dt_, y, acc, out);
}

// ---- one stage, marching form (quad topology): every spring ONCE ---------------------------------------------------------
// The gather kernel evaluates each spring at both of its ends (4 spring evaluations and, in strict arithmetic, 8 libm sincos
// per mass and stage): that made the strict RK4 step and the one- / two-evaluation steps of symplectic Euler and velocity
// Verlet FP64-bound at half of the HBM roofline.  Here a warp owns a strip of 32 columns (30 output columns + one halo lane
// on each side) and marches down a chunk of rows: per tick the lane evaluates the RIGHT spring of its body (neighbour by warp
// shuffle; the force and torque on the neighbour's end go back by a second shuffle) and the DOWN spring (the end-1 force and
// torque are carried in registers to the next tick, where they belong to the row below) -- two spring evaluations and four
// sincos per mass and stage, accumulated in the reference's visiting order left (as mass 1), right (as mass 0), upper (as
// mass 1), lower (as mass 0).  Force on end 1 = exact negation of the force on end 0, torque on end 1 from the negated force
// with the reference's own expression: the bits of the gather form (hence of ValidatedSystem) in strict mode.
template <bool S>
__device__ __forceinline__ void spring_both(const UGridDev& g, const Body& m0, const Body& m1, const double a0, const double a1,
                                            const double ca0, const double sa0, const double ca1, const double sa1, const double L0,
                                            Real<S>& fx, Real<S>& fy, Real<S>& tq0, Real<S>& tq1) {
    using R = Real<S>;
    if constexpr (S) {
        const R r(g.radius);
        R s0, c0, s1, c1;
        r_sincos(R(m0.th) + R(a0), s0, c0); r_sincos(R(m1.th) + R(a1), s1, c1);
        const R o0x = r * c0, o0y = r * s0, o1x = r * c1, o1y = r * s1;
        const R a0x = R(m0.x) + o0x, a0y = R(m0.y) + o0y, a1x = R(m1.x) + o1x, a1y = R(m1.y) + o1y;
        const R v0x = R(m0.vx) - R(m0.om) * o0y, v0y = R(m0.vy) + R(m0.om) * o0x;
        const R v1x = R(m1.vx) - R(m1.om) * o1y, v1y = R(m1.vy) + R(m1.om) * o1x;
        const R dx = a1x - a0x, dy = a1y - a0y;
        const R len = r_sqrt(dx * dx + dy * dy);
        const R len_safe = len.v < 1e-10 ? R(1e-10) : len;
        const R ux = dx / len_safe, uy = dy / len_safe;
        const R ext = len - R(L0);
        const R dvx = v1x - v0x, dvy = v1y - v0y;
        const R rel = dvx * ux + dvy * uy;
        R F = R(g.k) * ext + R(g.c) * rel;
        if (g.min_distance > 0.0 && len.v < g.min_distance)
            F = F + (-R(g.krep)) * (R(g.min_distance) / len_safe - R(1.0));
        fx = F * ux; fy = F * uy; tq0 = o0x * fy - o0y * fx;
        const R gx = -fx, gy = -fy;                         // (-F) * ux == -(F * ux) exactly
        tq1 = o1x * gy - o1y * gx;
    } else {
        const double r = g.radius;
        const double c0 = fma(m0.c, ca0, -(m0.s * sa0)), s0 = fma(m0.s, ca0, m0.c * sa0);
        const double c1 = fma(m1.c, ca1, -(m1.s * sa1)), s1 = fma(m1.s, ca1, m1.c * sa1);
        const double o0x = r * c0, o0y = r * s0, o1x = r * c1, o1y = r * s1;
        const double dx = (m1.x + o1x) - (m0.x + o0x), dy = (m1.y + o1y) - (m0.y + o0y);
        const double dvx = fma(-m1.om, o1y, m1.vx) - fma(-m0.om, o0y, m0.vx), dvy = fma(m1.om, o1x, m1.vy) - fma(m0.om, o0x, m0.vy);
        const double len2 = fma(dx, dx, dy * dy);
        double len, inv;
        if (len2 >= 1e-20) { sbtrig::sqrt_rsqrt_fast(len2, len, inv); }
        else { len = sqrt(len2); inv = 1e10; }
        const double ux = dx * inv, uy = dy * inv;
        double F = fma(g.k, len - L0, g.c * fma(dvx, ux, dvy * uy));
        if (g.min_distance > 0.0 && len < g.min_distance) F = fma(-g.krep, fma(g.min_distance, inv, -1.0), F);
        fx.v = F * ux; fy.v = F * uy; tq0.v = fma(o0x, fy.v, -(o0y * fx.v));
        tq1.v = fma(o1x, -fy.v, o1y * fx.v);
    }
}

__device__ __forceinline__ Body shfl_body_down(const Body& b) {
    constexpr unsigned M = 0xffffffffu;
    return Body{__shfl_down_sync(M, b.x, 1), __shfl_down_sync(M, b.y, 1), __shfl_down_sync(M, b.vx, 1), __shfl_down_sync(M, b.vy, 1),
                __shfl_down_sync(M, b.th, 1), __shfl_down_sync(M, b.om, 1), __shfl_down_sync(M, b.s, 1), __shfl_down_sync(M, b.c, 1)};
}

constexpr int kUMarchOwn = 30, kUMarchWarps = 4;
template <bool S, int STAGE>
__global__ void __launch_bounds__(32 * kUMarchWarps)
ugrid_stage_march(const UGridDev g, const UIn in, const double dt_, double* __restrict__ y, double* __restrict__ acc,
                  double* __restrict__ out, const size_t r_lo, const size_t r_hi, const unsigned rows_per_chunk) {
    using R = Real<S>;
    constexpr unsigned M = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long strip = (long)blockIdx.x * kUMarchWarps + (threadIdx.x >> 5);
    if (strip * kUMarchOwn >= (long)g.cols) return;
    const long col = strip * kUMarchOwn - 1 + lane;
    const bool col_ok = col >= 0 && col < (long)g.cols;
    const bool owner = lane >= 1 && lane <= kUMarchOwn && col_ok;
    const bool has_l = col > 0, has_r = col + 1 < (long)g.cols;
    const long lo = (long)r_lo + (long)blockIdx.y * rows_per_chunk;
    long hi = lo + rows_per_chunk;
    if (hi > (long)r_hi) hi = (long)r_hi;
    if (lo >= hi) return;
    // body of local row lr (ghost rows at -1 / rows_local); a cell outside the global grid is a placeholder on its own
    // lattice position (its springs are never used, and never of zero length)
    const auto load = [&](const long lr) -> Body {
        const long gr = g.row_begin + lr;
        if (!col_ok || gr < 0 || gr >= (long)g.rows) return Body{(double)col, (double)gr, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0};
        return body_at<S>(g, in, lr, (size_t)col);
    };
    Body cur = load(lo - 1), nxt = load(lo);
    R ufx(0.0), ufy(0.0), utq(0.0);                       // force / torque of the spring from the row above on THIS row's body
    for (long q = lo - 1; q < hi; ++q) {                   // the first tick only produces the up spring of row lo
        const Body pre = load(q + 2 <= hi ? q + 2 : hi);
        Body rt = shfl_body_down(cur);
        if (lane == 31) rt.x = cur.x + 1.0;                // no right neighbour in the warp: a unit-length pseudo-spring, discarded
        R rfx, rfy, rtq0, rtq1, dfx, dfy, dtq0, dtq1;
        spring_both<S>(g, cur, rt, g.ha0, g.ha1, g.cha0, g.sha0, g.cha1, g.sha1, g.L0, rfx, rfy, rtq0, rtq1);
        spring_both<S>(g, cur, nxt, g.va0, g.va1, g.cva0, g.sva0, g.cva1, g.sva1, g.L0, dfx, dfy, dtq0, dtq1);
        // what the left lane's right spring does to this body (its end 1): the negated force, its own torque
        const R lfx(-__shfl_up_sync(M, rfx.v, 1)), lfy(-__shfl_up_sync(M, rfy.v, 1)), ltq(__shfl_up_sync(M, rtq1.v, 1));
        const long gr = g.row_begin + q;
        if (q >= lo && owner) {
            R Fx(0.0), Fy(0.0), Tq(0.0);
            if (has_l) { Fx += lfx; Fy += lfy; Tq += ltq; }
            if (has_r) { Fx += rfx; Fy += rfy; Tq += rtq0; }
            if (gr > 0) { Fx += ufx; Fy += ufy; Tq += utq; }
            if (gr + 1 < (long)g.rows) { Fx += dfx; Fy += dfy; Tq += dtq0; }
            R d[6];
            d[0] = R(cur.vx); d[1] = R(cur.vy); d[4] = R(cur.om);
            d[2] = r_div_by<S>(Fx, g.mass, g.inv_mass); d[3] = r_div_by<S>(Fy, g.mass, g.inv_mass);
            d[5] = r_div_by<S>(Tq, g.inertia, g.inv_inertia);
            ustage_apply<S, STAGE>(cur, d, ((size_t)q * g.cols + (size_t)col) * 6, dt_, y, acc, out);
        }
        ufx = -dfx; ufy = -dfy; utq = dtq1;                // the down spring acts on the next row's body as its end 1
        cur = nxt; nxt = pre;
    }
}

// ---- two RK4 stages per kernel (single rank) ----------------------------------------------------------------------------
// The per-stage schedule moves 816 B per mass and step through HBM (every stage re-reads y and the accumulator).  Here a
// CTA loads a 32 x 8 tile with a 2-cell halo into shared memory (SoA: x, y, vx, vy, theta, omega and, in contract mode,
// sin/cos theta computed once per body), evaluates the first stage of the pair on the tile grown by one cell, keeps the
// intermediate state in a second shared tile, and evaluates the second stage on the interior:
//   FIRST  (stages 1+2): in = y            -> acc = k1 + 2 k2,  out = y + k2/2                       144 B per mass
//   !FIRST (stages 3+4): in = y + k2/2, y, acc -> out = y + ((acc + 2 k3) + k4)/6  (a fresh buffer)   192 B per mass
// 336 B per mass and step; the ring of the grown tile is evaluated redundantly (1.33x on one of the two stages).  Every
// operation and its order equal the per-stage kernels', so the strict results are bit-identical to that schedule.
// MEASURED (4096^2, B200): 2.43 ms per step in contract mode against 2.30 ms for the per-stage schedule, 6.8 against 3.9 ms in
// strict mode -- the per-stage kernels already run at 0.9 of the HBM roofline, while this kernel stalls on its two barriers,
// the exposed tile load (long_scoreboard) and the three warps that evaluate the ring; FP64 pipe 45 %, 24 warps per SM
// (profiles/r1_ugrid_stage_pairs.txt).  Hence opt-in (SOPOT_B200_FLAG_UGRID_STAGE_PAIRS); a winning version needs every
// spring evaluated once and the ring spread over all warps, as grid_rk4_fused does for the 4-state grid.
#ifndef SB_UGRID_PAIR_MINB
#define SB_UGRID_PAIR_MINB 3
#endif
constexpr int kPairTX = 32, kPairTY = 8, kPairW = kPairTX + 4, kPairH = kPairTY + 4, kPairCells = kPairW * kPairH;

struct PairTile {
    double* p;                                                    // [8][kPairCells]
    __device__ __forceinline__ Body get(int c) const {
        return Body{p[c], p[kPairCells + c], p[2 * kPairCells + c], p[3 * kPairCells + c], p[4 * kPairCells + c], p[5 * kPairCells + c],
                    p[6 * kPairCells + c], p[7 * kPairCells + c]};
    }
    __device__ __forceinline__ void put(int c, const Body& b) const {
        p[c] = b.x; p[kPairCells + c] = b.y; p[2 * kPairCells + c] = b.vx; p[3 * kPairCells + c] = b.vy; p[4 * kPairCells + c] = b.th;
        p[5 * kPairCells + c] = b.om; p[6 * kPairCells + c] = b.s; p[7 * kPairCells + c] = b.c;
    }
};

template <bool S, bool FIRST>
__global__ void __launch_bounds__(kPairTX * kPairTY, SB_UGRID_PAIR_MINB)
ugrid_pair(const UGridDev g, const double* __restrict__ in, const double* __restrict__ ybase, const double dt_,
           double* __restrict__ acc, double* __restrict__ out) {
    using R = Real<S>;
    extern __shared__ double s_pair[];
    const PairTile T1{s_pair}, T2{s_pair + 8 * kPairCells};
    const int tid = threadIdx.x;
    const long row0 = (long)blockIdx.y * kPairTY - 2, col0 = (long)blockIdx.x * kPairTX - 2;
    const auto in_grid = [&](long r, long c) { return r >= 0 && c >= 0 && (size_t)r < g.rows && (size_t)c < g.cols; };
    for (int c = tid; c < kPairCells; c += kPairTX * kPairTY) {
        const long r = row0 + c / kPairW, cc = col0 + c % kPairW;
        if (in_grid(r, cc)) T1.put(c, load_body_t<S>(in, (size_t)r * g.cols + (size_t)cc));
    }
    __syncthreads();
    const R dt(dt_);
    // first stage of the pair on one cell of the grown tile: k = dt f(T1), intermediate state -> T2
    const auto stage_a = [&](const int c, R (&k)[6], Body& base) {
        const long r = row0 + c / kPairW, cc = col0 + c % kPairW;
        const Body self = T1.get(c);
        R d[6];
        derivative_from<S>(g, self, (size_t)r, (size_t)cc, [&](int dr, int dc) { return T1.get(c + dr * kPairW + dc); }, d);
#pragma unroll
        for (int j = 0; j < 6; ++j) k[j] = d[j] * dt;
        if constexpr (FIRST) base = self;                                  // stage 2 reads y + k1/2
        else base = load_body(ybase, (size_t)r * g.cols + (size_t)cc);     // stage 4 reads y + k3
        const R h(FIRST ? 0.5 : 1.0);
        Body nb = base;
        if constexpr (FIRST) {
            nb.x = (R(base.x) + h * k[0]).v; nb.y = (R(base.y) + h * k[1]).v; nb.vx = (R(base.vx) + h * k[2]).v;
            nb.vy = (R(base.vy) + h * k[3]).v; nb.th = (R(base.th) + h * k[4]).v; nb.om = (R(base.om) + h * k[5]).v;
        } else {
            nb.x = (R(base.x) + k[0]).v; nb.y = (R(base.y) + k[1]).v; nb.vx = (R(base.vx) + k[2]).v;
            nb.vy = (R(base.vy) + k[3]).v; nb.th = (R(base.th) + k[4]).v; nb.om = (R(base.om) + k[5]).v;
        }
        if constexpr (!S) sbtrig::sincos(nb.th, nb.s, nb.c);
        T2.put(c, nb);
    };
    // the thread's own interior cell (its k stays in registers), then the ring of the grown tile spread over the threads
    const int own = (2 + tid / kPairTX) * kPairW + 2 + tid % kPairTX;
    const long orow = row0 + own / kPairW, ocol = col0 + own % kPairW;
    const bool own_valid = in_grid(orow, ocol);
    R ka[6];
    Body base;
    if (own_valid) stage_a(own, ka, base);
    constexpr int kRing = (kPairTX + 2) * (kPairTY + 2) - kPairTX * kPairTY;   // 84 cells
    if (tid < kRing) {
        // ring enumeration: top row, bottom row (kPairTX + 2 each), then the left and right columns
        int rr, rc;
        if (tid < kPairTX + 2) { rr = 1; rc = 1 + tid; }
        else if (tid < 2 * (kPairTX + 2)) { rr = kPairTY + 2; rc = 1 + tid - (kPairTX + 2); }
        else { const int t = tid - 2 * (kPairTX + 2); rr = 2 + t / 2; rc = (t & 1) ? kPairTX + 2 : 1; }
        const int c = rr * kPairW + rc;
        if (in_grid(row0 + rr, col0 + rc)) { R kr[6]; Body b; stage_a(c, kr, b); }
    }
    __syncthreads();
    if (!own_valid) return;
    // second stage of the pair on the interior
    R d[6];
    derivative_from<S>(g, T2.get(own), (size_t)orow, (size_t)ocol, [&](int dr, int dc) { return T2.get(own + dr * kPairW + dc); }, d);
    R kb[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) kb[j] = d[j] * dt;
    const size_t off = ((size_t)orow * g.cols + (size_t)ocol) * 6;
    const R yb[6] = {R(base.x), R(base.y), R(base.vx), R(base.vy), R(base.th), R(base.om)};
    double o[6];
    if constexpr (FIRST) {
        double a[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { a[j] = (ka[j] + 2.0 * kb[j]).v; o[j] = (yb[j] + 0.5 * kb[j]).v; }
        store6(acc + off, a);
        store6(out + off, o);
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const R s3 = R(acc[off + j]) + 2.0 * ka[j];
            o[j] = (yb[j] + r_div_const(s3 + kb[j], 6.0)).v;
        }
        store6(out + off, o);
    }
}

template <bool FIRST>
int launch_upair(sopot_b200_ctx* ctx, bool strict, const UGridDev& g, const double* in, const double* ybase, double dt, double* acc,
                 double* out) {
    const dim3 grid((unsigned)((g.cols + kPairTX - 1) / kPairTX), (unsigned)((g.rows + kPairTY - 1) / kPairTY));
    constexpr size_t smem = 2 * 8 * kPairCells * sizeof(double);
    if (strict) {
        SB_CUDA(ctx, cudaFuncSetAttribute(ugrid_pair<true, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ugrid_pair<true, FIRST><<<grid, kPairTX * kPairTY, smem, ctx->stream>>>(g, in, ybase, dt, acc, out);
    } else {
        SB_CUDA(ctx, cudaFuncSetAttribute(ugrid_pair<false, FIRST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ugrid_pair<false, FIRST><<<grid, kPairTX * kPairTY, smem, ctx->stream>>>(g, in, ybase, dt, acc, out);
    }
    ctx->launches++;
    return SOPOT_B200_OK;
}

__global__ void ugrid_init_kernel(const size_t rows, const size_t cols, const size_t row_begin, const double spacing,
                                  double* __restrict__ state) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x, row = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= cols || row >= rows) return;
    const double v[6] = {(double)col * spacing, (double)(row_begin + row) * spacing, 0.0, 0.0, 0.0, 0.0};     // wasm_grid2d.cpp:427-433
    store6(state + (row * cols + col) * 6, v);
}

int check_ugrid(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* p, UGridDev& g) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p) return sb_fail(ctx, SOPOT_B200_EINVAL, "ugrid params NULL");
    if (p->rows < 1 || p->cols < 1 || p->rows * p->cols < 2) return sb_fail(ctx, SOPOT_B200_EINVAL, "the grid needs at least two masses");
    if (!(p->mass > 0.0) || !(p->radius > 0.0))
        return sb_fail(ctx, SOPOT_B200_EINVAL, "mass and radius must be positive (I = m r^2 / 2 divides the torque)");
    g.row_begin = p->rows_local ? (long)p->row_begin : 0;
    g.rows_local = p->rows_local ? p->rows_local : p->rows;
    if ((size_t)g.row_begin + g.rows_local > p->rows) return sb_fail(ctx, SOPOT_B200_EINVAL, "slab [row_begin, row_begin + rows_local) exceeds the grid");
    if (ctx->nranks == 1 && g.rows_local != p->rows) return sb_fail(ctx, SOPOT_B200_EINVAL, "a single-rank ctx must own the whole grid");
    if (int rc = sb_check_slab(ctx, p->rows, (size_t)g.row_begin, g.rows_local)) return rc;
    g.rows = p->rows; g.cols = p->cols; g.mass = p->mass; g.radius = p->radius; g.k = p->stiffness; g.L0 = p->rest_length;
    g.c = p->damping; g.min_distance = p->min_distance;
    g.krep = p->repulsion_stiffness > 0.0 ? p->repulsion_stiffness : 10.0 * p->stiffness;       // components.hpp:186
    g.inertia = 0.5 * p->mass * p->radius * p->radius;                                          // :98
    g.ha0 = p->h_attach0; g.ha1 = p->h_attach1; g.va0 = p->v_attach0; g.va1 = p->v_attach1;
    g.inv_mass = 1.0 / g.mass; g.inv_inertia = 1.0 / g.inertia;
    g.cha0 = std::cos(g.ha0); g.sha0 = std::sin(g.ha0); g.cha1 = std::cos(g.ha1); g.sha1 = std::sin(g.ha1);
    g.cva0 = std::cos(g.va0); g.sva0 = std::sin(g.va0); g.cva1 = std::cos(g.va1); g.sva1 = std::sin(g.va1);
    if (p->topology != 0 && p->topology != 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "ugrid topology must be 0 (quad) or 1 (triangle)");
    g.topo = p->topology; g.Ld = p->diag_rest_length;
    g.da0 = p->d_attach0; g.da1 = p->d_attach1; g.aa0 = p->a_attach0; g.aa1 = p->a_attach1;
    g.cda0 = std::cos(g.da0); g.sda0 = std::sin(g.da0); g.cda1 = std::cos(g.da1); g.sda1 = std::sin(g.da1);
    g.caa0 = std::cos(g.aa0); g.saa0 = std::sin(g.aa0); g.caa1 = std::cos(g.aa1); g.saa1 = std::sin(g.aa1);
    return SOPOT_B200_OK;
}

// rows [r_lo, r_hi) of the local buffer (default: all of them)
template <int STAGE>
void launch_ustage(sopot_b200_ctx* ctx, bool strict, const UGridDev& g, const UIn& in, double dt, double* y, double* acc, double* out,
                   size_t r_lo = 0, size_t r_hi = (size_t)-1) {
    if (r_hi == (size_t)-1) r_hi = g.rows_local;
    if (r_hi <= r_lo) return;
    // Quad topology: the marching kernel (every spring once) wherever the gather kernel is FP64-bound -- strict arithmetic
    // (4096^2 RK4 step 3.96 -> 2.83 ms, symplectic Euler 0.89 -> 0.57, velocity Verlet 1.86 -> 1.24) and the contract-mode Euler /
    // Verlet stages (0.49 -> 0.37, 1.06 -> 1.02).  The contract-mode RK4 stages stay on the gather kernel: they are HBM-bound
    // there (0.88 of peak) and the marching form, with fewer loads in flight per SM, measured 3 % slower (2.44 against 2.38 ms).
    // SOPOT_B200_UGRID_GATHER=1 keeps the gather kernel everywhere, SOPOT_B200_UGRID_MARCH=1 marches everywhere.
    static const bool gather_only = std::getenv("SOPOT_B200_UGRID_GATHER") != nullptr;
    static const bool march_all = std::getenv("SOPOT_B200_UGRID_MARCH") != nullptr;
    static const unsigned env_rows = std::getenv("SOPOT_B200_UGRID_MARCH_ROWS") ? (unsigned)std::atoi(std::getenv("SOPOT_B200_UGRID_MARCH_ROWS")) : 0u;
    // ... and only where it can fill the GPU: a warp marches a 30-column strip in chunks of 16 rows, so a 512^2 grid is 160 CTAs
    // and the gather kernel (one thread per mass) is twice as fast there -- strict RK4 0.088 against 0.163 ms per step, contract
    // Euler 0.013 against 0.020; from 1024^2 (576 CTAs) the marching kernel wins (strict RK4 0.232 against 0.282).  The
    // contract-mode Verlet stages gain from it only on very large grids (4096^2: 1.02 against 1.06; 2048^2: 0.292 against 0.272).
    // Decided from the GLOBAL grid size, so every rank of a slab decomposition and the single-domain run pick the same kernel
    // (profiles/r2_ugrid_small_sizes.txt).
    const size_t march_ctas = (((size_t)g.cols + kUMarchOwn - 1) / kUMarchOwn + kUMarchWarps - 1) / kUMarchWarps * (((size_t)g.rows + 15) / 16);
    const bool verlet_contract = !strict && (STAGE == 6 || STAGE == 7);
    const bool big_enough = march_ctas >= (verlet_contract ? 8192u : 2u * (unsigned)ctx->sm_count);
    const bool force_march = march_all || (ctx->ugrid_kernel & SOPOT_B200_FLAG_UGRID_MARCH);
    const bool force_gather = gather_only || (ctx->ugrid_kernel & SOPOT_B200_FLAG_UGRID_GATHER);
    if (g.topo == 0 && STAGE != 0 && !force_gather && (force_march || (big_enough && (strict || STAGE >= 5)))) {
        // 16 rows per chunk (one extra tick each): 16 -> 2.83 ms, 32 -> 2.87, 64 -> 3.08, 128 -> 3.23 (strict RK4 step)
        const unsigned strips = (unsigned)((g.cols + kUMarchOwn - 1) / kUMarchOwn), chunk = env_rows ? env_rows : 16u;
        const dim3 mgrid((strips + kUMarchWarps - 1) / kUMarchWarps, (unsigned)((r_hi - r_lo + chunk - 1) / chunk));
        if (strict) ugrid_stage_march<true, STAGE><<<mgrid, 32 * kUMarchWarps, 0, ctx->stream>>>(g, in, dt, y, acc, out, r_lo, r_hi, chunk);
        else ugrid_stage_march<false, STAGE><<<mgrid, 32 * kUMarchWarps, 0, ctx->stream>>>(g, in, dt, y, acc, out, r_lo, r_hi, chunk);
        ctx->launches++;
        return;
    }
    const dim3 block(64, 4), grid((unsigned)((g.cols + 63) / 64), (unsigned)((r_hi - r_lo + 3) / 4));
    if (strict) ugrid_stage<true, STAGE><<<grid, block, 0, ctx->stream>>>(g, in, dt, y, acc, out, r_lo, r_hi);
    else ugrid_stage<false, STAGE><<<grid, block, 0, ctx->stream>>>(g, in, dt, y, acc, out, r_lo, r_hi);
    ctx->launches++;
}

// send this slab's first / last row of `in.slab` to the rank above / below and receive their boundary rows into the ghosts
int ugrid_halo(sopot_b200_ctx* ctx, const UGridDev& g, const UIn& in) {
    if (ctx->nranks == 1) return SOPOT_B200_OK;
    const size_t row_elems = g.cols * 6;
    const NcclApi* n = ctx->nccl;
    SB_NCCL(ctx, n->GroupStart());
    if (g.row_begin > 0) {
        SB_NCCL(ctx, n->Send(in.slab, row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->Recv(const_cast<double*>(in.top), row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
    }
    if ((size_t)g.row_begin + g.rows_local < g.rows) {
        SB_NCCL(ctx, n->Send(in.slab + (g.rows_local - 1) * row_elems, row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->Recv(const_cast<double*>(in.bot), row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
    }
    SB_NCCL(ctx, n->GroupEnd());
    return SOPOT_B200_OK;
}

}  // namespace

SB_EXPORT int sopot_b200_ugrid_initial_state(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* p, uint32_t mem, double* state) {
    UGridDev g;
    int rc = check_ugrid(ctx, p, g);
    if (rc) return rc;
    if (!state || mem > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad state/mem");
    const size_t bytes = g.rows_local * g.cols * 48;
    Staging sg(ctx, mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_align(bytes) + 4096))) return rc;
    void* d;
    if ((rc = sg.out(state, bytes, &d))) return rc;
    const dim3 block(64, 4), grid((unsigned)((g.cols + 63) / 64), (unsigned)((g.rows_local + 3) / 4));
    ugrid_init_kernel<<<grid, block, 0, ctx->stream>>>(g.rows_local, g.cols, (size_t)g.row_begin, p->spacing, (double*)d);
    SB_CHECK_LAUNCH(ctx);
    return sg.finish();
}

SB_EXPORT int sopot_b200_ugrid_rhs(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* p, const sopot_b200_rk4_opts* opts,
                                   const double* state, double* out) {
    UGridDev g;
    int rc = check_ugrid(ctx, p, g);
    if (rc) return rc;
    if (!opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    if (g.rows_local != g.rows) return sb_fail(ctx, SOPOT_B200_EINVAL, "ugrid_rhs evaluates a whole grid (no slab)");
    const size_t bytes = g.rows * g.cols * 48;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, 2 * sb_align(bytes) + 4096))) return rc;
    const void* d_in; void* d_out;
    if ((rc = sg.in(state, bytes, &d_in))) return rc;
    if ((rc = sg.out(out, bytes, &d_out))) return rc;
    launch_ustage<0>(ctx, opts->fp_mode == SOPOT_B200_FP_STRICT, g, UIn{(const double*)d_in, nullptr, nullptr}, 0.0, nullptr, nullptr,
                     (double*)d_out);
    SB_CUDA(ctx, cudaGetLastError());
    return sg.finish();
}

SB_EXPORT int sopot_b200_ugrid_rk4(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* p, double dt, size_t n_steps,
                                   const sopot_b200_rk4_opts* opts, double* state) {
    return sopot_b200_ugrid_integrate(ctx, p, SOPOT_B200_INTEGRATOR_RK4, dt, n_steps, opts, state);
}

SB_EXPORT int sopot_b200_ugrid_integrate(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* p, int integrator, double dt,
                                         size_t n_steps, const sopot_b200_rk4_opts* opts, double* state) {
    UGridDev g;
    int rc = check_ugrid(ctx, p, g);
    if (rc) return rc;
    if (!opts || !state) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    if (integrator < 0 || integrator > 2) return sb_fail(ctx, SOPOT_B200_EINVAL, "integrator must be RK4, SYMPLECTIC_EULER or VELOCITY_VERLET");
    if (!(dt > 0.0)) return sb_fail(ctx, SOPOT_B200_EINVAL, "dt must be positive");
    const bool strict = opts->fp_mode == SOPOT_B200_FP_STRICT;
    ctx->ugrid_kernel = opts->flags & (SOPOT_B200_FLAG_UGRID_MARCH | SOPOT_B200_FLAG_UGRID_GATHER);
    const size_t n = g.rows_local * g.cols * 6, row_elems = g.cols * 6;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_align(n * 8) + 4096))) return rc;
    double* y;
    if (sg.host) {
        const void* d_in;
        if ((rc = sg.in(state, n * 8, &d_in))) return rc;
        y = const_cast<double*>(static_cast<const double*>(d_in));
        sg.outs.push_back({state, y, n * 8});
    } else {
        y = state;
    }
    // Several ranks, slabs at least as tall as the step has stages: ONE exchange per step.  Every buffer carries H ghost rows
    // above and below the slab (H = derivative evaluations per step: 4 RK4, 2 Verlet, 1 Euler); the H boundary rows of the
    // state travel once, and stage s is evaluated on the buffer shrunk by s rows at both ends, so the rows next to the slab
    // boundary are computed redundantly by both neighbours (12 extra row-evaluations per RK4 step) instead of being
    // exchanged after every stage.  Same operations per mass, so the result equals the single-domain run bit for bit.
    const size_t H = integrator == SOPOT_B200_INTEGRATOR_RK4 ? 4 : (integrator == SOPOT_B200_INTEGRATOR_VELOCITY_VERLET ? 2 : 1);
    if (ctx->nranks > 1 && g.rows / (size_t)ctx->nranks >= H && !std::getenv("SOPOT_B200_UGRID_HALO_PER_STAGE")) {
        UGridDev ge = g;
        ge.row_begin = g.row_begin - (long)H;
        ge.rows_local = g.rows_local + 2 * H;
        const size_t ne = ge.rows_local * g.cols * 6;
        void* scr2;
        if ((rc = sb_scratch(ctx, 4 * ne * 8 + 1024, &scr2))) return rc;
        double* Y = static_cast<double*>(scr2);
        double* ACC = Y + ne;
        double* EA = ACC + ne;
        double* EB = EA + ne;
        SB_CUDA(ctx, cudaMemcpyAsync(Y + H * row_elems, y, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        const bool up = g.row_begin > 0, down = (size_t)g.row_begin + g.rows_local < g.rows;
        const NcclApi* nc = ctx->nccl;
        const auto exchange = [&](double* buf) -> int {                 // H boundary rows of the slab <-> the neighbours' ghost rows
            SB_NCCL(ctx, nc->GroupStart());
            if (up) {
                SB_NCCL(ctx, nc->Send(buf + H * row_elems, H * row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
                SB_NCCL(ctx, nc->Recv(buf, H * row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
            }
            if (down) {
                SB_NCCL(ctx, nc->Send(buf + g.rows_local * row_elems, H * row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
                SB_NCCL(ctx, nc->Recv(buf + (H + g.rows_local) * row_elems, H * row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
            }
            SB_NCCL(ctx, nc->GroupEnd());
            return SOPOT_B200_OK;
        };
        // rows of the extended buffer on which evaluation number s (1-based) of a step has valid inputs, inside the global grid
        const auto lo = [&](size_t s) { const long l = (long)s, gmin = -ge.row_begin; return (size_t)(l > gmin ? l : gmin); };
        const auto hi = [&](size_t s) { const long h = (long)(ge.rows_local - s), gmax = (long)g.rows - ge.row_begin; return (size_t)(h < gmax ? h : gmax); };
        const auto ext = [](const double* b) { return UIn{b, nullptr, nullptr}; };
        double* cur = Y;
        double* nxt = EA;
        // Overlap (opt-in, SOPOT_B200_UGRID_OVERLAP=1): the first evaluation of a step needs ghost rows only for the first and
        // last owned row and the redundant rows beyond them, so the exchange can run on the copy stream while the main stream
        // evaluates the rows in between; the few boundary rows follow once the ghost rows have arrived.  Row ranges of one
        // evaluation are independent launches of the same kernel, so the bits do not depend on the split.  Measured and left off:
        // 4096^2 on 2 GPUs 1.238 ms per RK4 step against 1.227 without, 1024^2 on 4 GPUs 0.080 against 0.071 -- the two extra
        // launches and the two stream hand-overs cost what the hidden ncclSend/Recv group (4 rows, ~10 us) saves, and more when
        // the step is launch-bound (profiles/r2_ugrid_overlap.txt).
        const bool overlap = g.rows_local >= 64 && std::getenv("SOPOT_B200_UGRID_OVERLAP");
        const size_t in_lo = H + 1, in_hi = H + g.rows_local - 1;      // rows whose 3x3 neighbourhood lies inside the owned rows
        cudaStream_t main_s = ctx->stream, copy_s = ctx->copy_stream;
        // first evaluation of a step on rows [a, b) of the extended buffer
        const auto first_eval = [&](size_t a, size_t b) {
            if (a >= b) return;
            if (integrator == SOPOT_B200_INTEGRATOR_RK4) launch_ustage<1>(ctx, strict, ge, ext(Y), dt, Y, ACC, EA, a, b);
            else if (integrator == SOPOT_B200_INTEGRATOR_SYMPLECTIC_EULER) launch_ustage<5>(ctx, strict, ge, ext(cur), dt, nullptr, nullptr, nxt, a, b);
            else launch_ustage<6>(ctx, strict, ge, ext(cur), dt, nullptr, ACC, EB, a, b);
        };
        for (size_t step = 0; step < n_steps; ++step) {
            if (overlap) {
                SB_CUDA(ctx, cudaEventRecord(ctx->ev_a, main_s));                 // `cur` is complete
                SB_CUDA(ctx, cudaStreamWaitEvent(copy_s, ctx->ev_a, 0));
                ctx->stream = copy_s;
                rc = exchange(cur);
                ctx->stream = main_s;
                if (rc) return rc;
                SB_CUDA(ctx, cudaEventRecord(ctx->ev_b, copy_s));
                first_eval(in_lo, in_hi);
                SB_CUDA(ctx, cudaStreamWaitEvent(main_s, ctx->ev_b, 0));
                first_eval(lo(1), in_lo);
                first_eval(in_hi, hi(1));
            } else {
                if ((rc = exchange(cur))) return rc;
                first_eval(lo(1), hi(1));
            }
            if (integrator == SOPOT_B200_INTEGRATOR_RK4) {
                launch_ustage<2>(ctx, strict, ge, ext(EA), dt, Y, ACC, EB, lo(2), hi(2));
                launch_ustage<3>(ctx, strict, ge, ext(EB), dt, Y, ACC, EA, lo(3), hi(3));
                launch_ustage<4>(ctx, strict, ge, ext(EA), dt, Y, ACC, nullptr, lo(4), hi(4));
            } else if (integrator == SOPOT_B200_INTEGRATOR_SYMPLECTIC_EULER) {
                double* t = cur; cur = nxt; nxt = t;
            } else {
                launch_ustage<7>(ctx, strict, ge, ext(EB), dt, nullptr, ACC, nxt, lo(2), hi(2));
                double* t = cur; cur = nxt; nxt = t;
            }
        }
        SB_CUDA(ctx, cudaMemcpyAsync(y, cur + H * row_elems, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        SB_CUDA(ctx, cudaGetLastError());
        return sg.finish();
    }
    void* scr;
    if ((rc = sb_scratch(ctx, (3 * n + 6 * row_elems) * 8 + 1024, &scr))) return rc;
    double* acc = static_cast<double*>(scr);
    double* A = acc + n;
    double* B = A + n;
    double* gh = B + n;                           // ghost rows: above / below for each of the three stencil inputs
    const UIn in_y{y, gh, gh + row_elems}, in_A{A, gh + 2 * row_elems, gh + 3 * row_elems}, in_B{B, gh + 4 * row_elems, gh + 5 * row_elems};
    if ((rc = ugrid_halo(ctx, g, in_y))) return rc;
    if (integrator == SOPOT_B200_INTEGRATOR_RK4 && ctx->nranks == 1 && (opts->flags & SOPOT_B200_FLAG_UGRID_STAGE_PAIRS)) {
        // opt-in: two stages per kernel (336 instead of 816 B per mass and step); the new state goes to a fresh buffer because the
        // second kernel still reads the old one for the ring cells of neighbouring tiles
        double* cur = y;
        double* nxt = A;
        for (size_t step = 0; step < n_steps; ++step) {
            if ((rc = launch_upair<true>(ctx, strict, g, cur, nullptr, dt, acc, B))) return rc;
            if ((rc = launch_upair<false>(ctx, strict, g, B, cur, dt, acc, nxt))) return rc;
            double* t = cur; cur = nxt; nxt = t;
        }
        if (cur != y) SB_CUDA(ctx, cudaMemcpyAsync(y, cur, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (integrator == SOPOT_B200_INTEGRATOR_RK4) {
        for (size_t step = 0; step < n_steps; ++step) {
            launch_ustage<1>(ctx, strict, g, in_y, dt, y, acc, A);
            if ((rc = ugrid_halo(ctx, g, in_A))) return rc;
            launch_ustage<2>(ctx, strict, g, in_A, dt, y, acc, B);
            if ((rc = ugrid_halo(ctx, g, in_B))) return rc;
            launch_ustage<3>(ctx, strict, g, in_B, dt, y, acc, A);
            if ((rc = ugrid_halo(ctx, g, in_A))) return rc;
            launch_ustage<4>(ctx, strict, g, in_A, dt, y, acc, nullptr);
            if (step + 1 < n_steps && (rc = ugrid_halo(ctx, g, in_y))) return rc;
        }
    } else {
        // one (symplectic Euler) or two (velocity Verlet) derivative evaluations per step; a step cannot update in place
        // (neighbours still read the old state), so the state ping-pongs between y and A
        const UIn* cur = &in_y;
        const UIn* nxt = &in_A;
        for (size_t step = 0; step < n_steps; ++step) {
            double* nxt_slab = const_cast<double*>(nxt->slab);
            if (integrator == SOPOT_B200_INTEGRATOR_SYMPLECTIC_EULER) {
                launch_ustage<5>(ctx, strict, g, *cur, dt, nullptr, nullptr, nxt_slab);
            } else {
                launch_ustage<6>(ctx, strict, g, *cur, dt, nullptr, acc, B);
                if ((rc = ugrid_halo(ctx, g, in_B))) return rc;
                launch_ustage<7>(ctx, strict, g, in_B, dt, nullptr, acc, nxt_slab);
            }
            const UIn* t = cur; cur = nxt; nxt = t;
            if (step + 1 < n_steps && (rc = ugrid_halo(ctx, g, *cur))) return rc;
        }
        if (cur->slab != y) SB_CUDA(ctx, cudaMemcpyAsync(y, cur->slab, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    SB_CUDA(ctx, cudaGetLastError());
    return sg.finish();
}
