// grid2d.cu -- K3: RK4 over the 2-D mass-spring grid (physics/connected_masses/), runtime sized.
//
// Reference semantics restated per mass (gather form):
//   spring (i -> j)       indexed_spring_2d.hpp:107-165   F = k*(len - L0) + c*(dv . u), force_on_i = F*u
//                         with len_safe = max(len, 1e-10) (:125) and no repulsion (min_distance = -1
//                         through makeGrid2DSystem, connectivity_matrix_2d.hpp:76-80)
//   aggregation           force_aggregator_2d.hpp:58-88   global edge-list order; "+=" when the mass is the
//                         edge's first endpoint, "-=" when it is the second
//   edge orders           grid_2d.hpp:92-143 (quad: horizontals, verticals, [main diagonals, anti diagonals])
//                         grid_2d.hpp:255-285 (triangular: per cell main then anti)
//   mass                  indexed_point_mass_2d.hpp:90-108  {vx, vy, Fx/m, Fy/m}
// State is AoS [rows][cols][x,y,vx,vy] (32 B per mass, row-major, grid_2d.hpp:97-99).
//
// One RK4 step = 4 stage kernels with the exact running accumulator of ensemble.cuh:
//   stage 1: k=dt*f(y);   S = k;        A = y + 0.5k      reads 32 B  writes 64 B
//   stage 2: k=dt*f(A);   S = S + 2k;   B = y + 0.5k      reads 96 B  writes 64 B
//   stage 3: k=dt*f(B);   S = S + 2k;   A = y + k         reads 96 B  writes 64 B
//   stage 4: k=dt*f(A);   y = y + (S + k)/6               reads 96 B  writes 32 B      => 544 B / mass-step
// (HBM-bound by the survey's count; the FP64 pipe is the co-limit because every
// spring costs one IEEE sqrt and two IEEE divisions, see DESIGN.md.)
//
// Multi-GPU: 1-D slab decomposition over rows.  The stencil-input buffer of every stage owns two
// ghost rows (above / below the slab) that are filled from the neighbouring ranks with
// ncclSend/ncclRecv in one group after the producing stage kernel.
#include "common.cuh"
#include <cstdlib>
#include "nccl_api.h"
#include "trig.cuh"

namespace {

struct GridDev {
    size_t rows, cols;        // global
    size_t row_begin, rows_local;
    double mass, k, c, L_orth, L_diag, inv_mass;
};

// stencil input: slab rows plus one ghost row on each side (ghosts unused at the global boundary)
struct StencilIn {
    const double* slab;       // [rows_local][cols][4]
    const double* ghost_top;  // [cols][4]  = global row row_begin - 1
    const double* ghost_bot;  // [cols][4]  = global row row_begin + rows_local
    __device__ __forceinline__ const double* row(long long r_local, size_t rows_local, size_t cols) const {
        if (r_local < 0) return ghost_top;
        if ((size_t)r_local >= rows_local) return ghost_bot;
        return slab + (size_t)r_local * cols * 4;
    }
};

struct Mass4 { double x, y, vx, vy; };

__device__ __forceinline__ Mass4 load_mass(const double* rowp, size_t col) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(rowp + 4 * col));
    const double2 b = __ldg(reinterpret_cast<const double2*>(rowp + 4 * col) + 1);
    return Mass4{a.x, a.y, b.x, b.y};
}

// force on the FIRST endpoint i of spring (i -> j)
template <bool S>
__device__ __forceinline__ void spring_force(const Mass4& pi, const Mass4& pj, const double k, const double c,
                                             const double L0, Real<S>& fx, Real<S>& fy) {
    using R = Real<S>;
    const R dx = R(pj.x) - R(pi.x), dy = R(pj.y) - R(pi.y);
    R length, ux, uy;
    if constexpr (S) {
        length = r_sqrt(dx * dx + dy * dy);
        const R length_safe = length.v < 1e-10 ? R(1e-10) : length;           // :125
        ux = dx / length_safe; uy = dy / length_safe;
    } else {
        // root and reciprocal root from one MUFU seed, no IEEE slow-path branches; coincident masses (the
        // len_safe clamp) take the generic route
        const double l2 = fma(dx.v, dx.v, dy.v * dy.v);
        double inv;
        if (l2 > 1e-20) sbtrig::sqrt_rsqrt_fast(l2, length.v, inv);
        else { length.v = sqrt(l2); inv = 1.0 / (length.v < 1e-10 ? 1e-10 : length.v); }
        ux = dx * R(inv); uy = dy * R(inv);
    }
    const R extension = length - R(L0);
    const R dvx = R(pj.vx) - R(pi.vx), dvy = R(pj.vy) - R(pi.vy);
    const R rel = dvx * ux + dvy * uy;
    const R F = R(k) * extension + R(c) * rel;                              // :140
    fx = F * ux; fy = F * uy;
}

// derivative of one mass, gather form (every incident spring evaluated here).  nb(dr, dc) returns the state of the
// neighbour at (row + dr, col + dc); it is only called for neighbours that exist in the global grid.
template <bool S, int TOPO, class Fetch>
__device__ __forceinline__ void mass_derivative_from(const GridDev& g, const Mass4& P, const bool has_l, const bool has_r,
                                                     const bool has_u, const bool has_d, Fetch nb, Real<S> (&d)[4]) {
    using R = Real<S>;
    R Fx(0.0), Fy(0.0), fx, fy;
    // horizontals then verticals (global edge order): left(-), right(+), up(-), down(+)
    if (has_l) { spring_force<S>(nb(0, -1), P, g.k, g.c, g.L_orth, fx, fy); Fx -= fx; Fy -= fy; }
    if (has_r) { spring_force<S>(P, nb(0, 1), g.k, g.c, g.L_orth, fx, fy); Fx += fx; Fy += fy; }
    if (has_u) { spring_force<S>(nb(-1, 0), P, g.k, g.c, g.L_orth, fx, fy); Fx -= fx; Fy -= fy; }
    if (has_d) { spring_force<S>(P, nb(1, 0), g.k, g.c, g.L_orth, fx, fy); Fx += fx; Fy += fy; }
    if constexpr (TOPO == 1) {
        // all main diagonals (r,c)-(r+1,c+1), then all anti-diagonals (r,c)-(r+1,c-1)
        if (has_u && has_l) { spring_force<S>(nb(-1, -1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_r) { spring_force<S>(P, nb(1, 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
        if (has_u && has_r) { spring_force<S>(nb(-1, 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_l) { spring_force<S>(P, nb(1, -1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
    } else if constexpr (TOPO == 2) {
        // per cell, row-major: main (r,c)-(r+1,c+1) then anti (r,c+1)-(r+1,c)
        if (has_u && has_l) { spring_force<S>(nb(-1, -1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_u && has_r) { spring_force<S>(nb(-1, 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_l) { spring_force<S>(P, nb(1, -1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
        if (has_d && has_r) { spring_force<S>(P, nb(1, 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
    }
    d[0] = R(P.vx); d[1] = R(P.vy);
    d[2] = r_div_by<S>(Fx, g.mass, g.inv_mass); d[3] = r_div_by<S>(Fy, g.mass, g.inv_mass);      // strict: correctly rounded
}

template <bool S, int TOPO>
__device__ __forceinline__ void mass_derivative(const GridDev& g, const StencilIn& in, const size_t rl,
                                                const size_t col, Real<S> (&d)[4], Mass4& self_out) {
    const size_t gr = g.row_begin + rl;
    const double* r0 = in.row((long long)rl, g.rows_local, g.cols);
    const Mass4 P = load_mass(r0, col);
    self_out = P;
    const bool has_l = col > 0, has_r = col + 1 < g.cols, has_u = gr > 0, has_d = gr + 1 < g.rows;
    const double* ru = has_u ? in.row((long long)rl - 1, g.rows_local, g.cols) : r0;
    const double* rd = has_d ? in.row((long long)rl + 1, g.rows_local, g.cols) : r0;
    auto nb = [&](int dr, int dc) { return load_mass(dr < 0 ? ru : (dr > 0 ? rd : r0), col + dc); };
    mass_derivative_from<S, TOPO>(g, P, has_l, has_r, has_u, has_d, nb, d);
}

__device__ __forceinline__ void store4(double* p, const double a, const double b, const double c, const double d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
}

// One RK stage for every mass of the slab rows [r_lo, r_hi).  STAGE 0 = derivative only (out <- f(in)).
template <bool S, int TOPO, int STAGE>
__global__ void __launch_bounds__(256)
grid_stage_gather(const GridDev g, const StencilIn in, const size_t r_lo, const size_t r_hi, const double dt_,
                  double* __restrict__ y, double* __restrict__ acc, double* __restrict__ out) {
    using R = Real<S>;
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t rl = r_lo + (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= g.cols || rl >= r_hi) return;
    R d[4];
    Mass4 self;
    mass_derivative<S, TOPO>(g, in, rl, col, d, self);
    const size_t off = (rl * g.cols + col) * 4;
    if constexpr (STAGE == 0) { store4(out + off, d[0].v, d[1].v, d[2].v, d[3].v); return; }
    const R dt(dt_);
    R k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = d[j] * dt;
    if constexpr (STAGE == 1) {
        // the stencil input IS y
        const R y0(self.x), y1(self.y), y2(self.vx), y3(self.vy);
        store4(acc + off, k[0].v, k[1].v, k[2].v, k[3].v);
        store4(out + off, (y0 + 0.5 * k[0]).v, (y1 + 0.5 * k[1]).v, (y2 + 0.5 * k[2]).v, (y3 + 0.5 * k[3]).v);
    } else {
        const double2 ya = *reinterpret_cast<const double2*>(y + off), yb = *(reinterpret_cast<const double2*>(y + off) + 1);
        const double2 sa = *reinterpret_cast<const double2*>(acc + off), sb = *(reinterpret_cast<const double2*>(acc + off) + 1);
        const R yv[4] = {R(ya.x), R(ya.y), R(yb.x), R(yb.y)};
        const R sv[4] = {R(sa.x), R(sa.y), R(sb.x), R(sb.y)};
        if constexpr (STAGE == 2 || STAGE == 3) {
            R s2[4], t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s2[j] = sv[j] + 2.0 * k[j];
                t[j] = (STAGE == 2) ? yv[j] + 0.5 * k[j] : yv[j] + k[j];
            }
            store4(acc + off, s2[0].v, s2[1].v, s2[2].v, s2[3].v);
            store4(out + off, t[0].v, t[1].v, t[2].v, t[3].v);
        } else {
            R yn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) yn[j] = yv[j] + r_div_const(sv[j] + k[j], 6.0);
            store4(y + off, yn[0].v, yn[1].v, yn[2].v, yn[3].v);
        }
    }
}

// ---- fused RK4 step: all four stages in one kernel on a shared-memory tile -----------------------------------
// A CTA owns TX x TY interior masses and loads them with a 4-cell halo (one cell per RK stage).  Stage s evaluates
// the derivative on the region shrunk by s cells, so the stencil input of stage s+1 is complete inside the tile:
//   Y (halo 4) -> k1 on halo 3 -> A1 = y + k1/2 (halo 3) -> k2 on halo 2 -> A2 (halo 2) -> k3 on halo 1 -> A3 = y + k3
//   (halo 1) -> k4 on the interior -> y_new = y + (S + k4)/6, S = k1 + 2 k2 + 2 k3 kept in registers.
// Halo cells are recomputed by the neighbouring CTA with the same operations on the same inputs, so the result is
// bit-identical to the four-kernel schedule (FP_STRICT: to the reference).  HBM traffic per mass and step: 32 B read
// (+ halo re-reads, served by L2) and 32 B written, instead of 544 B -- the step becomes FP64-bound, so the FP64
// work is halved as well: every spring is evaluated ONCE, by its first endpoint (the reference's edge orientation:
// left->right, up->down, and for the diagonal topologies up-left->down-right, up-right->down-left), and its force
// goes through shared memory to the second endpoint, which subtracts it -- exactly the reference's
// "+= on edge.first, -= on edge.second" (force_aggregator_2d.hpp:69-81; negation is exact).
//
// Each stage is two phases separated by __syncthreads():
//   spring phase  cells of margin >= s-1: forces of the cell's right/down (and diagonal) springs -> F planes
//   update phase  cells of margin >= s  : sum the incident forces in edge-list order, k = dt*f, write the next
//                 stage input (or, in stage 4, y_new to global memory)
// Ownership is static: thread t owns interior cell (4 + t/TX, 4 + t%TX) and, for t < #halo cells, ONE halo cell,
// enumerated ring by ring from the inside out, so "margin >= m" is a prefix of the enumeration and the halo cell's
// own y stays in registers for all stages.
// Shared memory: two SoA state buffers (ping-pong stage inputs) and one force buffer of (TX+8) x (TY+8) cells
// (x 4 planes; 8 planes of forces with diagonals); 64-bit plane accesses are bank-conflict free.
template <int TX, int TY, int TOPO>
struct FusedTile {
    static constexpr int H = 4;
    // WS = row stride of the planes in shared memory.  W = 40 doubles is 8 mod 16 bank pairs: the halo-ring cells of the two
    // side columns (one cell per row, consecutive lanes) hit 2-4 bank pairs -- 0.75 conflicts per shared access in round 1's
    // ncu capture.  An odd stride spreads any column over all banks.
    static constexpr int W = TX + 2 * H, HT = TY + 2 * H, WS = W + 1, CELLS = WS * HT;
    static constexpr int THREADS = TX * TY;
    static constexpr int FPLANES = TOPO == 0 ? 4 : 8;
    static constexpr size_t SMEM = (size_t)(8 + FPLANES) * CELLS * sizeof(double);
    // halo cells with margin >= m (margin = distance to the tile edge), m = 0..3; margin 4 = interior
    static constexpr int ring(int m) { return 2 * (W - 2 * m) + 2 * (HT - 2 * m - 2); }
    static constexpr int upto(int m) { int n = 0; for (int k = H - 1; k >= m; --k) n += ring(k); return n; }
    static_assert(upto(0) == W * HT - TX * TY && upto(0) <= THREADS, "one halo cell per thread");
};

// y of the local slab plus FOUR ghost rows on either side (ghost_top row j = global row row_begin - 4 + j,
// ghost_bot row j = global row row_begin + rows_local + j)
struct StepIn {
    const double* slab;
    const double* ghost_top;
    const double* ghost_bot;
};

template <int CELLS>
__device__ __forceinline__ Mass4 tile_cell(const double* __restrict__ pl, const int o) {
    return Mass4{pl[o], pl[CELLS + o], pl[2 * CELLS + o], pl[3 * CELLS + o]};
}

// spring phase for one cell: forces of the springs whose FIRST endpoint is this cell
template <bool S, int TOPO, int TX, int TY, bool FULL>
__device__ __forceinline__ void fused_springs(const GridDev& g, const double* __restrict__ in_pl, double* __restrict__ F,
                                              const int r, const int c, const long long gr, const long long gc) {
    using T = FusedTile<TX, TY, TOPO>;
    const int o = r * T::WS + c;
    const bool here = FULL || (gr >= 0 && gc >= 0 && gr < (long long)g.rows && gc < (long long)g.cols);
    const bool right = c + 1 < T::W && here && (FULL || gc + 1 < (long long)g.cols);
    const bool down = r + 1 < T::HT && here && (FULL || gr + 1 < (long long)g.rows);
    Mass4 P;
    if (right || down) P = tile_cell<T::CELLS>(in_pl, o);
    Real<S> fx, fy;
    if (right) { spring_force<S>(P, tile_cell<T::CELLS>(in_pl, o + 1), g.k, g.c, g.L_orth, fx, fy); F[o] = fx.v; F[T::CELLS + o] = fy.v; }
    if (down) { spring_force<S>(P, tile_cell<T::CELLS>(in_pl, o + T::WS), g.k, g.c, g.L_orth, fx, fy); F[2 * T::CELLS + o] = fx.v; F[3 * T::CELLS + o] = fy.v; }
    if constexpr (TOPO != 0) {
        const bool dr = down && c + 1 < T::W && (FULL || gc + 1 < (long long)g.cols);
        const bool dl = down && c >= 1 && (FULL || gc >= 1);
        if (dr) { spring_force<S>(P, tile_cell<T::CELLS>(in_pl, o + T::WS + 1), g.k, g.c, g.L_diag, fx, fy); F[4 * T::CELLS + o] = fx.v; F[5 * T::CELLS + o] = fy.v; }
        if (dl) { spring_force<S>(P, tile_cell<T::CELLS>(in_pl, o + T::WS - 1), g.k, g.c, g.L_diag, fx, fy); F[6 * T::CELLS + o] = fx.v; F[7 * T::CELLS + o] = fy.v; }
    }
}

// update phase for one cell: aggregate in the reference's edge-list order, then the stage arithmetic of rk4_step
template <bool S, int TOPO, int STAGE, int TX, int TY, bool FULL>
__device__ __forceinline__ void fused_update(const GridDev& g, const double dt_, const double* __restrict__ in_pl,
                                             const double* __restrict__ F, double* __restrict__ out_pl, const int r, const int c,
                                             const long long gr, const long long gc, const Real<S> (&y)[4], Real<S> (&Sacc)[4],
                                             const bool interior, double* __restrict__ y_out) {
    using R = Real<S>;
    using T = FusedTile<TX, TY, TOPO>;
    if (!FULL && (gr < 0 || gc < 0 || gr >= (long long)g.rows || gc >= (long long)g.cols)) return;
    const int o = r * T::WS + c;
    const bool has_l = FULL || gc > 0, has_r = FULL || gc + 1 < (long long)g.cols;
    const bool has_u = FULL || gr > 0, has_d = FULL || gr + 1 < (long long)g.rows;
    R Fx(0.0), Fy(0.0);
    // horizontals then verticals: left(-), right(+), up(-), down(+)
    if (has_l) { Fx -= R(F[o - 1]); Fy -= R(F[T::CELLS + o - 1]); }
    if (has_r) { Fx += R(F[o]); Fy += R(F[T::CELLS + o]); }
    if (has_u) { Fx -= R(F[2 * T::CELLS + o - T::WS]); Fy -= R(F[3 * T::CELLS + o - T::WS]); }
    if (has_d) { Fx += R(F[2 * T::CELLS + o]); Fy += R(F[3 * T::CELLS + o]); }
    if constexpr (TOPO == 1) {        // main diagonals, then anti-diagonals
        if (has_u && has_l) { Fx -= R(F[4 * T::CELLS + o - T::WS - 1]); Fy -= R(F[5 * T::CELLS + o - T::WS - 1]); }
        if (has_d && has_r) { Fx += R(F[4 * T::CELLS + o]); Fy += R(F[5 * T::CELLS + o]); }
        if (has_u && has_r) { Fx -= R(F[6 * T::CELLS + o - T::WS + 1]); Fy -= R(F[7 * T::CELLS + o - T::WS + 1]); }
        if (has_d && has_l) { Fx += R(F[6 * T::CELLS + o]); Fy += R(F[7 * T::CELLS + o]); }
    } else if constexpr (TOPO == 2) { // per cell: main then anti
        if (has_u && has_l) { Fx -= R(F[4 * T::CELLS + o - T::WS - 1]); Fy -= R(F[5 * T::CELLS + o - T::WS - 1]); }
        if (has_u && has_r) { Fx -= R(F[6 * T::CELLS + o - T::WS + 1]); Fy -= R(F[7 * T::CELLS + o - T::WS + 1]); }
        if (has_d && has_l) { Fx += R(F[6 * T::CELLS + o]); Fy += R(F[7 * T::CELLS + o]); }
        if (has_d && has_r) { Fx += R(F[4 * T::CELLS + o]); Fy += R(F[5 * T::CELLS + o]); }
    }
    R d[4];
    d[0] = R(in_pl[2 * T::CELLS + o]); d[1] = R(in_pl[3 * T::CELLS + o]);           // {vx, vy, Fx/m, Fy/m}
    d[2] = r_div_by<S>(Fx, g.mass, g.inv_mass); d[3] = r_div_by<S>(Fy, g.mass, g.inv_mass);
    const R dt(dt_);
    R k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = d[j] * dt;
    if constexpr (STAGE == 4) {
        const size_t off = (((size_t)gr - g.row_begin) * g.cols + (size_t)gc) * 4;
        R yn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) yn[j] = y[j] + r_div_const(Sacc[j] + k[j], 6.0);
        store4(y_out + off, yn[0].v, yn[1].v, yn[2].v, yn[3].v);
    } else {
        if (interior) {
#pragma unroll
            for (int j = 0; j < 4; ++j) Sacc[j] = (STAGE == 1) ? k[j] : Sacc[j] + 2.0 * k[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) out_pl[j * T::CELLS + o] = ((STAGE == 3) ? y[j] + k[j] : y[j] + 0.5 * k[j]).v;
    }
}

template <bool S, int TOPO, int TX, int TY, bool FULL>
__device__ __forceinline__ void fused_step_body(const GridDev& g, const long long r0, const long long c0, const double dt,
                                                double* __restrict__ A, double* __restrict__ B, double* __restrict__ F,
                                                double* __restrict__ y_out) {
    using T = FusedTile<TX, TY, TOPO>;
    using R = Real<S>;
    constexpr int H = T::H;
    const int tid = threadIdx.x;
    // interior cell of this thread
    const int ri = H + tid / TX, ci = H + tid % TX;
    // halo cell of this thread: rings enumerated from margin 3 (innermost) to margin 0
    int rr = 0, cr = 0, mr = -1;          // mr = margin of the halo cell, -1: none
    if (tid < T::upto(0)) {
        int q = tid;
        mr = H - 1;
#pragma unroll
        for (int m = H - 1; m >= 1; --m) if (tid >= T::upto(m)) { mr = m - 1; q = tid - T::upto(m); }
        const int w = T::W - 2 * mr;
        if (q < w) { rr = mr; cr = mr + q; }
        else if (q < 2 * w) { rr = T::HT - 1 - mr; cr = mr + q - w; }
        else { q -= 2 * w; rr = mr + 1 + (q >> 1); cr = (q & 1) ? T::W - 1 - mr : mr; }
    }
    const int oi = ri * T::WS + ci, orr = rr * T::WS + cr;
    R yi[4], yr[4], Sacc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { yi[j] = R(A[j * T::CELLS + oi]); yr[j] = R(mr >= 0 ? A[j * T::CELLS + orr] : 0.0); }
    // rows of the next rank's slab are computed as halo but never written
    const bool write_i = r0 + ri < (long long)(g.row_begin + g.rows_local);

#define SB_FUSED_STAGE(STAGE, IN, OUT)                                                                                  \
    fused_springs<S, TOPO, TX, TY, FULL>(g, IN, F, ri, ci, r0 + ri, c0 + ci);                                           \
    if (mr >= STAGE - 1) fused_springs<S, TOPO, TX, TY, FULL>(g, IN, F, rr, cr, r0 + rr, c0 + cr);                      \
    __syncthreads();                                                                                                    \
    if (STAGE < 4 || write_i)                                                                                           \
        fused_update<S, TOPO, STAGE, TX, TY, FULL>(g, dt, IN, F, OUT, ri, ci, r0 + ri, c0 + ci, yi, Sacc, true, y_out);  \
    if (STAGE < 4 && mr >= STAGE)                                                                                       \
        fused_update<S, TOPO, STAGE, TX, TY, FULL>(g, dt, IN, F, OUT, rr, cr, r0 + rr, c0 + cr, yr, Sacc, false, y_out); \
    if (STAGE < 4) __syncthreads();

    SB_FUSED_STAGE(1, A, B)
    SB_FUSED_STAGE(2, B, A)
    SB_FUSED_STAGE(3, A, B)
    SB_FUSED_STAGE(4, B, A)
#undef SB_FUSED_STAGE
}

template <bool S, int TOPO, int TX, int TY>
__global__ void __launch_bounds__(TX * TY, 2)
grid_rk4_fused(const GridDev g, const StepIn in, const double dt, double* __restrict__ y_out, const unsigned ty_first,
               const unsigned ty_stride) {
    using T = FusedTile<TX, TY, TOPO>;
    extern __shared__ double fused_smem[];
    double* A = fused_smem;
    double* B = A + 4 * T::CELLS;
    double* F = B + 4 * T::CELLS;
    // GLOBAL grid coordinates of cell (0, 0) of the shared-memory tile
    // tile row of this CTA: ty_first + blockIdx.y * ty_stride (a launch covers all tile rows, only the interior ones, or --
    // with stride = last tile row -- the two boundary ones: see grid_rk4_fused_run)
    const unsigned tile_row = ty_first + blockIdx.y * ty_stride;
    const long long r0 = (long long)g.row_begin + (long long)tile_row * TY - T::H;
    const long long c0 = (long long)blockIdx.x * TX - T::H;
    for (int idx = threadIdx.x; idx < T::W * T::HT; idx += T::THREADS) {
        const int r = idx / T::W, c = idx % T::W;
        const long long gr = r0 + r, gc = c0 + c;
        double2 a = make_double2(0.0, 0.0), b = a;
        const long long rl = gr - (long long)g.row_begin;
        // (only 4 ghost rows exist below the slab: a ragged last tile row must not read past them)
        if (gr >= 0 && gc >= 0 && gr < (long long)g.rows && gc < (long long)g.cols && rl < (long long)g.rows_local + T::H) {
            const double* rowp = rl < 0 ? in.ghost_top + (size_t)(rl + T::H) * g.cols * 4
                               : rl >= (long long)g.rows_local ? in.ghost_bot + (size_t)(rl - (long long)g.rows_local) * g.cols * 4
                               : in.slab + (size_t)rl * g.cols * 4;
            a = __ldg(reinterpret_cast<const double2*>(rowp + 4 * gc));
            b = __ldg(reinterpret_cast<const double2*>(rowp + 4 * gc) + 1);
        }
        const int o = r * T::WS + c;
        A[o] = a.x; A[T::CELLS + o] = a.y; A[2 * T::CELLS + o] = b.x; A[3 * T::CELLS + o] = b.y;
    }
    __syncthreads();
    // tiles that lie entirely inside the global grid (all but the boundary tiles) skip every existence test
    const bool full = r0 >= 0 && c0 >= 0 && r0 + T::HT <= (long long)g.rows && c0 + T::W <= (long long)g.cols;
    if (full) fused_step_body<S, TOPO, TX, TY, true>(g, r0, c0, dt, A, B, F, y_out);
    else fused_step_body<S, TOPO, TX, TY, false>(g, r0, c0, dt, A, B, F, y_out);
}

// ---- fused RK4 step, marching form (quad topology): one warp per strip of columns, rows streamed through the warp --------
// The tile kernel above spends most of its cycles on barriers, shared-memory latency and index arithmetic (ncu, round 1:
// FP64 pipe 34 % busy, 0.75 bank conflicts per shared access, ~700 integer / control instructions per mass-step).  Here a warp
// owns a strip of 32 consecutive columns and marches down the rows with the four RK stages in a software pipeline.  Tick r:
//     load y[r+1] (prefetch);  stage 1 on row r-1;  stage 2 on row r-3;  stage 3 on row r-5;  stage 4 on row r-7 -> store.
// Stage s on row q needs its input (y, A1, A2 or A3) on rows q and q+1 only: every spring is evaluated ONCE, by its first
// endpoint -- the right spring by the lane itself (neighbour state by warp shuffle, the force handed to the right lane by a
// second shuffle), the down spring from rows q and q+1, and its force is CARRIED in registers to the next tick, where it is the
// up spring of row q+1.  The reference's "+= on edge.first, -= on edge.second" in edge-list order left(-), right(+), up(-),
// down(+) (force_aggregator_2d.hpp:69-81) is kept, so FP_STRICT results equal the tile kernel's, the per-stage schedule's and
// the reference's bit for bit.  The lags 1, 3, 5, 7 make the four stages of one tick INDEPENDENT of each other (stage s reads
// what stage s-1 produced in the two previous ticks), which is where the instruction-level parallelism comes from: four
// spring / update chains interleave in one basic block, no barrier, no block-level communication at all.
// What a stage produces is handed on in registers (it is the `next` row one tick later and the `current` row two ticks later);
// the two things that live longer -- y of the last 8 rows (the base of every stage update) and the stage accumulator S of the
// rows between stage 1 and stage 4 -- sit in a warp-private shared-memory ring [8 rows][4][32 lanes]: every lane touches only
// its own column, so there is no synchronisation and no bank conflict (64-bit accesses, consecutive lanes).
// Redundancy instead of exchange: lanes 0-3 and 28-31 of a warp recompute the 4-column halo of the neighbouring strips (stage s
// is valid on lanes s .. 31-s, 24 of 32 lanes own an output column), and a chunk of rows runs its pipeline 10 ticks longer.
// ---- halo exchange over NVLink peer memory, fused into the step kernel (several ranks) -----------------------------------
// With NCCL the fused step costs, per RK4 step, a send/recv group (a proxy-driven kernel per peer) in front of the compute
// kernel.  Here the step kernel itself does the exchange: the warps that produce the first / last four rows of the slab store
// them a second time, straight into the ghost rows of the rank above / below (its buffer is mapped into this process with
// CUDA IPC, the stores travel over NVLink), every warp ends with  __threadfence_system + atomic count, and the last one to
// finish raises a sequence flag in both neighbours' memory (st.release.sys).  The only warps that ever WAIT are those of the
// first and the last chunk of rows, which read ghost rows: they poll their own flag (ld.acquire.sys) while all other chunks
// already compute -- the transfer hides behind the interior of the slab.  Ghost rows are double-buffered by the parity of the
// exchange number; a writer can only get one exchange ahead of the reader it writes to (it needs the reader's rows first),
// which is exactly what two buffers cover.  No collective call on the step path at all.
struct MarchP2P {
    int on = 0;
    unsigned wait_seq = 0, signal_seq = 0, done_target = 0;
    const unsigned* flag_up = nullptr;        // raised by the rank above when its rows for exchange wait_seq are in my ghost buffer
    const unsigned* flag_down = nullptr;
    unsigned* peer_up_flag = nullptr;         // the rank above: its "from below" flag;  the rank below: its "from above" flag
    unsigned* peer_down_flag = nullptr;
    double* peer_up_slot = nullptr;           // the rank above: its from-below ghost slot of the NEXT exchange's parity
    double* peer_down_slot = nullptr;         // the rank below: its from-above ghost slot
    unsigned* done = nullptr;                 // warps of this launch series that have finished
    unsigned* err = nullptr;
    long long timeout_clocks = 0;             // how long a CTA waits for a neighbouring rank before it raises err (p2p_timeout_clocks)
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// ONE thread of a CTA polls until *flag has reached seq (sequence numbers wrap: compare the difference), relaxed loads with a
// back-off and one acquire fence at the end -- hundreds of warps hammering a flag with system-scope acquire loads slow down the
// very stores they wait for; gives up after timeout_clocks (one minute by default: a neighbour may be held up on its host for
// seconds without anything being wrong, and a waiting CTA costs nothing but its slot) and raises err
__device__ __noinline__ void p2p_wait(const unsigned* flag, const unsigned seq, unsigned* err, const long long timeout_clocks) {
    const long long t0 = clock64();
    unsigned v;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int)(v - seq) >= 0) break;
        __nanosleep(500);
        if (clock64() - t0 > timeout_clocks) { atomicExch(err, 1u); break; }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}
// every warp of a chunk that borders a neighbouring rank calls this exactly once when it is done (the other warps write no
// peer memory and take no part): the last one signals both neighbours
__device__ __noinline__ void p2p_finish(const MarchP2P& x, const bool has_up, const bool has_down) {
    const bool first = blockIdx.y == 0, last = blockIdx.y == gridDim.y - 1;
    if (!((first && has_up) || (last && has_down))) return;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        __threadfence_system();                                  // this warp's peer stores before its count
        if (atomicAdd(x.done, 1u) + 1u == x.done_target) {
            __threadfence_system();
            if (has_up) st_release_sys(x.peer_up_flag, x.signal_seq);
            if (has_down) st_release_sys(x.peer_down_flag, x.signal_seq);
        }
    }
}

// The slab's first / last four rows of the NEW state are the neighbours' ghost rows of the next step: the lanes that just stored
// them read them back (L2) and store them a second time, straight into the neighbour's memory over NVLink.  A separate function
// behind a call on purpose: the same stores inside the tick loop cost the loop 11 %, and even inlined after the loop they changed
// its schedule (the row prefetch lost its distance: long_scoreboard 0 -> 1.5 warps per issue cycle, 2.44 -> 2.94 ms).
__device__ __noinline__ void p2p_push_rows(const MarchP2P& x, const GridDev& g, const double* y_out, const long long lo, const long long hi,
                                           const long long gc, const bool has_up, const bool has_down) {
    const int lane = threadIdx.x & 31;
    const long long slab_lo = (long long)g.row_begin, slab_hi = (long long)(g.row_begin + g.rows_local);
    if (!(lane >= 4 && lane < 28 && gc >= 0 && gc < (long long)g.cols)) return;
    for (int r = 0; r < 4; ++r) {
        const long long qu = slab_lo + r, qd = slab_hi - 4 + r;
        if (has_up && qu >= lo && qu < hi) {
            const double2* src = reinterpret_cast<const double2*>(y_out + ((size_t)r * g.cols + (size_t)gc) * 4);
            const double2 a = __ldcg(src), b = __ldcg(src + 1);
            store4(x.peer_up_slot + ((size_t)r * g.cols + (size_t)gc) * 4, a.x, a.y, b.x, b.y);
        }
        if (has_down && qd >= lo && qd < hi) {
            const double2* src = reinterpret_cast<const double2*>(y_out + ((size_t)(qd - slab_lo) * g.cols + (size_t)gc) * 4);
            const double2 a = __ldcg(src), b = __ldcg(src + 1);
            store4(x.peer_down_slot + ((size_t)r * g.cols + (size_t)gc) * 4, a.x, a.y, b.x, b.y);
        }
    }
}

// first exchange of a call: boundary rows of the caller's state -> the neighbours' ghost slots (after the neighbours have
// finished reading that slot in their previous call: their flags have reached wait_seq), then the same count-and-signal
__global__ void __launch_bounds__(256) p2p_publish_kernel(const MarchP2P x, const double* __restrict__ first_rows,
                                                          const double* __restrict__ last_rows, const size_t n2 /*double2 per slot*/,
                                                          const int has_up, const int has_down) {
    if (threadIdx.x == 0) {
        if (has_up) p2p_wait(x.flag_up, x.wait_seq, x.err, x.timeout_clocks);
        if (has_down) p2p_wait(x.flag_down, x.wait_seq, x.err, x.timeout_clocks);
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        if (has_up) reinterpret_cast<double2*>(x.peer_up_slot)[i] = reinterpret_cast<const double2*>(first_rows)[i];
        if (has_down) reinterpret_cast<double2*>(x.peer_down_slot)[i] = reinterpret_cast<const double2*>(last_rows)[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(x.done, 1u) + 1u == x.done_target) {
        __threadfence_system();
        if (has_up) st_release_sys(x.peer_up_flag, x.signal_seq);
        if (has_down) st_release_sys(x.peer_down_flag, x.signal_seq);
    }
}

constexpr int kMarchRing = 8;                    // rows of y / S kept in shared memory per warp
constexpr int kMarchWarps = 4;
constexpr size_t kMarchSmem = (size_t)kMarchWarps * 2 * kMarchRing * 4 * 32 * sizeof(double);     // 64 KiB per CTA

template <bool S>
struct MarchAcc { Real<S> v[4]; };

__device__ __forceinline__ void ring_store(double* ring, const long long row, const Mass4& m) {
    double* p = ring + (size_t)(row & (kMarchRing - 1)) * 128;
    p[0] = m.x; p[32] = m.y; p[64] = m.vx; p[96] = m.vy;
}
__device__ __forceinline__ Mass4 ring_load(const double* ring, const long long row) {
    const double* p = ring + (size_t)(row & (kMarchRing - 1)) * 128;
    return Mass4{p[0], p[32], p[64], p[96]};
}

// contract-mode spring in two halves, so that the two springs a lane evaluates per stage share ONE test for coincident
// masses: geometry (difference vector, length, reciprocal length by the fast route) and force
struct SpringGeom { double dx, dy, l2, length, inv; };
__device__ __forceinline__ SpringGeom march_geom(const Mass4& pi, const Mass4& pj) {
    SpringGeom s;
    s.dx = pj.x - pi.x; s.dy = pj.y - pi.y;
    s.l2 = fma(s.dx, s.dx, s.dy * s.dy);
    sbtrig::sqrt_rsqrt_fast(s.l2, s.length, s.inv);
    return s;
}
__device__ __forceinline__ void march_geom_generic(SpringGeom& s) {       // indexed_spring_2d.hpp:118-128, any l2 >= 0
    s.length = sqrt(s.l2);
    s.inv = 1.0 / (s.length < 1e-10 ? 1e-10 : s.length);
}
__device__ __forceinline__ void march_spring(const SpringGeom& s, const Mass4& pi, const Mass4& pj, const double k, const double c,
                                             const double L0, Real<false>& fx, Real<false>& fy) {
    const double ux = s.dx * s.inv, uy = s.dy * s.inv;
    const double rel = (pj.vx - pi.vx) * ux + (pj.vy - pi.vy) * uy;
    const double F = k * (s.length - L0) + c * rel;                                         // :140
    fx.v = F * ux; fy.v = F * uy;
}

template <bool S, bool EDGE>
__device__ __forceinline__ void march_forces(const GridDev& g, const Mass4& cur, const Mass4& nxt, const bool here, const bool has_l,
                                             const bool has_r, const bool has_u, const bool has_d, Real<S>& ufx, Real<S>& ufy,
                                             Real<S>& Fx, Real<S>& Fy) {
    using R = Real<S>;
    constexpr unsigned FULLMASK = 0xffffffffu;
    // right spring (this lane -> lane + 1): neighbour state by shuffle, evaluated once, handed on to the right neighbour
    Mass4 rt{__shfl_down_sync(FULLMASK, cur.x, 1), __shfl_down_sync(FULLMASK, cur.y, 1),
             __shfl_down_sync(FULLMASK, cur.vx, 1), __shfl_down_sync(FULLMASK, cur.vy, 1)};
    // lane 31 has no right neighbour in the warp (the shuffle hands it its own state): give its discarded pseudo-spring a
    // unit length, so that neither arithmetic mode sees a zero-length spring there (generic / IEEE slow paths)
    if ((threadIdx.x & 31) == 31) rt.x = cur.x + 1.0;
    R rfx, rfy, dfx, dfy;
    if constexpr (S) {
        spring_force<true>(cur, rt, g.k, g.c, g.L_orth, rfx, rfy);
        spring_force<true>(cur, nxt, g.k, g.c, g.L_orth, dfx, dfy);
    } else {
        // both springs by the fast route; a warp in which any lane sees (nearly) coincident masses -- in practice never --
        // redoes those lanes through the generic route.  The test is conservative and cheap: one integer compare of the high
        // words (l2 >= 0, so their order is the floating-point order; a length a hair above the threshold merely takes the
        // generic route, which is valid for every length).
        SpringGeom sr = march_geom(cur, rt), sd = march_geom(cur, nxt);
        constexpr int kHi = 0x3bc79ca1;                       // high word of 1e-20 (0x3bc79ca10c924223)
        const bool deg_r = __double2hiint(sr.l2) <= kHi, deg_d = __double2hiint(sd.l2) <= kHi;
        if (__any_sync(FULLMASK, deg_r || deg_d)) {
            if (deg_r) march_geom_generic(sr);
            if (deg_d) march_geom_generic(sd);
        }
        march_spring(sr, cur, rt, g.k, g.c, g.L_orth, rfx, rfy);
        march_spring(sd, cur, nxt, g.k, g.c, g.L_orth, dfx, dfy);
    }
    const double lfx = __shfl_up_sync(FULLMASK, rfx.v, 1), lfy = __shfl_up_sync(FULLMASK, rfy.v, 1);
    // aggregation with explicit, never-contracted additions: the forces are values in their own right (shuffled, carried), and
    // the predicated (EDGE) and plain bodies must round alike -- a mass may be computed by either, depending on how the rows
    // are cut into chunks and slabs, and the result must not depend on that in contract mode either.
    // A spring that does not exist contributes nothing: the reference skips it, and so does a select (never an added 0).
    double ax = 0.0, ay = 0.0;
    if (!EDGE || (here && has_l)) { ax = __dadd_rn(ax, -lfx); ay = __dadd_rn(ay, -lfy); }
    if (!EDGE || (here && has_r)) { ax = __dadd_rn(ax, rfx.v); ay = __dadd_rn(ay, rfy.v); }
    if (!EDGE || (here && has_u)) { ax = __dadd_rn(ax, -ufx.v); ay = __dadd_rn(ay, -ufy.v); }
    if (!EDGE || (here && has_d)) { ax = __dadd_rn(ax, dfx.v); ay = __dadd_rn(ay, dfy.v); }
    Fx = R(ax); Fy = R(ay);
    ufx = dfx; ufy = dfy;        // the down spring of this row is the up spring of the next one
}

// ---- strict arithmetic in the marching kernel --------------------------------------------------------------------------------
// spring_force<true> carries three guarded IEEE operations (a root, two quotients), each with a branch to an out-of-line
// fallback; eight springs per lane and tick made the strict marching kernel slower than the tile kernel (8.1 against 7.0 ms
// per step at 8192^2).  Here a whole RK stage of one mass is computed with the UNGUARDED inline sequences (ieee.cuh: exactly
// the guarded functions' fast branches, the reciprocal of the spring length shared by its two quotients), every operand whose
// range the guards would have tested raises one flag, the warp votes once per stage, and the rare warp with a raised flag
// recomputes the stage with the guarded functions.  Same bits on both routes: verified against the oracle like every strict
// path, and a parity test forces the guarded route (tiny and huge operands).
template <bool GUARDED>
__device__ __forceinline__ bool strict_spring(const Mass4& pi, const Mass4& pj, const double k, const double c, const double L0,
                                              double& fx, double& fy) {
    if constexpr (GUARDED) {
        Real<true> gx, gy;
        spring_force<true>(pi, pj, k, c, L0, gx, gy);
        fx = gx.v; fy = gy.v;
        return false;
    } else {
        const double dx = __dadd_rn(pj.x, -pi.x), dy = __dadd_rn(pj.y, -pi.y);
        const double l2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        // l2 in [2e-20, 2^500): the root is mid-exponent and at least 1.4e-10, so the len_safe clamp (:125) is inactive and
        // the divisor of the two quotients is the root itself; their numerators must be zero or mid-exponent
        constexpr unsigned kLo = 0x3bd79ca1u, kSpan = ((1023u + 500u) << 20) - kLo;       // high word of 2e-20 .. of 2^500
        const bool bad = ((unsigned)__double2hiint(l2) - kLo) >= kSpan || sb_out_of_range(dx) || sb_out_of_range(dy);
        const double length = sb_sqrt_seq(l2);
        const double y = sb_rcp_seq(length);
        const double ux = sb_div_seq_pos(dx, length, y), uy = sb_div_seq_pos(dy, length, y);
        const double extension = __dadd_rn(length, -L0);
        const double dvx = __dadd_rn(pj.vx, -pi.vx), dvy = __dadd_rn(pj.vy, -pi.vy);
        const double rel = __dadd_rn(__dmul_rn(dvx, ux), __dmul_rn(dvy, uy));
        const double F = __dadd_rn(__dmul_rn(k, extension), __dmul_rn(c, rel));             // :140
        fx = __dmul_rn(F, ux); fy = __dmul_rn(F, uy);
        return bad;
    }
}

// one RK stage of one mass in strict arithmetic: forces of its two springs, aggregation in edge-list order, stage update.
// Results go to (dfx, dfy) = the down spring's force, acc_out, out; returns the lane's out-of-range flag (always false when GUARDED).
template <bool EDGE, int STAGE, bool GUARDED>
__device__ __forceinline__ bool strict_stage(const GridDev& g, const double dt, const Mass4& cur, const Mass4& nxt, const Mass4& rt,
                                             const bool here, const bool has_l, const bool has_r, const bool has_u, const bool has_d,
                                             const double ufx, const double ufy, const Mass4& y, const MarchAcc<true>& acc,
                                             double& dfx, double& dfy, MarchAcc<true>& acc_out, Mass4& out) {
    constexpr unsigned FULLMASK = 0xffffffffu;
    double rfx, rfy;
    bool bad = !GUARDED && !sb_mid_exponent(g.mass);                // the divisor of Fx / m, Fy / m (uniform)
    bad |= strict_spring<GUARDED>(cur, rt, g.k, g.c, g.L_orth, rfx, rfy);
    bad |= strict_spring<GUARDED>(cur, nxt, g.k, g.c, g.L_orth, dfx, dfy);
    const double lfx = __shfl_up_sync(FULLMASK, rfx, 1), lfy = __shfl_up_sync(FULLMASK, rfy, 1);
    double Fx = 0.0, Fy = 0.0;
    if (!EDGE || (here && has_l)) { Fx = __dadd_rn(Fx, -lfx); Fy = __dadd_rn(Fy, -lfy); }
    if (!EDGE || (here && has_r)) { Fx = __dadd_rn(Fx, rfx); Fy = __dadd_rn(Fy, rfy); }
    if (!EDGE || (here && has_u)) { Fx = __dadd_rn(Fx, -ufx); Fy = __dadd_rn(Fy, -ufy); }
    if (!EDGE || (here && has_d)) { Fx = __dadd_rn(Fx, dfx); Fy = __dadd_rn(Fy, dfy); }
    double d[4] = {cur.vx, cur.vy, 0.0, 0.0};
    if constexpr (GUARDED) {
        d[2] = sb_div_by(Fx, g.mass, g.inv_mass); d[3] = sb_div_by(Fy, g.mass, g.inv_mass);
    } else {
        bad |= sb_out_of_range(Fx) || sb_out_of_range(Fy);
        d[2] = sb_div_seq_pos(Fx, g.mass, g.inv_mass); d[3] = sb_div_seq_pos(Fy, g.mass, g.inv_mass);
    }
    double k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = __dmul_rn(d[j], dt);
    const double yv[4] = {y.x, y.y, y.vx, y.vy};
    double o[4];
    if constexpr (STAGE == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double sum = __dadd_rn(acc.v[j].v, k[j]);
            double q;
            if constexpr (GUARDED) q = sb_div_by(sum, 6.0, 1.0 / 6.0);
            else { bad |= sb_out_of_range(sum); q = sb_div_seq_pos(sum, 6.0, 1.0 / 6.0); }
            o[j] = __dadd_rn(yv[j], q);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_out.v[j].v = (STAGE == 1) ? k[j] : __dadd_rn(acc.v[j].v, __dmul_rn(2.0, k[j]));
            o[j] = (STAGE == 3) ? __dadd_rn(yv[j], k[j]) : __dadd_rn(yv[j], __dmul_rn(0.5, k[j]));
        }
    }
    out = Mass4{o[0], o[1], o[2], o[3]};
    return bad;
}

// one stage of the marching pipeline for one lane: springs, aggregation, update; (ufx, ufy) carried force in / out, acc in / out
template <bool S, bool EDGE, int STAGE>
__device__ __forceinline__ void march_stage(const GridDev& g, const Real<S> dt, const Mass4& cur, const Mass4& nxt, const bool here,
                                            const bool has_l, const bool has_r, const bool has_u, const bool has_d, Real<S>& ufx,
                                            Real<S>& ufy, const Mass4& y, MarchAcc<S>& acc, Mass4& out);

// stage arithmetic of rk4_step (ensemble.cuh) for one mass: k = dt * {vx, vy, Fx/m, Fy/m}
template <bool S, int STAGE>
__device__ __forceinline__ void march_update(const GridDev& g, const Real<S> dt, const Mass4& cur, const Real<S> Fx, const Real<S> Fy,
                                             const Mass4& y, MarchAcc<S>& acc, Mass4& out) {
    using R = Real<S>;
    R d[4] = {R(cur.vx), R(cur.vy), r_div_by<S>(Fx, g.mass, g.inv_mass), r_div_by<S>(Fy, g.mass, g.inv_mass)};
    R k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = d[j] * dt;
    const R yv[4] = {R(y.x), R(y.y), R(y.vx), R(y.vy)};
    R o[4];
    if constexpr (STAGE == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = yv[j] + r_div_const(acc.v[j] + k[j], 6.0);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc.v[j] = (STAGE == 1) ? k[j] : acc.v[j] + 2.0 * k[j];
            o[j] = (STAGE == 3) ? yv[j] + k[j] : yv[j] + 0.5 * k[j];
        }
    }
    out = Mass4{o[0].v, o[1].v, o[2].v, o[3].v};
}

template <bool S, bool EDGE, int STAGE>
__device__ __forceinline__ void march_stage(const GridDev& g, const Real<S> dt, const Mass4& cur, const Mass4& nxt, const bool here,
                                            const bool has_l, const bool has_r, const bool has_u, const bool has_d, Real<S>& ufx,
                                            Real<S>& ufy, const Mass4& y, MarchAcc<S>& acc, Mass4& out) {
    if constexpr (!S) {
        Real<S> Fx, Fy;
        march_forces<S, EDGE>(g, cur, nxt, here, has_l, has_r, has_u, has_d, ufx, ufy, Fx, Fy);
        march_update<S, STAGE>(g, dt, cur, Fx, Fy, y, acc, out);
    } else {
        constexpr unsigned FULLMASK = 0xffffffffu;
        Mass4 rt{__shfl_down_sync(FULLMASK, cur.x, 1), __shfl_down_sync(FULLMASK, cur.y, 1),
                 __shfl_down_sync(FULLMASK, cur.vx, 1), __shfl_down_sync(FULLMASK, cur.vy, 1)};
        if ((threadIdx.x & 31) == 31) rt.x = cur.x + 1.0;            // see march_forces
        double dfx, dfy;
        MarchAcc<true> acc_new = acc;
        const bool bad = strict_stage<EDGE, STAGE, false>(g, dt.v, cur, nxt, rt, here, has_l, has_r, has_u, has_d, ufx.v, ufy.v, y, acc,
                                                          dfx, dfy, acc_new, out);
        if (__any_sync(FULLMASK, bad)) {
            acc_new = acc;
            strict_stage<EDGE, STAGE, true>(g, dt.v, cur, nxt, rt, here, has_l, has_r, has_u, has_d, ufx.v, ufy.v, y, acc, dfx, dfy,
                                            acc_new, out);
        }
        ufx.v = dfx; ufy.v = dfy;
        acc = acc_new;
    }
}

template <bool S, bool EDGE>
__device__ __forceinline__ void march_body(const GridDev& g, const StepIn& in, const double dt_, double* __restrict__ y_out,
                                           const long long row_lo, const long long row_hi, const long long gc, double* yring,
                                           double* sring) {
    using R = Real<S>;
    const R dt(dt_);
    const int lane = threadIdx.x & 31;
    const bool col_ok = gc >= 0 && gc < (long long)g.cols;
    const bool has_l = gc > 0, has_r = gc + 1 < (long long)g.cols;
    const bool owner = lane >= 4 && lane < 28 && col_ok;
    const long long slab_lo = (long long)g.row_begin, slab_hi = (long long)(g.row_begin + g.rows_local);
    const long long nrows = (long long)g.rows;
    const long long last_load = row_hi + 3;       // rows beyond it are outside the dependency cone of this chunk

    auto load_row = [&](long long gr) -> Mass4 {
        if (gr > last_load) gr = last_load;
        // a cell outside the grid: a placeholder at its own lattice position (never a zero-length pseudo-spring)
        if (EDGE && !(col_ok && gr >= 0 && gr < nrows)) return Mass4{(double)gc, (double)gr, 0.0, 0.0};
        const long long rl = gr - slab_lo;
        const double* rowp = rl < 0 ? in.ghost_top + (size_t)(rl + 4) * g.cols * 4
                           : gr >= slab_hi ? in.ghost_bot + (size_t)(gr - slab_hi) * g.cols * 4
                           : in.slab + (size_t)rl * g.cols * 4;
        return load_mass(rowp, (size_t)gc);
    };
    auto exists = [&](const long long q) { return !EDGE || (col_ok && q >= 0 && q < nrows); };

    // inputs of the four stages on the rows they process this tick (cur) and the row below (nxt)
    // (before the pipeline has filled they hold placeholders on distinct lattice positions, outside every dependency cone)
    const double px = (double)gc;
    Mass4 c1{}, n1{}, c2{px, -2.0, 0.0, 0.0}, n2{px, -1.0, 0.0, 0.0}, c3 = c2, n3 = n2, c4 = c2, n4 = n2;
    R u1x(0.0), u1y(0.0), u2x(0.0), u2y(0.0), u3x(0.0), u3y(0.0), u4x(0.0), u4y(0.0);   // carried up-spring forces per stage
    // first tick: r = row_lo - 3; stage 1 then processes row_lo - 4, whose down spring is the first force inside the cone
    long long r = row_lo - 3;
    n1 = load_row(r - 1);
    Mass4 pre = load_row(r);
    ring_store(yring, r - 1, n1);
    for (; r < row_hi + 7; ++r) {
        c1 = n1; n1 = pre;                                   // y[r-1], y[r]
        ring_store(yring, r, n1);
        pre = load_row(r + 1);
        Mass4 o1, o2, o3, o4;
        R Fx, Fy;
        {   // stage 1 on row r-1
            const long long q = r - 1;
            MarchAcc<S> acc;
            if constexpr (S) march_stage<S, EDGE, 1>(g, dt, c1, n1, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u1x, u1y, c1, acc, o1);
            else {
                march_forces<S, EDGE>(g, c1, n1, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u1x, u1y, Fx, Fy);
                march_update<S, 1>(g, dt, c1, Fx, Fy, c1, acc, o1);                   // o1 = A1[r-1]
            }
            ring_store(sring, q, Mass4{acc.v[0].v, acc.v[1].v, acc.v[2].v, acc.v[3].v});
        }
        {   // stage 2 on row r-3: A1 rows r-3 (c2), r-2 (n2)
            const long long q = r - 3;
            const Mass4 sv = ring_load(sring, q);
            MarchAcc<S> acc{{R(sv.x), R(sv.y), R(sv.vx), R(sv.vy)}};
            if constexpr (S) march_stage<S, EDGE, 2>(g, dt, c2, n2, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u2x, u2y, ring_load(yring, q), acc, o2);
            else {
                march_forces<S, EDGE>(g, c2, n2, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u2x, u2y, Fx, Fy);
                march_update<S, 2>(g, dt, c2, Fx, Fy, ring_load(yring, q), acc, o2);
            }      // o2 = A2[r-3]
            ring_store(sring, q, Mass4{acc.v[0].v, acc.v[1].v, acc.v[2].v, acc.v[3].v});
        }
        {   // stage 3 on row r-5: A2 rows r-5 (c3), r-4 (n3)
            const long long q = r - 5;
            const Mass4 sv = ring_load(sring, q);
            MarchAcc<S> acc{{R(sv.x), R(sv.y), R(sv.vx), R(sv.vy)}};
            if constexpr (S) march_stage<S, EDGE, 3>(g, dt, c3, n3, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u3x, u3y, ring_load(yring, q), acc, o3);
            else {
                march_forces<S, EDGE>(g, c3, n3, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u3x, u3y, Fx, Fy);
                march_update<S, 3>(g, dt, c3, Fx, Fy, ring_load(yring, q), acc, o3);
            }      // o3 = A3[r-5]
            ring_store(sring, q, Mass4{acc.v[0].v, acc.v[1].v, acc.v[2].v, acc.v[3].v});
        }
        {   // stage 4 on row r-7: A3 rows r-7 (c4), r-6 (n4)
            const long long q = r - 7;
            const Mass4 sv = ring_load(sring, q);
            MarchAcc<S> acc{{R(sv.x), R(sv.y), R(sv.vx), R(sv.vy)}};
            if constexpr (S) march_stage<S, EDGE, 4>(g, dt, c4, n4, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u4x, u4y, ring_load(yring, q), acc, o4);
            else {
                march_forces<S, EDGE>(g, c4, n4, exists(q), has_l, has_r, q > 0, q + 1 < nrows, u4x, u4y, Fx, Fy);
                march_update<S, 4>(g, dt, c4, Fx, Fy, ring_load(yring, q), acc, o4);
            }      // o4 = y_new[r-7]
            if (owner && q >= row_lo && q < row_hi)
                store4(y_out + ((size_t)(q - slab_lo) * g.cols + (size_t)gc) * 4, o4.x, o4.y, o4.vx, o4.vy);
        }
        // what stage s produced is the `next` row of stage s+1 in the coming tick and its `current` row in the one after
        c2 = n2; n2 = o1;
        c3 = n3; n3 = o2;
        c4 = n4; n4 = o3;
    }
}

#ifndef SB_MARCH_MINB
#define SB_MARCH_MINB 3
#endif
template <bool S, bool P2P>
__global__ void __launch_bounds__(32 * kMarchWarps, SB_MARCH_MINB)
grid_rk4_march(const GridDev g, const StepIn in, const double dt, double* __restrict__ y_out, const unsigned row_lo_local,
               const unsigned row_hi_local, const unsigned rows_per_chunk, const MarchP2P x) {
    constexpr int OWN = 24;                       // output columns per warp
    extern __shared__ double march_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* yring = march_smem + (size_t)warp * 2 * kMarchRing * 128 + lane;
    double* sring = yring + kMarchRing * 128;
    const long long strip = (long long)blockIdx.x * kMarchWarps + warp;
    const long long gc = strip * OWN - 4 + lane;
    const long long lo = (long long)g.row_begin + row_lo_local + (long long)blockIdx.y * rows_per_chunk;
    long long hi = lo + rows_per_chunk;
    const long long end = (long long)g.row_begin + row_hi_local;
    if (hi > end) hi = end;
    const bool has_up = g.row_begin > 0, has_down = g.row_begin + g.rows_local < g.rows;
    if constexpr (P2P) {
        // only the CTAs of the chunks that read ghost rows (the first and the last one) wait for the neighbour's rows of this
        // exchange: one thread polls, the CTA's one and only barrier releases the others
        const bool first = blockIdx.y == 0 && has_up, last = blockIdx.y == gridDim.y - 1 && has_down;
        if (first || last) {
            if (threadIdx.x == 0) {
                if (first) p2p_wait(x.flag_up, x.wait_seq, x.err, x.timeout_clocks);
                if (last) p2p_wait(x.flag_down, x.wait_seq, x.err, x.timeout_clocks);
            }
            __syncthreads();
        }
    }
    if (strip * OWN >= (long long)g.cols || lo >= hi) {      // whole warp: no column / no row to own
        if constexpr (P2P) p2p_finish(x, has_up, has_down);
        return;
    }
    // rows the pipeline reads before it has written them hold zeros, not whatever the previous CTA left: their (discarded)
    // results stay finite and ordinary, so the strict-mode operand votes are not failed by garbage
#pragma unroll
    for (int k = 0; k < 2 * kMarchRing * 4; ++k) yring[k * 32] = 0.0;
    // a strip / chunk that reaches outside the global grid takes the predicated body
    const bool edge = strip * OWN - 4 < 0 || strip * OWN - 4 + 32 > (long long)g.cols || lo - 11 < 0 || hi + 8 > (long long)g.rows;
    if (edge) march_body<S, true>(g, in, dt, y_out, lo, hi, gc, yring, sring);
    else march_body<S, false>(g, in, dt, y_out, lo, hi, gc, yring, sring);
    if constexpr (P2P) {
        p2p_push_rows(x, g, y_out, lo, hi, gc, has_up, has_down);
        p2p_finish(x, has_up, has_down);
    }
}

__global__ void grid_init_kernel(const size_t rows_local, const size_t cols, const size_t row_begin,
                                 const double spacing, double* __restrict__ state) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t rl = (size_t)blockIdx.y * blockDim.y + threadIdx.y;
    if (col >= cols || rl >= rows_local) return;
    // connectivity_matrix_2d.hpp:171-180: x = col*spacing, y = row*spacing, v = 0
    store4(state + (rl * cols + col) * 4, (double)col * spacing, (double)(row_begin + rl) * spacing, 0.0, 0.0);
}

int check_grid(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* p, GridDev& g) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p) return sb_fail(ctx, SOPOT_B200_EINVAL, "grid params NULL");
    if (p->rows < 2 || p->cols < 2) return sb_fail(ctx, SOPOT_B200_EINVAL, "Grid must have at least 2 rows and 2 columns");
    if (p->topology < 0 || p->topology > 2) return sb_fail(ctx, SOPOT_B200_EINVAL, "topology must be 0, 1 or 2");
    if (!(p->mass > 0.0)) return sb_fail(ctx, SOPOT_B200_EINVAL, "Mass must be positive");
    if (p->stiffness < 0.0 || p->damping < 0.0 || !(p->spacing > 0.0))
        return sb_fail(ctx, SOPOT_B200_EINVAL, "stiffness/damping must be non-negative and spacing positive");
    if (p->rows_local == 0 || p->row_begin + p->rows_local > p->rows)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "slab [%zu, %zu) outside grid of %zu rows", p->row_begin,
                       p->row_begin + p->rows_local, p->rows);
    if (int rc = sb_check_slab(ctx, p->rows, p->row_begin, p->rows_local)) return rc;
    g.rows = p->rows; g.cols = p->cols; g.row_begin = p->row_begin; g.rows_local = p->rows_local;
    g.mass = p->mass; g.k = p->stiffness; g.c = p->damping; g.inv_mass = 1.0 / p->mass;
    // rest lengths, connectivity_matrix_2d.hpp:190-192: spacing * sqrt(dx*dx + dy*dy)
    g.L_orth = p->spacing * std::sqrt(1.0 * 1.0 + 0.0 * 0.0);
    g.L_diag = p->spacing * std::sqrt(1.0 * 1.0 + 1.0 * 1.0);
    return SOPOT_B200_OK;
}

template <bool S, int TOPO, int STAGE>
void launch_stage(sopot_b200_ctx* ctx, cudaStream_t st, const GridDev& g, const StencilIn& in, size_t r_lo,
                  size_t r_hi, double dt, double* y, double* acc, double* out) {
    if (r_hi <= r_lo) return;
    const dim3 block(64, 4);
    const dim3 grid((unsigned)((g.cols + block.x - 1) / block.x), (unsigned)((r_hi - r_lo + block.y - 1) / block.y));
    grid_stage_gather<S, TOPO, STAGE><<<grid, block, 0, st>>>(g, in, r_lo, r_hi, dt, y, acc, out);
    ctx->launches++;
}

template <bool S, int TOPO>
void launch_stage_n(sopot_b200_ctx* ctx, cudaStream_t st, int stage, const GridDev& g, const StencilIn& in,
                    size_t r_lo, size_t r_hi, double dt, double* y, double* acc, double* out) {
    switch (stage) {
        case 0: launch_stage<S, TOPO, 0>(ctx, st, g, in, r_lo, r_hi, dt, y, acc, out); break;
        case 1: launch_stage<S, TOPO, 1>(ctx, st, g, in, r_lo, r_hi, dt, y, acc, out); break;
        case 2: launch_stage<S, TOPO, 2>(ctx, st, g, in, r_lo, r_hi, dt, y, acc, out); break;
        case 3: launch_stage<S, TOPO, 3>(ctx, st, g, in, r_lo, r_hi, dt, y, acc, out); break;
        default: launch_stage<S, TOPO, 4>(ctx, st, g, in, r_lo, r_hi, dt, y, acc, out); break;
    }
}

void launch_any(sopot_b200_ctx* ctx, cudaStream_t st, bool strict, int topo, int stage, const GridDev& g,
                const StencilIn& in, size_t r_lo, size_t r_hi, double dt, double* y, double* acc, double* out) {
#define SB_GRID_CASE(SS, TT) launch_stage_n<SS, TT>(ctx, st, stage, g, in, r_lo, r_hi, dt, y, acc, out)
    if (strict) { if (topo == 0) SB_GRID_CASE(true, 0); else if (topo == 1) SB_GRID_CASE(true, 1); else SB_GRID_CASE(true, 2); }
    else        { if (topo == 0) SB_GRID_CASE(false, 0); else if (topo == 1) SB_GRID_CASE(false, 1); else SB_GRID_CASE(false, 2); }
#undef SB_GRID_CASE
}

// exchange the first / last slab row of `buf` into the neighbours' ghost rows (and receive ours)
int halo_exchange(sopot_b200_ctx* ctx, const GridDev& g, const double* buf, double* ghost_top, double* ghost_bot,
                  const size_t nrows = 1, cudaStream_t st = nullptr) {
    if (!st) st = ctx->stream;
    if (ctx->nranks == 1) return SOPOT_B200_OK;
    const size_t row_elems = g.cols * 4 * nrows;      // nrows boundary rows are contiguous in the row-major slab
    const bool up = g.row_begin > 0, down = g.row_begin + g.rows_local < g.rows;
    const NcclApi* n = ctx->nccl;
    SB_NCCL(ctx, n->GroupStart());
    if (up) {
        SB_NCCL(ctx, n->Send(buf, row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, st));
        SB_NCCL(ctx, n->Recv(ghost_top, row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, st));
    }
    if (down) {
        SB_NCCL(ctx, n->Send(buf + (g.rows_local - nrows) * g.cols * 4, row_elems, SB_NCCL_FLOAT64, ctx->rank + 1,
                             ctx->nccl_comm, st));
        SB_NCCL(ctx, n->Recv(ghost_bot, row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, st));
    }
    SB_NCCL(ctx, n->GroupEnd());
    return SOPOT_B200_OK;
}

// fused schedule: per step one 4-row halo exchange (multi-rank) and one kernel, y ping-pongs with a scratch slab
constexpr int kFusedTX = 32, kFusedTY = 16;

// tile rows [ty_first, ty_first + n_ty * ty_stride) by steps of ty_stride, on `st`
template <bool S, int TOPO>
int launch_fused(sopot_b200_ctx* ctx, cudaStream_t st, const GridDev& g, const StepIn& in, double dt, double* y_out, unsigned ty_first,
                 unsigned n_ty, unsigned ty_stride) {
    using T = FusedTile<kFusedTX, kFusedTY, TOPO>;
    auto kern = grid_rk4_fused<S, TOPO, kFusedTX, kFusedTY>;
    // (a function attribute belongs to the current device: set it on every launch rather than caching a flag that a
    // second context on another GPU of the same process would inherit)
    SB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
    const dim3 grid((unsigned)((g.cols + kFusedTX - 1) / kFusedTX), n_ty);
    kern<<<grid, T::THREADS, T::SMEM, st>>>(g, in, dt, y_out, ty_first, ty_stride);
    SB_CHECK_LAUNCH(ctx);
    return SOPOT_B200_OK;
}

// quad topology: the marching kernel over local rows [ty_first * 16, (ty_first + n_ty) * 16) (callers count in tile rows of 16)
template <bool S>
int launch_march(sopot_b200_ctx* ctx, cudaStream_t st, const GridDev& g, const StepIn& in, double dt, double* y_out, unsigned ty_first,
                 unsigned n_ty, unsigned fixed_chunk, MarchP2P* p2p = nullptr) {
    const unsigned r_lo = ty_first * kFusedTY;
    unsigned r_hi = (ty_first + n_ty) * kFusedTY;
    if (r_hi > g.rows_local) r_hi = (unsigned)g.rows_local;
    if (r_hi <= r_lo) return SOPOT_B200_OK;
    // Rows per chunk.  Every chunk runs its pipeline 10 ticks longer than it has rows, and the CTAs of a launch are scheduled in
    // waves of (SMs x resident CTAs): the chunk length is chosen to minimise a simple model of the launch time over all ways
    // of cutting the rows into equal chunks of 64 .. 512 rows.  At 8192 rows on one GPU that is 128 +- a few rows (measured: 64 ->
    // 3.06 ms, 128 -> 2.90, 256 -> 2.97, 512 -> 3.14 at that stage of development); on a 1/8 slab of 1024 rows it is 205 rows = 430
    // CTAs = one wave on 148 SMs instead of 688 CTAs = 1.55 waves with 128 (8 GPUs, strict: 1.00 -> 0.85 ms per step).
    static const unsigned env_rows = std::getenv("SOPOT_B200_MARCH_ROWS") ? (unsigned)std::atoi(std::getenv("SOPOT_B200_MARCH_ROWS")) : 0u;
    const unsigned strips = (unsigned)((g.cols + 23) / 24), ctas_x = (strips + kMarchWarps - 1) / kMarchWarps;
    // (fixed_chunk: the interior launch of the overlapped multi-rank schedule keeps 128-row chunks -- a launch tuned to fill
    //  exactly one wave would leave the boundary CTAs on the other stream no slot to run beside it)
    unsigned chunk = env_rows ? env_rows : fixed_chunk;
    if (!chunk) {
        const unsigned rows = r_hi - r_lo, slots = (unsigned)ctx->sm_count * SB_MARCH_MINB;
        unsigned long long best = ~0ull;
        for (unsigned nch = 1; nch <= rows; ++nch) {
            const unsigned c = (rows + nch - 1) / nch;
            if (c > 512 && nch < rows) continue;
            if (c < 64 && nch > 1) break;
            // launch time in ticks x 1024: one CTA lives c + 10 ticks; up to one wave all CTAs run side by side, beyond that
            // the SMs stay busy (work-conserving) and about one CTA lifetime is lost in the tail
            const unsigned long long ctas = (unsigned long long)ctas_x * nch, life = c + 10;
            const unsigned long long cost = ctas <= slots ? life * 1024 : life * 1024 * ctas / slots + life * 1024;
            if (cost < best) { best = cost; chunk = c; }
        }
        if (!chunk) chunk = rows;
    }
    const dim3 grid(ctas_x, (r_hi - r_lo + chunk - 1) / chunk);
    SB_CUDA(ctx, cudaFuncSetAttribute(grid_rk4_march<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMarchSmem));
    SB_CUDA(ctx, cudaFuncSetAttribute(grid_rk4_march<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMarchSmem));
    MarchP2P x;
    if (p2p) {                                    // the warps of the bordering chunks count themselves: target = those of all launches so far
        x = *p2p;
        const bool up = g.row_begin > 0, down = g.row_begin + g.rows_local < g.rows;
        const unsigned bordering = grid.y == 1 ? ((up || down) ? 1u : 0u) : (up ? 1u : 0u) + (down ? 1u : 0u);
        x.done_target = p2p->done_target += grid.x * bordering * kMarchWarps;
    }
    static const bool force_p2p_variant = std::getenv("SOPOT_B200_MARCH_P2P_KERNEL") != nullptr;    // diagnosis: time that instantiation alone
    if (p2p || force_p2p_variant) grid_rk4_march<S, true><<<grid, 32 * kMarchWarps, kMarchSmem, st>>>(g, in, dt, y_out, r_lo, r_hi, chunk, x);
    else grid_rk4_march<S, false><<<grid, 32 * kMarchWarps, kMarchSmem, st>>>(g, in, dt, y_out, r_lo, r_hi, chunk, x);
    SB_CHECK_LAUNCH(ctx);
    return SOPOT_B200_OK;
}

// Which kernel runs the fused step of the quad topology.  The marching kernel (one warp per 24-column strip, rows streamed) is
// the fast one once it fills the GPU: 2048^2 0.195 ms per step against 0.281 for the tile kernel, 8192^2 2.4 against 4.3.  Its
// parallelism is strips x row chunks, though, and a chunk shorter than 64 rows spends more ticks filling its pipeline than
// working: below about one wave of CTAs the tile kernel (one CTA per 32 x 16 tile) wins -- 512^2: 0.025 ms per step against
// 0.082 (strict 0.040 against 0.212), 1024^2: 0.074 against 0.099 (strict 0.123 against 0.242); profiles/r2_grid_small_sizes.txt.
// Decided from the GLOBAL size (rows per rank of the even partition), so that all ranks of a slab decomposition agree.
// opts->flags SOPOT_B200_FLAG_GRID_MARCH / _TILE and the environment (SOPOT_B200_GRID_TILE_KERNEL=1, and
// SOPOT_B200_GRID_TILE_STRICT=1 for strict arithmetic only) override the choice.
bool grid_use_march(const sopot_b200_ctx* ctx, const GridDev& g, int topo, bool strict, uint32_t flags) {
    static const bool tile_only = std::getenv("SOPOT_B200_GRID_TILE_KERNEL") != nullptr;
    static const bool tile_strict = std::getenv("SOPOT_B200_GRID_TILE_STRICT") != nullptr;
    if (topo != 0 || tile_only || (strict && tile_strict) || (flags & SOPOT_B200_FLAG_GRID_TILE)) return false;
    static const bool march_only = std::getenv("SOPOT_B200_GRID_MARCH_KERNEL") != nullptr;
    if ((flags & SOPOT_B200_FLAG_GRID_MARCH) || march_only) return true;
    const size_t base_rows = g.rows / (size_t)(ctx->nranks > 0 ? ctx->nranks : 1);
    const size_t ctas = ((g.cols + 23) / 24 + kMarchWarps - 1) / kMarchWarps * ((base_rows + 63) / 64);
    // measured crossover (profiles/r2_grid_small_sizes.txt): strict arithmetic at one full wave of CTAs (between 1536^2 and 1792^2),
    // contract arithmetic at 9/16 of a wave (between 1152^2 and 1280^2) -- its tile kernel is barrier-bound, not FP64-bound
    const size_t slots = (size_t)ctx->sm_count * SB_MARCH_MINB;
    return strict ? ctas >= slots : 16 * ctas >= 9 * slots;
}

int launch_fused_any(sopot_b200_ctx* ctx, cudaStream_t st, bool strict, int topo, const GridDev& g, const StepIn& in, double dt,
                     double* y_out, unsigned ty_first, unsigned n_ty, unsigned ty_stride, bool march, unsigned fixed_chunk = 0) {
    if (topo == 0 && ty_stride == 1 && march)
        return strict ? launch_march<true>(ctx, st, g, in, dt, y_out, ty_first, n_ty, fixed_chunk)
                      : launch_march<false>(ctx, st, g, in, dt, y_out, ty_first, n_ty, fixed_chunk);
#define SB_FUSED_CASE(SS, TT) return launch_fused<SS, TT>(ctx, st, g, in, dt, y_out, ty_first, n_ty, ty_stride)
    if (strict) { if (topo == 0) SB_FUSED_CASE(true, 0); else if (topo == 1) SB_FUSED_CASE(true, 1); else SB_FUSED_CASE(true, 2); }
    else        { if (topo == 0) SB_FUSED_CASE(false, 0); else if (topo == 1) SB_FUSED_CASE(false, 1); else SB_FUSED_CASE(false, 2); }
#undef SB_FUSED_CASE
}

// ---- host side of the peer-memory exchange ----------------------------------------------------------------------------------
// p2p_prepare is COLLECTIVE (every rank of the ctx calls it with the same slot size): allocate the ghost slots and the flags,
// hand their CUDA IPC handles to the ranks above and below (two 128-byte messages over the existing NCCL communicator), map
// theirs, and agree with one all-reduce(min) whether every rank succeeded.  If any did not (IPC unavailable in the sandbox, no
// peer access), every rank keeps using the NCCL send/recv schedule.  SOPOT_B200_NO_P2P=1 skips the attempt.
int p2p_prepare(sopot_b200_ctx* ctx, const size_t slot_doubles) {
    auto& P = ctx->p2p;
    if (std::getenv("SOPOT_B200_NO_P2P")) return SOPOT_B200_OK;
    if (P.ready && P.slot_doubles >= slot_doubles) return SOPOT_B200_OK;
    if (P.tried && !P.ready) return SOPOT_B200_OK;                 // failed before: stay on NCCL
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; ++k) {
        if (P.peer_ghost[k]) cudaIpcCloseMemHandle(P.peer_ghost[k]);
        if (P.peer_flags[k]) cudaIpcCloseMemHandle(P.peer_flags[k]);
        P.peer_ghost[k] = nullptr; P.peer_flags[k] = nullptr;
    }
    if (P.ghost) cudaFree(P.ghost);
    if (P.flags) cudaFree(P.flags);
    P.ghost = nullptr; P.flags = nullptr; P.ready = false; P.tried = true; P.seq = 0;
    bool ok = cudaMalloc(&P.ghost, 4 * slot_doubles * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&P.flags, 128 * sizeof(unsigned)) == cudaSuccess &&
              cudaMemset(P.flags, 0, 128 * sizeof(unsigned)) == cudaSuccess;
    struct Msg { cudaIpcMemHandle_t ghost, flags; };               // 2 x 64 bytes = 16 doubles
    static_assert(sizeof(Msg) == 128, "IPC handles are 64 bytes each");
    Msg mine{}, from_up{}, from_down{};
    ok = ok && cudaIpcGetMemHandle(&mine.ghost, P.ghost) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, P.flags) == cudaSuccess;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof mine); }
    void* scr;
    int rc = sb_scratch(ctx, 4096, &scr);
    if (rc) return rc;
    double* d_msg = static_cast<double*>(scr);                     // [0..16) mine, [16..32) from above, [32..48) from below, [48] verdict
    SB_CUDA(ctx, cudaMemcpyAsync(d_msg, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    const bool up = ctx->rank > 0, down = ctx->rank + 1 < ctx->nranks;
    const NcclApi* n = ctx->nccl;
    SB_NCCL(ctx, n->GroupStart());
    if (up) { SB_NCCL(ctx, n->Send(d_msg, 16, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
              SB_NCCL(ctx, n->Recv(d_msg + 16, 16, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream)); }
    if (down) { SB_NCCL(ctx, n->Send(d_msg, 16, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
                SB_NCCL(ctx, n->Recv(d_msg + 32, 16, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream)); }
    SB_NCCL(ctx, n->GroupEnd());
    SB_CUDA(ctx, cudaMemcpyAsync(&from_up, d_msg + 16, sizeof from_up, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(&from_down, d_msg + 32, sizeof from_down, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const auto open = [&](const Msg& m, int k) {
        const Msg zero{};
        if (!memcmp(&m, &zero, sizeof m)) return false;            // that rank could not export its buffers
        return cudaIpcOpenMemHandle((void**)&P.peer_ghost[k], m.ghost, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess &&
               cudaIpcOpenMemHandle((void**)&P.peer_flags[k], m.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    };
    if (ok && up) ok = open(from_up, 0);
    if (ok && down) ok = open(from_down, 1);
    if (!ok) cudaGetLastError();
    const double verdict = ok ? 1.0 : 0.0;
    double all = 0.0;
    SB_CUDA(ctx, cudaMemcpyAsync(d_msg + 48, &verdict, 8, cudaMemcpyHostToDevice, ctx->stream));
    SB_NCCL(ctx, n->AllReduce(d_msg + 48, d_msg + 48, 1, SB_NCCL_FLOAT64, SB_NCCL_MIN, ctx->nccl_comm, ctx->stream));
    SB_CUDA(ctx, cudaMemcpyAsync(&all, d_msg + 48, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (all == 1.0) { P.ready = true; P.slot_doubles = slot_doubles; }
    return SOPOT_B200_OK;
}

// SM clocks a CTA waits for a neighbouring rank: SOPOT_B200_P2P_TIMEOUT_S seconds (default 60) at the device's clock rate
long long p2p_timeout_clocks(const sopot_b200_ctx* ctx) {
    static const double seconds = std::getenv("SOPOT_B200_P2P_TIMEOUT_S") ? std::atof(std::getenv("SOPOT_B200_P2P_TIMEOUT_S")) : 60.0;
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device) != cudaSuccess || khz <= 0) { cudaGetLastError(); khz = 2000000; }
    return (long long)((seconds > 0.0 ? seconds : 60.0) * 1e3 * (double)khz);
}

// the fused step with the exchange inside the kernel: publish the caller's boundary rows once, then one launch per step
template <bool S>
int grid_rk4_march_p2p(sopot_b200_ctx* ctx, const GridDev& g, double dt, size_t n_steps, double* y, double* alt) {
    auto& P = ctx->p2p;
    const size_t slot = P.slot_doubles, rows4 = 4 * g.cols * 4;
    const bool up = g.row_begin > 0, down = g.row_begin + g.rows_local < g.rows;
    const unsigned E0 = P.seq;
    SB_CUDA(ctx, cudaMemsetAsync(P.flags + 64, 0, 2 * sizeof(unsigned), ctx->stream));     // the two warp / block counters
    // slot(parity, side): side 0 = rows from the rank above (my ghost_top), 1 = from the rank below (my ghost_bot)
    const auto mine = [&](unsigned e, int side) { return P.ghost + ((size_t)(e & 1u) * 2 + side) * slot; };
    const auto peer = [&](int k, unsigned e, int side) { return P.peer_ghost[k] ? P.peer_ghost[k] + ((size_t)(e & 1u) * 2 + side) * slot : nullptr; };
    MarchP2P x;
    x.on = 1;
    x.timeout_clocks = p2p_timeout_clocks(ctx);
    x.flag_up = P.flags; x.flag_down = P.flags + 32; x.err = P.flags + 96;
    x.peer_up_flag = P.peer_flags[0] ? P.peer_flags[0] + 32 : nullptr;     // I am the rank BELOW my upper neighbour
    x.peer_down_flag = P.peer_flags[1] ? P.peer_flags[1] : nullptr;        // ... and the rank ABOVE my lower one
    {   // exchange E0 + 1: the caller's state; wait until both neighbours are past exchange E0 (their previous call)
        MarchP2P pub = x;
        pub.wait_seq = E0; pub.signal_seq = E0 + 1; pub.done = P.flags + 64; pub.done_target = 64;
        pub.peer_up_slot = peer(0, E0 + 1, 1); pub.peer_down_slot = peer(1, E0 + 1, 0);
        p2p_publish_kernel<<<64, 256, 0, ctx->stream>>>(pub, y, y + (g.rows_local - 4) * g.cols * 4, rows4 / 2, up ? 1 : 0, down ? 1 : 0);
        SB_CHECK_LAUNCH(ctx);
    }
    x.done = P.flags + 65;
    x.done_target = 0;
    double* cur = y;
    double* nxt = alt;
    const unsigned nty = (unsigned)((g.rows_local + kFusedTY - 1) / kFusedTY);
    for (size_t step = 0; step < n_steps; ++step) {
        const unsigned e = E0 + 1 + (unsigned)step;
        x.wait_seq = e; x.signal_seq = e + 1;
        x.peer_up_slot = peer(0, e + 1, 1); x.peer_down_slot = peer(1, e + 1, 0);
        const StepIn in{cur, mine(e, 0), mine(e, 1)};
        int rc = launch_march<S>(ctx, ctx->stream, g, in, dt, nxt, 0, nty, 0, &x);
        if (rc) return rc;
        double* t = cur; cur = nxt; nxt = t;
    }
    P.seq = E0 + 1 + (unsigned)n_steps;
    if (cur != y) SB_CUDA(ctx, cudaMemcpyAsync(y, cur, g.rows_local * g.cols * 4 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return SOPOT_B200_OK;
}

// One rank: one launch per step over all tiles.  Several ranks with slabs of more than 48 rows: the tiles that touch the
// slab boundary run on a second stream behind the halo exchange, the interior tiles on the main stream beside them --
//   copy stream:   wait I(s-1) -> ncclSend/Recv of the 4 boundary rows of the current state -> B(s) -> event
//   main stream:   wait B(s-1) -> I(s) -> event
// B = the first tile row and the last TWO (the last one may hold fewer than 4 rows, so the one before it can still reach
// the ghost rows and produce rows a neighbour needs), I = the tile rows between them.  B(s) and I(s) read the same state
// and write disjoint rows; the exchange only moves rows that B(s-1) wrote (same stream), and the ghost rows are read by
// B only.  The exchange and the boundary tiles hide behind the interior tiles; thinner slabs exchange first and launch
// once, as before.  (A first version with all kernels on one stream, B before I, was SLOWER than no overlap at all --
// 2.48 vs 2.37 ms per step on 2 GPUs: three partial waves instead of one launch cost more than the 30 us exchange.)
int grid_rk4_fused_run(sopot_b200_ctx* ctx, const GridDev& g, int topo, bool strict, double dt, size_t n_steps, double* y, uint32_t flags) {
    const bool march = grid_use_march(ctx, g, topo, strict, flags);
    const size_t slab = g.rows_local * g.cols * 4, ghost = 4 * g.cols * 4;
    if (ctx->nranks > 1 && g.rows_local < 4)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "the fused grid step needs >= 4 rows per rank (use SOPOT_B200_FLAG_GRID_PER_STAGE)");
    void* scr;
    int rc = sb_scratch(ctx, (slab + 2 * ghost) * 8 + 1024, &scr);
    if (rc) return rc;
    double* alt = static_cast<double*>(scr);
    double* g_top = alt + slab;
    double* g_bot = g_top + ghost;
    // several ranks, quad topology, slabs of at least 8 rows (a rank's first and last four rows are distinct): the exchange
    // goes over NVLink peer memory from inside the step kernel when every rank could set that up (collective decision)
    // (the marching kernel carries the exchange; slabs too small to fill the GPU with it take the tile kernel and NCCL below)
    if (ctx->nranks > 1 && march && g.rows / (size_t)ctx->nranks >= 8) {
        if ((rc = p2p_prepare(ctx, ghost))) return rc;
        if (ctx->p2p.ready)
            return strict ? grid_rk4_march_p2p<true>(ctx, g, dt, n_steps, y, alt) : grid_rk4_march_p2p<false>(ctx, g, dt, n_steps, y, alt);
    }
    double* cur = y;
    double* nxt = alt;
    const unsigned nty = (unsigned)((g.rows_local + kFusedTY - 1) / kFusedTY);
    // all ranks must take the same branch (slabs differ by at most one row): decide on the smallest slab
    // (check_grid admits only the even partition, so rows / nranks > 48 means rows_local >= 49 and nty >= 4 on every rank)
    // The overlap only pays when the interior launch spans several waves of CTAs, so that the boundary CTAs find free slots
    // beside it: a 1/8 slab of 8192^2 is ONE wave of the marching kernel, every slot is taken for the whole launch and the
    // boundary tiles end up BEHIND the interior ones (8 GPUs, strict: 0.98 ms per step against 0.85 for exchange-then-launch).
    // Decided from the global size alone (slabs differ by at most one row), so every rank takes the same branch.
    const size_t base_rows = g.rows / (size_t)ctx->nranks;
    const size_t interior_ctas = march
        ? ((g.cols + 23) / 24 + kMarchWarps - 1) / kMarchWarps * ((base_rows > 48 ? base_rows - 48 : 0) + 127) / 128
        : (g.cols + kFusedTX - 1) / kFusedTX * (base_rows / kFusedTY > 3 ? base_rows / kFusedTY - 3 : 0);
    const size_t slots = (size_t)ctx->sm_count * (march ? SB_MARCH_MINB : 2);
    const bool overlap = ctx->nranks > 1 && base_rows > 3 * (size_t)kFusedTY && nty >= 4 && interior_ctas >= 2 * slots &&
                         !(std::getenv("SOPOT_B200_NO_HALO_OVERLAP"));
    if (!overlap) {
        for (size_t step = 0; step < n_steps; ++step) {
            if ((rc = halo_exchange(ctx, g, cur, g_top, g_bot, 4))) return rc;
            if ((rc = launch_fused_any(ctx, ctx->stream, strict, topo, g, StepIn{cur, g_top, g_bot}, dt, nxt, 0, nty, 1, march))) return rc;
            double* t = cur; cur = nxt; nxt = t;
        }
    } else {
        cudaStream_t main_s = ctx->stream, copy_s = ctx->copy_stream;
        cudaEvent_t ev_i = ctx->ev_a, ev_b = ctx->ev_b;           // "interior tiles done" (main), "boundary tiles done" (copy)
        SB_CUDA(ctx, cudaEventRecord(ev_i, main_s));              // whatever produced `cur` counts as I(-1)
        for (size_t step = 0; step < n_steps; ++step) {
            const StepIn in{cur, g_top, g_bot};
            if (step > 0) SB_CUDA(ctx, cudaStreamWaitEvent(main_s, ev_b, 0));     // I(s) reads and overwrites rows B(s-1) used
            SB_CUDA(ctx, cudaStreamWaitEvent(copy_s, ev_i, 0));                   // B(s) reads rows I(s-1) wrote
            if ((rc = halo_exchange(ctx, g, cur, g_top, g_bot, 4, copy_s))) return rc;
            if ((rc = launch_fused_any(ctx, copy_s, strict, topo, g, in, dt, nxt, 0, 1, 1, march))) return rc;         // B(s): top
            if ((rc = launch_fused_any(ctx, copy_s, strict, topo, g, in, dt, nxt, nty - 2, 2, 1, march))) return rc;   //       bottom
            SB_CUDA(ctx, cudaEventRecord(ev_b, copy_s));
            if ((rc = launch_fused_any(ctx, main_s, strict, topo, g, in, dt, nxt, 1, nty - 3, 1, march, 128))) return rc;     // I(s)
            SB_CUDA(ctx, cudaEventRecord(ev_i, main_s));
            double* t = cur; cur = nxt; nxt = t;
        }
        SB_CUDA(ctx, cudaStreamWaitEvent(main_s, ev_b, 0));       // the caller's stream sees the whole new state
    }
    if (cur != y) SB_CUDA(ctx, cudaMemcpyAsync(y, cur, slab * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return SOPOT_B200_OK;
}

}  // namespace

SB_EXPORT int sopot_b200_grid2d_initial_state(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* p, uint32_t mem,
                                              double* state) {
    GridDev g;
    int rc = check_grid(ctx, p, g);
    if (rc) return rc;
    if (!state || mem > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad state/mem");
    const size_t bytes = g.rows_local * g.cols * 32;
    Staging sg(ctx, mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_align(bytes) + 4096))) return rc;
    void* d;
    if ((rc = sg.out(state, bytes, &d))) return rc;
    const dim3 block(64, 4), grid((unsigned)((g.cols + 63) / 64), (unsigned)((g.rows_local + 3) / 4));
    grid_init_kernel<<<grid, block, 0, ctx->stream>>>(g.rows_local, g.cols, g.row_begin, p->spacing, (double*)d);
    SB_CHECK_LAUNCH(ctx);
    return sg.finish();
}

SB_EXPORT int sopot_b200_grid2d_rhs(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* p,
                                    const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    GridDev g;
    int rc = check_grid(ctx, p, g);
    if (rc) return rc;
    if (!opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    if (g.rows_local != g.rows) return sb_fail(ctx, SOPOT_B200_EINVAL, "grid2d_rhs works on a whole grid only");
    const size_t bytes = g.rows_local * g.cols * 32;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, 2 * sb_align(bytes) + 4096))) return rc;
    const void* d_in; void* d_out;
    if ((rc = sg.in(state, bytes, &d_in))) return rc;
    if ((rc = sg.out(out, bytes, &d_out))) return rc;
    StencilIn in{(const double*)d_in, nullptr, nullptr};
    launch_any(ctx, ctx->stream, opts->fp_mode == SOPOT_B200_FP_STRICT, p->topology, 0, g, in, 0, g.rows_local, 0.0,
               nullptr, nullptr, (double*)d_out);
    SB_CUDA(ctx, cudaGetLastError());
    return sg.finish();
}

SB_EXPORT int sopot_b200_grid2d_rk4(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* p, double dt, size_t n_steps,
                                    const sopot_b200_rk4_opts* opts, double* state) {
    GridDev g;
    int rc = check_grid(ctx, p, g);
    if (rc) return rc;
    if (!opts || !state) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    if (!(dt > 0.0)) return sb_fail(ctx, SOPOT_B200_EINVAL, "dt must be positive");
    if (ctx->nranks == 1 && g.rows_local != g.rows)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "a single-rank ctx must own the whole grid");
    const bool strict = opts->fp_mode == SOPOT_B200_FP_STRICT;
    const size_t slab = g.rows_local * g.cols * 4;          // doubles
    const size_t row_elems = g.cols * 4;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_align(slab * 8) + 4096))) return rc;
    // y lives in the caller's buffer (device) or in the arena (host mode), updated in place
    double* y;
    if (sg.host) {
        const void* d_in;
        if ((rc = sg.in(state, slab * 8, &d_in))) return rc;
        y = const_cast<double*>(static_cast<const double*>(d_in));
        sg.outs.push_back({state, y, slab * 8});
    } else {
        y = state;
    }
    // The fused step exchanges 4 boundary rows at once, so every rank needs >= 4 rows; all ranks must take the same
    // branch (the exchanges are collective), hence the test on the GLOBAL size (slabs differ by at most one row).
    const bool thin_slabs = ctx->nranks > 1 && g.rows / (size_t)ctx->nranks < 4;
    if (!(opts->flags & SOPOT_B200_FLAG_GRID_PER_STAGE) && !thin_slabs) {
        rc = grid_rk4_fused_run(ctx, g, p->topology, strict, dt, n_steps, y, opts->flags);
        if (rc) return rc;
        return sg.finish();
    }
    // scratch: S, A, B slabs + ghost rows for y, A, B (top and bottom each)
    void* scr;
    if ((rc = sb_scratch(ctx, (3 * slab + 6 * row_elems) * 8 + 1024, &scr))) return rc;
    double* acc = static_cast<double*>(scr);
    double* bufA = acc + slab;
    double* bufB = bufA + slab;
    double* ghosts = bufB + slab;
    double *gy_t = ghosts, *gy_b = ghosts + row_elems, *gA_t = ghosts + 2 * row_elems, *gA_b = ghosts + 3 * row_elems,
           *gB_t = ghosts + 4 * row_elems, *gB_b = ghosts + 5 * row_elems;
    const StencilIn in_y{y, gy_t, gy_b}, in_A{bufA, gA_t, gA_b}, in_B{bufB, gB_t, gB_b};
    const int topo = p->topology;
    cudaStream_t st = ctx->stream;

    if ((rc = halo_exchange(ctx, g, y, gy_t, gy_b))) return rc;
    for (size_t step = 0; step < n_steps; ++step) {
        launch_any(ctx, st, strict, topo, 1, g, in_y, 0, g.rows_local, dt, y, acc, bufA);
        if ((rc = halo_exchange(ctx, g, bufA, gA_t, gA_b))) return rc;
        launch_any(ctx, st, strict, topo, 2, g, in_A, 0, g.rows_local, dt, y, acc, bufB);
        if ((rc = halo_exchange(ctx, g, bufB, gB_t, gB_b))) return rc;
        launch_any(ctx, st, strict, topo, 3, g, in_B, 0, g.rows_local, dt, y, acc, bufA);
        if ((rc = halo_exchange(ctx, g, bufA, gA_t, gA_b))) return rc;
        launch_any(ctx, st, strict, topo, 4, g, in_A, 0, g.rows_local, dt, y, acc, nullptr);
        if (step + 1 < n_steps && (rc = halo_exchange(ctx, g, y, gy_t, gy_b))) return rc;
    }
    SB_CUDA(ctx, cudaGetLastError());
    return sg.finish();
}
