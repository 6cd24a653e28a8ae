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
#include "nccl_api.h"

namespace {

struct GridDev {
    size_t rows, cols;        // global
    size_t row_begin, rows_local;
    double mass, k, c, L_orth, L_diag;
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
    const R length = r_sqrt(dx * dx + dy * dy);
    const R length_safe = length.v < 1e-10 ? R(1e-10) : length;            // :125
    R ux, uy;
    if constexpr (S) { ux = dx / length_safe; uy = dy / length_safe; }
    else { const R inv = R(1.0) / length_safe; ux = dx * inv; uy = dy * inv; }
    const R extension = length - R(L0);
    const R dvx = R(pj.vx) - R(pi.vx), dvy = R(pj.vy) - R(pi.vy);
    const R rel = dvx * ux + dvy * uy;
    const R F = R(k) * extension + R(c) * rel;                              // :140
    fx = F * ux; fy = F * uy;
}

// derivative of one mass, gather form (every incident spring evaluated here)
template <bool S, int TOPO>
__device__ __forceinline__ void mass_derivative(const GridDev& g, const StencilIn& in, const size_t rl,
                                                const size_t col, Real<S> (&d)[4], Mass4& self_out) {
    using R = Real<S>;
    const size_t gr = g.row_begin + rl;
    const double* r0 = in.row((long long)rl, g.rows_local, g.cols);
    const Mass4 P = load_mass(r0, col);
    self_out = P;
    const bool has_l = col > 0, has_r = col + 1 < g.cols, has_u = gr > 0, has_d = gr + 1 < g.rows;
    const double* ru = has_u ? in.row((long long)rl - 1, g.rows_local, g.cols) : r0;
    const double* rd = has_d ? in.row((long long)rl + 1, g.rows_local, g.cols) : r0;
    R Fx(0.0), Fy(0.0), fx, fy;
    // horizontals then verticals (global edge order): left(-), right(+), up(-), down(+)
    if (has_l) { spring_force<S>(load_mass(r0, col - 1), P, g.k, g.c, g.L_orth, fx, fy); Fx -= fx; Fy -= fy; }
    if (has_r) { spring_force<S>(P, load_mass(r0, col + 1), g.k, g.c, g.L_orth, fx, fy); Fx += fx; Fy += fy; }
    if (has_u) { spring_force<S>(load_mass(ru, col), P, g.k, g.c, g.L_orth, fx, fy); Fx -= fx; Fy -= fy; }
    if (has_d) { spring_force<S>(P, load_mass(rd, col), g.k, g.c, g.L_orth, fx, fy); Fx += fx; Fy += fy; }
    if constexpr (TOPO == 1) {
        // all main diagonals (r,c)-(r+1,c+1), then all anti-diagonals (r,c)-(r+1,c-1)
        if (has_u && has_l) { spring_force<S>(load_mass(ru, col - 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_r) { spring_force<S>(P, load_mass(rd, col + 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
        if (has_u && has_r) { spring_force<S>(load_mass(ru, col + 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_l) { spring_force<S>(P, load_mass(rd, col - 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
    } else if constexpr (TOPO == 2) {
        // per cell, row-major: main (r,c)-(r+1,c+1) then anti (r,c+1)-(r+1,c)
        if (has_u && has_l) { spring_force<S>(load_mass(ru, col - 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_u && has_r) { spring_force<S>(load_mass(ru, col + 1), P, g.k, g.c, g.L_diag, fx, fy); Fx -= fx; Fy -= fy; }
        if (has_d && has_l) { spring_force<S>(P, load_mass(rd, col - 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
        if (has_d && has_r) { spring_force<S>(P, load_mass(rd, col + 1), g.k, g.c, g.L_diag, fx, fy); Fx += fx; Fy += fy; }
    }
    d[0] = R(P.vx); d[1] = R(P.vy);
    d[2] = Fx / R(g.mass); d[3] = Fy / R(g.mass);
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
    g.rows = p->rows; g.cols = p->cols; g.row_begin = p->row_begin; g.rows_local = p->rows_local;
    g.mass = p->mass; g.k = p->stiffness; g.c = p->damping;
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
int halo_exchange(sopot_b200_ctx* ctx, const GridDev& g, const double* buf, double* ghost_top, double* ghost_bot) {
    if (ctx->nranks == 1) return SOPOT_B200_OK;
    const size_t row_elems = g.cols * 4;
    const bool up = g.row_begin > 0, down = g.row_begin + g.rows_local < g.rows;
    const NcclApi* n = ctx->nccl;
    SB_NCCL(ctx, n->GroupStart());
    if (up) {
        SB_NCCL(ctx, n->Send(buf, row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->Recv(ghost_top, row_elems, SB_NCCL_FLOAT64, ctx->rank - 1, ctx->nccl_comm, ctx->stream));
    }
    if (down) {
        SB_NCCL(ctx, n->Send(buf + (g.rows_local - 1) * row_elems, row_elems, SB_NCCL_FLOAT64, ctx->rank + 1,
                             ctx->nccl_comm, ctx->stream));
        SB_NCCL(ctx, n->Recv(ghost_bot, row_elems, SB_NCCL_FLOAT64, ctx->rank + 1, ctx->nccl_comm, ctx->stream));
    }
    SB_NCCL(ctx, n->GroupEnd());
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
