/*
 * sopot_b200.h -- C ABI of the B200-native batched ODE integration engine for SOPOT.
 *
 * The reference (jedrzejmichalczyk/sopot) is a header-only C++20 template library with no FFI of
 * its own: the boundary below is what a host-language binding for the
 * RK4-ensemble hot path would bind.  Each entry point names the reference interface it replaces
 * (file:line under the reference tree).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *  - One ctx per host thread / GPU.  Every call is enqueued on the ctx stream and returns
 *    immediately; results are valid after sopot_b200_sync().  Return value: 0 on success, a
 *    SOPOT_B200_E* code otherwise, text in sopot_b200_last_error().  Nothing throws across the ABI
 *    (the reference throws std::invalid_argument / std::runtime_error; the C++ header layer
 *    re-throws from these codes).
 *  - Caller owns every buffer.  `mem` in the options says whether the pointers of that call are
 *    device (SOPOT_B200_MEM_DEVICE) or host (SOPOT_B200_MEM_HOST) addresses; host buffers are
 *    staged through a ctx-owned device arena (H2D before, D2H after, same stream).
 *  - Ensemble data is SoA: element i of component d lives at base[d * n + i].
 *  - Per-instance parameters: a parameter pointer addresses n doubles in the call's memory space, or --
 *    when its bit is set in `broadcast` (bit index = position of the field in the params struct) --
 *    ONE double in HOST memory, whatever `mem` says: it is read at call time and passed to the kernel
 *    by value.
 *  - fp_mode: SOPOT_B200_FP_CONTRACT lets the compiler fuse a*b+c into DFMA and replaces x/6.0
 *    style divisions by reciprocal multiplies (<= 1e-12 relative per step against the reference);
 *    SOPOT_B200_FP_STRICT evaluates every operation with the reference's association and IEEE
 *    rounding (no contraction): bit-identical to the reference wherever no libm transcendental is
 *    involved (oscillator, grid2d), and differing only by the libm last-ulp elsewhere.
 *  - status[i] (optional, int32): 0 ok, SOPOT_B200_ST_NONFINITE, SOPOT_B200_ST_SINGULAR
 *    (the reference throws "mass matrix is singular", physics/control/cart_n_pendulum.hpp:331-333).
 */
#ifndef SOPOT_B200_H
#define SOPOT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOPOT_B200_ABI_VERSION 5

enum {
    SOPOT_B200_OK = 0,
    SOPOT_B200_EINVAL = 1,      /* bad argument (reference: std::invalid_argument in ctors) */
    SOPOT_B200_ECUDA = 2,       /* CUDA runtime error */
    SOPOT_B200_ENCCL = 3,       /* NCCL error / NCCL not loadable */
    SOPOT_B200_ENOMEM = 4,
    SOPOT_B200_ENODEVICE = 5    /* no usable sm_100 device: there is NO CPU fallback */
};

enum { SOPOT_B200_MEM_DEVICE = 0, SOPOT_B200_MEM_HOST = 1 };
enum { SOPOT_B200_FP_CONTRACT = 0, SOPOT_B200_FP_STRICT = 1 };
/* grid2d: run the four RK stages as four kernels with one 1-row halo exchange each (the literal per-stage
 * schedule) instead of the default fused step kernel with one 4-row halo exchange per step */
enum { SOPOT_B200_FLAG_GRID_PER_STAGE = 1, SOPOT_B200_FLAG_UGRID_STAGE_PAIRS = 2,
       /* fused grid2d step, quad topology: force the row-marching kernel / the shared-memory tile kernel.  Default: by size
        * (tile kernel while the marching kernel would not fill the GPU, about <= 1024^2 per rank).  Same bits in FP_STRICT. */
       SOPOT_B200_FLAG_GRID_MARCH = 4, SOPOT_B200_FLAG_GRID_TILE = 8,
       /* unified grid, quad topology: force the marching stage kernel (every spring once) / the gather kernel for every stage.
        * Default: by arithmetic mode, integrator and global grid size.  Same bits in FP_STRICT. */
       SOPOT_B200_FLAG_UGRID_MARCH = 16, SOPOT_B200_FLAG_UGRID_GATHER = 32 };
enum { SOPOT_B200_ST_OK = 0, SOPOT_B200_ST_NONFINITE = 1, SOPOT_B200_ST_SINGULAR = 2,
       SOPOT_B200_ST_RANGE = 4 /* |angle| > 5e8 rad in FP_CONTRACT mode: rerun with FP_STRICT */,
       SOPOT_B200_ST_NOT_CONVERGED = 8 /* LQR::solve returned false (Riccati iteration diverged) */ };

typedef struct sopot_b200_ctx sopot_b200_ctx;

/* ---------------------------------------------------------------------------------------------
 * Context.  nccl_unique_id: NULL for a single-GPU ctx; otherwise the 128 bytes produced by
 * sopot_b200_nccl_unique_id() on rank 0 and distributed by the host (torch.distributed, MPI, ...).
 * NCCL is dlopen()ed (libnccl.so.2) only when nranks > 1.
 * ------------------------------------------------------------------------------------------- */
int  sopot_b200_abi_version(void);
int  sopot_b200_nccl_unique_id(void* out128);
int  sopot_b200_init(sopot_b200_ctx** out, int device, const void* nccl_unique_id, int rank, int nranks);
void sopot_b200_destroy(sopot_b200_ctx* ctx);
const char* sopot_b200_last_error(const sopot_b200_ctx* ctx);   /* ctx may be NULL: last init error */
int  sopot_b200_sync(sopot_b200_ctx* ctx);
/* adopt an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream */
int  sopot_b200_set_stream(sopot_b200_ctx* ctx, void* cuda_stream);
/* the ctx stream as a cudaStream_t, for kernels a host program instantiates itself from sopot/b200/generic_ensemble.cuh
 * (user-defined component systems compiled with nvcc); sopot_b200_note_launches adds them to the launch count */
int  sopot_b200_get_stream(sopot_b200_ctx* ctx, void** cuda_stream);
int  sopot_b200_note_launches(sopot_b200_ctx* ctx, uint64_t count);
int  sopot_b200_device_info(sopot_b200_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                            size_t* global_mem_bytes, char* name, size_t name_len);

/* device / pinned-host memory so that a host language without a CUDA binding can drive the engine */
int  sopot_b200_malloc(sopot_b200_ctx* ctx, size_t bytes, void** dptr);
int  sopot_b200_free(sopot_b200_ctx* ctx, void* dptr);
int  sopot_b200_host_alloc(sopot_b200_ctx* ctx, size_t bytes, void** hptr);   /* pinned */
int  sopot_b200_host_free(sopot_b200_ctx* ctx, void* hptr);
int  sopot_b200_memcpy_h2d(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes);
int  sopot_b200_memcpy_d2h(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes);
int  sopot_b200_memcpy_d2d(sopot_b200_ctx* ctx, void* dst, const void* src, size_t bytes);
int  sopot_b200_memset(sopot_b200_ctx* ctx, void* dst, int byte, size_t bytes);

/* CUDA-event timing on the ctx stream (slots 0..15) and launch accounting */
int  sopot_b200_event_record(sopot_b200_ctx* ctx, int slot);
int  sopot_b200_event_elapsed_ms(sopot_b200_ctx* ctx, int slot_begin, int slot_end, float* ms); /* syncs on slot_end */
uint64_t sopot_b200_launch_count(const sopot_b200_ctx* ctx);   /* kernels this library launched on ctx */

/* RK4Solver::solve loop count, core/solver.hpp:91,127: while (t < t_end - dt*0.5) t += dt */
size_t sopot_b200_rk4_num_steps(double t0, double t1, double dt);

/* ---------------------------------------------------------------------------------------------
 * Ensemble RK4: replaces createRK4Solver().solve(derivs, dim, t0, t1, dt, y0) (core/solver.hpp:63-140)
 * over TypedODESystem<double, ...>::computeDerivatives (core/typed_component.hpp:410-414), for n
 * independent instances at once.  state_out[dim][n] = state after n_steps steps.
 * traj_out (optional): snapshot s (0-based) = state after (s+1)*store_every steps, layout
 * [n_steps / store_every][dim][n].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t mem;          /* SOPOT_B200_MEM_* for every pointer of the call          */
    uint32_t fp_mode;      /* SOPOT_B200_FP_*                                         */
    size_t   store_every;  /* 0: final state only                                     */
    uint32_t broadcast;    /* bit i: i-th params pointer addresses a single double    */
    uint32_t flags;        /* SOPOT_B200_FLAG_*                                       */
} sopot_b200_rk4_opts;

/* physics::HarmonicOscillator<double>(mass, k, c, x0, v0), physics/harmonic_oscillator.hpp:44-85 */
typedef struct { const double *m, *k, *c, *x0, *v0; } sopot_b200_ho_params;
int sopot_b200_ho_rk4(sopot_b200_ctx* ctx, const sopot_b200_ho_params* p, size_t n, double t0, double dt,
                      size_t n_steps, const sopot_b200_rk4_opts* opts, double* state_out /*[2][n]*/,
                      double* traj_out, int32_t* status);
/* one derivative evaluation per instance (TypedODESystem::computeDerivatives), state/out [2][n] */
int sopot_b200_ho_rhs(sopot_b200_ctx* ctx, const sopot_b200_ho_params* p, size_t n,
                      const sopot_b200_rk4_opts* opts, const double* state, double* out);

/* pendulum::DoublePendulum<double>(m1, m2, L1, L2, g, th1, th2, w1, w2),
 * physics/pendulum/double_pendulum.hpp:68-179 */
typedef struct { const double *m1, *m2, *L1, *L2, *g, *th1, *th2, *w1, *w2; } sopot_b200_dpend_params;
int sopot_b200_dpend_rk4(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n, double t0,
                         double dt, size_t n_steps, const sopot_b200_rk4_opts* opts,
                         double* state_out /*[4][n]*/, double* traj_out, int32_t* status);
/* one derivative evaluation per instance (computeDerivatives), state/out [4][n] */
int sopot_b200_dpend_rhs(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n,
                         const sopot_b200_rk4_opts* opts, const double* state, double* out);
/* Batched 4x4 state Jacobians d f_i / d x_j of the double pendulum by forward-mode AD: computeJacobian
 * (core/typed_component.hpp:486-512) on DoublePendulum<Dual<double,4>> (tests/double_pendulum_test.cpp:383-412).
 * Reads m1, m2, L1, L2, g of `p` (arrays of n, or broadcast scalars) and state[4][n]; writes J[16][n], row-major 4x4 per
 * point, SoA over the points. */
int sopot_b200_dpend_jacobian(sopot_b200_ctx* ctx, const sopot_b200_dpend_params* p, size_t n,
                              const sopot_b200_rk4_opts* opts, const double* state, double* J);

/* control::CartNPendulum<1,double>(mc, {m}, {L}, g, {x,th,xd,w}) with a constant control force
 * (setControlForce), physics/control/cart_n_pendulum.hpp:83-221 */
typedef struct { const double *mc, *m, *L, *g, *x, *th, *xd, *w, *force; } sopot_b200_cartpole_params;
int sopot_b200_cartpole_rk4(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p, size_t n, double t0,
                            double dt, size_t n_steps, const sopot_b200_rk4_opts* opts,
                            double* state_out /*[4][n]*/, double* traj_out, int32_t* status);
int sopot_b200_cartpole_rhs(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p, size_t n,
                            const sopot_b200_rk4_opts* opts, const double* state, double* out);

/* ---------------------------------------------------------------------------------------------
 * Batched linearization: replaces makeLinearizer<4,1>(f).linearize(t, x, u)
 * (core/linearization.hpp:75-126) around CartNPendulum<1, Dual<double,5>> for P operating points.
 * x0 [4][P], u0 [P]; A [4*4][P] (A[i][j] at (i*4+j)*P + p), B [4][P].
 * opts->broadcast bits: 0 mc, 1 m, 2 L, 3 g.   opts->store_every is ignored.
 * ------------------------------------------------------------------------------------------- */
typedef struct { const double *mc, *m, *L, *g; } sopot_b200_cartpole_plant;
int sopot_b200_cartpole_linearize(sopot_b200_ctx* ctx, const sopot_b200_cartpole_plant* plant,
                                  const double* x0, const double* u0, size_t P,
                                  const sopot_b200_rk4_opts* opts, double* A, double* B, int32_t* status);

/* ---------------------------------------------------------------------------------------------
 * Batched LQR: replaces LQR<4,1>::solve(A, B, Q, R, dt, max_iter, tol) + getGain()
 * (physics/control/lqr.hpp:260-396) for P linearizations at once -- the consumer of the Jacobians above
 * ("A/B Jacobians for LQR sweep").  A [16][P], B [4][P] exactly as sopot_b200_cartpole_linearize writes
 * them; Q16 (row-major 4x4) and R are HOST values shared by the sweep.  K [4][P] (u = -K x); optional
 * Pric [16][P] (Riccati solution), iters [P] (iterations used).  status[p]: 0 = the reference's solve()
 * returned true, SOPOT_B200_ST_NOT_CONVERGED = it returned false, SOPOT_B200_ST_SINGULAR = it threw
 * "LQR solveForGain: ... singular".  The reference's defaults are riccati_dt 0.01, max_iter 1000, tol 1e-8.
 * ------------------------------------------------------------------------------------------- */
int sopot_b200_lqr4_gain(sopot_b200_ctx* ctx, const double* A, const double* B, size_t P, const double* Q16,
                         double R, double riccati_dt, size_t max_iter, double tol,
                         const sopot_b200_rk4_opts* opts, double* K, double* Pric, int32_t* iters, int32_t* status);

/* The same for LQR<state_dim,1> with state_dim 4, 6 or 14 (cart + 1, 2 or 6 links; the reference's own tests use
 * LQR<14,1>): A [state_dim^2][P], B [state_dim][P], Q [state_dim^2] host row-major, K [state_dim][P].  state_dim 4 runs
 * one point per thread, 6 and 14 one CTA of state_dim^2 threads per point with the matrices in shared memory. */
int sopot_b200_lqr_gain(sopot_b200_ctx* ctx, int state_dim, const double* A, const double* B, size_t P, const double* Q,
                        double R, double riccati_dt, size_t max_iter, double tol, const sopot_b200_rk4_opts* opts,
                        double* K, double* Pric, int32_t* iters, int32_t* status);

/* Closed-loop cart-pole ensemble: StateFeedbackController<4,1>::compute (state_feedback_controller.hpp:88-104,
 * u = -K (x - x_ref), clamped to [u_min, u_max] when saturate != 0) drives CartNPendulum<1,double> through
 * setControlForce; the force is evaluated ONCE per RK4 step from the state at the start of the step and held
 * over the four stages (zero-order hold, as wasm/wasm_inverted_pendulum.cpp:110-135 does), the step itself is
 * RK4Solver's (core/solver.hpp:95-125).  K [4][n] in the call's memory space (e.g. straight from
 * sopot_b200_lqr4_gain); x_ref [4][n] or NULL (= 0).  p->force is ignored.  stats_out [2][n] (optional):
 * max |theta| and max |u| over the run (the quantities tests/inverted_pendulum_test.cpp:241-292 asserts on).
 * broadcast bits as for sopot_b200_cartpole_rk4.
 * ------------------------------------------------------------------------------------------- */
typedef struct { const double* K; const double* x_ref; double u_min, u_max; int saturate; } sopot_b200_feedback;
int sopot_b200_cartpole_feedback_rk4(sopot_b200_ctx* ctx, const sopot_b200_cartpole_params* p,
                                     const sopot_b200_feedback* fb, size_t n, double t0, double dt, size_t n_steps,
                                     const sopot_b200_rk4_opts* opts, double* state_out, double* traj_out,
                                     double* stats_out, int32_t* status);

/* ---------------------------------------------------------------------------------------------
 * 6-DOF rocket Monte Carlo: replaces rocket::simulateRocket(rocket, t_end, dt) +
 * SimulationResult::computeStatistics (rocket/rocket.hpp:268-303, rocket/simulation_result.hpp:151-189)
 * for n trajectories of Rocket<double>::SystemType (rocket/rocket.hpp:73-85, 14 states).
 * Tables are shared by the ensemble and always HOST pointers, row-major:
 *   engine [engine_rows][8]: time, throat_d, exit_d, p_comb, T_comb, gamma, mol_mass, efficiency
 *          (rocket/interpolated_engine.hpp:66-79; the isentropic precompute of :176-252 runs here,
 *          on the host, once);  mass/ix/iyz [rows][2]: time, value (rocket/rocket_body.hpp:55-85);
 *   rows == 0 means "not loaded" (constants 100 / 10 / 100, rocket_body.hpp:36-39; no thrust).
 * Per trajectory (broadcast bits 0..3): launcher elevation / azimuth [deg] (Rocket::setLauncher),
 * reference diameter [m] (setDiameter), gravity (Gravity<T>(g)).
 * state_out [14][n]; stats_out [8][n] (optional): apogee, apogee_time, max_speed, max_speed_time,
 * burnout_altitude, burnout_speed, final East, final North.
 * ------------------------------------------------------------------------------------------- */
/* 2-D coefficient table of io::BilinearInterpolator (io/interpolation.hpp:77-181): row_vals[rows], col_vals[cols],
 * data[rows][cols] row-major, all HOST pointers; rows == 0 means "not loaded".  Aerodynamic tables are laid out as the
 * reference reads them: rows = Mach, columns = angle of attack / sideslip in degrees for Cd, Cl, Cs
 * (rocket/axisymmetric_aero.hpp:133-167, interpolate(mach, angle_deg)); rows = signed angle of attack in degrees,
 * columns = Mach for the damping coefficient cmq_yz (:276-281, interpolate(aoa_deg, mach)). */
typedef struct { size_t rows, cols; const double *row_vals, *col_vals, *data; } sopot_b200_table2d;
typedef struct {
    sopot_b200_table2d cd_power_on, cd_power_off, cl_power_on, cl_power_off, cs_power_on, cs_power_off, cmq_yz;
    int engine_on;      /* AxisymmetricAerodynamics::setEngineOn; the reference's default (never changed by Rocket<T>) is 1 */
} sopot_b200_aero_tables;
typedef struct {
    const double *elevation_deg, *azimuth_deg, *diameter, *gravity;
    size_t engine_rows; const double* engine;
    size_t mass_rows;   const double* mass;
    size_t ix_rows;     const double* ix;
    size_t iyz_rows;    const double* iyz;
    const sopot_b200_aero_tables* aero;   /* NULL: no coefficient tables (Cd 0.3, Cl = Cs = 0, cmq -450) */
} sopot_b200_rocket_params;
int sopot_b200_rocket_rk4(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n, double dt,
                          size_t n_steps, const sopot_b200_rk4_opts* opts, double* state_out,
                          double* traj_out, double* stats_out, int32_t* status);
/* one derivative evaluation per instance, state/out [14][n] (per-trajectory params as above) */
int sopot_b200_rocket_rhs(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n,
                          const sopot_b200_rk4_opts* opts, const double* state, double* out);
/* The reference's state functions (rocket/rocket_tags.hpp) that are intermediates of the derivative, evaluated at
 * state[14][n] in strict arithmetic: replaces Rocket<T>::queryStateFunction<Tag> (rocket/rocket.hpp:222-228) and, through
 * it, SimulationResult::query<Tag>(t) (rocket/simulation_result.hpp:111-115).  out[SOPOT_B200_ROCKET_PROBE_VALUES][n], rows:
 *   0 dynamics::Mass | 1-3 dynamics::MomentOfInertia (Ix, Iyz, Iyz) | 4-6 dynamics::TotalForceENU | 7-9 dynamics::TotalTorqueBody
 *   10 dynamics::GravityAcceleration | 11-13 aero::AeroForceBody | 14-16 aero::AeroForceENU | 17-19 aero::AeroMomentBody
 *   20 environment::AtmosphericDensity | 21 ::AtmosphericPressure | 22 ::AtmosphericTemperature | 23 ::SpeedOfSound
 *   24 propulsion::MassFlowRate | 25-27 propulsion::ThrustForceBody
 * (kinematics::* and propulsion::Time are slices of the state vector and need no kernel.) */
#define SOPOT_B200_ROCKET_PROBE_VALUES 28
int sopot_b200_rocket_state_functions(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n,
                                      const sopot_b200_rk4_opts* opts, const double* state, double* out);

/* Dispersion statistics over per-trajectory rows (e.g. stats_out above): for each of `rows` rows of
 * values[rows][n_local]: count, sum, sum of squares, min, max -- reduced on the device, then summed /
 * min-maxed over all ranks of the ctx with NCCL (one allreduce(sum) + one allreduce(min) + one (max)).
 * out[rows] on the HOST, identical on every rank.  Non-finite entries are skipped and counted. */
typedef struct { double count, sum, sumsq, min, max, mean, stddev, nonfinite; } sopot_b200_dispersion;
int sopot_b200_dispersion_stats(sopot_b200_ctx* ctx, const double* values, size_t rows, size_t n_local,
                                uint32_t mem, sopot_b200_dispersion* out);

/* ---------------------------------------------------------------------------------------------
 * grid2d mass-spring cloth: replaces RK4Solver::solve over makeGrid2DSystem<double,R,C,Diag> /
 * makeTriangularGridSystem (physics/connected_masses/connectivity_matrix_2d.hpp:156-266) at
 * runtime sizes.  State is AoS [rows][cols][x,y,vx,vy], row-major (grid_2d.hpp:97-99).
 * topology: 0 quad, 1 quad + diagonals (IncludeDiagonals=true), 2 triangular.
 * Multi-GPU: each rank of the ctx owns the contiguous slab rows [row_begin, row_begin + rows_local)
 * of the global grid and `state` addresses only that slab; one boundary row is exchanged with each
 * neighbour before every RK stage (NCCL send/recv over NVLink) with SOPOT_B200_FLAG_GRID_PER_STAGE; by
 * default the whole RK4 step is one kernel (all four stages on a shared-memory tile with a 4-cell halo) and
 * FOUR boundary rows are exchanged once per step -- same values bit for bit, a quarter of the messages and
 * 64 instead of 544 bytes of HBM traffic per mass and step.
 * With the quad topology the fused step on several ranks carries the exchange itself: ghost rows and sequence flags live in
 * buffers that the neighbouring ranks map with CUDA IPC, and the step kernel stores its boundary rows straight into them over
 * NVLink (no collective call per step).  Set up on first use; if any rank cannot (no IPC / peer access) all ranks agree to use
 * ncclSend/ncclRecv instead.  A timed-out wait is reported by the next sopot_b200_sync().  Environment switches for
 * diagnosis: SOPOT_B200_NO_P2P=1 (NCCL schedule), SOPOT_B200_NO_HALO_OVERLAP=1, SOPOT_B200_GRID_TILE_KERNEL=1.
 * Slab contract (nranks > 1): rank r must own exactly the r-th block of the even contiguous partition of the
 * rows (base = rows / nranks rows each, the first rows % nranks ranks one more; sopot_b200/distributed.py::slab).
 * Every rank derives its exchange schedule from rows / nranks alone, so any other decomposition -- or fewer rows
 * than ranks -- is rejected with SOPOT_B200_EINVAL before the first exchange instead of hanging in it.  The same
 * holds for the unified grid below.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    size_t rows, cols;            /* GLOBAL grid size                         */
    size_t row_begin, rows_local; /* this rank's slab (whole grid if 1 rank)  */
    int    topology;
    double mass, spacing, stiffness, damping;
} sopot_b200_grid2d_params;
int sopot_b200_grid2d_initial_state(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* g, uint32_t mem,
                                    double* state);
int sopot_b200_grid2d_rk4(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* g, double dt, size_t n_steps,
                          const sopot_b200_rk4_opts* opts, double* state /* in/out */);
/* one derivative evaluation (computeDerivatives) of the local slab; single-rank ctx only */
int sopot_b200_grid2d_rhs(sopot_b200_ctx* ctx, const sopot_b200_grid2d_params* g,
                          const sopot_b200_rk4_opts* opts, const double* state, double* out);

/* ---------------------------------------------------------------------------------------------
 * "Unified" 6-state grid: PointMass2D [x, y, vx, vy, theta, omega] + Spring2D with attachment points on the rim of each
 * mass and torque coupling (physics/unified/components.hpp:52-138, 206-291), evaluated like
 * ValidatedSystem::computeDerivatives (physics/unified/validated_system.hpp:204-301) on makeGridTopology<R,C>
 * (constexpr_topology.hpp:368-414) -- the system behind the reference's Grid2DSimulator (wasm/wasm_grid2d.cpp:223).
 * State is AoS [rows][cols][6].  Attachment angles: horizontal springs h_attach0 on the left mass / h_attach1 on the
 * right one, vertical springs v_attach0 on the upper / v_attach1 on the lower mass (the wasm wrapper uses 0, pi, pi/2,
 * -pi/2; Spring2D's defaults are 0, pi for every spring).  repulsion_stiffness <= 0 means 10 * stiffness; min_distance
 * 0 disables the repulsion.  Multi-GPU like grid2d: each rank passes its slab of rows (row_begin, rows_local) and its
 * local state [rows_local][cols][6]; boundary rows travel over NCCL after every kernel that produces a stencil input.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    size_t rows, cols;
    double mass, radius, spacing;
    double stiffness, rest_length, damping, min_distance, repulsion_stiffness;
    double h_attach0, h_attach1, v_attach0, v_attach1;
    size_t row_begin, rows_local;   /* this rank's slab of rows; rows_local == 0: the whole grid (single rank) */
    /* topology 1 = makeTriangleGridTopology (constexpr_topology.hpp:436-505, the simulator's "triangle" grid type,
     * wasm/wasm_grid2d.cpp:466-525): after the horizontal and the vertical springs every cell adds its main diagonal
     * (r,c)->(r+1,c+1) and its anti-diagonal (r,c+1)->(r+1,c), with their own rest length and attachment angles
     * (the wrapper uses sqrt(2) * spacing and pi/4, -3pi/4, 3pi/4, -pi/4).  0 = makeGridTopology. */
    int topology;
    double diag_rest_length;
    double d_attach0, d_attach1, a_attach0, a_attach1;
} sopot_b200_ugrid_params;
int sopot_b200_ugrid_initial_state(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* g, uint32_t mem, double* state);
int sopot_b200_ugrid_rhs(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* g, const sopot_b200_rk4_opts* opts,
                         const double* state, double* out);
/* RK4: one kernel per stage (816 B per mass and step, 0.9 of the HBM roofline in contract mode).  With
 * SOPOT_B200_FLAG_UGRID_STAGE_PAIRS in opts->flags (single rank) two stages run per kernel on a shared-memory tile with a
 * 2-cell halo (336 B per mass and step); bit-identical in strict mode, measured slower on B200 (barrier / latency bound),
 * kept as an option. */
int sopot_b200_ugrid_rk4(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* g, double dt, size_t n_steps,
                         const sopot_b200_rk4_opts* opts, double* state /* in/out */);
/* the three integrators of the reference's Grid2DSimulator (wasm/wasm_grid2d.cpp:286-413): RK4 as above; symplectic Euler
 * (v += dt a(x_n, v_n), then x += dt v_new; one derivative evaluation per step); velocity Verlet (x += dt v + dt^2/2 a_n,
 * a_{n+1} = a(x_{n+1}, v_n), v += dt/2 (a_n + a_{n+1}); two evaluations per step) */
enum { SOPOT_B200_INTEGRATOR_RK4 = 0, SOPOT_B200_INTEGRATOR_SYMPLECTIC_EULER = 1, SOPOT_B200_INTEGRATOR_VELOCITY_VERLET = 2 };
int sopot_b200_ugrid_integrate(sopot_b200_ctx* ctx, const sopot_b200_ugrid_params* g, int integrator, double dt,
                               size_t n_steps, const sopot_b200_rk4_opts* opts, double* state /* in/out */);

/* ---------------------------------------------------------------------------------------------
 * Roofline denominators measured on this device: dependent-free DFMA issue rate (FP64 TFLOP/s,
 * FMA = 2 flop) and a device copy (GB/s, read + write).  Used by bench.py; MEASURED_PEAKS.json has
 * no FP64 figure.
 * ------------------------------------------------------------------------------------------- */
int sopot_b200_measure_fp64_peak(sopot_b200_ctx* ctx, double* tflops);
int sopot_b200_measure_hbm_peak(sopot_b200_ctx* ctx, size_t bytes, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* SOPOT_B200_H */
