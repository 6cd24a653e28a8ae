/*
 * oracle/sopot_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU algorithm for the RK4-ensemble hot path
 * (SURVEY.md section 8a).  Nothing under sopot_b200/ may include, link or call this file;
 * it is the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Build: oracle/Makefile (gcc -O3 -ffp-contract=off, glibc libm --
 * the same arithmetic as the reference's g++ -std=c++20 -O3 build, CMakeLists.txt:7-9,33-35).
 *
 * PINNING: every function below is checked bit-for-bit against the reference itself compiled
 * in the build container (oracle/_ref/libsopot_ref.so, built by oracle/Makefile from the sources
 * under /root/reference through oracle/ref_harness.cpp) by tests/test_oracle.py, and against the
 * golden vectors those runs produced (tests/golden/, generator tests/golden/make_golden.py).
 * The reference's own tests hold no numeric golden vectors for the rocket or for grids
 * larger than 6x6; for those the pin is the reference run here.
 *
 * All citations are file:line under /root/reference.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* minimal pthread fan-out (the image's gcc ships without libgomp): body(user, begin, end) over
 * [0, n) in dynamically claimed chunks.  Threads are harness plumbing, not reference behaviour:
 * the reference is single-threaded. */
typedef void (*orc_range_fn)(void* user, size_t begin, size_t end);
typedef struct { orc_range_fn fn; void* user; size_t n, chunk; size_t next; pthread_mutex_t mu; } orc_pf_t;

static void* orc_pf_worker(void* arg) {
    orc_pf_t* pf = (orc_pf_t*)arg;
    for (;;) {
        pthread_mutex_lock(&pf->mu);
        size_t b = pf->next; pf->next += pf->chunk;
        pthread_mutex_unlock(&pf->mu);
        if (b >= pf->n) break;
        size_t e = b + pf->chunk < pf->n ? b + pf->chunk : pf->n;
        pf->fn(pf->user, b, e);
    }
    return NULL;
}

static void orc_parallel_for(size_t n, size_t chunk, int nthreads, orc_range_fn fn, void* user) {
    if (nthreads <= 1 || n <= chunk) { if (n) fn(user, 0, n); return; }
    if (nthreads > 256) nthreads = 256;
    orc_pf_t pf; pf.fn = fn; pf.user = user; pf.n = n; pf.chunk = chunk ? chunk : 1; pf.next = 0;
    pthread_mutex_init(&pf.mu, NULL);
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads - 1; ++t)
        if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) ++started;
    orc_pf_worker(&pf);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    pthread_mutex_destroy(&pf.mu);
}

/* ------------------------------------------------------------------------------------------
 * RK4Solver::solve, core/solver.hpp:63-140.
 *   k_i = dt * f(...)            (:95-120, each derivative scaled in place by dt)
 *   tmp = y + 0.5*k1, y + 0.5*k2, y + k3
 *   y  += (k1 + 2.0*k2 + 2.0*k3 + k4) / 6.0          (:123-125, left-to-right)
 *   t  += dt  (accumulated, :127);  loop while (t < t_end - dt*0.5)   (:91)
 * ------------------------------------------------------------------------------------------ */
typedef void (*orc_rhs_fn)(const void* ctx, double t, const double* y, double* dy);

/* returns number of steps taken; snap(user, step_index, t, y) is called after every step */
typedef void (*orc_snap_fn)(void* user, size_t step, double t, const double* y);

static size_t orc_rk4(orc_rhs_fn f, const void* ctx, size_t dim, double t0, double t1, double dt,
                      double* y, double* work /* 5*dim */, orc_snap_fn snap, void* user,
                      double* t_last) {
    double *k1 = work, *k2 = work + dim, *k3 = work + 2 * dim, *k4 = work + 3 * dim,
           *tmp = work + 4 * dim;
    double t = t0;
    size_t step = 0;
    while (t < t1 - dt * 0.5) {
        f(ctx, t, y, k1);
        for (size_t i = 0; i < dim; ++i) { k1[i] *= dt; tmp[i] = y[i] + 0.5 * k1[i]; }
        f(ctx, t + 0.5 * dt, tmp, k2);
        for (size_t i = 0; i < dim; ++i) { k2[i] *= dt; tmp[i] = y[i] + 0.5 * k2[i]; }
        f(ctx, t + 0.5 * dt, tmp, k3);
        for (size_t i = 0; i < dim; ++i) { k3[i] *= dt; tmp[i] = y[i] + k3[i]; }
        f(ctx, t + dt, tmp, k4);
        for (size_t i = 0; i < dim; ++i) k4[i] *= dt;
        for (size_t i = 0; i < dim; ++i)
            y[i] += (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
        t += dt;
        ++step;
        if (snap) snap(user, step, t, y);
    }
    if (t_last) *t_last = t;
    return step;
}

/* number of steps RK4Solver::solve takes (core/solver.hpp:91,127: accumulated t) */
ORC_API size_t orc_rk4_num_steps(double t0, double t1, double dt) {
    double t = t0;
    size_t n = 0;
    while (t < t1 - dt * 0.5) { t += dt; ++n; }
    return n;
}

/* strided trajectory capture: traj[(snap*dim + d)*n + i] */
typedef struct {
    double* traj; size_t dim, n, i, every;
} orc_traj_t;

static void orc_traj_snap(void* user, size_t step, double t, const double* y) {
    orc_traj_t* tr = (orc_traj_t*)user;
    (void)t;
    if (!tr->traj || !tr->every || step % tr->every) return;
    size_t s = step / tr->every - 1;
    for (size_t d = 0; d < tr->dim; ++d) tr->traj[(s * tr->dim + d) * tr->n + tr->i] = y[d];
}

/* ------------------------------------------------------------------------------------------
 * Config 1 -- HarmonicOscillator::derivatives, physics/harmonic_oscillator.hpp:75-85
 *   {v, T(-k/m)*x - T(c/m)*v}
 * ------------------------------------------------------------------------------------------ */
typedef struct { double m, k, c; } orc_ho_t;

static void orc_ho_rhs(const void* ctx, double t, const double* y, double* dy) {
    const orc_ho_t* p = (const orc_ho_t*)ctx;
    (void)t;
    double x = y[0], v = y[1];
    dy[0] = v;
    dy[1] = (-p->k / p->m) * x - (p->c / p->m) * v;
}

typedef struct {
    size_t n; const double *m, *k, *c, *x0, *v0; double t0, t1, dt; size_t store_every;
    double *final_out, *traj_out, *t_last;
} orc_ho_job_t;

static void orc_ho_range(void* user, size_t b, size_t e) {
    const orc_ho_job_t* j = (const orc_ho_job_t*)user;
    for (size_t i = b; i < e; ++i) {
        orc_ho_t p = {j->m[i], j->k[i], j->c[i]};
        double y[2] = {j->x0[i], j->v0[i]}, work[10], tl;
        orc_traj_t tr = {j->traj_out, 2, j->n, i, j->store_every};
        orc_rk4(orc_ho_rhs, &p, 2, j->t0, j->t1, j->dt, y, work, orc_traj_snap, &tr, &tl);
        j->final_out[i] = y[0];
        j->final_out[j->n + i] = y[1];
        if (j->t_last) j->t_last[i] = tl;
    }
}

ORC_API int orc_ho_rk4(size_t n, const double* m, const double* k, const double* c,
                       const double* x0, const double* v0, double t0, double t1, double dt,
                       size_t store_every, double* final_out /*[2][n]*/, double* traj_out,
                       double* t_last, int nthreads) {
    orc_ho_job_t j = {n, m, k, c, x0, v0, t0, t1, dt, store_every, final_out, traj_out, t_last};
    orc_parallel_for(n, 64, nthreads, orc_ho_range, &j);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Config 2 -- DoublePendulum::derivatives, physics/pendulum/double_pendulum.hpp:126-179
 * ------------------------------------------------------------------------------------------ */
typedef struct { double m1, m2, L1, L2, g; } orc_dp_t;

static void orc_dpend_rhs_fn(const void* ctx, double t, const double* y, double* dy) {
    const orc_dp_t* p = (const orc_dp_t*)ctx;
    (void)t;
    double theta1 = y[0], theta2 = y[1], omega1 = y[2], omega2 = y[3];
    double delta = theta1 - theta2;                                   /* :145 */
    double sin_delta = sin(delta), cos_delta = cos(delta);
    double sin_theta1 = sin(theta1), sin_theta2 = sin(theta2);
    double m1 = p->m1, m2 = p->m2, L1 = p->L1, L2 = p->L2, g = p->g;
    double a11 = (m1 + m2) * L1;                                      /* :162-165 */
    double a12 = m2 * L2 * cos_delta;
    double a21 = L1 * cos_delta;
    double a22 = L2;
    double b1 = m2 * L2 * omega2 * omega2 * sin_delta - (m1 + m2) * g * sin_theta1; /* :168 */
    double b2 = -L1 * omega1 * omega1 * sin_delta - g * sin_theta2;                 /* :169 */
    double det = a11 * a22 - a12 * a21;                               /* :172 */
    dy[0] = omega1;
    dy[1] = omega2;
    dy[2] = (b1 * a22 - b2 * a12) / det;                              /* :174 */
    dy[3] = (a11 * b2 - a21 * b1) / det;                              /* :175 */
}

ORC_API int orc_dpend_rhs(size_t n, const double* m1, const double* m2, const double* L1,
                          const double* L2, const double* g, const double* state /*[4][n]*/,
                          double* out /*[4][n]*/) {
    for (size_t i = 0; i < n; ++i) {
        orc_dp_t p = {m1[i], m2[i], L1[i], L2[i], g[i]};
        double y[4] = {state[i], state[n + i], state[2 * n + i], state[3 * n + i]}, d[4];
        orc_dpend_rhs_fn(&p, 0.0, y, d);
        for (int j = 0; j < 4; ++j) out[j * n + i] = d[j];
    }
    return 0;
}

typedef struct {
    size_t n; const double *m1, *m2, *L1, *L2, *g, *th1, *th2, *w1, *w2; double t0, t1, dt;
    size_t store_every; double *final_out, *traj_out, *t_last;
} orc_dp_job_t;

static void orc_dpend_range(void* user, size_t b, size_t e) {
    const orc_dp_job_t* j = (const orc_dp_job_t*)user;
    for (size_t i = b; i < e; ++i) {
        orc_dp_t p = {j->m1[i], j->m2[i], j->L1[i], j->L2[i], j->g[i]};
        double y[4] = {j->th1[i], j->th2[i], j->w1[i], j->w2[i]}, work[20], tl;
        orc_traj_t tr = {j->traj_out, 4, j->n, i, j->store_every};
        orc_rk4(orc_dpend_rhs_fn, &p, 4, j->t0, j->t1, j->dt, y, work, orc_traj_snap, &tr, &tl);
        for (int q = 0; q < 4; ++q) j->final_out[q * j->n + i] = y[q];
        if (j->t_last) j->t_last[i] = tl;
    }
}

ORC_API int orc_dpend_rk4(size_t n, const double* m1, const double* m2, const double* L1,
                          const double* L2, const double* g, const double* th1, const double* th2,
                          const double* w1, const double* w2, double t0, double t1, double dt,
                          size_t store_every, double* final_out /*[4][n]*/, double* traj_out,
                          double* t_last, int nthreads) {
    orc_dp_job_t j = {n, m1, m2, L1, L2, g, th1, th2, w1, w2, t0, t1, dt, store_every,
                      final_out, traj_out, t_last};
    orc_parallel_for(n, 16, nthreads, orc_dpend_range, &j);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Config 3 -- Dual<double,5> (core/dual.hpp:15-157, sin/cos :181-201) and
 * CartNPendulum<1,T>::derivatives (physics/control/cart_n_pendulum.hpp:143-221) with
 * solveLinearSystem (:319-359), driven by SystemLinearizer<4,1>::linearize
 * (core/linearization.hpp:75-114).
 * ------------------------------------------------------------------------------------------ */
#define DN 5
typedef struct { double v; double d[DN]; } dual_t;

static dual_t d_const(double v) { dual_t r; r.v = v; for (int i = 0; i < DN; ++i) r.d[i] = 0.0; return r; }
static dual_t d_var(double v, int idx) { dual_t r = d_const(v); if (idx < DN) r.d[idx] = 1.0; return r; }
static dual_t d_add(dual_t a, dual_t b) {           /* core/dual.hpp:66-72 */
    dual_t r; r.v = a.v + b.v; for (int i = 0; i < DN; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
static dual_t d_sub(dual_t a, dual_t b) {           /* :74-80 */
    dual_t r; r.v = a.v - b.v; for (int i = 0; i < DN; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
static dual_t d_mul(dual_t a, dual_t b) {           /* :82-89  u*dv + du*v */
    dual_t r; r.v = a.v * b.v;
    for (int i = 0; i < DN; ++i) r.d[i] = a.v * b.d[i] + a.d[i] * b.v;
    return r; }
static dual_t d_div(dual_t a, dual_t b) {           /* :91-99 */
    dual_t r; double inv_v2 = 1.0 / (b.v * b.v);
    for (int i = 0; i < DN; ++i) r.d[i] = (b.v * a.d[i] - a.v * b.d[i]) * inv_v2;
    r.v = a.v / b.v; return r; }
static dual_t d_sin(dual_t x) {                     /* :181-190 */
    dual_t r; double c = cos(x.v);
    for (int i = 0; i < DN; ++i) r.d[i] = c * x.d[i];
    r.v = sin(x.v); return r; }
static dual_t d_cos(dual_t x) {                     /* :192-201 */
    dual_t r; double ns = -sin(x.v);
    for (int i = 0; i < DN; ++i) r.d[i] = ns * x.d[i];
    r.v = cos(x.v); return r; }

/* computeJacobian (core/typed_component.hpp:486-512) on DoublePendulum<Dual<double,4>> (physics/pendulum/
 * double_pendulum.hpp:126-179; the reference's test: tests/double_pendulum_test.cpp:383-412): state variable i carries
 * seed i, J[i][j] = derivative j of f_i.  dual_t holds DN = 5 partials; partials never mix, so the first four are exactly
 * those of Dual<double,4>.  J is [16][n], row-major 4x4 per point, SoA over the points. */
ORC_API void orc_dpend_jacobian(size_t n, const double* m1_, const double* m2_, const double* L1_, const double* L2_,
                                const double* g_, const double* state /*[4][n]*/, double* J /*[16][n]*/) {
    for (size_t p = 0; p < n; ++p) {
        dual_t y[4];
        for (int i = 0; i < 4; ++i) y[i] = d_var(state[(size_t)i * n + p], i);
        const dual_t theta1 = y[0], theta2 = y[1], omega1 = y[2], omega2 = y[3];
        const dual_t delta = d_sub(theta1, theta2);
        const dual_t sin_delta = d_sin(delta), cos_delta = d_cos(delta);
        const dual_t sin_theta1 = d_sin(theta1), sin_theta2 = d_sin(theta2);
        const dual_t m1 = d_const(m1_[p]), m2 = d_const(m2_[p]), L1 = d_const(L1_[p]), L2 = d_const(L2_[p]), g = d_const(g_[p]);
        const dual_t a11 = d_mul(d_add(m1, m2), L1);
        const dual_t a12 = d_mul(d_mul(m2, L2), cos_delta);
        const dual_t a21 = d_mul(L1, cos_delta);
        const dual_t a22 = L2;
        const dual_t b1 = d_sub(d_mul(d_mul(d_mul(d_mul(m2, L2), omega2), omega2), sin_delta), d_mul(d_mul(d_add(m1, m2), g), sin_theta1));
        const dual_t negL1 = d_sub(d_const(0.0), L1);                  /* unary minus of the reference: -L1 (value -L1) */
        const dual_t b2 = d_sub(d_mul(d_mul(d_mul(negL1, omega1), omega1), sin_delta), d_mul(g, sin_theta2));
        const dual_t det = d_sub(d_mul(a11, a22), d_mul(a12, a21));
        const dual_t f[4] = {omega1, omega2, d_div(d_sub(d_mul(b1, a22), d_mul(b2, a12)), det),
                             d_div(d_sub(d_mul(a11, b2), d_mul(a21, b1)), det)};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) J[(size_t)(i * 4 + j) * n + p] = f[i].d[j];
    }
}

/* CartNPendulum<1, Dual5>::derivatives. status: 0 ok, 1 singular mass matrix (:331-333 throws) */
static int orc_cartpole_dual_rhs(double mc, double m, double L, double g_, const dual_t local[4],
                                 dual_t F, dual_t deriv[4]) {
    const dual_t g = d_const(g_);
    dual_t theta = local[1];
    dual_t s = d_sin(theta), c = d_cos(theta);                 /* :155-161 */
    dual_t omega = local[3];
    double Mtail = 0.0; Mtail += m;                            /* tailMass(0), :134-138 */
    dual_t M[2][2], rhs[2];
    dual_t total_mass = d_const(mc);                           /* :172-174 */
    total_mass = d_add(total_mass, d_const(m));
    M[0][0] = total_mass;
    rhs[0] = F;                                                /* :176 */
    dual_t coupling = d_mul(d_mul(d_const(Mtail), d_const(L)), c);   /* :178 */
    M[0][1] = coupling; M[1][0] = coupling;
    rhs[0] = d_add(rhs[0], d_mul(d_mul(d_mul(d_mul(d_const(Mtail), d_const(L)), s), omega), omega)); /* :181 */
    dual_t rhs_a = d_mul(d_mul(d_mul(d_const(Mtail), g), d_const(L)), s);   /* :187 */
    dual_t cab = d_add(d_mul(c, c), d_mul(s, s));              /* :191 */
    M[1][1] = d_mul(d_mul(d_mul(d_const(Mtail), d_const(L)), d_const(L)), cab);  /* :192-193 */
    rhs[1] = rhs_a;
    /* solveLinearSystem, :319-359, num_dof = 2 */
    dual_t A[2][2] = {{M[0][0], M[0][1]}, {M[1][0], M[1][1]}}, b[2] = {rhs[0], rhs[1]};
    for (int col = 0; col < 2; ++col) {
        int pivot = col;
        double best = fabs(A[col][col].v);
        for (int row = col + 1; row < 2; ++row) {
            double mag = fabs(A[row][col].v);
            if (mag > best) { best = mag; pivot = row; }
        }
        if (best < 1e-15) return 1;
        if (pivot != col) {
            for (int j = 0; j < 2; ++j) { dual_t t = A[col][j]; A[col][j] = A[pivot][j]; A[pivot][j] = t; }
            dual_t t = b[col]; b[col] = b[pivot]; b[pivot] = t;
        }
        for (int row = col + 1; row < 2; ++row) {
            dual_t factor = d_div(A[row][col], A[col][col]);
            for (int j = col; j < 2; ++j) A[row][j] = d_sub(A[row][j], d_mul(factor, A[col][j]));
            b[row] = d_sub(b[row], d_mul(factor, b[col]));
        }
    }
    dual_t x[2];
    for (int i = 1; i >= 0; --i) {
        dual_t sum = b[i];
        for (int j = i + 1; j < 2; ++j) sum = d_sub(sum, d_mul(A[i][j], x[j]));
        x[i] = d_div(sum, A[i][i]);
    }
    deriv[0] = local[2]; deriv[1] = local[3]; deriv[2] = x[0]; deriv[3] = x[1];   /* :210-219 */
    return 0;
}

typedef struct {
    size_t P; const double *x, *u, *mc, *m, *L, *g; double *A, *B; int* status;
} orc_lin_job_t;

static void orc_lin_range(void* user, size_t b, size_t e) {
    const orc_lin_job_t* j = (const orc_lin_job_t*)user;
    const size_t P = j->P;
    for (size_t p = b; p < e; ++p) {
        dual_t xs[4], d[4];
        for (int i = 0; i < 4; ++i) xs[i] = d_var(j->x[i * P + p], i);    /* linearization.hpp:86-89 */
        dual_t us = d_var(j->u[p], 4);                                    /* :92-95 */
        int st = orc_cartpole_dual_rhs(j->mc[p], j->m[p], j->L[p], j->g[p], xs, us, d);
        if (j->status) j->status[p] = st;
        for (int i = 0; i < 4; ++i) {
            for (int q = 0; q < 4; ++q) j->A[(i * 4 + q) * P + p] = st ? NAN : d[i].d[q];  /* :102-111 */
            j->B[i * P + p] = st ? NAN : d[i].d[4];
        }
    }
}

ORC_API int orc_cartpole_linearize(size_t P, const double* x /*[4][P]*/, const double* u /*[P]*/,
                                   const double* mc, const double* m, const double* L,
                                   const double* g, double* A /*[16][P]*/, double* B /*[4][P]*/,
                                   int* status, int nthreads) {
    orc_lin_job_t j = {P, x, u, mc, m, L, g, A, B, status};
    orc_parallel_for(P, 1024, nthreads, orc_lin_range, &j);
    return 0;
}

/* plain-double CartNPendulum<1,double> with constant control force */
typedef struct { double mc, m, L, g, F; int status; } orc_cp_t;

static void orc_cartpole_rhs_fn(const void* ctx, double t, const double* y, double* dy) {
    orc_cp_t* p = (orc_cp_t*)ctx;
    (void)t;
    double s = sin(y[1]), c = cos(y[1]), omega = y[3];
    double Mtail = 0.0; Mtail += p->m;
    double total = p->mc; total = total + p->m;
    double A[2][2], b[2];
    A[0][0] = total;
    double coupling = Mtail * p->L * c;
    A[0][1] = coupling; A[1][0] = coupling;
    b[0] = p->F + Mtail * p->L * s * omega * omega;
    double rhs_a = Mtail * p->g * p->L * s;
    double cab = c * c + s * s;
    A[1][1] = Mtail * p->L * p->L * cab;
    b[1] = rhs_a;
    for (int col = 0; col < 2; ++col) {
        int pivot = col; double best = fabs(A[col][col]);
        for (int row = col + 1; row < 2; ++row) { double mag = fabs(A[row][col]); if (mag > best) { best = mag; pivot = row; } }
        if (best < 1e-15) { p->status = 1; dy[0] = dy[1] = dy[2] = dy[3] = NAN; return; }
        if (pivot != col) {
            for (int j = 0; j < 2; ++j) { double tt = A[col][j]; A[col][j] = A[pivot][j]; A[pivot][j] = tt; }
            double tt = b[col]; b[col] = b[pivot]; b[pivot] = tt;
        }
        for (int row = col + 1; row < 2; ++row) {
            double factor = A[row][col] / A[col][col];
            for (int j = col; j < 2; ++j) A[row][j] = A[row][j] - factor * A[col][j];
            b[row] = b[row] - factor * b[col];
        }
    }
    double x[2];
    for (int i = 1; i >= 0; --i) {
        double sum = b[i];
        for (int j = i + 1; j < 2; ++j) sum = sum - A[i][j] * x[j];
        x[i] = sum / A[i][i];
    }
    dy[0] = y[2]; dy[1] = y[3]; dy[2] = x[0]; dy[3] = x[1];
}

typedef struct {
    size_t n; const double *mc, *m, *L, *g, *s0, *force; double t0, t1, dt; double* final_out;
} orc_cp_job_t;

static void orc_cp_range(void* user, size_t b, size_t e) {
    const orc_cp_job_t* j = (const orc_cp_job_t*)user;
    const size_t n = j->n;
    for (size_t i = b; i < e; ++i) {
        orc_cp_t p = {j->mc[i], j->m[i], j->L[i], j->g[i], j->force[i], 0};
        double y[4] = {j->s0[i], j->s0[n + i], j->s0[2 * n + i], j->s0[3 * n + i]}, work[20];
        orc_rk4(orc_cartpole_rhs_fn, &p, 4, j->t0, j->t1, j->dt, y, work, NULL, NULL, NULL);
        for (int q = 0; q < 4; ++q) j->final_out[q * n + i] = y[q];
    }
}

ORC_API int orc_cartpole_rk4(size_t n, const double* mc, const double* m, const double* L,
                             const double* g, const double* s0 /*[4][n]*/, const double* force,
                             double t0, double t1, double dt, double* final_out, int nthreads) {
    orc_cp_job_t j = {n, mc, m, L, g, s0, force, t0, t1, dt, final_out};
    orc_parallel_for(n, 16, nthreads, orc_cp_range, &j);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Section 8f-1 -- LQR<4,1>::solve (physics/control/lqr.hpp:260-396) and the closed loop
 * StateFeedbackController<4,1>::compute (state_feedback_controller.hpp:88-104) -> setControlForce ->
 * one RK4Solver step with the force held (zero-order hold).
 * ------------------------------------------------------------------------------------------ */
#define LN 4
static double orc_frob(const double A[LN][LN]) {                     /* lqr.hpp:140-148 */
    double sum = 0.0;
    for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) sum += A[i][j] * A[i][j];
    return sqrt(sum);
}
static void orc_mm(const double A[LN][LN], const double B[LN][LN], double C[LN][LN]) {   /* multiply, :74-85 */
    for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) {
        C[i][j] = 0.0;
        for (int k = 0; k < LN; ++k) C[i][j] += A[i][k] * B[k][j];
    }
}
/* returns 0 solved, 8 not converged (solve() == false), 2 singular (solveForGain throws) */
static int orc_lqr4_one(const double A[LN][LN], const double B[LN], const double Q[LN][LN], double R, double dt,
                        size_t max_iter, double tol, double K[LN], double P[LN][LN], int* iters) {
    double Adt[LN][LN], Adt2[LN][LN], Ad[LN][LN], AdT[LN][LN], Bcoef[LN][LN], Bd[LN], Qd[LN][LN];
    for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) Adt[i][j] = A[i][j] * dt;
    orc_mm(Adt, Adt, Adt2);
    for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) {
        const double id = i == j ? 1.0 : 0.0;
        Ad[i][j] = (id + Adt[i][j]) + Adt2[i][j] * 0.5;               /* :274-275 */
        Bcoef[i][j] = id * dt + Adt[i][j] * (dt * 0.5);               /* :278 */
        Qd[i][j] = Q[i][j] * dt;                                      /* :282 */
        P[i][j] = Q[i][j] * 1.0;                                      /* :285 */
    }
    for (int i = 0; i < LN; ++i) { Bd[i] = 0.0; for (int k = 0; k < LN; ++k) Bd[i] += Bcoef[i][k] * B[k]; }
    for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) AdT[i][j] = Ad[j][i];
    *iters = 0;
    int converged = 0;
    for (size_t it = 0; it < max_iter; ++it) {
        double AdTP[LN][LN], BdTP[LN], g[LN], Kd[LN], AdTPAd[LN][LN], Pn[LN][LN], diff[LN][LN];
        orc_mm(AdT, P, AdTP);
        for (int j = 0; j < LN; ++j) { BdTP[j] = 0.0; for (int k = 0; k < LN; ++k) BdTP[j] += Bd[k] * P[k][j]; }
        double bpb = 0.0;
        for (int k = 0; k < LN; ++k) bpb += BdTP[k] * Bd[k];
        const double sden = R * dt + bpb;
        for (int j = 0; j < LN; ++j) { g[j] = 0.0; for (int k = 0; k < LN; ++k) g[j] += BdTP[k] * Ad[k][j]; }
        if (fabs(sden) < 1e-12) return 2;
        for (int j = 0; j < LN; ++j) Kd[j] = g[j] / sden;
        orc_mm(AdTP, Ad, AdTPAd);
        for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) {
            double ktrk = 0.0; ktrk += Kd[i] * g[j];
            Pn[i][j] = (AdTPAd[i][j] - ktrk) + Qd[i][j];
        }
        for (int i = 0; i < LN; ++i) for (int j = i + 1; j < LN; ++j) {
            const double avg = 0.5 * (Pn[i][j] + Pn[j][i]);
            Pn[i][j] = avg; Pn[j][i] = avg;
        }
        for (int i = 0; i < LN; ++i) for (int j = 0; j < LN; ++j) diff[i][j] = Pn[i][j] - P[i][j];
        const double change = orc_frob(diff), p_norm = orc_frob(Pn);
        const double rel = (p_norm > 1e-10) ? change / p_norm : change;
        memcpy(P, Pn, sizeof Pn);
        *iters = (int)it + 1;
        if (rel < tol || change < tol) { converged = 1; break; }
        if (p_norm > 1e15 || isnan(change)) break;
    }
    if (!converged) {
        const double fn = orc_frob(P);
        if (!(fn < 1e10) || isnan(fn)) return 8;
    }
    if (fabs(R) < 1e-12) return 2;
    for (int j = 0; j < LN; ++j) { double s = 0.0; for (int k = 0; k < LN; ++k) s += B[k] * P[k][j]; K[j] = s / R; }
    return 0;
}

typedef struct { size_t P; const double *A, *B, *Q; double R, dt; size_t max_iter; double tol; double *K, *Pric; int *iters, *status; } orc_lqr_job_t;
static void orc_lqr_range(void* user, size_t b, size_t e) {
    const orc_lqr_job_t* j = (const orc_lqr_job_t*)user;
    const size_t P = j->P;
    for (size_t p = b; p < e; ++p) {
        double A[LN][LN], B[LN], Q[LN][LN], K[LN] = {NAN, NAN, NAN, NAN}, Pm[LN][LN];
        int it = 0;
        for (int r = 0; r < LN; ++r) { for (int c = 0; c < LN; ++c) { A[r][c] = j->A[(r * LN + c) * P + p]; Q[r][c] = j->Q[r * LN + c]; } B[r] = j->B[r * P + p]; }
        const int st = orc_lqr4_one(A, B, Q, j->R, j->dt, j->max_iter, j->tol, K, Pm, &it);
        for (int c = 0; c < LN; ++c) j->K[c * P + p] = st ? NAN : K[c];
        if (j->Pric) for (int r = 0; r < LN; ++r) for (int c = 0; c < LN; ++c) j->Pric[(r * LN + c) * P + p] = Pm[r][c];
        if (j->iters) j->iters[p] = it;
        if (j->status) j->status[p] = st;
    }
}
ORC_API int orc_lqr4(size_t P, const double* A /*[16][P]*/, const double* B /*[4][P]*/, const double* Q16, double R,
                     double dt, size_t max_iter, double tol, double* K /*[4][P]*/, double* Pric, int* iters, int* status,
                     int nthreads) {
    orc_lqr_job_t j = {P, A, B, Q16, R, dt, max_iter, tol, K, Pric, iters, status};
    orc_parallel_for(P, 4, nthreads, orc_lqr_range, &j);
    return 0;
}

/* LQR<NS,1>::solve for any NS <= 14 (same statements as orc_lqr4_one with runtime bounds) */
#define LMAX 14
static double orc_frob_n(int ns, double A[LMAX][LMAX]) {
    double sum = 0.0;
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) sum += A[i][j] * A[i][j];
    return sqrt(sum);
}
static void orc_mm_n(int ns, double A[LMAX][LMAX], double B[LMAX][LMAX], double C[LMAX][LMAX]) {
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) {
        C[i][j] = 0.0;
        for (int k = 0; k < ns; ++k) C[i][j] += A[i][k] * B[k][j];
    }
}
static int orc_lqr_one(int ns, double A[LMAX][LMAX], const double* B, double Q[LMAX][LMAX], double R, double dt, size_t max_iter,
                       double tol, double* K, double P[LMAX][LMAX], int* iters) {
    static __thread double Adt[LMAX][LMAX], Adt2[LMAX][LMAX], Ad[LMAX][LMAX], AdT[LMAX][LMAX], Bcoef[LMAX][LMAX], Qd[LMAX][LMAX],
        AdTP[LMAX][LMAX], AdTPAd[LMAX][LMAX], Pn[LMAX][LMAX], diff[LMAX][LMAX];
    double Bd[LMAX], BdTP[LMAX], g[LMAX], Kd[LMAX];
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) Adt[i][j] = A[i][j] * dt;
    orc_mm_n(ns, Adt, Adt, Adt2);
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) {
        const double id = i == j ? 1.0 : 0.0;
        Ad[i][j] = (id + Adt[i][j]) + Adt2[i][j] * 0.5;
        Bcoef[i][j] = id * dt + Adt[i][j] * (dt * 0.5);
        Qd[i][j] = Q[i][j] * dt;
        P[i][j] = Q[i][j] * 1.0;
    }
    for (int i = 0; i < ns; ++i) { Bd[i] = 0.0; for (int k = 0; k < ns; ++k) Bd[i] += Bcoef[i][k] * B[k]; }
    for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) AdT[i][j] = Ad[j][i];
    *iters = 0;
    int converged = 0;
    for (size_t it = 0; it < max_iter; ++it) {
        orc_mm_n(ns, AdT, P, AdTP);
        for (int j = 0; j < ns; ++j) { BdTP[j] = 0.0; for (int k = 0; k < ns; ++k) BdTP[j] += Bd[k] * P[k][j]; }
        double bpb = 0.0;
        for (int k = 0; k < ns; ++k) bpb += BdTP[k] * Bd[k];
        const double sden = R * dt + bpb;
        for (int j = 0; j < ns; ++j) { g[j] = 0.0; for (int k = 0; k < ns; ++k) g[j] += BdTP[k] * Ad[k][j]; }
        if (fabs(sden) < 1e-12) return 2;
        for (int j = 0; j < ns; ++j) Kd[j] = g[j] / sden;
        orc_mm_n(ns, AdTP, Ad, AdTPAd);
        for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) {
            double ktrk = 0.0; ktrk += Kd[i] * g[j];
            Pn[i][j] = (AdTPAd[i][j] - ktrk) + Qd[i][j];
        }
        for (int i = 0; i < ns; ++i) for (int j = i + 1; j < ns; ++j) {
            const double avg = 0.5 * (Pn[i][j] + Pn[j][i]);
            Pn[i][j] = avg; Pn[j][i] = avg;
        }
        for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) diff[i][j] = Pn[i][j] - P[i][j];
        const double change = orc_frob_n(ns, diff), p_norm = orc_frob_n(ns, Pn);
        const double rel = (p_norm > 1e-10) ? change / p_norm : change;
        for (int i = 0; i < ns; ++i) for (int j = 0; j < ns; ++j) P[i][j] = Pn[i][j];
        *iters = (int)it + 1;
        if (rel < tol || change < tol) { converged = 1; break; }
        if (p_norm > 1e15 || isnan(change)) break;
    }
    if (!converged) {
        const double fn = orc_frob_n(ns, P);
        if (!(fn < 1e10) || isnan(fn)) return 8;
    }
    if (fabs(R) < 1e-12) return 2;
    for (int j = 0; j < ns; ++j) { double s = 0.0; for (int k = 0; k < ns; ++k) s += B[k] * P[k][j]; K[j] = s / R; }
    return 0;
}
ORC_API int orc_lqr_n(int ns, size_t P, const double* A, const double* B, const double* Q, double R, double dt, size_t max_iter,
                      double tol, double* K, double* Pric, int* iters, int* status, int nthreads) {
    (void)nthreads;
    if (ns < 1 || ns > LMAX) return -1;
    for (size_t p = 0; p < P; ++p) {
        static __thread double Am[LMAX][LMAX], Qm[LMAX][LMAX], Pm[LMAX][LMAX];
        double Bv[LMAX], Kv[LMAX];
        int it = 0;
        for (int r = 0; r < ns; ++r) { for (int c = 0; c < ns; ++c) { Am[r][c] = A[(r * ns + c) * P + p]; Qm[r][c] = Q[r * ns + c]; } Bv[r] = B[r * P + p]; }
        const int st = orc_lqr_one(ns, Am, Bv, Qm, R, dt, max_iter, tol, Kv, Pm, &it);
        for (int c = 0; c < ns; ++c) K[c * P + p] = st ? NAN : Kv[c];
        if (Pric) for (int r = 0; r < ns; ++r) for (int c = 0; c < ns; ++c) Pric[(r * ns + c) * P + p] = Pm[r][c];
        if (iters) iters[p] = it;
        if (status) status[p] = st;
    }
    return 0;
}

typedef struct {
    size_t n; const double *mc, *m, *L, *g, *s0, *K, *xref; double umin, umax; int saturate;
    double t0, dt; size_t n_steps; double *final_out, *stats;
} orc_fb_job_t;
static void orc_fb_range(void* user, size_t b, size_t e) {
    const orc_fb_job_t* j = (const orc_fb_job_t*)user;
    const size_t n = j->n;
    for (size_t i = b; i < e; ++i) {
        orc_cp_t p = {j->mc[i], j->m[i], j->L[i], j->g[i], 0.0, 0};
        double y[4] = {j->s0[i], j->s0[n + i], j->s0[2 * n + i], j->s0[3 * n + i]}, work[20];
        double t = j->t0, max_th = fabs(y[1]), max_u = 0.0;
        for (size_t step = 0; step < j->n_steps; ++step) {
            double u = 0.0;                                            /* state_feedback_controller.hpp:91-95 */
            for (int q = 0; q < 4; ++q) u -= j->K[q * n + i] * (y[q] - (j->xref ? j->xref[q * n + i] : 0.0));
            if (j->saturate) u = u < j->umin ? j->umin : (j->umax < u ? j->umax : u);   /* std::clamp, :98-100 */
            p.F = u;
            if (fabs(u) > max_u) max_u = fabs(u);
            orc_rk4(orc_cartpole_rhs_fn, &p, 4, t, t + j->dt, j->dt, y, work, NULL, NULL, NULL);   /* exactly one step */
            t += j->dt;
            if (fabs(y[1]) > max_th) max_th = fabs(y[1]);
        }
        for (int q = 0; q < 4; ++q) j->final_out[q * n + i] = y[q];
        if (j->stats) { j->stats[i] = max_th; j->stats[n + i] = max_u; }
    }
}
ORC_API int orc_cartpole_feedback_rk4(size_t n, const double* mc, const double* m, const double* L, const double* g,
                                      const double* s0, const double* K, const double* xref, double umin, double umax,
                                      int saturate, double t0, double dt, size_t n_steps, double* final_out,
                                      double* stats, int nthreads) {
    orc_fb_job_t j = {n, mc, m, L, g, s0, K, xref, umin, umax, saturate, t0, dt, n_steps, final_out, stats};
    orc_parallel_for(n, 16, nthreads, orc_fb_range, &j);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CartNPendulum<N,double> for any N <= 8 (physics/control/cart_n_pendulum.hpp:143-221, :319-359), constant force.
 * ------------------------------------------------------------------------------------------ */
#define ORC_MAXL 8
typedef struct { int N; double mc, m[ORC_MAXL], L[ORC_MAXL], g, F; int status; } orc_cpn_t;
static void orc_cartn_rhs_fn(const void* ctx, double t, const double* y, double* dy) {
    orc_cpn_t* p = (orc_cpn_t*)ctx;
    (void)t;
    const int N = p->N, ND = N + 1;
    double s[ORC_MAXL], c[ORC_MAXL], om[ORC_MAXL], Mt[ORC_MAXL], A[ORC_MAXL + 1][ORC_MAXL + 1], b[ORC_MAXL + 1], x[ORC_MAXL + 1];
    for (int i = 0; i < N; ++i) { s[i] = sin(y[1 + i]); c[i] = cos(y[1 + i]); om[i] = y[ND + 1 + i]; }
    for (int i = 0; i < N; ++i) { double sum = 0.0; for (int j = i; j < N; ++j) sum += p->m[j]; Mt[i] = sum; }
    double total = p->mc;
    for (int i = 0; i < N; ++i) total = total + p->m[i];
    A[0][0] = total;
    b[0] = p->F;
    for (int a = 0; a < N; ++a) {
        const double coupling = Mt[a] * p->L[a] * c[a];
        A[0][1 + a] = coupling; A[1 + a][0] = coupling;
        b[0] = b[0] + Mt[a] * p->L[a] * s[a] * om[a] * om[a];
    }
    for (int a = 0; a < N; ++a) {
        double rhs_a = Mt[a] * p->g * p->L[a] * s[a];
        for (int bb = 0; bb < N; ++bb) {
            const int mx = a > bb ? a : bb;
            const double cab = c[a] * c[bb] + s[a] * s[bb];
            A[1 + a][1 + bb] = Mt[mx] * p->L[a] * p->L[bb] * cab;
            if (bb != a) { const double sab = s[a] * c[bb] - c[a] * s[bb]; rhs_a = rhs_a - Mt[mx] * p->L[a] * p->L[bb] * sab * om[bb] * om[bb]; }
        }
        b[1 + a] = rhs_a;
    }
    for (int col = 0; col < ND; ++col) {
        int pivot = col; double best = fabs(A[col][col]);
        for (int row = col + 1; row < ND; ++row) { const double mag = fabs(A[row][col]); if (mag > best) { best = mag; pivot = row; } }
        if (best < 1e-15) { p->status = 1; for (int i = 0; i < 2 * ND; ++i) dy[i] = NAN; return; }
        if (pivot != col) {
            for (int j = 0; j < ND; ++j) { const double tt = A[col][j]; A[col][j] = A[pivot][j]; A[pivot][j] = tt; }
            const double tt = b[col]; b[col] = b[pivot]; b[pivot] = tt;
        }
        for (int row = col + 1; row < ND; ++row) {
            const double factor = A[row][col] / A[col][col];
            for (int j = col; j < ND; ++j) A[row][j] = A[row][j] - factor * A[col][j];
            b[row] = b[row] - factor * b[col];
        }
    }
    for (int i = ND - 1; i >= 0; --i) {
        double sum = b[i];
        for (int j = i + 1; j < ND; ++j) sum = sum - A[i][j] * x[j];
        x[i] = sum / A[i][i];
    }
    for (int i = 0; i < ND; ++i) { dy[i] = y[ND + i]; dy[ND + i] = x[i]; }
}
ORC_API int orc_cartn_rk4(int n_links, size_t n, const double* mc, const double* m, const double* L, const double* g,
                          const double* s0, const double* force, double t0, double t1, double dt, double* final_out, int nthreads) {
    (void)nthreads;
    if (n_links < 1 || n_links > ORC_MAXL) return -1;
    const int dim = 2 * (n_links + 1);
    for (size_t i = 0; i < n; ++i) {
        orc_cpn_t p; p.N = n_links; p.mc = mc[i]; p.g = g[i]; p.F = force[i]; p.status = 0;
        for (int q = 0; q < n_links; ++q) { p.m[q] = m[q * n + i]; p.L[q] = L[q * n + i]; }
        double y[2 * (ORC_MAXL + 1)], work[10 * (ORC_MAXL + 1)];
        for (int q = 0; q < dim; ++q) y[q] = s0[q * n + i];
        orc_rk4(orc_cartn_rhs_fn, &p, (size_t)dim, t0, t1, dt, y, work, NULL, NULL, NULL);
        for (int q = 0; q < dim; ++q) final_out[q * n + i] = y[q];
    }
    return 0;
}

/* closed loop for the N-link plant: controller sample -> force held over one RK4 step (as orc_cartpole_feedback_rk4) */
ORC_API int orc_cartn_feedback_rk4(int n_links, size_t n, const double* mc, const double* m, const double* L, const double* g,
                                   const double* s0, const double* K /*[2(N+1)][n]*/, const double* xref, double umin, double umax,
                                   int saturate, double t0, double dt, size_t n_steps, double* final_out, double* stats, int nthreads) {
    (void)nthreads;
    if (n_links < 1 || n_links > ORC_MAXL) return -1;
    const int dim = 2 * (n_links + 1);
    for (size_t i = 0; i < n; ++i) {
        orc_cpn_t p; p.N = n_links; p.mc = mc[i]; p.g = g[i]; p.F = 0.0; p.status = 0;
        for (int q = 0; q < n_links; ++q) { p.m[q] = m[q * n + i]; p.L[q] = L[q * n + i]; }
        double y[2 * (ORC_MAXL + 1)], work[10 * (ORC_MAXL + 1)], t = t0, max_th = 0.0, max_u = 0.0;
        for (int q = 0; q < dim; ++q) y[q] = s0[q * n + i];
        for (int q = 1; q <= n_links; ++q) if (fabs(y[q]) > max_th) max_th = fabs(y[q]);
        for (size_t step = 0; step < n_steps; ++step) {
            double u = 0.0;
            for (int q = 0; q < dim; ++q) u -= K[q * n + i] * (y[q] - (xref ? xref[q * n + i] : 0.0));
            if (saturate) u = u < umin ? umin : (umax < u ? umax : u);
            p.F = u;
            if (fabs(u) > max_u) max_u = fabs(u);
            orc_rk4(orc_cartn_rhs_fn, &p, (size_t)dim, t, t + dt, dt, y, work, NULL, NULL, NULL);
            t += dt;
            for (int q = 1; q <= n_links; ++q) if (fabs(y[q]) > max_th) max_th = fabs(y[q]);
        }
        for (int q = 0; q < dim; ++q) final_out[q * n + i] = y[q];
        if (stats) { stats[i] = max_th; stats[n + i] = max_u; }
    }
    return 0;
}


/* ------------------------------------------------------------------------------------------
 * Section 8f-2 -- unified 6-state grid: PointMass2D + Spring2D with attachment points and torque
 * (physics/unified/components.hpp:52-138, :206-291) driven as ValidatedSystem::computeDerivatives does
 * (validated_system.hpp:204-301) on makeGridTopology (constexpr_topology.hpp:368-414), runtime sized.
 * prm[18]: mass, radius, spacing, stiffness, rest_length, damping, min_distance, repulsion_stiffness, h_attach0/1, v_attach0/1,
 *          topology (0 quad, 1 triangle = makeTriangleGridTopology, constexpr_topology.hpp:436-505: after the horizontal and the
 *          vertical springs, per cell in row-major order the main diagonal (r,c)->(r+1,c+1) then the anti-diagonal
 *          (r,c+1)->(r+1,c)), diagonal rest length, main-diagonal attachment angles 0/1, anti-diagonal attachment angles 0/1.
 * ------------------------------------------------------------------------------------------ */
typedef struct { size_t rows, cols; double mass, radius, k, L0, c, mind, krep, ha0, ha1, va0, va1; int topo; double Ld, da0, da1, aa0, aa1; } orc_ugrid_t;
static void orc_uspring(const orc_ugrid_t* g, const double* m0, const double* m1, double a0, double a1, double* F /*[rows*cols][3]*/,
                        size_t i0, size_t i1, double L0) {
    const double r0 = g->radius, r1 = g->radius;
    const double wa0 = m0[4] + a0, wa1 = m1[4] + a1;
    const double o0x = r0 * cos(wa0), o0y = r0 * sin(wa0), o1x = r1 * cos(wa1), o1y = r1 * sin(wa1);
    const double a0x = m0[0] + o0x, a0y = m0[1] + o0y, a1x = m1[0] + o1x, a1y = m1[1] + o1y;
    const double v0x = m0[2] - m0[5] * o0y, v0y = m0[3] + m0[5] * o0x, v1x = m1[2] - m1[5] * o1y, v1y = m1[3] + m1[5] * o1x;
    const double dx = a1x - a0x, dy = a1y - a0y;
    const double len = sqrt(dx * dx + dy * dy), len_safe = len < 1e-10 ? 1e-10 : len;
    const double ux = dx / len_safe, uy = dy / len_safe, ext = len - L0;
    const double dvx = v1x - v0x, dvy = v1y - v0y, rel = dvx * ux + dvy * uy;
    double Fm = g->k * ext + g->c * rel;
    if (g->mind > 0.0 && len < g->mind) { const double rep = -g->krep * (g->mind / len_safe - 1.0); Fm += rep; }
    const double f0x = Fm * ux, f0y = Fm * uy, f1x = -Fm * ux, f1y = -Fm * uy;
    const double t0 = o0x * f0y - o0y * f0x, t1 = o1x * f1y - o1y * f1x;
    F[3 * i0] += f0x; F[3 * i0 + 1] += f0y; F[3 * i1] += f1x; F[3 * i1 + 1] += f1y;      /* addForce on both ends */
    F[3 * i0 + 2] += t0; F[3 * i1 + 2] += t1;                                          /* addTorque */
}
static void orc_ugrid_rhs_fn(const void* ctx, double t, const double* s, double* d) {
    const orc_ugrid_t* g = (const orc_ugrid_t*)ctx;
    (void)t;
    const size_t n = g->rows * g->cols;
    double* F = (double*)calloc(3 * n, sizeof(double));
    for (size_t r = 0; r < g->rows; ++r)                                               /* horizontal springs, row-major */
        for (size_t c = 0; c + 1 < g->cols; ++c) { const size_t i = r * g->cols + c; orc_uspring(g, s + 6 * i, s + 6 * (i + 1), g->ha0, g->ha1, F, i, i + 1, g->L0); }
    for (size_t r = 0; r + 1 < g->rows; ++r)                                           /* then vertical springs */
        for (size_t c = 0; c < g->cols; ++c) { const size_t i = r * g->cols + c; orc_uspring(g, s + 6 * i, s + 6 * (i + g->cols), g->va0, g->va1, F, i, i + g->cols, g->L0); }
    if (g->topo == 1)                                                                  /* then the X of every cell */
        for (size_t r = 0; r + 1 < g->rows; ++r)
            for (size_t c = 0; c + 1 < g->cols; ++c) {
                const size_t tl = r * g->cols + c, tr = tl + 1, bl = tl + g->cols, br = bl + 1;
                orc_uspring(g, s + 6 * tl, s + 6 * br, g->da0, g->da1, F, tl, br, g->Ld);
                orc_uspring(g, s + 6 * tr, s + 6 * bl, g->aa0, g->aa1, F, tr, bl, g->Ld);
            }
    const double I = 0.5 * g->mass * g->radius * g->radius;                            /* components.hpp:98 */
    for (size_t i = 0; i < n; ++i) {
        d[6 * i] = s[6 * i + 2]; d[6 * i + 1] = s[6 * i + 3];
        d[6 * i + 2] = F[3 * i] / g->mass; d[6 * i + 3] = F[3 * i + 1] / g->mass;
        d[6 * i + 4] = s[6 * i + 5]; d[6 * i + 5] = F[3 * i + 2] / I;
    }
    free(F);
}
ORC_API int orc_ugrid(size_t rows, size_t cols, const double* prm, double dt, size_t n_steps, const double* state_in, double* out, int mode) {
    orc_ugrid_t g = {rows, cols, prm[0], prm[1], prm[3], prm[4], prm[5], prm[6], prm[7] > 0.0 ? prm[7] : 10.0 * prm[3], prm[8], prm[9], prm[10], prm[11],
                     (int)prm[12], prm[13], prm[14], prm[15], prm[16], prm[17]};
    const size_t n = rows * cols * 6;
    if (mode == 2) {
        for (size_t r = 0; r < rows; ++r) for (size_t c = 0; c < cols; ++c) {
            double* q = out + 6 * (r * cols + c);
            q[0] = (double)c * prm[2]; q[1] = (double)r * prm[2]; q[2] = q[3] = q[4] = q[5] = 0.0;
        }
        return 0;
    }
    if (mode == 0) { orc_ugrid_rhs_fn(&g, 0.0, state_in, out); return 0; }
    double* y = (double*)malloc(n * sizeof(double));
    double* work = (double*)malloc(5 * n * sizeof(double));
    memcpy(y, state_in, n * sizeof(double));
    double t = 0.0;
    if (mode == 3 || mode == 4) {                    /* wasm/wasm_grid2d.cpp:329-358 (symplectic Euler), :375-410 (velocity Verlet) */
        const size_t nm = rows * cols;
        double *d = work, *dn = work + n;
        for (size_t s = 0; s < n_steps; ++s) {
            orc_ugrid_rhs_fn(&g, t, y, d);
            if (mode == 3) {
                for (size_t i = 0; i < nm; ++i) { y[i * 6 + 2] += dt * d[i * 6 + 2]; y[i * 6 + 3] += dt * d[i * 6 + 3]; y[i * 6 + 5] += dt * d[i * 6 + 5]; }
                for (size_t i = 0; i < nm; ++i) { y[i * 6 + 0] += dt * y[i * 6 + 2]; y[i * 6 + 1] += dt * y[i * 6 + 3]; y[i * 6 + 4] += dt * y[i * 6 + 5]; }
            } else {
                const double half_dt_sq = 0.5 * dt * dt, half_dt = 0.5 * dt;
                for (size_t i = 0; i < nm; ++i) {
                    y[i * 6 + 0] += dt * y[i * 6 + 2] + half_dt_sq * d[i * 6 + 2];
                    y[i * 6 + 1] += dt * y[i * 6 + 3] + half_dt_sq * d[i * 6 + 3];
                    y[i * 6 + 4] += dt * y[i * 6 + 5] + half_dt_sq * d[i * 6 + 5];
                }
                orc_ugrid_rhs_fn(&g, t + dt, y, dn);
                for (size_t i = 0; i < nm; ++i) {
                    y[i * 6 + 2] += half_dt * (d[i * 6 + 2] + dn[i * 6 + 2]);
                    y[i * 6 + 3] += half_dt * (d[i * 6 + 3] + dn[i * 6 + 3]);
                    y[i * 6 + 5] += half_dt * (d[i * 6 + 5] + dn[i * 6 + 5]);
                }
            }
            t += dt;
        }
        memcpy(out, y, n * sizeof(double));
        free(y); free(work);
        return 0;
    }
    for (size_t s = 0; s < n_steps; ++s) { orc_rk4(orc_ugrid_rhs_fn, &g, n, t, t + dt, dt, y, work, NULL, NULL, NULL); t += dt; }
    memcpy(out, y, n * sizeof(double));
    free(y); free(work);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Plugin-path check -- the reference's composition example, physics/coupled_oscillator/: two PointMass
 * (point_mass.hpp:70-84: {v, F/m}), one Spring (spring.hpp:96-131: F1 = -k (x1 - x2 - L0) [+ -c (v1 - v2) when
 * c > 0], F2 = -F1), state [x1, v1, x2, v2].
 * ------------------------------------------------------------------------------------------ */
typedef struct { double m1, m2, k, L0, c; } orc_coupled_t;
static void orc_coupled_rhs(const void* ctx, double t, const double* y, double* dy) {
    const orc_coupled_t* p = (const orc_coupled_t*)ctx;
    (void)t;
    double f1 = -p->k * (y[0] - y[2] - p->L0);
    if (p->c > 0.0) f1 += -p->c * (y[1] - y[3]);
    dy[0] = y[1]; dy[1] = f1 / p->m1;
    dy[2] = y[3]; dy[3] = (-f1) / p->m2;
}
ORC_API int orc_coupled_rk4(size_t n, const double* m1, const double* m2, const double* k, const double* L0, const double* c,
                            const double* x1, const double* x2, double t0, double t1, double dt, double* final_out,
                            double* energy_out, int nthreads) {
    (void)nthreads;
    for (size_t i = 0; i < n; ++i) {
        orc_coupled_t p = {m1[i], m2[i], k[i], L0[i], c[i]};
        double y[4] = {x1[i], 0.0, x2[i], 0.0}, work[20];
        orc_rk4(orc_coupled_rhs, &p, 4, t0, t1, dt, y, work, NULL, NULL, NULL);
        for (int q = 0; q < 4; ++q) final_out[q * n + i] = y[q];
        if (energy_out) {                                  /* energy_monitor.hpp: KE1 + KE2 + 0.5 k ext^2 */
            const double ext = y[0] - y[2] - p.L0;
            energy_out[i] = 0.5 * p.m1 * y[1] * y[1] + 0.5 * p.m2 * y[3] * y[3] + (0.5 * p.k) * ext * ext;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Config 4 -- 6-DOF rocket.
 * ------------------------------------------------------------------------------------------ */
/* LinearInterpolator::interpolate, io/interpolation.hpp:49-71 */
typedef struct { size_t n; const double* x; const double* y; } orc_lerp_t;

static double orc_lerp(const orc_lerp_t* tb, double x) {
    if (tb->n == 0) return 0.0;
    if (x <= tb->x[0]) return tb->y[0];
    if (x >= tb->x[tb->n - 1]) return tb->y[tb->n - 1];
    size_t lo = 0, hi = tb->n;                 /* std::lower_bound: first element >= x */
    while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (tb->x[mid] < x) lo = mid + 1; else hi = mid; }
    size_t i = lo; if (i > 0) --i;
    double x0 = tb->x[i], x1 = tb->x[i + 1], y0 = tb->y[i], y1 = tb->y[i + 1];
    double t = (x - x0) / (x1 - x0);
    return y0 + t * (y1 - y0);
}

/* StandardAtmosphere, rocket/standard_atmosphere.hpp:18-71 */
static const double ATM_R = 8.3144598, ATM_M = 0.0289644, ATM_GAMMA = 1.4, ATM_G0 = 9.80665,
                    ATM_P0 = 101325.0, ATM_T0 = 288.15;
typedef struct { double h_top, P_base, T_base, lapse, h_base; } orc_layer_t;
static const orc_layer_t ATM_LAYERS[7] = {
    {11000.0, 101325.0, 288.15, -0.0065, 0.0},   {20000.0, 22632.0, 216.65, 0.0, 11000.0},
    {32000.0, 5474.89, 216.65, 0.001, 20000.0},  {47000.0, 868.02, 228.65, 0.0028, 32000.0},
    {51000.0, 110.91, 270.65, 0.0, 47000.0},     {71000.0, 66.94, 270.65, -0.0028, 51000.0},
    {100000.0, 3.96, 214.65, -0.002, 71000.0}};

static double orc_atm_temperature(double h) {      /* :43-52 */
    if (h <= 0) return ATM_T0;
    if (h >= 100000.0) return 186.95;
    for (int i = 0; i < 7; ++i) {
        const orc_layer_t* l = &ATM_LAYERS[i];
        if (h <= l->h_top)
            return fabs(l->lapse) < 1e-10 ? l->T_base : l->T_base + l->lapse * (h - l->h_base);
    }
    return 186.95;
}
static double orc_atm_pressure(double h) {         /* :54-68 */
    if (h <= 0) return ATM_P0;
    if (h >= 100000.0) return 1e-4;
    for (int i = 0; i < 7; ++i) {
        const orc_layer_t* l = &ATM_LAYERS[i];
        if (h <= l->h_top) {
            if (fabs(l->lapse) < 1e-10)
                return l->P_base * exp(-(ATM_G0 * ATM_M) * (h - l->h_base) / (ATM_R * l->T_base));
            double temp_ratio = l->T_base / (l->T_base + l->lapse * (h - l->h_base));
            return l->P_base * pow(temp_ratio, ATM_G0 * ATM_M / (ATM_R * l->lapse));
        }
    }
    return 1e-4;
}
static double orc_atm_density(double h) {          /* :70 */
    return orc_atm_pressure(h) / ((ATM_R / ATM_M) * orc_atm_temperature(h));
}
static double orc_atm_sos(double h) {              /* :71 */
    return sqrt((ATM_GAMMA * (ATM_R / ATM_M)) * orc_atm_temperature(h));
}

ORC_API int orc_atmosphere(size_t n, const double* h, double* T, double* P, double* rho, double* a) {
    for (size_t i = 0; i < n; ++i) {
        T[i] = orc_atm_temperature(h[i]); P[i] = orc_atm_pressure(h[i]);
        rho[i] = orc_atm_density(h[i]); a[i] = orc_atm_sos(h[i]);
    }
    return 0;
}

/* InterpolatedEngine load-time precompute, rocket/interpolated_engine.hpp:176-252 */
static double orc_pressure_ratio(double k, double area_ratio) {   /* :225-252 */
    double pr = 0.01;
    for (int iter = 0; iter < 20; ++iter) {
        double M2 = (2.0 / (k - 1.0)) * (pow(1.0 / pr, (k - 1.0) / k) - 1.0);
        if (M2 < 0) M2 = 0;
        double M = sqrt(M2);
        double term = (2.0 / (k + 1.0)) * (1.0 + (k - 1.0) / 2.0 * M2);
        double exponent = (k + 1.0) / (2.0 * (k - 1.0));
        double area_calc = (1.0 / M) * pow(term, exponent);
        if (fabs(area_calc - area_ratio) < 0.001 * area_ratio) break;
        if (area_calc > area_ratio) pr *= 0.95; else pr *= 1.05;
        if (pr < 1e-6) pr = 1e-6;
        if (pr > 1.0) pr = 1.0;
    }
    return pr;
}

#define ORC_PI 3.14159265358979323846     /* interpolated_engine.hpp:34 */

/* engine: [rows][8] row-major = time, throat_d, exit_d, p_comb, T_comb, gamma, mol_mass, eff
 * derived: [3][rows] = pressure_ratio, mass_flow, engine_force */
ORC_API int orc_engine_precompute(size_t rows, const double* engine, double* derived) {
    if (rows == 0) return 0;
    double* col = (double*)malloc(sizeof(double) * rows * 8);
    for (size_t r = 0; r < rows; ++r)
        for (int c = 0; c < 8; ++c) col[c * rows + r] = engine[r * 8 + c];
    const double* times = col;
    orc_lerp_t T[8];
    for (int c = 0; c < 8; ++c) { T[c].n = rows; T[c].x = times; T[c].y = col + c * rows; }
    for (size_t i = 0; i < rows; ++i) {
        double t = times[i];
        double k = orc_lerp(&T[5], t), d_throat = orc_lerp(&T[1], t), d_exit = orc_lerp(&T[2], t);
        double p = orc_lerp(&T[3], t), temp = orc_lerp(&T[4], t), mol_mass = orc_lerp(&T[6], t);
        double eff = orc_lerp(&T[7], t);
        double throat_area = ORC_PI * d_throat * d_throat / 4.0;
        double exit_area = ORC_PI * d_exit * d_exit / 4.0;
        double area_ratio = exit_area / throat_area;
        double ratio = orc_pressure_ratio(k, area_ratio);
        double throat_pressure = p * pow(2.0 / (k + 1.0), k / (k - 1.0));
        double throat_temp = temp * (2.0 / (k + 1.0));
        double R_specific = ATM_R / mol_mass;
        double throat_density = throat_pressure / (R_specific * throat_temp);
        double throat_velocity = sqrt(k * R_specific * throat_temp);
        double mass_flow = throat_density * throat_velocity * throat_area;
        if (mass_flow < 0) mass_flow = 0;
        double exit_velocity = eff * sqrt(2.0 * k / (k - 1.0) * R_specific * temp *
                                          (1.0 - pow(ratio, (k - 1.0) / k)));
        derived[i] = ratio;
        derived[rows + i] = mass_flow;
        derived[2 * rows + i] = mass_flow * exit_velocity;
    }
    free(col);
    return 0;
}

typedef struct {
    /* engine */
    size_t eng_rows; const double* eng_t; const double* exit_d; const double* p_comb;
    const double* press_ratio; const double* eng_force; double burn_time;
    /* body */
    orc_lerp_t mass, ix, iyz; int use_interp;
    double const_mass, const_ix, const_iyz;
    /* aerodynamic coefficient tables (axisymmetric_aero.hpp:38-44), rows == 0: not loaded */
    struct orc_tab2d { size_t rows, cols; const double *rv, *cv, *data; } cd_on, cd_off, cl_on, cl_off, cs_on, cs_off, cmq_yz;
    int engine_on;
    /* per trajectory */
    double g, ref_area, ref_len;
} orc_rocket_t;

/* BilinearInterpolator::interpolate, io/interpolation.hpp:137-181 */
static double orc_bilinear(const struct orc_tab2d* t, double row, double col) {
    if (t->rows == 0) return 0.0;
    if (row <= t->rv[0]) row = t->rv[0]; else if (row >= t->rv[t->rows - 1]) row = t->rv[t->rows - 1];
    if (col <= t->cv[0]) col = t->cv[0]; else if (col >= t->cv[t->cols - 1]) col = t->cv[t->cols - 1];
    size_t ri = 0, ci = 0;
    { size_t lo = 0, hi = t->rows; while (lo < hi) { size_t mid = (lo + hi) / 2; if (t->rv[mid] < row) lo = mid + 1; else hi = mid; } ri = lo; }
    if (ri > 0) --ri;
    if (ri >= t->rows - 1) ri = t->rows - 2;
    { size_t lo = 0, hi = t->cols; while (lo < hi) { size_t mid = (lo + hi) / 2; if (t->cv[mid] < col) lo = mid + 1; else hi = mid; } ci = lo; }
    if (ci > 0) --ci;
    if (ci >= t->cols - 1) ci = t->cols - 2;
    const double r0 = t->rv[ri], r1 = t->rv[ri + 1], c0 = t->cv[ci], c1 = t->cv[ci + 1];
    const double tr = (row - r0) / (r1 - r0), tc = (col - c0) / (c1 - c0);
    const double v00 = t->data[ri * t->cols + ci], v01 = t->data[ri * t->cols + ci + 1];
    const double v10 = t->data[(ri + 1) * t->cols + ci], v11 = t->data[(ri + 1) * t->cols + ci + 1];
    const double v0 = v00 + tc * (v01 - v00), v1 = v10 + tc * (v11 - v10);
    return v0 + tr * (v1 - v0);
}
/* getCd / getCl / getCs, axisymmetric_aero.hpp:133-167 */
static double orc_coeff(const orc_rocket_t* p, const struct orc_tab2d* on, const struct orc_tab2d* off, double angle_deg,
                        double mach, double dflt) {
    if (p->engine_on && on->rows) return orc_bilinear(on, mach, angle_deg);
    if (off->rows) return orc_bilinear(off, mach, angle_deg);
    return dflt;
}
/* aero blob: engine_on, then 7 x (rows, cols, row_vals[rows], col_vals[cols], data[rows*cols]) in the order
 * cd_on, cd_off, cl_on, cl_off, cs_on, cs_off, cmq_yz */
static void orc_rocket_set_aero(orc_rocket_t* p, const double* blob) {
    p->engine_on = 1;
    if (!blob) return;
    p->engine_on = blob[0] != 0.0;
    const double* q = blob + 1;
    struct orc_tab2d* t[7] = {&p->cd_on, &p->cd_off, &p->cl_on, &p->cl_off, &p->cs_on, &p->cs_off, &p->cmq_yz};
    for (int k = 0; k < 7; ++k) {
        t[k]->rows = (size_t)q[0]; t[k]->cols = (size_t)q[1]; q += 2;
        t[k]->rv = q; q += t[k]->rows;
        t[k]->cv = q; q += t[k]->cols;
        t[k]->data = q; q += t[k]->rows * t[k]->cols;
    }
}

typedef struct { double x, y, z; } v3;

static void orc_rotmat(const double* q, double R[3][3]) {        /* rocket/quaternion.hpp:94-112 */
    double q1 = q[0], q2 = q[1], q3 = q[2], q4 = q[3];
    double q1_2 = q1 * q1, q2_2 = q2 * q2, q3_2 = q3 * q3, q4_2 = q4 * q4;
    double q1q2 = q1 * q2, q1q3 = q1 * q3, q1q4 = q1 * q4, q2q3 = q2 * q3, q2q4 = q2 * q4, q3q4 = q3 * q4;
    R[0][0] = q4_2 + q1_2 - q2_2 - q3_2; R[0][1] = 2.0 * (q1q2 - q3q4); R[0][2] = 2.0 * (q1q3 + q2q4);
    R[1][0] = 2.0 * (q1q2 + q3q4); R[1][1] = q4_2 - q1_2 + q2_2 - q3_2; R[1][2] = 2.0 * (q2q3 - q1q4);
    R[2][0] = 2.0 * (q1q3 - q2q4); R[2][1] = 2.0 * (q2q3 + q1q4); R[2][2] = q4_2 - q1_2 - q2_2 + q3_2;
}
static v3 orc_b2r(const double* q, v3 b) {                       /* :115-122 */
    double R[3][3]; orc_rotmat(q, R);
    v3 r = {R[0][0] * b.x + R[0][1] * b.y + R[0][2] * b.z,
            R[1][0] * b.x + R[1][1] * b.y + R[1][2] * b.z,
            R[2][0] * b.x + R[2][1] * b.y + R[2][2] * b.z};
    return r;
}
static v3 orc_r2b(const double* q, v3 v) {                       /* :125-133 */
    double R[3][3]; orc_rotmat(q, R);
    v3 r = {R[0][0] * v.x + R[1][0] * v.y + R[2][0] * v.z,
            R[0][1] * v.x + R[1][1] * v.y + R[2][1] * v.z,
            R[0][2] * v.x + R[1][2] * v.y + R[2][2] * v.z};
    return r;
}
static double orc_norm(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }  /* vector3.hpp:89-98 */

static double orc_rocket_mass(const orc_rocket_t* p, double time) {  /* rocket_body.hpp:146-152 */
    if (p->use_interp && p->mass.n) return orc_lerp(&p->mass, time);
    return p->const_mass;
}

/* InterpolatedEngine::compute(ThrustForceBody,...) :139-152 + computeThrustMagnitude :103-121 */
static double orc_thrust(const orc_rocket_t* p, double time, double alt) {
    if (p->eng_rows == 0) return 0.0;
    if (!(time <= p->burn_time)) return 0.0;
    double patm = orc_atm_pressure(alt);
    orc_lerp_t t;
    t.n = p->eng_rows; t.x = p->eng_t;
    t.y = p->exit_d; double d_exit = orc_lerp(&t, time);
    t.y = p->p_comb; double p_comb = orc_lerp(&t, time);
    t.y = p->press_ratio; double press_ratio = orc_lerp(&t, time);
    double exit_area = ORC_PI * d_exit * d_exit / 4.0;
    double exit_pressure = p_comb * press_ratio;
    t.y = p->eng_force; double engine_force = orc_lerp(&t, time);
    double pressure_force = exit_area * (exit_pressure - patm);
    return engine_force + pressure_force;
}

/* AxisymmetricAerodynamics::computeAeroForceBody, rocket/axisymmetric_aero.hpp:170-246 (no tables:
 * Cd=0.3 :142, Cl=0 :154, Cs=0 :166) */
static v3 orc_aero_force_body(const orc_rocket_t* p, const double* s) {
    v3 zero = {0, 0, 0};
    v3 vel = {s[4], s[5], s[6]};
    v3 vb = orc_r2b(s + 7, vel);
    double airspeed = orc_norm(vb);
    if (airspeed < 1.0) return zero;
    double alt = s[3];
    double density = orc_atm_density(alt);
    double sos = orc_atm_sos(alt);
    double mach = airspeed / sos;
    double dynp = 0.5 * density * airspeed * airspeed;
    double aoa = (vb.x * vb.x + vb.z * vb.z) < 1e-10 ? 0.0 : atan2(-vb.z, vb.x);   /* :116-122 */
    double aoss = (vb.x * vb.x + vb.y * vb.y) < 1e-10 ? 0.0 : atan2(vb.y, vb.x);  /* :125-131 */
    const double rad2deg = 180.0 / 3.14159265358979;                                /* :175 */
    double aoa_deg = fabs(aoa * rad2deg), aoss_deg = fabs(aoss * rad2deg);          /* :203-204 */
    double cd = orc_coeff(p, &p->cd_on, &p->cd_off, aoa_deg, mach, 0.3);
    double cl = orc_coeff(p, &p->cl_on, &p->cl_off, aoa_deg, mach, 0.0);
    double cs = orc_coeff(p, &p->cs_on, &p->cs_off, aoss_deg, mach, 0.0);
    double qS = dynp * p->ref_area;
    double n = orc_norm(vb);
    v3 vu = zero;
    if (!(n < 1e-15)) { vu.x = vb.x / n; vu.y = vb.y / n; vu.z = vb.z / n; }      /* vector3.hpp:101-105 */
    double drag = qS * cd, lift = qS * cl, side = qS * cs;
    v3 F_drag = {(-vu.x) * drag, (-vu.y) * drag, (-vu.z) * drag};
    v3 lift_dir = {-sin(aoa), 0.0, cos(aoa)};
    if (aoa < 0) lift_dir.z = -lift_dir.z;
    v3 F_lift = {lift_dir.x * lift, lift_dir.y * lift, lift_dir.z * lift};
    v3 side_dir = {0.0, -sin(aoss), 0.0};
    if (aoss < 0) side_dir.y = -side_dir.y;
    v3 F_side = {side_dir.x * side, side_dir.y * side, side_dir.z * side};
    v3 r = {F_drag.x + F_lift.x + F_side.x, F_drag.y + F_lift.y + F_side.y, F_drag.z + F_lift.z + F_side.z};
    return r;
}

/* computeAeroMomentBody :257-288 (default cmq = -450) */
static v3 orc_aero_moment_body(const orc_rocket_t* p, const double* s) {
    v3 zero = {0, 0, 0};
    v3 vel = {s[4], s[5], s[6]};
    v3 vb = orc_r2b(s + 7, vel);
    double airspeed = orc_norm(vb);
    if (airspeed < 1.0) return zero;
    double density = orc_atm_density(s[3]);
    double dynp = 0.5 * density * airspeed * airspeed;
    double qSd = dynp * p->ref_area * p->ref_len;
    double damping = p->ref_len / airspeed;
    double cmq = -450.0;
    if (p->cmq_yz.rows) {                                                           /* :276-281 */
        double mach = airspeed / orc_atm_sos(s[3]);
        double aoa = (vb.x * vb.x + vb.z * vb.z) < 1e-10 ? 0.0 : atan2(-vb.z, vb.x);
        double aoa_deg = aoa * 180.0 / 3.14159265358979;
        cmq = orc_bilinear(&p->cmq_yz, aoa_deg, mach);
    }
    v3 r = {cmq * qSd * damping * s[11], cmq * qSd * damping * s[12], cmq * qSd * damping * s[13]};
    return r;
}

static void orc_rocket_rhs_fn(const void* ctx, double t_unused, const double* s, double* d) {
    const orc_rocket_t* p = (const orc_rocket_t*)ctx;
    (void)t_unused;                       /* components read state[0], simulation_time.hpp:56-58 */
    double time = s[0], alt = s[3];
    d[0] = 1.0;                                              /* simulation_time.hpp:46-54 */
    d[1] = s[4]; d[2] = s[5]; d[3] = s[6];                   /* translation_kinematics.hpp:46-55 */
    /* ForceAggregator::compute(TotalForceENU), force_aggregator.hpp:61-79 */
    double mass = orc_rocket_mass(p, time);
    v3 Fg = {0.0, 0.0, -p->g * mass};
    v3 tb = {orc_thrust(p, time, alt), 0.0, 0.0};
    v3 Ft = orc_b2r(s + 7, tb);
    v3 Fa = orc_b2r(s + 7, orc_aero_force_body(p, s));
    v3 F = {Fg.x + Ft.x + Fa.x, Fg.y + Ft.y + Fa.y, Fg.z + Ft.z + Fa.z};
    double mass2 = orc_rocket_mass(p, time);                 /* translation_dynamics.hpp:39 */
    d[4] = F.x / mass2; d[5] = F.y / mass2; d[6] = F.z / mass2;
    /* RotationKinematics, rotation_kinematics.hpp:39-46; Quaternion::derivative quaternion.hpp:80-90 */
    double q1 = s[7], q2 = s[8], q3 = s[9], q4 = s[10], wx = s[11], wy = s[12], wz = s[13];
    double p1 = wx, p2 = wy, p3 = wz, p4 = 0.0;
    d[7]  = (q4 * p1 + q1 * p4 + q2 * p3 - q3 * p2) * 0.5;
    d[8]  = (q4 * p2 - q1 * p3 + q2 * p4 + q3 * p1) * 0.5;
    d[9]  = (q4 * p3 + q1 * p2 - q2 * p1 + q3 * p4) * 0.5;
    d[10] = (q4 * p4 - q1 * p1 - q2 * p2 - q3 * p3) * 0.5;
    /* RotationDynamics, rotation_dynamics.hpp:35-47 */
    v3 torque = orc_aero_moment_body(p, s);
    double Ix, Iy, Iz;
    if (p->use_interp) {                                     /* rocket_body.hpp:163-179,  :121-136 */
        Ix = p->ix.n ? orc_lerp(&p->ix, time) : p->const_ix;
        Iy = p->iyz.n ? orc_lerp(&p->iyz, time) : p->const_iyz;
        Iz = Iy;
    } else { Ix = p->const_ix; Iy = p->const_iyz; Iz = p->const_iyz; }
    d[11] = (torque.x - (Iz - Iy) * wy * wz) / Ix;
    d[12] = (torque.y - (Ix - Iz) * wx * wz) / Iy;
    d[13] = (torque.z - (Iy - Ix) * wx * wy) / Iz;
}

/* Quaternion::from_launcher_angles, rocket/quaternion.hpp:164-190 */
ORC_API void orc_launcher_quaternion(double elevation_deg, double azimuth_deg, double* q) {
    const double deg2rad = 3.14159265358979323846 / 180.0;
    double yaw = (90.0 - azimuth_deg) * deg2rad;
    double pitch = elevation_deg * deg2rad;
    double y = yaw * 0.5, p = pitch * 0.5;
    double cy = cos(y), sy = sin(y), cp = cos(p), sp = sin(p);
    q[0] = sy * sp; q[1] = -cy * sp; q[2] = sy * cp; q[3] = cy * cp;
}

/* statistics scan, rocket/simulation_result.hpp:151-189 (+ final downrange E/N) */
typedef struct {
    orc_traj_t tr; double burn;
    double apogee, apogee_t, vmax, vmax_t, bo_alt, bo_spd;
} orc_rstat_t;

static void orc_rstat_visit(orc_rstat_t* st, double t, const double* y) {
    double alt = y[3];
    double spd = sqrt(y[4] * y[4] + y[5] * y[5] + y[6] * y[6]);
    if (alt > st->apogee) { st->apogee = alt; st->apogee_t = t; }
    if (spd > st->vmax) { st->vmax = spd; st->vmax_t = t; }
    if (st->burn > 0 && fabs(t - st->burn) < 0.1) { st->bo_alt = alt; st->bo_spd = spd; }
}
static void orc_rstat_snap(void* user, size_t step, double t, const double* y) {
    orc_rstat_t* st = (orc_rstat_t*)user;
    orc_rstat_visit(st, t, y);
    orc_traj_snap(&st->tr, step, t, y);
}

static void orc_rocket_setup(orc_rocket_t* p, size_t engine_rows, const double* eng_cols,
                             const double* derived, size_t mass_rows, const double* mass_cols,
                             size_t ix_rows, const double* ix_cols, size_t iyz_rows,
                             const double* iyz_cols) {
    memset(p, 0, sizeof(*p));
    p->eng_rows = engine_rows;
    if (engine_rows) {
        p->eng_t = eng_cols; p->exit_d = eng_cols + 2 * engine_rows; p->p_comb = eng_cols + 3 * engine_rows;
        p->press_ratio = derived; p->eng_force = derived + 2 * engine_rows;
        p->burn_time = eng_cols[engine_rows - 1];
    }
    p->mass.n = mass_rows; p->mass.x = mass_cols; p->mass.y = mass_cols + mass_rows;
    p->ix.n = ix_rows; p->ix.x = ix_cols; p->ix.y = ix_cols + ix_rows;
    p->iyz.n = iyz_rows; p->iyz.x = iyz_cols; p->iyz.y = iyz_cols + iyz_rows;
    p->use_interp = (mass_rows || ix_rows || iyz_rows) ? 1 : 0;
    p->const_mass = 100.0; p->const_ix = 10.0; p->const_iyz = 100.0;   /* rocket_body.hpp:36-39 */
}

static double* orc_transpose(size_t rows, size_t cols, const double* a) {
    double* c = (double*)malloc(sizeof(double) * (rows * cols + 1));
    for (size_t r = 0; r < rows; ++r) for (size_t k = 0; k < cols; ++k) c[k * rows + r] = a[r * cols + k];
    return c;
}

typedef struct {
    size_t n; const double *elev, *az, *diam, *g0; const orc_rocket_t* base; double t_end, dt;
    size_t store_every; double *final_out, *traj_out, *stats_out;
} orc_rk_job_t;

static void orc_rocket_range(void* user, size_t b, size_t e) {
    const orc_rk_job_t* j = (const orc_rk_job_t*)user;
    const size_t n = j->n;
    for (size_t i = b; i < e; ++i) {
        orc_rocket_t p = *j->base;
        p.g = j->g0[i];
        p.ref_area = 3.14159265358979 * j->diam[i] * j->diam[i] / 4.0;  /* axisymmetric_aero.hpp:56 */
        p.ref_len = j->diam[i];
        double y[14] = {0}, work[70];
        orc_launcher_quaternion(j->elev[i], j->az[i], y + 7);
        orc_rstat_t st;
        memset(&st, 0, sizeof(st));
        st.tr.traj = j->traj_out; st.tr.dim = 14; st.tr.n = n; st.tr.i = i; st.tr.every = j->store_every;
        st.burn = j->base->burn_time;
        orc_rstat_visit(&st, 0.0, y);                       /* stored initial condition */
        orc_rk4(orc_rocket_rhs_fn, &p, 14, 0.0, j->t_end, j->dt, y, work, orc_rstat_snap, &st, NULL);
        for (int q = 0; q < 14; ++q) j->final_out[q * n + i] = y[q];
        if (j->stats_out) {
            double v[8] = {st.apogee, st.apogee_t, st.vmax, st.vmax_t, st.bo_alt, st.bo_spd, y[1], y[2]};
            for (int q = 0; q < 8; ++q) j->stats_out[q * n + i] = v[q];
        }
    }
}

ORC_API int orc_rocket_rk4_aero(size_t n, const double* elev_deg, const double* az_deg,
                           const double* diameter, const double* g0, size_t engine_rows,
                           const double* engine /*[rows][8]*/, size_t mass_rows,
                           const double* mass_tab /*[rows][2]*/, size_t ix_rows, const double* ix_tab,
                           size_t iyz_rows, const double* iyz_tab, double t_end, double dt,
                           size_t store_every, double* final_out /*[14][n]*/, double* traj_out,
                           double* stats_out /*[8][n]*/, int nthreads, const double* aero_blob) {
    double* eng_cols = orc_transpose(engine_rows, 8, engine);
    double* derived = (double*)malloc(sizeof(double) * (3 * engine_rows + 1));
    orc_engine_precompute(engine_rows, engine, derived);
    double* mass_cols = orc_transpose(mass_rows, 2, mass_tab);
    double* ix_cols = orc_transpose(ix_rows, 2, ix_tab);
    double* iyz_cols = orc_transpose(iyz_rows, 2, iyz_tab);
    orc_rocket_t base;
    orc_rocket_setup(&base, engine_rows, eng_cols, derived, mass_rows, mass_cols, ix_rows, ix_cols,
                     iyz_rows, iyz_cols);
    orc_rocket_set_aero(&base, aero_blob);
    orc_rk_job_t j = {n, elev_deg, az_deg, diameter, g0, &base, t_end, dt, store_every,
                      final_out, traj_out, stats_out};
    orc_parallel_for(n, 4, nthreads, orc_rocket_range, &j);
    free(eng_cols); free(derived); free(mass_cols); free(ix_cols); free(iyz_cols);
    return 0;
}

ORC_API int orc_rocket_rk4(size_t n, const double* elev_deg, const double* az_deg,
                           const double* diameter, const double* g0, size_t engine_rows,
                           const double* engine, size_t mass_rows, const double* mass_tab, size_t ix_rows, const double* ix_tab,
                           size_t iyz_rows, const double* iyz_tab, double t_end, double dt,
                           size_t store_every, double* final_out, double* traj_out, double* stats_out, int nthreads) {
    return orc_rocket_rk4_aero(n, elev_deg, az_deg, diameter, g0, engine_rows, engine, mass_rows, mass_tab, ix_rows, ix_tab,
                               iyz_rows, iyz_tab, t_end, dt, store_every, final_out, traj_out, stats_out, nthreads, NULL);
}

ORC_API int orc_rocket_rhs_aero(size_t n, double elev_deg, double az_deg, double diameter, double g0,
                           size_t engine_rows, const double* engine, size_t mass_rows,
                           const double* mass_tab, const double* state /*[14][n]*/, double* out, const double* aero_blob) {
    (void)elev_deg; (void)az_deg;
    double* eng_cols = orc_transpose(engine_rows, 8, engine);
    double* derived = (double*)malloc(sizeof(double) * (3 * engine_rows + 1));
    orc_engine_precompute(engine_rows, engine, derived);
    double* mass_cols = orc_transpose(mass_rows, 2, mass_tab);
    orc_rocket_t p;
    orc_rocket_setup(&p, engine_rows, eng_cols, derived, mass_rows, mass_cols, 0, NULL, 0, NULL);
    orc_rocket_set_aero(&p, aero_blob);
    p.g = g0; p.ref_area = 3.14159265358979 * diameter * diameter / 4.0; p.ref_len = diameter;
    for (size_t i = 0; i < n; ++i) {
        double s[14], d[14];
        for (int j = 0; j < 14; ++j) s[j] = state[j * n + i];
        orc_rocket_rhs_fn(&p, 0.0, s, d);
        for (int j = 0; j < 14; ++j) out[j * n + i] = d[j];
    }
    free(eng_cols); free(derived); free(mass_cols);
    return 0;
}

ORC_API int orc_rocket_rhs(size_t n, double elev_deg, double az_deg, double diameter, double g0, size_t engine_rows,
                           const double* engine, size_t mass_rows, const double* mass_tab, const double* state, double* out) {
    return orc_rocket_rhs_aero(n, elev_deg, az_deg, diameter, g0, engine_rows, engine, mass_rows, mass_tab, state, out, NULL);
}

ORC_API int orc_engine_thrust(size_t engine_rows, const double* engine, size_t nq, const double* tq,
                              const double* patm, double* thrust) {
    double* eng_cols = orc_transpose(engine_rows, 8, engine);
    double* derived = (double*)malloc(sizeof(double) * (3 * engine_rows + 1));
    orc_engine_precompute(engine_rows, engine, derived);
    orc_lerp_t t; t.n = engine_rows; t.x = eng_cols;
    double burn = engine_rows ? eng_cols[engine_rows - 1] : 0.0;
    for (size_t i = 0; i < nq; ++i) {
        if (!engine_rows || !(tq[i] <= burn)) { thrust[i] = 0.0; continue; }
        t.y = eng_cols + 2 * engine_rows; double d_exit = orc_lerp(&t, tq[i]);
        t.y = eng_cols + 3 * engine_rows; double p_comb = orc_lerp(&t, tq[i]);
        t.y = derived; double pr = orc_lerp(&t, tq[i]);
        double exit_area = ORC_PI * d_exit * d_exit / 4.0;
        double exit_pressure = p_comb * pr;
        t.y = derived + 2 * engine_rows; double ef = orc_lerp(&t, tq[i]);
        thrust[i] = ef + exit_area * (exit_pressure - patm[i]);
    }
    free(eng_cols); free(derived);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Config 5 -- grid2d mass-spring, runtime-sized restatement of the templated system
 * (physics/connected_masses/): spring indexed_spring_2d.hpp:107-165, aggregation order
 * force_aggregator_2d.hpp:58-88 over the edge list of grid_2d.hpp:92-143 (quad, quad+diag) or
 * :255-285 (triangular), mass indexed_point_mass_2d.hpp:90-108, layout/ICs
 * connectivity_matrix_2d.hpp:156-198.
 * topo: 0 quad, 1 quad + diagonals (makeGrid2DSystem<...,true>), 2 triangular.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    size_t rows, cols; int topo; double mass, spacing, k, c;
    size_t row_lo, row_hi;      /* rows whose derivative is produced: [row_lo, row_hi) */
} orc_grid_t;

/* force on mass `a` from the spring whose FIRST endpoint is i and second is j (a is one of them) */
static inline void orc_spring(const double* s, size_t i, size_t j, double k, double c, double L0,
                              double* fx, double* fy) {
    const double* pi = s + 4 * i; const double* pj = s + 4 * j;
    double dx = pj[0] - pi[0], dy = pj[1] - pi[1];
    double length = sqrt(dx * dx + dy * dy);
    double length_safe = length < 1e-10 ? 1e-10 : length;           /* :125 */
    double ux = dx / length_safe, uy = dy / length_safe;
    double extension = length - L0;
    double dvx = pj[2] - pi[2], dvy = pj[3] - pi[3];
    double rel = dvx * ux + dvy * uy;
    double F = k * extension + c * rel;                               /* :140 */
    *fx = F * ux; *fy = F * uy;
}

static void orc_grid_rhs_rows(const orc_grid_t* g, const double* s, double* d) {
    const size_t R = g->rows, C = g->cols;
    const double L_orth = g->spacing * sqrt(1.0 * 1.0 + 0.0 * 0.0);   /* connectivity_matrix_2d.hpp:190-192 */
    const double L_diag = g->spacing * sqrt(1.0 * 1.0 + 1.0 * 1.0);
    const double k = g->k, c = g->c;
    for (size_t r = g->row_lo; r < g->row_hi; ++r)
        for (size_t col = 0; col < C; ++col) {
            size_t a = r * C + col;
            double Fx = 0.0, Fy = 0.0, fx, fy;
#define AS_FIRST(j, L)  do { orc_spring(s, a, (j), k, c, (L), &fx, &fy); Fx += fx; Fy += fy; } while (0)
#define AS_SECOND(i, L) do { orc_spring(s, (i), a, k, c, (L), &fx, &fy); Fx -= fx; Fy -= fy; } while (0)
            /* horizontal edges (r,c)-(r,c+1), then vertical (r,c)-(r+1,c): global edge order */
            if (col > 0) AS_SECOND(a - 1, L_orth);
            if (col + 1 < C) AS_FIRST(a + 1, L_orth);
            if (r > 0) AS_SECOND(a - C, L_orth);
            if (r + 1 < R) AS_FIRST(a + C, L_orth);
            if (g->topo == 1) {
                /* all main diagonals (r,c)-(r+1,c+1), then all anti-diagonals (r,c)-(r+1,c-1) */
                if (r > 0 && col > 0) AS_SECOND(a - C - 1, L_diag);
                if (r + 1 < R && col + 1 < C) AS_FIRST(a + C + 1, L_diag);
                if (r > 0 && col + 1 < C) AS_SECOND(a - C + 1, L_diag);
                if (r + 1 < R && col > 0) AS_FIRST(a + C - 1, L_diag);
            } else if (g->topo == 2) {
                /* per cell (r,c): main (r,c)-(r+1,c+1) then anti (r,c+1)-(r+1,c), cells row-major */
                if (r > 0 && col > 0) AS_SECOND(a - C - 1, L_diag);           /* cell (r-1,c-1) main */
                if (r > 0 && col + 1 < C) AS_SECOND(a - C + 1, L_diag);       /* cell (r-1,c)   anti */
                if (r + 1 < R && col > 0) AS_FIRST(a + C - 1, L_diag);        /* cell (r,c-1)   anti */
                if (r + 1 < R && col + 1 < C) AS_FIRST(a + C + 1, L_diag);    /* cell (r,c)     main */
            }
#undef AS_FIRST
#undef AS_SECOND
            d[4 * a + 0] = s[4 * a + 2];
            d[4 * a + 1] = s[4 * a + 3];
            d[4 * a + 2] = Fx / g->mass;
            d[4 * a + 3] = Fy / g->mass;
        }
}

static void orc_grid_rhs_fn(const void* ctx, double t, const double* s, double* d) {
    (void)t;
    orc_grid_rhs_rows((const orc_grid_t*)ctx, s, d);
}

ORC_API int orc_grid2d_initial_state(size_t rows, size_t cols, double spacing, double* state) {
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) {
            double* p = state + 4 * (r * cols + c);
            p[0] = (double)c * spacing; p[1] = (double)r * spacing; p[2] = 0.0; p[3] = 0.0;
        }
    return 0;
}

ORC_API int orc_grid2d_rhs(size_t rows, size_t cols, int topo, double mass, double spacing, double k,
                           double c, const double* state, double* out) {
    orc_grid_t g = {rows, cols, topo, mass, spacing, k, c, 0, rows};
    orc_grid_rhs_fn(&g, 0.0, state, out);
    return 0;
}

/* Row-block threaded RK4 for the grid: each thread owns rows [lo,hi) of every vector; barriers
 * separate "write tmp" from "read neighbours' tmp".  Same arithmetic as orc_rk4, element by element. */
typedef struct {
    orc_grid_t g; double dt; size_t n_steps; double *y, *k1, *k2, *k3, *k4, *tmp;
    pthread_barrier_t* bar;
} orc_grid_job_t;

static void* orc_grid_worker(void* arg) {
    orc_grid_job_t* j = (orc_grid_job_t*)arg;
    const size_t lo = 4 * j->g.row_lo * j->g.cols, hi = 4 * j->g.row_hi * j->g.cols;
    const double dt = j->dt;
    double *y = j->y, *k1 = j->k1, *k2 = j->k2, *k3 = j->k3, *k4 = j->k4, *tmp = j->tmp;
    for (size_t step = 0; step < j->n_steps; ++step) {
        orc_grid_rhs_rows(&j->g, y, k1);
        for (size_t i = lo; i < hi; ++i) { k1[i] *= dt; tmp[i] = y[i] + 0.5 * k1[i]; }
        pthread_barrier_wait(j->bar);
        orc_grid_rhs_rows(&j->g, tmp, k2);
        pthread_barrier_wait(j->bar);
        for (size_t i = lo; i < hi; ++i) { k2[i] *= dt; tmp[i] = y[i] + 0.5 * k2[i]; }
        pthread_barrier_wait(j->bar);
        orc_grid_rhs_rows(&j->g, tmp, k3);
        pthread_barrier_wait(j->bar);
        for (size_t i = lo; i < hi; ++i) { k3[i] *= dt; tmp[i] = y[i] + k3[i]; }
        pthread_barrier_wait(j->bar);
        orc_grid_rhs_rows(&j->g, tmp, k4);
        for (size_t i = lo; i < hi; ++i) k4[i] *= dt;
        for (size_t i = lo; i < hi; ++i)
            y[i] += (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
        pthread_barrier_wait(j->bar);
    }
    return NULL;
}

/* in place: state <- state after n_steps RK4Solver::solve steps of size dt */
ORC_API int orc_grid2d_rk4(size_t rows, size_t cols, int topo, double mass, double spacing, double k,
                           double c, double dt, size_t n_steps, double* state, int nthreads) {
    size_t dim = 4 * rows * cols;
    double* work = (double*)malloc(sizeof(double) * 5 * dim);
    if (!work) return -1;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > rows) nthreads = (int)rows;
    if (nthreads > 256) nthreads = 256;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
    orc_grid_job_t jobs[256];
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) {
        orc_grid_t g = {rows, cols, topo, mass, spacing, k, c,
                        rows * (size_t)t / (size_t)nthreads, rows * (size_t)(t + 1) / (size_t)nthreads};
        orc_grid_job_t jb = {g, dt, n_steps, state, work, work + dim, work + 2 * dim, work + 3 * dim,
                             work + 4 * dim, &bar};
        jobs[t] = jb;
    }
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, orc_grid_worker, &jobs[t]);
    orc_grid_worker(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    free(work);
    return (int)n_steps;
}
