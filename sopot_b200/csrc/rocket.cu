// rocket.cu -- K1 instantiation for the 6-DOF rocket, Rocket<double>::SystemType
// (rocket/rocket.hpp:73-85): 11 components, 14 states
//   [0] t  [1..3] pos ENU  [4..6] vel ENU  [7..10] q1..q4 (scalar last)  [11..13] body rates.
//
// One trajectory per thread, state + RK4 accumulators register-resident.  Each distinct state
// function is evaluated ONCE per derivative call (the reference's registry re-evaluates its whole
// sub-tree on every query, core/typed_component.hpp:249-261; state functions are pure, so the
// de-duplicated evaluation is value-identical):
//   rotation matrix R(q)        1x  (reference: 4x  -- thrust, aero force in/out, aero moment)
//   atmosphere P(h), T(h)       1x  (reference: pow() 3x -- thrust, force density, moment density)
//   body velocity, airspeed     1x  (reference: 2x)
//   mass(t)                     1x  (reference: 2x)
// Terms the reference multiplies by an exact zero are dropped where that cannot change a value:
//   thrust_body = (T,0,0)  =>  R*thrust_body = (R00*T, R10*T, R20*T)      (quaternion.hpp:115-122)
//   omega_quat.q4 = 0      =>  the q*p4 products vanish                    (quaternion.hpp:80-90)
//   Cl = Cs = 0 without coefficient tables (axisymmetric_aero.hpp:154,166) => lift and side force
//   are +-0 and the two atan2 / sin / cos that only feed them are not evaluated.
// Engine / mass / inertia tables and the 7-layer atmosphere are staged in shared memory by every
// block; all lanes of a warp read the same time interval (state[0] advances identically in every
// instance), so table reads are shared-memory broadcasts.
#include "ensemble.cuh"
#include "trig.cuh"
#include <cmath>
#include <vector>

namespace {

// StandardAtmosphere, rocket/standard_atmosphere.hpp:18-32.  pow_exp = g0*M/(R*lapse) (:64) and
// exp_den = R*T_base (:62) are formed on the host in the reference's order.
struct AtmLayer { double h_top, P_base, T_base, lapse, h_base, pow_exp, exp_den; };
__constant__ AtmLayer c_atm[7];

constexpr double kAtmR = 8.3144598, kAtmM = 0.0289644, kAtmGamma = 1.4, kAtmG0 = 9.80665;
constexpr double kAtmP0 = 101325.0, kAtmT0 = 288.15;
constexpr double kAtmRs = kAtmR / kAtmM;
constexpr double kEnginePi = 3.14159265358979323846;   // interpolated_engine.hpp:34
constexpr double kAeroPi = 3.14159265358979;            // axisymmetric_aero.hpp:56 (truncated, literal)

// A divisor used several times: IEEE division per use when Strict (the reference divides each time); otherwise
// one fast reciprocal (MUFU seed + 3 FMA, <= 1 ulp, trig.cuh) and a multiply per use.
template <bool S>
struct Recip {
    Real<S> d, inv;
    __device__ __forceinline__ explicit Recip(Real<S> x) : d(x) { if constexpr (!S) inv = Real<S>(sbtrig::rcp_fast(x.v)); }
    __device__ __forceinline__ Real<S> of(Real<S> a) const { if constexpr (S) return a / d; else return a * inv; }
};

// pow(base, e) of standard_atmosphere.hpp:64.  Strict: libdevice pow (<= 1 ulp, the reference calls glibc pow).
// Otherwise exp(e * log(base)): base = T_b / T lies in (0.7, 1.4) and |e| <= 35, so the relative error is
// |e log(base)| * ulp-level ~ 1e-15, two orders inside the 1e-12 per-step promise, at a quarter of the instructions.
template <bool S>
__device__ __forceinline__ Real<S> r_pow(Real<S> base, double e) {
    if constexpr (S) { SB_COUNT_TRANSCENDENTAL(); return Real<S>(pow(base.v, e)); } else return Real<S>(exp(e * log(base.v)));
}

// one effective coefficient table inside the packed buffer: row_vals[rows], col_vals[cols], data[rows][cols] from `off`
struct Table2D { int rows, cols, off; };

struct RocketTablesDev {
    // one packed device buffer: engine columns time, exit_d, p_comb, press_ratio, engine_force
    // (5 * er), then mass time/value (2 * mr), ix (2 * xr), iyz (2 * yr), then the effective aero tables
    const double* packed;
    int total;                     // doubles in the packed buffer
    Table2D cd, cl, cs, cmq;       // rows == 0: reference default (host-side copy of the descriptors)
    int desc_off;                  // the same 4 descriptors as 12 doubles (rows, cols, off) inside the packed buffer
    int inv_off;                   // reciprocal interval widths 1 / (x[i+1] - x[i]) of the four time axes (er + mr + xr + yr)
    int mdot_off;                  // mass-flow column of the engine table (er values; read by the state-function probe only)
    int er, mr, xr, yr;
    double burn_time;
    int use_interp;
};

template <bool S>
struct Lerp {   // io/interpolation.hpp:49-71 on a shared time grid
    int i;          // interval index, or -1: clamp low, -2: clamp high
    Real<S> t;      // (x - x0) / (x1 - x0)
    __device__ __forceinline__ Real<S> eval(const double* y, int n) const {
        if (i == -1) return Real<S>(y[0]);
        if (i == -2) return Real<S>(y[n - 1]);
        const Real<S> y0(y[i]), y1(y[i + 1]);
        return y0 + t * (y1 - y0);
    }
};

// `hint` carries the interval of the previous lookup on this table: simulation time only moves forward, so the
// interval is almost always unchanged or the next one and the binary search is skipped (same index either way).
template <bool S>
__device__ __forceinline__ Lerp<S> lerp_locate(const double* x, const double* inv, int n, double xv, int& hint) {
    Lerp<S> L;
    L.t = Real<S>(0.0);
    if (xv <= x[0]) { L.i = -1; return L; }
    if (xv >= x[n - 1]) { L.i = -2; return L; }
    int i = hint;
    if (!(x[i] < xv && xv <= x[i + 1])) {         // lower_bound semantics: x[i] < xv <= x[i+1]
        if (i + 2 < n && x[i + 1] < xv && xv <= x[i + 2]) {
            i = i + 1;
        } else {
            int lo = 0, hi = n;                   // std::lower_bound: first element >= xv
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < xv) lo = mid + 1; else hi = mid; }
            i = lo > 0 ? lo - 1 : 0;
        }
        hint = i;
    }
    const Real<S> x0(x[i]), x1(x[i + 1]);
    L.i = i;
    if constexpr (S) L.t = (Real<S>(xv) - x0) / (x1 - x0);
    else L.t = Real<S>((xv - x0.v) * inv[i]);
    return L;
}

// BilinearInterpolator::interpolate, io/interpolation.hpp:137-181: clamp both coordinates to the table, lower_bound
// interval (index clamped to size-2), tc along the columns first, then tr along the rows.
__device__ __forceinline__ int interval_of(const double* x, int n, double v) {
    int lo = 0, hi = n;                                   // std::lower_bound
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (x[mid] < v) lo = mid + 1; else hi = mid; }
    int i = lo > 0 ? lo - 1 : 0;
    if (i >= n - 1) i = n - 2;
    return i;
}
template <bool S>
__device__ __forceinline__ Real<S> bilinear(const double* tb, const Table2D& t, double row, double col) {
    using R = Real<S>;
    const double* rv = tb + t.off;
    const double* cv = rv + t.rows;
    const double* data = cv + t.cols;
    if (row <= rv[0]) row = rv[0]; else if (row >= rv[t.rows - 1]) row = rv[t.rows - 1];
    if (col <= cv[0]) col = cv[0]; else if (col >= cv[t.cols - 1]) col = cv[t.cols - 1];
    const int ri = interval_of(rv, t.rows, row), ci = interval_of(cv, t.cols, col);
    const R tr = (R(row) - R(rv[ri])) / (R(rv[ri + 1]) - R(rv[ri]));
    const R tc = (R(col) - R(cv[ci])) / (R(cv[ci + 1]) - R(cv[ci]));
    const R v00(data[ri * t.cols + ci]), v01(data[ri * t.cols + ci + 1]);
    const R v10(data[(ri + 1) * t.cols + ci]), v11(data[(ri + 1) * t.cols + ci + 1]);
    const R v0 = v00 + tc * (v01 - v00), v1 = v10 + tc * (v11 - v10);
    return v0 + tr * (v1 - v0);
}

// one layer of the standard atmosphere (standard_atmosphere.hpp:50-64): gradient layers by the power law, isothermal
// ones by the exponential
template <bool S>
__device__ __forceinline__ void atmosphere_layer(const AtmLayer& l, const Real<S> alt, Real<S>& P, Real<S>& T) {
    const Real<S> dh = alt - Real<S>(l.h_base);
    if (fabs(l.lapse) < 1e-10) {
        T = Real<S>(l.T_base);
        // P_base * exp(-(g0*M) * (h - h_base) / (R*T_base))                    :62
        const Real<S> arg = Recip<S>(Real<S>(l.exp_den)).of(Real<S>(-(kAtmG0 * kAtmM)) * dh);
        if constexpr (S) { SB_COUNT_TRANSCENDENTAL(); P = Real<S>(l.P_base) * Real<S>(exp(arg.v)); }
        else P = Real<S>(l.P_base * sbtrig::exp_lean(arg.v));                   // |arg| <= 1.5 within a layer (host check)
    } else {
        const Real<S> Tl = Real<S>(l.T_base) + Real<S>(l.lapse) * dh;           // :50, :63
        T = Tl;
        if constexpr (S) {
            const Real<S> ratio = Real<S>(l.T_base) / Tl;
            P = Real<S>(l.P_base) * r_pow<S>(ratio, l.pow_exp);                 // :64
        } else {
            // (T_b / T)^e = exp(e log(T_b / T)) with the lean log-of-a-ratio and exp of trig.cuh (40 FP64 instructions
            // instead of ~95 through libdevice log + exp).  Their argument ranges hold for every layer of the built-in
            // table (checked on the host in upload_atmosphere), so there is no fallback path to keep registers alive.
            const double x = l.pow_exp * sbtrig::log_ratio_lean(l.T_base, Tl.v);
            P = Real<S>(l.P_base * sbtrig::exp_lean(x));
        }
    }
}

template <bool S>
__device__ __forceinline__ void atmosphere(const Real<S> alt, Real<S>& P, Real<S>& T) {
    const double h = alt.v;
    if (h <= 0) { P = Real<S>(kAtmP0); T = Real<S>(kAtmT0); return; }          // :45,:57
    // the lowest layer first (sounding-rocket altitudes): its constants are read with static constant-bank offsets, and the
    // search over the other layers (six FP64 compares and a register-indexed constant load) is skipped
    if (h <= c_atm[0].h_top) { atmosphere_layer<S>(c_atm[0], alt, P, T); return; }
    if (h >= 100000.0) { P = Real<S>(1e-4); T = Real<S>(186.95); return; }      // :46,:58
    int li = 6;
#pragma unroll
    for (int k = 5; k >= 1; --k) if (h <= c_atm[k].h_top) li = k;              // first layer with h <= h_top
    const AtmLayer l = c_atm[li];
    atmosphere_layer<S>(l, alt, P, T);
}

}  // namespace

#ifndef SB_ROCKET_BLOCK
#define SB_ROCKET_BLOCK 128
#endif
#ifndef SB_ROCKET_MINB
#define SB_ROCKET_MINB 1
#endif
#ifndef SB_ROCKET_SMEM_STATE
#define SB_ROCKET_SMEM_STATE 1
#endif
struct RocketDevParams {
    ParamRef elev, az, diam, g;
    RocketTablesDev tab;
    double* stats_out;   // [8][n] or nullptr
};

// AT: coefficient tables present (a separate instantiation: the table path needs atan2 / sincos / bilinear lookups and
// ~60 more registers, which would cost the default configuration one resident block per SM)
template <bool AT>
struct RocketSysT {
    static constexpr int DIM = 14;
    // default model: 3 x 128 threads per SM (<= 168 registers) in strict mode, 4 x 128 (<= 128 registers) in contract mode,
    // whose step keeps the base state and the stage sum in shared memory (measured: 220.7 -> 214.8 ms; 5 CTAs: 256 ms)
    static constexpr int BLOCK = SB_ROCKET_BLOCK, IPT = 1, MIN_BLOCKS = SB_ROCKET_MINB > 1 ? SB_ROCKET_MINB : (AT ? 1 : 3);
    static constexpr int MIN_BLOCKS_CONTRACT = SB_ROCKET_MINB > 1 ? SB_ROCKET_MINB : (AT ? 1 : (SB_ROCKET_SMEM_STATE ? 4 : 3));
    static constexpr bool FAST_STEP = true;
    using DevParams = RocketDevParams;
    template <bool S> struct Inst {
        Real<S> g, area, len;
        const double* tb;          // shared-memory copy of the packed tables
        int er, mr, xr, yr, use_interp;
        int desc;                  // offset of the aero table descriptors in the shared-memory tables (AT only)
        int inv;                   // offset of the reciprocal interval widths
        int mdot;                  // offset of the mass-flow column
        int soff;                  // offset of the per-thread state arrays behind the tables
        double burn;
        double apogee, apogee_t, vmax, vmax_t, bo_alt, bo_spd;
        mutable int h_eng, h_mass, h_ix, h_iyz;   // lerp_locate interval hints
    };
    // shared memory: the packed tables, then (contract-mode step) two [14][BLOCK] arrays holding the step's base state and
    // the running sum of the stage derivatives, so that neither occupies registers during the derivative evaluations
    static constexpr int kStateSmemDoubles = SB_ROCKET_SMEM_STATE ? 2 * 14 * BLOCK : 0;
    static size_t smem_bytes(const DevParams& P) {
        return sizeof(double) * ((size_t)((P.tab.total + 1) & ~1) + kStateSmemDoubles);
    }
    template <bool S> static __device__ __forceinline__ void block_setup(const DevParams& P) {
        extern __shared__ double s_tab[];
        for (int j = threadIdx.x; j < P.tab.total; j += blockDim.x) s_tab[j] = P.tab.packed[j];
        __syncthreads();
    }
    template <bool S>
    static __device__ __forceinline__ void load(const DevParams& P, size_t i, Inst<S>& in, Real<S> (&y)[14]) {
        extern __shared__ double s_tab[];
        in.tb = s_tab;
        in.er = P.tab.er; in.mr = P.tab.mr; in.xr = P.tab.xr; in.yr = P.tab.yr;
        in.use_interp = P.tab.use_interp; in.burn = P.tab.burn_time;
        in.desc = P.tab.desc_off; in.inv = P.tab.inv_off; in.mdot = P.tab.mdot_off; in.soff = (P.tab.total + 1) & ~1;
        const double d = P.diam.at(i);
        in.g = Real<S>(P.g.at(i));
        in.area = Real<S>(kAeroPi) * Real<S>(d) * Real<S>(d) / Real<S>(4.0);   // axisymmetric_aero.hpp:56
        in.len = Real<S>(d);
        in.h_eng = in.h_mass = in.h_ix = in.h_iyz = 0;
        in.apogee = 0; in.apogee_t = 0; in.vmax = 0; in.vmax_t = 0; in.bo_alt = 0; in.bo_spd = 0;
#pragma unroll
        for (int k = 0; k < 14; ++k) y[k] = Real<S>(0.0);
        // Quaternion::from_launcher_angles, rocket/quaternion.hpp:164-190
        const Real<S> deg2rad = Real<S>(3.14159265358979323846 / 180.0);
        const Real<S> yaw = (Real<S>(90.0) - Real<S>(P.az.at(i))) * deg2rad;
        const Real<S> pitch = Real<S>(P.elev.at(i)) * deg2rad;
        Real<S> sy, cy, sp, cp;
        r_sincos(yaw * Real<S>(0.5), sy, cy);
        r_sincos(pitch * Real<S>(0.5), sp, cp);
        y[7] = sy * sp; y[8] = (-cy) * sp; y[9] = sy * cp; y[10] = cy * cp;
    }

    // everything the derivative reads from the TIME-indexed tables: mass / inertia (rocket_body.hpp:121-179) and the engine
    // columns (interpolated_engine.hpp:103-152).  A pure function of the simulation time, so RK4 stages 2 and 3 (same time)
    // share one evaluation in the contract-mode step below.
    template <bool S> struct TimeTab { Real<S> mass, Ix, Iyz, exit_area, exit_pressure, engine_force, mass_flow; bool burning; };
    template <bool S, bool PROBE = false>
    static __device__ __forceinline__ TimeTab<S> time_tables(const Inst<S>& in, const double time) {
        using R = Real<S>;
        const double* tb = in.tb;
        const double* eng_t = tb;
        const double* mass_t = tb + 5 * in.er;
        const double* ix_t = mass_t + 2 * in.mr;
        const double* iyz_t = ix_t + 2 * in.xr;
        const double* inv_e = tb + in.inv;
        const double* inv_m = inv_e + in.er;
        const double* inv_x = inv_m + in.mr;
        const double* inv_y = inv_x + in.xr;
        TimeTab<S> tt{R(100.0), R(10.0), R(100.0), R(0.0), R(0.0), R(0.0), R(0.0), false};
        if (in.use_interp) {
            if (in.mr) tt.mass = lerp_locate<S>(mass_t, inv_m, in.mr, time, in.h_mass).eval(mass_t + in.mr, in.mr);
            if (in.xr) tt.Ix = lerp_locate<S>(ix_t, inv_x, in.xr, time, in.h_ix).eval(ix_t + in.xr, in.xr);
            if (in.yr) tt.Iyz = lerp_locate<S>(iyz_t, inv_y, in.yr, time, in.h_iyz).eval(iyz_t + in.yr, in.yr);
        }
        tt.burning = in.er > 0 && time <= in.burn;
        if (tt.burning) {
            const Lerp<S> L = lerp_locate<S>(eng_t, inv_e, in.er, time, in.h_eng);
            const R d_exit = L.eval(eng_t + in.er, in.er);
            const R p_comb = L.eval(eng_t + 2 * in.er, in.er);
            const R press_ratio = L.eval(eng_t + 3 * in.er, in.er);
            tt.exit_area = r_div_const(R(kEnginePi) * d_exit * d_exit, 4.0);
            tt.exit_pressure = p_comb * press_ratio;
            tt.engine_force = L.eval(eng_t + 4 * in.er, in.er);
            if constexpr (PROBE) tt.mass_flow = L.eval(tb + in.mdot, in.er);
        }
        return tt;
    }

    template <bool S>
    static __device__ __forceinline__ int rhs(const Inst<S>& in, const Real<S> (&s)[14], Real<S> (&d)[14]) {
        return rhs_tt<S>(in, time_tables<S>(in, s[0].v), s, d);
    }

    // Contract-mode RK4 step (rk4_step_fast of ensemble.cuh with the table lookups hoisted): the time component advances by
    // exactly dt/2, dt/2, dt, so stages 2 and 3 read the tables at the same instant.
#if SB_ROCKET_SMEM_STATE
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[14], const StepConsts& c, const size_t) {
        // y (base state) and F (sum of stage derivatives) live in shared memory, column threadIdx.x of [14][BLOCK]
        // (conflict-free): 56 fewer live registers across the four derivative evaluations -> one more CTA per SM
        double* sy = const_cast<double*>(in.tb) + in.soff + threadIdx.x;
        double* sF = sy + 14 * BLOCK;
        Real<false> f[14], tmp[14];
        int st = rhs_tt<false>(in, time_tables<false>(in, y[0].v), y, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { sy[d * BLOCK] = y[d].v; sF[d * BLOCK] = f[d].v; tmp[d].v = fma(c.hdt, f[d].v, y[d].v); }
        const TimeTab<false> mid = time_tables<false>(in, tmp[0].v);
        st |= rhs_tt<false>(in, mid, tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { sF[d * BLOCK] = fma(2.0, f[d].v, sF[d * BLOCK]); tmp[d].v = fma(c.hdt, f[d].v, sy[d * BLOCK]); }
        st |= rhs_tt<false>(in, mid, tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { sF[d * BLOCK] = fma(2.0, f[d].v, sF[d * BLOCK]); tmp[d].v = fma(c.dt, f[d].v, sy[d * BLOCK]); }
        st |= rhs_tt<false>(in, time_tables<false>(in, tmp[0].v), tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) y[d].v = fma(c.dt6, sF[d * BLOCK] + f[d].v, sy[d * BLOCK]);
        return st;
    }
#else
    static __device__ __forceinline__ int fast_step(Inst<false>& in, Real<false> (&y)[14], const StepConsts& c, const size_t) {
        Real<false> f[14], F[14], tmp[14];
        int st = rhs_tt<false>(in, time_tables<false>(in, y[0].v), y, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { F[d] = f[d]; tmp[d].v = fma(c.hdt, f[d].v, y[d].v); }
        const TimeTab<false> mid = time_tables<false>(in, tmp[0].v);
        st |= rhs_tt<false>(in, mid, tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { F[d].v = fma(2.0, f[d].v, F[d].v); tmp[d].v = fma(c.hdt, f[d].v, y[d].v); }
        st |= rhs_tt<false>(in, mid, tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) { F[d].v = fma(2.0, f[d].v, F[d].v); tmp[d].v = fma(c.dt, f[d].v, y[d].v); }
        st |= rhs_tt<false>(in, time_tables<false>(in, tmp[0].v), tmp, f);
#pragma unroll
        for (int d = 0; d < 14; ++d) y[d].v = fma(c.dt6, F[d].v + f[d].v, y[d].v);
        return st;
    }
#endif

    // PROBE: also write the values of the reference's state functions (rocket_tags.hpp) that are intermediates of the derivative
    // to probe[kProbeValues] -- what SimulationResult::query<Tag>(t) / Rocket::queryStateFunction<Tag> return.
    template <bool S, bool PROBE = false>
    static __device__ __forceinline__ int rhs_tt(const Inst<S>& in, const TimeTab<S>& tt, const Real<S> (&s)[14], Real<S> (&d)[14],
                                                 double* probe = nullptr) {
        using R = Real<S>;
        const R alt = s[3];
        const R q1 = s[7], q2 = s[8], q3 = s[9], q4 = s[10];
        const R wx = s[11], wy = s[12], wz = s[13];
        const double* tb = in.tb;
        (void)tb;

        // ---- rotation matrix, quaternion.hpp:94-112 (once) ----
        const R q1_2 = q1 * q1, q2_2 = q2 * q2, q3_2 = q3 * q3, q4_2 = q4 * q4;
        const R q1q2 = q1 * q2, q1q3 = q1 * q3, q1q4 = q1 * q4, q2q3 = q2 * q3, q2q4 = q2 * q4, q3q4 = q3 * q4;
        const R R00 = q4_2 + q1_2 - q2_2 - q3_2, R01 = R(2.0) * (q1q2 - q3q4), R02 = R(2.0) * (q1q3 + q2q4);
        const R R10 = R(2.0) * (q1q2 + q3q4), R11 = q4_2 - q1_2 + q2_2 - q3_2, R12 = R(2.0) * (q2q3 - q1q4);
        const R R20 = R(2.0) * (q1q3 - q2q4), R21 = R(2.0) * (q2q3 + q1q4), R22 = q4_2 - q1_2 - q2_2 + q3_2;

        // ---- body velocity and airspeed, axisymmetric_aero.hpp:177-188 ----
        const R vbx = R00 * s[4] + R10 * s[5] + R20 * s[6];
        const R vby = R01 * s[4] + R11 * s[5] + R21 * s[6];
        const R vbz = R02 * s[4] + R12 * s[5] + R22 * s[6];
        const R v2 = vbx * vbx + vby * vby + vbz * vbz;
        R airspeed, inv_airspeed;
        if constexpr (S) airspeed = r_sqrt(v2);
        else if (v2.v >= 1.0) sbtrig::sqrt_rsqrt_fast(v2.v, airspeed.v, inv_airspeed.v);   // else: aero off, root unused
        const bool aero_on = S ? !(airspeed.v < 1.0) : (v2.v >= 1.0 && !(airspeed.v < 1.0));

        const R mass = tt.mass, Ix = tt.Ix, Iyz = tt.Iyz;

        // ---- thrust, interpolated_engine.hpp:103-152 ----
        const bool burning = tt.burning;
        R P(0.0), T(0.0);
        if (PROBE || burning || aero_on) atmosphere<S>(alt, P, T);
        R thrust(0.0);
        if (burning) thrust = tt.engine_force + tt.exit_area * (tt.exit_pressure - P);

        // ---- aerodynamic force (ENU) and damping moment (body) ----
        R Fax(0.0), Fay(0.0), Faz(0.0), Mx(0.0), My(0.0), Mz(0.0);
        if (aero_on) {
            const R density = Recip<S>(R(kAtmRs) * T).of(P);                      // standard_atmosphere.hpp:70
            const R dynp = R(0.5) * density * airspeed * airspeed;                // :198
            const R qS = dynp * in.area;                                          // :212
            R ux(0.0), uy(0.0), uz(0.0);
            if constexpr (S) { if (!(airspeed.v < 1e-15)) { ux = vbx / airspeed; uy = vby / airspeed; uz = vbz / airspeed; } }
            else { ux = vbx * inv_airspeed; uy = vby * inv_airspeed; uz = vbz * inv_airspeed; }
            R cd(0.3), cmq(-450.0);                                               // defaults :142, :275
            R fbx, fby, fbz;
            if constexpr (!AT) {
                // Cl = Cs = 0 (:154,:166): lift and side force are +-0, the angles that only feed them are skipped
                const R drag = qS * cd;
                fbx = (-ux) * drag; fby = (-uy) * drag; fbz = (-uz) * drag;       // :228
            } else {
                // table-driven coefficients (:133-167): speed of sound, Mach, angles of attack / sideslip
                const R sos = r_sqrt(R(kAtmGamma * kAtmRs) * T);                  // standard_atmosphere.hpp:71
                const R mach = airspeed / sos;                                    // :197
                const R rad2deg = R(180.0) / R(kAeroPi);                          // :175
                const R aoa = (vbx * vbx + vbz * vbz).v < 1e-10 ? R(0.0) : R(atan2(-vbz.v, vbx.v));    // :116-122
                const R aoss = (vbx * vbx + vby * vby).v < 1e-10 ? R(0.0) : R(atan2(vby.v, vbx.v));   // :125-131
                const double aoa_deg = fabs((aoa * rad2deg).v), aoss_deg = fabs((aoss * rad2deg).v);  // :203-204
                auto desc = [&](int k) { const double* q = tb + in.desc + 3 * k; return Table2D{(int)q[0], (int)q[1], (int)q[2]}; };
                const Table2D t_cd = desc(0), t_cl = desc(1), t_cs = desc(2), t_cmq = desc(3);
                if (t_cd.rows) cd = bilinear<S>(tb, t_cd, mach.v, aoa_deg);
                const R cl = t_cl.rows ? bilinear<S>(tb, t_cl, mach.v, aoa_deg) : R(0.0);
                const R cs = t_cs.rows ? bilinear<S>(tb, t_cs, mach.v, aoss_deg) : R(0.0);
                const R drag = qS * cd, lift = qS * cl, side = qS * cs;
                R sa, ca, ss, cs_unused;
                r_sincos(aoa, sa, ca);
                r_sincos(aoss, ss, cs_unused);
                const R ldx = -sa, ldz = aoa.v < 0 ? -ca : ca;                    // :232-233
                const R sdy = aoss.v < 0 ? -(-ss) : -ss;                          // :237-238
                // F_drag + F_lift + F_side, component by component (:240)
                fbx = ((-ux) * drag + ldx * lift) + R(0.0) * side;
                fby = ((-uy) * drag + R(0.0) * lift) + sdy * side;
                fbz = ((-uz) * drag + ldz * lift) + R(0.0) * side;
                if (t_cmq.rows) {                                                 // :276-281, signed angle, arguments swapped
                    const R aoa_signed_deg = aoa * R(180.0) / R(kAeroPi);
                    cmq = bilinear<S>(tb, t_cmq, aoa_signed_deg.v, mach.v);
                }
            }
            if constexpr (PROBE) { probe[11] = fbx.v; probe[12] = fby.v; probe[13] = fbz.v; }
            Fax = R00 * fbx + R01 * fby + R02 * fbz;                              // :251-253
            Fay = R10 * fbx + R11 * fby + R12 * fbz;
            Faz = R20 * fbx + R21 * fby + R22 * fbz;
            const R qSd = qS * in.len;                                            // :271
            R damping;                                                            // :274
            if constexpr (S) damping = in.len / airspeed; else damping = in.len * inv_airspeed;
            const R cq = cmq * qSd * damping;                                     // :283-287
            Mx = cq * wx; My = cq * wy; Mz = cq * wz;
        }

        // ---- ForceAggregator, force_aggregator.hpp:61-79 : (F_gravity + F_thrust) + F_aero ----
        const R Fx = R00 * thrust + Fax;
        const R Fy = R10 * thrust + Fay;
        const R Fz = ((-in.g) * mass + R20 * thrust) + Faz;

        if constexpr (PROBE) {
            probe[0] = mass.v; probe[1] = Ix.v; probe[2] = Iyz.v; probe[3] = Iyz.v;
            probe[4] = Fx.v; probe[5] = Fy.v; probe[6] = Fz.v;                    // dynamics::TotalForceENU
            probe[7] = Mx.v; probe[8] = My.v; probe[9] = Mz.v;                    // dynamics::TotalTorqueBody (= the aero moment)
            probe[10] = in.g.v;                                                   // dynamics::GravityAcceleration
            if (!aero_on) { probe[11] = 0.0; probe[12] = 0.0; probe[13] = 0.0; }  // aero::AeroForceBody (:186: zero below 1 m/s)
            probe[14] = Fax.v; probe[15] = Fay.v; probe[16] = Faz.v;              // aero::AeroForceENU
            probe[17] = Mx.v; probe[18] = My.v; probe[19] = Mz.v;                 // aero::AeroMomentBody
            probe[20] = (P / (R(kAtmRs) * T)).v;                                  // environment::AtmosphericDensity (:70)
            probe[21] = P.v; probe[22] = T.v;
            probe[23] = r_sqrt(R(kAtmGamma * kAtmRs) * T).v;                      // environment::SpeedOfSound (:71)
            probe[24] = tt.mass_flow.v;                                           // propulsion::MassFlowRate
            probe[25] = thrust.v; probe[26] = 0.0; probe[27] = 0.0;               // propulsion::ThrustForceBody (along body +X)
        }
        d[0] = R(1.0);                                                            // simulation_time.hpp:46-54
        d[1] = s[4]; d[2] = s[5]; d[3] = s[6];                                    // translation_kinematics.hpp:46-55
        const Recip<S> by_mass(mass);
        d[4] = by_mass.of(Fx); d[5] = by_mass.of(Fy); d[6] = by_mass.of(Fz);      // translation_dynamics.hpp:35-42
        d[7] = ((q4 * wx + q2 * wz) - q3 * wy) * R(0.5);                          // quaternion.hpp:80-90
        d[8] = ((q4 * wy - q1 * wz) + q3 * wx) * R(0.5);
        d[9] = ((q4 * wz + q1 * wy) - q2 * wx) * R(0.5);
        d[10] = (((-(q1 * wx)) - q2 * wy) - q3 * wz) * R(0.5);
        const Recip<S> by_Ix(Ix), by_Iyz(Iyz);
        d[11] = by_Ix.of(Mx - (Iyz - Iyz) * wy * wz);                             // rotation_dynamics.hpp:42-44
        d[12] = by_Iyz.of(My - (Ix - Iyz) * wx * wz);
        d[13] = by_Iyz.of(Mz - (Iyz - Ix) * wx * wy);
        return 0;
    }

    // SimulationResult::computeStatistics, rocket/simulation_result.hpp:151-189, fused into the loop
    template <bool S>
    static __device__ __forceinline__ void observe(Inst<S>& in, double t, const Real<S> (&y)[14]) {
        const double alt = y[3].v;
        const Real<S> sp2 = y[4] * y[4] + y[5] * y[5] + y[6] * y[6];
#if defined(SB_COUNT_OPS) && defined(__CUDA_ARCH__)
        if (S) sb_count(3);                                   // the statistics' own root (plain sqrt below)
#endif
        const double spd = sqrt(sp2.v);
        if (alt > in.apogee) { in.apogee = alt; in.apogee_t = t; }
        if (spd > in.vmax) { in.vmax = spd; in.vmax_t = t; }
        if (in.burn > 0 && fabs(t - in.burn) < 0.1) { in.bo_alt = alt; in.bo_spd = spd; }
    }
    template <bool S>
    static __device__ __forceinline__ int finish(const DevParams& P, const Inst<S>& in, size_t i, size_t n,
                                                  const Real<S> (&y)[14]) {
        if (!P.stats_out) return 0;
        double* o = P.stats_out + i;
        o[0] = in.apogee; o[n] = in.apogee_t; o[2 * n] = in.vmax; o[3 * n] = in.vmax_t;
        o[4 * n] = in.bo_alt; o[5 * n] = in.bo_spd; o[6 * n] = y[1].v; o[7 * n] = y[2].v;
        return 0;
    }
};

constexpr int kProbeValues = 28;

// one evaluation of the state functions per instance: out[kProbeValues][n]
template <class Sys>
__global__ void __launch_bounds__(Sys::BLOCK)
rocket_probe_kernel(const typename Sys::DevParams P, const size_t n, const double* __restrict__ state, double* __restrict__ out) {
    Sys::template block_setup<true>(P);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename Sys::template Inst<true> in;
    Real<true> y[14], dy[14];
    Sys::template load<true>(P, i, in, y);
#pragma unroll
    for (int d = 0; d < 14; ++d) y[d] = Real<true>(state[(size_t)d * n + i]);
    double probe[kProbeValues];
    Sys::template rhs_tt<true, true>(in, Sys::template time_tables<true, true>(in, y[0].v), y, dy, probe);
#pragma unroll
    for (int k = 0; k < kProbeValues; ++k) out[(size_t)k * n + i] = probe[k];
}

using RocketSys = RocketSysT<false>;
using RocketSysTables = RocketSysT<true>;

// ---- host side ------------------------------------------------------------------------------------
namespace {

// LinearInterpolator::interpolate (io/interpolation.hpp:49-71) on host columns
double host_lerp(const std::vector<double>& x, const double* y, double xv) {
    const size_t n = x.size();
    if (n == 0) return 0.0;
    if (xv <= x.front()) return y[0];
    if (xv >= x.back()) return y[n - 1];
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (x[mid] < xv) lo = mid + 1; else hi = mid; }
    size_t i = lo ? lo - 1 : 0;
    const double t = (xv - x[i]) / (x[i + 1] - x[i]);
    return y[i] + t * (y[i + 1] - y[i]);
}

// InterpolatedEngine::computePressureRatio, interpolated_engine.hpp:225-252
double nozzle_pressure_ratio(double k, double area_ratio) {
    double pr = 0.01;
    for (int it = 0; it < 20; ++it) {
        double M2 = (2.0 / (k - 1.0)) * (std::pow(1.0 / pr, (k - 1.0) / k) - 1.0);
        if (M2 < 0) M2 = 0;
        const double M = std::sqrt(M2);
        const double term = (2.0 / (k + 1.0)) * (1.0 + (k - 1.0) / 2.0 * M2);
        const double area_calc = (1.0 / M) * std::pow(term, (k + 1.0) / (2.0 * (k - 1.0)));
        if (std::fabs(area_calc - area_ratio) < 0.001 * area_ratio) break;
        pr *= (area_calc > area_ratio) ? 0.95 : 1.05;
        if (pr < 1e-6) pr = 1e-6;
        if (pr > 1.0) pr = 1.0;
    }
    return pr;
}

// Pack the host tables for the device: the load-time precompute of InterpolatedEngine
// (interpolated_engine.hpp:176-222) happens here, once per call, on the host.
void pack_tables(const sopot_b200_rocket_params* p, std::vector<double>& out, RocketTablesDev& tab) {
    const size_t er = p->engine_rows, mr = p->mass_rows, xr = p->ix_rows, yr = p->iyz_rows;
    out.assign(5 * er + 2 * mr + 2 * xr + 2 * yr, 0.0);
    std::vector<double> mdot_col(er, 0.0);
    if (er) {
        std::vector<double> t(er), col(8 * er);
        for (size_t r = 0; r < er; ++r)
            for (int c = 0; c < 8; ++c) col[c * er + r] = p->engine[r * 8 + c];
        for (size_t r = 0; r < er; ++r) t[r] = col[r];
        double* o_t = out.data();
        double* o_exit = o_t + er; double* o_pc = o_exit + er; double* o_pr = o_pc + er; double* o_f = o_pr + er;
        for (size_t r = 0; r < er; ++r) {
            const double tt = t[r];
            const double k = host_lerp(t, &col[5 * er], tt), d_throat = host_lerp(t, &col[1 * er], tt);
            const double d_exit = host_lerp(t, &col[2 * er], tt), pc = host_lerp(t, &col[3 * er], tt);
            const double temp = host_lerp(t, &col[4 * er], tt), mol = host_lerp(t, &col[6 * er], tt);
            const double eff = host_lerp(t, &col[7 * er], tt);
            const double throat_area = kEnginePi * d_throat * d_throat / 4.0;
            const double exit_area = kEnginePi * d_exit * d_exit / 4.0;
            const double ratio = nozzle_pressure_ratio(k, exit_area / throat_area);
            const double throat_pressure = pc * std::pow(2.0 / (k + 1.0), k / (k - 1.0));
            const double throat_temp = temp * (2.0 / (k + 1.0));
            const double Rspec = kAtmR / mol;
            const double throat_density = throat_pressure / (Rspec * throat_temp);
            const double throat_velocity = std::sqrt(k * Rspec * throat_temp);
            double mdot = throat_density * throat_velocity * throat_area;
            if (mdot < 0) mdot = 0;
            const double exit_velocity =
                eff * std::sqrt(2.0 * k / (k - 1.0) * Rspec * temp * (1.0 - std::pow(ratio, (k - 1.0) / k)));
            o_t[r] = tt; o_exit[r] = col[2 * er + r]; o_pc[r] = col[3 * er + r];
            o_pr[r] = ratio; o_f[r] = mdot * exit_velocity;
            mdot_col[r] = mdot;
        }
    }
    double* o = out.data() + 5 * er;
    auto put2 = [&](size_t rows, const double* src) {
        for (size_t r = 0; r < rows; ++r) { o[r] = src[2 * r]; o[rows + r] = src[2 * r + 1]; }
        o += 2 * rows;
    };
    put2(mr, p->mass); put2(xr, p->ix); put2(yr, p->iyz);
    tab.er = (int)er; tab.mr = (int)mr; tab.xr = (int)xr; tab.yr = (int)yr;
    tab.burn_time = er ? p->engine[(er - 1) * 8] : 0.0;
    tab.use_interp = (mr || xr || yr) ? 1 : 0;
    // effective aero tables (getCd/getCl/getCs, axisymmetric_aero.hpp:133-167): power-on table when the engine flag is
    // set and the table is loaded, else the power-off table, else the default coefficient
    tab.cd = tab.cl = tab.cs = tab.cmq = Table2D{0, 0, 0};
    if (p->aero) {
        const sopot_b200_aero_tables& a = *p->aero;
        auto pick = [&](const sopot_b200_table2d& on, const sopot_b200_table2d& off) -> const sopot_b200_table2d* {
            if (a.engine_on && on.rows) return &on;
            if (off.rows) return &off;
            return nullptr;
        };
        auto put = [&](const sopot_b200_table2d* t, Table2D& dst) {
            if (!t) return;
            dst = Table2D{(int)t->rows, (int)t->cols, (int)out.size()};
            out.insert(out.end(), t->row_vals, t->row_vals + t->rows);
            out.insert(out.end(), t->col_vals, t->col_vals + t->cols);
            out.insert(out.end(), t->data, t->data + t->rows * t->cols);
        };
        put(pick(a.cd_power_on, a.cd_power_off), tab.cd);
        put(pick(a.cl_power_on, a.cl_power_off), tab.cl);
        put(pick(a.cs_power_on, a.cs_power_off), tab.cs);
        put(a.cmq_yz.rows ? &a.cmq_yz : nullptr, tab.cmq);
    }
    tab.desc_off = (int)out.size();
    for (const Table2D* t : {&tab.cd, &tab.cl, &tab.cs, &tab.cmq}) { out.push_back(t->rows); out.push_back(t->cols); out.push_back(t->off); }
    // contract mode: the interpolation weight is (x - x_i) * 1/(x_{i+1} - x_i) with the reciprocal read from here instead of
    // a MUFU seed + 3 dependent FMAs per lookup (n entries per axis, the last one unused)
    tab.inv_off = (int)out.size();
    {
        const size_t offs[4] = {0, 5 * er, 5 * er + 2 * mr, 5 * er + 2 * mr + 2 * xr}, ns[4] = {er, mr, xr, yr};
        for (int a = 0; a < 4; ++a)
            for (size_t r = 0; r < ns[a]; ++r)
                out.push_back(r + 1 < ns[a] ? 1.0 / (out[offs[a] + r + 1] - out[offs[a] + r]) : 0.0);
    }
    // propulsion::MassFlowRate (interpolated_engine.hpp:97-100,176-222): not part of the derivative, read by the probe only
    tab.mdot_off = (int)out.size();
    out.insert(out.end(), mdot_col.begin(), mdot_col.end());
    tab.total = (int)out.size();
}

int upload_atmosphere(sopot_b200_ctx* ctx) {
    static const double L[7][5] = {
        {11000.0, 101325.0, 288.15, -0.0065, 0.0},   {20000.0, 22632.0, 216.65, 0.0, 11000.0},
        {32000.0, 5474.89, 216.65, 0.001, 20000.0},  {47000.0, 868.02, 228.65, 0.0028, 32000.0},
        {51000.0, 110.91, 270.65, 0.0, 47000.0},     {71000.0, 66.94, 270.65, -0.0028, 51000.0},
        {100000.0, 3.96, 214.65, -0.002, 71000.0}};
    AtmLayer h[7];
    for (int i = 0; i < 7; ++i) {
        h[i] = AtmLayer{L[i][0], L[i][1], L[i][2], L[i][3], L[i][4], 0.0, kAtmR * L[i][2]};
        if (std::fabs(L[i][3]) < 1e-10 && std::fabs(kAtmG0 * kAtmM * (L[i][0] - L[i][4]) / h[i].exp_den) > sbtrig::kExpLeanMax)
            return sb_fail(ctx, SOPOT_B200_EINVAL, "isothermal atmosphere layer outside the range of the lean exp");
        if (std::fabs(L[i][3]) >= 1e-10) {
            h[i].pow_exp = kAtmG0 * kAtmM / (kAtmR * L[i][3]);
            // ranges of the lean log / exp used by the contract-mode power law, at the far end of the layer
            const double T_top = L[i][2] + L[i][3] * (L[i][0] - L[i][4]);
            const double z = (L[i][2] - T_top) / (L[i][2] + T_top);
            if (!(T_top > 0.0) || std::fabs(z) > sbtrig::kLogRatioMaxZ || std::fabs(h[i].pow_exp * std::log(L[i][2] / T_top)) > sbtrig::kExpLeanMax)
                return sb_fail(ctx, SOPOT_B200_EINVAL, "atmosphere layer outside the range of the lean power law");
        }
    }
    SB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_atm, h, sizeof h, 0, cudaMemcpyHostToDevice, ctx->stream));
    return SOPOT_B200_OK;
}

int rocket_prepare(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n, uint32_t broadcast,
                   Staging& sg, RocketSys::DevParams& P) {
    if ((p->engine_rows && !p->engine) || (p->mass_rows && !p->mass) || (p->ix_rows && !p->ix) ||
        (p->iyz_rows && !p->iyz))
        return sb_fail(ctx, SOPOT_B200_EINVAL, "table pointer is NULL with rows > 0");
    if (p->engine_rows == 1 || p->mass_rows == 1 || p->ix_rows == 1 || p->iyz_rows == 1)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "interpolation tables need at least 2 rows");
    if (p->aero) {
        const sopot_b200_table2d* t[7] = {&p->aero->cd_power_on, &p->aero->cd_power_off, &p->aero->cl_power_on, &p->aero->cl_power_off,
                                          &p->aero->cs_power_on, &p->aero->cs_power_off, &p->aero->cmq_yz};
        for (const sopot_b200_table2d* q : t) {
            if (q->rows == 0) continue;
            if (q->rows < 2 || q->cols < 2 || !q->row_vals || !q->col_vals || !q->data)
                return sb_fail(ctx, SOPOT_B200_EINVAL, "a loaded coefficient table needs >= 2 rows and columns and non-NULL arrays");
        }
    }
    int rc;
    if ((rc = upload_atmosphere(ctx))) return rc;
    std::vector<double> packed;
    pack_tables(p, packed, P.tab);
    if (packed.size() * 8 > 200 * 1024)
        return sb_fail(ctx, SOPOT_B200_EINVAL, "tables exceed the 200 KiB shared-memory staging budget");
    // the packed tables always travel host -> device (tables are host pointers by contract)
    void* d_tab = nullptr;
    if ((rc = sb_scratch(ctx, packed.size() * 8 + 256, &d_tab))) return rc;
    if (!packed.empty()) {
        SB_CUDA(ctx, cudaMemcpyAsync(d_tab, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // `packed` is a local
    }
    P.tab.packed = static_cast<const double*>(d_tab);
    const double* user[4] = {p->elevation_deg, p->azimuth_deg, p->diameter, p->gravity};
    ParamRef r[4];
    if ((rc = sb_stage_params(sg, user, 4, n, broadcast, r))) return rc;
    P.elev = r[0]; P.az = r[1]; P.diam = r[2]; P.g = r[3];
    P.stats_out = nullptr;
    return SOPOT_B200_OK;
}

template <class Sys, bool S>
int rocket_set_smem(sopot_b200_ctx* ctx, size_t bytes) {
    if (bytes > 48 * 1024) {
        SB_CUDA(ctx, cudaFuncSetAttribute(rk4_ensemble_kernel<Sys, S>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        SB_CUDA(ctx, cudaFuncSetAttribute(rhs_ensemble_kernel<Sys, S>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return SOPOT_B200_OK;
}

template <class Sys, bool S>
int rocket_launch_rk4(sopot_b200_ctx* ctx, const RocketSys::DevParams& P, size_t n, double dt, size_t n_steps, size_t store_every,
                      size_t smem, double* d_state, double* d_traj, int32_t* d_status) {
    int rc = rocket_set_smem<Sys, S>(ctx, smem);
    if (rc) return rc;
    rk4_ensemble_kernel<Sys, S><<<sb_grid_for(n, Sys::BLOCK), Sys::BLOCK, smem, ctx->stream>>>(
        P, n, 0.0, make_step_consts(dt), n_steps, store_every, d_state, d_traj, d_status);
    return SOPOT_B200_OK;
}
template <class Sys, bool S>
int rocket_launch_rhs(sopot_b200_ctx* ctx, const RocketSys::DevParams& P, size_t n, size_t smem, const double* d_state, double* d_out) {
    int rc = rocket_set_smem<Sys, S>(ctx, smem);
    if (rc) return rc;
    rhs_ensemble_kernel<Sys, S><<<sb_grid_for(n, Sys::BLOCK), Sys::BLOCK, smem, ctx->stream>>>(P, n, d_state, d_out);
    return SOPOT_B200_OK;
}

}  // namespace

SB_EXPORT int sopot_b200_rocket_rk4(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n, double dt,
                                    size_t n_steps, const sopot_b200_rk4_opts* opts, double* state_out,
                                    double* traj_out, double* stats_out, int32_t* status) {
    EnsembleArgs a{ctx, n, 0.0, dt, n_steps, opts, state_out, traj_out, status};
    int rc = sb_check_ensemble(a, p);
    if (rc) return rc;
    Staging sg(ctx, opts->mem);
    if (sg.host && (rc = sb_arena_begin(ctx, sb_ensemble_arena_bytes(4, 14, n, n_steps, opts, traj_out, status,
                                                                     sb_align(8 * n * 8)))))
        return rc;
    RocketSys::DevParams P;
    if ((rc = rocket_prepare(ctx, p, n, opts->broadcast, sg, P))) return rc;
    void* d_stats = nullptr;
    if ((rc = sg.out(stats_out, 8 * n * 8, &d_stats))) return rc;
    P.stats_out = static_cast<double*>(d_stats);

    const size_t smem = RocketSys::smem_bytes(P);
    const size_t snaps = (traj_out && opts->store_every) ? n_steps / opts->store_every : 0;
    void *d_state = nullptr, *d_traj = nullptr, *d_status = nullptr;
    if ((rc = sg.out(state_out, 14 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(traj_out, snaps * 14 * n * 8, &d_traj))) return rc;
    if ((rc = sg.out(status, n * 4, &d_status))) return rc;
    if (n) {
        const bool strict = opts->fp_mode == SOPOT_B200_FP_STRICT;
        const bool tables = P.tab.cd.rows || P.tab.cl.rows || P.tab.cs.rows || P.tab.cmq.rows;
        double* ds = (double*)d_state; double* dt_ = (double*)d_traj; int32_t* dst = (int32_t*)d_status;
        if (tables) rc = strict ? rocket_launch_rk4<RocketSysTables, true>(ctx, P, n, dt, n_steps, opts->store_every, smem, ds, dt_, dst)
                                : rocket_launch_rk4<RocketSysTables, false>(ctx, P, n, dt, n_steps, opts->store_every, smem, ds, dt_, dst);
        else rc = strict ? rocket_launch_rk4<RocketSys, true>(ctx, P, n, dt, n_steps, opts->store_every, smem, ds, dt_, dst)
                         : rocket_launch_rk4<RocketSys, false>(ctx, P, n, dt, n_steps, opts->store_every, smem, ds, dt_, dst);
        if (rc) return rc;
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

SB_EXPORT int sopot_b200_rocket_rhs(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n,
                                    const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 4 * sb_align(n * 8) + 2 * sb_align(14 * n * 8) + 4096))) return rc;
    RocketSys::DevParams P;
    if ((rc = rocket_prepare(ctx, p, n, opts->broadcast, sg, P))) return rc;
    const void* d_state; void* d_out;
    if ((rc = sg.in(state, 14 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(out, 14 * n * 8, &d_out))) return rc;
    const size_t smem = RocketSys::smem_bytes(P);
    if (n) {
        const bool strict = opts->fp_mode == SOPOT_B200_FP_STRICT;
        const bool tables = P.tab.cd.rows || P.tab.cl.rows || P.tab.cs.rows || P.tab.cmq.rows;
        if (tables) rc = strict ? rocket_launch_rhs<RocketSysTables, true>(ctx, P, n, smem, (const double*)d_state, (double*)d_out)
                                : rocket_launch_rhs<RocketSysTables, false>(ctx, P, n, smem, (const double*)d_state, (double*)d_out);
        else rc = strict ? rocket_launch_rhs<RocketSys, true>(ctx, P, n, smem, (const double*)d_state, (double*)d_out)
                         : rocket_launch_rhs<RocketSys, false>(ctx, P, n, smem, (const double*)d_state, (double*)d_out);
        if (rc) return rc;
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}


// State functions of the rocket at given states (strict arithmetic): the values behind Rocket::queryStateFunction<Tag> and
// SimulationResult::query<Tag>(t) (rocket/rocket.hpp:222-228, rocket/simulation_result.hpp:111-115) for every tag that is an
// intermediate of the derivative; out[SOPOT_B200_ROCKET_PROBE_VALUES][n], rows documented in include/sopot_b200.h.
SB_EXPORT int sopot_b200_rocket_state_functions(sopot_b200_ctx* ctx, const sopot_b200_rocket_params* p, size_t n,
                                                const sopot_b200_rk4_opts* opts, const double* state, double* out) {
    if (!ctx) return SOPOT_B200_EINVAL;
    if (!p || !opts || !state || !out) return sb_fail(ctx, SOPOT_B200_EINVAL, "NULL argument");
    if (opts->mem > 1 || opts->fp_mode > 1) return sb_fail(ctx, SOPOT_B200_EINVAL, "bad opts: mem and fp_mode must be 0 or 1");
    static_assert(kProbeValues == SOPOT_B200_ROCKET_PROBE_VALUES, "header and kernel disagree on the record size");
    Staging sg(ctx, opts->mem);
    int rc;
    if (sg.host && (rc = sb_arena_begin(ctx, 4 * sb_align(n * 8) + sb_align(14 * n * 8) + sb_align(kProbeValues * n * 8) + 4096))) return rc;
    RocketSys::DevParams P;
    if ((rc = rocket_prepare(ctx, p, n, opts->broadcast, sg, P))) return rc;
    const void* d_state; void* d_out;
    if ((rc = sg.in(state, 14 * n * 8, &d_state))) return rc;
    if ((rc = sg.out(out, (size_t)kProbeValues * n * 8, &d_out))) return rc;
    const size_t smem = RocketSys::smem_bytes(P);
    if (n) {
        const bool tables = P.tab.cd.rows || P.tab.cl.rows || P.tab.cs.rows || P.tab.cmq.rows;
        if (tables) {
            if (smem > 48 * 1024) SB_CUDA(ctx, cudaFuncSetAttribute(rocket_probe_kernel<RocketSysTables>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rocket_probe_kernel<RocketSysTables><<<sb_grid_for(n, RocketSysTables::BLOCK), RocketSysTables::BLOCK, smem, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        } else {
            if (smem > 48 * 1024) SB_CUDA(ctx, cudaFuncSetAttribute(rocket_probe_kernel<RocketSys>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            rocket_probe_kernel<RocketSys><<<sb_grid_for(n, RocketSys::BLOCK), RocketSys::BLOCK, smem, ctx->stream>>>(P, n, (const double*)d_state, (double*)d_out);
        }
        SB_CHECK_LAUNCH(ctx);
    }
    return sg.finish();
}

#ifdef SB_COUNT_OPS
// variant builds only (tools/rocket_opcount.py): read (and optionally clear) the operation tallies of thread 0 / block 0
SB_EXPORT int sopot_b200_debug_opcounts(sopot_b200_ctx* ctx, unsigned long long* out8, int reset) {
    if (!ctx || !out8) return SOPOT_B200_EINVAL;
    SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SB_CUDA(ctx, cudaMemcpyFromSymbol(out8, g_sb_ops, 8 * sizeof(unsigned long long)));
    if (reset) { unsigned long long z[8] = {}; SB_CUDA(ctx, cudaMemcpyToSymbol(g_sb_ops, z, sizeof z)); }
    return SOPOT_B200_OK;
}
#endif
