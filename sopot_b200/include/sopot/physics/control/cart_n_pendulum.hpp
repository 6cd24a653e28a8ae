// control::CartNPendulum<NLinks, T> -- cart with an N-link inverted pendulum, state [x, th_1..th_N, xd, w_1..w_N].
// Interface of physics/control/cart_n_pendulum.hpp:48-316 (constructor validation, setControlForce, state
// queries, energies, tags).  The Euler-Lagrange derivative with its Gaussian elimination (:143-221, :319-359)
// is CartPlant<NL, T> in csrc/cartpole.cu; libsopot_b200.so ships the NLinks = 1 instantiations:
//   T = double           -> sopot_b200_cartpole_rk4 / _rhs       (K1, constant control force)
//   T = Dual<double, K>  -> sopot_b200_cartpole_linearize        (K2) + chain rule over the caller's seeds
#pragma once
#include <array>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include "../../core/typed_component.hpp"
#include "control_tags.hpp"

namespace sopot::control {

template <size_t NLinks, Scalar T = double>
class CartNPendulum final : public TypedComponent<2 * (NLinks + 1), T> {
public:
    static_assert(NLinks >= 1, "CartNPendulum requires at least one link");
    static constexpr size_t num_links = NLinks;
    static constexpr size_t num_dof = NLinks + 1;
    static constexpr size_t num_state = 2 * num_dof;
    using Base = TypedComponent<num_state, T>;
    using typename Base::LocalState;
    using typename Base::LocalDerivative;

    SOPOT_HD CartNPendulum(double cart_mass, const std::array<double, NLinks>& link_masses,
                           const std::array<double, NLinks>& link_lengths, double gravity = 9.81,
                           const std::array<double, num_state>& initial_state = {})
        : m_cart_mass(cart_mass), m_mass(link_masses), m_length(link_lengths), m_gravity(gravity),
          m_initial_state(initial_state) {
#ifndef __CUDA_ARCH__        // validation throws on the host (reference :99-110); device factories skip it
        if (cart_mass <= 0.0) throw std::invalid_argument("Cart mass must be positive");
        for (size_t i = 0; i < NLinks; ++i) {
            if (m_mass[i] <= 0.0) throw std::invalid_argument("All link masses must be positive");
            if (m_length[i] <= 0.0) throw std::invalid_argument("All link lengths must be positive");
        }
#endif
    }
    // control input: const + mutable exactly like the reference (:70, :117) -- not thread-safe there either
    SOPOT_HD void setControlForce(T force) const { m_control_force = force; }
    SOPOT_HD T getControlForce() const { return m_control_force; }

    SOPOT_HD LocalState getInitialLocalState() const {
        LocalState s;
        for (size_t i = 0; i < num_state; ++i) s[i] = T(m_initial_state[i]);
        return s;
    }
    // Euler-Lagrange derivative, any number of links (:143-221 with the Gaussian elimination of :319-359): trig cache,
    // tail masses, (N+1)x(N+1) mass matrix with cos(a-b) = ca cb + sa sb, partial pivoting on the value.  One link runs
    // on the tuned kernels through its SystemBinding; N >= 2 (the reference's own tests use N = 2 and N = 6) compiles
    // into a generic kernel under nvcc.  Every loop has constant bounds and the row exchange is a predicated element
    // swap, so after unrolling all indices are static (register-resident on the device).  Where the reference throws
    // "mass matrix is singular" the derivative becomes NaN (device code cannot throw): the ensemble status reports it.
    template <typename Registry>
    SOPOT_HD LocalDerivative derivatives(T, std::span<const T> local, std::span<const T>, const Registry&) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        constexpr size_t ND = num_dof;
        T s[NLinks], c[NLinks], omega[NLinks];
        double Mtail[NLinks];
        SOPOT_UNROLL for (size_t i = 0; i < NLinks; ++i) {
            const T theta = local[1 + i];
            s[i] = sin(theta); c[i] = cos(theta);
            omega[i] = local[ND + 1 + i];
            Mtail[i] = tailMass(i);
        }
        const T g = T(m_gravity);
        T A[ND][ND], b[ND];
        T total = T(m_cart_mass);
        SOPOT_UNROLL for (size_t i = 0; i < NLinks; ++i) total = total + T(m_mass[i]);
        A[0][0] = total;
        b[0] = m_control_force;
        SOPOT_UNROLL for (size_t a = 0; a < NLinks; ++a) {
            const T coupling = T(Mtail[a]) * T(m_length[a]) * c[a];
            A[0][1 + a] = coupling; A[1 + a][0] = coupling;
            b[0] = b[0] + T(Mtail[a]) * T(m_length[a]) * s[a] * omega[a] * omega[a];
        }
        SOPOT_UNROLL for (size_t a = 0; a < NLinks; ++a) {
            T rhs_a = T(Mtail[a]) * g * T(m_length[a]) * s[a];
            SOPOT_UNROLL for (size_t bb = 0; bb < NLinks; ++bb) {
                const size_t mx = a > bb ? a : bb;
                A[1 + a][1 + bb] = T(Mtail[mx]) * T(m_length[a]) * T(m_length[bb]) * (c[a] * c[bb] + s[a] * s[bb]);
                if (bb != a)
                    rhs_a = rhs_a - T(Mtail[mx]) * T(m_length[a]) * T(m_length[bb]) * (s[a] * c[bb] - c[a] * s[bb]) * omega[bb] * omega[bb];
            }
            b[1 + a] = rhs_a;
        }
        bool singular = false;
        SOPOT_UNROLL for (size_t col = 0; col < ND; ++col) {
            size_t pivot = col;
            double best = std::abs(value_of(A[col][col]));
            SOPOT_UNROLL for (size_t row = col + 1; row < ND; ++row) {
                const double mag = std::abs(value_of(A[row][col]));
                if (mag > best) { best = mag; pivot = row; }
            }
            singular = singular || best < 1e-15;
            SOPOT_UNROLL for (size_t row = col + 1; row < ND; ++row) {
                if (row == pivot) {
                    SOPOT_UNROLL for (size_t j = 0; j < ND; ++j) { const T tmp = A[col][j]; A[col][j] = A[row][j]; A[row][j] = tmp; }
                    const T tmp = b[col]; b[col] = b[row]; b[row] = tmp;
                }
            }
            SOPOT_UNROLL for (size_t row = col + 1; row < ND; ++row) {
                const T factor = A[row][col] / A[col][col];
                SOPOT_UNROLL for (size_t j = col; j < ND; ++j) A[row][j] = A[row][j] - factor * A[col][j];
                b[row] = b[row] - factor * b[col];
            }
        }
        T x[ND];
        SOPOT_UNROLL for (size_t i = ND; i-- > 0;) {
            T sum = b[i];
            SOPOT_UNROLL for (size_t j = i + 1; j < ND; ++j) sum = sum - A[i][j] * x[j];
            x[i] = sum / A[i][i];
        }
        constexpr double kNaN = std::numeric_limits<double>::quiet_NaN();
        LocalDerivative d;
        SOPOT_UNROLL for (size_t i = 0; i < ND; ++i) { d[i] = singular ? T(kNaN) : T(local[ND + i]); d[ND + i] = singular ? T(kNaN) : x[i]; }
        return d;
    }
    std::string_view getComponentType() const { return "CartNPendulum"; }
    std::string_view getComponentName() const { return "CartNPendulum"; }
    SOPOT_HD double tailMass(size_t link) const {
        double sum = 0.0;
        SOPOT_UNROLL for (size_t i = 0; i < NLinks; ++i) if (i >= link) sum += m_mass[i];    // (same order, constant bounds)
        return sum;
    }

    T cartPosition(std::span<const T> s) const { return this->getGlobalState(s, 0); }
    T cartVelocity(std::span<const T> s) const { return this->getGlobalState(s, num_dof); }
    T linkAngle(size_t link, std::span<const T> s) const { return this->getGlobalState(s, 1 + link); }
    T linkAngularVelocity(size_t link, std::span<const T> s) const { return this->getGlobalState(s, num_dof + 1 + link); }
    std::array<T, 2> linkTipPosition(size_t link, std::span<const T> s) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        T px = cartPosition(s), py = T(0.0);
        for (size_t i = 0; i <= link; ++i) {
            const T th = linkAngle(i, s);
            px = px + T(m_length[i]) * sin(th);
            py = py + T(m_length[i]) * cos(th);
        }
        return {px, py};
    }
    T kineticEnergy(std::span<const T> s) const {
        using std::sin; using std::cos; using sopot::sin; using sopot::cos;
        const T xdot = cartVelocity(s);
        T KE = T(0.5) * T(m_cart_mass) * xdot * xdot;
        T vx = xdot, vy = T(0.0);
        for (size_t i = 0; i < NLinks; ++i) {
            const T th = linkAngle(i, s), w = linkAngularVelocity(i, s);
            vx = vx + T(m_length[i]) * cos(th) * w;
            vy = vy - T(m_length[i]) * sin(th) * w;
            KE = KE + T(0.5) * T(m_mass[i]) * (vx * vx + vy * vy);
        }
        return KE;
    }
    T potentialEnergy(std::span<const T> s) const {
        using std::cos; using sopot::cos;
        T h = T(0.0), PE = T(0.0);
        for (size_t i = 0; i < NLinks; ++i) {
            h = h + T(m_length[i]) * cos(linkAngle(i, s));
            PE = PE + T(m_mass[i]) * T(m_gravity) * h;
        }
        return PE;
    }
    T totalEnergy(std::span<const T> s) const { return kineticEnergy(s) + potentialEnergy(s); }

    T compute(cart::Position, std::span<const T> s) const { return cartPosition(s); }
    T compute(cart::Velocity, std::span<const T> s) const { return cartVelocity(s); }
    T compute(cart::Mass, std::span<const T>) const { return T(m_cart_mass); }
    T compute(controller::ControlInput, std::span<const T>) const { return m_control_force; }

    double getCartMass() const { return m_cart_mass; }
    double getMass(size_t link) const { return m_mass[link]; }
    double getLength(size_t link) const { return m_length[link]; }
    const std::array<double, NLinks>& getMasses() const { return m_mass; }
    const std::array<double, NLinks>& getLengths() const { return m_length; }
    double getGravity() const { return m_gravity; }

private:
    double m_cart_mass;
    std::array<double, NLinks> m_mass, m_length;
    double m_gravity;
    std::array<double, num_state> m_initial_state;
    mutable T m_control_force = T(0.0);
};

// SoA parameter block of a cart-pole (N = 1) ensemble for RK4Solver::solveEnsemble
struct CartPoleEnsemble {
    size_t n;
    b200::Param cart_mass, link_mass, link_length, gravity, x, theta, x_dot, omega, force;
    sopot_b200_cartpole_params params(uint32_t& bc) const {
        const b200::Param* f[9] = {&cart_mass, &link_mass, &link_length, &gravity, &x, &theta, &x_dot, &omega, &force};
        bc = 0;
        for (int j = 0; j < 9; ++j) bc |= (f[j]->broadcast(n) ? 1u : 0u) << j;
        return {f[0]->data(), f[1]->data(), f[2]->data(), f[3]->data(), f[4]->data(), f[5]->data(), f[6]->data(),
                f[7]->data(), f[8]->data()};
    }
    b200::EnsembleResult integrate(b200::Engine& eng, size_t n_steps, double t0, double dt, const b200::EnsembleOptions& o) const {
        uint32_t bc;
        const auto p = params(bc);
        const auto opts = b200::make_opts(o, bc);
        b200::EnsembleResult r;
        r.n = n; r.dim = 4; r.n_steps = n_steps; r.store_every = o.store_every;
        r.state.resize(4 * n);
        r.trajectory.resize(r.snapshots() * 4 * n);
        if (o.want_status) r.status.resize(n);
        eng.check(sopot_b200_cartpole_rk4(eng.ctx(), &p, n, t0, dt, n_steps, &opts, r.state.data(),
                                          r.trajectory.empty() ? nullptr : r.trajectory.data(),
                                          o.want_status ? r.status.data() : nullptr));
        eng.sync();
        return r;
    }
};

// Batched SystemLinearizer<4,1>::linearize over P operating points (K2): A[i][j] of point p at
// A[(i*4+j)*P + p], B[i] at B[i*P + p].
struct CartPoleJacobians {
    size_t P = 0;
    std::vector<double> A, B;
    std::vector<int32_t> status;
    double a(size_t i, size_t j, size_t p) const { return A[(i * 4 + j) * P + p]; }
    double b(size_t i, size_t p) const { return B[i * P + p]; }
};
inline CartPoleJacobians linearizeCartPoleBatch(b200::Engine& eng, const b200::Param& cart_mass, const b200::Param& link_mass,
                                                const b200::Param& link_length, const b200::Param& gravity,
                                                const std::vector<double>& x0 /*[4][P]*/, const std::vector<double>& u0 /*[P]*/,
                                                b200::FpMode fp = b200::FpMode::Contract) {
    const size_t P = u0.size();
    if (x0.size() != 4 * P) throw std::invalid_argument("x0 must hold 4*P values");
    const b200::Param* f[4] = {&cart_mass, &link_mass, &link_length, &gravity};
    uint32_t bc = 0;
    for (int j = 0; j < 4; ++j) bc |= ((f[j]->v.size() == 1) ? 1u : 0u) << j;
    for (int j = 0; j < 4; ++j)
        if (f[j]->v.size() != 1 && f[j]->v.size() != P) throw std::invalid_argument("plant parameter size mismatch");
    b200::EnsembleOptions o; o.fp_mode = fp;
    const auto opts = b200::make_opts(o, bc);
    sopot_b200_cartpole_plant plant{f[0]->data(), f[1]->data(), f[2]->data(), f[3]->data()};
    CartPoleJacobians r;
    r.P = P; r.A.resize(16 * P); r.B.resize(4 * P); r.status.resize(P);
    eng.check(sopot_b200_cartpole_linearize(eng.ctx(), &plant, x0.data(), u0.data(), P, &opts, r.A.data(), r.B.data(),
                                            r.status.data()));
    eng.sync();
    return r;
}

}  // namespace sopot::control

namespace sopot::b200 {
namespace detail {
inline void throw_if_singular(int32_t st) {
    if (st == SOPOT_B200_ST_SINGULAR)     // message of physics/control/cart_n_pendulum.hpp:331-333
        throw std::runtime_error("CartNPendulum: mass matrix is singular; cannot solve for accelerations.");
}
}  // namespace detail

template <>
struct SystemBinding<control::CartNPendulum<1, double>> {
    static constexpr bool available = true;
    using Plant = control::CartNPendulum<1, double>;
    using Tuple = std::tuple<Plant>;
    static control::CartPoleEnsemble one(const Plant& c, const std::vector<double>& y) {
        return {1, c.getCartMass(), c.getMass(0), c.getLength(0), c.getGravity(), y[0], y[1], y[2], y[3], c.getControlForce()};
    }
    static std::vector<double> rhs(Engine& eng, const Tuple& t, const std::vector<double>& s) {
        const auto e = one(std::get<0>(t), s);
        uint32_t bc;
        const auto p = e.params(bc);
        EnsembleOptions o; o.fp_mode = FpMode::Strict;
        const auto opts = make_opts(o, 0);
        std::vector<double> out(4);
        eng.check(sopot_b200_cartpole_rhs(eng.ctx(), &p, 1, &opts, s.data(), out.data()));
        eng.sync();
        return out;
    }
    static EnsembleResult solve(Engine& eng, const Tuple& t, const std::vector<double>& y0, size_t n_steps, double t0,
                                double dt, const EnsembleOptions& o) {
        return one(std::get<0>(t), y0).integrate(eng, n_steps, t0, dt, o);
    }
};

// Dual scalar: value and Jacobian from the device (one K2 point), then d(out_i)/d(seed_k) =
// sum_j A[i][j] * dx_j/d(seed_k) + B[i] * dF/d(seed_k).  With the linearizer's identity seeds the products are
// x*1 and x+0: the partials returned are the kernel's own bits.
template <size_t K>
struct SystemBinding<control::CartNPendulum<1, Dual<double, K>>> {
    static constexpr bool available = true;
    using D = Dual<double, K>;
    using Plant = control::CartNPendulum<1, D>;
    using Tuple = std::tuple<Plant>;
    static std::vector<D> rhs_dual(Engine& eng, const Tuple& t, const std::vector<D>& s) {
        const Plant& c = std::get<0>(t);
        const double mc = c.getCartMass(), m = c.getMass(0), L = c.getLength(0), g = c.getGravity();
        const D F = c.getControlForce();
        const double x0[4] = {s[0].value(), s[1].value(), s[2].value(), s[3].value()};
        const double u0 = F.value();
        EnsembleOptions o; o.fp_mode = FpMode::Strict;
        const auto opts = make_opts(o, 0);
        double A[16], B[4], f[4];
        int32_t st = 0;
        sopot_b200_cartpole_plant plant{&mc, &m, &L, &g};
        eng.check(sopot_b200_cartpole_linearize(eng.ctx(), &plant, x0, &u0, 1, &opts, A, B, &st));
        sopot_b200_cartpole_params p{&mc, &m, &L, &g, nullptr, nullptr, nullptr, nullptr, &u0};
        eng.check(sopot_b200_cartpole_rhs(eng.ctx(), &p, 1, &opts, x0, f));
        eng.sync();
        detail::throw_if_singular(st);
        std::vector<D> out(4);
        for (size_t i = 0; i < 4; ++i) {
            std::array<double, K> d{};
            for (size_t k = 0; k < K; ++k) {
                double acc = 0.0;
                for (size_t j = 0; j < 4; ++j) acc += A[i * 4 + j] * s[j].derivative(k);
                d[k] = acc + B[i] * F.derivative(k);
            }
            out[i] = D(f[i], d);
        }
        return out;
    }
};
}  // namespace sopot::b200
