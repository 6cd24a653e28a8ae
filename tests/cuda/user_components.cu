// user_components.cu -- the plugin path end to end: components written against the reference's TypedComponent contract
// (this is the reference's own composition example, physics/coupled_oscillator/: two point masses, a spring that
// provides both forces through the registry, an energy monitor that only provides state functions), compiled with nvcc
// into an RK4 ensemble kernel by sopot/b200/generic_ensemble.cuh.  Checks:
//   1. the reference's call shapes work unchanged (makeTypedODESystem, computeStateFunction<Tag>, computeDerivatives,
//      createRK4Solver().solve with the wrapper lambda) and run on the GPU;
//   2. an ensemble of jittered systems built by a device factory equals instance-by-instance single solves bit for bit;
//   3. -fmad=false builds reproduce the reference's CPU result of the same system (known answer below, from the
//      unmodified reference compiled by oracle/ref_harness.cpp) bit for bit; default builds to 1e-12.
// Build: nvcc -std=c++20 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a [-fmad=false] ... -lsopot_b200
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <sopot/core/solver.hpp>
#include <sopot/core/typed_component.hpp>
#include <sopot/physics/control/cart_n_linearize.cuh>
#include <sopot/physics/harmonic_oscillator.hpp>
#include <sopot/physics/pendulum/double_pendulum.hpp>

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "%s:%d CHECK FAILED: %s\n", __FILE__, __LINE__, #c); std::abort(); } } while (0)

namespace coupled {
using namespace sopot;
// tags as in physics/coupled_oscillator/tags.hpp
namespace mass1 { struct Position : StateFunction {}; struct Velocity : StateFunction {}; struct Force : StateFunction {}; struct Mass : StateFunction {};
                  struct TagSet { using Position = mass1::Position; using Velocity = mass1::Velocity; using Force = mass1::Force; using Mass = mass1::Mass; }; }
namespace mass2 { struct Position : StateFunction {}; struct Velocity : StateFunction {}; struct Force : StateFunction {}; struct Mass : StateFunction {};
                  struct TagSet { using Position = mass2::Position; using Velocity = mass2::Velocity; using Force = mass2::Force; using Mass = mass2::Mass; }; }
namespace spring { struct Extension : StateFunction {}; struct PotentialEnergy : StateFunction {}; }
namespace system { struct TotalEnergy : StateFunction {}; struct Momentum : StateFunction {}; struct CenterOfMass : StateFunction {}; }

// physics/coupled_oscillator/point_mass.hpp: state [x, v]; derivative {v, F/m} with F queried from the registry
template <typename TagSet, Scalar T = double>
class PointMass final : public TypedComponent<2, T> {
public:
    using Base = TypedComponent<2, T>;
    using typename Base::LocalState;
    using typename Base::LocalDerivative;
    SOPOT_HD PointMass(double mass, double x0 = 0.0, double v0 = 0.0) : m_mass(mass), m_x0(x0), m_v0(v0) {}
    SOPOT_HD LocalState getInitialLocalState() const { return {T(m_x0), T(m_v0)}; }
    template <typename Registry>
    SOPOT_HD LocalDerivative derivatives(T, std::span<const T> local, std::span<const T> global, const Registry& registry) const {
        const T force = registry.template computeFunction<typename TagSet::Force>(global);
        return {local[1], force / T(m_mass)};
    }
    SOPOT_HD T compute(typename TagSet::Position, std::span<const T> s) const { return this->getGlobalState(s, 0); }
    SOPOT_HD T compute(typename TagSet::Velocity, std::span<const T> s) const { return this->getGlobalState(s, 1); }
    SOPOT_HD T compute(typename TagSet::Mass, std::span<const T>) const { return T(m_mass); }
private:
    double m_mass, m_x0, m_v0;
};

// physics/coupled_oscillator/spring.hpp: no state; extension = x1 - x2 - L0, F1 = -k ext - c (v1 - v2), F2 = -F1
template <typename TS1, typename TS2, Scalar T = double>
class Spring final : public TypedComponent<0, T> {
public:
    SOPOT_HD Spring(double k, double L0 = 1.0, double c = 0.0) : m_k(k), m_L0(L0), m_c(c) {}
    typename TypedComponent<0, T>::LocalState getInitialLocalState() const { return {}; }
    template <typename Registry>
    SOPOT_HD T compute(spring::Extension, std::span<const T> s, const Registry& r) const {
        return r.template computeFunction<typename TS1::Position>(s) - r.template computeFunction<typename TS2::Position>(s) - T(m_L0);
    }
    template <typename Registry>
    SOPOT_HD T compute(typename TS1::Force, std::span<const T> s, const Registry& r) const {
        const T ext = compute(spring::Extension{}, s, r);
        const T rel = r.template computeFunction<typename TS1::Velocity>(s) - r.template computeFunction<typename TS2::Velocity>(s);
        T spring_force = -T(m_k) * ext;                                       // spring.hpp:105-112
        if (m_c > 0.0) spring_force += -T(m_c) * rel;
        return spring_force;
    }
    template <typename Registry>
    SOPOT_HD T compute(typename TS2::Force, std::span<const T> s, const Registry& r) const { return -compute(typename TS1::Force{}, s, r); }
    template <typename Registry>
    SOPOT_HD T compute(spring::PotentialEnergy, std::span<const T> s, const Registry& r) const {
        const T ext = compute(spring::Extension{}, s, r);
        return T(0.5 * m_k) * ext * ext;
    }
private:
    double m_k, m_L0, m_c;
};

// physics/coupled_oscillator/energy_monitor.hpp: state functions only
template <Scalar T = double>
class EnergyMonitor final : public TypedComponent<0, T> {
public:
    typename TypedComponent<0, T>::LocalState getInitialLocalState() const { return {}; }
    template <typename Registry>
    SOPOT_HD T compute(system::TotalEnergy, std::span<const T> s, const Registry& r) const {
        const T m1 = r.template computeFunction<mass1::Mass>(s), m2 = r.template computeFunction<mass2::Mass>(s);
        const T v1 = r.template computeFunction<mass1::Velocity>(s), v2 = r.template computeFunction<mass2::Velocity>(s);
        return T(0.5) * m1 * v1 * v1 + T(0.5) * m2 * v2 * v2 + r.template computeFunction<spring::PotentialEnergy>(s);
    }
    template <typename Registry>
    SOPOT_HD T compute(system::Momentum, std::span<const T> s, const Registry& r) const {
        return r.template computeFunction<mass1::Mass>(s) * r.template computeFunction<mass1::Velocity>(s) +
               r.template computeFunction<mass2::Mass>(s) * r.template computeFunction<mass2::Velocity>(s);
    }
};

inline auto make(double m1, double m2, double k, double L0, double c, double x1, double x2) {
    return makeTypedODESystem<double>(PointMass<mass1::TagSet>(m1, x1, 0.0), PointMass<mass2::TagSet>(m2, x2, 0.0),
                                      Spring<mass1::TagSet, mass2::TagSet>(k, L0, c), EnergyMonitor<double>());
}
using System = decltype(make(1, 1, 1, 1, 0, 0, 0));

// device factory: stiffness and damping jittered per instance
struct Factory {
    const double* k;
    const double* c;
    __device__ System operator()(size_t i) const {
        return makeTypedODESystem<double>(PointMass<mass1::TagSet>(1.0, 0.0, 0.0), PointMass<mass2::TagSet>(1.5, 2.0, 0.0),
                                          Spring<mass1::TagSet, mass2::TagSet>(k[i], 1.0, c[i]), EnergyMonitor<double>());
    }
};
}  // namespace coupled

int main() {
    using namespace sopot;
    using namespace coupled;
    // 1. the reference's call shapes on a user-defined pack
    auto sys = coupled::make(1.0, 1.5, 4.0, 1.0, 0.0, 0.0, 2.0);
    static_assert(decltype(sys)::hasFunction<coupled::system::TotalEnergy>() && decltype(sys)::hasFunction<mass2::Force>());
    static_assert(decltype(sys)::device_integrable && !decltype(sys)::Binding::available);
    CHECK(sys.getStateDimension() == 4 && sys.getComponentCount() == 4);
    auto y0 = sys.getInitialState();
    CHECK(y0[0] == 0.0 && y0[2] == 2.0);
    CHECK(sys.computeStateFunction<spring::Extension>(y0) == -3.0);
    CHECK(sys.computeStateFunction<mass1::Force>(y0) == 12.0 && sys.computeStateFunction<mass2::Force>(y0) == -12.0);
    const double e0 = sys.computeStateFunction<coupled::system::TotalEnergy>(y0);
    CHECK(e0 == 18.0);
    auto d = sys.computeDerivatives(0.0, y0);                                  // on the device
    CHECK(d[0] == 0.0 && d[1] == 12.0 && d[2] == 0.0 && d[3] == -12.0 / 1.5);
    auto derivs = [&sys](double t, StateView s) { return sys.computeDerivatives(t, std::vector<double>(s.begin(), s.end())); };
    auto r = createRK4Solver().solve(derivs, sys.getStateDimension(), 0.0, 5.0, 0.01, y0);
    CHECK(r.size() == 501);
    const double e1 = sys.computeStateFunction<coupled::system::TotalEnergy>(r.states.back());
    CHECK(std::abs(e1 - e0) / e0 < 1e-8);                                      // undamped: energy conserved
    CHECK(std::abs(sys.computeStateFunction<coupled::system::Momentum>(r.states.back())) < 1e-12);
    // known answer: the reference's createMass1/createMass2/createSpring system, RK4Solver 0..5 s, dt = 0.01 (CPU)
    const double ref[4] = {0.1051977204129609, 1.5655568262665147, 1.929868186391363, -1.043704550844343};
    for (int j = 0; j < 4; ++j) {
#ifdef SOPOT_TEST_STRICT
        CHECK(r.states.back()[j] == ref[j]);
#else
        CHECK(std::abs(r.states.back()[j] - ref[j]) <= 1e-12 * (1.0 + std::abs(ref[j])));
#endif
    }

    // 2. ensemble built by a device factory == single solves
    const size_t n = 4096;
    std::vector<double> k(n), c(n);
    for (size_t i = 0; i < n; ++i) { k[i] = 4.0 * (1.0 + 0.3 * std::sin(double(i))); c[i] = 0.05 * (1.0 + std::cos(double(i))); }
    double *dk, *dc;
    cudaMalloc(&dk, n * 8); cudaMalloc(&dc, n * 8);
    cudaMemcpy(dk, k.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dc, c.data(), n * 8, cudaMemcpyHostToDevice);
    b200::EnsembleOptions opt;
    opt.store_every = 100;
    auto ens = createRK4Solver().solveEnsemble(b200::ensembleOf(n, Factory{dk, dc}), 0.0, 5.0, 0.01, opt);
    CHECK(ens.snapshots() == 5 && ens.dim == 4 && ens.n == n);
    for (size_t i : {size_t(0), size_t(1234), n - 1}) {
        auto one = coupled::make(1.0, 1.5, k[i], 1.0, c[i], 0.0, 2.0);
        auto r1 = createRK4Solver().solve(one, 0.0, 5.0, 0.01);
        for (int j = 0; j < 4; ++j) {
            CHECK(ens.final(j, i) == r1.states.back()[j]);
            CHECK(ens.at(2, j, i) == r1.states[300][j]);
        }
        CHECK(ens.status[i] == 0);
    }
    cudaFree(dk); cudaFree(dc);

    // 3. shipped components inside a larger pack: HarmonicOscillator + DoublePendulum side by side (6 states) -- no
    //    tuned kernel exists for this pack, nvcc instantiates the generic one; each part evolves as it does alone
    {
        const double pi = 3.14159265358979323846;
        auto pack = makeTypedODESystem<double>(physics::createDampedOscillator<double>(2.0, 9.0, 0.1, 1.5),
                                               pendulum::DoublePendulum<double>(1.0, 1.0, 1.0, 1.0, 9.81, pi / 4, pi / 6, 0.0, 0.0));
        static_assert(!decltype(pack)::Binding::available && decltype(pack)::device_integrable);
        CHECK(pack.getStateDimension() == 6);
        auto rp = createRK4Solver().solve(pack, 0.0, 10.0, 0.01);
        CHECK(rp.size() == 1001);
        auto ho = makeTypedODESystem<double>(physics::createDampedOscillator<double>(2.0, 9.0, 0.1, 1.5));   // tuned kernel, strict
        auto rh = createRK4Solver().solve(ho, 0.0, 10.0, 0.01);
        for (int j = 0; j < 2; ++j) {
#ifdef SOPOT_TEST_STRICT
            CHECK(rp.states.back()[j] == rh.states.back()[j]);             // == the reference's -0.8212904505639914, -1.741913057783819
#else
            CHECK(std::abs(rp.states.back()[j] - rh.states.back()[j]) <= 1e-12 * 2.0);
#endif
        }
        auto dp = makeTypedODESystem<double>(pendulum::DoublePendulum<double>(1.0, 1.0, 1.0, 1.0, 9.81, pi / 4, pi / 6, 0.0, 0.0));
        auto rd = createRK4Solver().solve(dp, 0.0, 10.0, 0.01);
        for (int j = 0; j < 4; ++j) CHECK(std::abs(rp.states.back()[2 + j] - rd.states.back()[j]) <= 1e-9 * 10.0);
        CHECK(pack.computeStateFunction<energy::Total>(rp.states.back()) > 0.0);          // oscillator's tag, found in the pack
        CHECK(std::isfinite(pack.computeStateFunction<pendulum::system::TotalEnergy>(rp.states.back())));
    }

    // 4. N-link cart pendulum (the reference's tests use N = 2 and N = 6): generic kernel, constant force; known answers
    //    from the unmodified reference (oracle/ref_harness.cpp, 0..0.5 s, dt = 1e-3)
    {
        control::CartNPendulum<2, double> p2(1.0, {0.1, 0.2}, {0.5, 0.3}, 9.81, {0.0, -0.03, -0.06, 0.0, 0.0, 0.0});
        p2.setControlForce(0.4);
        auto s2 = makeTypedODESystem<double>(control::CartNPendulum<2, double>(p2));    // (lvalues are rejected, as in the reference)
        static_assert(!decltype(s2)::Binding::available && decltype(s2)::device_integrable);
        auto r2 = createRK4Solver().solve(s2, 0.0, 0.5, 0.001);
        const double ref2[6] = {0.05772152317547522, 0.1607029210223462, -1.260629618377168, 0.21007074548257548, 0.37351030122614226, -7.011211112024978};
        for (int j = 0; j < 6; ++j) CHECK(std::abs(r2.states.back()[j] - ref2[j]) <= 1e-10 * 8.0);
        control::CartNPendulum<6, double> p6(1.0, {0.1, 0.12000000000000001, 0.14, 0.16, 0.18, 0.2},
                                             {0.5, 0.46, 0.42, 0.38, 0.33999999999999997, 0.3}, 9.81,
                                             {0.0, -0.03, -0.06, -0.09, -0.12, -0.15, -0.18, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0});
        p6.setControlForce(0.4);
        auto s6 = makeTypedODESystem<double>(control::CartNPendulum<6, double>(p6));
        auto r6 = createRK4Solver().solve(s6, 0.0, 0.5, 0.001);
        const double ref6[14] = {0.04636578083826137, 0.8964738324552634, -1.288132050948551, -0.15209224959668097, -0.26008585952299235,
                                 -0.2880802460174711, -0.327630208430215, 0.1396841594166301, 3.034633113175984, -4.439563119100859,
                                 -1.2399832070583412, -0.25932876407463595, -0.33246554222609737, -0.3465573887257405};
        for (int j = 0; j < 14; ++j) CHECK(std::abs(r6.states.back()[j] - ref6[j]) <= 1e-9 * 5.0);
        // energy bookkeeping of the plant's own state queries on the device-integrated trajectory
        CHECK(std::isfinite(p6.totalEnergy(std::span<const double>(r6.states.back()))));
    }
    // 5. the reference's 6-link inverted-pendulum test on the GPU (tests/inverted_pendulum_test.cpp:208-292, configuration
    //    mc 2, m 0.1, L 0.3, Q = diag(1, 200 x6, 1, 10 x6), R 1e-3, Riccati dt 5e-3): lane-parallel Dual linearization at
    //    upright -> LQR<14,1> kernel -> closed loop of an ensemble through the generic kernel with a state-feedback controller
    {
        constexpr size_t NL = 6, NS = 14;
        const size_t n = 256;
        auto& eng = b200::Engine::shared();
        auto J = control::linearizeCartNPendulumBatch<NL>(eng, {2.0}, std::vector<double>(NL, 0.1), std::vector<double>(NL, 0.3), {9.81},
                                                         std::vector<double>(NS, 0.0), {0.0});
        double Q[NS * NS] = {};
        for (size_t i2 = 0; i2 < NS; ++i2) Q[i2 * NS + i2] = (i2 == 0 || i2 == NL + 1) ? 1.0 : (i2 <= NL ? 200.0 : 10.0);
        double K[NS];
        int32_t lqr_status = -1, lqr_iters = 0;
        b200::EnsembleOptions so; so.fp_mode = b200::FpMode::Strict;
        const auto lopts = b200::make_opts(so, 0);
        eng.check(sopot_b200_lqr_gain(eng.ctx(), NS, J.A.data(), J.B.data(), 1, Q, 0.001, 0.005, 1000, 1e-8, &lopts, K, nullptr, &lqr_iters, &lqr_status));
        eng.sync();
        CHECK(lqr_status == 0 && lqr_iters == 1000);
        // the reference's gain for this plant (LQR<14,1> on its analytic linearization; ours comes from the Dual Jacobians)
        const double Kref[NS] = {40.14077817166795, -24057.84045932211, 336162.24090843205, -1570953.7183741203, 3242784.7159446855,
                                 -3059859.116319377, 1078813.4568454213, 158.3014983927829, -215.6786313378815, 7627.8837693362075,
                                 -57888.385091549906, 169862.79596712242, -212059.4876367322, 95135.11617195724};
        for (size_t j = 0; j < NS; ++j) CHECK(std::abs(K[j] - Kref[j]) <= 1e-6 * std::abs(Kref[j]));
        // ensemble: every plant starts with all links at 0.03 rad (the reference test); link masses jittered by +-0.2 %
        // (the nominal gain has entries of 3e6: the six-link balance tolerates very little plant mismatch)
        std::vector<double> mc(n, 2.0), g(n, 9.81), m(NL * n), L(NL * n, 0.3), Kall(NS * n), y0(NS * n, 0.0);
        for (size_t i2 = 0; i2 < n; ++i2) {
            for (size_t a = 0; a < NL; ++a) m[a * n + i2] = 0.1 * (i2 == 0 ? 1.0 : 1.0 + 0.002 * std::sin(double(i2 + a)));
            for (size_t j = 0; j < NS; ++j) Kall[j * n + i2] = K[j];
            for (size_t a = 1; a <= NL; ++a) y0[a * n + i2] = 0.03;
        }
        using b200::generic::DeviceBuffer;
        DeviceBuffer<double> d_mc(n), d_g(n), d_m(NL * n), d_L(NL * n), d_K(NS * n);
        auto up = [](DeviceBuffer<double>& d, const std::vector<double>& h) { cudaMemcpy(d.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice); };
        up(d_mc, mc); up(d_g, g); up(d_m, m); up(d_L, L); up(d_K, Kall);
        auto loop = b200::ensembleOf(n, control::CartNFactory<NL>{d_mc.p, d_m.p, d_L.p, d_g.p, n})
                        .withInitialStates(y0)
                        .withController(control::CartNStateFeedback<NL>{d_K.p, nullptr, n, -500.0, 500.0, 1});
        b200::EnsembleOptions opt5;
        opt5.store_every = 50;
        auto r5 = createRK4Solver().solveEnsemble(loop, 0.0, 5.0, 0.001, opt5);
        for (size_t i2 = 0; i2 < n; ++i2) {
            CHECK(r5.status[i2] == 0);
            double max_angle = 0.03, final_angle = 0.0;
            for (size_t sidx = 0; sidx < r5.snapshots(); ++sidx)
                for (size_t a = 1; a <= NL; ++a) max_angle = std::max(max_angle, std::abs(r5.at(sidx, a, i2)));
            for (size_t a = 1; a <= NL; ++a) final_angle = std::max(final_angle, std::abs(r5.final(a, i2)));
            CHECK(max_angle < 0.2);          // never falls        (tests/inverted_pendulum_test.cpp:288)
            CHECK(final_angle < 0.01);       // converges back     (:289)
        }
        // plant 0 is the reference's nominal plant: state after 1.5 s of closed loop (snapshot 29) vs the reference stepping
        // (oracle/ref_harness.cpp: StateFeedbackController + CartNPendulum<6> + one RK4Solver step per sample, reference gain)
        const double ref15[NS] = {0.3469131415980191, -0.02512949479775311, -0.02363770986465978, -0.02219834441608054,
                                  -0.02078362623558745, -0.019390286965434777, -0.018027281622197824, 0.2430708885492833,
                                  -0.02852029197834663, -0.024081851393495852, -0.022182246865190935, -0.02182091176401473,
                                  -0.02227360805387352, -0.02304653899042201};
        for (size_t j = 0; j < NS; ++j) CHECK(std::abs(r5.at(29, j, 0) - ref15[j]) <= 1e-5);    // gains agree to 1e-6 relative
    }
    // 6. the same user-defined pack instantiated on Dual<double,4>: computeDerivatives in dual arithmetic and computeJacobian
    //    (core/typed_component.hpp:486-512; the reference's autodiff tests, tests/autodiff_test.cpp:387-427) run the components'
    //    own code in the forward-mode kernel.  d/dt [x1, v1, x2, v2] = [v1, F/m1, v2, -F/m2], F = -k (x1 - x2 - L0) - c (v1 - v2)
    {
        using D4 = Dual<double, 4>;
        const double m1 = 1.2, m2 = 0.7, k = 4.0, L0 = 1.0, c = 0.3;
        auto dsys = makeTypedODESystem<D4>(coupled::PointMass<coupled::mass1::TagSet, D4>(m1, 0.2, -0.1),
                                           coupled::PointMass<coupled::mass2::TagSet, D4>(m2, 1.9, 0.4),
                                           coupled::Spring<coupled::mass1::TagSet, coupled::mass2::TagSet, D4>(k, L0, c), coupled::EnergyMonitor<D4>());
        const std::array<double, 4> x{0.2, -0.1, 1.9, 0.4};
        const auto J = computeJacobian(dsys, 0.0, x);
        const double expect[4][4] = {{0, 1, 0, 0}, {-k / m1, -c / m1, k / m1, c / m1}, {0, 0, 0, 1}, {k / m2, c / m2, -k / m2, -c / m2}};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) CHECK(std::abs(J[i][j] - expect[i][j]) <= 1e-14);
        // seeds of the caller's choice travel through as well: directional derivative along (1, 0, 1, 0) (a rigid shift) is 0
        std::vector<D4> st(4);
        for (size_t i = 0; i < 4; ++i) st[i] = D4(x[i], {i % 2 == 0 ? 1.0 : 0.0, 0.0, 0.0, 0.0});
        const auto d = dsys.computeDerivatives(D4::constant(0.0), st);
        const double F = -k * (x[0] - x[2] - L0) - c * (x[1] - x[3]);
        CHECK(std::abs(d[1].value() - F / m1) <= 1e-14 && std::abs(d[3].value() + F / m2) <= 1e-14);
        for (size_t i = 0; i < 4; ++i) CHECK(std::abs(d[i].derivative(0)) <= 1e-14);
        // state functions of a Dual system stay host-side pure functions: dE/dv1 = m1 v1
        std::vector<D4> sv(4);
        for (size_t i = 0; i < 4; ++i) sv[i] = D4::variable(x[i], i);
        const D4 E = dsys.computeStateFunction<coupled::system::TotalEnergy>(sv);
        CHECK(std::abs(E.derivative(1) - m1 * x[1]) <= 1e-14 && std::abs(E.derivative(0) - k * (x[0] - x[2] - L0)) <= 1e-14);
    }
    std::puts("user_components: OK");
    return 0;
}
