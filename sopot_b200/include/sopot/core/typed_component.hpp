// sopot/core/typed_component.hpp -- composition API of the reference (core/typed_component.hpp) on top of
// the B200 engine: TypedComponent<N,T>, TypedRegistry, TypedODESystem, makeTypedODESystem, query<Tag>.
//
// What is the same: component base (state_size, LocalState, offsets, getGlobalState), compile-time
// registry dispatch of compute(Tag, state[, registry]) with "first provider in pack order wins, registry
// aware overload preferred" (reference :208-222, :249-261), constexpr state offsets (:306-325),
// getInitialState gathering (:377-398), hasFunction / computeStateFunction(s) / getComponent.
//
// What is different, on purpose: nothing integrates on the host.  computeDerivatives() and RK4Solver::solve()
// always launch a CUDA kernel, found in one of two ways:
//   1. b200::SystemBinding<Components...> -- a compile-time map from a shipped component pack to the hand-tuned
//      kernels of libsopot_b200.so behind the C ABI (oscillator, double pendulum, cart-pole, ...); these components
//      carry no derivative code at all.
//   2. any other pack whose stateful components implement the reference's
//          derivatives(T t, std::span<const T> local, std::span<const T> global, const Registry&) const
//      as SOPOT_HD functions: when the program is compiled with nvcc, b200/generic_ensemble.cuh instantiates the
//      RK4 ensemble kernel for that pack in the user's translation unit (state and registry queries resolve to
//      registers at compile time).  This is the plugin path: user-defined components run on the GPU.
// A pack with neither fails to compile with a message instead of silently running on the host (there is no CPU
// fallback).  State functions (positions, energies, ...) are pure functions of the state vector
// (docs/COMPONENT_GUIDELINES.md:396-407); they are SOPOT_HD too and serve both as host diagnostics and inside kernels.
#pragma once
#include <array>
#include <concepts>
#include <span>
#include <stdexcept>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../b200/engine.hpp"
#include "../b200/trace.hpp"
#include "macros.hpp"
#include "scalar.hpp"
#include "state_function_tags.hpp"

namespace sopot {

template <size_t N, Scalar T> class TypedComponent;

template <typename C>
concept TypedComponentConcept = requires {
    typename C::scalar_type;
    typename C::LocalState;
    typename C::LocalDerivative;
    { C::state_size } -> std::convertible_to<size_t>;
};
template <typename C>
concept HasInitialState = requires(const C& c) { { c.getInitialLocalState() } -> std::same_as<typename C::LocalState>; };
template <typename C>
concept HasIdentification = requires(const C& c) {
    { c.getComponentType() } -> std::convertible_to<std::string_view>;
    { c.getComponentName() } -> std::convertible_to<std::string_view>;
};
template <typename C, typename Tag, typename T>
concept TypedProvidesStateFunctionSpan = TypedComponentConcept<C> && StateTagConcept<Tag> &&
    requires(const C& c, std::span<const T> s) { { c.compute(Tag{}, s) }; };
template <typename C, typename Tag, typename T, typename Registry>
concept TypedProvidesRegistryAwareStateFunctionSpan = TypedComponentConcept<C> && StateTagConcept<Tag> &&
    requires(const C& c, std::span<const T> s, const Registry& r) { { c.compute(Tag{}, s, r) }; };
template <typename C, typename Tag, typename T, typename Registry>
concept TypedProvidesStateFunction =
    TypedProvidesStateFunctionSpan<C, Tag, T> || TypedProvidesRegistryAwareStateFunctionSpan<C, Tag, T, Registry>;

template <typename Registry, typename Tag>
concept RegistryProvides = requires { { Registry::template hasFunction<Tag>() } -> std::convertible_to<bool>; } &&
                           Registry::template hasFunction<Tag>();

template <StateTagConcept Tag, typename Registry, typename T>
    requires RegistryProvides<Registry, Tag>
inline auto query(const Registry& registry, std::span<const T> state) { return registry.template computeFunction<Tag>(state); }
template <StateTagConcept Tag, typename Registry, typename T>
    requires RegistryProvides<Registry, Tag>
inline auto query(const Registry& registry, const std::vector<T>& state) { return registry.template computeFunction<Tag>(state); }

template <size_t StateSize, Scalar T = double>
class TypedComponent {
public:
    using scalar_type = T;
    static constexpr size_t state_size = StateSize;
    using LocalState = ScalarState<T, StateSize>;
    using LocalDerivative = ScalarState<T, StateSize>;
    SOPOT_HD size_t getStateOffset() const noexcept { return m_state_offset; }
    SOPOT_HD void setStateOffset(size_t o) noexcept { m_state_offset = o; }
    SOPOT_HD void setOffset(size_t o) noexcept { m_state_offset = o; }
protected:
    size_t m_state_offset{0};
    SOPOT_HD T getGlobalState(std::span<const T> g, size_t i) const { return g[m_state_offset + i]; }
    SOPOT_HD LocalState extractLocalState(std::span<const T> g) const {
        LocalState l;
        for (size_t i = 0; i < StateSize; ++i) l[i] = g[m_state_offset + i];
        return l;
    }
};

// a stateful component that carries the reference's derivative method (device-compilable plugin path)
template <typename C, typename T, typename Registry>
concept HasDerivatives = requires(const C& c, T t, std::span<const T> s, const Registry& r) {
    { c.derivatives(t, s, s, r) } -> std::same_as<typename C::LocalDerivative>;
};

template <typename T, TypedComponentConcept... Components>
class TypedRegistry {
    std::tuple<const Components&...> m_components;
    using Self = TypedRegistry<T, Components...>;
    template <StateTagConcept Tag>
    static constexpr size_t providerIndex() {       // first component in pack order that provides Tag
        size_t idx = sizeof...(Components), i = 0;
        ((idx == sizeof...(Components) && TypedProvidesStateFunction<Components, Tag, T, Self> ? (idx = i) : 0, ++i), ...);
        return idx;
    }
public:
    SOPOT_HD explicit constexpr TypedRegistry(const Components&... c) : m_components(c...) {}
    template <StateTagConcept Tag>
    static constexpr bool hasFunction() { return (TypedProvidesStateFunction<Components, Tag, T, Self> || ...); }
    template <StateTagConcept Tag>
    SOPOT_HD auto computeFunction(std::span<const T> state) const {
        static_assert(hasFunction<Tag>(), "No component provides this state function");
        const auto& provider = std::get<providerIndex<Tag>()>(m_components);
        using P = std::decay_t<decltype(provider)>;
        if constexpr (TypedProvidesRegistryAwareStateFunctionSpan<P, Tag, T, Self>) return provider.compute(Tag{}, state, *this);
        else return provider.compute(Tag{}, state);
    }
    template <StateTagConcept Tag>
    auto computeFunction(const std::vector<T>& state) const { return computeFunction<Tag>(std::span<const T>(state)); }
    static constexpr size_t component_count() { return sizeof...(Components); }
    template <size_t I> SOPOT_HD constexpr const auto& getComponent() const { return std::get<I>(m_components); }
};

namespace b200 {
// Compile-time map: component pack -> device entry points.  Specialised next to the shipped components:
//   static constexpr bool available = true;  static constexpr size_t dim;
//   static std::vector<double> rhs(Engine&, const std::tuple<Cs...>&, const std::vector<double>& state);
//   static EnsembleResult solve(Engine&, const std::tuple<Cs...>&, const std::vector<double>& y0, size_t n_steps,
//                               double t0, double dt, const EnsembleOptions&);
template <typename... Components> struct SystemBinding { static constexpr bool available = false; };
}  // namespace b200

namespace b200::generic {
// defined in b200/generic_ensemble.cuh (nvcc only): N = 1 launches of the kernel instantiated for `System`
template <typename System> std::vector<double> rhs_one(const System& sys, const std::vector<double>& state);
// the same for a pack instantiated on Dual<double,K>: one derivative evaluation in dual arithmetic on the device
template <typename System> std::vector<typename System::scalar_type> rhs_one_dual(const System& sys, const std::vector<typename System::scalar_type>& state);
template <typename System>
EnsembleResult solve_one(Engine& eng, const System& sys, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                         const EnsembleOptions& o);
}  // namespace b200::generic

template <typename T, TypedComponentConcept... Components>
class TypedODESystem {
    std::tuple<Components...> m_components;
    static constexpr size_t m_total = (Components::state_size + ... + 0);
    using RegistryType = TypedRegistry<T, Components...>;
    static constexpr auto offsets() {
        std::array<size_t, sizeof...(Components) + 1> o{};
        size_t acc = 0, i = 0;
        ((o[i++] = acc, acc += Components::state_size), ...);
        o[sizeof...(Components)] = acc;
        return o;
    }
    static constexpr auto offset_array = offsets();
    template <typename C> static constexpr bool stateless_or_has_derivatives() {
        if constexpr (C::state_size == 0) return true; else return HasDerivatives<C, T, RegistryType>;
    }
public:
    using scalar_type = T;
    using Binding = b200::SystemBinding<Components...>;
    // every stateful component carries the reference's derivatives(): the pack can be compiled into a kernel
    static constexpr bool device_integrable = std::is_same_v<T, double> && (stateless_or_has_derivatives<Components>() && ...) &&
                                              (std::is_trivially_copyable_v<Components> && ...);
    // ... or, instantiated on Dual<double,K>, into the forward-mode derivative kernel (computeJacobian, linearizers)
    static constexpr bool device_differentiable = is_dual_v<T> && (stateless_or_has_derivatives<Components>() && ...) &&
                                                  (std::is_trivially_copyable_v<Components> && ...);

    SOPOT_HD explicit TypedODESystem(Components... c) : m_components(std::move(c)...) { initializeOffsets(); }
    // reference :306-325.  Public here: a kernel calls it on its local copy of a system that arrived through kernel
    // parameters, which turns every state offset back into a compile-time constant (register-resident state).
    SOPOT_HD void initializeOffsets() {
        [this]<size_t... I>(std::index_sequence<I...>) {
            (std::get<I>(m_components).setStateOffset(offset_array[I]), ...);
            (std::get<I>(m_components).setOffset(offset_array[I]), ...);
        }(std::make_index_sequence<sizeof...(Components)>{});
    }
    // Unlike the reference (whose stored registry refers to the tuple members) the registry is
    // rebuilt from the components on demand, so a system is an ordinary value: it can be copied, returned from a
    // factory and passed to a kernel.
    SOPOT_HD RegistryType getRegistry() const {
        return [this]<size_t... I>(std::index_sequence<I...>) { return RegistryType(std::get<I>(m_components)...); }(
            std::make_index_sequence<sizeof...(Components)>{});
    }

    // collectDerivatives (reference :333-365) on raw arrays: what the generic kernel calls.  With the system in
    // registers / kernel parameters every offset is a compile-time constant after inlining.
    SOPOT_HD void computeDerivativesInto(T t, const T (&state)[m_total ? m_total : 1], T (&out)[m_total ? m_total : 1]) const {
        const std::span<const T> global(state, m_total);
        const RegistryType reg = getRegistry();
        [&]<size_t... I>(std::index_sequence<I...>) {
            ([&] {
                using C = std::tuple_element_t<I, std::tuple<Components...>>;
                if constexpr (C::state_size > 0) {
                    const auto d = std::get<I>(m_components).derivatives(t, global.subspan(offset_array[I], C::state_size), global, reg);
                    for (size_t j = 0; j < C::state_size; ++j) out[offset_array[I] + j] = d[j];
                }
            }(), ...);
        }(std::make_index_sequence<sizeof...(Components)>{});
    }
    SOPOT_HD void getInitialStateInto(T (&s)[m_total ? m_total : 1]) const {
        [&]<size_t... I>(std::index_sequence<I...>) {
            ([&] {
                using C = std::tuple_element_t<I, std::tuple<Components...>>;
                if constexpr (C::state_size > 0) {
                    const auto l = std::get<I>(m_components).getInitialLocalState();
                    for (size_t j = 0; j < C::state_size; ++j) s[offset_array[I] + j] = l[j];
                }
            }(), ...);
        }(std::make_index_sequence<sizeof...(Components)>{});
    }

    // TypedODESystem::computeDerivatives (reference :410-414): one derivative evaluation, on the device.
    //   shipped pack, T = double          -> rhs_ensemble_kernel of libsopot_b200.so, N = 1
    //   shipped pack, T = Dual<double,K>  -> batched linearization kernel + chain rule over the caller's seeds
    //   user pack (nvcc)                  -> the kernel instantiated for this pack by b200/generic_ensemble.cuh
    std::vector<T> computeDerivatives(T /*t*/, const std::vector<T>& state) const {
        static_assert(Binding::available || ((device_integrable || device_differentiable) && SOPOT_B200_DEVICE_COMPILER),
                      "this component pack has no B200 kernel: it is not one of the shipped systems, and either its "
                      "stateful components lack SOPOT_HD derivatives() / are not trivially copyable or the program is not "
                      "compiled with nvcc (sopot-b200 has no host fallback)");
        if (state.size() != m_total) throw std::invalid_argument("state has the wrong dimension");
        if constexpr (Binding::available) {
            if constexpr (std::is_same_v<T, double>) {
                std::vector<double> out = Binding::rhs(b200::Engine::shared(), m_components, state);
                if (auto* tr = b200::detail::active_trace()) {           // see b200/trace.hpp
                    tr->calls++;
                    tr->returned = out;
                    const std::tuple<Components...> snap = m_components;
                    tr->solve = [snap](b200::Engine& e, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                                       const b200::EnsembleOptions& o) { return Binding::solve(e, snap, y0, n_steps, t0, dt, o); };
                    tr->rhs = [snap](b200::Engine& e, const std::vector<double>& y) { return Binding::rhs(e, snap, y); };
                }
                return out;
            } else {
                return Binding::rhs_dual(b200::Engine::shared(), m_components, state);
            }
        } else if constexpr (is_dual_v<T>) {
            return b200::generic::rhs_one_dual(*this, state);
        } else {
            std::vector<double> out = b200::generic::rhs_one(*this, state);
            if (auto* tr = b200::detail::active_trace()) {
                tr->calls++;
                tr->returned = out;
                const TypedODESystem snap = *this;
                tr->solve = [snap](b200::Engine& e, const std::vector<double>& y0, size_t n_steps, double t0, double dt,
                                   const b200::EnsembleOptions& o) { return b200::generic::solve_one(e, snap, y0, n_steps, t0, dt, o); };
                tr->rhs = [snap](b200::Engine&, const std::vector<double>& y) { return b200::generic::rhs_one(snap, y); };
            }
            return out;
        }
    }
    SOPOT_HD static constexpr size_t getStateDimension() noexcept { return m_total; }
    std::vector<T> getInitialState() const {
        std::vector<T> s(m_total);
        [this, &s]<size_t... I>(std::index_sequence<I...>) {
            ([&] {
                using C = std::tuple_element_t<I, std::tuple<Components...>>;
                if constexpr (C::state_size > 0) {
                    const auto& c = std::get<I>(m_components);
                    const auto l = c.getInitialLocalState();
                    for (size_t j = 0; j < C::state_size; ++j) s[c.getStateOffset() + j] = l[j];
                }
            }(), ...);
        }(std::make_index_sequence<sizeof...(Components)>{});
        return s;
    }
    template <StateTagConcept Tag> static constexpr bool hasFunction() { return RegistryType::template hasFunction<Tag>(); }
    template <StateTagConcept Tag> auto computeStateFunction(const std::vector<T>& s) const { return getRegistry().template computeFunction<Tag>(std::span<const T>(s)); }
    template <StateTagConcept Tag> SOPOT_HD auto computeStateFunction(std::span<const T> s) const { return getRegistry().template computeFunction<Tag>(s); }
    template <StateTagConcept... Tags>
    auto computeStateFunctions(const std::vector<T>& s) const {
        static_assert(sizeof...(Tags) > 0, "Must specify at least one function");
        static_assert((hasFunction<Tags>() && ...), "All requested functions must be available");
        const RegistryType reg = getRegistry();
        return std::tuple{reg.template computeFunction<Tags>(std::span<const T>(s))...};
    }
    template <size_t I>
        requires(I < sizeof...(Components))
    SOPOT_HD constexpr const auto& getComponent() const { return std::get<I>(m_components); }
    static constexpr size_t getComponentCount() noexcept { return sizeof...(Components); }
    const std::tuple<Components...>& components() const { return m_components; }
    std::vector<double> stateValues(const std::vector<T>& s) const {
        std::vector<double> v(s.size());
        for (size_t i = 0; i < s.size(); ++i) v[i] = value_of(s[i]);
        return v;
    }
};

template <typename T = double, TypedComponentConcept... Components>
SOPOT_HD auto makeTypedODESystem(Components&&... c) {
    return TypedODESystem<T, std::decay_t<Components>...>(std::forward<Components>(c)...);
}

// computeJacobian (reference :486-512): d f_i / d x_j of a system instantiated on Dual<double, N>.  Seeds the N state
// variables and evaluates the system's derivatives once -- for a shipped pack that evaluation is the batched forward-mode
// kernel behind its SystemBinding (rhs_dual), for a user pack the generic device kernel; nothing differentiates on the host.
template <size_t N, TypedComponentConcept... Components>
std::array<std::array<double, N>, N> computeJacobian(const TypedODESystem<Dual<double, N>, Components...>& system, double t,
                                                     const std::array<double, N>& state_values) {
    static_assert(N == (Components::state_size + ...), "State size mismatch");
    std::vector<Dual<double, N>> state(N);
    for (size_t i = 0; i < N; ++i) state[i] = Dual<double, N>::variable(state_values[i], i);
    const auto derivs = system.computeDerivatives(Dual<double, N>::constant(t), state);
    std::array<std::array<double, N>, N> jacobian{};
    for (size_t i = 0; i < N; ++i)
        for (size_t j = 0; j < N; ++j) jacobian[i][j] = derivs[i].derivative(j);
    return jacobian;
}

template <typename S> struct is_typed_ode_system : std::false_type {};
template <typename T, typename... C> struct is_typed_ode_system<TypedODESystem<T, C...>> : std::true_type {};

}  // namespace sopot

#if SOPOT_B200_DEVICE_COMPILER
#include "../b200/generic_ensemble.cuh"
#endif
