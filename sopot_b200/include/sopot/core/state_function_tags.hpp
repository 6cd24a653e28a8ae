// State-function tags: same names, categories and ids as the reference (core/state_function_tags.hpp:9-91),
// so user code that names sopot::kinematics::Position etc. compiles unchanged.
#pragma once
#include <concepts>
#include <cstddef>
#include <string_view>
#include <type_traits>

namespace sopot {

struct StateFunction {
    static constexpr std::string_view category() { return "base"; }
    static constexpr std::string_view name() { return "unknown"; }
    static constexpr size_t type_id() { return 0; }
};

#define SOPOT_B200_CATEGORY(NAME, TEXT)                                               \
    struct NAME : StateFunction { static constexpr std::string_view category() { return TEXT; } };
#define SOPOT_B200_TAG(NAME, BASE, TEXT, ID)                                          \
    struct NAME : BASE {                                                              \
        static constexpr std::string_view name() { return TEXT; }                     \
        static constexpr size_t type_id() { return ID; }                              \
    };

namespace categories {
SOPOT_B200_CATEGORY(Kinematics, "kinematics")
SOPOT_B200_CATEGORY(Dynamics, "dynamics")
SOPOT_B200_CATEGORY(Energy, "energy")
SOPOT_B200_CATEGORY(Analysis, "analysis")
}  // namespace categories
namespace kinematics {
SOPOT_B200_TAG(Position, categories::Kinematics, "position", 1)
SOPOT_B200_TAG(Velocity, categories::Kinematics, "velocity", 2)
SOPOT_B200_TAG(Acceleration, categories::Kinematics, "acceleration", 3)
}  // namespace kinematics
namespace energy {
SOPOT_B200_TAG(Kinetic, categories::Energy, "kinetic_energy", 101)
SOPOT_B200_TAG(Potential, categories::Energy, "potential_energy", 102)
SOPOT_B200_TAG(Total, categories::Energy, "total_energy", 103)
}  // namespace energy
namespace dynamics {
SOPOT_B200_TAG(Force, categories::Dynamics, "force", 201)
SOPOT_B200_TAG(Mass, categories::Dynamics, "mass", 202)
}  // namespace dynamics

template <typename T>
concept StateTagConcept = std::is_base_of_v<StateFunction, T> && requires {
    { T::name() } -> std::convertible_to<std::string_view>;
    { T::category() } -> std::convertible_to<std::string_view>;
    { T::type_id() } -> std::convertible_to<size_t>;
};

}  // namespace sopot
