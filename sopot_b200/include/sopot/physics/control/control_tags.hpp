// State-function tags of the control plants: names and namespaces of physics/control/control_tags.hpp:6-131.
#pragma once
#include <array>
#include "../../core/state_function_tags.hpp"

namespace sopot::control {
#define SOPOT_B200_CTAG(NAME, TEXT, RET)                                  \
    struct NAME : StateFunction {                                         \
        static constexpr auto description = TEXT;                         \
        using return_type = RET;                                          \
    };
namespace cart {
SOPOT_B200_CTAG(Position, "Cart position along track (m)", double)
SOPOT_B200_CTAG(Velocity, "Cart velocity (m/s)", double)
SOPOT_B200_CTAG(Mass, "Cart mass (kg)", double)
}  // namespace cart
namespace link1 {
SOPOT_B200_CTAG(Angle, "Angle of link 1 from vertical (rad), 0 = upright", double)
SOPOT_B200_CTAG(AngularVelocity, "Angular velocity of link 1 (rad/s)", double)
SOPOT_B200_CTAG(Mass, "Mass of link 1 (kg)", double)
SOPOT_B200_CTAG(Length, "Length of link 1 (m)", double)
using TipPositionValue = std::array<double, 2>;
SOPOT_B200_CTAG(TipPosition, "Cartesian position [x, y] of link 1 tip (m)", TipPositionValue)
}  // namespace link1
namespace link2 {
SOPOT_B200_CTAG(Angle, "Angle of link 2 from vertical (rad), 0 = upright", double)
SOPOT_B200_CTAG(AngularVelocity, "Angular velocity of link 2 (rad/s)", double)
SOPOT_B200_CTAG(Mass, "Mass of link 2 (kg)", double)
SOPOT_B200_CTAG(Length, "Length of link 2 (m)", double)
using TipPositionValue = std::array<double, 2>;
SOPOT_B200_CTAG(TipPosition, "Cartesian position [x, y] of link 2 tip (m)", TipPositionValue)
}  // namespace link2
namespace controller {
using Vector6 = std::array<double, 6>;
SOPOT_B200_CTAG(ControlInput, "Control force applied to cart (N)", double)
SOPOT_B200_CTAG(StateError, "Error from reference state", Vector6)
SOPOT_B200_CTAG(ReferenceState, "Reference/target state", Vector6)
}  // namespace controller
namespace system {
using Vector6 = std::array<double, 6>;
SOPOT_B200_CTAG(FullState, "Full state vector [x, th1, th2, xd, w1, w2]", Vector6)
SOPOT_B200_CTAG(TotalEnergy, "Total mechanical energy (J)", double)
SOPOT_B200_CTAG(KineticEnergy, "Total kinetic energy (J)", double)
SOPOT_B200_CTAG(PotentialEnergy, "Total potential energy (J)", double)
}  // namespace system
#undef SOPOT_B200_CTAG
}  // namespace sopot::control
