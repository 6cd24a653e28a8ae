// Double-pendulum state-function tags (names of physics/pendulum/tags.hpp).
#pragma once
#include "../../core/state_function_tags.hpp"

namespace sopot::pendulum {
namespace mass1 {
SOPOT_B200_TAG(Angle, categories::Kinematics, "mass1_angle", 1101)
SOPOT_B200_TAG(AngularVelocity, categories::Kinematics, "mass1_angular_velocity", 1102)
SOPOT_B200_TAG(CartesianPosition, categories::Kinematics, "mass1_position", 1103)
SOPOT_B200_TAG(CartesianVelocity, categories::Kinematics, "mass1_velocity", 1104)
SOPOT_B200_TAG(Mass, categories::Dynamics, "mass1_mass", 1105)
}  // namespace mass1
namespace mass2 {
SOPOT_B200_TAG(Angle, categories::Kinematics, "mass2_angle", 1201)
SOPOT_B200_TAG(AngularVelocity, categories::Kinematics, "mass2_angular_velocity", 1202)
SOPOT_B200_TAG(CartesianPosition, categories::Kinematics, "mass2_position", 1203)
SOPOT_B200_TAG(CartesianVelocity, categories::Kinematics, "mass2_velocity", 1204)
SOPOT_B200_TAG(Mass, categories::Dynamics, "mass2_mass", 1205)
}  // namespace mass2
namespace system {
SOPOT_B200_TAG(Length1, categories::Kinematics, "length1", 1301)
SOPOT_B200_TAG(Length2, categories::Kinematics, "length2", 1302)
SOPOT_B200_TAG(KineticEnergy, categories::Energy, "kinetic_energy", 1303)
SOPOT_B200_TAG(PotentialEnergy, categories::Energy, "potential_energy", 1304)
SOPOT_B200_TAG(TotalEnergy, categories::Energy, "total_energy", 1305)
}  // namespace system
}  // namespace sopot::pendulum
