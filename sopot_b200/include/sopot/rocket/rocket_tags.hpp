// State-function tags of the 6-DOF rocket: names, namespaces and value types of rocket/rocket_tags.hpp:6-239, plus the two
// small value types they return (rocket/vector3.hpp, rocket/quaternion.hpp: plain aggregates here -- the vector algebra of
// the model runs inside the kernel).  Every tag the reference's component registry can answer for Rocket<T>::SystemType is
// listed with where its value comes from: a slice of the 14-state vector, or a row of the record that
// sopot_b200_rocket_state_functions evaluates on the device (include/sopot_b200.h).  Rocket::queryStateFunction<Tag> and
// SimulationResult::query<Tag>(t) are generic over exactly this list.
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include "../core/scalar.hpp"
#include "../core/state_function_tags.hpp"

namespace sopot::rocket {

template <Scalar T = double>
struct Vector3 {
    T x{}, y{}, z{};
    constexpr Vector3() = default;
    constexpr Vector3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    explicit Vector3(const std::array<T, 3>& a) : x(a[0]), y(a[1]), z(a[2]) {}
    T& operator[](size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
    T norm() const { using std::sqrt; return sqrt(x * x + y * y + z * z); }
    std::array<T, 3> toArray() const { return {x, y, z}; }
};
template <Scalar T = double>
struct Quaternion {
    T q1{}, q2{}, q3{}, q4{1};                 // scalar last (rocket/quaternion.hpp:13-14)
    constexpr Quaternion() = default;
    constexpr Quaternion(T a, T b, T c, T d) : q1(a), q2(b), q3(c), q4(d) {}
    T& operator[](size_t i) { return i == 0 ? q1 : (i == 1 ? q2 : (i == 2 ? q3 : q4)); }
    const T& operator[](size_t i) const { return i == 0 ? q1 : (i == 1 ? q2 : (i == 2 ? q3 : q4)); }
    T norm() const { using std::sqrt; return sqrt(q1 * q1 + q2 * q2 + q3 * q3 + q4 * q4); }
};

// where a tag's value lives: Slice = state[first .. first + count), Probe = rows of the device record
enum class TagSource { Slice, Probe };
enum class TagShape { Scalar, Vec3, Quat };

#define SOPOT_B200_RTAG(NAME, TEXT, SOURCE, SHAPE, FIRST)                     \
    struct NAME : StateFunction {                                             \
        static constexpr auto description = TEXT;                             \
        static constexpr TagSource source = TagSource::SOURCE;                \
        static constexpr TagShape shape = TagShape::SHAPE;                    \
        static constexpr size_t first = FIRST;                                \
    };
namespace kinematics {
SOPOT_B200_RTAG(PositionENU, "Position in East-North-Up frame (m)", Slice, Vec3, 1)
SOPOT_B200_RTAG(VelocityENU, "Velocity in East-North-Up frame (m/s)", Slice, Vec3, 4)
SOPOT_B200_RTAG(AttitudeQuaternion, "Body-to-ENU attitude quaternion, scalar last", Slice, Quat, 7)
SOPOT_B200_RTAG(AngularVelocity, "Angular velocity in body frame (rad/s)", Slice, Vec3, 11)
SOPOT_B200_RTAG(Altitude, "Altitude above the launch site (m)", Slice, Scalar, 3)
}  // namespace kinematics
namespace propulsion {
SOPOT_B200_RTAG(Time, "Simulation time (s)", Slice, Scalar, 0)
SOPOT_B200_RTAG(MassFlowRate, "Propellant mass flow rate (kg/s)", Probe, Scalar, 24)
SOPOT_B200_RTAG(ThrustForceBody, "Thrust force in body frame (N)", Probe, Vec3, 25)
}  // namespace propulsion
namespace dynamics {
SOPOT_B200_RTAG(Mass, "Vehicle mass (kg)", Probe, Scalar, 0)
SOPOT_B200_RTAG(MomentOfInertia, "Principal moments of inertia (kg m^2)", Probe, Vec3, 1)
SOPOT_B200_RTAG(TotalForceENU, "Total force in ENU frame (N)", Probe, Vec3, 4)
SOPOT_B200_RTAG(TotalTorqueBody, "Total torque in body frame (N m)", Probe, Vec3, 7)
SOPOT_B200_RTAG(GravityAcceleration, "Gravitational acceleration (m/s^2)", Probe, Scalar, 10)
}  // namespace dynamics
namespace aero {
SOPOT_B200_RTAG(AeroForceBody, "Aerodynamic force in body frame (N)", Probe, Vec3, 11)
SOPOT_B200_RTAG(AeroForceENU, "Aerodynamic force in ENU frame (N)", Probe, Vec3, 14)
SOPOT_B200_RTAG(AeroMomentBody, "Aerodynamic moment in body frame (N m)", Probe, Vec3, 17)
}  // namespace aero
namespace environment {
SOPOT_B200_RTAG(AtmosphericDensity, "Air density (kg/m^3)", Probe, Scalar, 20)
SOPOT_B200_RTAG(AtmosphericPressure, "Static pressure (Pa)", Probe, Scalar, 21)
SOPOT_B200_RTAG(AtmosphericTemperature, "Static temperature (K)", Probe, Scalar, 22)
SOPOT_B200_RTAG(SpeedOfSound, "Speed of sound (m/s)", Probe, Scalar, 23)
}  // namespace environment
#undef SOPOT_B200_RTAG

template <typename Tag>
concept RocketTag = StateTagConcept<Tag> && requires { Tag::source; Tag::shape; Tag::first; };

// value of `Tag` from its source array (`v` = the 14-state vector or the 28-row device record, stride 1)
template <RocketTag Tag, Scalar T>
auto rocketTagValue(const T* v) {
    if constexpr (Tag::shape == TagShape::Scalar) return v[Tag::first];
    else if constexpr (Tag::shape == TagShape::Vec3) return Vector3<T>(v[Tag::first], v[Tag::first + 1], v[Tag::first + 2]);
    else return Quaternion<T>(v[Tag::first], v[Tag::first + 1], v[Tag::first + 2], v[Tag::first + 3]);
}

}  // namespace sopot::rocket
