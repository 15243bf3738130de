// rsb_consts.h -- gameplay constants of the tick (values of reference src/RLConst.h:9-528,
// written with the same float expressions so both compilers fold them identically).
#pragma once
#include "rsb_math.h"

namespace rsb {
namespace K {
constexpr float GRAVITY_Z = -650.f;
constexpr float ARENA_EXTENT_X = 4096, ARENA_EXTENT_Y = 5120, ARENA_HEIGHT = 2048;
constexpr float ARENA_EXTENT_X_HOOPS = (8900 / 3.f), ARENA_EXTENT_Y_HOOPS = 3581, ARENA_HEIGHT_HOOPS = 1820;
constexpr float ARENA_COLLISION_BASE_FRICTION = 0.6f, ARENA_COLLISION_BASE_RESTITUTION = 0.3f;
constexpr float CAR_MASS_BT = 180.f;
constexpr float BALL_MASS_BT = CAR_MASS_BT / 6.f;
constexpr float CAR_COLLISION_FRICTION = 0.3f, CAR_COLLISION_RESTITUTION = 0.1f;
constexpr float CARBALL_COLLISION_FRICTION = 2.0f, CARBALL_COLLISION_RESTITUTION = 0.0f;
constexpr float CARCAR_COLLISION_FRICTION = 0.09f, CARCAR_COLLISION_RESTITUTION = 0.1f;
constexpr float BALL_REST_Z = 93.15f;
constexpr float BALL_MAX_ANG_SPEED = 6.f;
constexpr float BALL_HOOPS_LAUNCH_Z_VEL = 1000, BALL_HOOPS_LAUNCH_DELAY = 0.265f;
constexpr float CAR_MAX_SPEED = 2300.f;
constexpr float BOOST_MAX = 100.f, BOOST_MIN_TIME = 0.1f;
constexpr float CAR_MAX_ANG_SPEED = 5.5f;
constexpr float SUPERSONIC_START_SPEED = 2200.f;
constexpr float SUPERSONIC_MAINTAIN_MIN_SPEED = SUPERSONIC_START_SPEED - 100.f;
constexpr float SUPERSONIC_MAINTAIN_MAX_TIME = 1.f;
constexpr float POWERSLIDE_RISE_RATE = 5, POWERSLIDE_FALL_RATE = 2;
constexpr float THROTTLE_TORQUE_AMOUNT = CAR_MASS_BT * 400.f;
constexpr float BRAKE_TORQUE_AMOUNT = CAR_MASS_BT * (14.25f + (1.f / 3.f));
constexpr float STOPPING_FORWARD_VEL = 25.f, COASTING_BRAKE_FACTOR = 0.15f, BRAKING_NO_THROTTLE_SPEED_THRESH = 0.01f;
constexpr float THROTTLE_DEADZONE = 0.001f;
constexpr float THROTTLE_AIR_ACCEL = 200 / 3.f;
constexpr float JUMP_IMMEDIATE_FORCE = 875.f / 3.f;
constexpr float JUMP_MIN_TIME = 0.025f, JUMP_RESET_TIME_PAD = (1 / 40.f), JUMP_MAX_TIME = 0.2f, DOUBLEJUMP_MAX_DELAY = 1.25f;
constexpr float FLIP_Z_DAMP_120 = 0.35f, FLIP_Z_DAMP_START = 0.15f, FLIP_Z_DAMP_END = 0.21f;
constexpr float FLIP_TORQUE_TIME = 0.65f, FLIP_PITCHLOCK_EXTRA_TIME = 0.3f, FLIP_INITIAL_VEL_SCALE = 500.f;
constexpr float FLIP_TORQUE_X = 260.f, FLIP_TORQUE_Y = 224.f;
constexpr float FLIP_FORWARD_IMPULSE_MAX_SPEED_SCALE = 1.f, FLIP_SIDE_IMPULSE_MAX_SPEED_SCALE = 1.9f;
constexpr float FLIP_BACKWARD_IMPULSE_MAX_SPEED_SCALE = 2.5f, FLIP_BACKWARD_IMPULSE_SCALE_X = 16.f / 15.f;
constexpr float HOOPS_GOAL_SCORE_THRESHOLD_Z = 270.f;
constexpr float CAR_TORQUE_SCALE = (float)(2 * 3.14159265358979323846 / (1 << 16) * 1000);
constexpr float CAR_AUTOFLIP_IMPULSE = 200, CAR_AUTOFLIP_TORQUE = 50, CAR_AUTOFLIP_TIME = 0.4f;
constexpr float CAR_AUTOFLIP_NORMZ_THRESH = (float)0.70710678118654752440;
constexpr float CAR_AUTOFLIP_ROLL_THRESH = 2.8f;
constexpr float CAR_AUTOROLL_FORCE = 100, CAR_AUTOROLL_TORQUE = 80;
constexpr float BALL_CAR_EXTRA_IMPULSE_Z_SCALE = 0.35f;
constexpr float BALL_CAR_EXTRA_IMPULSE_Z_SCALE_HOOPS_GROUND = BALL_CAR_EXTRA_IMPULSE_Z_SCALE * 1.55f;
constexpr float BALL_CAR_EXTRA_IMPULSE_FORWARD_SCALE = 0.65f;
constexpr float BALL_CAR_EXTRA_IMPULSE_MAXDELTAVEL_UU = 4600.f;
constexpr float BALL_CAR_EXTRA_IMPULSE_Z_SCALE_HOOPS_NORMAL_Z_THRESH = 0.1f;
constexpr float CAR_SPAWN_REST_Z = 17.f, CAR_RESPAWN_Z = 36.f;
constexpr float BUMP_MIN_FORWARD_DIST = 64.5f;
constexpr float AIR_TORQUE_P = 130, AIR_TORQUE_Y = 95, AIR_TORQUE_R = 400;   // CAR_AIR_CONTROL_TORQUE
constexpr float AIR_DAMP_P = 30, AIR_DAMP_Y = 20, AIR_DAMP_R = 50;           // CAR_AIR_CONTROL_DAMPING
// BTVehicle
constexpr float SUSPENSION_FORCE_SCALE_FRONT = 36.f - (1.f / 4.f);
constexpr float SUSPENSION_FORCE_SCALE_BACK = 54.f + (1.f / 4.f) + (1.5f / 100.f);
constexpr float SUSPENSION_STIFFNESS = 500.f, WHEELS_DAMPING_COMPRESSION = 25.f, WHEELS_DAMPING_RELAXATION = 40.f;
constexpr float MAX_SUSPENSION_TRAVEL = 12.f, SUSPENSION_SUBTRACTION = 0.05f;
// Heatseeker
constexpr float HS_INITIAL_TARGET_SPEED = 2900, HS_TARGET_SPEED_INCREMENT = 85, HS_MIN_SPEEDUP_INTERVAL = 1;
constexpr float HS_TARGET_Y = 5120, HS_TARGET_Z = 320, HS_HORIZONTAL_BLEND = 1.45f, HS_VERTICAL_BLEND = 0.78f, HS_SPEED_BLEND = 0.3f;
constexpr float HS_MAX_TURN_PITCH = (float)(7000 * 3.14159265358979323846 / (1 << 15));
constexpr float HS_MAX_SPEED = 4600, HS_WALL_BOUNCE_CHANGE_Y_THRESH = 300, HS_WALL_BOUNCE_CHANGE_Y_NORMAL = 0.5f;
constexpr float HS_WALL_BOUNCE_FORCE_SCALE = 1 / 3.f, HS_WALL_BOUNCE_UP_FRAC = 0.3f;
// Snowday
constexpr float PUCK_RADIUS = 114.25f, PUCK_HEIGHT = 62.5f, PUCK_GROUND_STICK_FORCE = 70;
// Boost pads
constexpr float PAD_CYL_HEIGHT = 95, PAD_CYL_RAD_BIG = 208, PAD_CYL_RAD_SMALL = 144;
constexpr float PAD_BOX_HEIGHT = 64, PAD_BOX_RAD_BIG = 160, PAD_BOX_RAD_SMALL = 120;
constexpr float PAD_BOOST_AMOUNT_BIG = 100, PAD_BOOST_AMOUNT_SMALL = 12;

struct SpawnPos { float x, y, yaw; };
}  // namespace K

// Piecewise-linear curves (RLConst.h:418-527)
#define RSB_CURVE(name, N, ...) \
	RSB_HD float name(float in) { const Curve<N> c = __VA_ARGS__; return curve_eval<N>(c, in); }

RSB_CURVE(curve_steer_angle, 6, {{0, 500, 1000, 1500, 1750, 3000}, {0.53356f, 0.31930f, 0.18203f, 0.10570f, 0.08507f, 0.03454f}})
RSB_CURVE(curve_steer_angle_threewheel, 2, {{0, 2300}, {0.342473f, 0.034837f}})
RSB_CURVE(curve_powerslide_steer, 2, {{0, 2500}, {0.39235f, 0.12610f}})
RSB_CURVE(curve_drive_torque, 3, {{0, 1400, 1410}, {1.0f, 0.1f, 0.0f}})
RSB_CURVE(curve_non_sticky_friction, 3, {{0, 0.7075f, 1}, {0.1f, 0.5f, 1.0f}})
RSB_CURVE(curve_lat_friction, 2, {{0, 1}, {1.0f, 0.2f}})
RSB_CURVE(curve_lat_friction_threewheel, 2, {{0, 1}, {0.30f, 0.25f}})
RSB_HD float curve_long_friction(float) { return 1.f; }  // empty curve -> default output 1
RSB_CURVE(curve_handbrake_lat, 1, {{0}, {0.1f}})
RSB_CURVE(curve_handbrake_long, 2, {{0, 1}, {0.5f, 0.9f}})
RSB_CURVE(curve_ball_car_extra_impulse, 4, {{0, 500.f, 2300.f, 4600.f}, {0.65f, 0.65f, 0.55f, 0.30f}})
RSB_CURVE(curve_bump_vel_ground, 3, {{0.f, 1400.f, 2200.f}, {(5.f / 6.f), 1100.f, 1530.f}})
RSB_CURVE(curve_bump_vel_air, 3, {{0.f, 1400.f, 2200.f}, {(5.f / 6.f), 1390.f, 1945.f}})
RSB_CURVE(curve_bump_upward, 3, {{0.f, 1400.f, 2200.f}, {(2.f / 6.f), 278.f, 417.f}})

// Respawn slots (RLConst.h:383-411), blue side; orange mirrors y and adds pi to yaw.
RSB_HD K::SpawnPos respawn_slot(bool hoops, int i) {
	const float H = (float)(3.14159265358979323846 / 2);
	if (hoops) {
		const K::SpawnPos t[4] = {{-1920, -3072, H}, {-1152, -3072, H}, {1920, -3072, H}, {1152, -3072, H}};
		return t[i];
	}
	const K::SpawnPos t[4] = {{-2304, -4608, H}, {-2688, -4608, H}, {2304, -4608, H}, {2688, -4608, H}};
	return t[i];
}

}  // namespace rsb
