// rsb_car.h -- per-car pre-tick: suspension rays, tyre friction and the jump/flip/boost
// game logic.  Restates (in our own structure, Bullet units internally):
//   Car::_PreTickUpdate and helpers      /root/reference/src/Sim/Car/Car.cpp:58-131, 339-853
//   btVehicleRL                          /root/reference/src/Sim/btVehicleRL/btVehicleRL.cpp:66-416
//   resolveSingleCollision/Bilateral     bt/BulletDynamics/ConstraintSolver/btContactConstraint.cpp:60-150
// Split in SPMD phases: car_pretick_begin (lanes = cars), wheel_trace (lanes = car x wheel),
// car_pretick_logic (lanes = cars).
#pragma once
#include "rsb_work.h"
#include "rsb_ray.h"

namespace rsb {

RSB_HD uint32_t minstd_next(uint32_t& x) {  // std::minstd_rand0 (std::default_random_engine in libstdc++)
	x = (uint32_t)(((uint64_t)x * 16807ull) % 2147483647ull);
	return x;
}
RSB_HD uint32_t minstd_seed(uint32_t s) { uint32_t x = s % 2147483647u; return x == 0 ? 1u : x; }
// std::default_random_engine(int seed): libstdc++ takes the seed as unsigned long, i.e. a negative int is sign-extended to 64 bits
// before the reduction mod 2^31-1 (bits/random.tcc, linear_congruential_engine::seed)
RSB_HD uint32_t minstd_seed_int(int seed) { uint32_t x = (uint32_t)((uint64_t)(int64_t)seed % 2147483647ull); return x == 0 ? 1u : x; }

// Angle(yaw,0,0).ToRotMat() -> btMatrix3x3::setEulerYPR(yaw,-0,-0) (MathTypes.cpp:72-77, btMatrix3x3.h:290-328)
RSB_HD M3 basis_from_ypr(float yaw, float pitch, float roll) {
	float eulerX = -roll, eulerY = -pitch, eulerZ = yaw;
	float ci = rs_cosf(eulerX), cj = rs_cosf(eulerY), ch = rs_cosf(eulerZ);
	float si = rs_sinf(eulerX), sj = rs_sinf(eulerY), sh = rs_sinf(eulerZ);
	float cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
	return M3(V3(cj * ch, sj * sc - cs, sj * cc + ss), V3(cj * sh, sj * ss + cc, sj * cs - sc), V3(-sj, cj * si, cj * ci));
}

// CarState() defaults + Car::SetState (Car.h:17-120, Car.cpp:23-36): everything except the
// wheel carry-over values and the controls, which SetState does not touch.
RSB_HD void car_reset_state(CarS& c, float boostAmount) {
	c.flipRelTorque = v3zero(); c.worldContactNormal = v3zero();
	c.jumpTime = c.flipTime = c.airTime = c.airTimeSinceJump = 0.f;
	c.boost = boostAmount; c.timeSinceBoosted = 0.f; c.boostingTime = 0.f; c.supersonicTime = 0.f;
	c.handbrakeVal = 0.f; c.autoFlipTimer = 0.f; c.autoFlipTorqueScale = 0.f;
	c.carContactCooldown = 0.f; c.demoRespawnTimer = 0.f; c.carContactOtherID = 0;
	c.isOnGround = 1; c.hasJumped = c.hasDoubleJumped = c.hasFlipped = 0;
	c.isFlipping = c.isJumping = c.isBoosting = c.isSupersonic = 0;
	c.isAutoFlipping = c.worldContactHas = c.isDemoed = c.bhValid = 0;
	for (int i = 0; i < 4; i++) c.wheelsWithContact[i] = 0;
	c.bhRelPos = c.bhBallPos = c.bhExtraVel = v3zero();
	c.bhTickHit = ~0ull; c.bhTickExtra = ~0ull;
	c.lastControls = CtrlS{0, 0, 0, 0, 0, 0, 0, 0, 0};
	c.tickCountSinceUpdate = 0;
	c.velImpulseCache = v3zero();
}

// Car::Respawn (Car.cpp:43-56)
RSB_HD void car_respawn(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	const BatchParams& P = *cx.p;
	int slot = (int)(minstd_next(cx.s->rng) % 4u);  // Math::RandInt(0, 4, -1)
	K::SpawnPos sp = respawn_slot(cx.mode == HOOPS, slot);
	bool blue = (P.team[ci] == 0);
	V3 posUU(sp.x, sp.y * (blue ? 1.f : -1.f), K::CAR_RESPAWN_Z);
	car_reset_state(c, P.mut.carSpawnBoostAmount);
	c.rb.pos = posUU * UU_TO_BT;
	c.rb.basis = P.respBasis[blue ? 0 : 1][slot];  // Angle(yaw + (blue ? 0 : pi), 0, 0).ToRotMat(), host-computed (BatchParams)
	c.rb.vel = v3zero();
	c.rb.angvel = v3zero();
	cx.body[ci + 1].invInertiaW = world_inv_inertia(c.rb.basis, cx.body[ci + 1].invInertiaLocal);
}

// Phase A (per car): input clamp, demo timer / respawn, activation flags (Car.cpp:58-87).
RSB_HD void car_pretick_begin(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	c.controls.throttle = rs_clamp(c.controls.throttle, -1.f, 1.f);
	c.controls.steer = rs_clamp(c.controls.steer, -1.f, 1.f);
	c.controls.pitch = rs_clamp(c.controls.pitch, -1.f, 1.f);
	c.controls.yaw = rs_clamp(c.controls.yaw, -1.f, 1.f);
	c.controls.roll = rs_clamp(c.controls.roll, -1.f, 1.f);
	if (c.isDemoed) {
		float t = c.demoRespawnTimer - cx.p->dt;
		c.demoRespawnTimer = (t > 0.f) ? t : 0.f;
		if (c.demoRespawnTimer == 0.f) car_respawn(cx, ci);  // clears isDemoed
		bw.activation = ACT_DISABLE_SIM;
		bw.noResponse = 1;
	} else {
		bw.activation = ACT_ACTIVE;
		bw.noResponse = 0;
	}
}

// Phase B (per wheel): updateWheelTransform + rayCast (btVehicleRL.cpp:66-94, 111-209).
// `skip` = the car was demoed at the start of the tick (Car.cpp:86-87 early return).
RSB_HD void wheel_trace(const Ctx& cx, int ci, int wi) {
	CarS& c = cx.s->cars[ci];
	const CarCfgD& cfg = cx.p->carCfg[ci];
	const BodyW& bw = cx.body[ci + 1];
	WheelW& w = cx.wheels[ci * 4 + wi];
	const RigidS& rb = c.rb;

	RSB_SUBPROF_BEGIN();
	// updateWheelTransformsWS
	w.isInContact = 0; w.isInContactWithWorld = 0;
	w.hardPointWS = rb.basis * cfg.wheelConn[wi] + rb.pos;
	w.wheelDirWS = rb.basis * V3(0.f, 0.f, -1.f);
	w.wheelAxleWS = rb.basis * V3(0.f, -1.f, 0.f);

	// updateWheelTransform: steering rotation about the suspension "up" axis (uses LAST tick's steer angle)
	{
		V3 up = -w.wheelDirWS;
		V3 right = w.wheelAxleWS;
		V3 fwd = normalized(cross(up, right));
		// btQuaternion(up, steer): sin(0) = 0 and cos(0) = 1 exactly, so the libm calls are skipped for an unsteered wheel
		const float steer = c.wSteer[wi];
		Q4 steeringOrn;
		if (steer == 0.f) { const float s0 = (steer * 0.5f) / len(up);  /* sinf(+-0) = +-0 */ steeringOrn = q4(up.x * s0, up.y * s0, up.z * s0, 1.f); }
		else steeringOrn = q_axis_angle(up, steer);
		M3 steeringMat = m3_from_quat(steeringOrn);
		// basis2 columns: [forward=0] fwd, [right=1] -right, [up=2] up
		M3 basis2(V3(fwd.x, -right.x, up.x), V3(fwd.y, -right.y, up.y), V3(fwd.z, -right.z, up.z));
		w.wtBasis = steeringMat * basis2;
	}

	// rayCast
	const float travelCm = (K::MAX_SUSPENSION_TRAVEL * UU_TO_BT) * 100.f;
	const float suspensionTravel = travelCm / 100.f;
	const float restLen = cfg.suspRest[wi];
	const float radius = cfg.wheelRadius[wi];
	const float realRayLength = restLen + suspensionTravel + radius - K::SUSPENSION_SUBTRACTION;
	V3 source = w.hardPointWS;
	V3 target = source + (w.wheelDirWS * realRayLength);
	w.contactPointWS = target;
	w.groundObj = OBJ_NONE;

	RayHit hit;
	RSB_SUBPROF_MARK(8);
	bool got = cast_vehicle_ray(cx, ci + 1, source, target, hit);
	RSB_SUBPROF_BEGIN2();
	const V3 upVec = rb.basis.col(2);
	if (got) {
		w.contactPointWS = hit.point;
		w.contactNormalWS = hit.normal;
		w.isInContact = 1;
		bool isStatic = hit.obj >= STATIC_BASE;
		w.isInContactWithWorld = isStatic ? 1 : 0;
		w.groundObj = hit.obj;
		float wheelTraceLen = dot(w.hardPointWS - w.contactPointWS, upVec);
		float sl = wheelTraceLen - radius;
		float lo = restLen - suspensionTravel, hi = restLen + suspensionTravel;
		w.suspLen = (sl < lo) ? lo : ((hi < sl) ? hi : sl);  // std::clamp
		float denominator = dot(w.contactNormalWS, upVec);
		V3 relpos = w.contactPointWS - rb.pos;
		V3 velAtContact = vel_at(rb, relpos);
		float projVel = dot(w.contactNormalWS, velAtContact);
		if (denominator > 0.1f) {
			float inv = 1.f / denominator;
			w.suspRelVel = projVel * inv;
			w.clippedInv = inv;
		} else {
			w.suspRelVel = 0.f;
			w.clippedInv = 10.f;
		}
		if (isStatic) {
			float rayPushbackThresh = (restLen + radius) - K::SUSPENSION_SUBTRACTION;
			if (wheelTraceLen < rayPushbackThresh) {
				float wheelTraceDistDelta = wheelTraceLen - rayPushbackThresh;
				// resolveSingleCollision(chassis, static, hitPoint, hitNormal, solverInfo, distance, apply=false)
				V3 rel_pos1 = hit.point - rb.pos;
				V3 vel = vel_at(rb, rel_pos1);  // static body: vel2 = 0
				float rel_vel = dot(hit.normal, vel);
				float restitution = 0.f * -rel_vel;
				const float erp = 0.2f;  // btContactSolverInfo::m_erp
				// solverInfo.m_timeStep is only written inside stepSimulation (btDiscreteDynamicsWorld.cpp:417): on an arena's
				// very first tick -- also the first tick after Arena::DeserializeNew, which builds a new world -- it still holds the
				// constructor default 1/60 (btContactSolverInfo.h:82)
				const float solverTimeStep = cx.s->worldFresh ? (1.f / 60.f) : cx.p->dt;
				float positionalError = erp * -wheelTraceDistDelta / solverTimeStep;
				float velocityError = -(1.0f + restitution) * rel_vel;
				float denom0 = impulse_denominator(rb, bw, hit.point, hit.normal);
				float denom1 = 0.f;
				float jacDiagABInv = 1.f / (denom0 + denom1);
				float normalImpulse = positionalError * jacDiagABInv + velocityError * jacDiagABInv;
				normalImpulse = 0.f > normalImpulse ? 0.f : normalImpulse;
				c.wExtraPushback[wi] = normalImpulse / 4.f;
			}
		}
	} else {
		w.suspLen = restLen + suspensionTravel;
		w.suspRelVel = 0.f;
		w.contactNormalWS = -w.wheelDirWS;
		w.clippedInv = 1.f;
		c.wExtraPushback[wi] = 0.f;
	}
	RSB_SUBPROF_MARK(11);
}

// btVehicleRL::calcFrictionImpulses (btVehicleRL.cpp:306-380)
// One wheel of btVehicleRL::calcFrictionImpulses (btVehicleRL.cpp:306-380).  It only reads the wheel's own ray result, last
// tick's wheel inputs and rigid-body state that nothing writes during the vehicle update, so it runs on the wheel's own thread
// right after its ray (PH_WHEELS) instead of in a 4-wheel loop on the car's thread.
RSB_HD void calc_friction_impulse(const Ctx& cx, int ci, int wi) {
	CarS& c = cx.s->cars[ci];
	const RigidS& rb = c.rb;
	const BodyW& bw = cx.body[ci + 1];
	const float dt = cx.p->dt;
	const float mass = 1.f / bw.invMass;  // getMass()
	const float frictionScale = mass / 3.f;
	{
		WheelW& w = cx.wheels[ci * 4 + wi];
		if (w.groundObj == OBJ_NONE) { w.impulse = v3zero(); return; }
		V3 axleDir = w.wtBasis.col(1);
		V3 surf = w.contactNormalWS;
		float proj = dot(axleDir, surf);
		axleDir -= surf * proj;
		axleDir = safe_normalized(axleDir);
		V3 forwardDir = safe_normalized(cross(surf, axleDir));

		// resolveSingleBilateral(chassis, cp, ground, cp, 0, axleDir, sideImpulse, dt)
		float sideImpulse;
		{
			float nl2 = len2(axleDir);
			if (nl2 > 1.1f) {
				sideImpulse = 0.f;
			} else {
				V3 rel_pos1 = w.contactPointWS - rb.pos;
				V3 vel1 = vel_at(rb, rel_pos1);
				V3 vel2 = v3zero();
				V3 rel_pos2;
				float jacDiagAB;
				M3 world2A = transpose(rb.basis);
				V3 aJ = world2A * cross(rel_pos1, axleDir);
				V3 MinvJtA = bw.invInertiaLocal * aJ;
				if (w.groundObj < STATIC_BASE) {
					const RigidS& g = rigid_of(cx, w.groundObj);
					const BodyW& gw = cx.body[w.groundObj];
					rel_pos2 = w.contactPointWS - g.pos;
					vel2 = vel_at(g, rel_pos2);
					M3 world2B = transpose(g.basis);
					V3 bJ = world2B * cross(rel_pos2, -axleDir);
					V3 MinvJtB = gw.invInertiaLocal * bJ;
					jacDiagAB = bw.invMass + dot(MinvJtA, aJ) + gw.invMass + dot(MinvJtB, bJ);
				} else {
					// static body: zero inverse mass / inertia, identity basis at its origin
					jacDiagAB = bw.invMass + dot(MinvJtA, aJ) + 0.f + 0.f;
				}
				V3 vel = vel1 - vel2;
				float jacDiagABInv = 1.f / jacDiagAB;
				float rel_vel = dot(axleDir, vel);
				const float contactDamping = 0.2f;
				sideImpulse = -contactDamping * rel_vel * jacDiagABInv;
			}
		}

		float rollingFriction;
		if (c.wEngine[wi] == 0.f) {
			if (c.wBrake[wi] != 0.f) {
				V3 contactPoint = w.contactPointWS;
				V3 carRel = contactPoint - rb.pos;
				V3 v1 = vel_at(rb, carRel);
				V3 v2 = v3zero();
				if (w.groundObj < STATIC_BASE) v2 = vel_at(rigid_of(cx, w.groundObj), carRel);  // (sic) car-relative point
				V3 contactVel = v1 - v2;
				float relVel = dot(contactVel, forwardDir);
				if (dt > (1 / 80.f)) {
					float threshold = -(1 / (dt * 150.f)) + 0.8f;
					if (fabsf(relVel) < threshold) relVel = 0.f;
				}
				const float ROLLING_FRICTION_SCALE_MAGIC = 113.73963f;
				rollingFriction = rs_clamp(-relVel * ROLLING_FRICTION_SCALE_MAGIC, -c.wBrake[wi], c.wBrake[wi]);
			} else {
				rollingFriction = 0.f;
			}
		} else {
			rollingFriction = -c.wEngine[wi] / frictionScale;
		}
		V3 total = (forwardDir * rollingFriction * c.wLongFriction[wi]) + (axleDir * sideImpulse * c.wLatFriction[wi]);
		w.impulse = total * frictionScale;
	}
}

// btVehicleRL::getUpwardsDirFromWheelContacts (btVehicleRL.cpp:397-410)
RSB_HD V3 upwards_dir_from_wheels(const Ctx& cx, int ci) {
	V3 sum = v3zero();
	for (int wi = 0; wi < 4; wi++) {
		const WheelW& w = cx.wheels[ci * 4 + wi];
		if (w.isInContact) sum += w.contactNormalWS;
	}
	if (is_zero(sum)) return cx.s->cars[ci].rb.basis.col(2);
	return safe_normalized(sum);
}

// Car::_UpdateWheels (Car.cpp:339-484)
RSB_HD void car_update_wheels(const Ctx& cx, int ci, int numWheelsInContact, float forwardSpeed_UU) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	const RigidS& rb = c.rb;
	const CarCfgD& cfg = cx.p->carCfg[ci];
	const float dt = cx.p->dt;
	float absForwardSpeed_UU = fabsf(forwardSpeed_UU);

	bool wheelsHaveWorldContact = false;
	for (int wi = 0; wi < 4; wi++) wheelsHaveWorldContact |= (cx.wheels[ci * 4 + wi].isInContactWithWorld != 0);

	if (c.controls.handbrake) c.handbrakeVal += K::POWERSLIDE_RISE_RATE * dt;
	else c.handbrakeVal -= K::POWERSLIDE_FALL_RATE * dt;
	c.handbrakeVal = rs_clamp(c.handbrakeVal, 0.f, 1.f);

	float realThrottle = c.controls.throttle;
	float realBrake = 0.f;
	if (c.controls.boost && c.boost > 0.f) realThrottle = 1.f;

	{
		float driveSpeedScale = curve_drive_torque(absForwardSpeed_UU);
		float engineThrottle = realThrottle;
		if (c.controls.handbrake) {
		} else {
			float absThrottle = fabsf(realThrottle);
			if (absThrottle >= K::THROTTLE_DEADZONE) {
				if (absForwardSpeed_UU > K::STOPPING_FORWARD_VEL && rs_sgn(realThrottle) != rs_sgn(forwardSpeed_UU)) {
					realBrake = 1.f;
					if (absForwardSpeed_UU > K::BRAKING_NO_THROTTLE_SPEED_THRESH) engineThrottle = 0.f;
				}
			} else {
				engineThrottle = 0.f;
				bool shouldFullStop = (absForwardSpeed_UU < K::STOPPING_FORWARD_VEL);
				realBrake = shouldFullStop ? 1.f : K::COASTING_BRAKE_FACTOR;
			}
		}
		if (numWheelsInContact < 3) driveSpeedScale /= 4.f;
		float driveEngineForce = engineThrottle * (K::THROTTLE_TORQUE_AMOUNT * UU_TO_BT) * driveSpeedScale;
		float driveBrakeForce = realBrake * (K::BRAKE_TORQUE_AMOUNT * UU_TO_BT);
		for (int wi = 0; wi < 4; wi++) { c.wEngine[wi] = driveEngineForce; c.wBrake[wi] = driveBrakeForce; }
	}

	{
		float steerAngle = cfg.threeWheels ? curve_steer_angle_threewheel(absForwardSpeed_UU) : curve_steer_angle(absForwardSpeed_UU);
		if (c.handbrakeVal != 0.f)
			steerAngle += (curve_powerslide_steer(absForwardSpeed_UU) - steerAngle) * c.handbrakeVal;
		steerAngle *= c.controls.steer;
		c.wSteer[0] = steerAngle;
		c.wSteer[1] = steerAngle;
	}

	for (int wi = 0; wi < 4; wi++) {
		const WheelW& w = cx.wheels[ci * 4 + wi];
		if (w.groundObj == OBJ_NONE) continue;
		V3 vel = rb.vel, angularVel = rb.angvel;
		V3 latDir = w.wtBasis.col(1);
		V3 longDir = cross(latDir, w.contactNormalWS);
		float frictionCurveInput = 0.f;
		V3 wheelDelta = w.hardPointWS - rb.pos;
		V3 crossVec = (cross(angularVel, wheelDelta) + vel) * BT_TO_UU;
		float baseFriction = fabsf(dot(crossVec, latDir));
		if (baseFriction > 5.f) frictionCurveInput = baseFriction / (fabsf(dot(crossVec, longDir)) + baseFriction);
		float latFriction = cfg.threeWheels ? curve_lat_friction_threewheel(frictionCurveInput) : curve_lat_friction(frictionCurveInput);
		float longFriction = curve_long_friction(frictionCurveInput);
		if (c.handbrakeVal != 0.f) {
			float hb = c.handbrakeVal;
			latFriction *= (curve_handbrake_lat(frictionCurveInput) - 1.f) * hb + 1.f;
			longFriction *= (curve_handbrake_long(frictionCurveInput) - 1.f) * hb + 1.f;
		} else {
			longFriction = 1.f;
		}
		bool isContactSticky = realThrottle != 0.f;
		if (!isContactSticky) {
			float nonStickyScale = curve_non_sticky_friction(w.contactNormalWS.z);
			latFriction *= nonStickyScale;
			longFriction *= nonStickyScale;
		}
		c.wLatFriction[wi] = latFriction;
		c.wLongFriction[wi] = longFriction;
	}

	if (wheelsHaveWorldContact) {
		V3 upwardsDir = upwards_dir_from_wheels(cx, ci);
		bool fullStick = (realThrottle != 0.f) || (absForwardSpeed_UU > K::STOPPING_FORWARD_VEL);
		float stickyForceScale = cfg.threeWheels ? 0.f : 0.5f;
		if (fullStick) stickyForceScale += 1.f - fabsf(upwardsDir.z);
		apply_central_force(bw, upwardsDir * stickyForceScale * (K::GRAVITY_Z * UU_TO_BT) * K::CAR_MASS_BT);
	}
}

// Car::_UpdateAirTorque (Car.cpp:584-669)
RSB_HD void car_update_air_torque(const Ctx& cx, int ci, bool updateAirControl) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	const RigidS& rb = c.rb;
	V3 fwdDir = rb.basis.col(0), rightDir = rb.basis.col(1), upDir = rb.basis.col(2);
	V3 dirPitch_right = -rightDir, dirYaw_up = upDir, dirRoll_forward = -fwdDir;

	bool doAirControl = false;
	if (c.isFlipping) c.isFlipping = (c.hasFlipped && c.flipTime < K::FLIP_TORQUE_TIME) ? 1 : 0;

	if (c.isFlipping) {
		V3 relDodgeTorque = c.flipRelTorque;
		if (!is_zero(c.flipRelTorque)) {
			float pitchScale = 1.f;
			if (relDodgeTorque.y != 0.f && c.controls.pitch != 0.f) {
				if (rs_sgn(relDodgeTorque.y) == rs_sgn(c.controls.pitch)) {
					float ap = fabsf(c.controls.pitch);
					pitchScale = 1.f - ((ap < 1.f) ? ap : 1.f);
					doAirControl = true;
				}
			}
			relDodgeTorque.y *= pitchScale;
			V3 dodgeTorque = relDodgeTorque * V3(K::FLIP_TORQUE_X, K::FLIP_TORQUE_Y, 0.f);
			apply_torque(bw, (inverse(bw.invInertiaW) * rb.basis) * dodgeTorque);
		} else {
			doAirControl = true;
		}
	} else {
		doAirControl = true;
	}

	doAirControl &= !c.isAutoFlipping;
	doAirControl &= updateAirControl;
	if (doAirControl) {
		float pitchTorqueScale = 1.f;
		V3 torque;
		if (c.controls.pitch != 0.f || c.controls.yaw != 0.f || c.controls.roll != 0.f) {
			if (c.isFlipping) {
				pitchTorqueScale = 0.f;
			} else if (c.hasFlipped) {
				if (c.flipTime < K::FLIP_TORQUE_TIME + K::FLIP_PITCHLOCK_EXTRA_TIME) pitchTorqueScale = 0.f;
			}
			torque = (c.controls.pitch * dirPitch_right * pitchTorqueScale * K::AIR_TORQUE_P) +
			         (c.controls.yaw * dirYaw_up * K::AIR_TORQUE_Y) +
			         (c.controls.roll * dirRoll_forward * K::AIR_TORQUE_R);
		} else {
			torque = v3zero();
		}
		V3 angVel = rb.angvel;
		float dampPitch = dot(dirPitch_right, angVel) * K::AIR_DAMP_P * (1.f - fabsf(c.controls.pitch * pitchTorqueScale));
		float dampYaw = dot(dirYaw_up, angVel) * K::AIR_DAMP_Y * (1.f - fabsf(c.controls.yaw));
		float dampRoll = dot(dirRoll_forward, angVel) * K::AIR_DAMP_R;
		V3 damping = (dirYaw_up * dampYaw) + (dirPitch_right * dampPitch) + (dirRoll_forward * dampRoll);
		apply_torque(bw, (inverse(bw.invInertiaW) * (torque - damping)) * K::CAR_TORQUE_SCALE);
	}

	if (c.controls.throttle != 0.f)
		apply_central_force(bw, fwdDir * c.controls.throttle * K::THROTTLE_AIR_ACCEL * UU_TO_BT * K::CAR_MASS_BT);
}

// Car::_UpdateJump (Car.cpp:535-582)
RSB_HD void car_update_jump(const Ctx& cx, int ci, bool jumpPressed) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	const MutD& mut = cx.p->mut;
	V3 upDir = c.rb.basis.col(2);
	if (c.isOnGround && !c.isJumping) {
		if (c.hasJumped && c.jumpTime < K::JUMP_MIN_TIME + K::JUMP_RESET_TIME_PAD) {
		} else {
			c.hasJumped = 0;
			c.jumpTime = 0.f;
		}
	}
	if (c.isJumping) {
		if (c.jumpTime < K::JUMP_MIN_TIME || (c.controls.jump && c.jumpTime < K::JUMP_MAX_TIME)) c.isJumping = 1;
		else c.isJumping = 0;
	} else if (c.isOnGround && jumpPressed) {
		c.isJumping = 1;
		c.jumpTime = 0.f;
		V3 jumpStartForce = upDir * mut.jumpImmediateForce * UU_TO_BT * K::CAR_MASS_BT;
		apply_central_impulse(c.rb, bw, jumpStartForce);
	}
	if (c.isJumping) {
		c.hasJumped = 1;
		V3 totalJumpForce = upDir * mut.jumpAccel;
		if (c.jumpTime < K::JUMP_MIN_TIME) totalJumpForce *= 0.62f;
		apply_central_force(bw, totalJumpForce * UU_TO_BT * K::CAR_MASS_BT);
	}
	if (c.isJumping || c.hasJumped) c.jumpTime += cx.p->dt;
}

// Angle::FromRotMat roll component (MathTypes.cpp:62-70; btMatrix3x3::getEulerYPR :527-552)
RSB_HD float roll_from_basis(const M3& m) {
	float yaw = rs_atan2f(m.r[1].x, m.r[0].x);
	float pitch = asinf(-m.r[2].x);
	float roll = rs_atan2f(m.r[2].y, m.r[2].z);
	if (fabsf(pitch) == kHalfPi) {
		if (roll > 0.f) roll -= kPi; else roll += kPi;
	}
	(void)yaw;
	return roll * -1.f;
}

// Car::_UpdateAutoFlip (Car.cpp:783-817)
RSB_HD void car_update_auto_flip(const Ctx& cx, int ci, bool jumpPressed) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	V3 upDir = c.rb.basis.col(2), fwdDir = c.rb.basis.col(0);
	if (jumpPressed && c.worldContactHas && c.worldContactNormal.z > K::CAR_AUTOFLIP_NORMZ_THRESH) {
		float roll = roll_from_basis(c.rb.basis);
		float absRoll = fabsf(roll);
		if (absRoll > K::CAR_AUTOFLIP_ROLL_THRESH) {
			c.autoFlipTimer = (float)((double)K::CAR_AUTOFLIP_TIME * ((double)absRoll / 3.14159265358979323846));
			c.autoFlipTorqueScale = (roll > 0.f) ? 1.f : -1.f;
			c.isAutoFlipping = 1;
			apply_central_impulse(c.rb, bw, (-upDir) * K::CAR_AUTOFLIP_IMPULSE * UU_TO_BT * K::CAR_MASS_BT);
		}
	}
	if (c.isAutoFlipping) {
		if (c.autoFlipTimer <= 0.f) {
			c.isAutoFlipping = 0;
			c.autoFlipTimer = 0.f;
		} else {
			c.rb.angvel += fwdDir * K::CAR_AUTOFLIP_TORQUE * c.autoFlipTorqueScale * cx.p->dt;
			c.autoFlipTimer -= cx.p->dt;
		}
	}
}

// Car::_UpdateDoubleJumpOrFlip (Car.cpp:671-781)
RSB_HD void car_update_double_jump_or_flip(const Ctx& cx, int ci, bool jumpPressed, float forwardSpeed_UU) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	const MutD& mut = cx.p->mut;
	const CarCfgD& cfg = cx.p->carCfg[ci];
	const float dt = cx.p->dt;
	V3 upDir = c.rb.basis.col(2), fwdDir = c.rb.basis.col(0);
	if (c.isOnGround) {
		c.hasDoubleJumped = 0; c.hasFlipped = 0;
		c.airTime = 0.f; c.airTimeSinceJump = 0.f; c.flipTime = 0.f;
	} else {
		c.airTime += dt;
		if (c.hasJumped && !c.isJumping) c.airTimeSinceJump += dt;
		else c.airTimeSinceJump = 0.f;

		if (jumpPressed && c.airTimeSinceJump < K::DOUBLEJUMP_MAX_DELAY) {
			float inputMagnitude = fabsf(c.controls.yaw) + fabsf(c.controls.pitch) + fabsf(c.controls.roll);
			bool isFlipInput = inputMagnitude >= cfg.dodgeDeadzone;
			bool canUse;
			if (isFlipInput) canUse = (!c.hasDoubleJumped && !c.hasFlipped) || mut.unlimitedFlips;
			else canUse = (!c.hasDoubleJumped && !c.hasFlipped) || mut.unlimitedDoubleJumps;
			if (c.isAutoFlipping) canUse = false;
			if (canUse) {
				if (isFlipInput) {
					c.flipTime = 0.f;
					c.hasFlipped = 1;
					c.isFlipping = 1;
					float forwardSpeedRatio = fabsf(forwardSpeed_UU) / K::CAR_MAX_SPEED;
					V3 dodgeDir(-c.controls.pitch, c.controls.yaw + c.controls.roll, 0.f);
					if (fabsf(c.controls.yaw + c.controls.roll) < 0.1f && fabsf(c.controls.pitch) < 0.1f) dodgeDir = v3zero();
					else dodgeDir = safe_normalized(dodgeDir);
					c.flipRelTorque = V3(-dodgeDir.y, dodgeDir.x, 0.f);
					if (fabsf(dodgeDir.x) < 0.1f) dodgeDir.x = 0.f;
					if (fabsf(dodgeDir.y) < 0.1f) dodgeDir.y = 0.f;
					if (!fuzzy_zero(dodgeDir)) {
						bool shouldDodgeBackwards;
						if (fabsf(forwardSpeed_UU) < 100.0f) shouldDodgeBackwards = dodgeDir.x < 0.0f;
						else shouldDodgeBackwards = (dodgeDir.x >= 0.0f) != (forwardSpeed_UU >= 0.0f);
						V3 initalDodgeVel = dodgeDir * K::FLIP_INITIAL_VEL_SCALE;
						float maxSpeedScaleX = shouldDodgeBackwards ? K::FLIP_BACKWARD_IMPULSE_MAX_SPEED_SCALE : K::FLIP_FORWARD_IMPULSE_MAX_SPEED_SCALE;
						initalDodgeVel.x *= ((maxSpeedScaleX - 1.f) * forwardSpeedRatio) + 1.f;
						initalDodgeVel.y *= ((K::FLIP_SIDE_IMPULSE_MAX_SPEED_SCALE - 1.f) * forwardSpeedRatio) + 1.f;
						if (shouldDodgeBackwards) initalDodgeVel.x *= K::FLIP_BACKWARD_IMPULSE_SCALE_X;
						V3 forwardDir2D = rs_normalized(V3(fwdDir.x, fwdDir.y, 0.f));
						V3 rightDir2D(-forwardDir2D.y, forwardDir2D.x, 0.f);
						V3 finalDeltaVel = initalDodgeVel.x * forwardDir2D + initalDodgeVel.y * rightDir2D;
						apply_central_impulse(c.rb, bw, finalDeltaVel * UU_TO_BT * K::CAR_MASS_BT);
					}
				} else {
					V3 jumpStartForce = upDir * K::JUMP_IMMEDIATE_FORCE * UU_TO_BT * K::CAR_MASS_BT;
					apply_central_impulse(c.rb, bw, jumpStartForce);
					c.hasDoubleJumped = 1;
				}
			}
		}
	}

	if (c.isFlipping) {
		c.flipTime += dt;
		if (c.flipTime <= K::FLIP_TORQUE_TIME) {
			if (c.flipTime >= K::FLIP_Z_DAMP_START && (c.rb.vel.z < 0.f || c.flipTime < K::FLIP_Z_DAMP_END))
				c.rb.vel.z *= cx.p->flipZDampFactor;  // powf(1 - FLIP_Z_DAMP_120, dt / (1 / 120)), a batch constant
		}
	} else if (c.hasFlipped) {
		c.flipTime += dt;
	}
}

// Car::_UpdateAutoRoll (Car.cpp:819-853)
RSB_HD void car_update_auto_roll(const Ctx& cx, int ci, int numWheelsInContact) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	V3 groundUpDir = (numWheelsInContact > 0) ? upwards_dir_from_wheels(cx, ci) : c.worldContactNormal;
	V3 groundDownDir = -groundUpDir;
	V3 forwardDir = c.rb.basis.col(0), rightDir = c.rb.basis.col(1);
	V3 crossRightDir = cross(groundUpDir, forwardDir);
	V3 crossForwardDir = cross(groundDownDir, crossRightDir);
	float rightTorqueFactor = 1.f - rs_clamp(dot(rightDir, crossRightDir), 0.f, 1.f);
	float forwardTorqueFactor = 1.f - rs_clamp(dot(forwardDir, crossForwardDir), 0.f, 1.f);
	V3 torqueDirRight = forwardDir * (dot(rightDir, groundUpDir) >= 0.f ? -1.f : 1.f);
	V3 torqueDirForward = rightDir * (dot(forwardDir, groundUpDir) >= 0.f ? 1.f : -1.f);
	V3 torqueRight = torqueDirRight * rightTorqueFactor;
	V3 torqueForward = torqueDirForward * forwardTorqueFactor;
	apply_central_force(bw, groundDownDir * K::CAR_AUTOROLL_FORCE * UU_TO_BT * K::CAR_MASS_BT);
	apply_torque(bw, (inverse(bw.invInertiaW) * (torqueForward + torqueRight)) * K::CAR_AUTOROLL_TORQUE);
}

// btVehicleRL::updateSuspension + applyFrictionImpulses (btVehicleRL.cpp:270-303, 383-395)
RSB_HD void vehicle_second(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	const BodyW& bw = cx.body[ci + 1];
	const CarCfgD& cfg = cx.p->carCfg[ci];
	const float dt = cx.p->dt;
	for (int wi = 0; wi < 4; wi++) {
		WheelW& w = cx.wheels[ci * 4 + wi];
		if (w.isInContact) {
			float force = (cfg.suspRest[wi] - w.suspLen) * K::SUSPENSION_STIFFNESS * w.clippedInv;
			float dampingVelScale = (w.suspRelVel < 0.f) ? K::WHEELS_DAMPING_COMPRESSION : K::WHEELS_DAMPING_RELAXATION;
			w.suspForce = force - (dampingVelScale * w.suspRelVel);
			w.suspForce *= cfg.forceScale[wi];
			if (w.suspForce < 0.f) w.suspForce = 0.f;
		} else {
			w.suspForce = 0.f;
		}
	}
	for (int wi = 0; wi < 4; wi++) {
		const WheelW& w = cx.wheels[ci * 4 + wi];
		if (w.suspForce != 0.f) {
			V3 contactPointOffset = w.contactPointWS - c.rb.pos;
			float baseForceScale = (w.suspForce * dt) + c.wExtraPushback[wi];
			V3 force = w.contactNormalWS * baseForceScale;
			apply_impulse(c.rb, bw, force, contactPointOffset);
		}
	}
	V3 upDir = c.rb.basis.col(2);
	for (int wi = 0; wi < 4; wi++) {
		const WheelW& w = cx.wheels[ci * 4 + wi];
		if (!is_zero(w.impulse)) {
			V3 wheelContactOffset = w.contactPointWS - c.rb.pos;
			float contactUpDot = dot(upDir, wheelContactOffset);
			V3 wheelRelPos = wheelContactOffset - upDir * contactUpDot;
			apply_impulse(c.rb, bw, w.impulse * dt, wheelRelPos);
		}
	}
}

// Car::_UpdateBoost (Car.cpp:486-533)
RSB_HD void car_update_boost(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	BodyW& bw = cx.body[ci + 1];
	const MutD& mut = cx.p->mut;
	const float dt = cx.p->dt;
	bool hasBoost = c.boost > 0.f;
	if (hasBoost) {
		if (c.isBoosting) {
			if (c.controls.boost || c.boostingTime < K::BOOST_MIN_TIME) c.isBoosting = 1;
			else c.isBoosting = 0;
		} else {
			if (c.controls.boost) c.isBoosting = c.controls.boost;
		}
	} else {
		c.isBoosting = 0;
	}
	if (c.isBoosting) c.boostingTime += dt;
	else c.boostingTime = 0.f;

	if (c.isBoosting) {
		float nb = c.boost - mut.boostUsedPerSecond * dt;
		c.boost = (nb > 0.f) ? nb : 0.f;
		float s = (c.isOnGround ? mut.boostAccelGround : mut.boostAccelAir) * UU_TO_BT;
		V3 f = (s * c.rb.basis.col(0)) * K::CAR_MASS_BT;
		apply_central_force(bw, f);
		c.timeSinceBoosted = 0.f;
	} else {
		c.timeSinceBoosted += dt;
		if (mut.rechargeBoostEnabled)
			if (c.timeSinceBoosted >= mut.rechargeBoostDelay) c.boost += mut.rechargeBoostPerSecond * dt;
	}
	c.boost = (c.boost < K::BOOST_MAX) ? c.boost : K::BOOST_MAX;
}

// Phase C (per car): everything of Car::_PreTickUpdate after the traces (Car.cpp:89-130).
RSB_HD void car_pretick_logic(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	RSB_SUBPROF_BEGIN();

	bool jumpPressed = c.controls.jump && !c.lastControls.jump;
	int numWheelsInContact = 0;
	for (int wi = 0; wi < 4; wi++) {
		int inContact = cx.wheels[ci * 4 + wi].isInContact;
		c.wheelsWithContact[wi] = (uint8_t)inContact;
		numWheelsInContact += inContact;
	}
	c.isOnGround = numWheelsInContact >= 3;

	float forwardSpeed_UU = dot(c.rb.vel, c.rb.basis.col(0)) * BT_TO_UU;
	car_update_wheels(cx, ci, numWheelsInContact, forwardSpeed_UU);
	RSB_SUBPROF_MARK(17);

	if (numWheelsInContact < 3) car_update_air_torque(cx, ci, numWheelsInContact == 0);
	else c.isFlipping = 0;

	car_update_jump(cx, ci, jumpPressed);
	car_update_auto_flip(cx, ci, jumpPressed);
	car_update_double_jump_or_flip(cx, ci, jumpPressed, forwardSpeed_UU);

	if (c.controls.throttle != 0.f && ((numWheelsInContact > 0 && numWheelsInContact < 4) || c.worldContactHas))
		car_update_auto_roll(cx, ci, numWheelsInContact);

	c.worldContactHas = 0;
	RSB_SUBPROF_MARK(18);
	vehicle_second(cx, ci);
	RSB_SUBPROF_MARK(19);
	car_update_boost(cx, ci);
	RSB_SUBPROF_MARK(20);
}

// Car::_PostTickUpdate + _FinishPhysicsTick (Car.cpp:133-193)
RSB_HD void car_posttick(const Ctx& cx, int ci) {
	CarS& c = cx.s->cars[ci];
	if (c.isDemoed) return;
	const float dt = cx.p->dt;
	{
		float speedSquared = len2(c.rb.vel * BT_TO_UU);
		if (c.isSupersonic && c.supersonicTime < K::SUPERSONIC_MAINTAIN_MAX_TIME)
			c.isSupersonic = (speedSquared >= K::SUPERSONIC_MAINTAIN_MIN_SPEED * K::SUPERSONIC_MAINTAIN_MIN_SPEED);
		else
			c.isSupersonic = (speedSquared >= K::SUPERSONIC_START_SPEED * K::SUPERSONIC_START_SPEED);
		if (c.isSupersonic) c.supersonicTime += dt;
		else c.supersonicTime = 0.f;
	}
	if (c.carContactCooldown > 0.f) {
		float t = c.carContactCooldown - dt;
		c.carContactCooldown = (t > 0.f) ? t : 0.f;
	}
	c.lastControls = c.controls;

	// _FinishPhysicsTick
	if (!is_zero(c.velImpulseCache)) {
		c.rb.vel += c.velImpulseCache;
		c.velImpulseCache = v3zero();
	}
	{
		V3 vel = c.rb.vel, angVel = c.rb.angvel;
		const float maxSpeed = K::CAR_MAX_SPEED * UU_TO_BT;
		if (len2(vel) > maxSpeed * maxSpeed) vel = normalized(vel) * maxSpeed;
		if (len2(angVel) > (K::CAR_MAX_ANG_SPEED * K::CAR_MAX_ANG_SPEED)) angVel = normalized(angVel) * K::CAR_MAX_ANG_SPEED;
		c.rb.vel = vel;
		c.rb.angvel = angVel;
	}
	c.tickCountSinceUpdate++;
}

}  // namespace rsb
