// rsb_ball.h -- ball game logic and boost pads.  Restates
//   Ball::_PreTickUpdate / _OnHit / _OnWorldCollision / _FinishPhysicsTick
//       /root/reference/src/Sim/Ball/Ball.cpp:117-143, 161-235, 237-324, 326-362
//   Angle::FromVec / GetForwardVec / NormalizeFix, Math::RoundAngleUE3
//       /root/reference/src/Math/MathTypes/MathTypes.cpp:79-116, src/Math/Math.cpp:66-91
//   BoostPad / BoostPadGrid   /root/reference/src/Sim/BoostPad/BoostPad.cpp:51-105,
//                             src/Sim/BoostPad/BoostPadGrid/BoostPadGrid.cpp:5-25
//   Arena::IsBallScored       /root/reference/src/Sim/Arena/Arena.cpp:760-769, 893-927
#pragma once
#include "rsb_work.h"

namespace rsb {

struct Ang { float yaw, pitch, roll; };

RSB_HD float wrap_normalize(float val, float minmax) {
	float result = fmodf(val, minmax * 2);
	if (result > minmax) result -= minmax * 2;
	else if (result < -minmax) result += minmax * 2;
	return result;
}
RSB_HD void ang_normalize_fix(Ang& a) {
	a.yaw = wrap_normalize(a.yaw, (float)3.14159265358979323846);
	a.pitch = wrap_normalize(a.pitch, (float)(3.14159265358979323846 / 2));
	a.roll = wrap_normalize(a.roll, (float)3.14159265358979323846);
}
RSB_HD Ang ang_from_vec(const V3& f) {
	Ang a; a.roll = 0.f;
	if (fabsf(f.y) > kEps || fabsf(f.x) > kEps) {
		a.yaw = rs_atan2f(f.y, f.x);
		float dist2D = sqrtf(f.x * f.x + f.y * f.y);
		a.pitch = -rs_atan2f(-f.z, dist2D);
	} else {
		a.yaw = 0.f;
		if (f.z > kEps) a.pitch = (float)(3.14159265358979323846 / 2);
		else if (f.z < -kEps) a.pitch = (float)(-3.14159265358979323846 / 2);
		else a.pitch = 0.f;
	}
	return a;
}
RSB_HD V3 ang_forward(const Ang& a) {
	float cp = rs_cosf(-a.pitch), cy = rs_cosf(a.yaw), sy = rs_sinf(a.yaw), sp = rs_sinf(-a.pitch);
	return V3(cp * cy, cp * sy, -sp);
}
RSB_HD Ang round_angle_ue3(Ang ang) {
	const float TO_INTS = (float)((double)(1 << 15) / 3.14159265358979323846);
	const float BACK_TO_RADIANS = (1.f / TO_INTS) * 4.f;
	const int ROUNDING_MASK = 0x4000 - 1;
	int rYaw = ((int)(ang.yaw * TO_INTS) >> 2) & ROUNDING_MASK;
	int rPitch = ((int)(ang.pitch * TO_INTS) >> 2) & ROUNDING_MASK;
	ang.yaw = rYaw * BACK_TO_RADIANS;
	ang.pitch = rPitch * BACK_TO_RADIANS;
	return ang;
}

// Ball::_PreTickUpdate
RSB_HD void ball_pretick(const Ctx& cx) {
	BallS& b = cx.s->ball;
	BodyW& bw = cx.body[0];
	const int mode = cx.mode;
	const float dt = cx.p->dt;
	if (mode == HEATSEEKER) {
		V3 pos = b.rb.pos * BT_TO_UU, vel = b.rb.vel * BT_TO_UU;
		float yTargetDir = b.hsYTargetDir;
		if (yTargetDir != 0.f) {
			Ang velAngle = ang_from_vec(vel);
			V3 goalTargetPos(0.f, K::HS_TARGET_Y * yTargetDir, K::HS_TARGET_Z);
			Ang angleToGoal = ang_from_vec(goalTargetPos - pos);
			// deltaAngle = angleToGoal - velAngle  (Angle::operator- -> other.GetDeltaTo(*this))
			Ang deltaAngle; deltaAngle.yaw = angleToGoal.yaw - velAngle.yaw; deltaAngle.pitch = angleToGoal.pitch - velAngle.pitch; deltaAngle.roll = angleToGoal.roll - velAngle.roll;
			ang_normalize_fix(deltaAngle);
			float l2 = len2(vel);
			float curSpeed = (l2 > 0.f) ? sqrtf(l2) : 0.f;
			float speedRatio = curSpeed / K::HS_MAX_SPEED;
			Ang newAngle = velAngle;
			float baseInterpFactor = speedRatio * dt;
			newAngle.yaw += deltaAngle.yaw * baseInterpFactor * K::HS_HORIZONTAL_BLEND;
			newAngle.pitch += deltaAngle.pitch * baseInterpFactor * K::HS_VERTICAL_BLEND;
			ang_normalize_fix(newAngle);
			newAngle.pitch = rs_clamp(newAngle.pitch, -K::HS_MAX_TURN_PITCH, K::HS_MAX_TURN_PITCH);
			newAngle = round_angle_ue3(newAngle);
			float newSpeed = curSpeed + ((b.hsCurTargetSpeed - curSpeed) * K::HS_SPEED_BLEND);
			V3 newDir = ang_forward(newAngle);
			V3 newVel = newDir * newSpeed;
			b.rb.vel = newVel * UU_TO_BT;
			b.hsTimeSinceHit += dt;
		}
	} else if (mode == SNOWDAY) {
		b.groundStickApplied = 0;
	} else if (mode == HOOPS || mode == DROPSHOT) {
		const bool isDropshot = (mode == DROPSHOT);
		float launchDelay = isDropshot ? 0.26f : K::BALL_HOOPS_LAUNCH_DELAY;  // Dropshot::BALL_LAUNCH_DELAY (RLConst.h:209)
		float curKickoffTime = b.tickCountSinceUpdate * dt;
		float prevKickoffTime = curKickoffTime - dt;
		if (prevKickoffTime < launchDelay && curKickoffTime >= launchDelay) {
			V3 pos = b.rb.pos * BT_TO_UU, vel = b.rb.vel * BT_TO_UU;
			if (is_zero(vel) && is_zero(b.rb.angvel) && pos.x == 0.f && pos.y == 0.f) {
				V3 imp = V3(0.f, 0.f, isDropshot ? 985.f : K::BALL_HOOPS_LAUNCH_Z_VEL) * cx.p->ballMass * UU_TO_BT;  // Dropshot::BALL_LAUNCH_Z_VEL
				b.rb.vel += imp * bw.invMass;
				bw.activation = ACT_ACTIVE;
			}
		}
	}
}

// Ball::_OnHit (car-ball contact hook).  relPosUU = contact point local to the ball, in uu.
RSB_HD_NOINLINE void ball_on_hit(Ctx cx, int ci, V3 relPosUU) {
	BallS& b = cx.s->ball;
	CarS& c = cx.s->cars[ci];
	const int mode = cx.mode;
	const uint64_t tickCount = cx.s->tickCount;
	V3 carPos = c.rb.pos * BT_TO_UU, carVel = c.rb.vel * BT_TO_UU;
	V3 ballPos = b.rb.pos * BT_TO_UU, ballVel = b.rb.vel * BT_TO_UU;
	c.bhValid = 1;
	c.bhRelPos = relPosUU;
	c.bhTickHit = tickCount;
	c.bhBallPos = ballPos;
	c.bhExtraVel = v3zero();
	if ((tickCount > c.bhTickExtra + 1) || (c.bhTickExtra > tickCount)) {
		c.bhTickExtra = tickCount;
		V3 carForward = c.rb.basis.col(0);
		V3 relPos = ballPos - carPos;
		V3 relVel = ballVel - carVel;
		float l2 = len2(relVel);
		float rl = (l2 > 0.f) ? sqrtf(l2) : 0.f;
		float relSpeed = (rl < K::BALL_CAR_EXTRA_IMPULSE_MAXDELTAVEL_UU) ? rl : K::BALL_CAR_EXTRA_IMPULSE_MAXDELTAVEL_UU;
		if (relSpeed > 0.f) {
			bool extraZScale = mode == HOOPS && c.isOnGround && c.rb.basis.col(2).z > K::BALL_CAR_EXTRA_IMPULSE_Z_SCALE_HOOPS_NORMAL_Z_THRESH;
			float zScale = extraZScale ? K::BALL_CAR_EXTRA_IMPULSE_Z_SCALE_HOOPS_GROUND : K::BALL_CAR_EXTRA_IMPULSE_Z_SCALE;
			V3 hitDir = rs_normalized(relPos * V3(1.f, 1.f, zScale));
			V3 forwardDirAdjustment = carForward * dot(hitDir, carForward) * (1.f - K::BALL_CAR_EXTRA_IMPULSE_FORWARD_SCALE);
			hitDir = rs_normalized(hitDir - forwardDirAdjustment);
			V3 addedVel = (hitDir * relSpeed) * curve_ball_car_extra_impulse(relSpeed) * cx.p->mut.ballHitExtraForceScale;
			c.bhExtraVel = addedVel;
			b.velImpulseCache += addedVel * UU_TO_BT;
		}
	} else {
		return;
	}
	if (mode == HEATSEEKER) {
		bool canIncrease = (b.hsTimeSinceHit > K::HS_MIN_SPEEDUP_INTERVAL) || (b.hsYTargetDir == 0.f);
		float newTargetDir = (cx.p->team[ci] == 0) ? 1.f : -1.f;
		if (canIncrease && (newTargetDir != b.hsYTargetDir)) {
			b.hsTimeSinceHit = 0.f;
			float t = b.hsCurTargetSpeed + K::HS_TARGET_SPEED_INCREMENT;
			b.hsCurTargetSpeed = (t < K::HS_MAX_SPEED) ? t : K::HS_MAX_SPEED;
		}
		b.hsYTargetDir = newTargetDir;
	} else if (mode == DROPSHOT) {
		// charge the ball with the speed the car drives into it (Ball.cpp:299-323); dsInfo lives in the record's cold tail
		ArenaS& G = *cx.g;
		V3 dirFromCar = rs_normalized(ballPos - carPos);
		V3 relVelFromCar = carVel - ballVel;
		float velIntoBall = dot(dirFromCar, relVelFromCar);
		if (velIntoBall >= 500.f) {  // MIN_CHARGE_HIT_SPEED
			G.dsAccumulatedHitForce += velIntoBall;
			if (G.dsAccumulatedHitForce >= 2500.f) G.dsChargeLevel = 2;   // MIN_ABSORBED_FORCE_FOR_CHARGE
			if (G.dsAccumulatedHitForce >= 11000.f) G.dsChargeLevel = 3;  // MIN_ABSORBED_FORCE_FOR_SUPERCHARGE
		}
		if (G.dsChargeLevel > 1) G.dsYTargetDir = (cx.p->team[ci] == 0) ? 1.f : -1.f;
	}
}

// Ball::_OnWorldCollision (ball-world contact hook)
RSB_HD_NOINLINE void ball_on_world_collision(Ctx cx, V3 normal) {
	BallS& b = cx.s->ball;
	const int mode = cx.mode;
	if (mode == HEATSEEKER) {
		if (b.hsYTargetDir != 0.f) {
			V3 pos = b.rb.pos * BT_TO_UU;
			float relNormalY = normal.y * b.hsYTargetDir;
			float relY = pos.y * b.hsYTargetDir;
			if (relNormalY <= -K::HS_WALL_BOUNCE_CHANGE_Y_NORMAL && relY >= K::ARENA_EXTENT_Y - K::HS_WALL_BOUNCE_CHANGE_Y_THRESH) {
				b.hsYTargetDir *= -1.f;
				V3 vel = b.rb.vel * BT_TO_UU;
				V3 goalTargetPos(0.f, K::HS_TARGET_Y * b.hsYTargetDir, K::HS_TARGET_Z);
				V3 dirToGoal = rs_normalized(goalTargetPos - pos);
				V3 bounceDir = dirToGoal * (1.f - K::HS_WALL_BOUNCE_UP_FRAC) + V3(0.f, 0.f, 1.f) * K::HS_WALL_BOUNCE_UP_FRAC;
				float l2 = len2(vel);
				float vl = (l2 > 0.f) ? sqrtf(l2) : 0.f;
				V3 bounceImpulse = bounceDir * vl * K::HS_WALL_BOUNCE_FORCE_SCALE;
				b.velImpulseCache += bounceImpulse * UU_TO_BT;
			}
		}
	} else if (mode == SNOWDAY) {
		if (!b.groundStickApplied) {
			apply_central_force(cx.body[0], (-normal) * K::PUCK_GROUND_STICK_FORCE);
			b.groundStickApplied = 1;
		}
	}
}

// Ball::_FinishPhysicsTick
RSB_HD void ball_finish_tick(const Ctx& cx) {
	BallS& b = cx.s->ball;
	if (!is_zero(b.velImpulseCache)) {
		b.rb.vel += b.velImpulseCache;
		b.velImpulseCache = v3zero();
	}
	V3 vel = b.rb.vel, angVel = b.rb.angvel;
	float ballMaxSpeedBT = cx.p->mut.ballMaxSpeed * UU_TO_BT;
	if (len2(vel) > ballMaxSpeedBT * ballMaxSpeedBT) vel = normalized(vel) * ballMaxSpeedBT;
	if (len2(angVel) > (K::BALL_MAX_ANG_SPEED * K::BALL_MAX_ANG_SPEED)) angVel = normalized(angVel) * K::BALL_MAX_ANG_SPEED;
	b.rb.vel = vel;
	b.rb.angvel = angVel;
	b.tickCountSinceUpdate++;
}

// Arena::IsBallScored
RSB_HD bool is_ball_scored(const Ctx& cx) {
	const BallS& b = cx.s->ball;
	switch (cx.mode) {
	case SOCCAR: case HEATSEEKER: case SNOWDAY: {
		float ballPosY = b.rb.pos.y * BT_TO_UU;
		return fabsf(ballPosY) > (cx.p->mut.goalBaseThresholdY + cx.p->mut.ballRadius);
	}
	case HOOPS: {
		if (b.rb.pos.z < K::HOOPS_GOAL_SCORE_THRESHOLD_Z * UU_TO_BT) {
			V3 p = b.rb.pos * BT_TO_UU;
			const float SCALE_Y = 0.9f, OFFSET_Y = 2770.f, RADIUS_SQ = 716 * 716;
			float dy = fabsf(p.y) * SCALE_Y - OFFSET_Y;
			float distSq = p.x * p.x + dy * dy;
			return (distSq - RADIUS_SQ) < 0.f;
		}
		return false;
	}
	case DROPSHOT:  // the ball fell through a broken tile (Arena.cpp:916-923)
		return (b.rb.pos.z * BT_TO_UU) < -(cx.p->mut.ballRadius * 1.75f);
	default: return false;
	}
}

// ---- boost pads ----
RSB_HD void pad_pretick(const Ctx& cx, int i) {
	PadS& p = cx.s->pads[i];
	if (p.cooldown > 0.f) { float t = p.cooldown - cx.p->dt; p.cooldown = (t > 0.f) ? t : 0.f; }
	p.isActive = (p.cooldown == 0.f);
	cx.padLocked[i] = -1;
}

// car AABB as btRigidBody::getAabb -> compound shape AABB of the CURRENT transform
// (btCompoundShape.cpp:159-179 with the local AABB of the single box child)
RSB_HD void compound_aabb(const CarCfgD& cfg, const M3& basis, const V3& origin, V3& mn, V3& mx);

RSB_HD void pad_check_collide(const Ctx& cx, int pi, int ci) {
	const PadCfgD& pc = cx.p->pads[pi];
	const CarS& c = cx.s->cars[ci];
	V3 carPosBT = c.rb.pos;
	bool colliding = false;
	if (cx.s->pads[pi].prevLockedCarID == (uint32_t)(ci + 1)) {
		V3 mn, mx;
		compound_aabb(cx.p->carCfg[ci], c.rb.basis, c.rb.pos, mn, mx);
		colliding = (pc.boxMax.x > mn.x && pc.boxMax.y > mn.y && pc.boxMax.z > mn.z) && (pc.boxMin.x < mx.x && pc.boxMin.y < mx.y && pc.boxMin.z < mx.z);
	} else {
		float rad = pc.cylRadBT;
		float dx = carPosBT.x - pc.posBT.x, dy = carPosBT.y - pc.posBT.y;
		if ((dx * dx + dy * dy) < (rad * rad)) colliding = fabsf(carPosBT.z - pc.posBT.z) < (K::PAD_CYL_HEIGHT * UU_TO_BT);
	}
	if (colliding) {  // cars are visited in index order by the reference loop: the highest index wins
#if defined(__CUDA_ARCH__)
		atomicMax(&cx.padLocked[pi], ci);
#else
		if (ci > cx.padLocked[pi]) cx.padLocked[pi] = ci;
#endif
	}
}

// BoostPadGrid::CheckCollision
RSB_HD void pad_grid_check(const Ctx& cx, int ci) {
	const CarS& c = cx.s->cars[ci];
	if (cx.p->customPads) {  // ArenaConfig::useCustomBoostPads: BoostPad::_CheckCollide for every pad, without the grid's early-outs
		for (int pad = 0; pad < cx.p->nPads; pad++) pad_check_collide(cx, pad, ci);
		return;
	}
	if (c.isDemoed || c.boost >= K::BOOST_MAX) return;
	V3 carPos = c.rb.pos * BT_TO_UU;
	const float EXTENT_Z = K::PAD_CYL_HEIGHT + 250.f;
	if (carPos.z > EXTENT_Z) return;
	const int CELLS_X = 8, CELLS_Y = 10, CELL_SIZE_X = 1024, CELL_SIZE_Y = 1024;
	int indexX = (int)(carPos.x / CELL_SIZE_X + (CELLS_X / 2));
	int indexY = (int)(carPos.y / CELL_SIZE_Y + (CELLS_Y / 2));
	int i0 = indexX - 1 > 0 ? indexX - 1 : 0, i1 = indexX + 1 < CELLS_X - 1 ? indexX + 1 : CELLS_X - 1;
	int j0 = indexY - 1 > 0 ? indexY - 1 : 0, j1 = indexY + 1 < CELLS_Y - 1 ? indexY + 1 : CELLS_Y - 1;
	for (int i = i0; i <= i1; i++)
		for (int j = j0; j <= j1; j++) {
			int pad = cx.p->padGrid[i * CELLS_Y + j];
			if (pad >= 0) pad_check_collide(cx, pad, ci);
		}
}

RSB_HD void pad_posttick(const Ctx& cx, int i) {
	PadS& p = cx.s->pads[i];
	const PadCfgD& pc = cx.p->pads[i];
	uint32_t lockedCarID = 0;
	int lc = cx.padLocked[i];
	if (lc >= 0) {
		lockedCarID = (uint32_t)(lc + 1);
		if (p.isActive) {
			float boostToAdd = pc.isBig ? K::PAD_BOOST_AMOUNT_BIG : K::PAD_BOOST_AMOUNT_SMALL;
			CarS& c = cx.s->cars[lc];
#if defined(__CUDA_ARCH__)
			// pads run one per thread: two pads feeding the same car in one tick must not lose an update
			unsigned int* addr = (unsigned int*)&c.boost;
			unsigned int old = *addr, assumed;
			do {
				assumed = old;
				float nb = __uint_as_float(assumed) + boostToAdd;
				nb = (nb < K::BOOST_MAX) ? nb : K::BOOST_MAX;
				old = atomicCAS(addr, assumed, __float_as_uint(nb));
			} while (old != assumed);
#else
			float nb = c.boost + boostToAdd;
			c.boost = (nb < K::BOOST_MAX) ? nb : K::BOOST_MAX;
#endif
			p.isActive = 0;
			p.cooldown = pc.isBig ? cx.p->mut.boostPadCooldown_Big : cx.p->mut.boostPadCooldown_Small;
		}
	}
	p.prevLockedCarID = lockedCarID;
}

}  // namespace rsb
