// rsb_tick.h -- one arena tick = Arena::Step body (reference src/Sim/Arena/Arena.cpp:689-756)
// + btDiscreteDynamicsWorld::stepSimulation / internalSingleStepSimulation
// (bt/BulletDynamics/Dynamics/btDiscreteDynamicsWorld.cpp:325-437, 889-991, 1029-1044),
// written as SPMD phases over the lanes of the warp that owns the arena.
#pragma once
#include "rsb_work.h"
#include "rsb_ray.h"
#include "rsb_car.h"
#include "rsb_ball.h"
#include "rsb_collide.h"
#include "rsb_solver.h"
#include "rsb_kickoff.h"

namespace rsb {

RSB_HD void init_body_work(const Ctx& cx, int b) {
	BodyW& bw = cx.w->body[b];
	const BatchParams& P = *cx.p;
	const RigidS& rb = rigid_of(cx, b);
	if (b == 0) {
		bw.invMass = P.ballInvMass; bw.invInertiaLocal = P.ballInvInertiaLocal;
		bw.friction = P.mut.ballWorldFriction; bw.restitution = P.mut.ballWorldRestitution;
		// Arena::Step ball sleep gate (Arena.cpp:694-700)
		bw.activation = (len2(rb.vel) == 0.f && len2(rb.angvel) == 0.f) ? ACT_SLEEPING : ACT_ACTIVE;
		bw.noResponse = 0;
	} else {
		bw.invMass = 1.f / K::CAR_MASS_BT; bw.invInertiaLocal = P.carCfg[b - 1].invInertiaLocal;
		bw.friction = K::CAR_COLLISION_FRICTION; bw.restitution = K::CAR_COLLISION_RESTITUTION;
		bw.activation = ACT_ACTIVE; bw.noResponse = 0;
	}
	bw.invInertiaW = world_inv_inertia(rb.basis, bw.invInertiaLocal);
	bw.totalForce = v3zero(); bw.totalTorque = v3zero();
	bw.islandTag = b;
}

RSB_HD void arena_tick(const Ctx& cx) {
	const BatchParams& P = *cx.p;
	ArenaS& S = *cx.s;
	ArenaW& W = *cx.w;
	const int nCars = P.nCars, nb = nCars + 1;
	const float dt = P.dt;
	const bool hasArenaStuff = (P.mode != THE_VOID);
	const bool ballOnly = (nCars == 0);

	// ---- T0: working rigid-body values + ball sleep gate
	RSB_FOR_LANES(b, nb) init_body_work(cx, b);
	RSB_SYNC();

	// ---- T1: cars pre-tick (Car::_PreTickUpdate)
	if (cx.lane == 0) {  // sequential: respawns draw from the arena RNG in car order
		for (int ci = 0; ci < nCars; ci++) car_pretick_begin(cx, ci);
	}
	RSB_SYNC();
	RSB_FOR_LANES(k, nCars * 4) {
		int ci = k >> 2, wi = k & 3;
		if (!S.cars[ci].isDemoed) wheel_trace(cx, ci, wi);
	}
	RSB_SYNC();
	RSB_FOR_LANES(ci, nCars) {
		if (!S.cars[ci].isDemoed) car_pretick_logic(cx, ci);
	}
	RSB_SYNC();

	// ---- T2: pads + ball pre-tick
	if (hasArenaStuff && !ballOnly) { RSB_FOR_LANES(i, P.nPads) pad_pretick(cx, i); }
	if (cx.lane == 0) ball_pretick(cx);
	RSB_SYNC();

	// ---- T3: world step
	// applyGravity (active bodies) + predictUnconstraintMotion (damping + predicted transform)
	RSB_FOR_LANES(b, nb) {
		BodyW& bw = W.body[b];
		RigidS& rb = rigid_of(cx, b);
		if (body_active(bw)) apply_central_force(bw, P.gravityBT * (1.f / bw.invMass));
		float linDamp = (b == 0) ? P.mut.ballDrag : 0.f;
		rb.vel *= powf(1.f - linDamp, dt);
		rb.angvel *= powf(1.f - 0.f, dt);
		Xform cur; cur.basis = rb.basis; cur.origin = rb.pos;
		integrate_transform(cur, rb.vel, rb.angvel, dt, (b == 0) && P.ballNoRot, bw.interp);
	}
	RSB_SYNC();
	if (cx.lane == 0) {
		collide_world(cx);                    // updateAabbs, pairs, narrowphase + contact hooks
		int nSolve = build_islands(cx);       // calculateSimulationIslands + buildIslands
		solve_arena(cx, nSolve);              // solveConstraints
	}
	RSB_SYNC();
	// integrateTransforms (active bodies) + clearForces
	RSB_FOR_LANES(b, nb) {
		BodyW& bw = W.body[b];
		RigidS& rb = rigid_of(cx, b);
		if (body_active(bw)) {
			Xform cur; cur.basis = rb.basis; cur.origin = rb.pos;
			Xform pred;
			integrate_transform(cur, rb.vel, rb.angvel, dt, (b == 0) && P.ballNoRot, pred);
			rb.pos = pred.origin; rb.basis = pred.basis;
		}
	}
	RSB_SYNC();

	// ---- T4: cars post-tick + pad checks
	RSB_FOR_LANES(ci, nCars) car_posttick(cx, ci);
	RSB_SYNC();
	if (cx.lane == 0) {
		if (hasArenaStuff) for (int ci = 0; ci < nCars; ci++) pad_grid_check(cx, ci);
		// ---- T5: pads post-tick, ball finish
		if (hasArenaStuff && !ballOnly) for (int i = 0; i < P.nPads; i++) pad_posttick(cx, i);
		ball_finish_tick(cx);
		// ---- T6: goal event, tick counter
		if (P.mode != THE_VOID && is_ball_scored(cx)) {
			float y = -S.ball.rb.pos.y;
			push_event(cx, 1u /*RSB_EVENT_GOAL*/, (y < 0.f) ? 0u : 1u, 0u);
			if (P.autoReset) kickoff_reset(S, P, (int)(P.autoResetSeed + (uint32_t)cx.arena + (uint32_t)S.tickCount));
		}
		S.tickCount++;
	}
	RSB_SYNC();
}

}  // namespace rsb
