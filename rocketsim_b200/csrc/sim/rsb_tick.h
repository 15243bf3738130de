// rsb_tick.h -- one arena tick = Arena::Step body (reference src/Sim/Arena/Arena.cpp:689-756)
// + btDiscreteDynamicsWorld::stepSimulation / internalSingleStepSimulation
// (bt/BulletDynamics/Dynamics/btDiscreteDynamicsWorld.cpp:325-437, 889-991, 1029-1044),
// written as a sequence of item-parallel phases (see Phase below).
#pragma once
#include "rsb_work.h"
#include "rsb_ray.h"
#include "rsb_car.h"
#include "rsb_ball.h"
#include "rsb_collide.h"
#include "rsb_solver.h"
#include "rsb_kickoff.h"
#include "rsb_tracker.h"

namespace rsb {

RSB_HD void init_body_work(const Ctx& cx, int b) {
	BodyW& bw = cx.body[b];
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

// ---------------------------------------------------------------- phases of one tick
// A tick is a fixed sequence of PHASES; every phase is a flat list of independent ITEMS
// (arena x wheel, arena x car, arena x body, arena) that any thread may execute.  The CUDA
// kernel (rsb_api.cu) keeps a chunk of arenas resident in shared memory, deals the items of a
// phase over all threads of the block and separates phases with __syncthreads(); the host
// check build runs the same phases with one thread.  Order = Arena::Step T0..T6.
enum Phase : int {
	PH_INIT = 0,       // T0  working rigid-body values + ball sleep gate                 item = body
	PH_PRE_BEGIN,      // T1  Car::_PreTickUpdate head (respawn RNG, sequential)           item = arena
	PH_WHEELS,         // T1  suspension rays                                              item = car x wheel
	PH_CAR_LOGIC,      // T1  friction, game logic, suspension/boost forces; T2 pads+ball  item = car | pad | ball
	PH_PREDICT,        // T3  gravity, damping, predicted transform, swept AABB            item = body
	PH_PAIRS,          // T3  broadphase cell bookkeeping + pair list                      item = arena
	PH_NARROW_STATIC,  // T3  pair walk vs static objects (sphere ball: contacts on the spot;   item = body | arena
	                   //     box / hull: candidate records) + dynamic-pair candidates
	PH_NARROW_DETECT,  // T3  GJK(+EPA) / box-box of one pending candidate record           item = queue entry
	PH_NARROW_COMMIT,  // T3  candidate records of static pairs -> manifolds + hooks        item = body
	PH_DYN_SOLVE,      // T3  ball-car / car-car records -> manifolds + hooks, islands, solver item = arena
	PH_FINISH_BODY,    // T3  contact-free solver write-back, integrateTransforms          item = body
	PH_CAR_POST,       // T4  Car::_PostTickUpdate / _FinishPhysicsTick, pad grid check    item = car
	PH_PADS_POST,      // T5  pads give boost / start cooldown                             item = pad
	PH_TAIL,           // T5-T6 ball finish, goal (+ kickoff auto reset), tick counter     item = arena
	PH_COUNT
};

RSB_HD int phase_items(const BatchParams& P, int phase) {
	const int nCars = P.nCars, nb = nCars + 1;
	const int nPadItems = (P.mode != THE_VOID && nCars != 0) ? P.nPads : 0;
	switch (phase) {
	case PH_INIT: case PH_PREDICT: case PH_NARROW_COMMIT: case PH_FINISH_BODY: return nb;
	case PH_NARROW_STATIC: return nb + ((nCars > 0 || P.mode == DROPSHOT) ? 1 : 0);  // + the arena's dynamic (and Dropshot tile) pairs
	case PH_NARROW_DETECT: return 0;                         // variable: the block's detector queue (dealt by the caller)
	case PH_WHEELS: return nCars * 4;
	case PH_CAR_LOGIC: return nCars + nPadItems + 1;
	case PH_CAR_POST: return nCars;
	case PH_PADS_POST: return nPadItems;
	default: return 1;
	}
}

RSB_HD void phase_run(const Ctx& cx, int phase, int item) {
	const BatchParams& P = *cx.p;
	ArenaS& S = *cx.s;
	const int nCars = P.nCars;
	const float dt = P.dt;
	const bool hasArenaStuff = (cx.mode != THE_VOID);
	const bool ballOnly = (nCars == 0);
	switch (phase) {
	case PH_INIT: init_body_work(cx, item); break;
	case PH_PRE_BEGIN:  // sequential: respawns draw from the arena RNG in car order
		for (int ci = 0; ci < nCars; ci++) car_pretick_begin(cx, ci);
		break;
	case PH_WHEELS: {
		int ci = item >> 2, wi = item & 3;
		if (!S.cars[ci].isDemoed) { wheel_trace(cx, ci, wi); calc_friction_impulse(cx, ci, wi); }
		break;
	}
	case PH_CAR_LOGIC: {
		// cars, pads and the ball do not read each other's pre-tick results: one phase
		const int nPadItems = (hasArenaStuff && !ballOnly) ? P.nPads : 0;
		if (item < nCars) { if (!S.cars[item].isDemoed) car_pretick_logic(cx, item); }
		else if (item < nCars + nPadItems) pad_pretick(cx, item - nCars);
		else ball_pretick(cx);
		break;
	}
	case PH_PREDICT: {
		// applyGravity (active bodies) + predictUnconstraintMotion (damping + predicted transform) + updateSingleAabb
		const int b = item;
		BodyW& bw = cx.body[b];
		RigidS& rb = rigid_of(cx, b);
		if (body_active(bw)) apply_central_force(bw, P.gravityBT * (1.f / bw.invMass));
		// applyDamping: vel *= powf(1 - linDamping, dt), angvel *= powf(1 - 0, dt) = 1 (btRigidBody.cpp:143-147); cars have no damping
		rb.vel *= (b == 0) ? P.ballLinDampFactor : 1.f;
		rb.angvel *= 1.f;
		Xform cur; cur.basis = rb.basis; cur.origin = rb.pos;
		integrate_transform(cur, rb.vel, rb.angvel, dt, (b == 0) && P.ballNoRot, bw.interp);
		update_aabb(cx, b);
		break;
	}
	case PH_PAIRS: collide_pairs(cx); break;
	case PH_NARROW_STATIC:
		if (item <= nCars) collide_static_pairs(cx, item);
		else enqueue_dynamic_pairs(cx);
		break;
	case PH_NARROW_DETECT: detect_candidate(cx, item); break;
	case PH_NARROW_COMMIT:
		if (!(RSB_SPHERE_INLINE && item == 0 && cx.ballIsSphere)) commit_candidates(cx, item);
		break;
	case PH_DYN_SOLVE: {
		RSB_SUBPROF_BEGIN();
		commit_candidates(cx, -1);
		RSB_SUBPROF_MARK(0);
		const int nSolve = build_islands(cx);  // calculateSimulationIslands + buildIslands
		RSB_SUBPROF_MARK(1);
		const bool rows = solver_has_rows(cx, nSolve);
		cx.misc[0] = rows ? 1 : 0; cx.misc[1] = nSolve; cx.misc[2] = 0;
		if (rows) solve_arena(cx, nSolve);     // solveConstraints
		RSB_SUBPROF_MARK(rows ? 7 : 1);
		break;
	}
	case PH_FINISH_BODY: {
		const int b = item;
		if (!((cx.misc[2] >> b) & 1)) solve_body_no_contacts(cx, b);  // bodies the contact solver did not write back
		// integrateTransforms (active bodies) + clearForces
		BodyW& bw = cx.body[b];
		RigidS& rb = rigid_of(cx, b);
		if (body_active(bw)) {
			Xform cur; cur.basis = rb.basis; cur.origin = rb.pos;
			Xform pred;
			integrate_transform(cur, rb.vel, rb.angvel, dt, (b == 0) && P.ballNoRot, pred);
			rb.pos = pred.origin; rb.basis = pred.basis;
		}
		break;
	}
	case PH_CAR_POST:
		car_posttick(cx, item);
		if (hasArenaStuff) pad_grid_check(cx, item);
		break;
	case PH_PADS_POST: pad_posttick(cx, item); break;
	case PH_TAIL:
		ball_finish_tick(cx);
		// Dropshot: the rigid bodies of tiles broken this tick lose their contact response now (Arena.cpp:741-746; a damage on
		// tick 0 is never synchronised there -- `lastDamageTick &&` -- and therefore not here either)
		if (cx.mode == DROPSHOT && cx.g->dsLastDamageTick && cx.g->dsLastDamageTick == S.tickCount) dropshot_sync_tiles(*cx.g);
		// ---- T6: goal event, tick counter
		{
			const bool scored = cx.mode != THE_VOID && is_ball_scored(cx);
			if (scored) {
				float y = -S.ball.rb.pos.y;
				push_event(cx, 1u /*RSB_EVENT_GOAL*/, (y < 0.f) ? 0u : 1u, 0u);
			}
			// the caller's loop is Step(); tracker.Update(); if (scored) Reset -- same order here
			if (P.trackerEnabled) tracker_update(cx, scored, S.tickCount + 1);
			// seed = f(global arena index, tick), masked to 31 bits: never negative, never the -1 that selects the arena RNG
			if (scored && P.autoReset) kickoff_reset(S, *cx.g, P, (int)((P.autoResetSeed + (uint32_t)(P.arenaOffset + cx.arena) + (uint32_t)S.tickCount) & 0x7fffffffu));
		}
		S.tickCount++;
		S.worldFresh = 0;
		break;
	default: break;
	}
}

// One whole tick of one arena on one thread (host check build, single-arena helpers).
RSB_HD void arena_tick(const Ctx& cx) {
	*cx.nq = 0;
	for (int ph = 0; ph < PH_COUNT; ph++) {
		if (ph == PH_NARROW_DETECT) {
			const unsigned nq = *cx.nq < (unsigned)cx.qcap ? *cx.nq : (unsigned)cx.qcap;
			for (unsigned i = 0; i < nq; i++) phase_run(cx, ph, cx.q[i] & 127);
			if (*cx.nq > (unsigned)cx.qcap)  // queue overflow: the records that are still pending
				for (int i = 0; i < cx.misc[3] && i < MAX_CANDS; i++) if (cx.w->cand[i].state == CAND_PENDING) phase_run(cx, ph, i);
			continue;
		}
		const int n = phase_items(*cx.p, ph);
		for (int i = 0; i < n; i++) phase_run(cx, ph, i);
	}
}

}  // namespace rsb
