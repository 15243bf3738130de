// rsb_solver.h -- islands, sequential-impulse (PGS) contact solve and velocity writeback for
// one arena (<= 7 dynamic bodies, all islands solved in one group).  Restates
//   island building / ordering   bt/BulletCollision/CollisionDispatch/btSimulationIslandManager.cpp:40-64,196-445
//                                bt/BulletCollision/CollisionDispatch/btUnionFind.h (path-compression variant)
//   btAlignedObjectArray::quickSort (non-stable)  bt/LinearMath/btAlignedObjectArray.h:311-348
//   solver setup/rows/iterations/writeback        bt/BulletDynamics/ConstraintSolver/btSequentialImpulseConstraintSolver.cpp
//                                                 :458-491, 517-607, 795-983, 1003-1211, 1467-1599, 1601-1829, 1878-1925
//   btSolverBody                                  bt/BulletDynamics/ConstraintSolver/btSolverBody.h:105-288
// including the RocketSim patches: "special" ball-world contacts are skipped in the velocity
// iterations and replaced by one averaged contact; no speculative penetration term; erp2 0.8;
// split impulse always on (threshold 1e30) (src/Sim/Arena/Arena.cpp:472-475).
#pragma once
#include "rsb_work.h"

namespace rsb {

// btAlignedObjectArray::quickSortInternal on (key,payload) pairs; comparator key<key
RSB_HD void quicksort_keys(int* key, int* val, int lo, int hi) {
	// explicit stack (max depth tiny: arrays <= MAX_MANIFOLDS)
	int stackLo[32], stackHi[32]; int sp = 0;
	stackLo[0] = lo; stackHi[0] = hi; sp = 1;
	while (sp > 0) {
		sp--; lo = stackLo[sp]; hi = stackHi[sp];
		int i = lo, j = hi;
		int x = key[(lo + hi) / 2];
		do {
			while (key[i] < x) i++;
			while (x < key[j]) j--;
			if (i <= j) {
				int tk = key[i]; key[i] = key[j]; key[j] = tk;
				int tv = val[i]; val[i] = val[j]; val[j] = tv;
				i++; j--;
			}
		} while (i <= j);
		// the reference recurses (lo,j) first, then (i,hi); the two ranges are disjoint so order is irrelevant
		if (i < hi && sp < 32) { stackLo[sp] = i; stackHi[sp] = hi; sp++; }
		if (lo < j && sp < 32) { stackLo[sp] = lo; stackHi[sp] = j; sp++; }
	}
}


// calculateSimulationIslands + buildIslands: island tags, wake-ups, and the manifold order
// handed to the solver.  Returns the number of manifolds to solve (order in pl->manifoldOrder).
// <= 7 bodies: the union-find forest lives in one register (4 bits per body), activation states and response flags are
// gathered into bit fields once, and only the manifolds that exist this tick (PairList::mfMask) are visited; the unite
// sequence, the roots it produces and the (non-stable) manifold sort are the reference's.
RSB_HD int build_islands(const Ctx& cx) {
	const int nb = n_bodies(cx);
	uint32_t par = 0x76543210u;  // par[b] = b
	uint32_t noResp = 0, acts = 0;  // bit b / 4 bits per body
	for (int b = 0; b < nb; b++) {
		noResp |= (cx.body[b].noResponse ? 1u : 0u) << b;
		acts |= (uint32_t)cx.body[b].activation << (4 * b);
	}
	// findUnions over ALL broadphase pairs whose bodies both merge islands (dynamic pairs sit behind each body's static pairs)
	for (uint32_t dm = cx.pl->dynMask; dm; dm &= dm - 1) {
#if defined(__CUDA_ARCH__)
		const int k = __ffs((int)dm) - 1;
#else
		const int k = __builtin_ctz(dm);
#endif
		const int a = cx.pl->a[k];
		const int b = cx.pl->b[k];
		if (((noResp >> a) | (noResp >> b)) & 1u) continue;
		int i = a, j = b;
		for (;;) { const int q = (par >> (4 * i)) & 15; if (q == i) break; i = q; }
		for (;;) { const int q = (par >> (4 * j)) & 15; if (q == j) break; j = q; }
		if (i != j) par = (par & ~(15u << (4 * i))) | ((uint32_t)j << (4 * i));  // id[root(a)] = root(b)
	}
	uint32_t tags = 0, rootsActive = 0;
	if (par == 0x76543210u) {  // no union: every body is its own island
		// Common case, short cut.  A one-body island changes no activation state (WANTS_DEACTIVATION only arises from a union
		// with a sleeper), islandTag[b] = b was set by init_body_work, and when at most ONE eligible manifold carries contact
		// points the order the sort would produce is irrelevant: the solver skips point-less manifolds.
		uint32_t notSl = 0, awake = 0;
		for (int b = 0; b < nb; b++) {
			const int act = (acts >> (4 * b)) & 15;
			if (act != ACT_SLEEPING) notSl |= 1u << b;
			if (act != ACT_SLEEPING && act != ACT_DISABLE_SIM) awake |= 1u << b;
		}
		int withPts = 0, last = -1;
		uint32_t mm = cx.pl->mfMask;
		while (mm) {
#if defined(__CUDA_ARCH__)
			const int m = __ffs((int)mm) - 1;
#else
			const int m = __builtin_ctz(mm);
#endif
			mm &= mm - 1;
			if (cx.mf[m].nPts <= 0) continue;
			const int b0 = cx.pl->b[m], b1 = cx.pl->a[m];
			const bool s1 = b1 >= STATIC_BASE;
			if (!(((notSl >> b0) & 1) || (!s1 && ((notSl >> b1) & 1)))) continue;
			if (((noResp >> b0) & 1) || (!s1 && ((noResp >> b1) & 1))) continue;
			if (s1 && static_no_response(cx, b1)) continue;  // broken Dropshot tile
			withPts++; last = m;
		}
		if (withPts == 0) return 0;
		if (withPts == 1) {
			if (!((awake >> cx.pl->b[last]) & 1)) return 0;
			cx.pl->manifoldOrder[0] = (uint8_t)last;
			return 1;
		}
		tags = par;
		for (int b = 0; b < nb; b++) {
			const int act = (acts >> (4 * b)) & 15;
			if (act == ACT_ACTIVE || act == ACT_DISABLE_DEACT) rootsActive |= 1u << b;
		}
	} else {
		for (int b = 0; b < nb; b++) {
			int r = b;
			for (;;) { const int q = (par >> (4 * r)) & 15; if (q == r) break; r = q; }
			tags |= (uint32_t)r << (4 * b);
			const int act = (acts >> (4 * b)) & 15;
			if (act == ACT_ACTIVE || act == ACT_DISABLE_DEACT) rootsActive |= 1u << r;
		}
	}
	// per island: all sleeping? else wake sleepers (ISLAND_SLEEPING -> WANTS_DEACTIVATION)
	uint32_t rootsAwake = 0, notSleeping = 0;
	for (int b = 0; b < nb; b++) {
		const int r = (tags >> (4 * b)) & 15;
		const int act0 = (acts >> (4 * b)) & 15;
		int act = act0;
		if ((rootsActive >> r) & 1) { if (act == ACT_SLEEPING) act = ACT_WANTS_DEACT; }
		else if (act != ACT_DISABLE_DEACT && act != ACT_DISABLE_SIM && act != ACT_SLEEPING) act = ACT_SLEEPING;
		if (act != act0) cx.body[b].activation = act;
		cx.body[b].islandTag = r;
		if (act != ACT_SLEEPING) notSleeping |= 1u << b;
		if (act != ACT_SLEEPING && act != ACT_DISABLE_SIM) rootsAwake |= 1u << r;  // body_active
	}
	// m_islandmanifold: manifolds with an awake body and needsResponse, then quickSort by island id
	int keys[MAX_MANIFOLDS], order[MAX_MANIFOLDS];
	int n = 0;
	uint32_t mm = cx.pl->mfMask;
	while (mm) {
#if defined(__CUDA_ARCH__)
		const int m = __ffs((int)mm) - 1;
#else
		const int m = __builtin_ctz(mm);
#endif
		mm &= mm - 1;
		// manifold of pair m: body0 is always dynamic (= pl->b[m]); body1 = pl->a[m], static for a static pair
		const int b0 = cx.pl->b[m], b1 = cx.pl->a[m];
		const bool s1 = b1 >= STATIC_BASE;
		const bool awake0 = (notSleeping >> b0) & 1;
		const bool awake1 = !s1 && ((notSleeping >> b1) & 1);
		if (!(awake0 || awake1)) continue;
		const bool resp = !((noResp >> b0) & 1) && (s1 || !((noResp >> b1) & 1));
		if (!resp) continue;
		if (s1 && static_no_response(cx, b1)) continue;  // broken Dropshot tile
		keys[n] = (int)((tags >> (4 * b0)) & 15);
		order[n] = m;
		n++;
	}
	if (n > 1) quicksort_keys(keys, order, 0, n - 1);
	// drop manifolds whose island is asleep (processIslands only calls the solver for awake islands)
	int outN = 0;
	for (int k = 0; k < n; k++)
		if ((rootsAwake >> keys[k]) & 1) cx.pl->manifoldOrder[outN++] = (uint8_t)order[k];
	return outN;
}

RSB_HD V3 sb_vel_no_delta(const SolverBody& b, const V3& rel) {  // getVelocityInLocalPointNoDelta
	if (b.orig < 0) return v3zero();
	return b.lin + b.extF + cross(b.ang + b.extT, rel);
}
RSB_HD void sb_apply_impulse(SolverBody& b, const V3& lin, const V3& ang, float mag) {
	if (b.orig >= 0) { b.dLin += lin * mag; b.dAng += ang * mag; }
}
// same, reporting whether the body's delta velocities changed at all (fixed-point detection of the PGS sweep)
RSB_HD bool sb_apply_impulse_chg(SolverBody& b, const V3& lin, const V3& ang, float mag) {
	if (b.orig < 0) return false;
	const V3 l0 = b.dLin, a0 = b.dAng;
	b.dLin += lin * mag; b.dAng += ang * mag;
	return (b.dLin.x != l0.x) | (b.dLin.y != l0.y) | (b.dLin.z != l0.z) | (b.dAng.x != a0.x) | (b.dAng.y != a0.y) | (b.dAng.z != a0.z);
}
RSB_HD void sb_apply_push(SolverBody& b, const V3& lin, const V3& ang, float mag) {
	if (b.orig >= 0) { b.push += lin * mag; b.turn += ang * mag; }
}
RSB_HD float simd_dot3(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }  // btSimdDot3 association

// setupContactConstraint (btSequentialImpulseConstraintSolver.cpp:795-983)
RSB_HD void setup_contact_row(const Ctx& cx, SolverRow& r, int ia, int ib, const V3& normalB, float distance, float restitutionCoef,
                              float frictionCoef, int isSpecial, const V3& rel1, const V3& rel2) {
	const SolverView sv = solver_view(cx);
	SolverBody& A = sv.sbody[ia];
	SolverBody& B = sv.sbody[ib];
	const bool hasA = A.orig >= 0, hasB = B.orig >= 0;
	const float dt = cx.p->dt;
	const float invTimeStep = 1.f / dt;
	const float erp = 0.8f;
	float cfm = 0.f; cfm *= invTimeStep;
	r.bodyA = ia; r.bodyB = ib;
	V3 torqueAxis0 = cross(rel1, normalB);
	r.angA = hasA ? (cx.body[A.orig].invInertiaW * torqueAxis0) : v3zero();
	V3 torqueAxis1 = cross(rel2, normalB);
	r.angB = hasB ? (cx.body[B.orig].invInertiaW * (-torqueAxis1)) : v3zero();
	float denom0 = 0.f, denom1 = 0.f;
	if (hasA) { V3 vec = cross(r.angA, rel1); denom0 = cx.body[A.orig].invMass + dot(normalB, vec); }
	if (hasB) { V3 vec = cross(-r.angB, rel2); denom1 = cx.body[B.orig].invMass + dot(normalB, vec); }
	r.jacDiagInv = 1.f / (denom0 + denom1 + cfm);
	if (hasA) { r.n1 = normalB; r.rp1xn = torqueAxis0; } else { r.n1 = v3zero(); r.rp1xn = v3zero(); }
	if (hasB) { r.n2 = -normalB; r.rp2xn = -torqueAxis1; } else { r.n2 = v3zero(); r.rp2xn = v3zero(); }
	float restitution = 0.f;
	float penetration = distance + 0.f;
	{
		V3 vel1 = hasA ? vel_at(rigid_of(cx, A.orig), rel1) : v3zero();
		V3 vel2 = hasB ? vel_at(rigid_of(cx, B.orig), rel2) : v3zero();
		V3 vel = vel1 - vel2;
		float rel_vel = dot(normalB, vel);
		r.friction = frictionCoef;
		// restitutionCurve with velocity threshold 0.2
		restitution = (fabsf(rel_vel) < 0.2f) ? 0.f : (restitutionCoef * -rel_vel);
		if (restitution <= 0.f) restitution = 0.f;
	}
	r.applied = 0.f;  // warm start of a fresh point: appliedImpulse 0 * 0.85
	r.appliedPush = 0.f;
	{
		V3 eFA = hasA ? A.extF : v3zero(), eTA = hasA ? A.extT : v3zero();
		V3 eFB = hasB ? B.extF : v3zero(), eTB = hasB ? B.extT : v3zero();
		float vel1Dotn = dot(r.n1, A.lin + eFA) + dot(r.rp1xn, A.ang + eTA);
		float vel2Dotn = dot(r.n2, B.lin + eFB) + dot(r.rp2xn, B.ang + eTB);
		float rel_vel = vel1Dotn + vel2Dotn;
		float positionalError = 0.f;
		float velocityError = restitution - rel_vel;
		if (penetration > 0.f) positionalError = 0.f;
		else positionalError = -penetration * erp * invTimeStep;
		float penetrationImpulse = positionalError * r.jacDiagInv;
		float velocityImpulse = velocityError * r.jacDiagInv;
		// split impulse: penetration (<= 1e30 threshold) always goes to rhsPenetration
		r.rhs = velocityImpulse;
		r.rhsPen = penetrationImpulse;
		r.cfm = cfm * r.jacDiagInv;
		r.lower = 0.f; r.upper = 1e10f;
	}
	r.isSpecial = isSpecial;
}

// setupFrictionConstraint (btSequentialImpulseConstraintSolver.cpp:517-607)
RSB_HD void setup_friction_row(const Ctx& cx, SolverRow& r, const V3& axis, int ia, int ib, float frictionCoef, const V3& rel1, const V3& rel2, int frictionIndex) {
	const SolverView sv = solver_view(cx);
	SolverBody& A = sv.sbody[ia];
	SolverBody& B = sv.sbody[ib];
	const bool hasA = A.orig >= 0, hasB = B.orig >= 0;
	r.bodyA = ia; r.bodyB = ib; r.frictionIndex = frictionIndex;
	r.friction = frictionCoef;
	r.applied = 0.f; r.appliedPush = 0.f;
	if (hasA) {
		r.n1 = axis; V3 ft = cross(rel1, r.n1); r.rp1xn = ft; r.angA = cx.body[A.orig].invInertiaW * ft;
	} else { r.n1 = v3zero(); r.rp1xn = v3zero(); r.angA = v3zero(); }
	if (hasB) {
		r.n2 = -axis; V3 ft = cross(rel2, r.n2); r.rp2xn = ft; r.angB = cx.body[B.orig].invInertiaW * ft;
	} else { r.n2 = v3zero(); r.rp2xn = v3zero(); r.angB = v3zero(); }
	float denom0 = 0.f, denom1 = 0.f;
	if (hasA) { V3 vec = cross(r.angA, rel1); denom0 = cx.body[A.orig].invMass + dot(axis, vec); }
	if (hasB) { V3 vec = cross(-r.angB, rel2); denom1 = cx.body[B.orig].invMass + dot(axis, vec); }
	r.jacDiagInv = 1.f / (denom0 + denom1);
	float vel1Dotn = dot(r.n1, hasA ? A.lin + A.extF : v3zero()) + dot(r.rp1xn, hasA ? A.ang : v3zero());
	float vel2Dotn = dot(r.n2, hasB ? B.lin + B.extF : v3zero()) + dot(r.rp2xn, hasB ? B.ang : v3zero());
	float rel_vel = vel1Dotn + vel2Dotn;
	float velocityError = 0.f - rel_vel;
	float velocityImpulse = velocityError * r.jacDiagInv;
	r.rhs = 0.f + velocityImpulse;
	r.rhsPen = 0.f;
	r.cfm = 0.f;
	r.lower = -r.friction; r.upper = r.friction;
	r.isSpecial = 0;
}

// convertContactInner: friction direction + friction row (btSequentialImpulseConstraintSolver.cpp:1069-1162)
RSB_HD void add_friction_for(const Ctx& cx, int ia, int ib, const V3& normalB, float frictionCoef, const V3& rel1, const V3& rel2, int frictionIndex) {
	const SolverView sv = solver_view(cx);
	V3 vel1 = sb_vel_no_delta(sv.sbody[ia], rel1);
	V3 vel2 = sb_vel_no_delta(sv.sbody[ib], rel2);
	V3 vel = vel1 - vel2;
	float rel_vel = dot(normalB, vel);
	V3 dir1 = vel - normalB * rel_vel;
	float lat_rel_vel = len2(dir1);
	if (lat_rel_vel > kEps) {
		dir1 *= 1.f / sqrtf(lat_rel_vel);
	} else {
		V3 dir2;
		plane_space1(normalB, dir1, dir2);
	}
	setup_friction_row(cx, sv.frow(sv.cnt[2]), dir1, ia, ib, frictionCoef, rel1, rel2, frictionIndex);
	sv.cnt[2]++;
}

RSB_HD int get_or_init_solver_body(const Ctx& cx, int obj, int* companion, int& fixedId) {
	const SolverView sv = solver_view(cx);
	if (obj < STATIC_BASE) {
		if (companion[obj] >= 0) return companion[obj];
		int id = sv.cnt[0]++;
		SolverBody& sb = sv.sbody[id];
		const RigidS& rb = rigid_of(cx, obj);
		const BodyW& bw = cx.body[obj];
		sb.dLin = sb.dAng = sb.push = sb.turn = v3zero();
		sb.xf.basis = rb.basis; sb.xf.origin = rb.pos;
		sb.invMassV = V3(bw.invMass, bw.invMass, bw.invMass);
		sb.orig = obj;
		sb.lin = rb.vel; sb.ang = rb.angvel;
		sb.extF = bw.totalForce * bw.invMass * cx.p->dt;
		sb.extT = vmul(bw.totalTorque, bw.invInertiaW) * cx.p->dt;
		companion[obj] = id;
		return id;
	}
	if (fixedId < 0) {
		fixedId = sv.cnt[0]++;
		SolverBody& sb = sv.sbody[fixedId];
		sb.dLin = sb.dAng = sb.push = sb.turn = v3zero();
		sb.xf.basis = m3identity(); sb.xf.origin = v3zero();
		sb.invMassV = v3zero(); sb.orig = -1;
		sb.lin = sb.ang = sb.extF = sb.extT = v3zero();
	}
	return fixedId;
}

// Row kernels (SSE2 variants, btSequentialImpulseConstraintSolver.cpp:149-232, 318-349)
// returns true when the row changed anything (its applied impulse or a body's delta velocity)
RSB_HD bool resolve_row(SolverBody& A, SolverBody& B, SolverRow& c, bool upperLimit) {
	const float applied0 = c.applied;
	float deltaImpulse = c.rhs - c.applied * c.cfm;
	float dv1 = simd_dot3(c.n1, A.dLin) + simd_dot3(c.rp1xn, A.dAng);
	float dv2 = simd_dot3(c.n2, B.dLin) + simd_dot3(c.rp2xn, B.dAng);
	deltaImpulse = deltaImpulse - dv1 * c.jacDiagInv;
	deltaImpulse = deltaImpulse - dv2 * c.jacDiagInv;
	float sum = c.applied + deltaImpulse;
	if (sum < c.lower) { deltaImpulse = c.lower - c.applied; c.applied = c.lower; }
	else if (upperLimit && !(sum < c.upper)) { deltaImpulse = c.upper - c.applied; c.applied = c.upper; }
	else c.applied = sum;
	bool chg = c.applied != applied0;
	chg |= sb_apply_impulse_chg(A, c.n1 * A.invMassV, c.angA, deltaImpulse);
	chg |= sb_apply_impulse_chg(B, c.n2 * B.invMassV, c.angB, deltaImpulse);
	return chg;
}
RSB_HD float resolve_split(SolverBody& A, SolverBody& B, SolverRow& c) {
	if (c.rhsPen == 0.f) return 0.f;
	float deltaImpulse = c.rhsPen - c.appliedPush * c.cfm;
	float dv1 = simd_dot3(c.n1, A.push) + simd_dot3(c.rp1xn, A.turn);
	float dv2 = simd_dot3(c.n2, B.push) + simd_dot3(c.rp2xn, B.turn);
	deltaImpulse = deltaImpulse - dv1 * c.jacDiagInv;
	deltaImpulse = deltaImpulse - dv2 * c.jacDiagInv;
	float sum = c.appliedPush + deltaImpulse;
	if (sum < c.lower) { deltaImpulse = c.lower - c.appliedPush; c.appliedPush = c.lower; }
	else c.appliedPush = sum;
	sb_apply_push(A, c.n1 * A.invMassV, c.angA, deltaImpulse);
	sb_apply_push(B, c.n2 * B.invMassV, c.angB, deltaImpulse);
	return deltaImpulse * (1.f / c.jacDiagInv);
}

// solveGroup for the whole arena: every awake dynamic body is converted (so that bodies
// without contacts still receive their external force impulse), contacts -> rows, special
// contacts, iterations, writeback.
RSB_HD void solve_arena(const Ctx& cx, int nSolveManifolds) {
	const SolverView sv = solver_view(cx);
	const int nb = n_bodies(cx);
	const float dt = cx.p->dt;
	int companion[MAX_BODIES];
	// special (ball-world) contact accumulation per body (btCollisionObject::m_specialResolveInfo)
	int spNum[MAX_BODIES]; float spFriction[MAX_BODIES], spRestitution[MAX_BODIES], spTotalDist[MAX_BODIES]; V3 spTotalNormal[MAX_BODIES];
	for (int b = 0; b < nb; b++) { companion[b] = -1; spNum[b] = 0; spTotalDist[b] = 0.f; spTotalNormal[b] = v3zero(); spFriction[b] = 0.f; spRestitution[b] = 0.f; }
	int fixedId = -1;
	sv.cnt[0] = 0; sv.cnt[1] = 0; sv.cnt[2] = 0;
	RSB_SUBPROF_BEGIN();

	// convertBodies: the reference converts every body of every awake island.  A body that ends up without a row only
	// receives v += externalForceImpulse in the write-back, which solve_body_no_contacts reproduces bit for bit on the
	// body's own thread (PH_FINISH_BODY); solver bodies are therefore created lazily, for bodies that carry a contact.
	// (Solver-body numbering is internal; the only order that matters -- special contacts per body -- involves the ball alone.)

	RSB_SUBPROF_MARK(2);
	// convertContacts
	for (int k = 0; k < nSolveManifolds; k++) {
		Manifold& mf = cx.mf[cx.pl->manifoldOrder[k]];
		if (mf.nPts == 0) continue;
		int ia = get_or_init_solver_body(cx, mf.body0, companion, fixedId);
		int ib = get_or_init_solver_body(cx, mf.body1, companion, fixedId);
		for (int j = 0; j < mf.nPts; j++) {
			ContactPt& cp = mf.pts[j];
			if (!(cp.dist <= mf.processingThresh)) continue;
			if (sv.cnt[1] >= MAX_ROWS) { set_error(cx, ERR_MANIFOLD_OVERFLOW); break; }
			V3 org0 = mf.body0 < STATIC_BASE ? rigid_of(cx, mf.body0).pos : static_origin(cx, mf.body0);
			V3 org1 = mf.body1 < STATIC_BASE ? rigid_of(cx, mf.body1).pos : static_origin(cx, mf.body1);
			V3 rel1 = cp.posA - org0;
			V3 rel2 = cp.posB - org1;
			int frictionIndex = sv.cnt[1];
			SolverRow& row = sv.row(sv.cnt[1]++);
			setup_contact_row(cx, row, ia, ib, cp.normalB, cp.dist, cp.restitution, cp.friction, cp.isSpecial, rel1, rel2);
			row.frictionIndex = sv.cnt[2];
			if (cp.isSpecial) {
				for (int i = 0; i < 2; i++) {
					int obj = i ? mf.body1 : mf.body0;
					if (obj < STATIC_BASE) {
						spNum[obj]++;
						spFriction[obj] = cp.friction;
						spRestitution[obj] = cp.restitution;
						spTotalNormal[obj] += cp.normalB;
						spTotalDist[obj] += len(i ? rel2 : rel1);
					}
				}
			}
			add_friction_for(cx, ia, ib, cp.normalB, cp.friction, rel1, rel2, frictionIndex);
		}
	}
	// special contacts, per converted body in solver-body order (convertContactSpecial, :1164-1211)
	for (int sbi = 0; sbi < sv.cnt[0]; sbi++) {
		int b = sv.sbody[sbi].orig;
		if (b < 0 || spNum[b] <= 0) continue;
		if (sv.cnt[1] >= MAX_ROWS) { set_error(cx, ERR_MANIFOLD_OVERFLOW); break; }
		float distance = spTotalDist[b] / (float)spNum[b];
		V3 normal = spTotalNormal[b] / (float)spNum[b];
		int frictionIndex = sv.cnt[1];
		int ia = companion[b];
		int ib = get_or_init_solver_body(cx, STATIC_BASE, companion, fixedId);
		V3 rel1 = normal * -distance;
		V3 rel2 = v3zero();
		SolverRow& row = sv.row(sv.cnt[1]++);
		setup_contact_row(cx, row, ia, ib, normal, distance, spRestitution[b], spFriction[b], 0, rel1, rel2);
		row.frictionIndex = sv.cnt[2];
		add_friction_for(cx, ia, ib, normal, spFriction[b], rel1, rel2, frictionIndex);
		spNum[b] = 0;
	}

	RSB_SUBPROF_MARK(3);
	// split impulse iterations (:1778-1810)
	for (int it = 0; it < 10; it++) {
		float lsr = 0.f;
		for (int j = 0; j < sv.cnt[1]; j++) {
			SolverRow& c = sv.row(j);
			float residual = resolve_split(sv.sbody[c.bodyA], sv.sbody[c.bodyB], c);
			float r2 = residual * residual;
			if (r2 > lsr) lsr = r2;
		}
		if (lsr <= 0.f || it >= 9) break;
	}
	// velocity iterations (:1601-1776): non-special contacts, then friction rows
	RSB_SUBPROF_MARK(4);
	// The reference always runs 10 sweeps (m_leastSquaresResidualThreshold = 0).  A sweep is a deterministic function of
	// (applied impulses, body delta velocities); once a sweep leaves all of them unchanged it is at a fixed point and the
	// remaining sweeps would reproduce it bit for bit, so they are skipped.
	for (int it = 0; it < 10; it++) {
		bool changed = false;
		for (int j = 0; j < sv.cnt[1]; j++) {
			SolverRow& c = sv.row(j);
			if (c.isSpecial) continue;
			changed |= resolve_row(sv.sbody[c.bodyA], sv.sbody[c.bodyB], c, false);
		}
		for (int j = 0; j < sv.cnt[2]; j++) {
			SolverRow& c = sv.frow(j);
			float totalImpulse = sv.row(c.frictionIndex).applied;
			if (totalImpulse > 0.f) {
				c.lower = -(c.friction * totalImpulse);
				c.upper = c.friction * totalImpulse;
				changed |= resolve_row(sv.sbody[c.bodyA], sv.sbody[c.bodyB], c, true);
			}
		}
		if (!changed) break;
	}
	RSB_SUBPROF_MARK(5);
	// writeBackBodies (:1878-1904, btSolverBody.h:265-287)
	for (int i = 0; i < sv.cnt[0]; i++) {
		SolverBody& sb = sv.sbody[i];
		if (sb.orig < 0) continue;
		RigidS& rb = rigid_of(cx, sb.orig);
		sb.lin += sb.dLin;
		sb.ang += sb.dAng;
		if (sb.push.x != 0.f || sb.push.y != 0.f || sb.push.z != 0.f || sb.turn.x != 0.f || sb.turn.y != 0.f || sb.turn.z != 0.f) {
			Xform nt;
			bool noRot = (sb.orig == 0) && cx.p->ballNoRot;
			integrate_transform(sb.xf, sb.push, sb.turn * 0.1f, dt, noRot, nt);
			sb.xf = nt;
		}
		rb.vel = sb.lin + sb.extF;
		rb.angvel = sb.ang + sb.extT;
		rb.pos = sb.xf.origin;
		rb.basis = sb.xf.basis;
		cx.misc[2] |= 1 << sb.orig;  // written back here: PH_FINISH_BODY must not touch this body's velocity again
	}
	RSB_SUBPROF_MARK(6);
}

// Does any manifold handed to the solver carry a contact point?  If not, solveGroup degenerates to the
// per-body write-back v += externalForceImpulse (no rows, no push): solve_body_no_contacts, one thread per body.
RSB_HD bool solver_has_rows(const Ctx& cx, int nSolveManifolds) {
	for (int k = 0; k < nSolveManifolds; k++) if (cx.mf[cx.pl->manifoldOrder[k]].nPts > 0) return true;
	return false;
}

RSB_HD void solve_body_no_contacts(const Ctx& cx, int b) {
	const int nb = n_bodies(cx);
	const int root = cx.body[b].islandTag;
	bool awake = false;  // convertBodies only visits islands with an active body
	for (int o = 0; o < nb; o++) if (cx.body[o].islandTag == root && body_active(cx.body[o])) { awake = true; break; }
	if (!awake) return;
	const BodyW& bw = cx.body[b];
	RigidS& rb = rigid_of(cx, b);
	V3 extF = bw.totalForce * bw.invMass * cx.p->dt;
	V3 extT = vmul(bw.totalTorque, bw.invInertiaW) * cx.p->dt;
	V3 lin = rb.vel; lin += v3zero();
	V3 ang = rb.angvel; ang += v3zero();
	rb.vel = lin + extF;
	rb.angvel = ang + extT;
}

}  // namespace rsb
