// rsb_collide.h -- per-tick collision detection of one arena: swept AABBs, grid broadphase
// pairs, narrowphase per pair and the game's contact hooks.  Restates
//   btCollisionWorld::updateAabbs / updateSingleAabb      bt/BulletCollision/CollisionDispatch/btCollisionWorld.cpp:143-193
//   btRSBroadphase::setAabb / calculateOverlappingPairs   bt/BulletCollision/BroadphaseCollision/btRSBroadphase.cpp:286-323, 392-468
//   algorithm table                                       bt/BulletCollision/CollisionDispatch/btDefaultCollisionConfiguration.cpp:265-342
//   compound / convex-concave / convex-plane / sphere-triangle algorithms (files cited at each function)
//   btManifoldResult::addContactPoint                     bt/BulletCollision/CollisionDispatch/btManifoldResult.cpp:109-207
//   btPersistentManifold (4-point cache)                  bt/BulletCollision/NarrowPhaseCollision/btPersistentManifold.cpp:110-329
//   Arena::_BulletContactAddedCallback and hooks          src/Sim/Arena/Arena.cpp:244-414
//   btAdjustInternalEdgeContacts                          bt/BulletCollision/CollisionDispatch/btInternalEdgeUtility.cpp:357-798
#pragma once
#include "rsb_work.h"
#include "rsb_ray.h"
#include "rsb_ball.h"
#include "rsb_gjk.h"

namespace rsb {

constexpr float kContactBreaking = 0.02f;  // gContactBreakingThreshold

// static object index k (= obj - STATIC_BASE): meshes [0, nMeshes), planes [nMeshes, nMeshes + nPlanes), Dropshot tiles behind them
RSB_HD int static_tile_index(const Scene& sc, int k) { return k - sc.nMeshes - sc.nPlanes; }  // >= 0: a Dropshot tile
RSB_HD V3 static_origin(const Ctx& cx, int obj) {
	int k = obj - STATIC_BASE;
	if (k < cx.sc->nMeshes) return v3zero();
	if (k < cx.sc->nMeshes + cx.sc->nPlanes) return cx.sc->planes[k - cx.sc->nMeshes].pos;
	return V3(0.f, 0.f, -0.04f);  // tile bodies sit one collision margin below the floor (Arena.cpp:1078)
}
RSB_HD void static_aabb(const Ctx& cx, int obj, V3& mn, V3& mx) {
	int k = obj - STATIC_BASE;
	if (k < cx.sc->nMeshes) { mn = cx.sc->meshes[k].aabbMin; mx = cx.sc->meshes[k].aabbMax; }
	else if (k < cx.sc->nMeshes + cx.sc->nPlanes) { mn = cx.sc->planes[k - cx.sc->nMeshes].aabbMin; mx = cx.sc->planes[k - cx.sc->nMeshes].aabbMax; }
	else { const TileD& t = cx.sc->tiles[static_tile_index(*cx.sc, k)]; mn = t.aabbMin; mx = t.aabbMax; }
}
// CF_NO_CONTACT_RESPONSE of a static object: only a broken Dropshot tile has it (Arena::SetDropshotTilesState, Arena.cpp:225-242)
RSB_HD bool static_no_response(const Ctx& cx, int obj) {
	if (cx.mode != DROPSHOT) return false;
	const int t = static_tile_index(*cx.sc, obj - STATIC_BASE);
	return t >= 0 && (cx.g->tiles[t] & 4u) != 0;
}
// broadphase proxy AABB of a static = shape AABB +- gContactBreakingThreshold (updateSingleAabb)
RSB_HD void static_proxy_aabb(const Ctx& cx, int obj, V3& mn, V3& mx) {
	static_aabb(cx, obj, mn, mx);
	const V3 ct(0.02f, 0.02f, 0.02f);
	mn -= ct; mx += ct;
}

RSB_HD M3 m3abs(const M3& m) { return M3(vabs(m.r[0]), vabs(m.r[1]), vabs(m.r[2])); }

// btTransformAabb(halfExtents, margin, t) (btAabbUtil2.h:172-180) -- box / hull child AABB
RSB_HD void transform_aabb_he(const V3& he, float margin, const M3& basis, const V3& origin, V3& mn, V3& mx) {
	V3 h = he + V3(margin, margin, margin);
	M3 ab = m3abs(basis);
	V3 extent(dot(h, ab.r[0]), dot(h, ab.r[1]), dot(h, ab.r[2]));
	mn = origin - extent; mx = origin + extent;
}
// btTransformAabb(localMin, localMax, margin, t) (btAabbUtil2.h:182-196) / btCompoundShape::getAabb
RSB_HD void transform_aabb_mm(const V3& lmin, const V3& lmax, float margin, const M3& basis, const V3& origin, V3& mn, V3& mx) {
	V3 lh = 0.5f * (lmax - lmin);
	lh += V3(margin, margin, margin);
	V3 lc = 0.5f * (lmax + lmin);
	M3 ab = m3abs(basis);
	V3 center = basis * lc + origin;
	V3 extent(dot(lh, ab.r[0]), dot(lh, ab.r[1]), dot(lh, ab.r[2]));
	mn = center - extent; mx = center + extent;
}

// local AABB of the compound = AABB of its single box child at the child offset
// (btCompoundShape::addChildShape, btCompoundShape.cpp:59-105)
RSB_HD void compound_local_aabb(const CarCfgD& cfg, V3& lmin, V3& lmax) {
	transform_aabb_he(cfg.hitboxImplicit, cfg.hitboxMargin, m3identity(), cfg.hitboxOffset, lmin, lmax);
}
RSB_HD void compound_aabb(const CarCfgD& cfg, const M3& basis, const V3& origin, V3& mn, V3& mx) {
	V3 lmin, lmax;
	compound_local_aabb(cfg, lmin, lmax);
	transform_aabb_mm(lmin, lmax, 0.f, basis, origin, mn, mx);
}

RSB_HD void body_shape_aabb(const Ctx& cx, int b, const M3& basis, const V3& origin, V3& mn, V3& mx) {
	const BatchParams& P = *cx.p;
	if (b == 0) {
		if (cx.ballIsSphere) {
			float margin = P.ballRadiusBT + 0.08f;  // btSphereShape::getAabb (btSphereShape.cpp:52-59)
			V3 e(margin, margin, margin);
			mn = origin - e; mx = origin + e;
		} else {
			transform_aabb_mm(P.puckLocalAabbMin, P.puckLocalAabbMax, 0.04f, basis, origin, mn, mx);
		}
	} else {
		compound_aabb(P.carCfg[b - 1], basis, origin, mn, mx);
	}
}

RSB_HD bool aabb_overlap(const V3& min1, const V3& max1, const V3& min2, const V3& max2) {
	bool overlap = true;
	overlap = (min1.x > max2.x || max1.x < min2.x) ? false : overlap;
	overlap = (min1.z > max2.z || max1.z < min2.z) ? false : overlap;
	overlap = (min1.y > max2.y || max1.y < min2.y) ? false : overlap;
	return overlap;
}

// body b leaves its place in the stamp order and becomes the newest entry
RSB_HD void dyn_order_move_to_back(ArenaS& S, int nb, int b) {
	int pos = 0;
	while (pos < nb && S.dynOrder[pos] != b) pos++;
	for (int q = pos; q + 1 < nb; q++) S.dynOrder[q] = S.dynOrder[q + 1];
	S.dynOrder[nb - 1] = (uint8_t)b;
}

// updateSingleAabb of one dynamic body (current U predicted transform, +-0.02) and the grid cell of its min corner
RSB_HD void body_filter(const Ctx& cx, int obj, int& group, int& mask);
RSB_HD Convex car_box(const CarCfgD& cfg);
RSB_HD Xform car_child_xform(const RigidS& rb, const CarCfgD& cfg);
RSB_HD void convex_world_aabb(const Ctx& cx, const Convex& s, const Xform& t, V3& mn, V3& mx);
RSB_HD void static_proxy_aabb(const Ctx& cx, int obj, V3& mn, V3& mx);
RSB_HD void update_aabb(const Ctx& cx, int b) {
	const V3 ct(kContactBreaking, kContactBreaking, kContactBreaking);
	BodyW& bw = cx.body[b];
	const RigidS& rb = rigid_of(cx, b);
	V3 mn, mx, mn2, mx2;
	body_shape_aabb(cx, b, rb.basis, rb.pos, mn, mx);
	mn -= ct; mx += ct;
	body_shape_aabb(cx, b, bw.interp.basis, bw.interp.origin, mn2, mx2);
	mn2 -= ct; mx2 += ct;
	mn = vmin(mn, mn2); mx = vmax(mx, mx2);
	bw.aabbMin = mn; bw.aabbMax = mx;
	int i, j, k;
	bw.newCell = grid_cell_of(*cx.sc, mn, i, j, k);
	bw.newI = i; bw.newJ = j; bw.newK = k;
	if (b > 0) {
		const CarCfgD& cfg = cx.p->carCfg[b - 1];
		bw.childXf = car_child_xform(rb, cfg);
		convex_world_aabb(cx, car_box(cfg), bw.childXf, bw.childMin, bw.childMax);
	}
	// static half of calculateOverlappingPairs for this body (its cell is final: PH_PAIRS only commits newCell)
	{
		const Scene& sc = *cx.sc;
		int gb, mb; body_filter(cx, b, gb, mb);
		int n = 0;
		const int sEnd = sc.cellStart[bw.newCell + 1];
		for (int s = sc.cellStart[bw.newCell]; s < sEnd; s++) {
			const int obj = STATIC_BASE + sc.cellStatics[s];
			V3 smn, smx; static_proxy_aabb(cx, obj, smn, smx);
			if (!aabb_overlap(mn, mx, smn, smx)) continue;
			int ga, ma; body_filter(cx, obj, ga, ma);
			if (!((ga & mb) && (gb & ma))) continue;  // needsBroadphaseCollision
			if (n >= MAX_CAND) { n = -1; break; }
			bw.candObj[n] = (uint8_t)(obj - STATIC_BASE);
			bw.cand[n++] = (uint16_t)s;
		}
		bw.nCand = n;
	}
}

// btRSBroadphase::setAabb cell bookkeeping, in body order (a body that changed cell is re-registered last)
RSB_HD void apply_cell_changes(const Ctx& cx) {
	const int nb = n_bodies(cx);
	for (int b = 0; b < nb; b++) {
		const BodyW& bw = cx.body[b];
		if (bw.newCell != cx.s->cellIdx[b]) {
			cx.s->cellIdx[b] = bw.newCell;
			cx.s->cellI[b] = (uint8_t)bw.newI; cx.s->cellJ[b] = (uint8_t)bw.newJ; cx.s->cellK[b] = (uint8_t)bw.newK;
			if (nb > 1) {
				cx.s->cellStamp[b] = ++cx.s->stampCounter;
				dyn_order_move_to_back(*cx.s, nb, b);
			}
		}
	}
}

RSB_HD void body_filter(const Ctx& cx, int obj, int& group, int& mask) {
	if (obj >= STATIC_BASE) {
		int k = obj - STATIC_BASE;
		if (k < cx.sc->nMeshes) { group = cx.sc->meshes[k].group; mask = cx.sc->meshes[k].mask; }
		else if (k < cx.sc->nMeshes + cx.sc->nPlanes) { group = cx.sc->planes[k - cx.sc->nMeshes].group; mask = cx.sc->planes[k - cx.sc->nMeshes].mask; }
		else { group = FILT_DROPSHOT_TILE; mask = FILT_DROPSHOT_TILE; }
	} else if (obj == 0) { group = FILT_DEFAULT | FILT_HOOPS_NET | FILT_DROPSHOT_TILE; mask = FILT_ALL; }
	else { group = FILT_DEFAULT | FILT_DROPSHOT_FLOOR; mask = FILT_ALL; }
}

RSB_HD void add_pair(const Ctx& cx, int a, int b, int entry = 0) {
	if (cx.pl->nPairs >= MAX_MANIFOLDS) { set_error(cx, ERR_MANIFOLD_OVERFLOW); return; }
	cx.pl->a[cx.pl->nPairs] = a; cx.pl->b[cx.pl->nPairs] = b; cx.pl->entry[cx.pl->nPairs] = (uint16_t)entry; cx.pl->nPairs++;
}

// calculateOverlappingPairs: pair list in creation order; (proxy0, proxy1) sorted by handle id
// (statics first; ball before cars; lower car index first).
// The collision filter (needsBroadphaseCollision) of a (static, body) pair was applied when the body's candidate list was
// built (update_aabb); two dynamic bodies always pass it (ball: group DEFAULT|..., mask ALL; cars likewise).  A dynamic pair
// (lo, hi) is created when the LOWER body is visited -- adjacency and AABB overlap are symmetric, so by the time the higher
// body meets the lower one the pair exists already (the reference's pair-cache lookup) -- in the order of the cell's
// dynamic-handle list.
RSB_HD void compute_pairs(const Ctx& cx) {
	const Scene& sc = *cx.sc;
	const int nb = n_bodies(cx);
	// The pair counter lives in a register (stores to the list are fire-and-forget, a counter in shared memory would be a
	// dependent load per pair), and the cell coordinates / stamp order of the <= 7 bodies are fetched as four 8-byte words.
	int nPairs = 0;
	bool overflow = false;
	uint32_t dynMask = 0;
	PairList& pl = *cx.pl;
	const uint64_t ordW = *(const uint64_t*)cx.s->dynOrder, ciW = *(const uint64_t*)cx.s->cellI, cjW = *(const uint64_t*)cx.s->cellJ, ckW = *(const uint64_t*)cx.s->cellK;
#define RSB_PAIR_ADD(A_, B_, E_) do { if (nPairs >= MAX_MANIFOLDS) overflow = true; else { pl.a[nPairs] = (uint8_t)(A_); pl.b[nPairs] = (uint8_t)(B_); pl.entry[nPairs] = (uint16_t)(E_); nPairs++; } } while (0)
	for (int b = 0; b < nb; b++) {
		const BodyW& bw = cx.body[b];
		if (bw.nCand >= 0) {
			for (int q = 0; q < bw.nCand; q++) RSB_PAIR_ADD(STATIC_BASE + bw.candObj[q], b, bw.cand[q]);
		} else {  // candidate list overflowed: full scan
			int gb, mb; body_filter(cx, b, gb, mb);
			int cell = cx.s->cellIdx[b];
			for (int s = sc.cellStart[cell]; s < sc.cellStart[cell + 1]; s++) {
				int obj = STATIC_BASE + sc.cellStatics[s];
				V3 mn, mx; static_proxy_aabb(cx, obj, mn, mx);
				int ga, ma; body_filter(cx, obj, ga, ma);
				if (aabb_overlap(bw.aabbMin, bw.aabbMax, mn, mx) && (ga & mb) && (gb & ma)) RSB_PAIR_ADD(obj, b, s);
			}
		}
		if (b + 1 < nb) {
			// dynamic proxies registered in b's cell, in the cell list's order (ascending stamp)
			const int ci = (int)((ciW >> (8 * b)) & 255), cj = (int)((cjW >> (8 * b)) & 255), ck = (int)((ckW >> (8 * b)) & 255);
			for (int q = 0; q < nb; q++) {
				const int o = (int)((ordW >> (8 * q)) & 255);
				if (o <= b) continue;
				if (abs((int)((ciW >> (8 * o)) & 255) - ci) > 1 || abs((int)((cjW >> (8 * o)) & 255) - cj) > 1 || abs((int)((ckW >> (8 * o)) & 255) - ck) > 1) continue;
				const BodyW& ow = cx.body[o];
				if (!aabb_overlap(bw.aabbMin, bw.aabbMax, ow.aabbMin, ow.aabbMax)) continue;
				if (nPairs < MAX_MANIFOLDS) dynMask |= 1u << nPairs;
				RSB_PAIR_ADD(b, o, 0);
			}
		}
	}
#undef RSB_PAIR_ADD
	if (overflow) set_error(cx, ERR_MANIFOLD_OVERFLOW);
	pl.nPairs = nPairs;
	pl.dynMask = dynMask;
}

// ------------------------------------------------------------------ manifolds + contact hooks
RSB_HD float body_angular_motion_disc(const Ctx& cx, int obj) {
	// btCollisionShape::getAngularMotionDisc (btCollisionShape.cpp:126-158)
	if (obj >= STATIC_BASE) return kLargeFloat;
	const BatchParams& P = *cx.p;
	if (obj == 0 && cx.ballIsSphere) return (P.ballRadiusBT + 0.08f) + 0.f;
	V3 mn, mx;
	body_shape_aabb(cx, obj, m3identity(), v3zero(), mn, mx);
	float radius = len(mx - mn) * 0.5f;
	V3 center = (mn + mx) * 0.5f;
	return radius + len(center);
}

// manifold of broadphase pair `slot` (btCollisionDispatcher::getNewManifold, btCollisionDispatcher.cpp:70-104)
RSB_HD int new_manifold(const Ctx& cx, int slot, int body0, int body1) {
	Manifold& m = cx.mf[slot];
	m.body0 = body0; m.body1 = body1; m.nPts = 0; m.pts = cx.w->points + slot * 4;
#if defined(__CUDA_ARCH__)
	atomicOr(&cx.pl->mfMask, 1u << slot);
#else
	cx.pl->mfMask |= 1u << slot;
#endif
	float t0 = (body0 >= STATIC_BASE ? kLargeFloat : cx.p->motionDisc[body0]) * kContactBreaking;
	float t1 = (body1 >= STATIC_BASE ? kLargeFloat : cx.p->motionDisc[body1]) * kContactBreaking;
	m.breakingThresh = t0 < t1 ? t0 : t1;
	m.processingThresh = kLargeFloat;
	return slot;
}

RSB_HD Xform obj_xform(const Ctx& cx, int obj) {
	Xform t;
	if (obj >= STATIC_BASE) { t.basis = m3identity(); t.origin = static_origin(cx, obj); }
	else { const RigidS& rb = rigid_of(cx, obj); t.basis = rb.basis; t.origin = rb.pos; }
	return t;
}

// sortCachedPoints (btPersistentManifold.cpp:110-189), gContactCalcArea3Points = true
RSB_HD int sort_cached_points(const Manifold& m, const ContactPt& pt) {
	int maxPenetrationIndex = -1;
	float maxPenetration = pt.dist;
	for (int i = 0; i < 4; i++) if (m.pts[i].dist < maxPenetration) { maxPenetrationIndex = i; maxPenetration = m.pts[i].dist; }
	float res0 = 0.f, res1 = 0.f, res2 = 0.f, res3 = 0.f;
	if (maxPenetrationIndex != 0) { V3 a0 = pt.localA - m.pts[1].localA; V3 b0 = m.pts[3].localA - m.pts[2].localA; res0 = len2(cross(a0, b0)); }
	if (maxPenetrationIndex != 1) { V3 a1 = pt.localA - m.pts[0].localA; V3 b1 = m.pts[3].localA - m.pts[2].localA; res1 = len2(cross(a1, b1)); }
	if (maxPenetrationIndex != 2) { V3 a2 = pt.localA - m.pts[0].localA; V3 b2 = m.pts[3].localA - m.pts[1].localA; res2 = len2(cross(a2, b2)); }
	if (maxPenetrationIndex != 3) { V3 a3 = pt.localA - m.pts[0].localA; V3 b3 = m.pts[2].localA - m.pts[1].localA; res3 = len2(cross(a3, b3)); }
	// btVector4::closestAxis4 = absolute4().maxAxis4(): first index of the maximum
	float v[4] = {fabsf(res0), fabsf(res1), fabsf(res2), fabsf(res3)};
	int maxIndex = -1; float maxVal = -kLargeFloat;
	for (int i = 0; i < 4; i++) if (v[i] > maxVal) { maxIndex = i; maxVal = v[i]; }
	return maxIndex;
}

RSB_HD float get_angle(const V3& edgeA, const V3& normalA, const V3& normalB) { return rs_atan2f(dot(normalB, edgeA), dot(normalB, normalA)); }
RSB_HD V3 quat_rotate(const Q4& q, const V3& v) {  // quatRotate (btQuaternion.h): q * v * q^-1
	// q * v (quaternion times vector)
	Q4 t = q4(q.w * v.x + q.y * v.z - q.z * v.y, q.w * v.y + q.z * v.x - q.x * v.z, q.w * v.z + q.x * v.y - q.y * v.x, -q.x * v.x - q.y * v.y - q.z * v.z);
	Q4 qi = q4(-q.x, -q.y, -q.z, q.w);
	Q4 r = qmul(t, qi);
	return V3(r.x, r.y, r.z);
}
RSB_HD V3 nearest_point_in_segment(const V3& point, const V3& line0, const V3& line1) {
	V3 lineDelta = line1 - line0;
	if (fuzzy_zero(lineDelta)) return line0;
	float delta = dot(point - line0, lineDelta) / dot(lineDelta, lineDelta);
	if (delta < 0.f) delta = 0.f; else if (delta > 1.f) delta = 1.f;
	return line0 + lineDelta * delta;
}
RSB_HD bool clamp_normal(const V3& edge, const V3& tri_normal, const V3& localContactNormalOnB, float correctedEdgeAngle, V3& clamped) {
	V3 edgeCross = normalized(cross(edge, tri_normal));
	float curAngle = get_angle(edgeCross, tri_normal, localContactNormalOnB);
	if (correctedEdgeAngle < 0.f) {
		if (curAngle < correctedEdgeAngle) {
			float diffAngle = correctedEdgeAngle - curAngle;
			clamped = m3_from_quat(q_axis_angle(edge, diffAngle)) * localContactNormalOnB;
			return true;
		}
	}
	if (correctedEdgeAngle >= 0.f) {
		if (curAngle > correctedEdgeAngle) {
			float diffAngle = correctedEdgeAngle - curAngle;
			clamped = m3_from_quat(q_axis_angle(edge, diffAngle)) * localContactNormalOnB;
			return true;
		}
	}
	return false;
}

// btAdjustInternalEdgeContacts with normalAdjustFlags = 0; the triangle's mesh sits at the
// world origin with identity basis, so "local" == world for B.
RSB_HD_NOINLINE void adjust_internal_edge(Ctx cx, ContactPt& cp, const MeshD& mesh, int triIndex) {
	const TriInfo info = mesh.info[triIndex];
	if (info.flags < 0) return;  // no entry in the triangle info map
	V3 v0, v1, v2;
	mesh_triangle(mesh, triIndex, v0, v1, v2);
	V3 tri_normal = normalized(cross(v1 - v0, v2 - v0));
	V3 nearest = nearest_point_in_segment(cp.localB, v0, v1);
	const V3 contact = cp.localB;
	bool isNearEdge = false;
	int numConcaveEdgeHits = 0, numConvexEdgeHits = 0;
	V3 localContactNormalOnB = normalized(transpose(m3identity()) * cp.normalB);
	const float maxEdgeAngle = k2Pi, edgeDistThresh = 0.1f, convexEps = 0.f;
	int bestedge = -1;
	float disttobestedge = kLargeFloat;
	if (fabsf(info.a01) < maxEdgeAngle) { V3 n = nearest_point_in_segment(cp.localB, v0, v1); float l = len(contact - n); if (l < disttobestedge) { bestedge = 0; disttobestedge = l; } }
	if (fabsf(info.a12) < maxEdgeAngle) { V3 n = nearest_point_in_segment(cp.localB, v1, v2); float l = len(contact - n); if (l < disttobestedge) { bestedge = 1; disttobestedge = l; } }
	if (fabsf(info.a20) < maxEdgeAngle) { V3 n = nearest_point_in_segment(cp.localB, v2, v0); float l = len(contact - n); if (l < disttobestedge) { bestedge = 2; disttobestedge = l; } }

	for (int e = 0; e < 3; e++) {
		float angle = e == 0 ? info.a01 : (e == 1 ? info.a12 : info.a20);
		V3 ea = e == 0 ? v0 : (e == 1 ? v1 : v2);
		V3 eb = e == 0 ? v1 : (e == 1 ? v2 : v0);
		if (e > 0) nearest = nearest_point_in_segment(contact, ea, eb);
		if (!(fabsf(angle) < maxEdgeAngle)) continue;
		float l = len(contact - nearest);
		if (!(l < edgeDistThresh)) continue;
		if (bestedge != e) continue;
		V3 edge = ea - eb;
		isNearEdge = true;
		if (angle == 0.f) { numConcaveEdgeHits++; continue; }
		bool isEdgeConvex = (info.flags & (1 << e)) != 0;          // TRI_INFO_VxVy_CONVEX = 1,2,4
		float swapFactor = isEdgeConvex ? 1.f : -1.f;
		V3 nA = swapFactor * tri_normal;
		Q4 orn = q_axis_angle(edge, angle);
		V3 computedNormalB = quat_rotate(orn, tri_normal);
		if (info.flags & (8 << e)) computedNormalB *= -1.f;        // TRI_INFO_VxVy_SWAP_NORMALB = 8,16,32
		V3 nB = swapFactor * computedNormalB;
		float NdotA = dot(localContactNormalOnB, nA);
		float NdotB = dot(localContactNormalOnB, nB);
		bool backFacingNormal = (NdotA < convexEps) && (NdotB < convexEps);
		if (backFacingNormal) { numConcaveEdgeHits++; continue; }
		numConvexEdgeHits++;
		// edges 1 and 2 re-derive the local normal WITHOUT normalising it (shadowed variable in the reference)
		V3 lcn = (e == 0) ? localContactNormalOnB : (transpose(m3identity()) * cp.normalB);
		V3 clamped;
		bool isClamped = clamp_normal(edge, swapFactor * tri_normal, lcn, angle, clamped);
		if (isClamped) {
			if (dot(clamped, 1.f * tri_normal) > 0.f) {
				V3 newNormal = m3identity() * clamped;
				cp.normalB = newNormal;
				cp.posB = cp.posA - cp.normalB * cp.dist;
				cp.localB = vmul(cp.posB - v3zero(), m3identity());
			}
		}
	}
	(void)numConvexEdgeHits;
	if (isNearEdge) {
		if (numConcaveEdgeHits > 0) {
			V3 newNormal = tri_normal * 1.f;
			float d = dot(newNormal, localContactNormalOnB);
			if (d < 0.f) return;
			cp.normalB = m3identity() * newNormal;
			cp.posB = cp.posA - cp.normalB * cp.dist;
			cp.localB = vmul(cp.posB - v3zero(), m3identity());
		}
	}
}

// Car::Demolish + bump logic: Arena::_BtCallback_OnCarCarCollision (Arena.cpp:323-405)
RSB_HD_NOINLINE void on_car_car_collision(Ctx cx, int c1, int c2, ContactPt& cp) {
	cp.friction = K::CARCAR_COLLISION_FRICTION;
	cp.restitution = K::CARCAR_COLLISION_RESTITUTION;
	const MutD& mut = cx.p->mut;
	for (int i = 0; i < 2; i++) {
		bool isSwapped = (i == 1);
		if (isSwapped) { int t = c1; c1 = c2; c2 = t; }
		CarS& car1 = cx.s->cars[c1];
		CarS& car2 = cx.s->cars[c2];
		if (car1.isDemoed || car2.isDemoed) return;
		if ((car1.carContactOtherID == (uint32_t)(c2 + 1)) && (car1.carContactCooldown > 0.f)) continue;
		V3 pos1 = car1.rb.pos * BT_TO_UU, vel1 = car1.rb.vel * BT_TO_UU;
		V3 pos2 = car2.rb.pos * BT_TO_UU, vel2 = car2.rb.vel * BT_TO_UU;
		V3 deltaPos = pos2 - pos1;
		if (dot(vel1, deltaPos) > 0.f) {
			V3 velDir = rs_normalized(vel1);
			V3 dirToOtherCar = rs_normalized(deltaPos);
			float speedTowardsOtherCar = dot(vel1, dirToOtherCar);
			float otherCarAwaySpeed = dot(vel2, velDir);
			if (speedTowardsOtherCar > otherCarAwaySpeed) {
				V3 localPoint = isSwapped ? cp.localB : cp.localA;
				bool hitWithBumper = (localPoint.x * BT_TO_UU) > K::BUMP_MIN_FORWARD_DIST;
				if (hitWithBumper) {
					bool isDemo;
					switch (mut.demoMode) {
					case 1: isDemo = true; break;
					case 2: isDemo = false; break;
					default: isDemo = car1.isSupersonic != 0;
					}
					if (isDemo && !mut.enableTeamDemos) isDemo = cx.p->team[c1] != cx.p->team[c2];
					if (isDemo) {
						car2.isDemoed = 1;
						car2.demoRespawnTimer = mut.respawnDelay;
						cx.g->demoRotMat[c2] = car2.rb.basis;  // what GetState().rotMat keeps showing until the respawn (Car.cpp:133-138)
					} else {
						bool groundHit = car2.isOnGround != 0;
						float baseScale = groundHit ? curve_bump_vel_ground(speedTowardsOtherCar) : curve_bump_vel_air(speedTowardsOtherCar);
						V3 hitUpDir = car2.isOnGround ? car2.rb.basis.col(2) : V3(0.f, 0.f, 1.f);
						V3 bumpImpulse = velDir * baseScale + hitUpDir * curve_bump_upward(speedTowardsOtherCar) * mut.bumpForceScale;
						car2.velImpulseCache += bumpImpulse * UU_TO_BT;
					}
					car1.carContactOtherID = (uint32_t)(c2 + 1);
					car1.carContactCooldown = mut.bumpCooldownTime;
					push_event(cx, isDemo ? 3u : 2u, (uint32_t)(c1 + 1), (uint32_t)(c2 + 1));
				}
			}
		}
	}
}

// btManifoldResult::addContactPoint + gContactAddedCallback.  body0/body1 of the manifold
// are "A"/"B"; (normalOnB, pointOnB, depth) as reported by the detector.  triMesh/triIndex
// identify the triangle when B is an arena mesh (else triMesh < 0).
// dynPair: the manifold joins two dynamic bodies (car-ball / car-car); false: (dynamic body, static object).  A constant at
// every inlined call site, so each site keeps only the hooks it can reach.
RSB_ACP_ATTR void add_contact_point(const Ctx& cx, int mi, const V3& normalOnB, const V3& pointInWorld, float depth, int triMesh, int triIndex, bool dynPair) {
	Manifold& m = cx.mf[mi];
	if (depth > m.breakingThresh) return;
	ContactPt pt;
	V3 pointA = pointInWorld + normalOnB * depth;
	Xform t0 = obj_xform(cx, m.body0), t1 = obj_xform(cx, m.body1);
	pt.localA = xf_inv_apply(t0, pointA);
	pt.localB = xf_inv_apply(t1, pointInWorld);
	pt.normalB = normalOnB; pt.dist = depth;
	pt.posA = pointA; pt.posB = pointInWorld;
	pt.isSpecial = 0;
	pt.partId0 = -1; pt.index0 = (m.body0 < STATIC_BASE && m.body0 > 0) ? 0 : -1;
	pt.partId1 = triMesh >= 0 ? 0 : -1; pt.index1 = triMesh >= 0 ? triIndex : -1;
	// combined friction / restitution (btManifoldResult.cpp:56-77)
	{
		const bool s0 = false, s1 = !dynPair;
		float fa = s0 ? K::ARENA_COLLISION_BASE_FRICTION : cx.body[m.body0].friction;
		float fb = s1 ? K::ARENA_COLLISION_BASE_FRICTION : cx.body[m.body1].friction;
		float ra = s0 ? K::ARENA_COLLISION_BASE_RESTITUTION : cx.body[m.body0].restitution;
		float rb = s1 ? K::ARENA_COLLISION_BASE_RESTITUTION : cx.body[m.body1].restitution;
		if (s0 || s1) { pt.friction = (fa < fb) ? fa : fb; pt.restitution = (ra > rb) ? ra : rb; }
		else { pt.friction = fa * fb; pt.restitution = ra * rb; }
	}
	int insertIndex = m.nPts;
	if (insertIndex == 4) insertIndex = sort_cached_points(m, pt);
	else m.nPts++;
	if (insertIndex < 0) insertIndex = 0;
	m.pts[insertIndex] = pt;
	ContactPt& cp = m.pts[insertIndex];

	// ---- Arena::_BulletContactAddedCallback ----
	const bool d0 = true, d1 = dynPair;  // body0 of a manifold is always dynamic; body1 is static unless dynPair
	if ((d0 && cx.body[m.body0].noResponse) || (d1 && cx.body[m.body1].noResponse)) return;
	// manifold orders used here: (car, ball) (carHi, carLo) (car, world) (ball, world)
	if (d0 && m.body0 > 0) {
		int car = m.body0 - 1;
		if (d1 && m.body1 == 0) {
			V3 relBallPos = cp.localB * BT_TO_UU;
			cp.friction = K::CARBALL_COLLISION_FRICTION;
			cp.restitution = K::CARBALL_COLLISION_RESTITUTION;
			ball_on_hit(cx, car, relBallPos);
		} else if (d1) {
			on_car_car_collision(cx, car, m.body1 - 1, cp);
		} else {
			cx.s->cars[car].worldContactHas = 1;
			cx.s->cars[car].worldContactNormal = cp.normalB;
			cp.friction = cx.p->mut.carWorldFriction;
			cp.restitution = cx.p->mut.carWorldRestitution;
		}
	} else if (d0 && m.body0 == 0 && !d1) {
		ball_on_world_collision(cx, cp.normalB);
		if (cx.mode != SNOWDAY) cp.isSpecial = 1;
	}
	if (triMesh >= 0) adjust_internal_edge(cx, cp, cx.sc->meshes[triMesh], triIndex);
}

// Ball::_OnDropshotTileCollision (Ball.cpp:364-409), called from the (ball, tile) branch of the contact callback (Arena.cpp:291-296).
// Tile states live in the cold tail of the arena record; the tile's rigid-body flag (bit 2) is only synchronised at the end of
// the tick (tick_tail), so that the ball still bounces off the tile it has just broken.
RSB_HD_NOINLINE void ball_on_dropshot_tile(Ctx cx, int tile) {
	ArenaS& G = *cx.g;
	const int team = tile / TILES_PER_TEAM, idx = tile - team * TILES_PER_TEAM;
	if ((G.tiles[tile] & 3u) == 2u) return;  // already broken (two tiles hit at once)
	const uint64_t tickCount = cx.s->tickCount;
	if (G.dsHasDamaged) {
		float timeSinceDamage = (tickCount - G.dsLastDamageTick) * cx.p->dt;
		if (timeSinceDamage <= 0.1f) return;  // MIN_DAMAGE_INTERVAL
	}
	const V3 vel = cx.s->ball.rb.vel * BT_TO_UU;
	if (vel.z > -250.f) return;  // MIN_DOWNWARD_SPEED_TO_DAMAGE
	const TileD& td = cx.sc->tiles[tile];
	if (G.dsChargeLevel > 1 && G.dsYTargetDir != 0.f) {
		const int sgn = (td.posY > 0.f) - (td.posY < 0.f);
		if ((float)sgn != G.dsYTargetDir) return;  // wrong side of the arena
	}
	const int lvl = G.dsChargeLevel;
	const int n = lvl <= 1 ? 1 : (lvl == 2 ? td.n1 : td.n2);
	for (int q = 0; q < n; q++) {
		const int i = lvl <= 1 ? idx : (lvl == 2 ? td.nb1[q] : td.nb2[q]);
		uint8_t& st = G.tiles[team * TILES_PER_TEAM + i];
		if ((st & 3u) != 2u) st = (uint8_t)((st & ~3u) | ((st & 3u) + 1u));
	}
	G.dsHasDamaged = 1u;
	G.dsLastDamageTick = tickCount;
	G.dsAccumulatedHitForce = 0.f;
	G.dsChargeLevel = 1;
	G.dsYTargetDir = 0.f;
}

// addContactPoint for a (tile, ball) manifold: convex-convex keeps the broadphase order, so body0 is the STATIC tile and body1
// the ball (btConvexConvexAlgorithm.cpp:268-276); combined material as for any static pair; the callback only runs the tile hook.
RSB_HD void add_contact_point_tile(const Ctx& cx, int mi, const V3& normalOnB, const V3& pointInWorld, float depth) {
	Manifold& m = cx.mf[mi];
	if (depth > m.breakingThresh) return;
	ContactPt pt;
	V3 pointA = pointInWorld + normalOnB * depth;
	Xform t0 = obj_xform(cx, m.body0), t1 = obj_xform(cx, m.body1);
	pt.localA = xf_inv_apply(t0, pointA);
	pt.localB = xf_inv_apply(t1, pointInWorld);
	pt.normalB = normalOnB; pt.dist = depth;
	pt.posA = pointA; pt.posB = pointInWorld;
	pt.isSpecial = 0;
	pt.partId0 = -1; pt.index0 = -1; pt.partId1 = -1; pt.index1 = -1;
	{
		const float fa = K::ARENA_COLLISION_BASE_FRICTION, fb = cx.body[m.body1].friction;
		const float ra = K::ARENA_COLLISION_BASE_RESTITUTION, rb = cx.body[m.body1].restitution;
		pt.friction = (fa < fb) ? fa : fb; pt.restitution = (ra > rb) ? ra : rb;
	}
	int insertIndex = m.nPts;
	if (insertIndex == 4) insertIndex = sort_cached_points(m, pt);
	else m.nPts++;
	if (insertIndex < 0) insertIndex = 0;
	m.pts[insertIndex] = pt;
	if (static_no_response(cx, m.body0) || cx.body[m.body1].noResponse) return;
	ball_on_dropshot_tile(cx, static_tile_index(*cx.sc, m.body0 - STATIC_BASE));
}

// btPersistentManifold::refreshContactPoints (btPersistentManifold.cpp:263-329)
RSB_HD void refresh_contact_points(const Ctx& cx, int mi) {
	Manifold& m = cx.mf[mi];
	Xform tA = obj_xform(cx, m.body0), tB = obj_xform(cx, m.body1);
	for (int i = m.nPts - 1; i >= 0; i--) {
		ContactPt& p = m.pts[i];
		p.posA = xf_apply(tA, p.localA);
		p.posB = xf_apply(tB, p.localB);
		p.dist = dot(p.posA - p.posB, p.normalB);
	}
	for (int i = m.nPts - 1; i >= 0; i--) {
		ContactPt& p = m.pts[i];
		bool remove = false;
		if (!(p.dist <= m.breakingThresh)) remove = true;
		else {
			V3 projectedPoint = p.posA - p.normalB * p.dist;
			V3 projectedDifference = p.posB - projectedPoint;
			float distance2d = dot(projectedDifference, projectedDifference);
			if (distance2d > m.breakingThresh * m.breakingThresh) remove = true;
		}
		if (remove) {
			int last = m.nPts - 1;
			if (i != last) m.pts[i] = m.pts[last];
			m.nPts--;
		}
	}
}

// ------------------------------------------------------------------ narrowphase algorithms
// SphereTriangleDetector::collide (SphereTriangleDetector.cpp:139-243)
RSB_HD bool sphere_triangle(const V3& sphereCenter, float radius, const V3 v[3], float contactBreakingThreshold, V3& point, V3& resultNormal, float& depth) {
	float radiusWithThreshold = radius + contactBreakingThreshold;
	V3 normal = cross(v[1] - v[0], v[2] - v[0]);
	float l2 = len2(normal);
	bool hasContact = false;
	V3 contactPoint = v3zero();
	if (l2 >= kEps * kEps) {
		normal = normal / sqrtf(l2);
		V3 p1ToCentre = sphereCenter - v[0];
		float distanceFromPlane = dot(p1ToCentre, normal);
		if (distanceFromPlane < 0.f) { distanceFromPlane *= -1.f; normal *= -1.f; }
		bool isInsideContactPlane = distanceFromPlane < radiusWithThreshold;
		if (isInsideContactPlane) {
			// pointInTriangle (barycentric variant, :274-296)
			V3 u = v[1] - v[0], vv = v[2] - v[0];
			V3 n = cross(u, vv);
			float nLenSq = dot(n, n);
			V3 wv = sphereCenter - v[0];
			float gamma = dot(cross(u, wv), n) / nLenSq;
			float beta = dot(cross(wv, vv), n) / nLenSq;
			float alpha = 1 - gamma - beta;
			bool inside = ((0 <= alpha) && (alpha <= 1) && (0 <= beta) && (beta <= 1) && (0 <= gamma) && (gamma <= 1));
			if (inside) {
				hasContact = true;
				contactPoint = sphereCenter - normal * distanceFromPlane;
			} else {
				float contactCapsuleRadiusSqr = radiusWithThreshold * radiusWithThreshold;
				// closestPointTriangle (:88-131)
				const V3 a = v[0], b = v[1], c = v[2], p = sphereCenter;
				const V3 ab = b - a, ac = c - a, ap = p - a;
				V3 q;
				const float d1 = dot(ab, ap), d2 = dot(ac, ap);
				bool done = false;
				if (d1 <= 0.f && d2 <= 0.f) { q = a; done = true; }
				float d3 = 0, d4 = 0, d5 = 0, d6 = 0;
				if (!done) { const V3 bp = p - b; d3 = dot(ab, bp); d4 = dot(ac, bp); if (d3 >= 0.f && d4 <= d3) { q = b; done = true; } }
				if (!done) { const V3 cp = p - c; d5 = dot(ab, cp); d6 = dot(ac, cp); if (d6 >= 0.f && d5 <= d6) { q = c; done = true; } }
				if (!done) { const float vc = d1 * d4 - d3 * d2; if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { const float t = d1 / (d1 - d3); q = a + t * ab; done = true; }
					if (!done) { const float vb = d5 * d2 - d1 * d6; if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { const float t = d2 / (d2 - d6); q = a + t * ac; done = true; }
						if (!done) { const float va = d3 * d6 - d5 * d4; if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) { const float t = (d4 - d3) / ((d4 - d3) + (d5 - d6)); q = b + t * (c - b); done = true; }
							if (!done) { const float denom = 1.f / (va + vb + vc); const float t = vb * denom; const float s = vc * denom; q = a + t * ab + s * ac; } } } }
				float distanceSqr = len2(q - sphereCenter);
				if (distanceSqr < contactCapsuleRadiusSqr) { hasContact = true; contactPoint = q; }
			}
		}
	}
	if (hasContact) {
		V3 contactToCentre = sphereCenter - contactPoint;
		float distanceSqr = len2(contactToCentre);
		if (distanceSqr < radiusWithThreshold * radiusWithThreshold) {
			if (distanceSqr > kEps) {
				float distance = sqrtf(distanceSqr);
				resultNormal = normalized(contactToCentre);
				point = contactPoint;
				depth = -(radius - distance);
			} else {
				resultNormal = normal;
				point = contactPoint;
				depth = -radius;
			}
			return true;
		}
	}
	return false;
}

RSB_HD bool triangle_vs_aabb(const V3 t[3], const V3& mn, const V3& mx) {  // TestTriangleAgainstAabb2 (Z, X, Y)
	const int order[3] = {2, 0, 1};
	for (int k = 0; k < 3; k++) {
		int i = order[k];
		if (fminf_(fminf_(t[0][i], t[1][i]), t[2][i]) > mx[i]) return false;
		if (fmaxf_(fmaxf_(t[0][i], t[1][i]), t[2][i]) < mn[i]) return false;
	}
	return true;
}

// ------------------------------------------------------------------ staged narrowphase: candidate records + detector queue
RSB_HD int cand_alloc(const Ctx& cx, int n) {
#if defined(__CUDA_ARCH__)
	const int i = atomicAdd(&cx.misc[3], n);
#else
	const int i = cx.misc[3]; cx.misc[3] += n;
#endif
	if (i + n > MAX_CANDS) { set_error(cx, ERR_TRI_OVERFLOW); return -1; }
	return i;
}
// A record that does not fit the executing block's queue simply stays CAND_PENDING: PH_NARROW_DETECT notices the overflow (the
// counter keeps counting) and sweeps the resident arenas' records for the stragglers after the queued ones are done.  (Running
// the detector right here instead was measured at -18 %: every emit site then carries an inlined GJK + EPA + box-box, and the
// record walk of PH_NARROW_STATIC pays for it in registers and spills whether the path is ever taken or not.)
RSB_HD void detect_queue_push(const Ctx& cx, int idx) {
#if defined(__CUDA_ARCH__)
	const unsigned q = atomicAdd(cx.nq, 1u);
#else
	const unsigned q = (*cx.nq)++;
#endif
	if (q < (unsigned)cx.qcap) cx.q[q] = (uint16_t)(cx.arLocal * 128 + idx);
}
RSB_HD void cand_emit(const Ctx& cx, int pair, int mesh, int tri, int state, const V3& normal, const V3& point, float depth) {
	const int i = cand_alloc(cx, 1);
	if (i < 0) return;
	CandRec r; r.pair = (int16_t)pair; r.mesh = (int16_t)mesh; r.tri = tri; r.state = state; r.depth = depth; r.normal = normal; r.point = point;
	cx.w->cand[i] = r;
	if (state == CAND_PENDING) detect_queue_push(cx, i);
}

RSB_HD void convex_local_aabb(const Ctx& cx, const Convex& shape, const Xform& xf, V3& amin, V3& amax) {
	if (shape.kind == SH_SPHERE) { float mg = shape.margin + 0.08f; V3 e(mg, mg, mg); amin = xf.origin - e; amax = xf.origin + e; }
	else if (shape.kind == SH_BOX) transform_aabb_he(shape.he, shape.margin, xf.basis, xf.origin, amin, amax);
	else transform_aabb_mm(cx.p->puckLocalAabbMin, cx.p->puckLocalAabbMax, shape.margin, xf.basis, xf.origin, amin, amax);
}

// Convex (ball sphere / puck hull / car box) vs an arena mesh:
// btConvexConcaveCollisionAlgorithm::processCollision + btConvexTriangleCallback::processTriangle
// (btConvexConcaveCollisionAlgorithm.cpp:71-250) over the BVH's AABB query
// (btQuantizedBvh.cpp:655-729: subtree headers in order, stackless walk inside each).
// The walk finds the triangles whose box overlaps the convex and that pass the supporting-plane early-out.  A sphere is
// resolved on the spot (closed form); any other shape gets one candidate record per triangle (DEFER), in leaf order.
template <bool DEFER>
RSB_HD void convex_vs_mesh(const Ctx& cx, int mi, const Convex& shape, const Xform& xf, int meshIdx, int entry) {
	const MeshD& mesh = cx.sc->meshes[meshIdx];
	Manifold& mf = cx.mf[mi];
	// convex AABB in triangle space (mesh at origin, identity): shape->getAabb(convexInTriangleSpace)
	V3 amin, amax;
	convex_local_aabb(cx, shape, xf, amin, amax);
	// extraMargin = mesh margin (0) + closestPointDistanceThreshold (0)
	uint16_t qmin[3], qmax[3];
	bvh_quantize_clamped(mesh, amin, 0, qmin);
	bvh_quantize_clamped(mesh, amax, 1, qmax);
	const M3 invB = inverse(xf.basis);
	// only the headers that can overlap a query from this broadphase cell (same order as the full header loop)
	const int hEnd = cx.sc->cellHdrStart[entry + 1];
	for (int hi = cx.sc->cellHdrStart[entry]; hi < hEnd; hi++) {
		const SubtreeHdr hd = mesh.hdrs[cx.sc->cellHdrs[hi]];
		if (!quant_overlap(qmin, qmax, hd.qmin, hd.qmax)) continue;
		int cur = hd.root;
		const int end = hd.root + hd.size;
		while (cur < end) {
			const BvhNode nd = mesh.nodes[cur];
			bool ov = quant_overlap(qmin, qmax, nd.qmin, nd.qmax);
			bool isLeaf = nd.idx >= 0;
			if (isLeaf && ov) {
				int tri = nd.idx & 0x1fffff;
				V3 t[3];
				mesh_triangle(mesh, tri, t[0], t[1], t[2]);
				if (triangle_vs_aabb(t, amin, amax)) {
					// early out on the triangle's supporting plane, both sides (:102-138)
					bool skip = false;
					{
						V3 tn = normalized(cross(t[1] - t[0], t[2] - t[0]));
						for (int side = 0; side < 2 && !skip; side++) {
							V3 localPt = support_with_margin(shape, invB * tn);
							V3 worldPt = xf_apply(xf, localPt);
							float proj_distPt = dot(tn, worldPt);
							float proj_distTr = dot(tn, t[0]);
							float dist = proj_distTr - proj_distPt;
							if (dist > mf.breakingThresh) skip = true;
							tn *= -1.f;
						}
					}
					if (!skip) {
						if (!DEFER) {
							V3 point, normal; float depth;
							// sphereInTr origin = (center - 0) * I
							V3 c = vmul(xf.origin - v3zero(), m3identity());
							if (sphere_triangle(c, shape.margin, t, mf.breakingThresh, point, normal, depth))
								add_contact_point(cx, mi, m3identity() * normal, m3identity() * point + v3zero(), depth, meshIdx, tri, false);
						} else {
							cand_emit(cx, mi, meshIdx, tri, CAND_PENDING, v3zero(), v3zero(), 0.f);
						}
					}
				}
			}
			if (ov || isLeaf) cur++;
			else cur += -nd.idx;
		}
	}
	if (!DEFER) refresh_contact_points(cx, mi);
}

// btConvexPlaneCollisionAlgorithm::processCollision (btConvexPlaneCollisionAlgorithm.cpp:92-161)
template <bool DEFER>
RSB_HD void convex_vs_plane(const Ctx& cx, int mi, const Convex& shape, const Xform& xf, const PlaneD& pl) {
	Manifold& mf = cx.mf[mi];
	const V3 planeNormal = pl.n;
	M3 planeInConvexBasis = transpose(xf.basis) * m3identity();
	V3 convexInPlaneOrigin = (m3identity() * xf.origin) + (m3identity() * (-pl.pos));
	M3 convexInPlaneBasis = m3identity() * xf.basis;
	V3 vtx = support_with_margin(shape, planeInConvexBasis * (-planeNormal));
	V3 vtxInPlane = convexInPlaneBasis * vtx + convexInPlaneOrigin;
	float distance = dot(planeNormal, vtxInPlane) - 0.f;
	bool hasCollision = distance < mf.breakingThresh;
	if (hasCollision) {
		V3 vtxInPlaneProjected = vtxInPlane - distance * planeNormal;
		V3 vtxInPlaneWorld = m3identity() * vtxInPlaneProjected + pl.pos;
		V3 normalOnSurfaceB = m3identity() * planeNormal;
		if (!DEFER) add_contact_point(cx, mi, normalOnSurfaceB, vtxInPlaneWorld, distance, -1, -1, false);
		else cand_emit(cx, mi, -1, -1, CAND_CONTACT, normalOnSurfaceB, vtxInPlaneWorld, distance);
	}
	if (!DEFER && mf.nPts) refresh_contact_points(cx, mi);
}

RSB_HD Convex car_box(const CarCfgD& cfg) {
	Convex c; c.kind = SH_BOX; c.margin = cfg.hitboxMargin; c.he = cfg.hitboxImplicit; c.pts = nullptr; c.npts = 0;
	return c;
}
RSB_HD Xform car_child_xform(const RigidS& rb, const CarCfgD& cfg) {
	Xform t; t.basis = rb.basis * m3identity(); t.origin = rb.basis * cfg.hitboxOffset + rb.pos;
	return t;
}
RSB_HD void convex_world_aabb(const Ctx& cx, const Convex& s, const Xform& t, V3& mn, V3& mx) { convex_local_aabb(cx, s, t, mn, mx); }

// performDiscreteCollisionDetection in four phases (rsb_tick.h):
//   collide_pairs            PH_PAIRS           (arena)   cell bookkeeping + broadphase pair list; manifold slot k <-> pair k
//   collide_static_pairs     PH_NARROW_STATIC   (body b)  b's pairs with static objects: the sphere ball is resolved on the spot
//                                                         (closed-form detectors + contact hooks); a car box / the puck hull only
//                                                         records candidates
//   enqueue_dynamic_pairs    PH_NARROW_STATIC   (arena)   ball-car and car-car pairs that pass the AABB checks -> candidates
//   detect_candidate         PH_NARROW_DETECT   (record)  GJK(+EPA) / box-box for one pending record
//   commit_static_candidates PH_NARROW_COMMIT   (body b)  b's records -> manifolds + contact hooks, in record order
//   commit_dynamic_candidates PH_DYN_SOLVE      (arena)   same for ball-car / car-car (their hooks touch two bodies)
// Manifold order (= solver row order) is the pair order whatever the execution order; within a manifold the
// points arrive in BVH leaf order / detector order, as in the reference.
RSB_HD void collide_pairs(const Ctx& cx) {
	RSB_SUBPROF_BEGIN();
	apply_cell_changes(cx);
	RSB_SUBPROF_MARK(21);
	compute_pairs(cx);
	RSB_SUBPROF_MARK(22);
	cx.pl->nManifolds = cx.pl->nPairs;
	cx.pl->deferMask = 0u; cx.pl->mfMask = 0u;
	cx.misc[3] = 0;
	for (int k = 0; k < cx.pl->nPairs; k++) { cx.mf[k].body0 = -1; cx.mf[k].nPts = 0; }
}

RSB_HD bool pair_needs_collision(const Ctx& cx, int a, int b) {
	// btCollisionDispatcher::needsCollision: skip when both are inactive (statics are inactive)
	bool actA = a >= STATIC_BASE ? false : body_active(cx.body[a]);
	bool actB = body_active(cx.body[b]);
	return actA || actB;
}

RSB_HD void mark_deferred(const Ctx& cx, int k) {
#if defined(__CUDA_ARCH__)
	atomicOr(&cx.pl->deferMask, 1u << k);
#else
	cx.pl->deferMask |= 1u << k;
#endif
}

RSB_HD void collide_static_pairs(const Ctx& cx, int b) {
	const BatchParams& P = *cx.p;
	const Scene& sc = *cx.sc;
	const int nPairs = cx.pl->nPairs;
#ifndef RSB_SPHERE_INLINE
#define RSB_SPHERE_INLINE 0  // developer A/B: 1 = round-1 behaviour, the sphere ball resolves its triangles inside the walk
#endif
	// The sphere ball used to run its closed-form sphere-triangle tests inside the walk, one after the other on the ball's thread --
	// next to a ramp that thread was the slowest of the whole phase.  It now records its triangles like every other shape and
	// the tests run one per thread in PH_NARROW_DETECT.
	const bool inlineBody = RSB_SPHERE_INLINE && (b == 0) && cx.ballIsSphere;
	RSB_SUBPROF_BEGIN();
	for (int k = 0; k < nPairs; k++) {
		if (cx.pl->b[k] != b) continue;
		const int a = cx.pl->a[k];
		if (a < STATIC_BASE) continue;
		if (!pair_needs_collision(cx, a, b)) continue;
		const int sk = a - STATIC_BASE;
		const bool isMesh = sk < sc.nMeshes;
		const RigidS& rb = rigid_of(cx, b);
		if (cx.nTiles && static_tile_index(sc, sk) >= 0) continue;  // (tile, ball): a pending record of the arena item (enqueue_dynamic_pairs)
		if (inlineBody) {
			Convex shape = ball_convex(cx);
			Xform xf; xf.basis = rb.basis; xf.origin = rb.pos;
			int mi = new_manifold(cx, k, 0, a);
			if (isMesh) convex_vs_mesh<false>(cx, mi, shape, xf, sk, cx.pl->entry[k]);
			else convex_vs_plane<false>(cx, mi, shape, xf, sc.planes[sk - sc.nMeshes]);
			continue;
		}
		Convex shape; Xform xf;
		if (b == 0) { shape = ball_convex(cx); xf.basis = rb.basis; xf.origin = rb.pos; }
		else {
			// btCompoundLeafCallback::ProcessChildShape: child AABB vs the other shape's AABB
			V3 omn, omx;
			static_aabb(cx, a, omn, omx);
			if (!aabb_overlap(cx.body[b].childMin, cx.body[b].childMax, omn, omx)) continue;
			shape = car_box(P.carCfg[b - 1]);
			xf = cx.body[b].childXf;
		}
		int mi = new_manifold(cx, k, b, a);
		mark_deferred(cx, k);
		RSB_SUBPROF_MARK(12);
		if (isMesh) convex_vs_mesh<true>(cx, mi, shape, xf, sk, cx.pl->entry[k]);
		else convex_vs_plane<true>(cx, mi, shape, xf, sc.planes[sk - sc.nMeshes]);
		RSB_SUBPROF_MARK(isMesh ? 14 : 13);
	}
}

// ball-car and car-car pairs, in pair order: the compound / child AABB gates of the reference, a manifold, and a pending record.
// Dropshot (tile, ball) pairs travel the same way -- one GJK query per pair, committed with the dynamic pairs in PH_DYN_SOLVE, in
// pair order: the ball's tile pairs precede its car pairs, as in the reference's pair array, so a tile hit sees the charge level
// from before this tick's car hits.  (Sharing the ball-car emit site also keeps the queue-full fallback at one inlined detector.)
RSB_HD void enqueue_dynamic_pairs(const Ctx& cx) {
	const BatchParams& P = *cx.p;
	const bool hasTiles = cx.nTiles != 0;
	for (int k = 0; k < cx.pl->nPairs; k++) {
		const int a = cx.pl->a[k], b = cx.pl->b[k];
		const bool tilePair = hasTiles && a >= STATIC_BASE && static_tile_index(*cx.sc, a - STATIC_BASE) >= 0;
		if (a >= STATIC_BASE && !tilePair) continue;
		if (!pair_needs_collision(cx, a, b)) continue;
		if (a == 0 || tilePair) {
			if (!tilePair) {
				// (ball, car): swapped compound -> convex-convex(box of the car, ball shape); manifold (car, ball)
				const RigidS& rball = rigid_of(cx, 0);
				Convex ball = ball_convex(cx); Xform xfb; xfb.basis = rball.basis; xfb.origin = rball.pos;
				V3 omn, omx;
				convex_world_aabb(cx, ball, xfb, omn, omx);
				if (!aabb_overlap(cx.body[b].childMin, cx.body[b].childMax, omn, omx)) continue;
			}
			// (tile, ball): btConvexConvexAlgorithm on (hull, sphere) in broadphase order -> manifold (tile, ball)
			new_manifold(cx, k, tilePair ? a : b, tilePair ? b : 0);
			mark_deferred(cx, k);
			cand_emit(cx, k, -1, -1, CAND_PENDING, v3zero(), v3zero(), 0.f);
		} else {
			// (carLo=a, carHi=b): compound-compound -> ... -> box-box(body0 = box of b, body1 = box of a); manifold (b, a)
			const CarCfgD& cfgB = P.carCfg[b - 1];
			const RigidS& rbb = rigid_of(cx, b);
			const V3 amn = cx.body[a].childMin, amx = cx.body[a].childMax;   // child of A
			V3 cmn, cmx;
			compound_aabb(cfgB, rbb.basis, rbb.pos, cmn, cmx);  // whole compound B
			if (!aabb_overlap(amn, amx, cmn, cmx)) continue;
			if (!aabb_overlap(cx.body[b].childMin, cx.body[b].childMax, amn, amx)) continue;   // child of B vs box A
			new_manifold(cx, k, b, a);
			mark_deferred(cx, k);
			// the box-box detector reports up to 4 points: 4 consecutive records, the first one carries the request
			const int i = cand_alloc(cx, 4);
			if (i < 0) continue;
			for (int j = 0; j < 4; j++) {
				CandRec r; r.pair = (int16_t)k; r.mesh = -1; r.tri = -1; r.state = j == 0 ? CAND_PENDING : CAND_NONE; r.depth = 0.f; r.normal = v3zero(); r.point = v3zero();
				cx.w->cand[i + j] = r;
			}
			detect_queue_push(cx, i);
		}
	}
}

// (Dropshot tile hull = A, ball = B): convex-convex keeps the broadphase order (btConvexConvexAlgorithm.cpp:268-276).  Out of line
// on the device: the two pair-detector call sites of detect_candidate are specialised for their shape kinds by inlining; a third
// inlined copy would add 22k instructions to every mode's kernel for a pair that only exists in Dropshot.
#ifndef RSB_TILE_INLINE
#define RSB_TILE_INLINE 0
#endif
#if RSB_TILE_INLINE
RSB_HD
#else
RSB_HD_NOINLINE
#endif
void detect_tile_ball(Ctx cx, int idx) {
	const BatchParams& P = *cx.p;
	CandRec* rec = &cx.w->cand[idx];
	const int k = rec->pair;
	const int a = cx.pl->a[k], b = cx.pl->b[k];
	const TileD& td = cx.sc->tiles[static_tile_index(*cx.sc, a - STATIC_BASE)];
	Convex hull; hull.kind = SH_HULL; hull.margin = 0.04f; hull.he = v3zero(); hull.pts = td.pts; hull.npts = 6;
	Xform xfT; xfT.basis = m3identity(); xfT.origin = static_origin(cx, a);
	const RigidS& rball = rigid_of(cx, b);
	Convex ball = ball_convex(cx); Xform xfb; xfb.basis = rball.basis; xfb.origin = rball.pos;
	V3 n, pt; float d;
	const bool hit = convex_convex_detect(cx, cx.mf[k].breakingThresh, hull, xfT, ball, xfb, n, pt, d);
	if (hit) { rec->normal = n; rec->point = pt; rec->depth = d; }
	rec->state = hit ? CAND_CONTACT : CAND_NONE;
}

// PH_NARROW_DETECT: the detector of one pending record, on its own thread
RSB_HD void detect_candidate(const Ctx& cx, int idx) {
	const BatchParams& P = *cx.p;
	CandRec* rec = &cx.w->cand[idx];
	const int k = rec->pair;
	const int a = cx.pl->a[k], b = cx.pl->b[k];
	V3 n, pt; float d;
	if (a >= STATIC_BASE) {
		if (cx.nTiles && static_tile_index(*cx.sc, a - STATIC_BASE) >= 0) { detect_tile_ball(cx, idx); return; }
		// (body b, triangle of an arena mesh)
		const RigidS& rb = rigid_of(cx, b);
		if (b == 0 && cx.ballIsSphere) {
			// sphere vs triangle in closed form (SphereTriangleDetector, the reference's choice for this shape pair)
			V3 t[3];
			mesh_triangle(cx.sc->meshes[rec->mesh], rec->tri, t[0], t[1], t[2]);
			const V3 c = vmul(rb.pos - v3zero(), m3identity());
			const bool hit = sphere_triangle(c, P.ballRadiusBT, t, cx.mf[k].breakingThresh, pt, n, d);
			if (hit) { rec->normal = m3identity() * n; rec->point = m3identity() * pt + v3zero(); rec->depth = d; }
			rec->state = hit ? CAND_CONTACT : CAND_NONE;
			return;
		}
		Convex shape; Xform xf;
		if (b == 0) { shape = ball_convex(cx); xf.basis = rb.basis; xf.origin = rb.pos; }
		else { shape = car_box(P.carCfg[b - 1]); xf = cx.body[b].childXf; }
		Convex triShape; triShape.kind = SH_TRI; triShape.margin = 0.f; triShape.he = v3zero(); triShape.pts = nullptr; triShape.npts = 0;
		mesh_triangle(cx.sc->meshes[rec->mesh], rec->tri, triShape.tri[0], triShape.tri[1], triShape.tri[2]);
		Xform xfT; xfT.basis = m3identity(); xfT.origin = v3zero();
		const bool hit = convex_convex_detect(cx, cx.mf[k].breakingThresh, shape, xf, triShape, xfT, n, pt, d);
		if (hit) { rec->normal = n; rec->point = pt; rec->depth = d; }
		rec->state = hit ? CAND_CONTACT : CAND_NONE;
	} else if (a == 0) {
		const RigidS& rball = rigid_of(cx, 0);
		Convex ball = ball_convex(cx); Xform xfb; xfb.basis = rball.basis; xfb.origin = rball.pos;
		Convex box = car_box(P.carCfg[b - 1]); const Xform xfc = cx.body[b].childXf;
		const bool hit = convex_convex_detect(cx, cx.mf[k].breakingThresh, box, xfc, ball, xfb, n, pt, d);
		if (hit) { rec->normal = n; rec->point = pt; rec->depth = d; }
		rec->state = hit ? CAND_CONTACT : CAND_NONE;
	} else {
		Convex boxA = car_box(P.carCfg[a - 1]), boxB = car_box(P.carCfg[b - 1]);
		const Xform xfA = cx.body[a].childXf, xfB = cx.body[b].childXf;
		V3 pts[4]; float dep[4];
		const int nOut = box_box_detect(cx, boxB, xfB, boxA, xfA, n, pts, dep);
		for (int j = 0; j < 4; j++) {
			if (j < nOut) { rec[j].normal = n; rec[j].point = pts[j]; rec[j].depth = dep[j]; }
			rec[j].state = j < nOut ? CAND_CONTACT : CAND_NONE;
		}
	}
}

// PH_NARROW_COMMIT (body b) / PH_DYN_SOLVE (dynamic pairs, b = -1): records -> manifolds + contact hooks, in record order.
// The records of one thread's walk carry ascending indices, so one pass in index order filtered by owner restores the order
// in which the reference's detectors add their points.
RSB_HD void commit_candidates(const Ctx& cx, int b) {
	const int n = cx.misc[3] < MAX_CANDS ? cx.misc[3] : MAX_CANDS;
	const uint32_t deferred = cx.pl->deferMask;
	int curPair = -1;
	for (int i = 0; i < n; i++) {
		const CandRec r = cx.w->cand[i];
		const int k = r.pair;
		const bool dyn = cx.pl->a[k] < STATIC_BASE || (cx.nTiles && static_tile_index(*cx.sc, cx.pl->a[k] - STATIC_BASE) >= 0);  // + (tile, ball)
		if (b < 0 ? !dyn : (dyn || cx.pl->b[k] != b)) continue;
		if (k != curPair) { if (curPair >= 0) refresh_contact_points(cx, curPair); curPair = k; }
		if (r.state != CAND_CONTACT) continue;
		if (b < 0 && cx.pl->a[k] >= STATIC_BASE) add_contact_point_tile(cx, k, r.normal, r.point, r.depth);
		else if (b < 0) add_contact_point(cx, k, r.normal, r.point, r.depth, -1, -1, true);
		else add_contact_point(cx, k, r.normal, r.point, r.depth, r.mesh, r.tri, false);
	}
	if (curPair >= 0) refresh_contact_points(cx, curPair);
	(void)deferred;
}

}  // namespace rsb
