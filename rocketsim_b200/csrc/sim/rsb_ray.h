// rsb_ray.h -- the suspension ray query: closest hit of a short segment against the arena.
// Restates the query stack the reference runs for each wheel:
//   btDefaultVehicleRaycaster::castRay      bt/BulletDynamics/Vehicle/btDefaultVehicleRaycaster.cpp:32-53
//   btRSBroadphase::rayTest (ray-origin cell) bt/BulletCollision/BroadphaseCollision/btRSBroadphase.cpp:325-357
//   rayTestSingleInternal                   bt/BulletCollision/CollisionDispatch/btCollisionWorld.cpp:267-547
//   quantized BVH ray walk                  bt/BulletCollision/BroadphaseCollision/btQuantizedBvh.cpp:531-652
//   triangle ray test                       bt/BulletCollision/NarrowPhaseCollision/btRaycastCallback.cpp:35-120
//   plane as two triangles                  bt/BulletCollision/CollisionShapes/btStaticPlaneShape.cpp:56-82
//   ray vs convex (sub-simplex cast)        bt/BulletCollision/NarrowPhaseCollision/btSubSimplexConvexCast.cpp:31-142
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"

namespace rsb {

struct RayHit { int obj; float fraction; V3 point, normal; };

// collision filter groups (btBroadphaseProxy::CollisionFilterGroups + src/Sim/CollisionMasks.h)
constexpr int FILT_DEFAULT = 1, FILT_STATIC = 2, FILT_ALL = -1;
constexpr int FILT_HOOPS_NET = (1 << 8), FILT_DROPSHOT_TILE = (1 << 9), FILT_DROPSHOT_FLOOR = (1 << 10);

RSB_HD int grid_cell_of(const Scene& sc, const V3& pos, int& i, int& j, int& k) {
	V3 f = (pos - sc.gridMin) / sc.cellSize;
	i = (int)f.x; j = (int)f.y; k = (int)f.z;
	i = i < 0 ? 0 : (i > sc.cellsX - 1 ? sc.cellsX - 1 : i);
	j = j < 0 ? 0 : (j > sc.cellsY - 1 ? sc.cellsY - 1 : j);
	k = k < 0 ? 0 : (k > sc.cellsZ - 1 ? sc.cellsZ - 1 : k);
	return i * sc.cellsY * sc.cellsZ + j * sc.cellsZ + k;
}

// btTriangleRaycastCallback::processTriangle; updates hitFraction/normal when closer.
RSB_HD bool ray_triangle(const V3& from, const V3& to, const V3& vert0, const V3& vert1, const V3& vert2, float& hitFraction, V3& outNormal) {
	V3 v10 = vert1 - vert0, v20 = vert2 - vert0;
	V3 triangleNormal = cross(v10, v20);
	const float dist = dot(vert0, triangleNormal);
	float dist_a = dot(triangleNormal, from); dist_a -= dist;
	float dist_b = dot(triangleNormal, to); dist_b -= dist;
	if (dist_a * dist_b >= 0.f) return false;
	const float proj_length = dist_a - dist_b;
	const float distance = dist_a / proj_length;
	if (distance < hitFraction) {
		float edge_tolerance = len2(triangleNormal);
		edge_tolerance *= -0.0001f;
		V3 point = lerp(from, to, distance);
		V3 v0p = vert0 - point, v1p = vert1 - point;
		V3 cp0 = cross(v0p, v1p);
		if (dot(cp0, triangleNormal) >= edge_tolerance) {
			V3 v2p = vert2 - point;
			V3 cp1 = cross(v1p, v2p);
			if (dot(cp1, triangleNormal) >= edge_tolerance) {
				V3 cp2 = cross(v2p, v0p);
				if (dot(cp2, triangleNormal) >= edge_tolerance) {
					triangleNormal = normalized(triangleNormal);
					outNormal = (dist_a <= 0.f) ? -triangleNormal : triangleNormal;
					hitFraction = distance;
					return true;
				}
			}
		}
	}
	return false;
}

RSB_HD void mesh_triangle(const MeshD& m, int tri, V3& a, V3& b, V3& c) {
	const int32_t* ix = m.tris + 3 * tri;
	const float* pa = m.verts + 4 * ix[0];
	const float* pb = m.verts + 4 * ix[1];
	const float* pc = m.verts + 4 * ix[2];
	a = V3(pa[0], pa[1], pa[2]); b = V3(pb[0], pb[1], pb[2]); c = V3(pc[0], pc[1], pc[2]);
}

// btQuantizedBvh::quantize / quantizeWithClamp / unQuantize (btQuantizedBvh.h:331-415)
RSB_HD void bvh_quantize_clamped(const MeshD& m, const V3& p, int isMax, uint16_t out[3]) {
	V3 cp = vmin(vmax(p, m.bvhMin), m.bvhMax);
	V3 v = (cp - m.bvhMin) * m.bvhQuant;
	if (isMax) {
		out[0] = (uint16_t)(((uint16_t)(v.x + 1.f)) | 1);
		out[1] = (uint16_t)(((uint16_t)(v.y + 1.f)) | 1);
		out[2] = (uint16_t)(((uint16_t)(v.z + 1.f)) | 1);
	} else {
		out[0] = (uint16_t)(((uint16_t)(v.x)) & 0xfffe);
		out[1] = (uint16_t)(((uint16_t)(v.y)) & 0xfffe);
		out[2] = (uint16_t)(((uint16_t)(v.z)) & 0xfffe);
	}
}
RSB_HD V3 bvh_unquantize(const MeshD& m, const uint16_t q[3]) {
	return V3((float)q[0] / m.bvhQuant.x, (float)q[1] / m.bvhQuant.y, (float)q[2] / m.bvhQuant.z) + m.bvhMin;
}
RSB_HD bool quant_overlap(const uint16_t* min1, const uint16_t* max1, const uint16_t* min2, const uint16_t* max2) {
	return (min1[0] <= max2[0]) & (max1[0] >= min2[0]) & (min1[2] <= max2[2]) & (max1[2] >= min2[2]) & (min1[1] <= max2[1]) & (max1[1] >= min2[1]);
}

// btRayAabb2 (btAabbUtil2.h:82-115)
RSB_HD bool ray_aabb2(const V3& rayFrom, const V3& invDir, const unsigned sign[3], const V3 bounds[2], float lambda_min, float lambda_max) {
	float tmin = (bounds[sign[0]].x - rayFrom.x) * invDir.x;
	float tmax = (bounds[1 - sign[0]].x - rayFrom.x) * invDir.x;
	float tymin = (bounds[sign[1]].y - rayFrom.y) * invDir.y;
	float tymax = (bounds[1 - sign[1]].y - rayFrom.y) * invDir.y;
	if ((tmin > tymax) || (tymin > tmax)) return false;
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (bounds[sign[2]].z - rayFrom.z) * invDir.z;
	float tzmax = (bounds[1 - sign[2]].z - rayFrom.z) * invDir.z;
	if ((tmin > tzmax) || (tzmin > tmax)) return false;
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	return (tmin < lambda_max) && (tmax > lambda_min);
}

// walkStacklessQuantizedTreeAgainstRay + triangle tests.  The reference walks the whole node array from node 0; a
// subtree is only entered through ancestors whose boxes contain it, so walking just the candidate subtrees of the
// ray-origin cell (Scene::cellHdrsRay, ascending root = the same leaf order) visits the same leaves.
RSB_HD bool ray_mesh(const Scene& sc, const MeshD& m, int entry, const V3& from, const V3& to, float& hitFraction, V3& outNormal) {
	V3 rayDirection = safe_normalized(to - from);
	float lambda_max = dot(rayDirection, to - from);
	V3 inv(rayDirection.x == 0.f ? kLargeFloat : 1.f / rayDirection.x,
	       rayDirection.y == 0.f ? kLargeFloat : 1.f / rayDirection.y,
	       rayDirection.z == 0.f ? kLargeFloat : 1.f / rayDirection.z);
	unsigned sign[3] = {inv.x < 0.f, inv.y < 0.f, inv.z < 0.f};
	V3 rayAabbMin = vmin(from, to), rayAabbMax = vmax(from, to);
	uint16_t qmin[3], qmax[3];
	bvh_quantize_clamped(m, rayAabbMin, 0, qmin);
	bvh_quantize_clamped(m, rayAabbMax, 1, qmax);
	bool any = false;
	if (lambda_max <= sc.rayLeafReach) {
		// short ray (every suspension ray): the leaves this cell's rays can reach, in leaf order, each gated by its OWN box tests
		// -- the same leaves the hierarchical walk reports (see Scene::cellRayLeaves), without the dependent node-to-node loads
		const int lEnd = sc.cellRayLeafStart[entry + 1];
		for (int li = sc.cellRayLeafStart[entry]; li < lEnd; li++) {
			const int cur = sc.cellRayLeaves[li];
			const BvhNode nd = m.nodes[cur];
			if (!quant_overlap(qmin, qmax, nd.qmin, nd.qmax)) continue;
			const float* nb = m.nodeBounds + 8 * (size_t)cur;
			V3 bounds[2];
			bounds[0] = V3(nb[0], nb[1], nb[2]);
			bounds[1] = V3(nb[4], nb[5], nb[6]);
			if (!ray_aabb2(from, inv, sign, bounds, 0.0f, lambda_max)) continue;
			V3 a, b, c;
			mesh_triangle(m, nd.idx & 0x1fffff, a, b, c);
			if (ray_triangle(from, to, a, b, c, hitFraction, outNormal)) any = true;
		}
		return any;
	}
	const int hEnd = sc.cellHdrStart[entry + 1];
	for (int hi = sc.cellHdrStart[entry]; hi < hEnd; hi++) {
	const SubtreeHdr hd = m.hdrs[sc.cellHdrsRay[hi]];
	if (!quant_overlap(qmin, qmax, hd.qmin, hd.qmax)) continue;
	int cur = hd.root;
	const int end = hd.root + hd.size;
	while (cur < end) {
		const BvhNode nd = m.nodes[cur];
		bool rayBoxOverlap = false;
		bool boxBoxOverlap = quant_overlap(qmin, qmax, nd.qmin, nd.qmax);
		bool isLeaf = nd.idx >= 0;
		if (boxBoxOverlap) {
			// unQuantize(node corners): six IEEE divisions per node in the reference, tabulated at scene build (MeshD::nodeBounds)
			const float* nb = m.nodeBounds + 8 * (size_t)cur;
			V3 bounds[2];
			bounds[0] = V3(nb[0], nb[1], nb[2]);
			bounds[1] = V3(nb[4], nb[5], nb[6]);
			rayBoxOverlap = ray_aabb2(from, inv, sign, bounds, 0.0f, lambda_max);
		}
		if (isLeaf && rayBoxOverlap) {
			V3 a, b, c;
			mesh_triangle(m, nd.idx & 0x1fffff, a, b, c);
			if (ray_triangle(from, to, a, b, c, hitFraction, outNormal)) any = true;
		}
		if (rayBoxOverlap || isLeaf) cur++;
		else cur += -nd.idx;
	}
	}
	return any;
}

// Static plane through the generic concave path (btCollisionWorld.cpp:383-434).
RSB_HD bool ray_plane(const PlaneD& pl, const V3& from, const V3& to, float& hitFraction, V3& outNormal) {
	V3 fromL = from + (-pl.pos), toL = to + (-pl.pos);
	V3 aabbMin = vmin(fromL, toL), aabbMax = vmax(fromL, toL);
	V3 halfExtents = (aabbMax - aabbMin) * 0.5f;
	float radius = len(halfExtents);
	V3 center = (aabbMax + aabbMin) * 0.5f;
	V3 t0, t1;
	plane_space1(pl.n, t0, t1);
	V3 projectedCenter = center - (dot(pl.n, center) - 0.f) * pl.n;
	V3 tri[3];
	bool any = false;
	tri[0] = projectedCenter + t0 * radius + t1 * radius;
	tri[1] = projectedCenter + t0 * radius - t1 * radius;
	tri[2] = projectedCenter - t0 * radius - t1 * radius;
	if (ray_triangle(fromL, toL, tri[0], tri[1], tri[2], hitFraction, outNormal)) any = true;
	tri[0] = projectedCenter - t0 * radius - t1 * radius;
	tri[1] = projectedCenter - t0 * radius + t1 * radius;
	tri[2] = projectedCenter + t0 * radius + t1 * radius;
	if (ray_triangle(fromL, toL, tri[0], tri[1], tri[2], hitFraction, outNormal)) any = true;
	return any;
}

// btSubsimplexConvexCast::calcTimeOfImpact with A = point moving from->to, B = static convex.
RSB_HD_NOINLINE bool ray_convex(Convex shapeB, Xform xfB, V3 from, V3 to, float& hitFraction, V3& outNormal) {
	Simplex sx;
	sx.reset();
	V3 linVelA = to - from;
	const V3 originB0 = xfB.origin;
	float lambda = 0.f;
	V3 originA = from;
	V3 r = linVelA;  // linVelB = 0
	V3 supA = originA;
	V3 supB = xf_apply(xfB, support_with_margin(shapeB, vmul(r, xfB.basis)));
	V3 v = supA - supB;
	int maxIter = 32;
	V3 n = v3zero();
	float dist2 = len2(v);
	V3 w;
	while ((dist2 > 0.0001f) && maxIter--) {
		supA = originA;
		supB = xf_apply(xfB, support_with_margin(shapeB, vmul(v, xfB.basis)));
		w = supA - supB;
		float VdotW = dot(v, w);
		if (lambda > 1.0f) return false;
		if (VdotW > 0.f) {
			float VdotR = dot(v, r);
			if (VdotR >= -(kEps * kEps)) return false;
			lambda = lambda - VdotW / VdotR;
			originA = lerp(from, to, lambda);
			// interpolatedTransB is re-interpolated too, between two IDENTICAL origins (fromB == toB): (1 - l) * o + l * o is not
			// always o in fp32, and the reference casts against that rounded origin from here on (btSubSimplexConvexCast.cpp:97-98)
			xfB.origin = lerp(originB0, originB0, lambda);
			w = supA - supB;
			n = v;
		}
		if (!sx.in_simplex(w)) sx.add_vertex(w, supA, supB);
		if (simplex_closest(sx, v)) dist2 = len2(v);
		else dist2 = 0.f;
	}
	V3 normal = (len2(n) >= (kEps * kEps)) ? normalized(n) : v3zero();
	if (dot(normal, r) >= -0.f) return false;
	// rayTestSingleInternal acceptance (btCollisionWorld.cpp:292-314)
	if (len2(normal) > 0.0001f) {
		if (lambda < hitFraction) {
			outNormal = normalized(normal);
			hitFraction = lambda;
			return true;
		}
	}
	return false;
}

RSB_HD Convex ball_convex(const Ctx& cx) {
	const BatchParams& P = *cx.p;
	Convex c;
	if (cx.ballIsSphere) { c.kind = SH_SPHERE; c.margin = P.ballRadiusBT; c.he = v3zero(); c.pts = nullptr; c.npts = 0; }
	else { c.kind = SH_HULL; c.margin = 0.04f; c.he = v3zero(); c.pts = P.puckPts; c.npts = PUCK_POINTS; }
	return c;
}

// Full query for one wheel of dynamic body `self` (a car).
RSB_HD bool cast_vehicle_ray(const Ctx& cx, int self, const V3& from, const V3& to, RayHit& out) {
	const BatchParams& P = *cx.p;
	float closest = 1.0f;
	int obj = OBJ_NONE;
	V3 normal = v3zero();
	const int rayGroup = FILT_DEFAULT | FILT_DROPSHOT_FLOOR;
	int ci = 0, cj = 0, ck = 0;
	const Scene& sc = *cx.sc;
	RSB_SUBPROF_BEGIN();
	{
		int cell = grid_cell_of(sc, from, ci, cj, ck);
		for (int s = sc.cellStart[cell]; s < sc.cellStart[cell + 1]; s++) {
			if (closest == 0.f) break;
			int k = sc.cellStatics[s];
			if (k < sc.nMeshes) {
				const MeshD& m = sc.meshes[k];
				if (!((m.group & FILT_ALL) != 0 && (rayGroup & m.mask))) continue;
				// conservative cull (ours): every triangle lies inside the mesh AABB, so a segment whose box misses it by more
				// than the triangle test's edge tolerance cannot be reported by the BVH walk
				{
					const float slack = 0.1f;
					const V3 smn = vmin(from, to), smx = vmax(from, to);
					if (smn.x > m.aabbMax.x + slack || smx.x < m.aabbMin.x - slack || smn.y > m.aabbMax.y + slack || smx.y < m.aabbMin.y - slack ||
					    smn.z > m.aabbMax.z + slack || smx.z < m.aabbMin.z - slack) continue;
				}
				if (ray_mesh(sc, m, s, from, to, closest, normal)) obj = STATIC_BASE + k;
			} else if (k < sc.nMeshes + sc.nPlanes) {
				const PlaneD& pl = sc.planes[k - sc.nMeshes];
				if (!((pl.group & FILT_ALL) != 0 && (rayGroup & pl.mask))) continue;
				// conservative cull (ours): the triangle test only reports a hit when the end points lie on opposite sides of the
				// triangle's plane, which is the static plane up to rounding (< 1e-4 bt at arena scale); 1e-2 bt of slack
				const float da = dot(pl.n, from - pl.pos), db = dot(pl.n, to - pl.pos);
				if ((da > 1e-2f && db > 1e-2f) || (da < -1e-2f && db < -1e-2f)) continue;
				if (ray_plane(pl, from, to, closest, normal)) obj = STATIC_BASE + k;
			}
		}
	}
	RSB_SUBPROF_MARK(9);
	// dynamic proxies registered in the ray-origin cell, in cell-list order (= ascending stamp)
	{
		const int nb = n_bodies(cx);
		for (int q = 0; q < nb; q++) {
			const int b = cx.s->dynOrder[q];
			if (b == self) continue;
			if (closest == 0.f) break;
			if (abs((int)cx.s->cellI[b] - ci) > 1 || abs((int)cx.s->cellJ[b] - cj) > 1 || abs((int)cx.s->cellK[b] - ck) > 1) continue;
			const RigidS& rb = rigid_of(cx, b);
			Xform xf; Convex shp;
			{
				// conservative cull (ours, not in the reference): the cast can only hit when the segment passes within the
				// body's bounding sphere; 0.25 bt of slack covers the cast's own 1e-2 termination tolerance many times over
				const float R = P.boundRadius[b] + 0.25f;
				const V3 centre = (b == 0) ? rb.pos : (rb.basis * P.carCfg[b - 1].hitboxOffset + rb.pos);
				const V3 d = to - from, m = centre - from;
				const float dd = dot(d, d);
				float t = dd > 0.f ? dot(m, d) / dd : 0.f;
				t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
				const V3 c = m - d * t;
				if (dot(c, c) > R * R) continue;
			}
			if (b == 0) {
				xf.basis = rb.basis; xf.origin = rb.pos;
				shp = ball_convex(cx);
			} else {
				const CarCfgD& cfg = P.carCfg[b - 1];
				xf.basis = rb.basis * m3identity();
				xf.origin = rb.basis * cfg.hitboxOffset + rb.pos;
				shp.kind = SH_BOX; shp.margin = cfg.hitboxMargin; shp.he = cfg.hitboxImplicit; shp.pts = nullptr; shp.npts = 0;
			}
			if (ray_convex(shp, xf, from, to, closest, normal)) obj = b;
		}
	}
	RSB_SUBPROF_MARK(10);
	if (obj == OBJ_NONE) return false;
	if (obj < STATIC_BASE && cx.body[obj].noResponse) return false;
	out.obj = obj;
	out.fraction = closest;
	out.point = lerp(from, to, closest);
	out.normal = normalized(normal);
	return true;
}

}  // namespace rsb
