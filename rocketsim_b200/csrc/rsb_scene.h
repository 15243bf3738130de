// rsb_scene.h -- HOST-side, load-time half of the hot path: turns the .cmf collision meshes of a
// game mode into the flattened static scene the kernels traverse:
//   .cmf parse                          reference src/CollisionMeshFile/CollisionMeshFile.cpp:11-61
//   quantised BVH (node order exact)    bt/BulletCollision/CollisionShapes/btOptimizedBvh.cpp:28-180,
//                                       bt/BulletCollision/BroadphaseCollision/btQuantizedBvh.cpp:76-305
//   internal-edge (triangle info) map   bt/BulletCollision/CollisionDispatch/btInternalEdgeUtility.cpp:47-353
//   arena planes                        reference src/Sim/Arena/Arena.cpp:1030-1068
//   broadphase static cell lists        bt/BulletCollision/BroadphaseCollision/btRSBroadphase.cpp:43-182
// The reference rebuilds the cell lists for every Arena (18k brute-force mesh scans each);
// here they are built once per game mode and shared by every arena of every batch.
#pragma once
#include <algorithm>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sim/rsb_state.h"
#include "sim/rsb_ray.h"

namespace rsb {

struct MeshHost {
	std::vector<int32_t> tris;     // 3 per triangle
	std::vector<float> verts;      // 4 per vertex
	std::vector<BvhNode> nodes;
	std::vector<float> nodeBounds;  // 8 floats per node: unquantize(qmin), 0, unquantize(qmax), 0 (the ray walk's box, rsb_ray.h)
	std::vector<SubtreeHdr> hdrs;
	std::vector<TriInfo> info;
	V3 bvhMin, bvhMax, bvhQuant;
	V3 aabbMin, aabbMax;
	int group = 0, mask = 0;
	int nTris() const { return (int)tris.size() / 3; }
	int nVerts() const { return (int)verts.size() / 4; }
};

struct SceneHost {
	int mode = THE_VOID;
	std::vector<MeshHost> meshes;
	std::vector<PlaneD> planes;
	std::vector<TileD> tiles;  // Dropshot floor tiles, team 0 then team 1 (Arena.cpp:1070-1085)
	V3 gridMin, gridMax;
	float cellSize = 0.f;
	int cellsX = 0, cellsY = 0, cellsZ = 0;
	std::vector<int32_t> cellStart;
	std::vector<uint8_t> cellStatics;
	// per cell-list entry (= position in cellStatics) of a MESH: the subtree headers of its BVH that can overlap a query
	// issued from that cell (box = the cell's 27-neighbourhood), in header creation order (AABB queries) and sorted by
	// root node (ray queries).  Pure culling: skipped headers cannot overlap the quantised query box.
	std::vector<int32_t> cellHdrStart;
	std::vector<uint16_t> cellHdrs, cellHdrsRay;
	// per cell-list entry of a mesh: leaf nodes reachable by a short ray cast from inside the cell (Scene::cellRayLeaves)
	std::vector<int32_t> cellRayLeafStart, cellRayLeaves;
	float rayLeafReach = 0.f;
};

namespace scene_detail {

inline void quantize(const MeshHost& m, const V3& p, int isMax, uint16_t out[3]) {
	V3 v = (p - m.bvhMin) * m.bvhQuant;
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
inline V3 unquantize(const MeshHost& m, const uint16_t q[3]) {
	return V3((float)q[0] / m.bvhQuant.x, (float)q[1] / m.bvhQuant.y, (float)q[2] / m.bvhQuant.z) + m.bvhMin;
}
inline V3 vdivv(const V3& a, const V3& b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }

// btQuantizedBvh::setQuantizationValues (three-pass refinement)
inline void set_quantization(MeshHost& m, const V3& aabbMin, const V3& aabbMax) {
	const V3 clampValue(1.f, 1.f, 1.f);
	const V3 k(65533.0f, 65533.0f, 65533.0f);
	m.bvhMin = aabbMin - clampValue;
	m.bvhMax = aabbMax + clampValue;
	V3 aabbSize = m.bvhMax - m.bvhMin;
	m.bvhQuant = vdivv(k, aabbSize);
	uint16_t q[3];
	quantize(m, m.bvhMin, 0, q);
	V3 v = unquantize(m, q);
	m.bvhMin = vmin(m.bvhMin, v - clampValue);
	aabbSize = m.bvhMax - m.bvhMin;
	m.bvhQuant = vdivv(k, aabbSize);
	quantize(m, m.bvhMax, 1, q);
	v = unquantize(m, q);
	m.bvhMax = vmax(m.bvhMax, v + clampValue);
	aabbSize = m.bvhMax - m.bvhMin;
	m.bvhQuant = vdivv(k, aabbSize);
}

inline void tri_verts(const MeshHost& m, int t, V3 out[3]) {
	for (int j = 0; j < 3; j++) {
		const float* p = &m.verts[4 * m.tris[3 * t + j]];
		out[j] = V3(p[0], p[1], p[2]);
	}
}

struct Builder {
	MeshHost& m;
	std::vector<BvhNode> leaves;
	int cur = 0;
	explicit Builder(MeshHost& mm) : m(mm) {}

	V3 leaf_min(int i) const { return unquantize(m, leaves[i].qmin); }
	V3 leaf_max(int i) const { return unquantize(m, leaves[i].qmax); }

	int calc_splitting_axis(int start, int end) {
		V3 means(0, 0, 0), variance(0, 0, 0);
		int n = end - start;
		for (int i = start; i < end; i++) means += 0.5f * (leaf_max(i) + leaf_min(i));
		means *= (1.f / (float)n);
		for (int i = start; i < end; i++) {
			V3 center = 0.5f * (leaf_max(i) + leaf_min(i));
			V3 d = center - means;
			d = d * d;
			variance += d;
		}
		variance *= (1.f / ((float)n - 1));
		// btVector3::maxAxis
		return variance.x < variance.y ? (variance.y < variance.z ? 2 : 1) : (variance.x < variance.z ? 2 : 0);
	}
	int sort_and_calc_splitting_index(int start, int end, int axis) {
		int splitIndex = start;
		int n = end - start;
		V3 means(0, 0, 0);
		for (int i = start; i < end; i++) means += 0.5f * (leaf_max(i) + leaf_min(i));
		means *= (1.f / (float)n);
		float splitValue = means[axis];
		for (int i = start; i < end; i++) {
			V3 center = 0.5f * (leaf_max(i) + leaf_min(i));
			if (center[axis] > splitValue) {
				BvhNode t = leaves[i]; leaves[i] = leaves[splitIndex]; leaves[splitIndex] = t;
				splitIndex++;
			}
		}
		int rangeBalanced = n / 3;
		bool unbalanced = ((splitIndex <= (start + rangeBalanced)) || (splitIndex >= (end - 1 - rangeBalanced)));
		if (unbalanced) splitIndex = start + (n >> 1);
		return splitIndex;
	}
	void update_subtree_headers(int left, int right) {
		const BvhNode& l = m.nodes[left];
		int lsz = l.idx >= 0 ? 1 : -l.idx;
		const BvhNode& r = m.nodes[right];
		int rsz = r.idx >= 0 ? 1 : -r.idx;
		if (lsz * 16 <= 2048) { SubtreeHdr h; memcpy(h.qmin, l.qmin, 6); memcpy(h.qmax, l.qmax, 6); h.root = left; h.size = lsz; m.hdrs.push_back(h); }
		if (rsz * 16 <= 2048) { SubtreeHdr h; memcpy(h.qmin, r.qmin, 6); memcpy(h.qmax, r.qmax, 6); h.root = right; h.size = rsz; m.hdrs.push_back(h); }
	}
	void build_tree(int start, int end) {
		int n = end - start;
		int curIndex = cur;
		if (n == 1) { m.nodes[cur] = leaves[start]; cur++; return; }
		int axis = calc_splitting_axis(start, end);
		int split = sort_and_calc_splitting_index(start, end, axis);
		int internal = cur;
		quantize(m, m.bvhMax, 0, m.nodes[cur].qmin);
		quantize(m, m.bvhMin, 1, m.nodes[cur].qmax);
		for (int i = start; i < end; i++) {
			uint16_t qmn[3], qmx[3];
			quantize(m, leaf_min(i), 0, qmn);
			quantize(m, leaf_max(i), 1, qmx);
			for (int k = 0; k < 3; k++) {
				if (m.nodes[cur].qmin[k] > qmn[k]) m.nodes[cur].qmin[k] = qmn[k];
				if (m.nodes[cur].qmax[k] < qmx[k]) m.nodes[cur].qmax[k] = qmx[k];
			}
		}
		cur++;
		int left = cur;
		build_tree(start, split);
		int right = cur;
		build_tree(split, end);
		int escape = cur - curIndex;
		if (escape * 16 > 2048) update_subtree_headers(left, right);
		m.nodes[internal].idx = -escape;
	}
};

// AABB query in report order (subtree headers, stackless walk): calls f(triIndex)
template <class F>
inline void bvh_query(const MeshHost& m, const V3& aabbMin, const V3& aabbMax, F&& f) {
	uint16_t qmin[3], qmax[3];
	V3 cmin = vmin(vmax(aabbMin, m.bvhMin), m.bvhMax), cmax = vmin(vmax(aabbMax, m.bvhMin), m.bvhMax);
	quantize(m, cmin, 0, qmin);
	quantize(m, cmax, 1, qmax);
	for (const SubtreeHdr& hd : m.hdrs) {
		if (!quant_overlap(qmin, qmax, hd.qmin, hd.qmax)) continue;
		int cur = hd.root, end = hd.root + hd.size;
		while (cur < end) {
			const BvhNode& nd = m.nodes[cur];
			bool ov = quant_overlap(qmin, qmax, nd.qmin, nd.qmax);
			bool leaf = nd.idx >= 0;
			if (leaf && ov) f(nd.idx & 0x1fffff);
			if (ov || leaf) cur++; else cur += -nd.idx;
		}
	}
}

inline V3 calc_normal(const V3 t[3]) { return normalized(cross(t[1] - t[0], t[2] - t[0])); }
inline V3 quat_rotate_h(const Q4& q, const V3& v) {
	Q4 t = q4(q.w * v.x + q.y * v.z - q.z * v.y, q.w * v.y + q.z * v.x - q.x * v.z, q.w * v.z + q.x * v.y - q.y * v.x, -q.x * v.x - q.y * v.y - q.z * v.z);
	Q4 qi = q4(-q.x, -q.y, -q.z, q.w);
	Q4 r = qmul(t, qi);
	return V3(r.x, r.y, r.z);
}

// btConnectivityProcessor::processTriangle for (A = triA, B = triB)
inline void connect(MeshHost& m, int triA, const V3 A[3], int triB) {
	if (triA == triB) return;
	V3 B[3];
	tri_verts(m, triB, B);
	const float equalVertexThreshold = 0.0001f * 0.0001f;
	const float planarEpsilon = 0.0001f;
	int numshared = 0;
	int sharedA[3] = {-1, -1, -1}, sharedB[3] = {-1, -1, -1};
	float crossBSqr = len2(cross(B[1] - B[0], B[2] - B[0]));
	if (crossBSqr < equalVertexThreshold) return;
	float crossASqr = len2(cross(A[1] - A[0], A[2] - A[0]));
	if (crossASqr < equalVertexThreshold) return;
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) {
			if (len2(A[i] - B[j]) < equalVertexThreshold) {
				sharedA[numshared] = i; sharedB[numshared] = j; numshared++;
				if (numshared >= 3) return;
			}
		}
		if (numshared >= 3) return;
	}
	if (numshared != 2) return;
	if (sharedA[0] == 0 && sharedA[1] == 2) {
		sharedA[0] = 2; sharedA[1] = 0;
		int tmp = sharedB[1]; sharedB[1] = sharedB[0]; sharedB[0] = tmp;
	}
	TriInfo& info = m.info[triA];
	if (info.flags < 0) { info.flags = 0; info.a01 = info.a12 = info.a20 = k2Pi; }
	int sumvertsA = sharedA[0] + sharedA[1];
	int otherIndexA = 3 - sumvertsA;
	V3 edge = A[sharedA[1]] - A[sharedA[0]];
	int otherIndexB = 3 - (sharedB[0] + sharedB[1]);
	V3 tB[3] = {B[sharedB[1]], B[sharedB[0]], B[otherIndexB]};
	V3 normalA = calc_normal(A);
	V3 normalB = calc_normal(tB);
	edge = normalized(edge);
	V3 edgeCrossA = normalized(cross(edge, normalA));
	{
		V3 tmp = A[otherIndexA] - A[sharedA[0]];
		if (dot(edgeCrossA, tmp) < 0.f) edgeCrossA *= -1.f;
	}
	V3 edgeCrossB = normalized(cross(edge, normalB));
	{
		V3 tmp = B[otherIndexB] - B[sharedB[0]];
		if (dot(edgeCrossB, tmp) < 0.f) edgeCrossB *= -1.f;
	}
	float angle2 = 0.f, ang4 = 0.f;
	V3 calculatedEdge = cross(edgeCrossA, edgeCrossB);
	float l2 = len2(calculatedEdge);
	float correctedAngle = 0.f;
	bool isConvex = false;
	if (l2 < planarEpsilon) {
		angle2 = 0.f; ang4 = 0.f;
	} else {
		calculatedEdge = normalized(calculatedEdge);
		V3 calculatedNormalA = normalized(cross(calculatedEdge, edgeCrossA));
		angle2 = get_angle(calculatedNormalA, edgeCrossA, edgeCrossB);
		ang4 = kPi - angle2;
		float dotA = dot(normalA, edgeCrossB);
		isConvex = (dotA < 0.f);
		correctedAngle = isConvex ? ang4 : -ang4;
	}
	(void)angle2;
	switch (sumvertsA) {
	case 1: {
		V3 e = A[0] - A[1];
		V3 computedNormalB = quat_rotate_h(q_axis_angle(e, -correctedAngle), normalA);
		if (dot(computedNormalB, normalB) < 0.f) info.flags |= 8;
		info.a01 = -correctedAngle;
		if (isConvex) info.flags |= 1;
		break;
	}
	case 2: {
		V3 e = A[2] - A[0];
		V3 computedNormalB = quat_rotate_h(q_axis_angle(e, -correctedAngle), normalA);
		if (dot(computedNormalB, normalB) < 0.f) info.flags |= 32;
		info.a20 = -correctedAngle;
		if (isConvex) info.flags |= 4;
		break;
	}
	case 3: {
		V3 e = A[1] - A[2];
		V3 computedNormalB = quat_rotate_h(q_axis_angle(e, -correctedAngle), normalA);
		if (dot(computedNormalB, normalB) < 0.f) info.flags |= 16;
		info.a12 = -correctedAngle;
		if (isConvex) info.flags |= 2;
		break;
	}
	}
}

}  // namespace scene_detail

inline MeshHost build_mesh(const uint8_t* data, size_t size) {
	using namespace scene_detail;
	if (size < 8) throw std::runtime_error("collision mesh file too small");
	int32_t nt, nv;
	memcpy(&nt, data, 4); memcpy(&nv, data + 4, 4);
	const int MAXC = 1000 * 1000;
	if (nt <= 0 || nv <= 0 || nt > MAXC || nv > MAXC)
		throw std::runtime_error("Invalid collision mesh file (bad triangle/vertex count: [" + std::to_string(nt) + ", " + std::to_string(nv) + "])");
	if (size < 8 + (size_t)nt * 12 + (size_t)nv * 12) throw std::runtime_error("Invalid collision mesh file (input data overflown)");
	MeshHost m;
	m.tris.resize((size_t)nt * 3);
	memcpy(m.tris.data(), data + 8, (size_t)nt * 12);
	for (int32_t ix : m.tris) if (ix < 0 || ix >= nv) throw std::runtime_error("Invalid collision mesh file (bad triangle vertex index)");
	m.verts.assign((size_t)nv * 4, 0.f);
	const uint8_t* vp = data + 8 + (size_t)nt * 12;
	for (int i = 0; i < nv; i++) memcpy(&m.verts[4 * i], vp + 12 * (size_t)i, 12);

	// local AABB = extremes over triangle vertices (mesh margin 0)
	V3 mn(kLargeFloat, kLargeFloat, kLargeFloat), mx(-kLargeFloat, -kLargeFloat, -kLargeFloat);
	// btTriangleMeshShape::recalcLocalAabb: per axis support vertex with maxDot starting at -BT_LARGE_FLOAT
	for (int t = 0; t < nt; t++) { V3 v[3]; tri_verts(m, t, v); for (int j = 0; j < 3; j++) { mn = vmin(mn, v[j]); mx = vmax(mx, v[j]); } }
	m.aabbMin = mn; m.aabbMax = mx;

	set_quantization(m, mn, mx);
	Builder b(m);
	b.leaves.resize(nt);
	for (int t = 0; t < nt; t++) {
		V3 v[3]; tri_verts(m, t, v);
		V3 amin(kLargeFloat, kLargeFloat, kLargeFloat), amax(-kLargeFloat, -kLargeFloat, -kLargeFloat);
		for (int j = 0; j < 3; j++) { amin = vmin(amin, v[j]); amax = vmax(amax, v[j]); }
		const float MIN_DIM = 0.002f, MIN_HALF = 0.001f;
		if (amax.x - amin.x < MIN_DIM) { amax.x = amax.x + MIN_HALF; amin.x = amin.x - MIN_HALF; }
		if (amax.y - amin.y < MIN_DIM) { amax.y = amax.y + MIN_HALF; amin.y = amin.y - MIN_HALF; }
		if (amax.z - amin.z < MIN_DIM) { amax.z = amax.z + MIN_HALF; amin.z = amin.z - MIN_HALF; }
		quantize(m, amin, 0, b.leaves[t].qmin);
		quantize(m, amax, 1, b.leaves[t].qmax);
		b.leaves[t].idx = t;  // partId 0
	}
	m.nodes.resize((size_t)2 * nt);
	memset(m.nodes.data(), 0, m.nodes.size() * sizeof(BvhNode));
	b.build_tree(0, nt);
	m.nodes.resize(b.cur);
	if (m.hdrs.empty()) {
		SubtreeHdr h; memcpy(h.qmin, m.nodes[0].qmin, 6); memcpy(h.qmax, m.nodes[0].qmax, 6);
		h.root = 0; h.size = m.nodes[0].idx >= 0 ? 1 : -m.nodes[0].idx;
		m.hdrs.push_back(h);
	}
	// btGenerateInternalEdgeInfo
	m.info.assign(nt, TriInfo{-1, k2Pi, k2Pi, k2Pi});
	for (int t = 0; t < nt; t++) {
		V3 A[3]; tri_verts(m, t, A);
		V3 amin(kLargeFloat, kLargeFloat, kLargeFloat), amax(-kLargeFloat, -kLargeFloat, -kLargeFloat);
		for (int j = 0; j < 3; j++) { amin = vmin(amin, A[j]); amax = vmax(amax, A[j]); }
		bvh_query(m, amin, amax, [&](int tb) { connect(m, t, A, tb); });
	}
	// unquantized node boxes for the ray walk (btQuantizedBvh::unQuantize: (float)q / quantization + bvhAabbMin, IEEE division)
	m.nodeBounds.resize(m.nodes.size() * 8);
	for (size_t n = 0; n < m.nodes.size(); n++) {
		const V3 lo = scene_detail::unquantize(m, m.nodes[n].qmin), hi = scene_detail::unquantize(m, m.nodes[n].qmax);
		float* o = &m.nodeBounds[n * 8];
		o[0] = lo.x; o[1] = lo.y; o[2] = lo.z; o[3] = 0.f; o[4] = hi.x; o[5] = hi.y; o[6] = hi.z; o[7] = 0.f;
	}
	return m;
}

// btStaticPlaneShape stores planeNormal.normalized() (btStaticPlaneShape.cpp:19-25): with the reference's rsqrtss-based
// btVector3::normalize() the arena planes carry e.g. (0, 0, 0.99999994) rather than (0, 0, 1); that value reaches every plane
// contact normal and friction direction.  normalized() (rsb_math.h) reproduces that instruction sequence on any host.
inline V3 bullet_normalized_host(const V3& v) { return normalized(v); }

inline PlaneD make_plane(const V3& posUU, const V3& nIn, int filter = 0) {
	PlaneD p;
	// _AddStaticCollisionShape (Arena.cpp:976-987): group = mask = `filter` when given, else the static defaults of addRigidBody
	p.group = filter ? filter : FILT_STATIC;
	p.mask = filter ? filter : (FILT_ALL ^ FILT_STATIC);
	const V3 n = bullet_normalized_host(nIn);
	p.n = n; p.pos = posUU * UU_TO_BT;
	p.d = dot(n, p.pos);
	// btStaticPlaneShape::getAabb (single-axis slab, half-infinite behind the plane)
	p.aabbMin = V3(-kLargeFloat, -kLargeFloat, -kLargeFloat);
	p.aabbMax = V3(kLargeFloat, kLargeFloat, kLargeFloat);
	const float OFF = 0.2f;
	int axis = (n.x != 0.f) ? 0 : ((n.y != 0.f) ? 1 : 2);
	bool backwards = n[axis] < 0.f;
	float o = p.pos[axis];
	p.aabbMin.set(axis, o + (0.f - OFF));
	p.aabbMax.set(axis, o + (0.f + OFF));
	if (backwards) p.aabbMax.set(axis, kLargeFloat); else p.aabbMin.set(axis, -kLargeFloat);
	return p;
}

// DropshotTiles::Init + MakeTileShapes (DropshotTiles.cpp:22-112) and the tile bodies of Arena::_SetupArenaCollisionShapes
// (Arena.cpp:1070-1085), in the reference's float arithmetic.
inline std::vector<TileD> make_dropshot_tiles() {
	const float BT_TO_UU_ = 50.f, UU_TO_BT_ = (1.f / 50.f);
	const V3 HEX[6] = {V3(0.0f, -8.85f, 0.f), V3(7.6643f, -4.425f, 0.f), V3(7.6643f, 4.425f, 0.f), V3(0.0f, 8.85f, 0.f), V3(-7.6643f, 4.425f, 0.f), V3(-7.6643f, -4.425f, 0.f)};
	const float TILE_WIDTH_X = HEX[1].x * 2 * BT_TO_UU_;
	const float ROW_OFFSET_Y = (HEX[3].y + HEX[4].y) * BT_TO_UU_;
	const float TILE_OFFSET_Y = 2.54736f * BT_TO_UU_;
	const int FIRST_ROW = 13, LAST_ROW = 7, ROWS = FIRST_ROW - LAST_ROW + 1;
	V3 pos[TILES_PER_TEAM];
	int cur = 0;
	float y = TILE_OFFSET_Y;
	for (int i = 0, numTiles = FIRST_ROW; i < ROWS; i++, numTiles--) {
		float rowSizeX = TILE_WIDTH_X * numTiles;
		float rowStartX = -(rowSizeX / 2.f) + (TILE_WIDTH_X / 2);
		for (int j = 0; j < numTiles; j++, cur++) {
			if (cur >= TILES_PER_TEAM) throw std::runtime_error("dropshot tile table overflow");
			float x = rowStartX + TILE_WIDTH_X * j;
			pos[cur] = V3(x, y, 0.f);
		}
		y += ROW_OFFSET_Y;
	}
	if (cur != TILES_PER_TEAM) throw std::runtime_error("dropshot tile table incomplete");
	std::vector<TileD> tiles(MAX_TILES);
	const float margin = 0.04f;  // btConvexHullShape default (CONVEX_DISTANCE_MARGIN)
	const float NEIGHBOR_MAX_RADIUS = TILE_WIDTH_X * 1.2f;
	for (int team = 0; team < 2; team++) {
		for (int i = 0; i < TILES_PER_TEAM; i++) {
			TileD& t = tiles[team * TILES_PER_TEAM + i];
			memset(&t, 0, sizeof(t));
			const V3 tp = pos[i] * ((team == 0) ? -1.f : 1.f);  // GetTilePos
			t.posY = tp.y;
			const float CLAMP_Y = TILE_OFFSET_Y * UU_TO_BT_;
			for (int j = 0; j < 6; j++) {
				V3 vert = tp * UU_TO_BT_ + HEX[j];
				if (team == 0) vert.y = vert.y < -CLAMP_Y ? vert.y : -CLAMP_Y;  // RS_MIN(vert.y, -CLAMP_Y)
				else vert.y = vert.y > CLAMP_Y ? vert.y : CLAMP_Y;              // RS_MAX(vert.y, CLAMP_Y)
				t.pts[j] = vert;
			}
			// recalcLocalAabb (btPolyhedralConvexShape.cpp:499-529): support points (first maximum of the dot product) along +-axes
			V3 lmin, lmax;
			for (int a = 0; a < 3; a++) {
				V3 dp(0, 0, 0), dn(0, 0, 0); dp.set(a, 1.f); dn.set(a, -1.f);
				float bp = -3.402823466e+38f, bn = -3.402823466e+38f; int ip = -1, in = -1;
				for (int j = 0; j < 6; j++) {
					float d1 = dot(t.pts[j], dp), d2 = dot(t.pts[j], dn);
					if (d1 > bp) { bp = d1; ip = j; }
					if (d2 > bn) { bn = d2; in = j; }
				}
				lmax.set(a, t.pts[ip][a] + margin);
				lmin.set(a, t.pts[in][a] - margin);
			}
			// getAabb(body transform) = btTransformAabb(localMin, localMax, margin, t) with t = (identity, (0, 0, -margin))
			const V3 origin(0.f, 0.f, -margin);
			V3 lh = 0.5f * (lmax - lmin);
			lh += V3(margin, margin, margin);
			const V3 lc = 0.5f * (lmax + lmin);
			const M3 I = m3identity();
			const V3 center = I * lc + origin;
			const V3 extent(dot(lh, I.r[0]), dot(lh, I.r[1]), dot(lh, I.r[2]));
			t.aabbMin = center - extent; t.aabbMax = center + extent;
			// neighbour lists (team-independent: computed from team 0 positions, DropshotTiles.cpp:92-112)
			const V3 p0 = pos[i] * -1.f;
			for (int j = 0; j < TILES_PER_TEAM; j++) {
				const V3 d = p0 - pos[j] * -1.f;
				const float dist = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
				if (dist < NEIGHBOR_MAX_RADIUS) { if (t.n1 >= 7) throw std::runtime_error("dropshot neighbour table"); t.nb1[t.n1++] = (uint8_t)j; }
				if (dist < NEIGHBOR_MAX_RADIUS * 2) { if (t.n2 >= 20) throw std::runtime_error("dropshot neighbour table"); t.nb2[t.n2++] = (uint8_t)j; }
			}
		}
	}
	return tiles;
}

inline void grid_indices(const SceneHost& s, const V3& pos, int& i, int& j, int& k) {
	V3 f = (pos - s.gridMin) / s.cellSize;
	i = (int)f.x; j = (int)f.y; k = (int)f.z;
	i = i < 0 ? 0 : (i > s.cellsX - 1 ? s.cellsX - 1 : i);
	j = j < 0 ? 0 : (j > s.cellsY - 1 ? s.cellsY - 1 : j);
	k = k < 0 ? 0 : (k > s.cellsZ - 1 ? s.cellsZ - 1 : k);
}

inline SceneHost build_scene(int mode, std::vector<MeshHost> meshes) {
	using namespace scene_detail;
	SceneHost s;
	s.mode = mode;
	s.meshes = std::move(meshes);
	const bool isHoops = mode == HOOPS;
	if ((int)s.meshes.size() > MAX_MESHES) throw std::runtime_error("too many collision meshes for one game mode");
	for (MeshHost& m : s.meshes) {
		bool net = isHoops && m.nVerts() == 505;  // Arena.cpp:1016-1019
		m.group = net ? FILT_HOOPS_NET : FILT_STATIC;
		m.mask = net ? FILT_HOOPS_NET : (FILT_ALL ^ FILT_STATIC);
	}
	const bool isDropshot = mode == DROPSHOT;
	if (mode != THE_VOID) {
		float extentX = isHoops ? K::ARENA_EXTENT_X_HOOPS : K::ARENA_EXTENT_X;
		float extentY = isHoops ? K::ARENA_EXTENT_Y_HOOPS : K::ARENA_EXTENT_Y;
		float height = isDropshot ? 2024.f : (isHoops ? K::ARENA_HEIGHT_HOOPS : K::ARENA_HEIGHT);  // ARENA_HEIGHT_DROPSHOT (RLConst.h:22)
		// Dropshot: the floor plane sits at FLOOR_HEIGHT_DROPSHOT = 1.5 uu and only collides with cars and wheel rays
		// (DROPSHOT_FLOOR filter); the ball lands on the tiles.  No side walls (Arena.cpp:1051-1061).
		s.planes.push_back(make_plane(V3(0, 0, isDropshot ? 1.5f : 0.f), V3(0, 0, 1), isDropshot ? FILT_DROPSHOT_FLOOR : 0));
		s.planes.push_back(make_plane(V3(0, 0, height), V3(0, 0, -1)));
		if (!isDropshot) {
			s.planes.push_back(make_plane(V3(-extentX, 0, height / 2), V3(1, 0, 0)));
			s.planes.push_back(make_plane(V3(extentX, 0, height / 2), V3(-1, 0, 0)));
		}
		if (isHoops) {
			s.planes.push_back(make_plane(V3(0, -extentY, height / 2), V3(0, 1, 0)));
			s.planes.push_back(make_plane(V3(0, extentY, height / 2), V3(0, -1, 0)));
		}
	}
	// btRSBroadphase grid from the default ArenaConfig (ArenaConfig.h:23-41, Arena.cpp:446-458)
	s.gridMin = V3(-5600, -6000, 0) * UU_TO_BT;
	s.gridMax = V3(5600, 6000, 2200) * UU_TO_BT;
	s.cellSize = 370 * UU_TO_BT * 1.f;
	V3 range = s.gridMax - s.gridMin;
	auto cl = [](float v) { int c = (int)ceilf(v); return c < 1 ? 1 : c; };
	s.cellsX = cl(range.x / s.cellSize); s.cellsY = cl(range.y / s.cellSize); s.cellsZ = cl(range.z / s.cellSize);
	const int total = s.cellsX * s.cellsY * s.cellsZ;
	std::vector<std::vector<uint8_t>> cells(total);
	if (isDropshot) s.tiles = make_dropshot_tiles();
	const int nStatics = (int)s.meshes.size() + (int)s.planes.size() + (int)s.tiles.size();
	if (nStatics > 255) throw std::runtime_error("too many static objects");
	const V3 ct(0.02f, 0.02f, 0.02f);
	for (int o = 0; o < nStatics; o++) {
		const bool isMesh = o < (int)s.meshes.size();
		const int op = o - (int)s.meshes.size(), ot = op - (int)s.planes.size();
		V3 amin, amax;
		if (isMesh) { amin = s.meshes[o].aabbMin - ct; amax = s.meshes[o].aabbMax + ct; }
		else if (ot < 0) { amin = s.planes[op].aabbMin - ct; amax = s.planes[op].aabbMax + ct; }
		else { amin = s.tiles[ot].aabbMin - ct; amax = s.tiles[ot].aabbMax + ct; }
		V3 cmax = vmin(amax, s.gridMax);
		int iMin, jMin, kMin, iMax, jMax, kMax;
		grid_indices(s, amin, iMin, jMin, kMin);
		grid_indices(s, cmax, iMax, jMax, kMax);
		for (int i = iMin; i <= iMax; i++) for (int j = jMin; j <= jMax; j++) for (int k = kMin; k <= kMax; k++) {
			if (isMesh) {
				const MeshHost& m = s.meshes[o];
				V3 cellMin = s.gridMin + V3((float)i, (float)j, (float)k) * s.cellSize;
				V3 cellMax = cellMin + V3(s.cellSize, s.cellSize, s.cellSize);
				bool hit = false;
				for (int t = 0; t < m.nTris() && !hit; t++) {
					V3 tv[3]; tri_verts(m, t, tv);
					bool ok = true;
					const int order[3] = {2, 0, 1};
					for (int q = 0; q < 3 && ok; q++) {
						int a = order[q];
						if (fminf_(fminf_(tv[0][a], tv[1][a]), tv[2][a]) > cellMax[a]) ok = false;
						else if (fmaxf_(fmaxf_(tv[0][a], tv[1][a]), tv[2][a]) < cellMin[a]) ok = false;
					}
					if (ok) hit = true;
				}
				if (!hit) continue;
			}
			for (int i1 = -1; i1 <= 1; i1++) for (int j1 = -1; j1 <= 1; j1++) for (int k1 = -1; k1 <= 1; k1++) {
				int ci = i + i1, cj = j + j1, ck = k + k1;
				if (ci < 0 || cj < 0 || ck < 0 || ci >= s.cellsX || cj >= s.cellsY || ck >= s.cellsZ) continue;
				auto& cell = cells[ci * s.cellsY * s.cellsZ + cj * s.cellsZ + ck];
				bool exists = false;
				for (uint8_t e : cell) if (e == (uint8_t)o) { exists = true; break; }
				if (!exists) cell.push_back((uint8_t)o);
			}
		}
	}
	s.cellStart.resize(total + 1);
	s.cellStart[0] = 0;
	for (int c = 0; c < total; c++) {
		s.cellStatics.insert(s.cellStatics.end(), cells[c].begin(), cells[c].end());
		s.cellStart[c + 1] = (int32_t)s.cellStatics.size();
	}
	// header candidate lists per (cell, mesh) entry
	s.cellHdrStart.assign(1, 0);
	const float INF = 1e30f;
	for (int c = 0; c < total; c++) {
		const int ci = c / (s.cellsY * s.cellsZ), cj = (c / s.cellsZ) % s.cellsY, ck = c % s.cellsZ;
		V3 cellMin = s.gridMin + V3((float)ci, (float)cj, (float)ck) * s.cellSize;
		V3 nMin = cellMin - V3(s.cellSize, s.cellSize, s.cellSize) - ct;
		V3 nMax = cellMin + V3(s.cellSize, s.cellSize, s.cellSize) * 2.f + ct;
		// positions outside the grid are clamped into the border cells
		if (ci == 0) nMin.x = -INF; if (cj == 0) nMin.y = -INF; if (ck == 0) nMin.z = -INF;
		if (ci == s.cellsX - 1) nMax.x = INF; if (cj == s.cellsY - 1) nMax.y = INF; if (ck == s.cellsZ - 1) nMax.z = INF;
		for (int e = s.cellStart[c]; e < s.cellStart[c + 1]; e++) {
			const int o = s.cellStatics[e];
			if (o < (int)s.meshes.size()) {
				const MeshHost& m = s.meshes[o];
				if (m.hdrs.size() > 65535) throw std::runtime_error("mesh has too many BVH subtree headers");
				uint16_t qmin[3], qmax[3];
				V3 lo = vmin(vmax(nMin, m.bvhMin), m.bvhMax), hi = vmin(vmax(nMax, m.bvhMin), m.bvhMax);  // quantizeWithClamp
				quantize(m, lo, 0, qmin); quantize(m, hi, 1, qmax);
				std::vector<uint16_t> list;
				for (size_t h = 0; h < m.hdrs.size(); h++) {
					const SubtreeHdr& hd = m.hdrs[h];
					bool ov = (qmin[0] <= hd.qmax[0]) && (qmax[0] >= hd.qmin[0]) && (qmin[1] <= hd.qmax[1]) && (qmax[1] >= hd.qmin[1]) &&
					          (qmin[2] <= hd.qmax[2]) && (qmax[2] >= hd.qmin[2]);
					if (ov) list.push_back((uint16_t)h);
				}
				s.cellHdrs.insert(s.cellHdrs.end(), list.begin(), list.end());
				std::sort(list.begin(), list.end(), [&](uint16_t a, uint16_t b) { return m.hdrs[a].root < m.hdrs[b].root; });
				s.cellHdrsRay.insert(s.cellHdrsRay.end(), list.begin(), list.end());
			}
			s.cellHdrStart.push_back((int32_t)s.cellHdrs.size());
		}
	}
	// leaf candidates of short rays (suspension rays are <= 1.28 bt long; 1.4 bt of reach + 0.1 bt of slack for the conservative
	// quantised boxes).  Border cells also serve every position outside the grid (grid_cell_of clamps).
	s.rayLeafReach = 1.4f;  // every car preset casts <= 1.28 bt (rest 39.5 + travel 12 + radius 15 - 2.5 uu); longer rays of custom configs take the hierarchical walk
	s.cellRayLeafStart.assign(1, 0);
	for (int c = 0; c < total; c++) {
		const int ci = c / (s.cellsY * s.cellsZ), cj = (c / s.cellsZ) % s.cellsY, ck = c % s.cellsZ;
		const float grow = s.rayLeafReach + 0.1f;
		V3 lo = s.gridMin + V3((float)ci, (float)cj, (float)ck) * s.cellSize - V3(grow, grow, grow);
		V3 hi = s.gridMin + V3((float)(ci + 1), (float)(cj + 1), (float)(ck + 1)) * s.cellSize + V3(grow, grow, grow);
		if (ci == 0) lo.x = -INF; if (cj == 0) lo.y = -INF; if (ck == 0) lo.z = -INF;
		if (ci == s.cellsX - 1) hi.x = INF; if (cj == s.cellsY - 1) hi.y = INF; if (ck == s.cellsZ - 1) hi.z = INF;
		for (int e = s.cellStart[c]; e < s.cellStart[c + 1]; e++) {
			const int o = s.cellStatics[e];
			if (o < (int)s.meshes.size()) {
				const MeshHost& m = s.meshes[o];
				for (size_t n = 0; n < m.nodes.size(); n++) {
					if (m.nodes[n].idx < 0) continue;  // internal node
					const float* b = &m.nodeBounds[n * 8];
					if (b[0] > hi.x || b[4] < lo.x || b[1] > hi.y || b[5] < lo.y || b[2] > hi.z || b[6] < lo.z) continue;
					s.cellRayLeaves.push_back((int32_t)n);
				}
			}
			s.cellRayLeafStart.push_back((int32_t)s.cellRayLeaves.size());
		}
	}
	return s;
}

}  // namespace rsb
