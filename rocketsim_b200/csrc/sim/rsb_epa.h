// rsb_epa.h -- penetration depth of two overlapping convex cores: the reference's second GJK
// (simplex projection by sub-determinants) followed by the expanding-polytope algorithm, plus
// the 9-direction retry loop and the distance fallback.  Restates
//   btGjkEpaPenetrationDepthSolver::calcPenDepth   bt/BulletCollision/NarrowPhaseCollision/btGjkEpaPenetrationDepthSolver.cpp:23-79
//   gjkepa2_impl::GJK / EPA / MinkowskiDiff        bt/BulletCollision/NarrowPhaseCollision/btGjkEpa2.cpp:40-1017
// Differences of form, not arithmetic: simplex vertices are values instead of pooled pointers,
// hull faces are indexed records in the arena's scratch (18 KB of per-thread stack would be
// reserved for every resident thread otherwise), and the two recursive routines
// (EncloseOrigin, expand) are unrolled onto explicit stacks so the kernel needs no
// recursion stack.
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"

namespace rsb {

struct Mink {
	Convex s0, s1;
	M3 toshape1;
	Xform toshape0;
	bool margins;
};

RSB_HD M3 transpose_times(const M3& a, const M3& b) {  // a^T * b
	M3 o;
	for (int i = 0; i < 3; i++)
		o.r[i] = V3((a.r[0][i] * b.r[0].x + a.r[1][i] * b.r[1].x) + a.r[2][i] * b.r[2].x,
		            (a.r[0][i] * b.r[0].y + a.r[1][i] * b.r[1].y) + a.r[2][i] * b.r[2].y,
		            (a.r[0][i] * b.r[0].z + a.r[1][i] * b.r[1].z) + a.r[2][i] * b.r[2].z);
	return o;
}
// localGetSupportVertexNonVirtual (btConvexShape.cpp:245-...): support of the core along the
// NORMALISED direction plus margin along it
RSB_HD V3 support_nv(const Convex& c, const V3& dir, bool margins) {
	if (!margins) return support_no_margin(c, dir);
	V3 n = dir;
	if (len2(n) < kEps * kEps) n = V3(-1.f, -1.f, -1.f);
	n = normalized(n);
	return support_no_margin(c, n) + c.margin * n;
}
RSB_HD V3 mink_support0(const Mink& m, const V3& d) { return support_nv(m.s0, d, m.margins); }
RSB_HD V3 mink_support1(const Mink& m, const V3& d) { return xf_apply(m.toshape0, support_nv(m.s1, m.toshape1 * d, m.margins)); }
RSB_HD V3 mink_support(const Mink& m, const V3& d) { return mink_support0(m, d) - mink_support1(m, -d); }
RSB_HD void mink_init(Mink& m, const Convex& A, const Xform& tA, const Convex& B, const Xform& tB, bool margins) {
	m.s0 = A; m.s1 = B;
	m.toshape1 = transpose_times(tB.basis, tA.basis);
	m.toshape0.basis = transpose_times(tA.basis, tB.basis);
	m.toshape0.origin = vmul(tB.origin - tA.origin, tA.basis);
	m.margins = margins;
}

// ---- closest point of a simplex to the origin, with the barycentric weights of the supporting sub-simplex -------------------
// Same decisions and arithmetic as the reference's sub-determinant projection (gjkepa2_impl::GJK::projectorigin,
// btGjkEpa2.cpp:430-560), arranged as value-returning routines: an OriginFit carries the squared distance (< 0: degenerate
// simplex), the weights and the bit mask of the vertices that support the closest point.  A k-simplex first asks its
// (k-1)-faces that can see the origin; only when none does is the origin's projection interior to the simplex itself.
struct OriginFit { float d2; float w[4]; unsigned mask; };

RSB_HD float triple3(const V3& a, const V3& b, const V3& c) {  // signed volume, the reference's term order
	return (a.y * b.z * c.x + a.z * b.x * c.y - a.x * b.z * c.y - a.y * b.x * c.z + a.x * b.y * c.z - a.z * b.y * c.x);
}
RSB_HD OriginFit fit_segment(const V3& a, const V3& b) {
	OriginFit f; f.d2 = -1.f; f.w[0] = f.w[1] = f.w[2] = f.w[3] = 0.f; f.mask = 0;
	const V3 ab = b - a;
	const float l2 = len2(ab);
	if (!(l2 > 0.f)) return f;
	const float t = -dot(a, ab) / l2;
	if (t >= 1.f) { f.w[1] = 1.f; f.mask = 2; f.d2 = len2(b); }
	else if (t <= 0.f) { f.w[0] = 1.f; f.mask = 1; f.d2 = len2(a); }
	else { f.w[1] = t; f.w[0] = 1 - t; f.mask = 3; f.d2 = len2(a + ab * t); }
	return f;
}
RSB_HD OriginFit fit_triangle(const V3& a, const V3& b, const V3& c) {
	OriginFit f; f.d2 = -1.f; f.w[0] = f.w[1] = f.w[2] = f.w[3] = 0.f; f.mask = 0;
	const V3 vt[3] = {a, b, c};
	const V3 edge[3] = {a - b, b - c, c - a};
	const V3 n = cross(edge[0], edge[1]);
	const float l2 = len2(n);
	if (!(l2 > 0.f)) return f;
	for (unsigned i = 0; i < 3; ++i) {
		if (!(dot(vt[i], cross(edge[i], n)) > 0.f)) continue;   // the origin is not outside edge (i, i+1)
		const unsigned j = (i + 1) % 3, k = (j + 1) % 3;
		const OriginFit e = fit_segment(vt[i], vt[j]);
		if (f.d2 < 0.f || e.d2 < f.d2) {
			f.d2 = e.d2;
			f.mask = ((e.mask & 1) ? 1u << i : 0u) + ((e.mask & 2) ? 1u << j : 0u);
			f.w[i] = e.w[0]; f.w[j] = e.w[1]; f.w[k] = 0.f;
		}
	}
	if (f.d2 < 0.f) {  // interior: foot of the perpendicular, weights from the sub-triangle areas
		const float d = dot(a, n);
		const float s = sqrtf(l2);
		const V3 p = n * (d / l2);
		f.d2 = len2(p);
		f.mask = 7;
		f.w[0] = len(cross(edge[1], b - p)) / s;
		f.w[1] = len(cross(edge[2], c - p)) / s;
		f.w[2] = 1 - (f.w[0] + f.w[1]);
	}
	return f;
}
RSB_HD OriginFit fit_tetra(const V3& a, const V3& b, const V3& c, const V3& d) {
	OriginFit f; f.d2 = -1.f; f.w[0] = f.w[1] = f.w[2] = f.w[3] = 0.f; f.mask = 0;
	const V3 vt[3] = {a, b, c};
	const V3 apex[3] = {a - d, b - d, c - d};
	const float vol = triple3(apex[0], apex[1], apex[2]);
	const bool wellOriented = (vol * dot(a, cross(b - c, a - b))) <= 0.f;
	if (!(wellOriented && fabsf(vol) > 0.f)) return f;
	for (unsigned i = 0; i < 3; ++i) {
		const unsigned j = (i + 1) % 3, k = (j + 1) % 3;
		if (!(vol * dot(d, cross(apex[i], apex[j])) > 0.f)) continue;   // the origin is not outside face (i, j, d)
		const OriginFit t = fit_triangle(vt[i], vt[j], d);
		if (f.d2 < 0.f || t.d2 < f.d2) {
			f.d2 = t.d2;
			f.mask = ((t.mask & 1) ? 1u << i : 0u) + ((t.mask & 2) ? 1u << j : 0u) + ((t.mask & 4) ? 8u : 0u);
			f.w[i] = t.w[0]; f.w[j] = t.w[1]; f.w[k] = 0.f; f.w[3] = t.w[2];
		}
	}
	if (f.d2 < 0.f) {  // the origin is inside the tetrahedron
		f.d2 = 0.f; f.mask = 15;
		f.w[0] = triple3(c, b, d) / vol;
		f.w[1] = triple3(a, c, d) / vol;
		f.w[2] = triple3(b, a, d) / vol;
		f.w[3] = 1 - (f.w[0] + f.w[1] + f.w[2]);
	}
	return f;
}

enum Gjk2Status { G2_VALID = 0, G2_INSIDE = 1, G2_FAILED = 2 };

struct Gjk2 {
	Mink shape;
	V3 ray;
	float distance;
	Simplex2 simplex;  // result simplex
	int status;
};

RSB_HD void gjk2_getsupport(const Mink& sh, const V3& d, SV& sv) { sv.d = d / len(d); sv.w = mink_support(sh, sv.d); }
RSB_HD void gjk2_append(const Mink& sh, Simplex2& s, const V3& v) { s.p[s.rank] = 0.f; gjk2_getsupport(sh, v, s.c[s.rank]); s.rank++; }

// GJK::Evaluate (btGjkEpa2.cpp:198-340)
RSB_HD_NOINLINE int gjk2_evaluate(Gjk2& g, V3 guess) {
	unsigned iterations = 0;
	float sqdist = 0.f, alpha = 0.f;
	V3 lastw[4];
	unsigned clastw = 0;
	Simplex2 simp[2];
	unsigned current = 0;
	g.status = G2_VALID;
	g.distance = 0.f;
	simp[0].rank = 0;
	g.ray = guess;
	const float sqrl = len2(g.ray);
	gjk2_append(g.shape, simp[0], sqrl > 0.f ? -g.ray : V3(1.f, 0.f, 0.f));
	simp[0].p[0] = 1.f;
	g.ray = simp[0].c[0].w;
	sqdist = sqrl;
	lastw[0] = lastw[1] = lastw[2] = lastw[3] = g.ray;
	do {
		const unsigned next = 1 - current;
		Simplex2& cs = simp[current];
		Simplex2& ns = simp[next];
		const float rl = len(g.ray);
		if (rl < GJK2_MIN_DISTANCE) { g.status = G2_INSIDE; break; }
		gjk2_append(g.shape, cs, -g.ray);
		const V3 w = cs.c[cs.rank - 1].w;
		bool found = false;
		for (unsigned i = 0; i < 4; ++i) if (len2(w - lastw[i]) < GJK2_DUPLICATED_EPS) { found = true; break; }
		if (found) { cs.rank--; break; }
		else { lastw[clastw = (clastw + 1) & 3] = w; }
		const float omega = dot(g.ray, w) / rl;
		alpha = omega > alpha ? omega : alpha;
		if (((rl - alpha) - (GJK2_ACCURACY * rl)) <= 0.f) { cs.rank--; break; }
		OriginFit fit; fit.d2 = sqdist; fit.mask = 0; fit.w[0] = fit.w[1] = fit.w[2] = fit.w[3] = 0.f;
		switch (cs.rank) {
		case 2: fit = fit_segment(cs.c[0].w, cs.c[1].w); break;
		case 3: fit = fit_triangle(cs.c[0].w, cs.c[1].w, cs.c[2].w); break;
		case 4: fit = fit_tetra(cs.c[0].w, cs.c[1].w, cs.c[2].w, cs.c[3].w); break;
		}
		sqdist = fit.d2;
		const float* weights = fit.w;
		const unsigned mask = fit.mask;
		if (sqdist >= 0.f) {
			ns.rank = 0;
			g.ray = v3zero();
			current = next;
			for (int i = 0, ni = cs.rank; i < ni; ++i) {
				if (mask & (1u << i)) {
					ns.c[ns.rank] = cs.c[i];
					ns.p[ns.rank++] = weights[i];
					g.ray += cs.c[i].w * weights[i];
				}
			}
			if (mask == 15) g.status = G2_INSIDE;
		} else { cs.rank--; break; }
		g.status = ((++iterations) < 128) ? g.status : G2_FAILED;
	} while (g.status == G2_VALID);
	g.simplex = simp[current];
	switch (g.status) {
	case G2_VALID: g.distance = len(g.ray); break;
	case G2_INSIDE: g.distance = 0.f; break;
	default: break;
	}
	return g.status;
}

// GJK::EncloseOrigin (btGjkEpa2.cpp:341-418), depth-first search unrolled onto option counters
RSB_HD_NOINLINE bool gjk2_enclose_origin(Gjk2& g) {
	Simplex2& s = g.simplex;
	const int base = s.rank;
	int opt[5] = {0, 0, 0, 0, 0};
	for (int guard = 0; guard < 4096; guard++) {
		const int r = s.rank;
		if (r == 4) {
			if (fabsf(triple3(s.c[0].w - s.c[3].w, s.c[1].w - s.c[3].w, s.c[2].w - s.c[3].w)) > 0.f) return true;
			if (r == base) return false;
			s.rank--;
			continue;
		}
		if (r < 1) return false;
		const int nOpts = (r == 3) ? 2 : 6;
		bool advanced = false;
		while (opt[r] < nOpts) {
			const int k = opt[r]++;
			V3 dir;
			if (r == 1) {
				V3 axis(0.f, 0.f, 0.f); axis.set(k >> 1, 1.f);
				dir = (k & 1) ? -axis : axis;
			} else if (r == 2) {
				const V3 d = s.c[1].w - s.c[0].w;
				V3 axis(0.f, 0.f, 0.f); axis.set(k >> 1, 1.f);
				const V3 p = cross(d, axis);
				if (!(len2(p) > 0.f)) continue;
				dir = (k & 1) ? -p : p;
			} else {
				const V3 n = cross(s.c[1].w - s.c[0].w, s.c[2].w - s.c[0].w);
				if (!(len2(n) > 0.f)) continue;
				dir = (k & 1) ? -n : n;
			}
			gjk2_append(g.shape, s, dir);
			opt[r + 1] = 0;
			advanced = true;
			break;
		}
		if (!advanced) {
			if (r == base) return false;
			s.rank--;
		}
	}
	return false;
}

enum EpaStatus { EPA_VALID = 0, EPA_TOUCHING, EPA_DEGENERATED, EPA_NONCONVEX, EPA_INVALIDHULL, EPA_OUTOFFACES, EPA_OUTOFVERTICES, EPA_ACCURACYREACHED, EPA_FALLBACK, EPA_FAILED };

struct Epa {
	EpaScratch* m;
	int status;
	V3 normal; float depth;
	int nextsv;
	int hullRoot, hullCount, stockRoot, stockCount;
	// result
	SV rc[3]; float rp[3]; int rrank;
};

RSB_HD void epa_list_append(Epa& e, int& root, int& count, int face) {
	EpaFace* fc = e.m->fc;
	fc[face].l[0] = -1; fc[face].l[1] = (int16_t)root;
	if (root >= 0) fc[root].l[0] = (int16_t)face;
	root = face; ++count;
}
RSB_HD void epa_list_remove(Epa& e, int& root, int& count, int face) {
	EpaFace* fc = e.m->fc;
	if (fc[face].l[1] >= 0) fc[fc[face].l[1]].l[0] = fc[face].l[0];
	if (fc[face].l[0] >= 0) fc[fc[face].l[0]].l[1] = fc[face].l[1];
	if (face == root) root = fc[face].l[1];
	--count;
}
RSB_HD void epa_bind(Epa& e, int fa, unsigned ea, int fb, unsigned eb) {
	EpaFace* fc = e.m->fc;
	fc[fa].e[ea] = (uint8_t)eb; fc[fa].f[ea] = (int16_t)fb;
	fc[fb].e[eb] = (uint8_t)ea; fc[fb].f[eb] = (int16_t)fa;
}
RSB_HD bool epa_getedgedist(const V3& fn, const V3& aw, const V3& bw, float& dist) {
	const V3 ba = bw - aw;
	const V3 n_ab = cross(ba, fn);
	const float a_dot_nab = dot(aw, n_ab);
	if (a_dot_nab < 0.f) {
		const float ba_l2 = len2(ba);
		const float a_dot_ba = dot(aw, ba);
		const float b_dot_ba = dot(bw, ba);
		if (a_dot_ba > 0.f) dist = len(aw);
		else if (b_dot_ba < 0.f) dist = len(bw);
		else {
			const float a_dot_b = dot(aw, bw);
			float v = (len2(aw) * len2(bw) - a_dot_b * a_dot_b) / ba_l2;
			dist = sqrtf(v > 0.f ? v : 0.f);
		}
		return true;
	}
	return false;
}
RSB_HD int epa_newface(Epa& e, int a, int b, int c, bool forced) {
	if (e.stockRoot >= 0) {
		int face = e.stockRoot;
		epa_list_remove(e, e.stockRoot, e.stockCount, face);
		epa_list_append(e, e.hullRoot, e.hullCount, face);
		EpaFace& f = e.m->fc[face];
		const V3 aw = e.m->sv[a].w, bw = e.m->sv[b].w, cw = e.m->sv[c].w;
		f.pass = 0; f.c[0] = (int16_t)a; f.c[1] = (int16_t)b; f.c[2] = (int16_t)c;
		f.n = cross(bw - aw, cw - aw);
		const float l = len(f.n);
		const bool v = l > EPA_ACCURACY;
		if (v) {
			if (!(epa_getedgedist(f.n, aw, bw, f.d) || epa_getedgedist(f.n, bw, cw, f.d) || epa_getedgedist(f.n, cw, aw, f.d)))
				f.d = dot(aw, f.n) / l;
			f.n = f.n / l;
			if (forced || (f.d >= -EPA_PLANE_EPS)) return face;
			else e.status = EPA_NONCONVEX;
		} else e.status = EPA_DEGENERATED;
		epa_list_remove(e, e.hullRoot, e.hullCount, face);
		epa_list_append(e, e.stockRoot, e.stockCount, face);
		return -1;
	}
	e.status = EPA_OUTOFFACES;
	return -1;
}
RSB_HD int epa_findbest(Epa& e) {
	const EpaFace* fc = e.m->fc;
	int minf = e.hullRoot;
	float mind = fc[minf].d * fc[minf].d;
	for (int f = fc[minf].l[1]; f >= 0; f = fc[f].l[1]) {
		const float sqd = fc[f].d * fc[f].d;
		if (sqd < mind) { minf = f; mind = sqd; }
	}
	return minf;
}
// EPA::expand (btGjkEpa2.cpp:867-908) with an explicit frame stack
RSB_HD bool epa_expand(Epa& e, unsigned pass, int w, int f0, unsigned e0, int& hcf, int& hff, unsigned& hnf) {
	const unsigned i1m3[3] = {1, 2, 0}, i2m3[3] = {2, 0, 1};
	EpaScratch& m = *e.m;
	int sp = 0;
	m.stackF[0] = (int16_t)f0; m.stackE[0] = (uint8_t)e0; m.stackS[0] = 0; sp = 1;
	bool ret = false;
	while (sp > 0) {
		const int f = m.stackF[sp - 1]; const unsigned ed = m.stackE[sp - 1]; const int stage = m.stackS[sp - 1];
		EpaFace& F = m.fc[f];
		if (stage == 0) {
			if (F.pass == (uint8_t)pass) { ret = false; sp--; continue; }
			const unsigned e1 = i1m3[ed];
			if ((dot(F.n, m.sv[w].w) - F.d) < -EPA_PLANE_EPS) {
				int nf = epa_newface(e, F.c[e1], F.c[ed], w, false);
				if (nf >= 0) {
					epa_bind(e, nf, 0, f, ed);
					if (hcf >= 0) epa_bind(e, hcf, 1, nf, 2); else hff = nf;
					hcf = nf; ++hnf;
					ret = true;
				} else ret = false;
				sp--; continue;
			}
			F.pass = (uint8_t)pass;
			m.stackS[sp - 1] = 1;
			if (sp >= EPA_MAX_FACES + 8) { ret = false; sp--; continue; }
			m.stackF[sp] = F.f[e1]; m.stackE[sp] = F.e[e1]; m.stackS[sp] = 0; sp++;
			continue;
		}
		if (stage == 1) {
			if (!ret) { sp--; continue; }
			const unsigned e2 = i2m3[ed];
			m.stackS[sp - 1] = 2;
			if (sp >= EPA_MAX_FACES + 8) { ret = false; sp--; continue; }
			m.stackF[sp] = F.f[e2]; m.stackE[sp] = F.e[e2]; m.stackS[sp] = 0; sp++;
			continue;
		}
		if (ret) {
			epa_list_remove(e, e.hullRoot, e.hullCount, f);
			epa_list_append(e, e.stockRoot, e.stockCount, f);
		}
		sp--;
	}
	return ret;
}

// EPA::Evaluate (btGjkEpa2.cpp:645-772)
RSB_HD_NOINLINE int epa_evaluate(Epa& e, Gjk2& gjk, V3 guess) {
	EpaScratch& m = *e.m;
	e.status = EPA_FAILED; e.normal = v3zero(); e.depth = 0.f; e.nextsv = 0;
	e.hullRoot = -1; e.hullCount = 0; e.stockRoot = -1; e.stockCount = 0;
	for (int i = 0; i < EPA_MAX_FACES; ++i) epa_list_append(e, e.stockRoot, e.stockCount, EPA_MAX_FACES - i - 1);
	Simplex2& simplex = gjk.simplex;
	if ((simplex.rank > 1) && gjk2_enclose_origin(gjk)) {
		e.status = EPA_VALID;
		e.nextsv = 0;
		if (triple3(simplex.c[0].w - simplex.c[3].w, simplex.c[1].w - simplex.c[3].w, simplex.c[2].w - simplex.c[3].w) < 0.f) {
			SV t = simplex.c[0]; simplex.c[0] = simplex.c[1]; simplex.c[1] = t;
			float tp = simplex.p[0]; simplex.p[0] = simplex.p[1]; simplex.p[1] = tp;
		}
		// simplex vertices live in sv[0..3]; EPA's own vertices follow
		for (int i = 0; i < 4; i++) m.sv[i] = simplex.c[i];
		int tetra[4] = {epa_newface(e, 0, 1, 2, true), epa_newface(e, 1, 0, 3, true), epa_newface(e, 2, 1, 3, true), epa_newface(e, 0, 2, 3, true)};
		if (e.hullCount == 4) {
			int best = epa_findbest(e);
			EpaFace outer = m.fc[best];
			unsigned pass = 0;
			unsigned iterations = 0;
			epa_bind(e, tetra[0], 0, tetra[1], 0);
			epa_bind(e, tetra[0], 1, tetra[2], 0);
			epa_bind(e, tetra[0], 2, tetra[3], 0);
			epa_bind(e, tetra[1], 1, tetra[3], 2);
			epa_bind(e, tetra[1], 2, tetra[2], 1);
			epa_bind(e, tetra[2], 2, tetra[3], 1);
			e.status = EPA_VALID;
			for (; iterations < (unsigned)EPA_MAX_ITERATIONS; ++iterations) {
				if (e.nextsv < EPA_MAX_VERTICES) {
					int hcf = -1, hff = -1; unsigned hnf = 0;
					const int w = 4 + e.nextsv++;
					bool valid = true;
					m.fc[best].pass = (uint8_t)(++pass);
					gjk2_getsupport(gjk.shape, m.fc[best].n, m.sv[w]);
					const float wdist = dot(m.fc[best].n, m.sv[w].w) - m.fc[best].d;
					if (wdist > EPA_ACCURACY) {
						for (unsigned j = 0; (j < 3) && valid; ++j)
							valid &= epa_expand(e, pass, w, m.fc[best].f[j], m.fc[best].e[j], hcf, hff, hnf);
						if (valid && (hnf >= 3)) {
							epa_bind(e, hcf, 1, hff, 2);
							epa_list_remove(e, e.hullRoot, e.hullCount, best);
							epa_list_append(e, e.stockRoot, e.stockCount, best);
							best = epa_findbest(e);
							outer = m.fc[best];
						} else { e.status = EPA_INVALIDHULL; break; }
					} else { e.status = EPA_ACCURACYREACHED; break; }
				} else { e.status = EPA_OUTOFVERTICES; break; }
			}
			const V3 projection = outer.n * outer.d;
			e.normal = outer.n;
			e.depth = outer.d;
			e.rrank = 3;
			e.rc[0] = m.sv[outer.c[0]]; e.rc[1] = m.sv[outer.c[1]]; e.rc[2] = m.sv[outer.c[2]];
			e.rp[0] = len(cross(e.rc[1].w - projection, e.rc[2].w - projection));
			e.rp[1] = len(cross(e.rc[2].w - projection, e.rc[0].w - projection));
			e.rp[2] = len(cross(e.rc[0].w - projection, e.rc[1].w - projection));
			const float sum = e.rp[0] + e.rp[1] + e.rp[2];
			e.rp[0] /= sum; e.rp[1] /= sum; e.rp[2] /= sum;
			return e.status;
		}
	}
	e.status = EPA_FALLBACK;
	e.normal = -guess;
	const float nl = len(e.normal);
	if (nl > 0.f) e.normal = e.normal / nl; else e.normal = V3(1.f, 0.f, 0.f);
	e.depth = 0.f;
	e.rrank = 1;
	e.rc[0] = simplex.c[0];
	e.rp[0] = 1.f;
	return e.status;
}

// btGjkEpaPenetrationDepthSolver::calcPenDepth
RSB_HD_NOINLINE bool epa_penetration(const Ctx& cx, Convex A, Xform tA, Convex B, Xform tB, V3& v, V3& wa, V3& wb) {
#if defined(RSB_PROFILE)
	if (cx.epa) cx.epa->profCalls += 256;
#endif
	const V3 guessVectors[9] = {
		safe_normalized(tB.origin - tA.origin), safe_normalized(tA.origin - tB.origin),
		V3(0, 0, 1), V3(0, 1, 0), V3(1, 0, 0), V3(1, 1, 0), V3(1, 1, 1), V3(0, 1, 1), V3(1, 0, 1)};
	for (int i = 0; i < 9; i++) {
		const V3 guess = guessVectors[i];
		{   // Penetration(..., usemargins = true)
			Gjk2 gjk;
			mink_init(gjk.shape, A, tA, B, tB, true);
			int st = gjk2_evaluate(gjk, -guess);
			if (st == G2_INSIDE) {
				Epa epa; epa.m = cx.epa;
				int es = epa_evaluate(epa, gjk, -guess);
				if (es != EPA_FAILED) {
					V3 w0 = v3zero();
					for (int k = 0; k < epa.rrank; ++k) w0 += mink_support0(gjk.shape, epa.rc[k].d) * epa.rp[k];
					wa = xf_apply(tA, w0);
					wb = xf_apply(tA, w0 - epa.normal * epa.depth);
					v = -epa.normal;
					return true;
				}
			}
		}
		{   // Distance(...) without margins
			Gjk2 gjk;
			mink_init(gjk.shape, A, tA, B, tB, false);
			int st = gjk2_evaluate(gjk, guess);
			if (st == G2_VALID) {
				V3 w0 = v3zero(), w1 = v3zero();
				for (int k = 0; k < gjk.simplex.rank; ++k) {
					const float p = gjk.simplex.p[k];
					w0 += mink_support0(gjk.shape, gjk.simplex.c[k].d) * p;
					w1 += mink_support1(gjk.shape, -gjk.simplex.c[k].d) * p;
				}
				wa = xf_apply(tA, w0);
				wb = xf_apply(tA, w1);
				V3 n = w0 - w1;
				float dist = len(n);
				v = n / (dist > GJK2_MIN_DISTANCE ? dist : 1.f);
				return false;
			}
		}
	}
	wa = v3zero(); wb = v3zero(); v = v3zero();
	return false;
}

}  // namespace rsb
