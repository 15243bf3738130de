// rsb_epa.h -- penetration depth for overlapping convex cores (GJK + EPA).
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"
namespace rsb {
RSB_HD_NOINLINE bool epa_penetration(const Ctx& cx, const Convex& A, const Xform& tA, const Convex& B, const Xform& tB, V3& v, V3& wa, V3& wb) {
	cx.s->errorFlags |= ERR_EPA_FAIL;  // TODO: EPA
	(void)A; (void)tA; (void)B; (void)tB; (void)v; wa = v3zero(); wb = v3zero();
	return false;
}
}
