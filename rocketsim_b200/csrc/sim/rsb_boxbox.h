// rsb_boxbox.h -- box-box contact generation (car-car).
#pragma once
#include "rsb_work.h"
#include "rsb_simplex.h"
namespace rsb {
RSB_HD_NOINLINE void box_box(const Ctx& cx, int mi, const Convex& boxA, const Xform& xfA, const Convex& boxB, const Xform& xfB) {
	(void)cx; (void)mi; (void)boxA; (void)xfA; (void)boxB; (void)xfB;  // TODO
}
}
