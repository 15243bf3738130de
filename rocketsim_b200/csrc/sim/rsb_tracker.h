// rsb_tracker.h -- shot / goal-credit / save detection on the device, one arena per call.
// Restates
//   Arena::IsBallProbablyGoingIn   src/Sim/Arena/Arena.cpp:771-891 (+ BallWithinHoopsGoalXYMarginSq :760-769)
//   GameEventTracker::Update       src/Sim/GameEventTracker/GameEventTracker.cpp:5-157, state GameEventTracker.h:72-87
// The reference runs the tracker on the host after Step(); here it runs at the end of every tick of an arena whose batch
// enabled it (rsb_set_event_tracker), i.e. "Update() every tick", and reports through the batch's event buffer.
#pragma once
#include "rsb_work.h"

namespace rsb {

RSB_HD float hoops_goal_xy_margin_sq(float x, float y) {
	const float SCALE_Y = 0.9f, OFFSET_Y = 2770.f, RADIUS_SQ = 716 * 716;
	float dy = fabsf(y) * SCALE_Y - OFFSET_Y;
	float distSq = x * x + dy * dy;
	return distSq - RADIUS_SQ;
}

// team ids: 0 = BLUE, 1 = ORANGE; RS_TEAM_FROM_Y(y) = y < 0 ? BLUE : ORANGE
RSB_HD bool is_ball_probably_going_in(const Ctx& cx, float maxTime, float extraMargin, int& goalTeamOut, bool& supported) {
	const float kFltEps = 1.1920929e-7f;
	const MutD& mut = cx.p->mut;
	const V3 ballPos = cx.s->ball.rb.pos * BT_TO_UU;
	const V3 ballVel = cx.s->ball.rb.vel * BT_TO_UU;
	supported = true;
	const int mode = cx.mode;
	if (mode == SOCCAR || mode == SNOWDAY) {
		if (fabsf(ballVel.y) < kFltEps) return false;
		float scoreDirSgn = (float)rs_sgn(ballVel.y);
		float goalY = mut.goalBaseThresholdY * scoreDirSgn;
		float distToGoal = fabsf(ballPos.y - goalY);
		float timeToGoal = distToGoal / fabsf(ballVel.y);
		if (timeToGoal > maxTime) return false;
		V3 extrap = ballPos + (ballVel * timeToGoal) + (mut.gravity * timeToGoal * timeToGoal) / 2.f;
		const float APPROX_GOAL_HALF_WIDTH = 892.755f, APPROX_GOAL_HEIGHT = 642.775f;
		float scoreMargin = mut.ballRadius * 0.1f + extraMargin;
		if (extrap.z > APPROX_GOAL_HEIGHT + scoreMargin) return false;
		if (fabsf(extrap.x) > APPROX_GOAL_HALF_WIDTH + scoreMargin) return false;
		goalTeamOut = scoreDirSgn < 0.f ? 0 : 1;
		return true;
	} else if (mode == HOOPS) {
		const float APPROX_RIM_HEIGHT = 365.f;
		float minHeight = APPROX_RIM_HEIGHT + mut.ballRadius * 1.2f;
		if (ballVel.z < -kFltEps && ballPos.z < minHeight) {
			if (hoops_goal_xy_margin_sq(ballPos.x, ballPos.y) < 0.f) { goalTeamOut = ballPos.y < 0.f ? 0 : 1; return true; }
		}
		float margin = mut.ballRadius * 1.0f;
		float marginSq = margin * margin;
		float upQuadIntercept, downQuadIntercept;
		{
			float g = mut.gravity.z;
			if (g > -kFltEps) return false;
			float v = ballVel.z;
			float h = ballPos.z - minHeight;
			float sqrtInput = v * v - 2.f * g * h;
			if (sqrtInput > 0.f) {
				float sqrtOutput = sqrtf(sqrtInput);
				upQuadIntercept = (-v + sqrtOutput) / g;
				downQuadIntercept = (-v - sqrtOutput) / g;
			} else {
				return hoops_goal_xy_margin_sq(ballPos.x, ballPos.y) < -marginSq;  // goalTeamOut untouched, as in the reference
			}
		}
		if (upQuadIntercept >= 0.f) {
			V3 up = ballPos + (ballVel * upQuadIntercept);
			float upMarginSq = hoops_goal_xy_margin_sq(up.x, up.y);
			float minClearanceMargin = 60.f + mut.ballRadius;
			if (upMarginSq > -marginSq && upMarginSq < (minClearanceMargin * minClearanceMargin)) return false;
		}
		V3 down = ballPos + (ballVel * downQuadIntercept);
		down.y = fabsf(down.y);
		{
			float wallBounceY = K::ARENA_EXTENT_Y_HOOPS - mut.ballRadius;
			if (down.y > wallBounceY) {
				float m2 = down.y - wallBounceY;
				down.y -= m2 * (1.f + mut.ballWorldRestitution);
			}
		}
		if (hoops_goal_xy_margin_sq(down.x, down.y) < -marginSq) { goalTeamOut = down.y < 0.f ? 0 : 1; return true; }
		return false;
	}
	supported = false;  // the reference raises for the other modes
	return false;
}

// GetShooterPasser (GameEventTracker.cpp:5-47).  Cars are visited in slot order (the reference walks an unordered_set, which
// only matters when two team-mates touched the ball in the very same tick).  Returns car indices or -1.
RSB_HD bool tracker_shooter_passer(const Ctx& cx, uint64_t now, int team, int& shooter, bool findPasser, int& passer, uint64_t maxShooterTicks, uint64_t maxPasserTicks) {
	shooter = passer = -1;
	const int nCars = cx.p->nCars;
	for (int i = 0; i < nCars; i++) {
		const CarS& c = cx.s->cars[i];
		if (cx.p->team[i] != team || !c.bhValid) continue;
		if (c.bhTickHit + maxShooterTicks >= now)
			if (shooter < 0 || c.bhTickHit > cx.s->cars[shooter].bhTickHit) shooter = i;
	}
	if (shooter >= 0 && findPasser) {
		const uint64_t shootTick = cx.s->cars[shooter].bhTickHit;
		for (int i = 0; i < nCars; i++) {
			const CarS& c = cx.s->cars[i];
			if (cx.p->team[i] != team || !c.bhValid || i == shooter) continue;
			if (c.bhTickHit + maxPasserTicks >= shootTick)
				if (passer < 0 || c.bhTickHit > cx.s->cars[passer].bhTickHit) passer = i;
		}
	}
	return shooter >= 0;
}

// GameEventTracker::Update with autoStateSetDetection = true.  `now` = Arena::tickCount as Update() sees it after Step().
RSB_HD void tracker_update(const Ctx& cx, bool scored, uint64_t now) {
	ArenaS& S = *cx.s;
	const TrackerCfgD& cfg = cx.p->tracker;
	const float tickrate = 1.f / cx.p->dt;  // Arena::GetTickRate
	const uint64_t ballUpdateCount = S.ball.tickCountSinceUpdate;
	if (ballUpdateCount > S.trkLastBallUpdate) {
		const uint64_t deltaTicks = ballUpdateCount - S.trkLastBallUpdate;
		const float deltaTime = (float)deltaTicks * cx.p->dt;
		if (scored && !S.trkScoredLast) {
			int shooter, passer;
			const float by = -S.ball.rb.pos.y;
			if (tracker_shooter_passer(cx, now, by < 0.f ? 0 : 1, shooter, true, passer, (uint64_t)(cfg.goalMaxTouchTime * tickrate), (uint64_t)(cfg.passMaxTouchTime * tickrate)))
				push_event_at(cx, 5u /*RSB_EVENT_GOAL_CREDIT*/, (uint32_t)(shooter + 1), (uint32_t)(passer + 1), now);
		} else if (!S.trkBallShot) {
			if (S.trkShotCooldown > 0.f) {
				S.trkShotCooldown = fmaxf_(S.trkShotCooldown - deltaTime, 0.f);
			} else {
				const float speedSq = len2(S.ball.rb.vel * BT_TO_UU);
				if (speedSq >= cfg.shotMinSpeed * cfg.shotMinSpeed) {
					int goalTeam = 0; bool supported;
					if (is_ball_probably_going_in(cx, cfg.shotMinScoreTime, cfg.predScoreExtraMargin, goalTeam, supported)) {
						const int shooterTeam = goalTeam == 0 ? 1 : 0;
						const uint64_t shotMinTouchDelayTicks = (uint64_t)(cfg.shotTouchMinDelay * tickrate);
						int shooter, passer;
						if (tracker_shooter_passer(cx, now, shooterTeam, shooter, true, passer, deltaTicks + shotMinTouchDelayTicks, (uint64_t)(cfg.passMaxTouchTime * tickrate))) {
							const uint64_t ticksSinceHit = now - S.cars[shooter].bhTickHit;
							if (ticksSinceHit >= shotMinTouchDelayTicks) {
								S.trkBallShot = 1; S.trkShotGoalTeam = (uint8_t)goalTeam; S.trkShotCooldown = cfg.shotEventCooldown;
								push_event_at(cx, 4u /*RSB_EVENT_SHOT*/, (uint32_t)(shooter + 1), (uint32_t)(passer + 1), now);
							}
						}
					}
				}
			}
		} else {
			int team; bool supported;
			const bool willScore = is_ball_probably_going_in(cx, cfg.shotMinScoreTime, cfg.predScoreExtraMargin, team, supported);
			if (!willScore) {
				int saver, unused;
				if (tracker_shooter_passer(cx, now, S.trkShotGoalTeam, saver, false, unused, deltaTicks, 0))
					push_event_at(cx, 6u /*RSB_EVENT_SAVE*/, (uint32_t)(saver + 1), 0u, now);
				S.trkBallShot = 0;
			}
		}
	} else if (ballUpdateCount == S.trkLastBallUpdate) {
		return;
	} else {
		S.trkScoredLast = 0; S.trkBallShot = 0; S.trkShotCooldown = 0.f;  // ResetPersistentInfo
	}
	S.trkScoredLast = scored ? 1 : 0;
	S.trkLastBallUpdate = ballUpdateCount;
}

}  // namespace rsb
