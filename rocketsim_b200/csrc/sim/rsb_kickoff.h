// rsb_kickoff.h -- Arena::ResetToRandomKickoff (reference src/Sim/Arena/Arena.cpp:113-223) as a
// device function: used by the reset kernel behind rsb_reset_to_random_kickoff and, when the
// batch has auto-reset enabled, by the tick itself right after a goal (the README bench loop
// "reset on IsBallScored()" without a host round trip).
#pragma once
#include "rsb_work.h"
#include "rsb_car.h"

namespace rsb {

// std::uniform_int_distribution<size_t>(0, range-1) on minstd_rand0 (libstdc++ downscaling path)
RSB_HD uint32_t minstd_uniform(uint32_t& x, uint32_t range) {
	const uint64_t urngrange = 2147483646ull - 1ull;  // max - min
	const uint64_t scaling = urngrange / range;
	const uint64_t past = range * scaling;
	uint64_t ret;
	do { ret = (uint64_t)minstd_next(x) - 1ull; } while (ret >= past);
	return (uint32_t)(ret / scaling);
}

// Kickoff spawn table (RLConst.h:343-381), blue side
RSB_HD K::SpawnPos kickoff_slot(int mode, int i) {
	const float P4 = (float)(3.14159265358979323846 / 4), P2 = (float)(3.14159265358979323846 / 2);
	if (mode == HOOPS) {
		const K::SpawnPos t[5] = {{-1536, -3072, P4 * 2}, {1536, -3072, P4 * 2}, {-256, -2816, P4 * 2}, {256, -2816, P4 * 2}, {0, -3200, P4 * 2}};
		return t[i];
	}
	if (mode == HEATSEEKER) {
		const K::SpawnPos t[5] = {{-1000, -4620, P2}, {1000, -4620, P2}, {-2000, -4620, P2}, {2000, -4620, P2}, {0, 0, 0}};
		return t[i];
	}
	if (mode == DROPSHOT) {  // CAR_SPAWN_LOCATIONS_DROPSHOT (RLConst.h:381-388)
		const K::SpawnPos t[5] = {{-1867, -2380, P4 * 1}, {1867, -2380, P4 * 3}, {-256, -3576, P4 * 2}, {256, -3576, P4 * 2}, {0, -4088, P4 * 2}};
		return t[i];
	}
	const K::SpawnPos t[5] = {{-2048, -2560, P4 * 1}, {2048, -2560, P4 * 3}, {-256, -3840, P4 * 2}, {256, -3840, P4 * 2}, {0, -4608, P4 * 2}};
	return t[i];
}

RSB_HD void dropshot_reset(ArenaS& G) {
	G.dsChargeLevel = 1; G.dsAccumulatedHitForce = 0.f; G.dsYTargetDir = 0.f; G.dsHasDamaged = 0u; G.dsLastDamageTick = 0ull;
	for (int t = 0; t < MAX_TILES; t++) G.tiles[t] = 0;
}
// Arena::SetDropshotTilesState(_dropshotTilesState): tile bodies take their no-contact-response flag from the damage states
RSB_HD void dropshot_sync_tiles(ArenaS& G) {
	for (int t = 0; t < MAX_TILES; t++) G.tiles[t] = (uint8_t)((G.tiles[t] & 3u) | (((G.tiles[t] & 3u) == 2u) ? 4u : 0u));
}

RSB_HD void kickoff_reset(ArenaS& A, ArenaS& G, const BatchParams& P, int seed) {
	uint32_t localRng = minstd_seed_int(seed);
	uint32_t& rng = (seed == -1) ? A.rng : localRng;
	int order[5] = {0, 1, 2, 3, 4};
	const int L = (P.mode == HEATSEEKER) ? 4 : 5;
	// std::shuffle(first, first+L, g) (libstdc++ bits/stl_algo.h: pairwise draws when range*range fits)
	{
		const uint64_t urngrange = 2147483646ull - 1ull;
		const uint64_t urange = (uint64_t)L;
		if (urngrange / urange >= urange) {
			int it = 1;
			if ((urange % 2) == 0) { uint32_t j = minstd_uniform(rng, 2); int t = order[it]; order[it] = order[j]; order[j] = t; it++; }
			while (it < L) {
				const uint32_t swap_range = (uint32_t)it + 1;
				uint32_t x = minstd_uniform(rng, swap_range * (swap_range + 1));
				uint32_t p0 = x / (swap_range + 1), p1 = x % (swap_range + 1);
				int t = order[it]; order[it] = order[p0]; order[p0] = t; it++;
				t = order[it]; order[it] = order[p1]; order[p1] = t; it++;
			}
		} else {
			for (int it = 1; it < L; it++) { uint32_t j = minstd_uniform(rng, (uint32_t)it + 1); int t = order[it]; order[it] = order[j]; order[j] = t; }
		}
	}
	int nBlue = 0, nOrange = 0;
	int blue[MAX_CARS], orange[MAX_CARS];
	for (int c = 0; c < P.nCars; c++) { if (P.team[c] == 0) blue[nBlue++] = c; else orange[nOrange++] = c; }
	int numAtRespawn[4] = {0, 0, 0, 0};
	int amount = nBlue > nOrange ? nBlue : nOrange;
	for (int k = 0; k < amount; k++) {
		K::SpawnPos sp;
		int kickIdx = -1, respIdx = -1;
		if (k < L) {
			int idx = order[k] < L - 1 ? order[k] : L - 1;
			sp = kickoff_slot(P.mode, idx);
			kickIdx = idx;
		} else {
			int ridx = (k - L) % L;
			respIdx = ridx;
			// CAR_RESPAWN_LOCATIONS_DROPSHOT (RLConst.h:407-413) is only used here; Car::Respawn takes the Soccar table in Dropshot
			if (P.mode == DROPSHOT) {
				const float H = (float)(3.14159265358979323846 / 2);
				const K::SpawnPos t[4] = {{-2176, -3410, H}, {-1152, -3100, H}, {2176, -3410, H}, {1152, -3100, H}};
				sp = t[ridx];
			} else sp = respawn_slot(P.mode == HOOPS, ridx);
			sp.y += 250.f * numAtRespawn[ridx];
			numAtRespawn[ridx]++;
		}
		for (int team = 0; team < 2; team++) {
			bool isBlue = team == 0;
			int cnt = isBlue ? nBlue : nOrange;
			if (k >= cnt) continue;
			CarS& c = A.cars[isBlue ? blue[k] : orange[k]];
			car_reset_state(c, P.mut.carSpawnBoostAmount);
			V3 pos(sp.x, sp.y, K::CAR_SPAWN_REST_Z);
			if (!isBlue) pos = pos * V3(-1.f, -1.f, 1.f);
			c.rb.pos = pos * UU_TO_BT;
			// Angle(yaw + (blue ? 0 : pi), 0, 0).ToRotMat(), host-computed (BatchParams)
			c.rb.basis = kickIdx >= 0 ? P.kickBasis[isBlue ? 0 : 1][kickIdx] : P.respBasis[isBlue ? 0 : 1][respIdx];
			c.rb.vel = v3zero(); c.rb.angvel = v3zero();
			c.isOnGround = 1;
		}
	}
	BallS& b = A.ball;
	b.rb.basis = m3identity(); b.rb.angvel = v3zero();
	V3 bpos(0.f, 0.f, K::BALL_REST_Z), bvel(0.f, 0.f, 0.f);
	b.hsYTargetDir = 0.f; b.hsCurTargetSpeed = K::HS_INITIAL_TARGET_SPEED; b.hsTimeSinceHit = 0.f;
	if (P.mode == HEATSEEKER) {
		int nextRand = (int)minstd_next(rng);
		V3 scale(1.f, (nextRand % 2) ? 1.f : -1.f, 1.f);
		bpos = V3(-1000, -2220, 92.75f) * scale;
		bvel = V3(0, -65, 650) * scale;
	} else if (P.mode == SNOWDAY) {
		bvel.z = kEps;
	}
	b.rb.pos = bpos * UU_TO_BT; b.rb.vel = bvel * UU_TO_BT;
	b.velImpulseCache = v3zero(); b.tickCountSinceUpdate = 0;
	for (int p = 0; p < P.nPads; p++) { A.pads[p].isActive = 1; A.pads[p].cooldown = 0.f; A.pads[p].prevLockedCarID = 0; }
	// ball->SetState(BallState()) resets dsInfo; SetDropshotTilesState({}) restores every tile (Arena.cpp:205-217).  G = the arena's
	// record in HBM (== A outside the step kernel)
	if (P.mode == DROPSHOT) dropshot_reset(G);
}

}  // namespace rsb
