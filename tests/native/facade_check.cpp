// Drives the C++ front end (include/rocketsim_b200.hpp) the way reference user code drives RocketSim::Arena and dumps
// the resulting states; tests/test_cpp_facade.py runs the same script on the oracle and compares.
//   usage: facade_check <mode> <ticks> <seed> <mesh.cmf>...      (meshes belong to <mode>)
#include <cstdio>
#include <cstdlib>
#include "rocketsim_b200.hpp"

using namespace RocketSimB200;

static void dump(const char* tag, const void* p, size_t n) {
	printf("%s ", tag);
	for (size_t i = 0; i < n; i++) printf("%02x", ((const unsigned char*)p)[i]);
	printf("\n");
}

int main(int argc, char** argv) {
	if (argc < 4) return 2;
	const GameMode mode = (GameMode)atoi(argv[1]);
	const int ticks = atoi(argv[2]), seed = atoi(argv[3]);
	try {
		std::map<GameMode, std::vector<std::string>> paths;
		for (int i = 4; i < argc; i++) paths[mode].push_back(argv[i]);
		Init(paths, true);

		int goals = 0, bumps = 0; Team lastScorer = Team::BLUE;
		Arena* arena = Arena::Create(mode, ArenaConfig(), 120);  // the reference's signature (Arena.h:111)
		{ Car* tmp = arena->AddCar(Team::ORANGE); if (!arena->RemoveCar(tmp->id) || arena->RemoveCar(77)) throw std::runtime_error("RemoveCar"); }
		Car* a = arena->AddCar(Team::BLUE);
		Car* b = arena->AddCar(Team::ORANGE, CAR_CONFIG_DOMINUS());
		arena->SetGoalScoreCallback([&](Arena*, Team t, void*) { goals++; lastScorer = t; });
		arena->SetCarBumpCallback([&](Arena*, Car*, Car*, bool, void*) { bumps++; });
		arena->ResetToRandomKickoff(seed);
		for (int t = 0; t < ticks; t++) {
			a->controls = {}; b->controls = {};
			a->controls.throttle = 1; a->controls.boost = 1; a->controls.steer = (t % 40 < 20) ? 0.3f : -0.2f;
			b->controls.throttle = 1; b->controls.jump = (t % 50) < 5; b->controls.pitch = -0.5f;
			arena->Step(1);
		}
		CarState sa = a->GetState(), sb = b->GetState();
		BallState bs = arena->ball->GetState();
		dump("carA", &sa, offsetof(CarState, lastControls));
		dump("carB", &sb, offsetof(CarState, lastControls));
		dump("ball", &bs, 104);
		printf("ticks %llu scored %d\n", (unsigned long long)arena->tickCount, (int)arena->IsBallScored());

		// goal callback: fire the ball into the orange goal
		bs.pos = Vec{600, 4000, 200, 0}; bs.vel = Vec{0, 5000, 0, 0};
		arena->ball->SetState(bs);
		for (int t = 0; t < 60 && goals == 0; t++) arena->Step(1);
		printf("goals %d scorer %d bumps %d\n", goals, (int)lastScorer, bumps);

		// error behaviour: AddCar after the first Step throws, bad tick rate throws (Arena.cpp:419)
		int errs = 0;
		try { arena->AddCar(Team::BLUE); } catch (const std::runtime_error&) { errs++; }
		try { Arena::Create(mode, ArenaConfig(), 10); } catch (const std::runtime_error&) { errs++; }
		try { ArenaConfig bad; bad.maxAABBLen = 500; Arena* x = Arena::Create(mode, bad); x->Step(1); } catch (const std::runtime_error&) { errs++; }
		printf("errors %d\n", errs);
		delete arena;
	} catch (const std::exception& e) {
		printf("EXCEPTION %s\n", e.what());
		return 1;
	}
	return 0;
}
