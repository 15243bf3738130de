// rsb_api.cu -- C ABI (include/rsb_c_api.h), batch management and the CUDA kernels of the
// batched Arena::Step for sm_100a.
//
// One thread block owns a chunk of arenas for a whole launch: their persistent records (ArenaS)
// are staged HBM -> shared memory once, n ticks run as block-synchronous item-parallel phases
// (rsb_tick.h: one thread per wheel / car / body / arena depending on the phase), and the
// records go back to HBM once.  rsb_step(n) is ONE launch that loops n ticks inside the kernel.
//
// Build modes:
//   nvcc (product):  the only execution path is CUDA; every entry point fails loudly
//                    when no device is available.  No CPU fallback exists in this build.
//   g++ -DRSB_HOSTCHECK (tests only, tests/_hostbuild/): the same device functions compiled
//                    for the host so the CPU-only test tier can diff the kernel arithmetic
//                    against the oracle.  Never built or loaded by the rocketsim_b200 package.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rsb_c_api.h"
#include "../../include/rsb_bench_inputs.h"
#include "sim/rsb_tick.h"
#include "rsb_scene.h"
#include "rsb_serialize.h"

using namespace rsb;

#ifndef RSB_HOSTCHECK
#include <cuda_runtime.h>
#define RSB_GLOBAL __global__
#else
#define RSB_GLOBAL
#endif

static thread_local std::string g_err;
static int fail(const std::string& s) { g_err = s; return -1; }

// ------------------------------------------------------------------ device memory shim
#ifndef RSB_HOSTCHECK
#define CUDA_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
static void* dev_alloc(size_t n) { void* p = nullptr; CUDA_OK(cudaMalloc(&p, n ? n : 1)); CUDA_OK(cudaMemset(p, 0, n ? n : 1)); return p; }
static void dev_free(void* p) { if (p) cudaFree(p); }
static void h2d(void* d, const void* h, size_t n, cudaStream_t s) { CUDA_OK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
static void d2h(void* h, const void* d, size_t n, cudaStream_t s) { CUDA_OK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); CUDA_OK(cudaStreamSynchronize(s)); }
typedef cudaStream_t stream_t;
#else
static void* dev_alloc(size_t n) { void* p = calloc(n ? n : 1, 1); if (!p) throw std::runtime_error("alloc"); return p; }
static void dev_free(void* p) { free(p); }
typedef int stream_t;
static void h2d(void* d, const void* h, size_t n, stream_t) { memcpy(d, h, n); }
static void d2h(void* h, const void* d, size_t n, stream_t) { memcpy(h, d, n); }
#endif

// ------------------------------------------------------------------ scenes (per game mode)
struct SceneDev {
	SceneHost host;
	Scene* dScene = nullptr;          // device copy of the Scene header
	std::vector<void*> allocs;
};
static std::mutex g_initMutex;
static bool g_initialized = false;
static std::map<int, std::vector<std::vector<uint8_t>>> g_meshFiles;
static std::map<int, SceneHost> g_sceneHost;                 // built at init
static std::map<std::pair<int, int>, SceneDev*> g_sceneDev;  // (device, mode) -> uploaded scene

static std::mutex g_sceneMutex;  // g_sceneDev / g_sceneHost lookups, uploads and inserts (batches may be created from several threads)
static SceneDev* get_scene_dev(int device, int mode) {
	std::lock_guard<std::mutex> lk(g_sceneMutex);
	auto key = std::make_pair(device, mode);
	auto it = g_sceneDev.find(key);
	if (it != g_sceneDev.end()) return it->second;
	auto hs = g_sceneHost.find(mode);
	if (hs == g_sceneHost.end()) throw std::runtime_error("No arena meshes registered for game mode " + std::to_string(mode) + " (call rsb_init_from_mem first)");
	SceneDev* sd = new SceneDev();
	sd->host = hs->second;
	Scene sc;
	memset(&sc, 0, sizeof(sc));
	const SceneHost& h = sd->host;
	sc.nMeshes = (int)h.meshes.size(); sc.nPlanes = (int)h.planes.size();
	auto up = [&](const void* src, size_t n) { void* d = dev_alloc(n); if (n) h2d(d, src, n, 0); sd->allocs.push_back(d); return d; };
	for (int i = 0; i < sc.nMeshes; i++) {
		const MeshHost& m = h.meshes[i];
		MeshD& d = sc.meshes[i];
		d.nodes = (const BvhNode*)up(m.nodes.data(), m.nodes.size() * sizeof(BvhNode)); d.nNodes = (int)m.nodes.size();
		d.nodeBounds = (const float*)up(m.nodeBounds.data(), m.nodeBounds.size() * 4);
		d.hdrs = (const SubtreeHdr*)up(m.hdrs.data(), m.hdrs.size() * sizeof(SubtreeHdr)); d.nHdrs = (int)m.hdrs.size();
		d.tris = (const int32_t*)up(m.tris.data(), m.tris.size() * 4);
		d.verts = (const float*)up(m.verts.data(), m.verts.size() * 4);
		d.info = (const TriInfo*)up(m.info.data(), m.info.size() * sizeof(TriInfo));
		d.nTris = m.nTris(); d.nVerts = m.nVerts();
		d.bvhMin = m.bvhMin; d.bvhMax = m.bvhMax; d.bvhQuant = m.bvhQuant;
		d.aabbMin = m.aabbMin; d.aabbMax = m.aabbMax; d.group = m.group; d.mask = m.mask;
	}
	for (int i = 0; i < sc.nPlanes; i++) sc.planes[i] = h.planes[i];
	sc.nTiles = (int)h.tiles.size();
	sc.tiles = sc.nTiles ? (const TileD*)up(h.tiles.data(), h.tiles.size() * sizeof(TileD)) : nullptr;
	sc.gridMin = h.gridMin; sc.gridMax = h.gridMax; sc.cellSize = h.cellSize;
	sc.cellsX = h.cellsX; sc.cellsY = h.cellsY; sc.cellsZ = h.cellsZ;
	sc.cellStart = (const int32_t*)up(h.cellStart.data(), h.cellStart.size() * 4);
	sc.cellStatics = (const uint8_t*)up(h.cellStatics.data(), h.cellStatics.size());
	sc.cellHdrStart = (const int32_t*)up(h.cellHdrStart.data(), h.cellHdrStart.size() * 4);
	sc.cellHdrs = (const uint16_t*)up(h.cellHdrs.data(), h.cellHdrs.size() * 2);
	sc.cellHdrsRay = (const uint16_t*)up(h.cellHdrsRay.data(), h.cellHdrsRay.size() * 2);
	sc.cellRayLeafStart = (const int32_t*)up(h.cellRayLeafStart.data(), h.cellRayLeafStart.size() * 4);
	sc.cellRayLeaves = (const int32_t*)up(h.cellRayLeaves.data(), h.cellRayLeaves.size() * 4);
	sc.rayLeafReach = h.rayLeafReach;
	sd->dScene = (Scene*)up(&sc, sizeof(sc));
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaDeviceSynchronize());
#endif
	g_sceneDev[key] = sd;
	return sd;
}

// ------------------------------------------------------------------ batch
struct RsbBatch {
	int device = 0;
	BatchParams hp;
	BatchParams* dP = nullptr;
	ArenaS* dS = nullptr;
	ArenaW* dW = nullptr;
	ArenaH* dH = nullptr;          // global mirror of the hot per-tick scratch
	EpaScratch* dE = nullptr;      // EPA working sets, one per thread of the step grid
	int* dChunkNext = nullptr;     // work counter of a multi-wave launch (chunks beyond the resident grid are fetched dynamically)
	int detectQueueCap = 0;        // 0 = the full queue; rsb_debug_set_detect_queue_cap (tests shrink it to exercise the overflow sweep)
	RsbBallState* traceOut = nullptr;  // consumed by the next launch (rsb_predict_ball)
	RsbBatch* pred = nullptr;      // ball-only sibling batch used by rsb_predict_ball (created on first use)
	RsbArenaConfig acfg;           // the ArenaConfig the batch was created with (customBoostPads points into acfgPads)
	std::vector<RsbBoostPadConfig> acfgPads;
	RsbMutatorConfig mutApi;       // the mutator config the batch was created with
	std::vector<RsbCarConfig> carCfgApi;  // the car configs the batch was created with (wire format, rsb_serialize_arena)
	float tickRate = 120.f;
	const Scene* dScene = nullptr;
	EventSink::Ev* dEvents = nullptr;
	uint32_t* dEventCount = nullptr;
	uint32_t eventCap = 0;
	CtrlS* dCtrl = nullptr;        // staging for rsb_set_controls
	CtrlS* dCtrlSeq = nullptr;     // staging for rsb_step_with_controls (host sequence)
	size_t ctrlSeqCap = 0;
	std::vector<EventSink::Ev> pendingEvents;  // downloaded, not yet handed to the caller (rsb_get_events with a small buffer)
	void* dScratch = nullptr;      // staging for get/set
	size_t scratchBytes = 0;
	stream_t stream = 0;
#ifndef RSB_HOSTCHECK
	cudaStream_t copyStream = nullptr;   // H2D of control-sequence chunks, overlapped with the previous chunk's launch
	cudaEvent_t chunkReady[8] = {};      // chunk c of the sequence has landed in dCtrlSeq
	cudaEvent_t seqFree = nullptr;       // every launch that reads dCtrlSeq has finished
#endif
	int numSMs = 1;
	size_t smemBudget = 0;
	void ensure_scratch(size_t n) { if (n > scratchBytes) { dev_free(dScratch); dScratch = dev_alloc(n); scratchBytes = n; } }
};

static void set_device(const RsbBatch* b) {
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaSetDevice(b->device));
#else
	(void)b;
#endif
}

// ------------------------------------------------------------------ conversions (device code)
RSB_HD RsbVec to_vec(const V3& v) { RsbVec r; r.x = v.x; r.y = v.y; r.z = v.z; r._w = 0.f; return r; }
RSB_HD V3 from_vec(const RsbVec& v) { return V3(v.x, v.y, v.z); }
RSB_HD void basis_to_rotmat(const M3& b, RsbRotMat& r) { r.forward = to_vec(b.col(0)); r.right = to_vec(b.col(1)); r.up = to_vec(b.col(2)); }
RSB_HD M3 rotmat_to_basis(const RsbRotMat& r) {
	return M3(V3(r.forward.x, r.right.x, r.up.x), V3(r.forward.y, r.right.y, r.up.y), V3(r.forward.z, r.right.z, r.up.z));
}
RSB_HD void ctrl_to_api(const CtrlS& c, RsbCarControls& o) { o.throttle = c.throttle; o.steer = c.steer; o.pitch = c.pitch; o.yaw = c.yaw; o.roll = c.roll; o.jump = c.jump; o.boost = c.boost; o.handbrake = c.handbrake; o._pad = 0; }
RSB_HD void ctrl_from_api(const RsbCarControls& c, CtrlS& o) { o.throttle = c.throttle; o.steer = c.steer; o.pitch = c.pitch; o.yaw = c.yaw; o.roll = c.roll; o.jump = c.jump != 0; o.boost = c.boost != 0; o.handbrake = c.handbrake != 0; o._pad = 0; }

// Car::GetState (Car.cpp:10-20)
RSB_HD void car_to_api(const ArenaS& A, int ci, RsbCarState& o) {
	const CarS& c = A.cars[ci];
	unsigned char* raw = (unsigned char*)&o;
	for (size_t i = 0; i < sizeof(RsbCarState); i++) raw[i] = 0;
	o.pos = to_vec(c.rb.pos * BT_TO_UU);
	// _internalState.rotMat is refreshed at the top of Car::_PostTickUpdate, which a demoed car skips (Car.cpp:133-138)
	basis_to_rotmat(c.isDemoed ? A.demoRotMat[ci] : c.rb.basis, o.rotMat);
	o.vel = to_vec(c.rb.vel * BT_TO_UU);
	o.angVel = to_vec(c.rb.angvel);
	o.tickCountSinceUpdate = c.tickCountSinceUpdate;
	o.isOnGround = c.isOnGround;
	for (int i = 0; i < 4; i++) o.wheelsWithContact[i] = c.wheelsWithContact[i];
	o.hasJumped = c.hasJumped; o.hasDoubleJumped = c.hasDoubleJumped; o.hasFlipped = c.hasFlipped;
	o.flipRelTorque = to_vec(c.flipRelTorque);
	o.jumpTime = c.jumpTime; o.flipTime = c.flipTime; o.isFlipping = c.isFlipping; o.isJumping = c.isJumping;
	o.airTime = c.airTime; o.airTimeSinceJump = c.airTimeSinceJump; o.boost = c.boost; o.timeSinceBoosted = c.timeSinceBoosted;
	o.isBoosting = c.isBoosting; o.boostingTime = c.boostingTime; o.isSupersonic = c.isSupersonic; o.supersonicTime = c.supersonicTime;
	o.handbrakeVal = c.handbrakeVal; o.isAutoFlipping = c.isAutoFlipping; o.autoFlipTimer = c.autoFlipTimer; o.autoFlipTorqueScale = c.autoFlipTorqueScale;
	o.worldContact.hasContact = c.worldContactHas; o.worldContact.contactNormal = to_vec(c.worldContactNormal);
	o.carContact.otherCarID = c.carContactOtherID; o.carContact.cooldownTimer = c.carContactCooldown;
	o.isDemoed = c.isDemoed; o.demoRespawnTimer = c.demoRespawnTimer;
	o.ballHitInfo.isValid = c.bhValid; o.ballHitInfo.relativePosOnBall = to_vec(c.bhRelPos); o.ballHitInfo.ballPos = to_vec(c.bhBallPos);
	o.ballHitInfo.extraHitVel = to_vec(c.bhExtraVel); o.ballHitInfo.tickCountWhenHit = c.bhTickHit; o.ballHitInfo.tickCountWhenExtraImpulseApplied = c.bhTickExtra;
	ctrl_to_api(c.lastControls, o.lastControls);
}
// Car::SetState (Car.cpp:23-36)
RSB_HD void car_from_api(const RsbCarState& s, ArenaS& A, int ci) {
	CarS& c = A.cars[ci];
	c.rb.pos = from_vec(s.pos) * UU_TO_BT;
	c.rb.basis = rotmat_to_basis(s.rotMat);
	A.demoRotMat[ci] = c.rb.basis;  // _internalState = state
	c.rb.vel = from_vec(s.vel) * UU_TO_BT;
	c.rb.angvel = from_vec(s.angVel);
	c.velImpulseCache = v3zero();
	c.tickCountSinceUpdate = 0;
	c.isOnGround = s.isOnGround != 0;
	for (int i = 0; i < 4; i++) c.wheelsWithContact[i] = s.wheelsWithContact[i] != 0;
	c.hasJumped = s.hasJumped != 0; c.hasDoubleJumped = s.hasDoubleJumped != 0; c.hasFlipped = s.hasFlipped != 0;
	c.flipRelTorque = from_vec(s.flipRelTorque);
	c.jumpTime = s.jumpTime; c.flipTime = s.flipTime; c.isFlipping = s.isFlipping != 0; c.isJumping = s.isJumping != 0;
	c.airTime = s.airTime; c.airTimeSinceJump = s.airTimeSinceJump; c.boost = s.boost; c.timeSinceBoosted = s.timeSinceBoosted;
	c.isBoosting = s.isBoosting != 0; c.boostingTime = s.boostingTime; c.isSupersonic = s.isSupersonic != 0; c.supersonicTime = s.supersonicTime;
	c.handbrakeVal = s.handbrakeVal; c.isAutoFlipping = s.isAutoFlipping != 0; c.autoFlipTimer = s.autoFlipTimer; c.autoFlipTorqueScale = s.autoFlipTorqueScale;
	c.worldContactHas = s.worldContact.hasContact != 0; c.worldContactNormal = from_vec(s.worldContact.contactNormal);
	c.carContactOtherID = s.carContact.otherCarID; c.carContactCooldown = s.carContact.cooldownTimer;
	c.isDemoed = s.isDemoed != 0; c.demoRespawnTimer = s.demoRespawnTimer;
	c.bhValid = s.ballHitInfo.isValid != 0; c.bhRelPos = from_vec(s.ballHitInfo.relativePosOnBall); c.bhBallPos = from_vec(s.ballHitInfo.ballPos);
	c.bhExtraVel = from_vec(s.ballHitInfo.extraHitVel); c.bhTickHit = s.ballHitInfo.tickCountWhenHit; c.bhTickExtra = s.ballHitInfo.tickCountWhenExtraImpulseApplied;
	ctrl_from_api(s.lastControls, c.lastControls);
}
// Ball::GetState / SetState (Ball.cpp:27-51)
RSB_HD void ball_to_api(const BallS& b, RsbBallState& o) {
	unsigned char* raw = (unsigned char*)&o;
	for (size_t i = 0; i < sizeof(RsbBallState); i++) raw[i] = 0;
	o.pos = to_vec(b.rb.pos * BT_TO_UU);
	basis_to_rotmat(b.rb.basis, o.rotMat);
	o.vel = to_vec(b.rb.vel * BT_TO_UU);
	o.angVel = to_vec(b.rb.angvel);
	o.tickCountSinceUpdate = b.tickCountSinceUpdate;
	o.hsInfo.yTargetDir = b.hsYTargetDir; o.hsInfo.curTargetSpeed = b.hsCurTargetSpeed; o.hsInfo.timeSinceHit = b.hsTimeSinceHit;
	o.dsInfo.chargeLevel = 1;
}
RSB_HD void ball_from_api(const RsbBallState& s, BallS& b) {
	b.rb.pos = from_vec(s.pos) * UU_TO_BT;
	b.rb.basis = rotmat_to_basis(s.rotMat);
	b.rb.vel = from_vec(s.vel) * UU_TO_BT;
	b.rb.angvel = from_vec(s.angVel);
	b.hsYTargetDir = s.hsInfo.yTargetDir; b.hsCurTargetSpeed = s.hsInfo.curTargetSpeed; b.hsTimeSinceHit = s.hsInfo.timeSinceHit;
	b.velImpulseCache = v3zero();
	b.tickCountSinceUpdate = 0;
}

// BallState::dsInfo (Ball.h:34-47) lives in the cold tail of the arena record (Dropshot only)
RSB_HD void ds_to_api(const ArenaS& A, RsbBallState& o) {
	o.dsInfo.chargeLevel = A.dsChargeLevel; o.dsInfo.accumulatedHitForce = A.dsAccumulatedHitForce; o.dsInfo.yTargetDir = A.dsYTargetDir;
	o.dsInfo.hasDamaged = A.dsHasDamaged ? 1 : 0; o.dsInfo.lastDamageTick = A.dsLastDamageTick;
}
RSB_HD void ds_from_api(const RsbBallState& s, ArenaS& A) {
	A.dsChargeLevel = s.dsInfo.chargeLevel; A.dsAccumulatedHitForce = s.dsInfo.accumulatedHitForce; A.dsYTargetDir = s.dsInfo.yTargetDir;
	A.dsHasDamaged = s.dsInfo.hasDamaged ? 1u : 0u; A.dsLastDamageTick = s.dsInfo.lastDamageTick;
}

// ------------------------------------------------------------------ kernels
#ifndef RSB_HOSTCHECK
#define KERNEL_1D(i, n) int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= (n)) return;
#define LAUNCH_1D(kern, n, b, ...) do { int _ln = (n); if (_ln > 0) { kern<<<(_ln + 127) / 128, 128, 0, (b)->stream>>>(__VA_ARGS__); CUDA_OK(cudaGetLastError()); } } while (0)
#else
static thread_local int g_hostIdx = 0;
#define KERNEL_1D(i, n) int i = g_hostIdx; if (i >= (n)) return;
#define LAUNCH_1D(kern, n, b, ...) do { int _ln = (n); for (g_hostIdx = 0; g_hostIdx < _ln; g_hostIdx++) kern(__VA_ARGS__); } while (0)
#endif

// ---- the hot kernel -------------------------------------------------------------------------
// Block-synchronous phase execution: a block keeps a CHUNK of arenas (their ArenaS records plus the
// batch constants and the scene header) resident in shared memory for all n_ticks ticks of the
// launch, deals the items of each tick phase (rsb_tick.h) over its threads and separates phases
// with __syncthreads().  All warps of an SM therefore run the same phase code at the same time
// (one instruction-cache working set), suspension rays / car logic / body integration use one
// thread per wheel / car / body, and HBM sees each arena record once per launch in and once out.
#ifndef RSB_STEP_THREADS
#define RSB_STEP_THREADS 448  // 14 warps, 128 registers: the 448 wheel rays of 28 resident 2v2 arenas take one round (A/B on B200: +6..9 % over 256 threads, 512 threads / 128 registers lose it again to spills)
#endif
constexpr int STEP_THREADS = RSB_STEP_THREADS;
// resident blocks per SM: each block keeps its own chunk of arenas; blocks drift through the tick phases independently
#ifndef RSB_BLOCKS_PER_SM
#define RSB_BLOCKS_PER_SM 1
#endif
constexpr int BLOCKS_PER_SM = RSB_BLOCKS_PER_SM;
static_assert(offsetof(ArenaS, cars) % 8 == 0 && sizeof(CarS) % 8 == 0 && sizeof(BodyW) % 8 == 0 && sizeof(WheelW) % 8 == 0 && sizeof(Manifold) % 8 == 0, "records are staged in 8-byte words");
constexpr size_t SMEM_P_BYTES = (sizeof(BatchParams) + 15) & ~(size_t)15;
constexpr size_t SMEM_SC_BYTES = (sizeof(Scene) + 15) & ~(size_t)15;
// Scheduling of the tick phases inside a block (see arena_step_kernel):
//   0 (default)  block-synchronous: all warps run one phase at a time, items dealt over the whole block, __syncthreads() in
//                between.  Every warp of the SM is in the same few KB of code at the same time.
//   1            warp-autonomous (developer A/B, -DRSB_WARP_SCHED=1): every warp owns a fixed subset of the block's arenas and
//                runs the whole tick for them on its own, phases separated by __syncwarp() only, so that no warp waits for the
//                slowest arena of the block.  MEASURED ON B200: 4.7x SLOWER (4.4 M vs 20.8 M arena-ticks/s, bit-identical
//                results) -- 14 warps in 14 different places of a 1.2 MB kernel thrash the 32 KB instruction cache.  The tick
//                is instruction-fetch bound as soon as the warps of an SM stop sharing their instruction stream; this is what
//                keeps the block-synchronous design (and its barrier stalls) in place.
#ifndef RSB_WARP_SCHED
#define RSB_WARP_SCHED 0
#endif
constexpr int WARP_QUEUE_CAP = 96;  // detector queue entries per warp (warp-autonomous scheduling)
constexpr size_t SMEM_Q_BYTES = RSB_WARP_SCHED ? (size_t)(RSB_STEP_THREADS / 32) * (16 + WARP_QUEUE_CAP * 2)
                                               : (size_t)(16 + (512 / RSB_BLOCKS_PER_SM) * 2);  // detector queue(s): counter + entries
constexpr size_t SMEM_MAX = (228 * 1024) / BLOCKS_PER_SM - 1024;  // 228 KB per SM, 1 KB reserved per resident block

// Shared-memory layout of one resident arena: [ArenaS prefix (header, ball, pads, nCars cars) | BodyW x nb |
// WheelW x 4 nCars | padLocked x nPads], padded so that neighbouring arenas start in different banks.
struct ArenaLayout {
	int sBytes, bodyOff, wheelOff, wheelBytes, solvBytes, padOff, miscOff, pairOff, mfOff, stride;
	RSB_HD static ArenaLayout make(int nCars, int nPads) {
		ArenaLayout L;
		L.sBytes = (int)offsetof(ArenaS, cars) + nCars * (int)sizeof(CarS);
		L.bodyOff = L.sBytes;
		L.wheelOff = L.bodyOff + (nCars + 1) * (int)sizeof(BodyW);
		// the wheel scratch and the solver scratch share one region (the wheels are dead after the car pre-tick)
		L.wheelBytes = nCars * 4 * (int)sizeof(WheelW);
		L.solvBytes = (solv_bytes(nCars + 1) + 7) & ~7;
		L.padOff = L.wheelOff + (L.wheelBytes > L.solvBytes ? L.wheelBytes : L.solvBytes);
		L.miscOff = L.padOff + ((nPads * 4 + 7) & ~7);
		L.pairOff = L.miscOff + 16;
		L.mfOff = L.pairOff + (((int)sizeof(PairList) + 7) & ~7);
		int end = L.mfOff + MAX_MANIFOLDS * (int)sizeof(Manifold);
		L.stride = end + 8;  // 8-byte aligned, and (stride / 4) odd multiples de-phase the banks
		if (((L.stride / 4) & 3) == 0) L.stride += 8;
		return L;
	}
};

struct StepArgs {
	ArenaS* S; ArenaH* H; ArenaW* W; EpaScratch* E; const BatchParams* P; const Scene* sc; EventSink ev;
	int nArenas, nTicks; const CtrlS* ctrlSeq; int nCtrl; int apb;
	int* chunkNext;  // device counter (starts at 0): a block that has finished its chunk fetches chunk gridDim.x + counter++
	int qcap;    // usable entries of the block's detector queue (DETECT_QUEUE_CAP; smaller only through rsb_debug_set_detect_queue_cap)
	int mirror;  // keep the full debug mirror (ArenaH) up to date: single-tick launches only
	RsbBallState* ballTrace;  // optional [nTicks + 1][nArenas]: Ball::GetState() after every tick (row t + 1), see rsb_predict_ball
	ArenaLayout lay;
};

#ifndef RSB_DETECT_QUEUE_CAP
#define RSB_DETECT_QUEUE_CAP (512 / RSB_BLOCKS_PER_SM)  // developer builds shrink it to exercise the overflow sweep
#endif
constexpr int DETECT_QUEUE_CAP = RSB_DETECT_QUEUE_CAP;
static_assert(DETECT_QUEUE_CAP >= 1 && DETECT_QUEUE_CAP <= 512 / RSB_BLOCKS_PER_SM, "the detector queue's shared-memory slice holds 512 / blocks-per-SM entries");

#ifndef RSB_HOSTCHECK
#if defined(RSB_PROFILE)
// developer builds: per phase [0..15] block wall cycles, [16..31] sum of busy-thread cycles, [32..47] busy-thread count,
// [48..63] sum over blocks of the slowest thread's cycles
__device__ unsigned long long g_prof[64];
__device__ unsigned long long g_blockCycles[1024];
__device__ unsigned long long g_blockPhase[256][16];  // per block: wall cycles per phase, accumulated over the launch  // elapsed cycles of each block of the last launch (imbalance between SMs)

// [phase NS=0 / DYN=1][bucket: min(7, detector calls of the slowest item) + 8 if it ran EPA][0 = sum of its cycles, 1 = count]
__device__ unsigned long long g_prof2[2][16][2];
#define PHASE_PROF_BEGIN() long long _pt0 = clock64()
#define PHASE_PROF_ITEMS(i, had) do { if (had) { unsigned long long _d = (unsigned long long)(clock64() - _pt0); atomicAdd(&g_prof[16 + i], _d); atomicAdd(&g_prof[32 + i], 1ull); atomicMax(&s_profMax, _d); } } while (0)
#define PHASE_PROF_END(i) do { if (threadIdx.x == 0) { atomicAdd(&g_prof[i], (unsigned long long)(clock64() - _pt0)); if (blockIdx.x < 256) g_blockPhase[blockIdx.x][i] += (unsigned long long)(clock64() - _pt0); atomicAdd(&g_prof[48 + i], s_profMax); s_profMax = 0; \
	if (i == PH_NARROW_DETECT || i == PH_DYN_SOLVE) { int c = (int)(s_profMax2 & 0xfffff); int bk = min(7, c & 255) + ((c >> 8) ? 8 : 0); int ph = (i == PH_DYN_SOLVE); atomicAdd(&g_prof2[ph][bk][0], s_profMax2 >> 20); atomicAdd(&g_prof2[ph][bk][1], 1ull); s_profMax2 = 0; } } __syncthreads(); } while (0)
#else
#define PHASE_PROF_BEGIN() ((void)0)
#define PHASE_PROF_ITEMS(i, had) ((void)0)
#define PHASE_PROF_END(i) ((void)0)
#endif

// MODE >= 0: the kernel instantiation of one game mode -- the code-path selectors of the context are compile-time constants.
// MODE < 0: taken from the batch at run time (developer builds with -DRSB_ONE_KERNEL compile only this one).
template <int MODE>
__device__ __forceinline__ Ctx make_ctx(const StepArgs& a, const BatchParams* sP, const Scene* sSc, unsigned char* sS, int a0, int ar) {
	extern __shared__ __align__(16) unsigned char smem[];
	Ctx cx;
	if (MODE >= 0) { cx.mode = MODE; cx.ballIsSphere = (MODE != SNOWDAY) ? 1 : 0; cx.nTiles = (MODE == DROPSHOT) ? MAX_TILES : 0; }
	else { cx.mode = sP->mode; cx.ballIsSphere = sP->ballIsSphere; cx.nTiles = sSc->nTiles; }
#if RSB_WARP_SCHED
	cx.nq = (unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES + (size_t)(threadIdx.x >> 5) * (16 + WARP_QUEUE_CAP * 2)); cx.q = (uint16_t*)(cx.nq + 4); cx.qcap = WARP_QUEUE_CAP; cx.arLocal = ar;
#else
	cx.nq = (unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES); cx.q = (uint16_t*)(cx.nq + 4); cx.qcap = a.qcap; cx.arLocal = ar;
#endif
	unsigned char* base = sS + (size_t)ar * a.lay.stride;
	cx.s = (ArenaS*)base; cx.g = a.S + (a0 + ar); cx.body = (BodyW*)(base + a.lay.bodyOff); cx.wheels = (WheelW*)(base + a.lay.wheelOff);
	cx.padLocked = (int*)(base + a.lay.padOff); cx.misc = (int*)(base + a.lay.miscOff);
	cx.pl = (PairList*)(base + a.lay.pairOff); cx.mf = (Manifold*)(base + a.lay.mfOff); cx.solv = base + a.lay.wheelOff;
	cx.w = a.W + (a0 + ar); cx.epa = a.E + ((size_t)blockIdx.x * STEP_THREADS + threadIdx.x); cx.p = sP; cx.sc = sSc; cx.ev = a.ev; cx.arena = a0 + ar;
	return cx;
}

// Item -> thread assignment of a phase.  Items of a phase are independent, so any assignment gives the same results; the
// assignment decides how much a warp diverges:
//   * arena-level phases (few, long, data-dependent items): consecutive items go to different warps (ptid permutation),
//     so a narrow phase spreads over all schedulers and at most 4 unrelated items share a warp;
//   * wheel rays: natural order -- the four wheels of a car and the cars of an arena sit in neighbouring lanes and mostly
//     take the same path (same broadphase cell, same static list);
//   * body phases: all balls in the first warp(s), cars dealt over the other warps -- the sphere and the box code paths
//     never share a warp;
//   * car logic: cars first, dealt evenly over all warps in low lanes; pad and ball items fill the remaining lanes/rounds.
// i / d for the small numbers of the item maps (i < 65,536): one multiply-high with m = div_magic(d) instead of a 32-bit division
// (~30 instructions), paid twice per phase and tick by every thread -- 5.6 % of all warp instructions in the round-2 profile.
struct ItemDiv { unsigned mNA, mPer; };
__device__ __forceinline__ unsigned div_magic(int d) { return d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; }
__device__ __forceinline__ int fast_div(int i, int d, unsigned m) { return d > 1 ? (int)__umulhi((unsigned)i, m) : i; }

template <int PH>
__device__ __forceinline__ bool map_item(int k, int tid, int ptid, int nA, const BatchParams& P, const ItemDiv dv, int& ar, int& item) {
	constexpr int NW = STEP_THREADS / 32;
	const int nCars = P.nCars;
	const int warp = tid >> 5, lane = tid & 31;
#ifndef RSB_MAP
#define RSB_MAP 1  // measured: natural order helps the wheel rays (-13 % phase time); the body / car groupings do not
#endif
	if (PH == PH_WHEELS && (RSB_MAP & 1)) {
		const int i = tid + k * STEP_THREADS, per = nCars * 4;
		if (i >= nA * per) return false;
		ar = fast_div(i, per, dv.mPer); item = i - ar * per;
		return true;
	}
	if ((PH == PH_INIT || PH == PH_PREDICT || PH == PH_NARROW_STATIC || PH == PH_FINISH_BODY) && (RSB_MAP & 2)) {
		const int ballWarps = (nA + 31) / 32;
		if (ballWarps < NW) {
			if (warp < ballWarps) {
				if (k > 0) return false;
				ar = warp * 32 + lane; item = 0;
				return ar < nA;
			}
			const int carWarps = NW - ballWarps;
			const int c = (warp - ballWarps) + carWarps * (lane + 32 * k);
			if (c >= nA * nCars) return false;
			ar = c / nCars; item = 1 + (c - ar * nCars);
			return true;
		}
	}
	if ((PH == PH_CAR_LOGIC || PH == PH_CAR_POST) && (RSB_MAP & 4)) {
		const int nCarItems = nA * nCars;
		const int L0 = (nCarItems + NW - 1) / NW;  // lanes per warp that hold a car in round 0
		if (L0 <= 32) {
			if (k == 0 && lane < L0) {
				const int c = lane * NW + warp;
				if (c >= nCarItems) return false;
				ar = c / nCars; item = c - ar * nCars;
				return true;
			}
			if (PH == PH_CAR_POST) return false;
			const int others = (phase_items(P, PH) - nCars) * nA;
			const int slots0 = (32 - L0) * NW;
			int j;
			if (k == 0) j = (lane - L0) * NW + warp;
			else j = slots0 + (k - 1) * STEP_THREADS + tid;
			if (j >= others) return false;
			ar = j % nA; item = nCars + j / nA;
			return true;
		}
	}
#ifndef RSB_NARROW_NATURAL
#define RSB_NARROW_NATURAL 0  // developer A/B: deal the default-mapped phases in thread order (few warps, dense lanes) instead of across warps
#endif
	const int i = (RSB_NARROW_NATURAL ? tid : ptid) + k * STEP_THREADS;
	if (i >= phase_items(P, PH) * nA) return false;
	item = fast_div(i, nA, dv.mNA); ar = i - item * nA;
	return true;
}

template <int PH>
__device__ __forceinline__ bool phase_thread_busy(int tid, int ptid, int nA, const BatchParams& P, const ItemDiv dv) { int ar, item; return map_item<PH>(0, tid, ptid, nA, P, dv, ar, item); }

template <int MODE, int PH, int SLOT = PH>
__device__ __forceinline__ void run_phase_block(const StepArgs& a, const BatchParams* sP, const Scene* sSc, unsigned char* sS, int a0, int nA, int ptid, const ItemDiv dv) {
#if defined(RSB_PROFILE)
	__shared__ unsigned long long s_profMax;
#endif
	PHASE_PROF_BEGIN();
	if (PH == PH_NARROW_DETECT) {
		// the block's detector queue: entry i goes to thread ptid == i, i.e. consecutive entries to different warps -- a GJK / EPA
		// call is a long data-dependent loop, two of them in one warp would run one after the other
		extern __shared__ __align__(16) unsigned char smem[];
		const unsigned* nqp = (const unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES);
		const uint16_t* q = (const uint16_t*)(nqp + 4);
		const int pushed = (int)*nqp;
		const int nq = min(pushed, a.qcap);
		// pass 0: the queued records.  pass 1 (only after a queue overflow, block-uniform): every record of every resident arena
		// that is still pending -- the ones that did not fit the queue.  One detector call site for both.
		const int nPass = pushed > a.qcap ? 2 : 1;
		for (int pass = 0; pass < nPass; pass++) {
			const int n = pass == 0 ? nq : nA * MAX_CANDS;
			for (int i = ptid; i < n; i += STEP_THREADS) {
				int arl, idx;
				if (pass == 0) { const int e = q[i]; arl = e >> 7; idx = e & 127; }
				else { arl = i / MAX_CANDS; idx = i - arl * MAX_CANDS; }
				const Ctx cx = make_ctx<MODE>(a, sP, sSc, sS, a0, arl);
				if (pass == 1 && !(idx < cx.misc[3] && cx.w->cand[idx].state == CAND_PENDING)) continue;
				phase_run(cx, PH, idx);
			}
			if (pass + 1 < nPass) __syncthreads();
		}
		PHASE_PROF_ITEMS(SLOT, ptid < nq);
	} else {
	const int n = phase_items(*sP, PH) * nA;
	const int nRounds = (n + STEP_THREADS - 1) / STEP_THREADS + ((RSB_MAP & 6) ? 1 : 0);  // the grouped developer maps leave holes: one more round
	for (int k = 0; k < nRounds; k++) {
		int ar, item;
		if (!map_item<PH>(k, threadIdx.x, ptid, nA, *sP, dv, ar, item)) continue;
		const Ctx cx = make_ctx<MODE>(a, sP, sSc, sS, a0, ar);
		phase_run(cx, PH, item);
	}
	PHASE_PROF_ITEMS(SLOT, phase_thread_busy<PH>(threadIdx.x, ptid, nA, *sP, dv));
	}
#if defined(RSB_PROFILE)
	__shared__ unsigned long long s_profMax2;
	if (PH == PH_NARROW_DETECT || PH == PH_DYN_SOLVE) {
		EpaScratch* es = a.E + ((size_t)blockIdx.x * STEP_THREADS + threadIdx.x);
		if (phase_thread_busy<PH>(threadIdx.x, ptid, nA, *sP, dv)) { unsigned long long d = (unsigned long long)(clock64() - _pt0); int c = es->profCalls; atomicMax(&s_profMax2, (d << 20) | (unsigned long long)(c & 0xfffff)); }
		es->profCalls = 0;
	}
#endif
	__syncthreads();
	if (a.mirror && (PH == PH_CAR_LOGIC || PH == PH_DYN_SOLVE)) {
		// single-tick launches keep the debug mirror complete: the wheel scratch and the solver scratch alias each other
		const int words = (PH == PH_CAR_LOGIC ? a.lay.wheelBytes : a.lay.solvBytes) / 4;
		for (int i = threadIdx.x; i < nA * words; i += STEP_THREADS) {
			const int ar = i / words, w = i - ar * words;
			const uint32_t v = ((const uint32_t*)(sS + (size_t)ar * a.lay.stride + a.lay.wheelOff))[w];
			ArenaH* dst = a.H + a0 + ar;
			if (PH == PH_CAR_LOGIC) ((uint32_t*)dst->wheels)[w] = v; else ((uint32_t*)dst->solv)[w] = v;
		}
		__syncthreads();
	}
	PHASE_PROF_END(SLOT);
}


#if RSB_WARP_SCHED
// One phase of the tick for the arenas owned by the calling warp (block-local indices warp, warp + NW, ...): its items go to the
// warp's lanes in rounds of 32, and the warp alone synchronises.  Items of a phase only touch their own arena, so nothing here
// depends on what the other warps of the block are doing.
template <int MODE, int PH>
__device__ __forceinline__ void run_phase_warp(const StepArgs& a, const BatchParams* sP, const Scene* sSc, unsigned char* sS, int a0, int nA, int warp, int lane) {
	constexpr int NW = STEP_THREADS / 32;
	const int nMine = warp < nA ? (nA - warp + NW - 1) / NW : 0;
	if (PH == PH_NARROW_DETECT) {
		extern __shared__ __align__(16) unsigned char smem[];
		const unsigned* nqp = (const unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES + (size_t)warp * (16 + WARP_QUEUE_CAP * 2));
		const uint16_t* q = (const uint16_t*)(nqp + 4);
		const int pushed = (int)*nqp;
		const int nq = min(pushed, WARP_QUEUE_CAP);
		const int nPass = pushed > WARP_QUEUE_CAP ? 2 : 1;  // pass 1: the records that did not fit the warp's queue (see run_phase_block)
		for (int pass = 0; pass < nPass; pass++) {
			const int n = pass == 0 ? nq : nMine * MAX_CANDS;
			for (int i = lane; i < n; i += 32) {
				int arl, idx;
				if (pass == 0) { const int e = q[i]; arl = e >> 7; idx = e & 127; }
				else { const int k = i / MAX_CANDS; arl = warp + k * NW; idx = i - k * MAX_CANDS; }
				const Ctx cx = make_ctx<MODE>(a, sP, sSc, sS, a0, arl);
				if (pass == 1 && !(idx < cx.misc[3] && cx.w->cand[idx].state == CAND_PENDING)) continue;
				phase_run(cx, PH, idx);
			}
			__syncwarp();
		}
	} else {
		const int per = phase_items(*sP, PH), n = per * nMine;
		for (int i = lane; i < n; i += 32) {
			int k, item;
			if (PH == PH_WHEELS) { k = i / per; item = i - k * per; }  // the wheels of a car sit in neighbouring lanes
			else { item = i / nMine; k = i - item * nMine; }           // same item kind in neighbouring lanes
			const Ctx cx = make_ctx<MODE>(a, sP, sSc, sS, a0, warp + k * NW);
			phase_run(cx, PH, item);
		}
	}
	__syncwarp();
	if (a.mirror && (PH == PH_CAR_LOGIC || PH == PH_DYN_SOLVE)) {
		// single-tick launches keep the debug mirror complete: the wheel scratch and the solver scratch alias each other
		const int words = (PH == PH_CAR_LOGIC ? a.lay.wheelBytes : a.lay.solvBytes) / 4;
		for (int k = 0; k < nMine; k++) {
			const int ar = warp + k * NW;
			ArenaH* dst = a.H + a0 + ar;
			for (int w = lane; w < words; w += 32) {
				const uint32_t v = ((const uint32_t*)(sS + (size_t)ar * a.lay.stride + a.lay.wheelOff))[w];
				if (PH == PH_CAR_LOGIC) ((uint32_t*)dst->wheels)[w] = v; else ((uint32_t*)dst->solv)[w] = v;
			}
		}
		__syncwarp();
	}
}
#endif

// ptxas derives 128 registers from __launch_bounds__(448); the register file holds 144 per thread at this block size
template <int MODE>
#if defined(RSB_MAXNREG)
__global__ void __maxnreg__(RSB_MAXNREG) arena_step_kernel(const StepArgs a) {
#else
__global__ void __launch_bounds__(STEP_THREADS, BLOCKS_PER_SM) arena_step_kernel(const StepArgs a) {
#endif
	extern __shared__ __align__(16) unsigned char smem[];
	BatchParams* sP = (BatchParams*)smem;
	Scene* sSc = (Scene*)(smem + SMEM_P_BYTES);
	unsigned char* sS = smem + SMEM_P_BYTES + SMEM_SC_BYTES + SMEM_Q_BYTES;
	const int tid = threadIdx.x;
#if defined(RSB_PROFILE)
	const long long _kt0 = clock64();
#endif
	for (int i = tid; i < (int)(sizeof(BatchParams) / 4); i += STEP_THREADS) ((uint32_t*)sP)[i] = ((const uint32_t*)a.P)[i];
	for (int i = tid; i < (int)(sizeof(Scene) / 4); i += STEP_THREADS) ((uint32_t*)sSc)[i] = ((const uint32_t*)a.sc)[i];
	// item -> thread permutation: consecutive items go to different warps, so that a phase with few
	// items (one per arena) spreads over all four schedulers instead of serialising in one warp
	const int ptid = (tid & 31) * (STEP_THREADS / 32) + (tid >> 5);
	const int nCars = a.P->nCars, nPads = a.P->nPads, nb = nCars + 1;
	const int sWords = a.lay.sBytes / 8, strideWords = a.lay.stride / 8;
	const int nChunks = (a.nArenas + a.apb - 1) / a.apb;
	// A block starts with chunk blockIdx.x; when the batch has more chunks than resident blocks the remaining ones are handed out
	// first come, first served: chunks differ in cost (contact-heavy games), a fixed round-robin would wait for the unluckiest block.
	__shared__ int s_nextChunk;
	for (int chunk = blockIdx.x; chunk < nChunks;) {
		const int a0 = chunk * a.apb;
		const int nA = min(a.apb, a.nArenas - a0);
		ItemDiv dv; dv.mNA = div_magic(nA); dv.mPer = div_magic(nCars * 4);
		__syncthreads();
		// stage the records: persistent prefix from HBM, per-tick scratch zeroed
		for (int i = tid; i < nA * strideWords; i += STEP_THREADS) {
			const int ar = i / strideWords, w = i - ar * strideWords;
			uint2 v = make_uint2(0u, 0u);
			if (w < sWords) v = ((const uint2*)(a.S + a0 + ar))[w];
			((uint2*)(sS + (size_t)ar * a.lay.stride))[w] = v;
		}
		__syncthreads();
#if RSB_WARP_SCHED
		{
			constexpr int NW = STEP_THREADS / 32;
			const int warp = tid >> 5, lane = tid & 31;
			const int nMine = warp < nA ? (nA - warp + NW - 1) / NW : 0;
			unsigned* myq = (unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES + (size_t)warp * (16 + WARP_QUEUE_CAP * 2));
			for (int t = 0; t < a.nTicks; t++) {
				if (a.ctrlSeq) {
					const int ct = t < a.nCtrl ? t : a.nCtrl - 1;
					for (int i = lane; i < nMine * nCars; i += 32) {
						const int k = i / nCars, ci = i - k * nCars, ar = warp + k * NW;
						((ArenaS*)(sS + (size_t)ar * a.lay.stride))->cars[ci].controls = a.ctrlSeq[((size_t)ct * a.nArenas + a0 + ar) * nCars + ci];
					}
				}
				if (lane == 0) *myq = 0u;  // this warp's detector queue
				__syncwarp();
				run_phase_warp<MODE, PH_INIT>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_PRE_BEGIN>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_WHEELS>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_CAR_LOGIC>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_PREDICT>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_PAIRS>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_NARROW_STATIC>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_NARROW_DETECT>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_NARROW_COMMIT>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_DYN_SOLVE>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_FINISH_BODY>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_CAR_POST>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_PADS_POST>(a, sP, sSc, sS, a0, nA, warp, lane);
				run_phase_warp<MODE, PH_TAIL>(a, sP, sSc, sS, a0, nA, warp, lane);
				if (a.ballTrace && lane < nMine) {
					const int ar = warp + lane * NW;
					ball_to_api(((const ArenaS*)(sS + (size_t)ar * a.lay.stride))->ball, a.ballTrace[(size_t)(t + 1) * a.nArenas + a0 + ar]);
				}
			}
			__syncthreads();
		}
#else
		for (int t = 0; t < a.nTicks; t++) {
			if (a.ctrlSeq) {
				const int ct = t < a.nCtrl ? t : a.nCtrl - 1;
				const CtrlS* src = a.ctrlSeq + ((size_t)ct * a.nArenas + a0) * nCars;
				for (int i = tid; i < nA * nCars; i += STEP_THREADS) ((ArenaS*)(sS + (size_t)(i / nCars) * a.lay.stride))->cars[i % nCars].controls = src[i];
				__syncthreads();
			}
			run_phase_block<MODE, PH_INIT>(a, sP, sSc, sS, a0, nA, ptid, dv);
#if defined(RSB_EXP_TWICE)
			run_phase_block<MODE, PH_INIT, 15>(a, sP, sSc, sS, a0, nA, ptid, dv);
#endif
			run_phase_block<MODE, PH_PRE_BEGIN>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_WHEELS>(a, sP, sSc, sS, a0, nA, ptid, dv);
#if defined(RSB_EXP_TWICE)  // developer experiment: the wheel phase is idempotent; its second run measures the same code WARM in the instruction cache
			run_phase_block<MODE, PH_WHEELS, 14>(a, sP, sSc, sS, a0, nA, ptid, dv);
#endif
			run_phase_block<MODE, PH_CAR_LOGIC>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_PREDICT>(a, sP, sSc, sS, a0, nA, ptid, dv);
			if (tid == 0) *(unsigned*)(smem + SMEM_P_BYTES + SMEM_SC_BYTES) = 0u;  // detector queue; the next write to it is two barriers away
			run_phase_block<MODE, PH_PAIRS>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_NARROW_STATIC>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_NARROW_DETECT>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_NARROW_COMMIT>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_DYN_SOLVE>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_FINISH_BODY>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_CAR_POST>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_PADS_POST>(a, sP, sSc, sS, a0, nA, ptid, dv);
			run_phase_block<MODE, PH_TAIL>(a, sP, sSc, sS, a0, nA, ptid, dv);
			if (a.ballTrace)  // the ball is next written in PH_CAR_LOGIC, several barriers from here
				for (int i = tid; i < nA; i += STEP_THREADS) {
					RsbBallState& o = a.ballTrace[(size_t)(t + 1) * a.nArenas + a0 + i];
					ball_to_api(((const ArenaS*)(sS + (size_t)i * a.lay.stride))->ball, o);
					if (sP->mode == DROPSHOT) ds_to_api(a.S[a0 + i], o);
				}
		}
#endif
		// records back to HBM; the last tick's hot scratch goes to its global mirror (debug probes)
		for (int i = tid; i < nA * sWords; i += STEP_THREADS) {
			const int ar = i / sWords, w = i - ar * sWords;
			((uint2*)(a.S + a0 + ar))[w] = ((const uint2*)(sS + (size_t)ar * a.lay.stride))[w];
		}
		if (a.mirror) {
			const int bodyWords = nb * (int)sizeof(BodyW) / 4;
			const int tailWords = (a.lay.stride - 8 - a.lay.padOff) / 4;  // padLocked, misc, pairs, manifold headers
			const int hotWords = bodyWords + tailWords;
			for (int i = tid; i < nA * hotWords; i += STEP_THREADS) {
				const int ar = i / hotWords; int w = i - ar * hotWords;
				const unsigned char* base = sS + (size_t)ar * a.lay.stride;
				ArenaH* dst = a.H + a0 + ar;
				if (w < bodyWords) { ((uint32_t*)dst->body)[w] = ((const uint32_t*)(base + a.lay.bodyOff))[w]; continue; }
				w -= bodyWords;
				const int off = a.lay.padOff + w * 4;
				const uint32_t v = *(const uint32_t*)(base + off);
				if (off < a.lay.miscOff) { if (w < nPads) ((uint32_t*)dst->padLocked)[w] = v; }
				else if (off < a.lay.pairOff) ((uint32_t*)dst->misc)[(off - a.lay.miscOff) / 4] = v;
				else if (off < a.lay.mfOff) { if (off - a.lay.pairOff < (int)sizeof(PairList)) ((uint32_t*)&dst->pairs)[(off - a.lay.pairOff) / 4] = v; }
				else if (off - a.lay.mfOff < MAX_MANIFOLDS * (int)sizeof(Manifold)) ((uint32_t*)dst->mf)[(off - a.lay.mfOff) / 4] = v;
			}
		}
		if (tid == 0) s_nextChunk = (int)gridDim.x + atomicAdd(a.chunkNext, 1);
		__syncthreads();
		chunk = s_nextChunk;
	}
#if defined(RSB_PROFILE)
	if (tid == 0 && blockIdx.x < 1024) g_blockCycles[blockIdx.x] = (unsigned long long)(clock64() - _kt0);
#endif
}
// One instantiation per game mode (the product build); -DRSB_ONE_KERNEL (developer A/B builds: one seventh of the compile time)
// keeps a single kernel that reads the mode at run time.
typedef void (*StepKernel)(const StepArgs);
static StepKernel step_kernel_of(int mode) {
#if defined(RSB_ONE_KERNEL)
	(void)mode;
	return arena_step_kernel<-1>;
#else
	switch (mode) {
	case SOCCAR: return arena_step_kernel<SOCCAR>;
	case HOOPS: return arena_step_kernel<HOOPS>;
	case HEATSEEKER: return arena_step_kernel<HEATSEEKER>;
	case SNOWDAY: return arena_step_kernel<SNOWDAY>;
	case DROPSHOT: return arena_step_kernel<DROPSHOT>;
	default: return arena_step_kernel<THE_VOID>;
	}
#endif
}
#else
// host check build (tests only): the same phases, one thread, state in place
static void arena_step_host(const StepArgs& a) {
	const int nCars = a.P->nCars;
	for (int arena = 0; arena < a.nArenas; arena++) {
		uint16_t hq[512]; unsigned hnq = 0;
		Ctx cx; cx.q = hq; cx.nq = &hnq; cx.qcap = a.qcap; cx.arLocal = 0; ctx_set_mode(cx, *a.P, a.sc);
		cx.s = a.S + arena; cx.g = cx.s; cx.w = a.W + arena; cx.epa = a.E; cx.p = a.P; cx.sc = a.sc; cx.ev = a.ev; cx.arena = arena;
		cx.body = a.H[arena].body; cx.wheels = a.H[arena].wheels; cx.padLocked = a.H[arena].padLocked; cx.misc = a.H[arena].misc;
		cx.pl = &a.H[arena].pairs; cx.mf = a.H[arena].mf; cx.solv = a.H[arena].solv;
		for (int t = 0; t < a.nTicks; t++) {
			if (a.ctrlSeq) {
				const int ct = t < a.nCtrl ? t : a.nCtrl - 1;
				const CtrlS* src = a.ctrlSeq + ((size_t)ct * a.nArenas + arena) * nCars;
				for (int ci = 0; ci < nCars; ci++) cx.s->cars[ci].controls = src[ci];
			}
			arena_tick(cx);
			if (a.ballTrace) { RsbBallState& o = a.ballTrace[(size_t)(t + 1) * a.nArenas + arena]; ball_to_api(cx.s->ball, o); if (a.P->mode == DROPSHOT) ds_to_api(*cx.s, o); }
		}
	}
}
#endif

RSB_GLOBAL void set_controls_kernel(ArenaS* S, const RsbCarControls* in, int nArenas, int nCars) {
	KERNEL_1D(i, nArenas * nCars);
	ctrl_from_api(in[i], S[i / nCars].cars[i % nCars].controls);
}
RSB_GLOBAL void get_car_states_kernel(const ArenaS* S, RsbCarState* out, int nArenas, int nCars) {
	KERNEL_1D(i, nArenas * nCars);
	car_to_api(S[i / nCars], i % nCars, out[i]);
}
RSB_GLOBAL void set_car_states_kernel(ArenaS* S, const uint32_t* ai, const uint32_t* ci, const RsbCarState* in, int n) {
	KERNEL_1D(i, n);
	car_from_api(in[i], S[ai[i]], (int)ci[i]);
}
RSB_GLOBAL void get_ball_states_kernel(const ArenaS* S, RsbBallState* out, int nArenas, int dropshot) {
	KERNEL_1D(i, nArenas);
	ball_to_api(S[i].ball, out[i]);
	if (dropshot) ds_to_api(S[i], out[i]);
}
RSB_GLOBAL void set_ball_states_kernel(ArenaS* S, const uint32_t* ai, const RsbBallState* in, int n, int dropshot) {
	KERNEL_1D(i, n);
	ball_from_api(in[i], S[ai[i]].ball);
	if (dropshot) ds_from_api(in[i], S[ai[i]]);
}
// Arena::GetDropshotTilesState / SetDropshotTilesState (Arena.h:178-179, Arena.cpp:225-242): one damage state per tile,
// [team][70]; setting the states also updates the tile bodies' contact response, as the reference does.
RSB_GLOBAL void get_tiles_kernel(const ArenaS* S, uint8_t* out, int nArenas) {
	KERNEL_1D(i, nArenas * MAX_TILES);
	out[i] = S[i / MAX_TILES].tiles[i % MAX_TILES] & 3u;
}
RSB_GLOBAL void set_tiles_kernel(ArenaS* S, const uint32_t* ai, const uint8_t* in, int n) {
	KERNEL_1D(i, n * MAX_TILES);
	const uint8_t st = in[i] & 3u;
	S[ai[i / MAX_TILES]].tiles[i % MAX_TILES] = (uint8_t)(st | (st == 2u ? 4u : 0u));
}
RSB_GLOBAL void get_pad_states_kernel(const ArenaS* S, RsbBoostPadState* out, int nArenas, int nPads) {
	KERNEL_1D(i, nArenas * nPads);
	const PadS& p = S[i / nPads].pads[i % nPads];
	RsbBoostPadState o; o.isActive = p.isActive; o._pad[0] = o._pad[1] = o._pad[2] = 0; o.cooldown = p.cooldown; o.prevLockedCarID = p.prevLockedCarID;
	out[i] = o;
}
RSB_GLOBAL void set_pad_states_kernel(ArenaS* S, const uint32_t* ai, const RsbBoostPadState* in, int n, int nPads) {
	KERNEL_1D(i, n * nPads);
	PadS& p = S[ai[i / nPads]].pads[i % nPads];
	p.isActive = in[i].isActive != 0; p.cooldown = in[i].cooldown; p.prevLockedCarID = in[i].prevLockedCarID;
}
RSB_GLOBAL void get_misc_kernel(const ArenaS* S, const BatchParams* P, const Scene* sc, uint64_t* ticks, uint8_t* scored, uint32_t* errs, int nArenas) {
	KERNEL_1D(i, nArenas);
	if (ticks) ticks[i] = S[i].tickCount;
	if (errs) errs[i] = S[i].errorFlags;
	if (scored) { Ctx cx; cx.s = (ArenaS*)&S[i]; cx.w = nullptr; cx.epa = nullptr; cx.body = nullptr; cx.wheels = nullptr; cx.padLocked = nullptr; cx.misc = nullptr; cx.pl = nullptr; cx.mf = nullptr; cx.solv = nullptr; cx.p = P; cx.sc = sc; cx.arena = i; cx.ev.buf = nullptr; ctx_set_mode(cx, *P, sc); scored[i] = is_ball_scored(cx) ? 1 : 0; }
}
RSB_GLOBAL void get_car_internals_kernel(const ArenaS* S, RsbCarInternal* out, int nArenas, int nCars) {
	KERNEL_1D(i, nArenas * nCars);
	const CarS& c = S[i / nCars].cars[i % nCars];
	RsbCarInternal o;
	for (int k = 0; k < 4; k++) { o.steerAngle[k] = c.wSteer[k]; o.engineForce[k] = c.wEngine[k]; o.brake[k] = c.wBrake[k]; o.latFriction[k] = c.wLatFriction[k]; o.longFriction[k] = c.wLongFriction[k]; o.extraPushback[k] = c.wExtraPushback[k]; }
	o.velocityImpulseCache[0] = c.velImpulseCache.x; o.velocityImpulseCache[1] = c.velImpulseCache.y; o.velocityImpulseCache[2] = c.velImpulseCache.z; o.velocityImpulseCache[3] = 0.f;
	o.rbPos[0] = c.rb.pos.x; o.rbPos[1] = c.rb.pos.y; o.rbPos[2] = c.rb.pos.z; o.rbPos[3] = 0.f;
	o.rbVel[0] = c.rb.vel.x; o.rbVel[1] = c.rb.vel.y; o.rbVel[2] = c.rb.vel.z; o.rbVel[3] = 0.f;
	o.tickCountSinceUpdate = c.tickCountSinceUpdate; o.hasRigidBody = 3; o._pad = 0;
	for (int r = 0; r < 3; r++) { o.rbBasis[4 * r] = c.rb.basis.r[r].x; o.rbBasis[4 * r + 1] = c.rb.basis.r[r].y; o.rbBasis[4 * r + 2] = c.rb.basis.r[r].z; o.rbBasis[4 * r + 3] = 0.f; }
	out[i] = o;
}
RSB_GLOBAL void set_car_internals_kernel(ArenaS* S, const uint32_t* ai, const uint32_t* ci, const RsbCarInternal* in, int n) {
	KERNEL_1D(i, n);
	CarS& c = S[ai[i]].cars[ci[i]];
	const RsbCarInternal& o = in[i];
	for (int k = 0; k < 4; k++) { c.wSteer[k] = o.steerAngle[k]; c.wEngine[k] = o.engineForce[k]; c.wBrake[k] = o.brake[k]; c.wLatFriction[k] = o.latFriction[k]; c.wLongFriction[k] = o.longFriction[k]; c.wExtraPushback[k] = o.extraPushback[k]; }
	c.velImpulseCache = V3(o.velocityImpulseCache[0], o.velocityImpulseCache[1], o.velocityImpulseCache[2]);
	if (o.hasRigidBody & 1) {
		c.rb.pos = V3(o.rbPos[0], o.rbPos[1], o.rbPos[2]); c.rb.vel = V3(o.rbVel[0], o.rbVel[1], o.rbVel[2]);
		c.tickCountSinceUpdate = o.tickCountSinceUpdate;
	}
	// the body's own basis; CarState::rotMat (kept in demoRotMat while demoed) is not touched
	if (o.hasRigidBody & 2) for (int r = 0; r < 3; r++) c.rb.basis.r[r] = V3(o.rbBasis[4 * r], o.rbBasis[4 * r + 1], o.rbBasis[4 * r + 2]);
}
RSB_GLOBAL void get_ball_internals_kernel(const ArenaS* S, RsbBallInternal* out, int nArenas) {
	KERNEL_1D(i, nArenas);
	RsbBallInternal o; memset(&o, 0, sizeof(o));
	o.velocityImpulseCache[0] = S[i].ball.velImpulseCache.x; o.velocityImpulseCache[1] = S[i].ball.velImpulseCache.y; o.velocityImpulseCache[2] = S[i].ball.velImpulseCache.z;
	o.groundStickApplied = S[i].ball.groundStickApplied;
	const BallS& bb = S[i].ball;
	o.rbPos[0] = bb.rb.pos.x; o.rbPos[1] = bb.rb.pos.y; o.rbPos[2] = bb.rb.pos.z;
	o.rbVel[0] = bb.rb.vel.x; o.rbVel[1] = bb.rb.vel.y; o.rbVel[2] = bb.rb.vel.z;
	o.tickCountSinceUpdate = bb.tickCountSinceUpdate; o.hasRigidBody = 1;
	out[i] = o;
}
RSB_GLOBAL void set_ball_internals_kernel(ArenaS* S, const uint32_t* ai, const RsbBallInternal* in, int n) {
	KERNEL_1D(i, n);
	BallS& b = S[ai[i]].ball;
	b.velImpulseCache = V3(in[i].velocityImpulseCache[0], in[i].velocityImpulseCache[1], in[i].velocityImpulseCache[2]);
	b.groundStickApplied = in[i].groundStickApplied != 0;
	if (in[i].hasRigidBody) {
		b.rb.pos = V3(in[i].rbPos[0], in[i].rbPos[1], in[i].rbPos[2]); b.rb.vel = V3(in[i].rbVel[0], in[i].rbVel[1], in[i].rbVel[2]);
		b.tickCountSinceUpdate = in[i].tickCountSinceUpdate;
	}
}
RSB_GLOBAL void get_obs_kernel(const ArenaS* S, float* out, int nArenas, int nCars) {
	KERNEL_1D(i, nArenas);
	float* o = out + (size_t)i * (9 + 17 * nCars);
	const BallS& b = S[i].ball;
	V3 p = b.rb.pos * BT_TO_UU, v = b.rb.vel * BT_TO_UU;
	o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = v.x; o[4] = v.y; o[5] = v.z; o[6] = b.rb.angvel.x; o[7] = b.rb.angvel.y; o[8] = b.rb.angvel.z;
	for (int c = 0; c < nCars; c++) {
		const CarS& car = S[i].cars[c];
		float* q = o + 9 + 17 * c;
		V3 cp = car.rb.pos * BT_TO_UU, cv = car.rb.vel * BT_TO_UU, f = car.rb.basis.col(0), u = car.rb.basis.col(2);
		q[0] = cp.x; q[1] = cp.y; q[2] = cp.z; q[3] = f.x; q[4] = f.y; q[5] = f.z; q[6] = u.x; q[7] = u.y; q[8] = u.z;
		q[9] = cv.x; q[10] = cv.y; q[11] = cv.z; q[12] = car.rb.angvel.x; q[13] = car.rb.angvel.y; q[14] = car.rb.angvel.z; q[15] = car.boost;
		bool hasFlipOrJump = car.isOnGround || (!car.hasFlipped && !car.hasDoubleJumped && car.airTimeSinceJump < K::DOUBLEJUMP_MAX_DELAY);
		q[16] = (float)((car.isOnGround ? 1 : 0) | (car.isDemoed ? 2 : 0) | (car.isSupersonic ? 4 : 0) | (hasFlipOrJump ? 8 : 0));
	}
}
RSB_GLOBAL void seed_rng_kernel(ArenaS* S, uint32_t arena, uint32_t seed) {
	KERNEL_1D(i, 1);
	S[arena].rng = minstd_seed(seed);
}
RSB_GLOBAL void fill_controls_kernel(CtrlS* out, int nArenas, int nCars, int nTicks, uint32_t tick0, uint64_t seed) {
	KERNEL_1D(i, nArenas * nCars);
	int arena = i / nCars, car = i % nCars;
	RsbInputSched s; rsb_inp_init(&s);
	for (uint32_t t = 0; t < tick0 + (uint32_t)nTicks; t++) {
		rsb_inp_update(&s, seed, (uint32_t)arena, (uint32_t)car, t);
		if (t >= tick0) {
			CtrlS c; c.throttle = s.throttle; c.steer = s.steer; c.pitch = s.pitch; c.yaw = s.yaw; c.roll = s.roll;
			c.jump = s.jump; c.boost = s.boost; c.handbrake = s.handbrake; c._pad = 0;
			out[((size_t)(t - tick0) * nArenas + arena) * nCars + car] = c;
		}
	}
}

// Arena::ResetToRandomKickoff (Arena.cpp:113-223) on device, one thread per listed arena.
RSB_GLOBAL void reset_kickoff_kernel(ArenaS* S, const BatchParams* P, const uint32_t* ai, const int32_t* seeds, int n) {
	KERNEL_1D(i, n);
	const int arena = ai ? (int)ai[i] : i;
	kickoff_reset(S[arena], S[arena], *P, seeds[i]);
}

// ------------------------------------------------------------------ parameter setup (host)
static void fill_mut(const RsbMutatorConfig& m, MutD& o) {
	o.gravity = V3(m.gravity.x, m.gravity.y, m.gravity.z);
	o.carMass = m.carMass; o.carWorldFriction = m.carWorldFriction; o.carWorldRestitution = m.carWorldRestitution;
	o.ballMass = m.ballMass; o.ballMaxSpeed = m.ballMaxSpeed; o.ballDrag = m.ballDrag; o.ballWorldFriction = m.ballWorldFriction; o.ballWorldRestitution = m.ballWorldRestitution;
	o.jumpAccel = m.jumpAccel; o.jumpImmediateForce = m.jumpImmediateForce; o.boostAccelGround = m.boostAccelGround; o.boostAccelAir = m.boostAccelAir; o.boostUsedPerSecond = m.boostUsedPerSecond;
	o.respawnDelay = m.respawnDelay; o.bumpCooldownTime = m.bumpCooldownTime; o.boostPadCooldown_Big = m.boostPadCooldown_Big; o.boostPadCooldown_Small = m.boostPadCooldown_Small; o.carSpawnBoostAmount = m.carSpawnBoostAmount;
	o.ballHitExtraForceScale = m.ballHitExtraForceScale; o.bumpForceScale = m.bumpForceScale; o.ballRadius = m.ballRadius;
	o.rechargeBoostPerSecond = m.rechargeBoostPerSecond; o.rechargeBoostDelay = m.rechargeBoostDelay; o.goalBaseThresholdY = m.goalBaseThresholdY;
	o.unlimitedFlips = m.unlimitedFlips; o.unlimitedDoubleJumps = m.unlimitedDoubleJumps; o.rechargeBoostEnabled = m.rechargeBoostEnabled; o.demoMode = m.demoMode; o.enableTeamDemos = m.enableTeamDemos;
}

// MutatorConfig(GameMode) (MutatorConfig.cpp:5-40)
static void default_mutator(int mode, RsbMutatorConfig& m) {
	memset(&m, 0, sizeof(m));
	m.gravity.x = 0; m.gravity.y = 0; m.gravity.z = K::GRAVITY_Z;
	m.carMass = K::CAR_MASS_BT; m.carWorldFriction = 0.3f; m.carWorldRestitution = 0.3f;
	m.ballMaxSpeed = 6000.f; m.ballDrag = 0.03f;
	m.jumpAccel = 4375.f / 3.f; m.jumpImmediateForce = 875.f / 3.f;
	m.boostAccelGround = 2975 / 3.f; m.boostAccelAir = 3175 / 3.f; m.boostUsedPerSecond = K::BOOST_MAX / 3;
	m.respawnDelay = 3.f; m.bumpCooldownTime = 0.25f; m.boostPadCooldown_Big = 10; m.boostPadCooldown_Small = 4;
	m.carSpawnBoostAmount = K::BOOST_MAX / 3;
	m.ballHitExtraForceScale = 1; m.bumpForceScale = 1;
	m.rechargeBoostPerSecond = 10; m.rechargeBoostDelay = 0.25f;
	m.demoMode = RSB_DEMO_NORMAL; m.enableTeamDemos = 0; m.goalBaseThresholdY = 5124.25f;
	switch (mode) {
	case HOOPS: m.ballRadius = 96.3831f; break;
	case SNOWDAY: m.ballRadius = K::PUCK_RADIUS; break;
	default: m.ballRadius = 91.25f;
	}
	if (mode == SNOWDAY) { m.ballWorldFriction = 0.1f; m.ballWorldRestitution = 0.3f; m.ballMass = 50; }
	else { m.ballWorldFriction = 0.35f; m.ballWorldRestitution = 0.6f; m.ballMass = K::BALL_MASS_BT; }
	if (mode == HEATSEEKER) { m.carSpawnBoostAmount = 100; m.boostUsedPerSecond = 0; }
	else if (mode == DROPSHOT) { m.carSpawnBoostAmount = 100; m.rechargeBoostEnabled = 1; }
}

// Car::_BulletSetup (Car.cpp:195-292)
static void fill_car_cfg(const RsbCarConfig& c, CarCfgD& o) {
	V3 size(c.hitboxSize.x, c.hitboxSize.y, c.hitboxSize.z);
	o.hitboxHalf = (size * UU_TO_BT) / 2.f;
	const float defMargin = 0.04f;
	o.hitboxImplicit = o.hitboxHalf - V3(defMargin, defMargin, defMargin);
	float minDim = fminf_(fminf_(o.hitboxHalf.x, o.hitboxHalf.y), o.hitboxHalf.z);
	// btVector3::minAxis picks the axis; value is the same
	float safeMargin = 0.1f * minDim;
	o.hitboxMargin = (safeMargin < defMargin) ? safeMargin : defMargin;
	o.hitboxOffset = V3(c.hitboxPosOffset.x, c.hitboxPosOffset.y, c.hitboxPosOffset.z) * UU_TO_BT;
	// btBoxShape::calculateLocalInertia on half extents with margin
	V3 he = o.hitboxImplicit + V3(o.hitboxMargin, o.hitboxMargin, o.hitboxMargin);
	float lx = 2.f * he.x, ly = 2.f * he.y, lz = 2.f * he.z;
	const float mass = K::CAR_MASS_BT;
	V3 inertia(mass / 12.0f * (ly * ly + lz * lz), mass / 12.0f * (lx * lx + lz * lz), mass / 12.0f * (lx * lx + ly * ly));
	o.invInertiaLocal = V3(inertia.x != 0.f ? 1.f / inertia.x : 0.f, inertia.y != 0.f ? 1.f / inertia.y : 0.f, inertia.z != 0.f ? 1.f / inertia.z : 0.f);
	for (int i = 0; i < 4; i++) {
		bool front = i < 2, left = (i % 2) != 0;
		const RsbWheelPairConfig& wp = front ? c.frontWheels : c.backWheels;
		V3 off(wp.connectionPointOffset.x, wp.connectionPointOffset.y, wp.connectionPointOffset.z);
		if (left) off.y *= -1.f;
		float rest = wp.suspensionRestLength;
		rest -= K::MAX_SUSPENSION_TRAVEL;
		o.wheelConn[i] = off * UU_TO_BT;
		o.suspRest[i] = rest * UU_TO_BT;
		o.wheelRadius[i] = wp.wheelRadius * UU_TO_BT;
		o.forceScale[i] = front ? K::SUSPENSION_FORCE_SCALE_FRONT : K::SUSPENSION_FORCE_SCALE_BACK;
	}
	o.dodgeDeadzone = c.dodgeDeadzone;
	o.threeWheels = c.threeWheels;
}

static const float LOCS_SMALL_SOCCAR[28][3] = {{0, -4240, 70}, {-1792, -4184, 70}, {1792, -4184, 70}, {-940, -3308, 70}, {940, -3308, 70}, {0, -2816, 70}, {-3584, -2484, 70}, {3584, -2484, 70}, {-1788, -2300, 70}, {1788, -2300, 70}, {-2048, -1036, 70}, {0, -1024, 70}, {2048, -1036, 70}, {-1024, 0, 70}, {1024, 0, 70}, {-2048, 1036, 70}, {0, 1024, 70}, {2048, 1036, 70}, {-1788, 2300, 70}, {1788, 2300, 70}, {-3584, 2484, 70}, {3584, 2484, 70}, {0, 2816, 70}, {-940, 3308, 70}, {940, 3308, 70}, {-1792, 4184, 70}, {1792, 4184, 70}, {0, 4240, 70}};
static const float LOCS_BIG_SOCCAR[6][3] = {{-3584, 0, 73}, {3584, 0, 73}, {-3072, 4096, 73}, {3072, 4096, 73}, {-3072, -4096, 73}, {3072, -4096, 73}};
static const float LOCS_SMALL_HOOPS[14][3] = {{1536, -1024, 64}, {-1280, -2304, 64}, {0, -2816, 64}, {-1536, -1024, 64}, {1280, -2304, 64}, {-512, 512, 64}, {-1536, 1024, 64}, {1536, 1024, 64}, {1280, 2304, 64}, {0, 2816, 64}, {512, 512, 64}, {512, -512, 64}, {-512, -512, 64}, {-1280, 2304, 64}};
static const float LOCS_BIG_HOOPS[6][3] = {{-2176, 2944, 72}, {2176, -2944, 72}, {-2176, -2944, 72}, {-2432, 0, 72}, {2432, 0, 72}, {2175.99f, 2944, 72}};

// Arena ctor pad setup (Arena.cpp:498-534), BoostPad::_Setup (BoostPad.cpp:37-49), BoostPadGrid::Add
static void fill_pads(int mode, BatchParams& P, const RsbArenaConfig& ac) {
	for (int i = 0; i < 80; i++) P.padGrid[i] = -1;
	P.nPads = 0; P.customPads = 0;
	if (mode == THE_VOID || mode == DROPSHOT) return;
	bool hoops = mode == HOOPS;
	int nSmall = hoops ? 14 : 28;
	P.nPads = 6 + nSmall;
	if (ac.useCustomBoostPads) {
		if (ac.numCustomBoostPads > (uint32_t)MAX_PADS) throw std::runtime_error("ArenaConfig::customBoostPads: at most " + std::to_string(MAX_PADS) + " pads are supported by the batched back end");
		if (ac.numCustomBoostPads && !ac.customBoostPads) throw std::runtime_error("ArenaConfig::customBoostPads is NULL");
		P.nPads = (int)ac.numCustomBoostPads; P.customPads = 1;
	}
	for (int i = 0; i < P.nPads; i++) {
		bool big = i < 6;
		const float* p = hoops ? (big ? LOCS_BIG_HOOPS[i] : LOCS_SMALL_HOOPS[i - 6]) : (big ? LOCS_BIG_SOCCAR[i] : LOCS_SMALL_SOCCAR[i - 6]);
		float cp[3];
		if (P.customPads) { big = ac.customBoostPads[i].isBig != 0; cp[0] = ac.customBoostPads[i].pos.x; cp[1] = ac.customBoostPads[i].pos.y; cp[2] = ac.customBoostPads[i].pos.z; p = cp; }
		PadCfgD& d = P.pads[i];
		V3 posUU(p[0], p[1], p[2]);
		d.posBT = posUU * UU_TO_BT;
		d.isBig = big;
		float boxRad = (big ? K::PAD_BOX_RAD_BIG : K::PAD_BOX_RAD_SMALL) * UU_TO_BT;
		d.boxMin = d.posBT - V3(boxRad, boxRad, 0.f);
		d.boxMax = d.posBT + V3(boxRad, boxRad, K::PAD_BOX_HEIGHT * UU_TO_BT);
		d.cylRadBT = (big ? K::PAD_CYL_RAD_BIG : K::PAD_CYL_RAD_SMALL) * UU_TO_BT;
		if (P.customPads) continue;  // no BoostPadGrid::Add for custom pads (Arena.cpp:505-512)
		int ix = (int)(posUU.x / 1024 + 4), iy = (int)(posUU.y / 1024 + 5);
		if (ix < 0 || ix >= 8 || iy < 0 || iy >= 10 || P.padGrid[ix * 10 + iy] != -1) throw std::runtime_error("BoostPadGrid::Add(): pad cell collision");
		P.padGrid[ix * 10 + iy] = i;
	}
}

// Ball::_BulletSetup / MakeBallCollisionShape (Ball.cpp:53-115)
static void fill_ball(int mode, BatchParams& P, bool noBallRot) {
	const MutD& m = P.mut;
	P.ballInvMass = 1.0f / m.ballMass;
	P.ballMass = 1.0f / P.ballInvMass;  // btRigidBody::getMass()
	P.ballIsSphere = mode != SNOWDAY;
	P.ballNoRot = (P.ballIsSphere && noBallRot) ? 1 : 0;  // ArenaConfig::noBallRot, spheres only (Ball.cpp:53-115)
	P.ballRadiusBT = m.ballRadius * UU_TO_BT;
	V3 inertia;
	if (P.ballIsSphere) {
		float elem = 0.4f * m.ballMass * P.ballRadiusBT * P.ballRadiusBT;
		inertia = V3(elem, elem, elem);
		P.puckLocalAabbMin = P.puckLocalAabbMax = v3zero();
	} else {
		float angStep = (float)((3.14159265358979323846 * 2) / 20.0);
		float curAng = 0.f;
		for (int i = 0; i < 20; i++) {
			V3 pt(cosf(curAng) * m.ballRadius * UU_TO_BT, sinf(curAng) * m.ballRadius * UU_TO_BT, K::PUCK_HEIGHT / 2 * UU_TO_BT);
			P.puckPts[2 * i] = pt;
			pt.z *= -1.f;
			P.puckPts[2 * i + 1] = pt;
			curAng += angStep;
		}
		// btPolyhedralConvexAabbCachingShape::recalcLocalAabb: support (with margin) along +-axes
		Convex hull; hull.kind = SH_HULL; hull.margin = 0.04f; hull.pts = P.puckPts; hull.npts = PUCK_POINTS; hull.he = v3zero();
		const V3 dirs[6] = {V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1), V3(-1, 0, 0), V3(0, -1, 0), V3(0, 0, -1)};
		V3 sup[6];
		for (int i = 0; i < 6; i++) sup[i] = support_no_margin(hull, dirs[i]);
		for (int i = 0; i < 3; i++) {
			P.puckLocalAabbMax.set(i, sup[i][i] + hull.margin);
			P.puckLocalAabbMin.set(i, sup[i + 3][i] - hull.margin);
		}
		// btPolyhedralConvexShape::calculateLocalInertia (AABB box approximation incl. margin)
		float margin = hull.margin;
		V3 amin, amax;
		transform_aabb_mm(P.puckLocalAabbMin, P.puckLocalAabbMax, margin, m3identity(), v3zero(), amin, amax);
		V3 halfExtents = (amax - amin) * 0.5f;
		float lx = 2.f * (halfExtents.x + margin), ly = 2.f * (halfExtents.y + margin), lz = 2.f * (halfExtents.z + margin);
		float x2 = lx * lx, y2 = ly * ly, z2 = lz * lz;
		float scaledmass = m.ballMass * 0.08333333f;
		inertia = scaledmass * V3(y2 + z2, x2 + z2, x2 + y2);
	}
	P.ballInvInertiaLocal = V3(inertia.x != 0.f ? 1.f / inertia.x : 0.f, inertia.y != 0.f ? 1.f / inertia.y : 0.f, inertia.z != 0.f ? 1.f / inertia.z : 0.f);
}

// The seven hitbox / wheel presets (values of src/Sim/Car/CarConfig/CarConfig.cpp:20-100; order = RSB_CAR_*).
struct PresetRow { float size[3], off[3], fr, br, fs, bs, fo[3], bo[3]; int three; };
static const PresetRow CAR_PRESETS[7] = {
	{{120.507f, 86.6994f, 38.6591f}, {13.8757f, 0, 20.755f}, 12.50f, 15.00f, 38.755f, 37.055f, {51.25f, 25.90f, 20.755f}, {-33.75f, 29.50f, 20.755f}, 0},    // OCTANE
	{{130.427f, 85.7799f, 33.8f}, {9.f, 0, 15.75f}, 12.00f, 13.50f, 33.95f, 33.85f, {50.30f, 31.10f, 15.75f}, {-34.75f, 33.00f, 15.75f}, 0},                // DOMINUS
	{{131.32f, 87.1704f, 31.8944f}, {9.00857f, 0, 12.0942f}, 12.50f, 17.00f, 31.9242f, 27.9242f, {49.97f, 27.80f, 12.0942f}, {-35.43f, 20.28f, 12.0942f}, 0}, // PLANK
	{{133.992f, 83.021f, 32.8f}, {12.5f, 0, 11.75f}, 13.50f, 15.00f, 29.7f, 29.666f, {51.50f, 26.67f, 11.75f}, {-35.75f, 35.00f, 11.75f}, 0},               // BREAKOUT
	{{129.519f, 84.6879f, 36.6591f}, {13.8757f, 0, 20.755f}, 12.50f, 15.00f, 38.755f, 37.055f, {51.25f, 25.90f, 20.755f}, {-34.00f, 29.50f, 20.755f}, 0},   // HYBRID
	{{123.22f, 79.2103f, 44.1591f}, {11.3757f, 0, 21.505f}, 15.00f, 15.00f, 39.505f, 39.105f, {51.25f, 25.90f, 21.505f}, {-33.75f, 29.50f, 21.505f}, 0},    // MERC
	{{120.507f + 0.134f, 86.6994f + 0.134f, 38.6591f + 0.134f}, {13.8757f, 0, 15.0f}, 12.50f, 15.00f, 33.0f, 31.3f, {51.25f, 5.000f, 15.000f}, {-33.75f, 29.50f, 15.000f}, 1},  // PSYCLOPS
};
static RsbCarConfig car_preset(int k) {
	const PresetRow& r = CAR_PRESETS[k];
	RsbCarConfig c; memset(&c, 0, sizeof(c));
	c.hitboxSize = RsbVec{r.size[0], r.size[1], r.size[2], 0}; c.hitboxPosOffset = RsbVec{r.off[0], r.off[1], r.off[2], 0};
	c.frontWheels.wheelRadius = r.fr; c.frontWheels.suspensionRestLength = r.fs; c.frontWheels.connectionPointOffset = RsbVec{r.fo[0], r.fo[1], r.fo[2], 0};
	c.backWheels.wheelRadius = r.br; c.backWheels.suspensionRestLength = r.bs; c.backWheels.connectionPointOffset = RsbVec{r.bo[0], r.bo[1], r.bo[2], 0};
	c.threeWheels = (uint8_t)r.three; c.dodgeDeadzone = 0.5f;  // CarConfig.h:36
	return c;
}
static const RsbCarConfig& preset_octane() { static const RsbCarConfig c = car_preset(0); return c; }

// scratch layout for the indexed setters: [arena idx | car idx | payload]
template <class T>
static void stage_set(RsbBatch* b, const uint32_t* ai, const uint32_t* ci, int n, const T* payload, size_t payloadCount, uint32_t** dai, uint32_t** dci, T** dpl) {
	size_t idxBytes = ((size_t)n * 4 + 15) & ~(size_t)15;
	size_t total = 2 * idxBytes + payloadCount * sizeof(T);
	b->ensure_scratch(total);
	char* base = (char*)b->dScratch;
	*dai = (uint32_t*)base; *dci = (uint32_t*)(base + idxBytes); *dpl = (T*)(base + 2 * idxBytes);
	h2d(*dai, ai, (size_t)n * 4, b->stream);
	if (ci) h2d(*dci, ci, (size_t)n * 4, b->stream);
	h2d(*dpl, payload, payloadCount * sizeof(T), b->stream);
}
static void check_idx(RsbBatch* b, const uint32_t* ai, const uint32_t* ci, int n) {
	for (int i = 0; i < n; i++) {
		if (ai[i] >= (uint32_t)b->hp.nArenas) throw std::runtime_error("arena index out of range");
		if (ci && ci[i] >= (uint32_t)b->hp.nCars) throw std::runtime_error("car index out of range");
	}
}

// ------------------------------------------------------------------ C ABI
#define API_BEGIN try {
#define API_END } catch (const std::exception& e) { return fail(e.what()); } return 0;

extern "C" {

const char* rsb_last_error(void) { return g_err.c_str(); }

int rsb_sizeof(const char* n) {
	std::string s(n);
	if (s == "RsbVec") return sizeof(RsbVec); if (s == "RsbRotMat") return sizeof(RsbRotMat);
	if (s == "RsbCarControls") return sizeof(RsbCarControls); if (s == "RsbBallHitInfo") return sizeof(RsbBallHitInfo);
	if (s == "RsbCarState") return sizeof(RsbCarState); if (s == "RsbBallState") return sizeof(RsbBallState);
	if (s == "RsbMutatorConfig") return sizeof(RsbMutatorConfig); if (s == "RsbCarConfig") return sizeof(RsbCarConfig);
	if (s == "RsbBoostPadState") return sizeof(RsbBoostPadState); if (s == "RsbCarInternal") return sizeof(RsbCarInternal);
	if (s == "RsbBallInternal") return sizeof(RsbBallInternal); if (s == "RsbEvent") return sizeof(RsbEvent);
	if (s == "RsbBoostPadConfig") return sizeof(RsbBoostPadConfig); if (s == "RsbArenaConfig") return sizeof(RsbArenaConfig);
	if (s == "ArenaS") return sizeof(ArenaS); if (s == "ArenaW") return sizeof(ArenaW);
	return -1;
}

int rsb_car_config_preset(int preset, RsbCarConfig* out) {
	if (preset < 0 || preset > 6 || !out) return fail("car preset must be 0..6 (RSB_CAR_*)");
	*out = car_preset(preset);
	return 0;
}
int rsb_default_mutator_config(int game_mode, RsbMutatorConfig* out) {
	if (game_mode < 0 || game_mode > THE_VOID || !out) return fail("bad game mode");
	default_mutator(game_mode, *out);
	return 0;
}

int rsb_init_from_mem(int n_meshes, const int* mesh_modes, const uint8_t* const* mesh_ptrs, const size_t* mesh_sizes) {
	API_BEGIN
	std::lock_guard<std::mutex> lk(g_initMutex);
	if (g_initialized) return 0;  // idempotent like RocketSim::InitFromMem (RocketSim.cpp:98-106)
	std::map<int, std::vector<MeshHost>> meshes;
	for (int i = 0; i < n_meshes; i++) meshes[mesh_modes[i]].push_back(build_mesh(mesh_ptrs[i], mesh_sizes[i]));
	for (auto& kv : meshes) g_sceneHost[kv.first] = build_scene(kv.first, std::move(kv.second));
	g_sceneHost[THE_VOID] = build_scene(THE_VOID, {});
	g_initialized = true;
	API_END
}

void rsb_default_arena_config(RsbArenaConfig* o) {
	if (!o) return;
	memset(o, 0, sizeof(*o));
	o->memWeightMode = RSB_ARENA_MEM_HEAVY; o->noBallRot = 1; o->useCustomBroadphase = 1; o->useCustomBoostPads = 0;
	o->maxAABBLen = 370.f; o->maxObjects = 512;
	o->minPos = RsbVec{-5600, -6000, 0, 0}; o->maxPos = RsbVec{5600, 6000, 2200, 0};
}

int rsb_batch_create(int device, int game_mode, int n_arenas, int n_cars, const int* teams, const RsbCarConfig* car_configs,
                     const RsbMutatorConfig* mutator, float tick_rate, uint32_t seed, RsbBatch** out) {
	return rsb_batch_create_ex(device, game_mode, nullptr, n_arenas, n_cars, teams, car_configs, mutator, tick_rate, seed, out);
}

int rsb_batch_create_ex(int device, int game_mode, const RsbArenaConfig* arena_config, int n_arenas, int n_cars, const int* teams,
                        const RsbCarConfig* car_configs, const RsbMutatorConfig* mutator, float tick_rate, uint32_t seed, RsbBatch** out) {
	API_BEGIN
	if (!out) throw std::runtime_error("out is NULL");
	*out = nullptr;
	RsbArenaConfig ac; rsb_default_arena_config(&ac);
	if (arena_config) {
		ac = *arena_config;
		RsbArenaConfig d; rsb_default_arena_config(&d);
		// the broadphase grid is part of the static scene built per game mode: only the default ArenaConfig geometry is available
		if (ac.memWeightMode != d.memWeightMode) throw std::runtime_error("ArenaConfig::memWeightMode = LIGHT is not supported by the batched back end (the broadphase grid is built once per game mode with the HEAVY cell size)");
		if (memcmp(&ac.minPos, &d.minPos, 12) || memcmp(&ac.maxPos, &d.maxPos, 12)) throw std::runtime_error("ArenaConfig::minPos / maxPos other than the defaults are not supported by the batched back end");
		if (ac.maxAABBLen != d.maxAABBLen) throw std::runtime_error("ArenaConfig::maxAABBLen other than 370 is not supported by the batched back end");
		if (!ac.useCustomBroadphase) throw std::runtime_error("ArenaConfig::useCustomBroadphase = false (btDbvtBroadphase) is not supported by the batched back end");
	}
	if (n_arenas <= 0) throw std::runtime_error("n_arenas must be positive");
	if (n_cars < 0 || n_cars > MAX_CARS) throw std::runtime_error("n_cars must be in [0, " + std::to_string(MAX_CARS) + "]");
	if (!(tick_rate >= 15 && tick_rate <= 120)) throw std::runtime_error("Tickrate must be from 15 to 120tps");  // Arena.cpp:419
	{
		std::lock_guard<std::mutex> lk(g_initMutex);
		if (!g_initialized) throw std::runtime_error("Cannot create Arena, RocketSim has not been initialized, call rsb_init_from_mem() first");
		const auto hs = g_sceneHost.find(game_mode);
		if (hs == g_sceneHost.end() || (game_mode != THE_VOID && hs->second.meshes.empty()))
			throw std::runtime_error("No arena meshes found for gamemode " + std::to_string(game_mode));  // Arena.cpp:997-1002
	}
#ifndef RSB_HOSTCHECK
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw std::runtime_error("no CUDA device: rocketsim_b200 has no CPU path");
	if (device < 0 || device >= ndev) throw std::runtime_error("bad CUDA device index");
	CUDA_OK(cudaSetDevice(device));
#endif
	RsbBatch* b = new RsbBatch();
	b->device = device;
	b->acfg = ac;
	if (ac.useCustomBoostPads && ac.numCustomBoostPads <= (uint32_t)MAX_PADS && ac.customBoostPads) b->acfgPads.assign(ac.customBoostPads, ac.customBoostPads + ac.numCustomBoostPads);
	b->acfg.customBoostPads = b->acfgPads.empty() ? nullptr : b->acfgPads.data();
#ifndef RSB_HOSTCHECK
	{
		cudaDeviceProp prop;
		CUDA_OK(cudaGetDeviceProperties(&prop, device));
		b->numSMs = prop.multiProcessorCount;
		if ((size_t)prop.sharedMemPerBlockOptin < 227 * 1024) throw std::runtime_error("device offers less than 227 KB of shared memory per block: rocketsim_b200 targets sm_100a");
		static std::mutex attrMutex; std::lock_guard<std::mutex> lk(attrMutex);
		cudaFuncAttributes fa;
		const StepKernel kern = step_kernel_of(game_mode);
		CUDA_OK(cudaFuncGetAttributes(&fa, kern));
		b->smemBudget = SMEM_MAX - fa.sharedSizeBytes;  // dynamic shared memory left beside the kernel's static part
		CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smemBudget));
	}
#endif
	BatchParams& P = b->hp;
	memset(&P, 0, sizeof(P));
	P.mode = game_mode; P.nCars = n_cars; P.nArenas = n_arenas;
	P.dt = 1 / tick_rate;
	RsbMutatorConfig mc;
	if (mutator) mc = *mutator; else default_mutator(game_mode, mc);
	b->mutApi = mc; b->tickRate = tick_rate;
	fill_mut(mc, P.mut);
	P.ballLinDampFactor = powf(1.f - P.mut.ballDrag, P.dt);
	P.flipZDampFactor = powf(1.f - K::FLIP_Z_DAMP_120, P.dt / (1 / 120.f));
	P.gravityBT = P.mut.gravity * UU_TO_BT;
	fill_ball(game_mode, P, ac.noBallRot != 0);
	for (int i = 0; i < n_cars; i++) {
		fill_car_cfg(car_configs ? car_configs[i] : preset_octane(), P.carCfg[i]);
		b->carCfgApi.push_back(car_configs ? car_configs[i] : preset_octane());
		P.team[i] = teams ? teams[i] : (i < (n_cars + 1) / 2 ? 0 : 1);
	}
	fill_pads(game_mode, P, ac);
	for (int team = 0; team < 2; team++) {
		auto yaw_of = [&](float yaw) { return team == 0 ? yaw : (float)((double)yaw + 3.14159265358979323846); };
		for (int i = 0; i < 5; i++) P.kickBasis[team][i] = basis_from_ypr(yaw_of(kickoff_slot(game_mode, i).yaw), 0.f, 0.f);
		for (int i = 0; i < 4; i++) P.respBasis[team][i] = basis_from_ypr(yaw_of(respawn_slot(game_mode == HOOPS, i).yaw), 0.f, 0.f);
	}
	{
		ArenaS tmpS; memset(&tmpS, 0, sizeof(tmpS));
		Ctx cx; cx.s = &tmpS; cx.w = nullptr; cx.epa = nullptr; cx.body = nullptr; cx.wheels = nullptr; cx.padLocked = nullptr; cx.misc = nullptr; cx.pl = nullptr; cx.mf = nullptr; cx.solv = nullptr; cx.p = &P; cx.sc = nullptr; cx.arena = 0; cx.ev.buf = nullptr; ctx_set_mode(cx, P, nullptr);
		for (int bidx = 0; bidx <= n_cars; bidx++) P.motionDisc[bidx] = body_angular_motion_disc(cx, bidx);
		P.boundRadius[0] = P.ballIsSphere ? P.ballRadiusBT : (sqrtf(P.ballRadiusBT * P.ballRadiusBT + (K::PUCK_HEIGHT * UU_TO_BT / 2) * (K::PUCK_HEIGHT * UU_TO_BT / 2)) + 0.04f);
		for (int c = 0; c < n_cars; c++) P.boundRadius[c + 1] = len(P.carCfg[c].hitboxHalf) + 0.04f;
	}
	SceneDev* const sdev = get_scene_dev(device, game_mode);
	b->dScene = sdev->dScene;
	const SceneHost& sh = sdev->host;
	b->dP = (BatchParams*)dev_alloc(sizeof(BatchParams));
	h2d(b->dP, &P, sizeof(P), 0);
	b->dS = (ArenaS*)dev_alloc(sizeof(ArenaS) * (size_t)n_arenas);
	b->dW = (ArenaW*)dev_alloc(sizeof(ArenaW) * (size_t)n_arenas);
	b->dH = (ArenaH*)dev_alloc(sizeof(ArenaH) * (size_t)n_arenas);
#ifndef RSB_HOSTCHECK
	b->dE = (EpaScratch*)dev_alloc(sizeof(EpaScratch) * (size_t)b->numSMs * BLOCKS_PER_SM * STEP_THREADS);
#else
	b->dE = (EpaScratch*)dev_alloc(sizeof(EpaScratch));
#endif
	b->eventCap = (uint32_t)std::max(4096, n_arenas * 4);
	b->dChunkNext = (int*)dev_alloc(sizeof(int));
	b->dEvents = (EventSink::Ev*)dev_alloc(sizeof(EventSink::Ev) * b->eventCap);
	b->dEventCount = (uint32_t*)dev_alloc(sizeof(uint32_t));
	b->dCtrl = (CtrlS*)dev_alloc(sizeof(CtrlS) * (size_t)n_arenas * std::max(1, n_cars));

	// initial arena record: Arena ctor + AddCar/Respawn (Arena.cpp:50-62, 416-540)
	std::vector<ArenaS> init((size_t)n_arenas);
	ArenaW tmpW; memset(&tmpW, 0, sizeof(tmpW));
	ArenaH tmpH; memset(&tmpH, 0, sizeof(tmpH));
	Scene hsc; memset(&hsc, 0, sizeof(hsc));
	hsc.gridMin = sh.gridMin; hsc.gridMax = sh.gridMax; hsc.cellSize = sh.cellSize; hsc.cellsX = sh.cellsX; hsc.cellsY = sh.cellsY; hsc.cellsZ = sh.cellsZ;
	for (int a = 0; a < n_arenas; a++) {
		ArenaS& A = init[a];
		memset(&A, 0, sizeof(A));
		A.rng = minstd_seed(seed + (uint32_t)a);
		A.ball.rb.pos = V3(0, 0, K::BALL_REST_Z) * UU_TO_BT;  // ball->SetState(BallState())
		A.ball.rb.basis = m3identity();
		A.ball.hsCurTargetSpeed = K::HS_INITIAL_TARGET_SPEED;
		A.dsChargeLevel = 1;  // BallState::DropshotInfo default; every tile starts intact (state 0)
		for (int p = 0; p < P.nPads; p++) { A.pads[p].isActive = 1; }
		Ctx cx; cx.s = &A; cx.g = &A; cx.w = &tmpW; cx.epa = nullptr; cx.p = &P; cx.sc = &hsc; cx.arena = a; cx.ev.buf = nullptr; ctx_set_mode(cx, P, &hsc);
		cx.body = tmpH.body; cx.wheels = tmpH.wheels; cx.padLocked = tmpH.padLocked; cx.misc = tmpH.misc; cx.pl = &tmpH.pairs; cx.mf = tmpH.mf; cx.solv = tmpH.solv;
		for (int c = 0; c < n_cars; c++) {
			tmpH.body[c + 1].invInertiaLocal = P.carCfg[c].invInertiaLocal;
			car_respawn(cx, c);  // AddCar -> Respawn(gameMode, -1, carSpawnBoostAmount)
		}
		// broadphase proxies are created from the body transforms at construction time:
		// ball at (0,0,radius), cars at the origin (Ball.cpp:92-93, Car.cpp:211-215)
		for (int bidx = 0; bidx <= n_cars; bidx++) {
			V3 mn, mx;
			V3 org = (bidx == 0) ? V3(0, 0, P.mut.ballRadius * UU_TO_BT) : v3zero();
			body_shape_aabb(cx, bidx, m3identity(), org, mn, mx);
			int i, j, k;
			A.cellIdx[bidx] = grid_cell_of(hsc, mn, i, j, k);
			A.cellI[bidx] = (uint8_t)i; A.cellJ[bidx] = (uint8_t)j; A.cellK[bidx] = (uint8_t)k;
			A.cellStamp[bidx] = (uint32_t)bidx + 1;
			A.dynOrder[bidx] = (uint8_t)bidx;
		}
		A.stampCounter = (uint32_t)n_cars + 2;
		A.worldFresh = 1;
	}
	h2d(b->dS, init.data(), sizeof(ArenaS) * (size_t)n_arenas, 0);
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaDeviceSynchronize());
#endif
	*out = b;
	API_END
}

int rsb_batch_destroy(RsbBatch* b) {
	API_BEGIN
	if (!b) return 0;
	set_device(b);
#ifndef RSB_HOSTCHECK
	cudaDeviceSynchronize();
#endif
#ifndef RSB_HOSTCHECK
	if (b->copyStream) { cudaStreamDestroy(b->copyStream); for (auto& e : b->chunkReady) cudaEventDestroy(e); cudaEventDestroy(b->seqFree); }
#endif
	if (b->pred) { rsb_batch_destroy(b->pred); b->pred = nullptr; }
	dev_free(b->dP); dev_free(b->dS); dev_free(b->dW); dev_free(b->dH); dev_free(b->dE); dev_free(b->dChunkNext); dev_free(b->dEvents); dev_free(b->dEventCount); dev_free(b->dCtrl); dev_free(b->dCtrlSeq); dev_free(b->dScratch);
	delete b;
	API_END
}

int rsb_batch_info(RsbBatch* b, int* na, int* nc, int* np, int* mode) {
	if (!b) return fail("null batch");
	if (na) *na = b->hp.nArenas; if (nc) *nc = b->hp.nCars; if (np) *np = b->hp.nPads; if (mode) *mode = b->hp.mode;
	return 0;
}

int rsb_set_stream(RsbBatch* b, void* s) {
	if (!b) return fail("null batch");
#ifndef RSB_HOSTCHECK
	b->stream = (cudaStream_t)s;
#else
	(void)s;
#endif
	return 0;
}

static EventSink sink_of(RsbBatch* b) { EventSink e; e.buf = b->dEvents; e.count = b->dEventCount; e.cap = b->eventCap; return e; }

static void launch_step(RsbBatch* b, int n_ticks, const CtrlS* seq, int nCtrl) {
	const BatchParams& P = b->hp;
	StepArgs a; a.S = b->dS; a.H = b->dH; a.W = b->dW; a.E = b->dE; a.P = b->dP;
	a.lay = ArenaLayout::make(P.nCars, P.nPads); a.sc = b->dScene; a.ev = sink_of(b);
	a.nArenas = P.nArenas; a.nTicks = n_ticks; a.ctrlSeq = seq; a.nCtrl = nCtrl; a.apb = 1; a.mirror = (n_ticks == 1) ? 1 : 0;
	a.ballTrace = b->traceOut; b->traceOut = nullptr;
	a.qcap = b->detectQueueCap > 0 ? b->detectQueueCap : DETECT_QUEUE_CAP;
#ifndef RSB_HOSTCHECK
	// chunk size: fill every SM once if the batch fits (one wave), else the fewest equal waves
	const int sms = b->numSMs * BLOCKS_PER_SM;  // resident block slots
	const int APB_MAX = (int)((b->smemBudget - SMEM_P_BYTES - SMEM_SC_BYTES - SMEM_Q_BYTES) / a.lay.stride);
	const int chunksMin = (P.nArenas + APB_MAX - 1) / APB_MAX;
	const int waves = (chunksMin + sms - 1) / sms;
	a.apb = (P.nArenas + waves * sms - 1) / (waves * sms);
	if (a.apb < 1) a.apb = 1;
	if (a.apb > APB_MAX) a.apb = APB_MAX;
	const int nChunks = (P.nArenas + a.apb - 1) / a.apb;
	const int grid = nChunks < sms ? nChunks : sms;
	const size_t smem = SMEM_P_BYTES + SMEM_SC_BYTES + SMEM_Q_BYTES + (size_t)a.apb * a.lay.stride;
	a.chunkNext = b->dChunkNext;
	CUDA_OK(cudaMemsetAsync(b->dChunkNext, 0, sizeof(int), b->stream));
	step_kernel_of(P.mode)<<<grid, STEP_THREADS, smem, b->stream>>>(a);
	CUDA_OK(cudaGetLastError());
#else
	arena_step_host(a);
#endif
}

int rsb_set_controls(RsbBatch* b, const RsbCarControls* controls) {
	API_BEGIN
	set_device(b);
	size_t n = (size_t)b->hp.nArenas * b->hp.nCars;
	if (n == 0) return 0;
	b->ensure_scratch(n * sizeof(RsbCarControls));
	h2d(b->dScratch, controls, n * sizeof(RsbCarControls), b->stream);
	LAUNCH_1D(set_controls_kernel, (int)n, b, b->dS, (const RsbCarControls*)b->dScratch, b->hp.nArenas, b->hp.nCars);
	API_END
}
int rsb_set_controls_device(RsbBatch* b, const void* d) {
	API_BEGIN
	set_device(b);
	size_t n = (size_t)b->hp.nArenas * b->hp.nCars;
	LAUNCH_1D(set_controls_kernel, (int)n, b, b->dS, (const RsbCarControls*)d, b->hp.nArenas, b->hp.nCars);
	API_END
}
int rsb_step(RsbBatch* b, int n_ticks) {
	API_BEGIN
	set_device(b);
	if (n_ticks > 0) launch_step(b, n_ticks, nullptr, 0);
	API_END
}
int rsb_step_with_controls_device(RsbBatch* b, const void* d_seq, int n_ticks) {
	API_BEGIN
	set_device(b);
	if (n_ticks > 0) launch_step(b, n_ticks, (const CtrlS*)d_seq, n_ticks);
	API_END
}
int rsb_step_with_controls(RsbBatch* b, const RsbCarControls* seq, int n_ticks) {
	API_BEGIN
	set_device(b);
	if (n_ticks <= 0) return 0;
	static_assert(sizeof(CtrlS) == sizeof(RsbCarControls), "control layouts must match");
	size_t n = (size_t)n_ticks * b->hp.nArenas * b->hp.nCars;
	if (n > b->ctrlSeqCap) { dev_free(b->dCtrlSeq); b->dCtrlSeq = (CtrlS*)dev_alloc(n * sizeof(CtrlS)); b->ctrlSeqCap = n; }
#ifndef RSB_HOSTCHECK
	// Pipeline: the sequence goes up in 3 chunks on a copy stream while the step stream already runs the ticks of the
	// chunks that have landed, so only the first chunk's copy is exposed.
	const size_t perTick = (size_t)b->hp.nArenas * b->hp.nCars;
	// a short first chunk (so compute starts after ~1/16 of the copy) and two long ones: every extra launch costs a grid
	// tail (its slowest block), so the sequence is not cut finer than the overlap needs
	const int nChunks = n_ticks >= 16 ? 3 : 1;
	auto chunk_end = [&](int c) { return nChunks == 1 ? n_ticks : (c == 0 ? n_ticks / 16 : (c == 1 ? n_ticks / 16 + (n_ticks - n_ticks / 16) / 2 : n_ticks)); };
	if (!b->copyStream) {
		CUDA_OK(cudaStreamCreateWithFlags(&b->copyStream, cudaStreamNonBlocking));
		for (auto& e : b->chunkReady) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		CUDA_OK(cudaEventCreateWithFlags(&b->seqFree, cudaEventDisableTiming));
	}
	CUDA_OK(cudaEventRecord(b->seqFree, b->stream));           // earlier launches may still read the staging buffer
	CUDA_OK(cudaStreamWaitEvent(b->copyStream, b->seqFree, 0));
	int t0 = 0;
	for (int c = 0; c < nChunks; c++) {
		const int t1 = chunk_end(c);
		if (t1 > t0) CUDA_OK(cudaMemcpyAsync(b->dCtrlSeq + t0 * perTick, seq + t0 * perTick, (size_t)(t1 - t0) * perTick * sizeof(CtrlS), cudaMemcpyHostToDevice, b->copyStream));
		CUDA_OK(cudaEventRecord(b->chunkReady[c], b->copyStream));
		t0 = t1;
	}
	t0 = 0;
	for (int c = 0; c < nChunks; c++) {
		const int t1 = chunk_end(c);
		CUDA_OK(cudaStreamWaitEvent(b->stream, b->chunkReady[c], 0));
		if (t1 > t0) launch_step(b, t1 - t0, b->dCtrlSeq + t0 * perTick, t1 - t0);
		t0 = t1;
	}
#else
	h2d(b->dCtrlSeq, seq, n * sizeof(CtrlS), b->stream);
	launch_step(b, n_ticks, b->dCtrlSeq, n_ticks);
#endif
	API_END
}
int rsb_set_auto_reset(RsbBatch* b, int enable, uint32_t seed) {
	API_BEGIN
	set_device(b);
	b->hp.autoReset = enable ? 1 : 0; b->hp.autoResetSeed = seed;
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaStreamSynchronize(b->stream));
#endif
	h2d(b->dP, &b->hp, sizeof(BatchParams), b->stream);
	API_END
}
int rsb_set_arena_offset(RsbBatch* b, uint32_t off) {
	API_BEGIN
	set_device(b);
	b->hp.arenaOffset = (int)off;
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaStreamSynchronize(b->stream));
#endif
	h2d(b->dP, &b->hp, sizeof(BatchParams), b->stream);
	API_END
}
int rsb_set_event_tracker(RsbBatch* b, int enable, const RsbEventTrackerConfig* cfg) {
	API_BEGIN
	set_device(b);
	if (enable && (b->hp.mode == HEATSEEKER || b->hp.mode == DROPSHOT || b->hp.mode == THE_VOID))
		throw std::runtime_error("Arena::IsBallProbablyGoingIn() is not supported for gamemode " + std::to_string(b->hp.mode));  // Arena.cpp:887
	b->hp.trackerEnabled = enable ? 1 : 0;
	TrackerCfgD& t = b->hp.tracker;
	if (cfg) { t.shotMinSpeed = cfg->shotMinSpeed; t.shotTouchMinDelay = cfg->shotTouchMinDelay; t.predScoreExtraMargin = cfg->predScoreExtraMargin; t.shotEventCooldown = cfg->shotEventCooldown;
	           t.shotMinScoreTime = cfg->shotMinScoreTime; t.goalMaxTouchTime = cfg->goalMaxTouchTime; t.passMaxTouchTime = cfg->passMaxTouchTime; }
	else { t.shotMinSpeed = 1750; t.shotTouchMinDelay = 0.3f; t.predScoreExtraMargin = 0; t.shotEventCooldown = 1.0f; t.shotMinScoreTime = 2.0f; t.goalMaxTouchTime = 4.0f; t.passMaxTouchTime = 2.0f; }
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaStreamSynchronize(b->stream));
#endif
	h2d(b->dP, &b->hp, sizeof(BatchParams), b->stream);
	API_END
}
RSB_GLOBAL void going_in_kernel(ArenaS* S, const BatchParams* P, float maxTime, float extraMargin, uint8_t* in, uint8_t* team, int nArenas) {
	KERNEL_1D(i, nArenas);
	Ctx cx; cx.s = &S[i]; cx.w = nullptr; cx.epa = nullptr; cx.body = nullptr; cx.wheels = nullptr; cx.padLocked = nullptr; cx.misc = nullptr; cx.pl = nullptr; cx.mf = nullptr; cx.solv = nullptr; cx.p = P; cx.sc = nullptr; cx.arena = i; cx.ev.buf = nullptr; ctx_set_mode(cx, *P, nullptr);
	int t = 0; bool sup;
	in[i] = is_ball_probably_going_in(cx, maxTime, extraMargin, t, sup) ? 1 : 0;
	if (team) team[i] = (uint8_t)t;
}
int rsb_is_ball_probably_going_in(RsbBatch* b, float max_time, float extra_margin, uint8_t* out, uint8_t* goal_team_out) {
	API_BEGIN
	set_device(b);
	if (b->hp.mode == HEATSEEKER || b->hp.mode == DROPSHOT || b->hp.mode == THE_VOID)  // Arena.cpp:775-889: Soccar, Snowday and Hoops only
		throw std::runtime_error("Arena::IsBallProbablyGoingIn() is not supported for gamemode " + std::to_string(b->hp.mode));
	const int n = b->hp.nArenas;
	b->ensure_scratch((size_t)n * 2);
	uint8_t* din = (uint8_t*)b->dScratch; uint8_t* dteam = din + n;
	LAUNCH_1D(going_in_kernel, n, b, b->dS, b->dP, max_time, extra_margin, din, dteam, n);
	d2h(out, din, (size_t)n, b->stream);
	if (goal_team_out) d2h(goal_team_out, dteam, (size_t)n, b->stream);
	API_END
}
// Ball::GetState -> Ball::SetState across two arenas (BallPredTracker::ForceUpdateAllPred's first line): through the UU-valued
// BallState, as the reference does, so the x50 / x0.02 round trip of the position is reproduced
RSB_GLOBAL void copy_ball_kernel(const ArenaS* src, ArenaS* dst, RsbBallState* row0, int nArenas) {
	KERNEL_1D(i, nArenas);
	RsbBallState st; ball_to_api(src[i].ball, st); ds_to_api(src[i], st);
	ball_from_api(st, dst[i].ball); ds_from_api(st, dst[i]);
	row0[i] = st;
}
int rsb_predict_ball(RsbBatch* b, int n_ticks, RsbBallState* out) {
	API_BEGIN
	set_device(b);
	if (n_ticks < 1) throw std::runtime_error("n_ticks must be >= 1");
	if (!b->pred) {
		RsbBatch* p = nullptr;
		if (rsb_batch_create_ex(b->device, b->hp.mode, &b->acfg, b->hp.nArenas, 0, nullptr, nullptr, &b->mutApi, b->tickRate, 0, &p) != 0) throw std::runtime_error(g_err);
		b->pred = p;
	}
	RsbBatch* p = b->pred;
	p->stream = b->stream;
	const size_t rows = (size_t)n_ticks * b->hp.nArenas;
	b->ensure_scratch(rows * sizeof(RsbBallState));
	RsbBallState* dTrace = (RsbBallState*)b->dScratch;
	LAUNCH_1D(copy_ball_kernel, b->hp.nArenas, b, b->dS, p->dS, dTrace, b->hp.nArenas);
	if (n_ticks > 1) { p->traceOut = dTrace; launch_step(p, n_ticks - 1, nullptr, 0); }
	d2h(out, dTrace, rows * sizeof(RsbBallState), b->stream);
	API_END
}
RSB_GLOBAL void get_controls_kernel(const ArenaS* S, RsbCarControls* out, int nArenas, int nCars) {
	KERNEL_1D(i, nArenas * nCars);
	ctrl_to_api(S[i / nCars].cars[i % nCars].controls, out[i]);
}
// The arena becomes what Arena::DeserializeNew leaves behind: a NEW physics world (first-step solver quirk, broadphase proxies
// registered at the construction-time poses: ball at (0,0,radius), cars at the origin -- Ball.cpp:92-93, Car.cpp:211-215) and new
// cars / ball whose state outside CarState/BallState is at its defaults.
RSB_GLOBAL void fresh_world_kernel(ArenaS* S, const BatchParams* P, const Scene* sc, uint64_t tickCount) {
	KERNEL_1D(i, 1);
	ArenaS& A = S[0];
	A.tickCount = tickCount;
	A.worldFresh = 1;
	A.trkLastBallUpdate = 0; A.trkShotCooldown = 0.f; A.trkBallShot = 0; A.trkScoredLast = 0;
	Ctx cx; cx.s = &A; cx.w = nullptr; cx.epa = nullptr; cx.body = nullptr; cx.wheels = nullptr; cx.padLocked = nullptr; cx.misc = nullptr; cx.pl = nullptr; cx.mf = nullptr; cx.solv = nullptr; cx.p = P; cx.sc = sc; cx.arena = 0; cx.ev.buf = nullptr; ctx_set_mode(cx, *P, sc);
	const int nCars = P->nCars;
	for (int b = 0; b <= nCars; b++) {
		V3 mn, mx;
		V3 org = (b == 0) ? V3(0, 0, P->mut.ballRadius * UU_TO_BT) : v3zero();
		body_shape_aabb(cx, b, m3identity(), org, mn, mx);
		int ci, cj, ck;
		A.cellIdx[b] = grid_cell_of(*sc, mn, ci, cj, ck);
		A.cellI[b] = (uint8_t)ci; A.cellJ[b] = (uint8_t)cj; A.cellK[b] = (uint8_t)ck;
		A.cellStamp[b] = (uint32_t)b + 1;
		A.dynOrder[b] = (uint8_t)b;
	}
	A.stampCounter = (uint32_t)nCars + 2;
	A.ball.groundStickApplied = 0;
	for (int t = 0; t < MAX_TILES; t++) A.tiles[t] = 0;  // a new Arena has every Dropshot tile intact (tile states are not serialised)
	for (int c = 0; c < nCars; c++) {
		CarS& car = A.cars[c];
		for (int k = 0; k < 4; k++) { car.wSteer[k] = car.wEngine[k] = car.wBrake[k] = car.wLatFriction[k] = car.wLongFriction[k] = car.wExtraPushback[k] = 0.f; }
	}
}
// Arena::Clone (Arena.cpp:646-674) from arena `src` into arena `dst` of the same batch: a NEW world whose ball, cars and pads are
// set through GetState -> SetState (so positions take the UU round trip and everything outside the state structs restarts at its
// defaults), plus the velocity-impulse caches, the controls and the tick counter, which Clone copies explicitly.
RSB_GLOBAL void clone_arena_kernel(ArenaS* S, const BatchParams* P, const Scene* sc, uint32_t src, uint32_t dst) {
	KERNEL_1D(i, 1);
	const ArenaS& A = S[src];
	ArenaS& B = S[dst];
	const int nCars = P->nCars;
	RsbBallState bs; ball_to_api(A.ball, bs); ds_to_api(A, bs); ball_from_api(bs, B.ball); ds_from_api(bs, B);
	for (int t = 0; t < MAX_TILES; t++) B.tiles[t] = 0;  // a new Arena starts with every tile intact; Clone does not copy tile states
	B.ball.velImpulseCache = A.ball.velImpulseCache;
	B.ball.groundStickApplied = 0;
	for (int c = 0; c < nCars; c++) {
		RsbCarState cs; car_to_api(A, c, cs); car_from_api(cs, B, c);
		B.cars[c].controls = A.cars[c].controls;
		B.cars[c].velImpulseCache = A.cars[c].velImpulseCache;
		for (int k = 0; k < 4; k++) { B.cars[c].wSteer[k] = B.cars[c].wEngine[k] = B.cars[c].wBrake[k] = B.cars[c].wLatFriction[k] = B.cars[c].wLongFriction[k] = B.cars[c].wExtraPushback[k] = 0.f; }
	}
	for (int p = 0; p < P->nPads; p++) { B.pads[p].isActive = A.pads[p].isActive; B.pads[p].cooldown = A.pads[p].cooldown; B.pads[p].prevLockedCarID = A.pads[p].prevLockedCarID; }
	B.rng = A.rng;
	B.errorFlags = 0;
	(void)sc;
}
int rsb_clone_arena(RsbBatch* b, uint32_t src, uint32_t dst) {
	API_BEGIN
	set_device(b);
	if (src >= (uint32_t)b->hp.nArenas || dst >= (uint32_t)b->hp.nArenas) throw std::runtime_error("arena index out of range");
	if (src == dst) return 0;
	uint64_t tick = 0;
	{
		b->ensure_scratch(16);
		LAUNCH_1D(get_misc_kernel, 1, b, b->dS + src, b->dP, b->dScene, (uint64_t*)b->dScratch, (uint8_t*)nullptr, (uint32_t*)nullptr, 1);
		d2h(&tick, b->dScratch, 8, b->stream);
	}
	LAUNCH_1D(clone_arena_kernel, 1, b, b->dS, b->dP, b->dScene, src, dst);
	LAUNCH_1D(fresh_world_kernel, 1, b, b->dS + dst, b->dP, b->dScene, tick);
	API_END
}

// Replaces Arena::Serialize (Arena.cpp:546-574) for one arena of the batch.
int rsb_serialize_arena(RsbBatch* b, uint32_t arena, uint8_t* out, size_t cap, size_t* n_out) {
	API_BEGIN
	set_device(b);
	const BatchParams& P = b->hp;
	if (arena >= (uint32_t)P.nArenas) throw std::runtime_error("arena index out of range");
	rsbser::ArenaImage img;
	img.gameMode = (uint8_t)P.mode; img.tickTime = P.dt; img.lastCarID = (uint32_t)P.nCars; img.mut = b->mutApi;
	img.noBallRot = b->acfg.noBallRot != 0;
	const size_t need = (size_t)std::max(1, P.nCars) * (sizeof(RsbCarState) + sizeof(RsbCarControls)) + sizeof(RsbBallState) + (size_t)MAX_PADS * sizeof(RsbBoostPadState) + 64;
	b->ensure_scratch(need);
	char* base = (char*)b->dScratch;
	RsbCarState* dCars = (RsbCarState*)base; RsbCarControls* dCtl = (RsbCarControls*)(dCars + std::max(1, P.nCars));
	RsbBallState* dBall = (RsbBallState*)(((uintptr_t)(dCtl + std::max(1, P.nCars)) + 15) & ~(uintptr_t)15);
	RsbBoostPadState* dPads = (RsbBoostPadState*)(dBall + 1); uint64_t* dTick = (uint64_t*)(dPads + MAX_PADS);
	ArenaS* S1 = b->dS + arena;
	std::vector<RsbCarState> cars(P.nCars); std::vector<RsbCarControls> ctl(P.nCars); img.pads.resize(P.nPads);
	if (P.nCars) {
		LAUNCH_1D(get_car_states_kernel, P.nCars, b, S1, dCars, 1, P.nCars);
		LAUNCH_1D(get_controls_kernel, P.nCars, b, S1, dCtl, 1, P.nCars);
		d2h(cars.data(), dCars, sizeof(RsbCarState) * P.nCars, b->stream); d2h(ctl.data(), dCtl, sizeof(RsbCarControls) * P.nCars, b->stream);
	}
	LAUNCH_1D(get_ball_states_kernel, 1, b, S1, dBall, 1, b->hp.mode == DROPSHOT);
	d2h(&img.ball, dBall, sizeof(RsbBallState), b->stream);
	if (P.nPads) { LAUNCH_1D(get_pad_states_kernel, P.nPads, b, S1, dPads, 1, P.nPads); d2h(img.pads.data(), dPads, sizeof(RsbBoostPadState) * P.nPads, b->stream); }
	LAUNCH_1D(get_misc_kernel, 1, b, S1, b->dP, b->dScene, dTick, (uint8_t*)nullptr, (uint32_t*)nullptr, 1);
	d2h(&img.tickCount, dTick, 8, b->stream);
	for (int i = 0; i < P.nCars; i++) {
		rsbser::ArenaImage::CarImg c; c.team = (uint8_t)P.team[i]; c.id = (uint32_t)i + 1; c.controls = ctl[i]; c.config = b->carCfgApi[i]; c.state = cars[i];
		img.cars.push_back(c);
	}
	const std::vector<uint8_t> bytes = rsbser::write_arena(img);
	if (n_out) *n_out = bytes.size();
	if (out) { if (cap < bytes.size()) throw std::runtime_error("output buffer too small"); memcpy(out, bytes.data(), bytes.size()); }
	API_END
}
// Replaces Arena::DeserializeNew (Arena.cpp:576-644) INTO an existing arena of the batch.  What a batch fixes for all of its
// arenas (game mode, tick rate, car slots with their teams and configs, mutators) must match the stream.
int rsb_deserialize_arena(RsbBatch* b, uint32_t arena, const uint8_t* data, size_t n) {
	API_BEGIN
	set_device(b);
	const BatchParams& P = b->hp;
	if (arena >= (uint32_t)P.nArenas) throw std::runtime_error("arena index out of range");
	rsbser::ArenaImage img = rsbser::read_arena(data, n, [](int mode) { return mode == HOOPS ? 20 : ((mode == THE_VOID || mode == DROPSHOT) ? 0 : 34); });
	if (img.gameMode != P.mode) throw std::runtime_error("Arena::Deserialize(): the stream's game mode differs from the batch's");
	if (img.tickTime != P.dt) throw std::runtime_error("Arena::Deserialize(): the stream's tick rate differs from the batch's");
	if (img.noBallRot != (b->acfg.noBallRot != 0)) throw std::runtime_error("Arena::Deserialize(): the stream's ArenaConfig::noBallRot differs from the batch's");
	if ((int)img.cars.size() != P.nCars) throw std::runtime_error("Arena::Deserialize(): the stream has " + std::to_string(img.cars.size()) + " cars, the batch has " + std::to_string(P.nCars) + " car slots");
	if (memcmp(&img.mut, &b->mutApi, offsetof(RsbMutatorConfig, _pad2)) != 0) {
		RsbMutatorConfig x = img.mut, y = b->mutApi; x._pad0 = y._pad0 = 0; x._pad1[0] = x._pad1[1] = y._pad1[0] = y._pad1[1] = 0;
		if (memcmp(&x, &y, offsetof(RsbMutatorConfig, _pad2)) != 0) throw std::runtime_error("Arena::Deserialize(): the stream's mutator config differs from the batch's");
	}
	std::vector<RsbCarState> st(P.nCars); std::vector<RsbCarControls> ctl(P.nCars); std::vector<uint32_t> ai(std::max(1, P.nCars), arena), ci(std::max(1, P.nCars));
	for (const auto& c : img.cars) {  // the stream lists cars in any order (unordered_set); ids are the slots + 1
		if (c.id < 1 || c.id > (uint32_t)P.nCars) throw std::runtime_error("Arena::Deserialize(): car ID " + std::to_string(c.id) + " has no slot in the batch");
		const int slot = (int)c.id - 1;
		if (c.team != P.team[slot]) throw std::runtime_error("Arena::Deserialize(): team of car " + std::to_string(c.id) + " differs from the batch's slot");
		const RsbCarConfig& k = b->carCfgApi[slot];
		if (memcmp(&c.config.hitboxSize, &k.hitboxSize, 12) || memcmp(&c.config.hitboxPosOffset, &k.hitboxPosOffset, 12) || c.config.dodgeDeadzone != k.dodgeDeadzone ||
		    c.config.frontWheels.wheelRadius != k.frontWheels.wheelRadius || c.config.backWheels.wheelRadius != k.backWheels.wheelRadius ||
		    c.config.frontWheels.suspensionRestLength != k.frontWheels.suspensionRestLength || c.config.backWheels.suspensionRestLength != k.backWheels.suspensionRestLength ||
		    memcmp(&c.config.frontWheels.connectionPointOffset, &k.frontWheels.connectionPointOffset, 12) || memcmp(&c.config.backWheels.connectionPointOffset, &k.backWheels.connectionPointOffset, 12))
			throw std::runtime_error("Arena::Deserialize(): config of car " + std::to_string(c.id) + " differs from the batch's slot");
		st[slot] = c.state; ctl[slot] = c.controls; ci[slot] = (uint32_t)slot;
		// Arena::DeserializeNewCar runs Car::_BulletSetup AFTER the state was read, and _BulletSetup ends with
		// `_internalState.boost = carSpawnBoostAmount` (Car.cpp:291): a loaded car always starts with the spawn boost
		st[slot].boost = b->mutApi.carSpawnBoostAmount;
	}
	if (P.nCars) {
		if (rsb_set_car_states(b, ai.data(), ci.data(), P.nCars, st.data()) != 0) throw std::runtime_error(g_err);  // Car::SetState(_internalState)
		b->ensure_scratch(sizeof(RsbCarControls) * P.nCars);
		h2d(b->dScratch, ctl.data(), sizeof(RsbCarControls) * P.nCars, b->stream);
		LAUNCH_1D(set_controls_kernel, P.nCars, b, b->dS + arena, (const RsbCarControls*)b->dScratch, 1, P.nCars);
	}
	if (P.nPads && rsb_set_pad_states(b, ai.data(), 1, img.pads.data()) != 0) throw std::runtime_error(g_err);
	if (rsb_set_ball_states(b, ai.data(), 1, &img.ball) != 0) throw std::runtime_error(g_err);
	LAUNCH_1D(fresh_world_kernel, 1, b, b->dS + arena, b->dP, b->dScene, img.tickCount);
	API_END
}
int rsb_sync(RsbBatch* b) {
	API_BEGIN
	set_device(b);
#ifndef RSB_HOSTCHECK
	CUDA_OK(cudaStreamSynchronize(b->stream));
#endif
	API_END
}

#define GET_IMPL(kern, T, count, ...) \
	API_BEGIN set_device(b); size_t _n = (count); if (_n == 0) return 0; b->ensure_scratch(_n * sizeof(T)); \
	LAUNCH_1D(kern, (int)_n, b, __VA_ARGS__); d2h(out, b->dScratch, _n * sizeof(T), b->stream); API_END

int rsb_get_car_states(RsbBatch* b, RsbCarState* out) { GET_IMPL(get_car_states_kernel, RsbCarState, (size_t)b->hp.nArenas * b->hp.nCars, b->dS, (RsbCarState*)b->dScratch, b->hp.nArenas, b->hp.nCars) }
int rsb_get_ball_states(RsbBatch* b, RsbBallState* out) { GET_IMPL(get_ball_states_kernel, RsbBallState, (size_t)b->hp.nArenas, b->dS, (RsbBallState*)b->dScratch, b->hp.nArenas, b->hp.mode == DROPSHOT) }
int rsb_get_dropshot_tiles(RsbBatch* b, uint8_t* out) {
	if (b->hp.mode != DROPSHOT) return fail("rsb_get_dropshot_tiles: the batch is not a Dropshot batch");
	GET_IMPL(get_tiles_kernel, uint8_t, (size_t)b->hp.nArenas * MAX_TILES, b->dS, (uint8_t*)b->dScratch, b->hp.nArenas)
}
int rsb_get_pad_states(RsbBatch* b, RsbBoostPadState* out) { GET_IMPL(get_pad_states_kernel, RsbBoostPadState, (size_t)b->hp.nArenas * b->hp.nPads, b->dS, (RsbBoostPadState*)b->dScratch, b->hp.nArenas, b->hp.nPads) }
int rsb_get_car_internals(RsbBatch* b, RsbCarInternal* out) { GET_IMPL(get_car_internals_kernel, RsbCarInternal, (size_t)b->hp.nArenas * b->hp.nCars, b->dS, (RsbCarInternal*)b->dScratch, b->hp.nArenas, b->hp.nCars) }
int rsb_get_ball_internals(RsbBatch* b, RsbBallInternal* out) { GET_IMPL(get_ball_internals_kernel, RsbBallInternal, (size_t)b->hp.nArenas, b->dS, (RsbBallInternal*)b->dScratch, b->hp.nArenas) }
int rsb_observation_floats(RsbBatch* b) { return 9 + 17 * b->hp.nCars; }
int rsb_get_observations(RsbBatch* b, float* out) {
	API_BEGIN set_device(b);
	size_t n = (size_t)b->hp.nArenas * (9 + 17 * b->hp.nCars);
	b->ensure_scratch(n * 4);
	LAUNCH_1D(get_obs_kernel, b->hp.nArenas, b, b->dS, (float*)b->dScratch, b->hp.nArenas, b->hp.nCars);
	d2h(out, b->dScratch, n * 4, b->stream);
	API_END
}

static void misc_get(RsbBatch* b, uint64_t* ticks, uint8_t* scored, uint32_t* errs) {
	set_device(b);
	int n = b->hp.nArenas;
	b->ensure_scratch((size_t)n * 16);
	uint64_t* dt = (uint64_t*)b->dScratch; uint32_t* de = (uint32_t*)((char*)b->dScratch + (size_t)n * 8); uint8_t* ds = (uint8_t*)((char*)b->dScratch + (size_t)n * 12);
	LAUNCH_1D(get_misc_kernel, n, b, b->dS, b->dP, b->dScene, ticks ? dt : nullptr, scored ? ds : nullptr, errs ? de : nullptr, n);
	if (ticks) d2h(ticks, dt, (size_t)n * 8, b->stream);
	if (errs) d2h(errs, de, (size_t)n * 4, b->stream);
	if (scored) d2h(scored, ds, (size_t)n, b->stream);
}
int rsb_get_tick_counts(RsbBatch* b, uint64_t* out) { API_BEGIN misc_get(b, out, nullptr, nullptr); API_END }
int rsb_is_ball_scored(RsbBatch* b, uint8_t* out) { API_BEGIN misc_get(b, nullptr, out, nullptr); API_END }
int rsb_get_error_flags(RsbBatch* b, uint32_t* out) { API_BEGIN misc_get(b, nullptr, nullptr, out); API_END }

int rsb_set_car_states(RsbBatch* b, const uint32_t* ai, const uint32_t* ci, int n, const RsbCarState* st) {
	API_BEGIN set_device(b); if (n <= 0) return 0; check_idx(b, ai, ci, n);
	uint32_t *dai, *dci; RsbCarState* dp;
	stage_set(b, ai, ci, n, st, (size_t)n, &dai, &dci, &dp);
	LAUNCH_1D(set_car_states_kernel, n, b, b->dS, dai, dci, dp, n);
	API_END
}
int rsb_set_ball_states(RsbBatch* b, const uint32_t* ai, int n, const RsbBallState* st) {
	API_BEGIN set_device(b); if (n <= 0) return 0; check_idx(b, ai, nullptr, n);
	uint32_t *dai, *dci; RsbBallState* dp;
	stage_set(b, ai, (const uint32_t*)nullptr, n, st, (size_t)n, &dai, &dci, &dp);
	LAUNCH_1D(set_ball_states_kernel, n, b, b->dS, dai, dp, n, b->hp.mode == DROPSHOT);
	API_END
}
int rsb_set_dropshot_tiles(RsbBatch* b, const uint32_t* ai, int n, const uint8_t* st) {
	API_BEGIN set_device(b);
	if (b->hp.mode != DROPSHOT) throw std::runtime_error("rsb_set_dropshot_tiles: the batch is not a Dropshot batch");
	if (n <= 0) return 0; check_idx(b, ai, nullptr, n);
	for (size_t i = 0; i < (size_t)n * MAX_TILES; i++) if (st[i] > 2) throw std::runtime_error("rsb_set_dropshot_tiles: damage state must be 0 (full), 1 (damaged) or 2 (broken)");
	uint32_t *dai, *dci; uint8_t* dp;
	stage_set(b, ai, (const uint32_t*)nullptr, n, st, (size_t)n * MAX_TILES, &dai, &dci, &dp);
	LAUNCH_1D(set_tiles_kernel, n * MAX_TILES, b, b->dS, dai, dp, n);
	API_END
}
int rsb_set_pad_states(RsbBatch* b, const uint32_t* ai, int n, const RsbBoostPadState* st) {
	API_BEGIN set_device(b); if (n <= 0 || b->hp.nPads == 0) return 0; check_idx(b, ai, nullptr, n);
	uint32_t *dai, *dci; RsbBoostPadState* dp;
	stage_set(b, ai, (const uint32_t*)nullptr, n, st, (size_t)n * b->hp.nPads, &dai, &dci, &dp);
	LAUNCH_1D(set_pad_states_kernel, n * b->hp.nPads, b, b->dS, dai, dp, n, b->hp.nPads);
	API_END
}
int rsb_set_car_internals(RsbBatch* b, const uint32_t* ai, const uint32_t* ci, int n, const RsbCarInternal* in) {
	API_BEGIN set_device(b); if (n <= 0) return 0; check_idx(b, ai, ci, n);
	uint32_t *dai, *dci; RsbCarInternal* dp;
	stage_set(b, ai, ci, n, in, (size_t)n, &dai, &dci, &dp);
	LAUNCH_1D(set_car_internals_kernel, n, b, b->dS, dai, dci, dp, n);
	API_END
}
int rsb_set_ball_internals(RsbBatch* b, const uint32_t* ai, int n, const RsbBallInternal* in) {
	API_BEGIN set_device(b); if (n <= 0) return 0; check_idx(b, ai, nullptr, n);
	uint32_t *dai, *dci; RsbBallInternal* dp;
	stage_set(b, ai, (const uint32_t*)nullptr, n, in, (size_t)n, &dai, &dci, &dp);
	LAUNCH_1D(set_ball_internals_kernel, n, b, b->dS, dai, dp, n);
	API_END
}
int rsb_seed_rng(RsbBatch* b, uint32_t arena, uint32_t seed) {
	API_BEGIN set_device(b);
	if (arena >= (uint32_t)b->hp.nArenas) throw std::runtime_error("arena index out of range");
	LAUNCH_1D(seed_rng_kernel, 1, b, b->dS, arena, seed);
	API_END
}

int rsb_get_events(RsbBatch* b, RsbEvent* out, int cap, int* n_out) {
	API_BEGIN set_device(b);
	// events already downloaded but not yet handed out (the caller's buffer was smaller than what was pending) come first
	uint32_t cnt = 0;
	d2h(&cnt, b->dEventCount, 4, b->stream);
	uint32_t n = std::min(cnt, b->eventCap);
	if (n) {
		std::vector<EventSink::Ev> ev(n);
		d2h(ev.data(), b->dEvents, n * sizeof(EventSink::Ev), b->stream);
		uint32_t zero = 0;
		h2d(b->dEventCount, &zero, 4, b->stream);
		std::stable_sort(ev.begin(), ev.end(), [](const EventSink::Ev& x, const EventSink::Ev& y) { return x.arena != y.arena ? x.arena < y.arena : x.tick < y.tick; });
		b->pendingEvents.insert(b->pendingEvents.end(), ev.begin(), ev.end());
	}
	const size_t m = std::min<size_t>(b->pendingEvents.size(), (size_t)std::max(cap, 0));
	for (size_t i = 0; i < m; i++) { const EventSink::Ev& e = b->pendingEvents[i]; out[i].arena = e.arena; out[i].type = e.type; out[i].a = e.a; out[i].b = e.b; out[i].tick = e.tick; }
	b->pendingEvents.erase(b->pendingEvents.begin(), b->pendingEvents.begin() + (long)m);
	if (n_out) *n_out = (int)m;
	if (!b->pendingEvents.empty()) { g_err = std::to_string(b->pendingEvents.size()) + " more events pending: call rsb_get_events again"; return 1; }
	API_END
}

int rsb_reset_to_random_kickoff(RsbBatch* b, const uint32_t* ai, int n, const int32_t* seeds) {
	API_BEGIN set_device(b);
	if (!ai) n = b->hp.nArenas; else check_idx(b, ai, nullptr, n);
	if (n <= 0) return 0;
	size_t idxBytes = ((size_t)n * 4 + 15) & ~(size_t)15;
	b->ensure_scratch(2 * idxBytes);
	uint32_t* dai = (uint32_t*)b->dScratch; int32_t* dseed = (int32_t*)((char*)b->dScratch + idxBytes);
	if (ai) h2d(dai, ai, (size_t)n * 4, b->stream);
	h2d(dseed, seeds, (size_t)n * 4, b->stream);
	LAUNCH_1D(reset_kickoff_kernel, n, b, b->dS, b->dP, ai ? dai : nullptr, dseed, n);
	API_END
}

int rsb_bench_fill_controls_device(RsbBatch* b, void* d_seq, int n_ticks, uint32_t tick0, uint64_t seed) {
	API_BEGIN set_device(b);
	LAUNCH_1D(fill_controls_kernel, b->hp.nArenas * b->hp.nCars, b, (CtrlS*)d_seq, b->hp.nArenas, b->hp.nCars, n_ticks, tick0, seed);
	API_END
}

// ---- debug probe: wheel values of the last tick (layout of oracle ref_get_wheel_debug) ----
int rsb_debug_get_wheels(RsbBatch* b, int arena, float* out) {
	API_BEGIN set_device(b);
	std::vector<ArenaH> w(1);
	d2h(w.data(), b->dH + arena, sizeof(ArenaH), b->stream);
	for (int k = 0; k < b->hp.nCars * 4; k++) {
		const WheelW& x = w[0].wheels[k]; float* p = out + 14 * k;
		p[0] = (float)x.isInContact; p[1] = (float)x.isInContactWithWorld; p[2] = x.suspLen;
		p[3] = x.contactPointWS.x; p[4] = x.contactPointWS.y; p[5] = x.contactPointWS.z;
		p[6] = x.contactNormalWS.x; p[7] = x.contactNormalWS.y; p[8] = x.contactNormalWS.z;
		p[9] = x.suspRelVel; p[10] = x.clippedInv; p[11] = x.impulse.x; p[12] = x.impulse.y; p[13] = x.impulse.z;
	}
	API_END
}

// ---- debug probe: contact points of the last tick (layout of oracle ref_get_manifolds) ----
// The debug probes read the global mirror of the per-tick scratch, which the kernel keeps complete only for
// single-tick launches (rsb_step(b, 1) / a 1-tick control sequence).
struct DebugSnap { std::vector<ArenaH> h; std::vector<ArenaW> w; };
static DebugSnap debug_snap(RsbBatch* b, int arena) {
	if (arena < 0 || arena >= b->hp.nArenas) throw std::runtime_error("arena index out of range");
	DebugSnap s; s.h.resize(1); s.w.resize(1);
	d2h(s.h.data(), b->dH + arena, sizeof(ArenaH), b->stream);
	d2h(s.w.data(), b->dW + arena, sizeof(ArenaW), b->stream);
	return s;
}
static SolverView debug_solver_view(DebugSnap& s, int nCars) {
	SolverView v;
	v.cnt = (int*)s.h[0].solv; v.sbody = (SolverBody*)(s.h[0].solv + 16);
	v.rows = (SolverRow*)(v.sbody + (nCars + 2)); v.frows = v.rows + SOLVER_ROWS_FAST;
	v.orows = s.w[0].rows; v.ofrows = s.w[0].frows;
	return v;
}

int rsb_debug_get_manifolds(RsbBatch* b, int arena, float* out, int cap_points, int* n_out) {
	API_BEGIN set_device(b);
	DebugSnap s = debug_snap(b, arena);
	const SceneHost& sh = get_scene_dev(b->device, b->hp.mode)->host;
	const int nStatics = (int)sh.meshes.size() + (int)sh.planes.size();
	auto world_index = [&](int obj) { return obj >= STATIC_BASE ? obj - STATIC_BASE : nStatics + obj; };
	int n = 0;
	for (int m = 0; m < s.h[0].pairs.nManifolds; m++) {
		const Manifold& mf = s.h[0].mf[m];
		if (mf.body0 < 0) continue;
		for (int k = 0; k < mf.nPts && n < cap_points; k++, n++) {
			const ContactPt& p = s.w[0].points[m * 4 + k]; float* o = out + 16 * n;
			o[0] = (float)world_index(mf.body0); o[1] = (float)world_index(mf.body1); o[2] = (float)k; o[3] = p.dist;
			o[4] = p.normalB.x; o[5] = p.normalB.y; o[6] = p.normalB.z; o[7] = p.posA.x; o[8] = p.posA.y; o[9] = p.posA.z;
			o[10] = p.posB.x; o[11] = p.posB.y; o[12] = p.posB.z; o[13] = p.friction; o[14] = p.restitution; o[15] = (float)p.isSpecial;
		}
	}
	if (n_out) *n_out = n;
	API_END
}

// ---- debug probe: solver rows of the last tick (layout of oracle ref_get_solver_rows) ----
int rsb_debug_get_solver_rows(RsbBatch* b, int arena, int which, float* out, int cap, int* n_out) {
	API_BEGIN set_device(b);
	DebugSnap s = debug_snap(b, arena);
	int n = 0;
	if (s.h[0].misc[0]) {  // the general solver ran this tick
		const SolverView sv = debug_solver_view(s, b->hp.nCars);
		n = which ? sv.cnt[2] : sv.cnt[1];
		if (n > cap) n = cap;
		for (int i = 0; i < n; i++) {
			const SolverRow& c = which ? sv.frow(i) : sv.row(i); float* o = out + 24 * i;
			const V3* v[6] = {&c.n1, &c.rp1xn, &c.n2, &c.rp2xn, &c.angA, &c.angB};
			for (int k = 0; k < 6; k++) { o[3 * k] = v[k]->x; o[3 * k + 1] = v[k]->y; o[3 * k + 2] = v[k]->z; }
			o[18] = c.jacDiagInv; o[19] = c.rhs; o[20] = c.rhsPen; o[21] = c.applied; o[22] = c.appliedPush; o[23] = c.friction;
		}
	}
	if (n_out) *n_out = n;
	API_END
}

int rsb_debug_get_solver_bodies(RsbBatch* b, int arena, float* out, int cap, int* n_out) {
	API_BEGIN set_device(b);
	DebugSnap s = debug_snap(b, arena);
	int n = 0;
	if (s.h[0].misc[0]) {
		const SolverView sv = debug_solver_view(s, b->hp.nCars);
		n = sv.cnt[0]; if (n > cap) n = cap;
		for (int i = 0; i < n; i++) {
			const SolverBody& sb = sv.sbody[i]; float* o = out + 12 * i;
			const V3* v[4] = {&sb.lin, &sb.ang, &sb.extF, &sb.extT};
			for (int k = 0; k < 4; k++) { o[3 * k] = v[k]->x; o[3 * k + 1] = v[k]->y; o[3 * k + 2] = v[k]->z; }
		}
	}
	if (n_out) *n_out = n;
	API_END
}

int rsb_debug_get_misc(RsbBatch* b, int arena, int* out4) {
	API_BEGIN set_device(b);
	DebugSnap s = debug_snap(b, arena);
	for (int i = 0; i < 4; i++) out4[i] = s.h[0].misc[i];
	API_END
}

// ---- test hook: usable capacity of the per-block detector queue (1 .. DETECT_QUEUE_CAP).  Results do not depend on it: records
// that do not fit are picked up by the overflow sweep of PH_NARROW_DETECT; tests shrink it to run that sweep on every tick.
int rsb_debug_set_detect_queue_cap(RsbBatch* b, int cap) {
	API_BEGIN
	if (cap < 1 || cap > DETECT_QUEUE_CAP) throw std::runtime_error("detector queue capacity must be in [1, " + std::to_string(DETECT_QUEUE_CAP) + "]");
	b->detectQueueCap = cap;
	API_END
}

// ---- developer profile counters (non-zero only in -DRSB_PROFILE builds) ----
int rsb_debug_profile(uint64_t* out64, int reset) {
	API_BEGIN
	for (int i = 0; i < 64; i++) out64[i] = 0;
#if defined(RSB_PROFILE) && !defined(RSB_HOSTCHECK)
	CUDA_OK(cudaDeviceSynchronize());
	CUDA_OK(cudaMemcpyFromSymbol(out64, g_prof, sizeof(unsigned long long) * 64));
	{ static unsigned long long h2[2][16][2]; CUDA_OK(cudaMemcpyFromSymbol(h2, g_prof2, sizeof(h2)));
	  if (getenv("RSB_PROF2")) for (int ph = 0; ph < 2; ph++) for (int k = 0; k < 16; k++) if (h2[ph][k][1]) printf("prof2 %s slowest-item detector calls %d%s epa %d: n %llu mean cycles %llu\n", ph ? "DYN_SOLVE" : "NARROW_DETECT", k & 7, (k & 7) == 7 ? "+" : "", k >> 3, h2[ph][k][1], h2[ph][k][0] / h2[ph][k][1]);
	  if (reset) { memset(h2, 0, sizeof(h2)); CUDA_OK(cudaMemcpyToSymbol(g_prof2, h2, sizeof(h2))); } }
	{ unsigned long long h3[64]; CUDA_OK(cudaMemcpyFromSymbol(h3, g_subprof, sizeof(h3)));
	  const char* nm[32] = {"DYN dyn pairs", "DYN islands (+has_rows)", "DYN solver: bodies", "DYN solver: rows", "DYN solver: split sweeps", "DYN solver: velocity sweeps", "DYN solver: writeback", "DYN solver: whole",
	                        "WHEEL transforms", "WHEEL cast: statics", "WHEEL cast: dynamics", "WHEEL post", "NS pair loop head", "NS plane pair", "NS mesh pair", "NS 15",
	                        "CAR friction", "CAR update_wheels", "CAR air/jump/flip", "CAR vehicle_second", "CAR boost", "PAIRS cells", "PAIRS statics", "PAIRS dynamics", "24", "25", "26", "27", "28", "29", "30", "31"};
	  if (getenv("RSB_PROF2")) for (int k = 0; k < 32; k++) if (h3[32 + k]) printf("subprof %-28s n %llu mean cycles %llu total Mcycles %llu\n", nm[k], h3[32 + k], h3[k] / h3[32 + k], h3[k] / 1000000);
	  if (reset) { memset(h3, 0, sizeof(h3)); CUDA_OK(cudaMemcpyToSymbol(g_subprof, h3, sizeof(h3))); } }
	if (getenv("RSB_PROF2")) {
		static unsigned long long bc[1024]; CUDA_OK(cudaMemcpyFromSymbol(bc, g_blockCycles, sizeof(bc)));
		unsigned long long mn = ~0ull, mx = 0, sum = 0; int n = 0;
		for (int i = 0; i < 1024; i++) if (bc[i]) { mn = std::min(mn, bc[i]); mx = std::max(mx, bc[i]); sum += bc[i]; n++; }
		if (n) printf("blocks of the last launch: %d, cycles min %llu mean %llu max %llu (max/mean %.3f)\n", n, mn, sum / n, mx, (double)mx * n / (double)sum);
		static unsigned long long bp[256][16]; CUDA_OK(cudaMemcpyFromSymbol(bp, g_blockPhase, sizeof(bp)));
		if (n) {
			std::vector<int> idx; for (int i = 0; i < 256; i++) if (bc[i]) idx.push_back(i);
			std::sort(idx.begin(), idx.end(), [&](int x, int y) { return bc[x] < bc[y]; });
			const int pick[5] = {idx.front(), idx[idx.size() / 4], idx[idx.size() / 2], idx[idx.size() * 3 / 4], idx.back()};
			printf("per-block phase walls (Mcycles since reset) for the fastest / q1 / median / q3 / slowest block of the last launch:\n");
			for (int ph = 0; ph < 16; ph++) { bool any = false; for (int q = 0; q < 5; q++) any |= bp[pick[q]][ph] != 0; if (!any) continue;
				printf("  phase %2d:", ph); for (int q = 0; q < 5; q++) printf(" %8.1f", bp[pick[q]][ph] / 1e6); printf("\n"); }
		}
		if (reset) { memset(bp, 0, sizeof(bp)); CUDA_OK(cudaMemcpyToSymbol(g_blockPhase, bp, sizeof(bp))); }
	}
	if (reset) { unsigned long long z[64] = {0}; CUDA_OK(cudaMemcpyToSymbol(g_prof, z, sizeof(z))); }
#else
	(void)reset;
#endif
	API_END
}

// ---- scene introspection ----
int rsb_scene_num_meshes(int mode) { auto it = g_sceneHost.find(mode); return it == g_sceneHost.end() ? 0 : (int)it->second.meshes.size(); }
int rsb_scene_counts(int mode, int mesh, int* c) {
	auto it = g_sceneHost.find(mode);
	if (it == g_sceneHost.end() || mesh < 0 || mesh >= (int)it->second.meshes.size()) return fail("no such mesh");
	const MeshHost& m = it->second.meshes[mesh];
	c[0] = (int)m.nodes.size(); c[1] = (int)m.hdrs.size(); c[2] = m.nTris(); c[3] = m.nVerts();
	int ni = 0; for (const TriInfo& t : m.info) if (t.flags >= 0) ni++;
	c[4] = ni;
	return 0;
}
int rsb_scene_dump(int mode, int mesh, void* nodes, void* headers, float* quant9, void* triinfo) {
	auto it = g_sceneHost.find(mode);
	if (it == g_sceneHost.end() || mesh < 0 || mesh >= (int)it->second.meshes.size()) return fail("no such mesh");
	const MeshHost& m = it->second.meshes[mesh];
	memcpy(nodes, m.nodes.data(), m.nodes.size() * 16);
	for (size_t i = 0; i < m.hdrs.size(); i++) {
		char* p = (char*)headers + 20 * i;
		memcpy(p, m.hdrs[i].qmin, 6); memcpy(p + 6, m.hdrs[i].qmax, 6); memcpy(p + 12, &m.hdrs[i].root, 4); memcpy(p + 16, &m.hdrs[i].size, 4);
	}
	uint16_t q0[3] = {0, 0, 0}, q1[3] = {1, 1, 1};
	V3 mn = scene_detail::unquantize(m, q0), one = scene_detail::unquantize(m, q1);
	quant9[0] = mn.x; quant9[1] = mn.y; quant9[2] = mn.z; quant9[3] = quant9[4] = quant9[5] = 0.f;
	quant9[6] = one.x - mn.x; quant9[7] = one.y - mn.y; quant9[8] = one.z - mn.z;
	memcpy(triinfo, m.info.data(), m.info.size() * 16);
	return 0;
}
int rsb_scene_cells(int mode, int32_t* cell_start, int cap_start, uint8_t* statics, int cap_statics, int* n_cells, int* n_entries) {
	auto it = g_sceneHost.find(mode);
	if (it == g_sceneHost.end()) return fail("no such scene");
	const SceneHost& s = it->second;
	if (n_cells) *n_cells = (int)s.cellStart.size() - 1;
	if (n_entries) *n_entries = (int)s.cellStatics.size();
	if (cell_start && cap_start >= (int)s.cellStart.size()) memcpy(cell_start, s.cellStart.data(), s.cellStart.size() * 4);
	if (statics && cap_statics >= (int)s.cellStatics.size()) memcpy(statics, s.cellStatics.data(), s.cellStatics.size());
	return 0;
}

}  // extern "C"
