// Shared declarations for libhns_b200: error handling, launch accounting, device-side grid view.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <string>

#include "../../include/hns_b200.h"

namespace hns {

// ---- errors -------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define HNS_CUDA(call)                                                                                           \
	do {                                                                                                         \
		cudaError_t e_ = (call);                                                                                 \
		if (e_ != cudaSuccess)                                                                                   \
			return ::hns::fail(HNS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
	} while (0)

void release_grid_pool();  // topology.cu: device blocks of destroyed index grids kept for reuse

// ---- launch accounting (hns_launch_count) ---------------------------------------------------------------
extern std::atomic<uint64_t> g_launches;
#define HNS_LAUNCH(kernel, grid, block, smem, stream, ...)              \
	do {                                                                \
		kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);     \
		::hns::g_launches.fetch_add(1, std::memory_order_relaxed);      \
	} while (0)

// ---- NanoVDB 32.7 ValueOnIndex buffer layout (own POD mirror) -------------------
// [GridData 672][TreeData 64][RootData 96][Tile 32 x T][Upper 270400 x T][Lower 33856 x nLower][Leaf 96 x L]
namespace nvdb {
constexpr uint64_t kGrid = 672, kTree = 64, kRoot = 96, kTile = 32, kUpper = 270400, kLower = 33856, kLeaf = 96;
constexpr uint64_t kUpperChildMask = 4128, kUpperTable = 8256;  // InternalData<.,5>: bbox 0, flags 24, valueMask 32, childMask 4128, stats 8224.., table 8256
constexpr uint64_t kLowerChildMask = 544, kLowerTable = 1088;   // InternalData<.,4>: bbox 0, flags 24, valueMask 32, childMask 544, stats 1056.., table 1088
constexpr uint64_t kLeafMask = 16, kLeafOffset = 80, kLeafPrefix = 88;  // LeafData<ValueOnIndex>: bboxMin 0, bboxDif 12, flags 15, mask 16, mOffset 80, mPrefixSum 88
constexpr uint64_t kRootTableSize = 24;
}  // namespace nvdb

// Device view of an index grid.
struct GridView {
	const uint8_t* nvdb;      // the NanoVDB buffer
	const int4* origin;       // [L] leaf origins (x, y, z, 0)
	const int32_t* nbr;       // [L][27] neighbour leaf ids, -1 = none; slot (dx+1)*9+(dy+1)*3+(dz+1)
	uint32_t num_leaves;
	uint32_t num_tiles;
	uint64_t off_root, off_leaf;  // byte offsets of RootData and of the first leaf inside nvdb
	// optional work list: when set, a launch processes leaves list[0..num_list) instead of 0..num_leaves (sharded runs: owned leaves
	// only, or the boundary / interior split that lets a ghost exchange overlap the interior sweep)
	const int32_t* list;
	uint32_t num_list;
	// optional companion of `list`: the neighbour rows gathered in list order, [num_list][27] with slot 13 = the leaf id itself. One
	// dependent load then yields the leaf id AND its neighbours (list -> nbr -> data would be three hops; this is two, like no list).
	const int32_t* list_nbr;
	// optional (sharded runs): bit 8 is set when a semi-Lagrangian sample leaves the 3x3x3 leaf neighbourhood, i.e. may land beyond
	// the shard's ghost layer where the local tree knows nothing about leaves other ranks own
	uint32_t* far_flag;
	__host__ __device__ uint32_t count() const { return list ? num_list : num_leaves; }
#ifdef __CUDACC__
	__device__ __forceinline__ uint32_t leaf_at(uint32_t i) const { return list ? uint32_t(__ldg(list + i)) : i; }
#endif
};

// ---- L2 residency of the pressure field during the solve ----------------------------------------------------------------------
// The two pressure halves (one allocation) are read by every one of the 2 x iterations half-sweeps; everything else a sweep touches
// streams. An access-policy window marks a fraction of the pressure lines "persisting" in the L2 set-aside for the launches
// enqueued while it is set, so those lines are served from L2 in every sweep instead of from HBM. Scoped: the previous window of
// the caller's stream is put back when the object dies (attributes are captured per launch at enqueue time).
// MEASURED ON B200 AND SWITCHED OFF BY DEFAULT: on the 512^3 workload (161 MB of pressure) the 40-iteration solve takes 4.59 ms
// without a window, 4.77 / 4.80 / 6.50 / 7.54 ms with 32 / 48 / 64 / 80 MB set aside, and a large set-aside also slows the
// streaming kernels around it (the normal L2 shrinks). The knob stays for experiments.
struct L2PressureWindow {
	cudaStream_t st = nullptr;
	bool active = false;
	cudaStreamAttrValue saved{};
	static long long& requested_mb() {  // HNS_L2_PERSIST_MB (default 0 = off, see below); hns_set_l2_persist_mb() overrides it
		static long long mb = [] {
			const char* e = std::getenv("HNS_L2_PERSIST_MB");
			return e ? std::atoll(e) : 0ll;
		}();
		return mb;
	}
	static long long& applied() {
		static long long bytes = -1;  // -1: device limit not set yet
		return bytes;
	}
	static size_t persist_bytes() {  // set-aside size, clamped to the device limit
		long long& cached = applied();
		if (cached < 0) {
			int dev = 0, max_persist = 0;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
			cached = std::max(0ll, std::min<long long>(requested_mb() << 20, max_persist));
			if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(cached)) != cudaSuccess) cached = 0;
			cudaGetLastError();
		}
		return size_t(cached);
	}
	L2PressureWindow(cudaStream_t stream, const void* base, size_t bytes) : st(stream) {
		const size_t persist = persist_bytes();
		if (!persist || !bytes || bytes <= persist / 2) return;  // small fields live in the L2 anyway
		int dev = 0, max_window = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
		if (max_window <= 0) return;
		if (cudaStreamGetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &saved) != cudaSuccess) {
			cudaGetLastError();
			return;
		}
		cudaStreamAttrValue v{};
		v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
		v.accessPolicyWindow.num_bytes = std::min(bytes, size_t(max_window));
		v.accessPolicyWindow.hitRatio = float(std::min(1.0, double(persist) / double(v.accessPolicyWindow.num_bytes)));
		v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
		if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) == cudaSuccess) active = true;
		else cudaGetLastError();
	}
	~L2PressureWindow() {
		if (active) cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &saved);
	}
	L2PressureWindow(const L2PressureWindow&) = delete;
	L2PressureWindow& operator=(const L2PressureWindow&) = delete;
};

// packed field groups of the advection kernels (kernels.cuh, advect.cu)
struct AdvectGroups {
	float4* g[5] = {};
	unsigned valid = 0;
	uint64_t n = 0;  // voxels per group
};

// slot ids of the six face neighbours
constexpr int kSlotXm = 4, kSlotXp = 22, kSlotYm = 10, kSlotYp = 16, kSlotZm = 12, kSlotZp = 14, kSlotSelf = 13;

#ifdef __CUDACC__
// Leaf ordinal containing voxel (x,y,z), or -1: the ReadAccessor walk (root tile scan -> upper -> lower) on the emitted buffer.
__device__ __forceinline__ int probe_leaf(const GridView& g, int x, int y, int z) {
	const uint8_t* root = g.nvdb + g.off_root;
	const uint64_t key = (uint64_t(uint32_t(z) >> 12)) | (uint64_t(uint32_t(y) >> 12) << 21) | (uint64_t(uint32_t(x) >> 12) << 42);
	const uint8_t* upper = nullptr;
	for (uint32_t t = 0; t < g.num_tiles; ++t) {
		const uint8_t* tile = root + nvdb::kRoot + nvdb::kTile * t;
		if (*reinterpret_cast<const uint64_t*>(tile) == key) {
			upper = root + *reinterpret_cast<const int64_t*>(tile + 8);
			break;
		}
	}
	if (!upper) return -1;
	const uint32_t uo = uint32_t(((x & 4095) >> 7) << 10 | ((y & 4095) >> 7) << 5 | ((z & 4095) >> 7));
	if (!((*reinterpret_cast<const uint64_t*>(upper + nvdb::kUpperChildMask + 8 * (uo >> 6)) >> (uo & 63)) & 1)) return -1;
	const uint8_t* lower = upper + *reinterpret_cast<const int64_t*>(upper + nvdb::kUpperTable + 8ull * uo);
	const uint32_t lo = uint32_t(((x & 127) >> 3) << 8 | ((y & 127) >> 3) << 4 | ((z & 127) >> 3));
	if (!((*reinterpret_cast<const uint64_t*>(lower + nvdb::kLowerChildMask + 8 * (lo >> 6)) >> (lo & 63)) & 1)) return -1;
	const uint8_t* leaf = lower + *reinterpret_cast<const int64_t*>(lower + nvdb::kLowerTable + 8ull * lo);
	return int((leaf - (g.nvdb + g.off_leaf)) / nvdb::kLeaf);
}
#endif

}  // namespace hns

struct hns_mg;
struct hns_state;
namespace hns {
// packed advection groups in a sharded frame (api.cu): which groups the kernels of this frame have written for the owned leaves, and
// -- after the ghost exchange that precedes the advection pass -- the same for the listed (ghost) leaves, from the brick fields
struct GroupsCurrent {
	bool g0 = false, g1 = false;
};
GroupsCurrent groups_current(const hns_state* s);
int groups_refresh_leaves(hns_state* s, const int32_t* ids, uint64_t n_ids, GroupsCurrent which, cudaStream_t st);
void groups_stamp(hns_state* s, GroupsCurrent which);
int mg_pressure_solve(hns_state* s, hns_mg* mg, int max_cycles, double rel_tol, int nu_pre, int nu_post, float omega, cudaStream_t st);  // multigrid.cu
}
// ---- opaque handle definitions ----------------------------------------------------------------------------
struct hns_grid {
	int device = 0;
	float voxel_size = 0.f;
	uint64_t num_leaves = 0, num_lower = 0, num_upper = 0;
	uint64_t nvdb_bytes = 0;
	uint8_t* d_block = nullptr;  // one allocation: [NanoVDB buffer | origins | neighbour table] (topology.cu)
	uint64_t block_bytes = 0;
	uint8_t* d_nvdb = nullptr;
	int4* d_origin = nullptr;
	int32_t* d_nbr = nullptr;
	hns::GridView view{};
};

struct hns_state {
	const hns_grid* grid = nullptr;
	uint64_t n = 0;  // voxels
	int n_scalars = 0;
	float* vel[3] = {};   // current velocity, SoA bricks float[L][512]
	float* adv[3] = {};   // advected velocity
	float* div[2] = {};   // divergence, colour-split: [0] red = (x+y+z) even, [1] black; float[L][256] each
	float* p[2] = {};     // pressure, colour-split like div
	float* sc[16] = {};   // scalar fields (current)
	float* sc_out[16] = {};
	float* aos = nullptr; // staging float[N][3] for host <-> device velocity transfers
	uint8_t* cold = nullptr;  // [L] leaf flags of the advection kernels (advect.cu), zero between launches
	// packed groups the advection kernels stage from, allocated on first use, and what they were built from (api.cu, "packed groups")
	hns::AdvectGroups grp;
	uint64_t sc_version = 1;  // bumped by every entry point that (may) overwrite a scalar field, like vel_version for the velocity
	uint64_t grp0_vel_version = 0, grp0_sc_version = 0, grp1_sc_version = 0;  // 0 never matches: the counters start at 1
	const float* grp0_s0 = nullptr;                                          // the scalar buffer in group 0's fourth lane
	const float* grp1_src[4] = {nullptr, nullptr, nullptr, nullptr};         // the buffers group 1 mirrors
	bool grp_dist = false;  // a sharded frame (dist.cu) re-packs the ghost leaves of the groups after its exchanges
	float* vort[4] = {nullptr, nullptr, nullptr, nullptr};  // vorticity confinement scratch, allocated on first use: output planes u, v, w and |curl|
	// optional combustion + buoyancy stage of the all-in-one frame
	bool comb_enabled = false;
	int comb_idx[4] = {-1, -1, -1, -1};  // fuel, waste, temperature, flame
	hns_combustion_params comb{};
	int skip_scalar = -1;  // scalar that is carried but not advected ("collision_sdf")
	bool collision = false;  // hasCollision with collision data: scalar `skip_scalar` is the SDF the kernels collide against
	const float* collision_sdf() const { return collision && skip_scalar >= 0 ? sc[skip_scalar] : nullptr; }
	const float* elem0 = nullptr;  // device float[3 + n_scalars]: element 0 of the global arrays (sharded runs), else null
	uint64_t vel_version = 1;  // bumped by every entry point that (may) overwrite the velocity planes; lets a sharded run skip a ghost exchange of unchanged data
	const int32_t* active = nullptr;  // device list of the leaves the kernels process (sharded runs: the owned leaves), null = all
	uint32_t n_active = 0;
	uint32_t* far_flag = nullptr;  // GridView::far_flag
	// optional multigrid pressure solve of the frame (hns_state_set_pressure_solver); null = the reference's fixed-count red-black SOR
	struct hns_mg* mg = nullptr;
	int mg_cycles = 0, mg_nu[2] = {2, 2};
	float mg_omega = 1.0f;
	double* d_sums = nullptr;  // device double[2] of the norm reductions, allocated on first use
	// The reference's solve is 2 x iterations dependent launches; on grids of a few thousand leaves each one is launch latency, not
	// work (4.4 us per launch for 0.9 us of sweep at 1.2 k leaves). The sequence is captured once into a CUDA graph and replayed.
	struct SolveGraph {
		cudaGraphExec_t exec = nullptr;
		hns::GridView view{};  // everything the captured launches depend on: the grid view (pointers, counts, work list) ...
		const void *p = nullptr, *div = nullptr;  // ... the fields ...
		int iterations = -1;                      // ... and the solve's parameters
		unsigned flags = 0;
		float omega = 0.f, dx = 0.f;
		uint64_t launches = 0;
	} solve_graph;
	cudaStream_t capture_stream = nullptr;
	hns::GridView view() const {
		hns::GridView v = grid->view;
		v.list = active, v.num_list = n_active, v.far_flag = far_flag;
		return v;
	}
};
