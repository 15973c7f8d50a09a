// Index-grid construction for HNS's dense-leaf sidecar.
//
// Replaces CreateIndexGrid -> nanovdb::tools::cuda::voxelsToGrid<ValueOnIndex> (reference src/Cuda/HNanoSolver.cu:375-390,
// externals/nanovdb/tools/cuda/PointsToGrid.cuh:566-1201). The reference uploads all N voxel coordinates and radix-sorts
// N 64-bit voxel keys twice per cook. Because every leaf of the sidecar is a dense, already ordered 8^3 brick
// (src/Utils/GridBuilder.hpp:156-166,229) the same buffer follows from the L leaf origins alone: the node hierarchy is
// assembled on the host from L keys (L = N/512), written straight into the NanoVDB byte layout, and uploaded once.
// The per-leaf 27-neighbour table that replaces per-voxel ReadAccessor walks in the kernels is filled on the device by
// walking that buffer.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace hns {

// sort key of a leaf: (root tile, upper-node offset, lower-node offset), PointsToGrid.cuh:596-602,640-645
struct LeafKey {
	uint64_t tile;  // 21 bits per axis of the 4096^3 tile, biased by 2^31 so that signed order == unsigned order
	uint32_t node;  // upper offset (15 bits) << 12 | lower offset (12 bits)
};
static inline LeafKey leaf_key(const int32_t* o) {
	const int64_t bias = int64_t(1) << 31;
	LeafKey k;
	k.tile = (uint64_t(uint32_t(int64_t(o[2]) + bias) >> 12)) | (uint64_t(uint32_t(int64_t(o[1]) + bias) >> 12) << 21) |
	         (uint64_t(uint32_t(int64_t(o[0]) + bias) >> 12) << 42);
	const uint32_t up = uint32_t(((o[0] & 4095) >> 7) << 10 | ((o[1] & 4095) >> 7) << 5 | ((o[2] & 4095) >> 7));
	const uint32_t lo = uint32_t(((o[0] & 127) >> 3) << 8 | ((o[1] & 127) >> 3) << 4 | ((o[2] & 127) >> 3));
	k.node = up << 12 | lo;
	return k;
}

template <typename T>
static inline void put(uint8_t* p, T v) {
	std::memcpy(p, &v, sizeof(T));
}

struct BBox {
	int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
	void add(const int32_t* a, const int32_t* b) {
		for (int d = 0; d < 3; ++d) lo[d] = std::min(lo[d], a[d]), hi[d] = std::max(hi[d], b[d]);
	}
	void store(uint8_t* p) const {
		std::memcpy(p, lo, 12);
		std::memcpy(p + 12, hi, 12);
	}
};

// Writes the NanoVDB buffer for L dense leaves into `buf` (zero-initialised, `bytes` long).
static void emit_nanovdb(const int32_t* origins, uint64_t L, const std::vector<LeafKey>& keys, uint64_t T, uint64_t nLower, float voxelSize,
                         uint8_t* buf, uint64_t bytes) {
	using namespace nvdb;
	const uint64_t oTree = kGrid, oRoot = oTree + kTree, oUpper = oRoot + kRoot + kTile * T, oLower = oUpper + kUpper * T,
	               oLeaf = oLower + kLower * nLower;
	const double s = double(voxelSize);
	// GridData
	uint8_t* g = buf;
	put<uint64_t>(g + 0, 0x304244566f6e614eull);           // "NanoVDB0"
	put<uint64_t>(g + 8, ~uint64_t(0));                    // checksum: disabled
	put<uint32_t>(g + 16, (32u << 21) | (7u << 10));       // version 32.7.0
	put<uint32_t>(g + 20, 0x22u);                          // HasBBox | IsBreadthFirst
	put<uint32_t>(g + 24, 0u);
	put<uint32_t>(g + 28, 1u);
	put<uint64_t>(g + 32, bytes);
	uint8_t* m = g + 296;                                  // Map: matF[9] invMatF[9] vecF[3] taperF matD[9] invMatD[9] vecD[3] taperD
	for (int d = 0; d < 3; ++d) {
		put<float>(m + 16 * d, float(s));
		put<float>(m + 36 + 16 * d, 1.0f / float(s));
		put<double>(m + 88 + 32 * d, s);
		put<double>(m + 160 + 32 * d, 1.0 / s);
	}
	put<float>(m + 84, 1.0f);
	put<double>(m + 256, 1.0);
	for (int d = 0; d < 3; ++d) put<double>(g + 608 + 8 * d, s);
	put<uint32_t>(g + 632, 0u);                            // GridClass::Unknown -- voxelsToGrid only tags off-index grids as IndexGrid
	put<uint32_t>(g + 636, 20u);                           // GridType::OnIndex
	put<int64_t>(g + 640, int64_t(bytes));                 // blind meta offset: end of leaves
	put<uint64_t>(g + 656, 1u + 512u * L);                 // value count incl. background
	put<uint64_t>(g + 664, 0x314244566f6e614eull);         // "NanoVDB1"
	// TreeData
	uint8_t* t = buf + oTree;
	put<int64_t>(t + 0, int64_t(oLeaf - oTree));
	put<int64_t>(t + 8, int64_t(oLower - oTree));
	put<int64_t>(t + 16, int64_t(oUpper - oTree));
	put<int64_t>(t + 24, int64_t(oRoot - oTree));
	for (int rep = 0; rep < 2; ++rep) {                    // node counts, then "tile counts" (the builder sets them equal)
		put<uint32_t>(t + 32 + 12 * rep, uint32_t(L));
		put<uint32_t>(t + 36 + 12 * rep, uint32_t(nLower));
		put<uint32_t>(t + 40 + 12 * rep, uint32_t(T));
	}
	put<uint64_t>(t + 56, 512u * L);
	// nodes, walking the sorted leaf list once
	BBox rootBox, upBox, loBox;
	uint8_t *U = nullptr, *Lo = nullptr;
	int64_t iu = -1, il = -1;
	constexpr uint64_t kPrefix = 64ull | 128ull << 9 | 192ull << 18 | 256ull << 27 | 320ull << 36 | 384ull << 45 | 448ull << 54;
	for (uint64_t l = 0; l < L; ++l) {
		const int32_t* o = origins + 3 * l;
		const bool newUpper = l == 0 || keys[l].tile != keys[l - 1].tile;
		const bool newLower = newUpper || (keys[l].node >> 12) != (keys[l - 1].node >> 12);
		if (newLower && Lo) loBox.store(Lo), upBox.add(loBox.lo, loBox.hi);
		if (newUpper && U) upBox.store(U), rootBox.add(upBox.lo, upBox.hi);
		if (newUpper) {
			++iu;
			U = buf + oUpper + kUpper * uint64_t(iu);
			upBox = BBox();
			uint8_t* tile = buf + oRoot + kRoot + kTile * uint64_t(iu);
			const uint32_t tx = uint32_t(o[0]) >> 12, ty = uint32_t(o[1]) >> 12, tz = uint32_t(o[2]) >> 12;
			put<uint64_t>(tile, uint64_t(tz) | uint64_t(ty) << 21 | uint64_t(tx) << 42);
			put<int64_t>(tile + 8, int64_t(U - (buf + oRoot)));
		}
		if (newLower) {
			++il;
			Lo = buf + oLower + kLower * uint64_t(il);
			loBox = BBox();
			const uint32_t uo = keys[l].node >> 12;
			reinterpret_cast<uint64_t*>(U + kUpperChildMask)[uo >> 6] |= uint64_t(1) << (uo & 63);
			put<int64_t>(U + kUpperTable + 8ull * uo, int64_t(Lo - U));
		}
		const uint32_t lo = keys[l].node & 4095u;
		uint8_t* F = buf + oLeaf + kLeaf * l;
		reinterpret_cast<uint64_t*>(Lo + kLowerChildMask)[lo >> 6] |= uint64_t(1) << (lo & 63);
		put<int64_t>(Lo + kLowerTable + 8ull * lo, int64_t(F - Lo));
		std::memcpy(F, o, 12);
		F[12] = F[13] = F[14] = 7;
		F[15] = 0x22;
		std::memset(F + kLeafMask, 0xFF, 64);
		put<uint64_t>(F + kLeafOffset, 1u + 512u * l);
		put<uint64_t>(F + kLeafPrefix, kPrefix);
		const int32_t hi[3] = {o[0] + 7, o[1] + 7, o[2] + 7};
		loBox.add(o, hi);
	}
	if (Lo) loBox.store(Lo), upBox.add(loBox.lo, loBox.hi);
	if (U) upBox.store(U), rootBox.add(upBox.lo, upBox.hi);
	uint8_t* r = buf + oRoot;
	rootBox.store(r);
	put<uint32_t>(r + kRootTableSize, uint32_t(T));
	for (int d = 0; d < 3; ++d) {                          // world bbox = index bbox (inclusive max) x scale
		const double a = double(rootBox.lo[d]) * s, b = double(rootBox.hi[d]) * s;
		put<double>(g + 560 + 8 * d, std::min(a, b));
		put<double>(g + 584 + 8 * d, std::max(a, b));
	}
}

// ---- device kernels -------------------------------------------------------------------------------------------
__global__ void k_neighbor_table(GridView g, int32_t* __restrict__ nbr) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= g.num_leaves * 27u) return;
	const uint32_t leaf = t / 27u, slot = t % 27u;
	const int4 o = g.origin[leaf];
	const int dx = int(slot / 9u) - 1, dy = int((slot / 3u) % 3u) - 1, dz = int(slot % 3u) - 1;
	// int32 wrap-around at the coordinate limits is the same as in nanovdb::Coord arithmetic
	nbr[t] = slot == 13u ? int32_t(leaf) : probe_leaf(g, int(uint32_t(o.x) + uint32_t(dx * 8)), int(uint32_t(o.y) + uint32_t(dy * 8)), int(uint32_t(o.z) + uint32_t(dz * 8)));
}

__global__ void k_get_values(GridView g, const int32_t* __restrict__ ijk, uint64_t n, uint64_t* __restrict__ out) {
	const uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
	if (t >= n) return;
	const int x = ijk[3 * t], y = ijk[3 * t + 1], z = ijk[3 * t + 2];
	const int leaf = probe_leaf(g, x, y, z);
	if (leaf < 0) {
		out[t] = 0;
		return;
	}
	// LeafData<ValueOnIndex>::getValue on the emitted record (mask, mOffset, mPrefixSum)
	const uint8_t* F = g.nvdb + g.off_leaf + nvdb::kLeaf * uint64_t(leaf);
	const uint32_t v = uint32_t((x & 7) << 6 | (y & 7) << 3 | (z & 7));
	uint32_t w = v >> 6;
	const uint64_t word = reinterpret_cast<const uint64_t*>(F + nvdb::kLeafMask)[w], bit = uint64_t(1) << (v & 63);
	if (!(word & bit)) {
		out[t] = 0;
		return;
	}
	uint64_t sum = *reinterpret_cast<const uint64_t*>(F + nvdb::kLeafOffset) + uint64_t(__popcll(word & (bit - 1)));
	if (w--) sum += (*reinterpret_cast<const uint64_t*>(F + nvdb::kLeafPrefix) >> (9u * w)) & 511u;
	out[t] = sum;
}

// ---- device block pool + pinned staging of build_grid ---------------------------------------------------------------------------
static std::mutex g_build_mu, g_pool_mu;
static uint8_t* g_stage = nullptr;
static uint64_t g_stage_bytes = 0;
struct PoolBlock {
	uint8_t* p;
	uint64_t bytes;
	int device;
};
static std::vector<PoolBlock> g_pool;  // at most kPoolMax blocks
constexpr size_t kPoolMax = 4;
static cudaError_t pool_take(int device, uint64_t bytes, uint8_t** p, uint64_t* got) {
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		int best = -1;
		for (size_t i = 0; i < g_pool.size(); ++i)
			if (g_pool[i].device == device && g_pool[i].bytes >= bytes && g_pool[i].bytes <= 2 * bytes + (1u << 20) &&
			    (best < 0 || g_pool[i].bytes < g_pool[size_t(best)].bytes))
				best = int(i);
		if (best >= 0) {
			*p = g_pool[size_t(best)].p, *got = g_pool[size_t(best)].bytes;
			g_pool.erase(g_pool.begin() + best);
			return cudaSuccess;
		}
	}
	*got = bytes;
	return cudaMalloc(reinterpret_cast<void**>(p), bytes);
}
static void pool_give(int device, uint8_t* p, uint64_t bytes) {
	if (!p) return;
	uint8_t* drop = nullptr;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		g_pool.push_back(PoolBlock{p, bytes, device});
		if (g_pool.size() > kPoolMax) drop = g_pool.front().p, g_pool.erase(g_pool.begin());
	}
	if (drop) cudaFree(drop);
}
void release_grid_pool() {
	std::vector<PoolBlock> all;
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		all.swap(g_pool);
	}
	for (auto& b : all) cudaFree(b.p);
}

static int build_grid(const int32_t* origins, uint64_t L, float voxel_size, hns_grid** out) {
	if (!out) return fail(HNS_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if (!(voxel_size > 0.0f)) return fail(HNS_ERR_INVALID_ARGUMENT, "voxelSize must be positive.");
	if (L > 0 && !origins) return fail(HNS_ERR_INVALID_ARGUMENT, "origins is null");
	if (L * 512ull >= (1ull << 32)) return fail(HNS_ERR_UNSUPPORTED, "more than 2^32 voxels (the reference's voxel ids are uint32_t)");
	std::vector<LeafKey> keys(L);
	uint64_t T = 0, nLower = 0;
	for (uint64_t l = 0; l < L; ++l) {
		const int32_t* o = origins + 3 * l;
		if ((o[0] | o[1] | o[2]) & 7) return fail(HNS_ERR_TOPOLOGY, "leaf origin is not a multiple of 8 at leaf " + std::to_string(l));
		keys[l] = leaf_key(o);
		if (l) {
			const LeafKey &a = keys[l - 1], &b = keys[l];
			if (!(a.tile < b.tile || (a.tile == b.tile && a.node < b.node)))
				return fail(HNS_ERR_TOPOLOGY, "leaves are not in strictly increasing NanoVDB order at leaf " + std::to_string(l) +
				                                  " (the reference kernels index the sidecar both by input position and by sorted position, so they require it too)");
		}
		if (l == 0 || keys[l].tile != keys[l - 1].tile) ++T;
		if (l == 0 || keys[l].tile != keys[l - 1].tile || (keys[l].node >> 12) != (keys[l - 1].node >> 12)) ++nLower;
	}
	if (L >= (uint64_t(1) << 24)) return fail(HNS_ERR_UNSUPPORTED, "more than 2^24 leaves (8.6 G voxels): the kernels index half-brick fields with 32 bits");
	auto* g = new hns_grid();
	cudaGetDevice(&g->device);
	g->voxel_size = voxel_size;
	g->num_leaves = L, g->num_lower = nLower, g->num_upper = T;
	g->nvdb_bytes = nvdb::kGrid + nvdb::kTree + nvdb::kRoot + nvdb::kTile * T + nvdb::kUpper * T + nvdb::kLower * nLower + nvdb::kLeaf * L;
	// One device block per grid: [NanoVDB buffer | leaf origins int4[L] | neighbour table int32[L][27]], taken from a small pool of
	// blocks of destroyed grids when one fits (CreateIndexGrid runs on every cook: cudaMalloc / cudaFree of ~20 MB each time, the latter
	// a device-wide synchronisation, were a third of its cost), filled through a pinned staging buffer that is kept as well.
	const uint64_t off_origin = (g->nvdb_bytes + 255) & ~uint64_t(255), off_nbr = off_origin + ((L * sizeof(int4) + 255) & ~uint64_t(255));
	const uint64_t block_bytes = off_nbr + std::max<uint64_t>(L, 1) * 27 * sizeof(int32_t);
	auto cleanup = [&](cudaError_t e, const char* what) {
		const std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
		hns_grid_destroy(g);
		return fail(HNS_ERR_CUDA, msg);
	};
	cudaError_t e;
	std::lock_guard<std::mutex> lock(g_build_mu);  // the staging buffer is shared; building is host-bound anyway
	if ((e = pool_take(g->device, block_bytes, &g->d_block, &g->block_bytes)) != cudaSuccess) return cleanup(e, "cudaMalloc(index grid)");
	const uint64_t stage_bytes = off_origin + L * sizeof(int4);
	if (g_stage_bytes < stage_bytes) {
		if (g_stage) cudaFreeHost(g_stage);
		g_stage = nullptr, g_stage_bytes = 0;
		if ((e = cudaMallocHost(reinterpret_cast<void**>(&g_stage), stage_bytes + stage_bytes / 4)) != cudaSuccess) return cleanup(e, "cudaMallocHost(staging)");
		g_stage_bytes = stage_bytes + stage_bytes / 4;
	}
	std::memset(g_stage, 0, off_origin);
	emit_nanovdb(origins, L, keys, T, nLower, voxel_size, g_stage, g->nvdb_bytes);
	int4* org = reinterpret_cast<int4*>(g_stage + off_origin);
	for (uint64_t l = 0; l < L; ++l) org[l] = make_int4(origins[3 * l], origins[3 * l + 1], origins[3 * l + 2], 0);
	g->d_nvdb = g->d_block;
	g->d_origin = L ? reinterpret_cast<int4*>(g->d_block + off_origin) : nullptr;
	g->d_nbr = L ? reinterpret_cast<int32_t*>(g->d_block + off_nbr) : nullptr;
	if ((e = cudaMemcpyAsync(g->d_block, g_stage, stage_bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess) return cleanup(e, "cudaMemcpy(index grid)");
	g->view.nvdb = g->d_nvdb;
	g->view.origin = g->d_origin;
	g->view.nbr = g->d_nbr;
	g->view.num_leaves = uint32_t(L);
	g->view.num_tiles = uint32_t(T);
	g->view.off_root = nvdb::kGrid + nvdb::kTree;
	g->view.off_leaf = g->nvdb_bytes - nvdb::kLeaf * L;
	g->view.list = nullptr;
	g->view.num_list = 0;
	g->view.list_nbr = nullptr;
	if (L) {
		const uint32_t n = uint32_t(L) * 27u;
		HNS_LAUNCH(k_neighbor_table, (n + 255) / 256, 256, 0, 0, g->view, g->d_nbr);
	}
	if ((e = cudaStreamSynchronize(0)) != cudaSuccess) return cleanup(e, "index grid upload / neighbour table kernel");  // the staging buffer is free again
	*out = g;
	return HNS_OK;
}

}  // namespace hns

using namespace hns;

extern "C" {

int hns_grid_create_from_origins(const int32_t* origins, uint64_t n_leaves, float voxel_size, hns_grid** out) {
	return build_grid(origins, n_leaves, voxel_size, out);
}

int hns_grid_create_from_coords(const int32_t* coords, uint64_t n_voxels, float voxel_size, int validate, hns_grid** out) {
	if (!out) return fail(HNS_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if (n_voxels % 512) return fail(HNS_ERR_TOPOLOGY, "voxel count is not a multiple of 512 (dense 8^3 leaves expected, GridBuilder.hpp:229)");
	if (n_voxels && !coords) return fail(HNS_ERR_INVALID_ARGUMENT, "coords is null");
	const uint64_t L = n_voxels / 512;
	std::vector<int32_t> origins(3 * L);
	for (uint64_t l = 0; l < L; ++l) std::memcpy(&origins[3 * l], coords + 3 * 512 * l, 12);
	if (validate) {
		// validate == 1: every coordinate; otherwise a spot check of eight voxels per block that pins the stride of each axis and the last
		// voxel (0, 1, 7, 8, 63, 64, 448, 511) -- O(leaves) instead of O(voxels), cheap enough to run on every cook (compat/hns_compat.cu)
		static const int kSpot[8] = {0, 1, 7, 8, 63, 64, 448, 511};
		for (uint64_t l = 0; l < L; ++l) {
			const int32_t* o = &origins[3 * l];
			const int32_t* c = coords + 3 * 512 * l;
			const int n = validate == 1 ? 512 : 8;
			for (int q = 0; q < n; ++q) {
				const int j = validate == 1 ? q : kSpot[q];
				if (c[3 * j] != o[0] + (j >> 6) || c[3 * j + 1] != o[1] + ((j >> 3) & 7) || c[3 * j + 2] != o[2] + (j & 7))
					return fail(HNS_ERR_TOPOLOGY, "coords block " + std::to_string(l) + " is not a dense leaf in offset order");
			}
		}
	}
	return build_grid(origins.data(), L, voxel_size, out);
}

void hns_grid_destroy(hns_grid* g) {
	if (!g) return;
	pool_give(g->device, g->d_block, g->block_bytes);  // kept for the next grid of similar size (hns_release_scratch frees the pool)
	delete g;
}

uint64_t hns_grid_num_leaves(const hns_grid* g) { return g ? g->num_leaves : 0; }
uint64_t hns_grid_num_voxels(const hns_grid* g) { return g ? g->num_leaves * 512 : 0; }
float hns_grid_voxel_size(const hns_grid* g) { return g ? g->voxel_size : 0.f; }
uint64_t hns_grid_nanovdb_bytes(const hns_grid* g) { return g ? g->nvdb_bytes : 0; }
const void* hns_grid_nanovdb_device_ptr(const hns_grid* g) { return g ? g->d_nvdb : nullptr; }

int hns_grid_nanovdb_download(const hns_grid* g, void* dst) {
	if (!g || !dst) return fail(HNS_ERR_INVALID_ARGUMENT, "null argument");
	HNS_CUDA(cudaMemcpy(dst, g->d_nvdb, g->nvdb_bytes, cudaMemcpyDeviceToHost));
	return HNS_OK;
}

int hns_grid_get_values(const hns_grid* g, const int32_t* ijk, uint64_t n, uint64_t* out) {
	if (!g || (n && (!ijk || !out))) return fail(HNS_ERR_INVALID_ARGUMENT, "null argument");
	if (!n) return HNS_OK;
	int32_t* d_ijk = nullptr;
	uint64_t* d_out = nullptr;
	HNS_CUDA(cudaMalloc(&d_ijk, n * 12));
	HNS_CUDA(cudaMalloc(&d_out, n * 8));
	HNS_CUDA(cudaMemcpy(d_ijk, ijk, n * 12, cudaMemcpyHostToDevice));
	HNS_LAUNCH(k_get_values, unsigned((n + 255) / 256), 256, 0, 0, g->view, d_ijk, n, d_out);
	HNS_CUDA(cudaMemcpy(out, d_out, n * 8, cudaMemcpyDeviceToHost));
	cudaFree(d_ijk);
	cudaFree(d_out);
	return HNS_OK;
}

int hns_grid_neighbors_download(const hns_grid* g, int32_t* dst) {
	if (!g || !dst) return fail(HNS_ERR_INVALID_ARGUMENT, "null argument");
	if (g->num_leaves) HNS_CUDA(cudaMemcpy(dst, g->d_nbr, g->num_leaves * 27 * sizeof(int32_t), cudaMemcpyDeviceToHost));
	return HNS_OK;
}

}  // extern "C"
