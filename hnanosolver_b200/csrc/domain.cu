// Device-side construction of the simulation domain: which 8^3 leaves the sidecar covers this frame.
//
// Reference: SOP_HNanoSolverVerb::cook builds it on the CPU with OpenVDB every cook (src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:188-199):
//     domain = MaskGrid; domain.topologyUnion(velocity tree)
//     Morphology(domain).dilateVoxels(padding, NN_FACE_EDGE_VERTEX, IGNORE_TILES)
//     if (has_collision && sdf) domain.topologyUnion(sdf tree)
// and IndexGridBuilder then treats every leaf NODE of that tree as a dense brick (src/Utils/GridBuilder.hpp:221-239, LeafManager order).
// So the result is a leaf list: every leaf node of the velocity tree, every leaf an active velocity voxel reaches when it is dilated
// `padding` times with the 26-neighbourhood (= by `padding` in the Chebyshev metric), every leaf node of the SDF tree -- in NanoVDB /
// LeafManager order. With the step itself at a few milliseconds this CPU pass is what a cook would wait for, hence on the device:
//   1. one thread per (velocity leaf, candidate neighbour leaf within ceil(padding / 8) leaves): the candidate is reached iff the leaf's
//      512-bit voxel mask has a bit inside the box of voxels whose +-padding cube overlaps the candidate (three interval tests);
//      reached candidates emit their 63-bit NanoVDB-order key, the others a sentinel
//   2. radix sort (CUB), unique, decode.
// Coordinates are limited to |c| < 2^23 voxels so that root tile, upper and lower offsets of a leaf fit one 64-bit key in NanoVDB order.
// OpenVDB is not vendored with the reference, so this block is pinned to a host restatement (oracle/oracle.py domain_leaves), not to
// OpenVDB itself: PARITY UNPINNED at the OpenVDB boundary.
#include <cub/cub.cuh>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace hns {
namespace {

constexpr int kBias = 1 << 23;
constexpr uint64_t kNoKey = ~uint64_t(0);

// key of the leaf with origin (x, y, z) (multiples of 8): tile (12 bits per axis), upper offset (5 per axis), lower offset (4 per axis)
__host__ __device__ inline uint64_t leaf_key(int x, int y, int z) {
	const uint32_t bx = uint32_t(x + kBias), by = uint32_t(y + kBias), bz = uint32_t(z + kBias);
	const uint64_t tile = (uint64_t(bx >> 12) << 24) | (uint64_t(by >> 12) << 12) | uint64_t(bz >> 12);
	const uint64_t up = (uint64_t((bx >> 7) & 31u) << 10) | (uint64_t((by >> 7) & 31u) << 5) | uint64_t((bz >> 7) & 31u);
	const uint64_t lo = (uint64_t((bx >> 3) & 15u) << 8) | (uint64_t((by >> 3) & 15u) << 4) | uint64_t((bz >> 3) & 15u);
	return (tile << 27) | (up << 12) | lo;
}
__host__ __device__ inline void key_origin(uint64_t k, int* o) {
	const uint32_t lo = uint32_t(k & 4095u), up = uint32_t((k >> 12) & 32767u);
	const uint64_t tile = k >> 27;
	const uint32_t t[3] = {uint32_t((tile >> 24) & 4095u), uint32_t((tile >> 12) & 4095u), uint32_t(tile & 4095u)};
	const uint32_t u[3] = {(up >> 10) & 31u, (up >> 5) & 31u, up & 31u}, l[3] = {(lo >> 8) & 15u, (lo >> 4) & 15u, lo & 15u};
	for (int a = 0; a < 3; ++a) o[a] = int((t[a] << 12) | (u[a] << 7) | (l[a] << 3)) - kBias;
}

// voxels v in [0, 8) of a leaf whose dilation by p reaches the leaf d leaves away along one axis: [lo, hi], empty when lo > hi
__device__ inline void reach(int d, int p, int& lo, int& hi) {
	lo = 0, hi = 7;
	if (d > 0) lo = max(0, 8 * d - p);
	else if (d < 0) hi = min(7, p + 8 * d + 7);
}

// masks: [n][8] words, word x holds bit (y * 8 + z) (NanoVDB / OpenVDB leaf mask order: bit n = x<<6 | y<<3 | z); null = every voxel active
__global__ void __launch_bounds__(256) k_dilate_candidates(const int32_t* __restrict__ origins, const uint64_t* __restrict__ masks, uint64_t n, int padding,
                                                           int R, uint64_t* __restrict__ keys) {
	const int side = 2 * R + 1, per = side * side * side;
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n * uint64_t(per)) return;
	const uint64_t leaf = t / uint64_t(per);
	const int c = int(t % uint64_t(per));
	const int dx = c / (side * side) - R, dy = (c / side) % side - R, dz = c % side - R;
	const int ox = origins[3 * leaf], oy = origins[3 * leaf + 1], oz = origins[3 * leaf + 2];
	bool hit = (dx | dy | dz) == 0;  // the leaf node itself is part of the domain whatever its mask (topologyUnion copies nodes)
	if (!hit) {
		int xl, xh, yl, yh, zl, zh;
		reach(dx, padding, xl, xh), reach(dy, padding, yl, yh), reach(dz, padding, zl, zh);
		if (xl <= xh && yl <= yh && zl <= zh) {
			uint64_t any = 0;
			for (int x = xl; x <= xh; ++x) any |= masks ? masks[leaf * 8 + x] : ~uint64_t(0);
			uint64_t box = 0;
			const uint64_t zbits = (uint64_t(0xff) >> (7 - zh)) & (uint64_t(0xff) << zl) & 0xffu;
			for (int y = yl; y <= yh; ++y) box |= zbits << (8 * y);
			hit = (any & box) != 0;
		}
	}
	const int X = ox + 8 * dx, Y = oy + 8 * dy, Z = oz + 8 * dz;
	const bool in_range = X >= -kBias && X < kBias && Y >= -kBias && Y < kBias && Z >= -kBias && Z < kBias;
	keys[t] = hit && in_range ? leaf_key(X, Y, Z) : kNoKey;
}
__global__ void __launch_bounds__(256) k_leaf_keys(const int32_t* __restrict__ origins, uint64_t n, uint64_t* __restrict__ keys) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t < n) keys[t] = leaf_key(origins[3 * t], origins[3 * t + 1], origins[3 * t + 2]);
}
__global__ void __launch_bounds__(256) k_decode_keys(const uint64_t* __restrict__ keys, uint64_t n, int32_t* __restrict__ origins) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n) return;
	int o[3];
	key_origin(keys[t], o);
	origins[3 * t] = o[0], origins[3 * t + 1] = o[1], origins[3 * t + 2] = o[2];
}

struct DevBuf {
	void* p = nullptr;
	~DevBuf() { cudaFree(p); }
	template <typename T>
	T* as() {
		return static_cast<T*>(p);
	}
};

int check_origins(const int32_t* o, uint64_t n, const char* what) {
	for (uint64_t i = 0; i < n; ++i)
		for (int a = 0; a < 3; ++a) {
			const int32_t c = o[3 * i + a];
			if (c & 7) return fail(HNS_ERR_TOPOLOGY, std::string(what) + ": leaf origin is not a multiple of 8");
			if (c < -kBias + 128 || c >= kBias - 128) return fail(HNS_ERR_UNSUPPORTED, std::string(what) + ": leaf origin beyond +-(2^23 - 128) voxels");
		}
	return HNS_OK;
}

}  // namespace
}  // namespace hns

using namespace hns;

struct hns_domain {
	std::vector<int32_t> origins;  // [L][3], NanoVDB order
};

extern "C" {

void hns_domain_destroy(hns_domain* d) { delete d; }
uint64_t hns_domain_num_leaves(const hns_domain* d) { return d ? d->origins.size() / 3 : 0; }
int hns_domain_origins(const hns_domain* d, int32_t* origins_out) {
	if (!d || (!origins_out && !d->origins.empty())) return fail(HNS_ERR_INVALID_ARGUMENT, "null argument");
	if (!d->origins.empty()) std::memcpy(origins_out, d->origins.data(), d->origins.size() * sizeof(int32_t));
	return HNS_OK;
}
int hns_domain_create_grid(const hns_domain* d, float voxel_size, hns_grid** out) {
	if (!d) return fail(HNS_ERR_INVALID_ARGUMENT, "null domain");
	return hns_grid_create_from_origins(d->origins.data(), d->origins.size() / 3, voxel_size, out);
}

int hns_domain_build(const int32_t* vel_origins, const uint64_t* vel_masks, uint64_t n_vel, int padding, const int32_t* sdf_origins, uint64_t n_sdf,
                     hns_domain** out) {
	if (!out) return fail(HNS_ERR_INVALID_ARGUMENT, "out is null");
	*out = nullptr;
	if ((n_vel && !vel_origins) || (n_sdf && !sdf_origins)) return fail(HNS_ERR_INVALID_ARGUMENT, "null origins");
	if (padding < 0 || padding > 64) return fail(HNS_ERR_INVALID_ARGUMENT, "padding must be in [0, 64] voxels");
	int rc;
	if ((rc = check_origins(vel_origins, n_vel, "velocity topology"))) return rc;
	if ((rc = check_origins(sdf_origins, n_sdf, "sdf topology"))) return rc;
	auto* dom = new hns_domain();
	if (n_vel + n_sdf == 0) {
		*out = dom;
		return HNS_OK;
	}
	const int R = (padding + 7) / 8, side = 2 * R + 1;
	const uint64_t per = uint64_t(side) * side * side, n_keys = n_vel * per + n_sdf;
	DevBuf d_org, d_mask, d_sorg, d_keys, d_sorted, d_unique, d_count, d_tmp, d_out;
	auto bail = [&](cudaError_t e, const char* what) {
		delete dom;
		return fail(HNS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
	};
	cudaError_t e;
#define DOM_CUDA(call) \
	if ((e = (call)) != cudaSuccess) return bail(e, #call)
	DOM_CUDA(cudaMalloc(&d_keys.p, n_keys * 8));
	DOM_CUDA(cudaMalloc(&d_sorted.p, n_keys * 8));
	DOM_CUDA(cudaMalloc(&d_unique.p, n_keys * 8));
	DOM_CUDA(cudaMalloc(&d_count.p, 8));
	if (n_vel) {
		DOM_CUDA(cudaMalloc(&d_org.p, n_vel * 12));
		DOM_CUDA(cudaMemcpy(d_org.p, vel_origins, n_vel * 12, cudaMemcpyHostToDevice));
		if (vel_masks) {
			DOM_CUDA(cudaMalloc(&d_mask.p, n_vel * 64));
			DOM_CUDA(cudaMemcpy(d_mask.p, vel_masks, n_vel * 64, cudaMemcpyHostToDevice));
		}
		const uint64_t threads = n_vel * per;
		HNS_LAUNCH(k_dilate_candidates, unsigned((threads + 255) / 256), 256, 0, 0, d_org.as<int32_t>(), d_mask.as<uint64_t>(), n_vel, padding, R,
		           d_keys.as<uint64_t>());
	}
	if (n_sdf) {
		DOM_CUDA(cudaMalloc(&d_sorg.p, n_sdf * 12));
		DOM_CUDA(cudaMemcpy(d_sorg.p, sdf_origins, n_sdf * 12, cudaMemcpyHostToDevice));
		HNS_LAUNCH(k_leaf_keys, unsigned((n_sdf + 255) / 256), 256, 0, 0, d_sorg.as<int32_t>(), n_sdf, d_keys.as<uint64_t>() + n_vel * per);
	}
	size_t tmp_sort = 0, tmp_unique = 0;
	DOM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, d_keys.as<uint64_t>(), d_sorted.as<uint64_t>(), n_keys, 0, 64));
	DOM_CUDA(cub::DeviceSelect::Unique(nullptr, tmp_unique, d_sorted.as<uint64_t>(), d_unique.as<uint64_t>(), d_count.as<uint64_t>(), n_keys));
	DOM_CUDA(cudaMalloc(&d_tmp.p, std::max(tmp_sort, tmp_unique) + 16));
	DOM_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp.p, tmp_sort, d_keys.as<uint64_t>(), d_sorted.as<uint64_t>(), n_keys, 0, 64));
	DOM_CUDA(cub::DeviceSelect::Unique(d_tmp.p, tmp_unique, d_sorted.as<uint64_t>(), d_unique.as<uint64_t>(), d_count.as<uint64_t>(), n_keys));
	uint64_t n_unique = 0;
	DOM_CUDA(cudaMemcpy(&n_unique, d_count.p, 8, cudaMemcpyDeviceToHost));
	if (n_unique) {  // the sentinel sorts last
		uint64_t last = 0;
		DOM_CUDA(cudaMemcpy(&last, d_unique.as<uint64_t>() + (n_unique - 1), 8, cudaMemcpyDeviceToHost));
		if (last == kNoKey) --n_unique;
	}
	dom->origins.resize(3 * n_unique);
	if (n_unique) {
		DOM_CUDA(cudaMalloc(&d_out.p, n_unique * 12));
		HNS_LAUNCH(k_decode_keys, unsigned((n_unique + 255) / 256), 256, 0, 0, d_unique.as<uint64_t>(), n_unique, d_out.as<int32_t>());
		DOM_CUDA(cudaMemcpy(dom->origins.data(), d_out.p, n_unique * 12, cudaMemcpyDeviceToHost));
	}
#undef DOM_CUDA
	*out = dom;
	return HNS_OK;
}

}  // extern "C"
