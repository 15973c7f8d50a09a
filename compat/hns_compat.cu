// hns_compat.cu -- the reference's seven extern "C" launcher symbols, re-exported on top of the libhns_b200 C ABI.
//
// This is the reference-side binding a maintainer adds to make the Houdini SOPs (src/SOP/**) or any other caller of
// src/Cuda's launchers use the B200-native path without touching their code: build this file against the SAME headers the
// callers use (the reference's src/Utils/GridData.hpp, its vendored externals/nanovdb, and OpenVDB's <openvdb/Types.h> -- or
// the 12-line POD stand-in when building headless) and link the resulting libhns_compat.so instead of the reference's
// `Kernels` static library (src/Cuda/CMakeLists.txt:6-18). Signatures below are verbatim those of
//   src/Cuda/HNanoSolver.cu:387-396, src/Cuda/Advection.cu:169-175, src/Cuda/PressureProjection.cu:127-135, src/Cuda/Combustion.cu:67-70
// Error behaviour: the C ABI's status codes are turned back into the exception types the reference throws
// (std::invalid_argument for argument validation, std::runtime_error otherwise).
//
// No reference source is copied here: the only reference artefacts involved are its public headers, included at build time.
#include <openvdb/Types.h>

#include <cuda_runtime.h>

#include <deque>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "Utils/GridData.hpp"
#include "nanovdb/GridHandle.h"
#include "nanovdb/NanoVDB.h"
#include "nanovdb/cuda/DeviceBuffer.h"
#include "nanovdb/cuda/GridHandle.cuh"

#include "../include/hns_b200.h"

// layout-identical to the reference's CombustionParams (src/Cuda/Kernels.cuh:6-13; duplicated by the caller in
// src/SOP/HNanoSolver/SOP_HNanoSolver.hpp:21-28)
struct CombustionParams {
	float expansionRate;
	float temperatureRelease;
	float buoyancyStrength;
	float ambientTemp;
	float vorticityScale;
	float factorScale;
};
static_assert(sizeof(CombustionParams) == sizeof(hns_combustion_params), "CombustionParams layout");

using HandleT = nanovdb::GridHandle<nanovdb::cuda::DeviceBuffer>;

namespace {

[[noreturn]] void raise(int rc) {
	const std::string msg = hns_last_error();
	if (rc == HNS_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
	throw std::runtime_error(msg);
}
inline void check(int rc) {
	if (rc != HNS_OK) raise(rc);
}

// CreateIndexGrid hands the caller a nanovdb::GridHandle; Compute_Sim gets it back. The leaf tables that the kernels use live in
// an hns_grid, remembered here by the device address of the NanoVDB buffer (bounded: the SOP builds one grid per cook).
struct Entry {
	const void* key;
	uint64_t n_voxels;
	hns_grid* grid;
};
std::mutex g_mu;
std::deque<Entry> g_cache;

void remember(const void* key, uint64_t n, hns_grid* g) {
	std::lock_guard<std::mutex> lk(g_mu);
	for (auto it = g_cache.begin(); it != g_cache.end();)
		if (it->key == key) {
			hns_grid_destroy(it->grid);
			it = g_cache.erase(it);
		} else {
			++it;
		}
	g_cache.push_back({key, n, g});
	while (g_cache.size() > 4) {
		hns_grid_destroy(g_cache.front().grid);
		g_cache.pop_front();
	}
}
hns_grid* lookup(const void* key, uint64_t n) {
	std::lock_guard<std::mutex> lk(g_mu);
	for (const auto& e : g_cache)
		if (e.key == key && e.n_voxels == n) return e.grid;
	return nullptr;
}

const int32_t* coords_of(const HNS::GridIndexedData& d) { return reinterpret_cast<const int32_t*>(d.pCoords()); }

}  // namespace

extern "C" void CreateIndexGrid(HNS::GridIndexedData& data, HandleT& handle, const float voxelSize) {
	hns_grid* g = nullptr;
	// validate = 2: eight voxels of every 512-block are checked against the dense-brick order the kernels assume (O(leaves)); anything
	// else is refused with a std::runtime_error instead of being simulated wrongly
	check(hns_grid_create_from_coords(coords_of(data), data.size(), voxelSize, 2, &g));
	const uint64_t bytes = hns_grid_nanovdb_bytes(g);
	auto buffer = nanovdb::cuda::DeviceBuffer::create(bytes, nullptr, false);  // device only, like voxelsToGrid (PointsToGrid.cuh:763)
	if (cudaMemcpy(buffer.deviceData(), hns_grid_nanovdb_device_ptr(g), bytes, cudaMemcpyDeviceToDevice) != cudaSuccess) {
		hns_grid_destroy(g);
		throw std::runtime_error("CreateIndexGrid: device copy of the NanoVDB buffer failed");
	}
	handle = HandleT(std::move(buffer));
	remember(handle.deviceData(), data.size(), g);
}

extern "C" void Compute_Sim(HNS::GridIndexedData& data, const HandleT& handle, int iteration, float dt, float voxelSize,
                            const CombustionParams& params, bool hasCollision, const cudaStream_t& stream) {
	// validation order of Compute() (src/Cuda/HNanoSolver.cu:12-83)
	if (voxelSize <= 0.0f) throw std::invalid_argument("voxelSize must be positive.");
	if (dt < 0.0f) throw std::invalid_argument("dt (time step) cannot be negative.");
	if (iteration <= 0) throw std::invalid_argument("Number of pressure iterations must be positive.");
	if (handle.isEmpty()) throw std::invalid_argument("Invalid nanovdb::GridHandle provided (null grid).");
	if (data.size() == 0) return;
	if (!handle.deviceGrid<nanovdb::ValueOnIndex>()) throw std::runtime_error("Failed to get device grid pointer of type ValueOnIndex from handle.");
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity), found " + std::to_string(vec.size()));
	float* vel = reinterpret_cast<float*>(data.pValues<openvdb::Vec3f>(vec[0]));
	const auto names = data.getBlocksOfType<float>();
	if (names.empty()) throw std::runtime_error("No float blocks found in input data.");
	std::vector<const char*> cnames;
	std::vector<float*> ptrs;
	for (const auto& n : names) cnames.push_back(n.c_str()), ptrs.push_back(data.pValues<float>(n));
	hns_grid* g = lookup(handle.deviceData(), data.size());
	hns_grid* temp = nullptr;
	if (!g) {  // a handle that did not come from our CreateIndexGrid (e.g. a real voxelsToGrid): rebuild the leaf tables from the coords
		check(hns_grid_create_from_coords(coords_of(data), data.size(), voxelSize, 2, &temp));
		g = temp;
	}
	hns_combustion_params p{params.expansionRate, params.temperatureRelease, params.buoyancyStrength, params.ambientTemp, params.vorticityScale,
	                        params.factorScale};
	const int rc = hns_compute_sim(g, vel, int(names.size()), cnames.data(), ptrs.data(), iteration, dt, voxelSize, &p, hasCollision ? 1 : 0, stream);
	if (temp) hns_grid_destroy(temp);
	check(rc);
}

extern "C" void AdvectIndexGrid(HNS::GridIndexedData& data, const float dt, const float voxelSize, const cudaStream_t& stream) {
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity)");
	const auto names = data.getBlocksOfType<float>();
	if (names.empty()) throw std::runtime_error("No float blocks found");
	std::vector<float*> ptrs;
	for (const auto& n : names) ptrs.push_back(data.pValues<float>(n));
	check(hns_advect_index_grid(coords_of(data), data.size(), reinterpret_cast<const float*>(data.pValues<openvdb::Vec3f>(vec[0])), int(ptrs.size()),
	                            ptrs.data(), dt, voxelSize, stream));
}

extern "C" void AdvectIndexGridVelocity(HNS::GridIndexedData& data, const float dt, const float voxelSize, const cudaStream_t& stream) {
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity)");
	check(hns_advect_index_grid_velocity(coords_of(data), data.size(), reinterpret_cast<float*>(data.pValues<openvdb::Vec3f>(vec[0])), dt, voxelSize,
	                                     stream));
}

extern "C" void ProjectNonDivergent(HNS::GridIndexedData& data, const size_t iterations, const float voxelSize, const cudaStream_t& stream) {
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity)");
	check(hns_project_non_divergent(coords_of(data), data.size(), reinterpret_cast<float*>(data.pValues<openvdb::Vec3f>(vec[0])), iterations, voxelSize,
	                                stream));
}

extern "C" void Divergence(HNS::GridIndexedData& data, const float voxelSize, const cudaStream_t& stream) {
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity)");
	check(hns_divergence(coords_of(data), data.size(), reinterpret_cast<const float*>(data.pValues<openvdb::Vec3f>(vec[0])),
	                     data.pValues<float>("divergence"), voxelSize, stream));
}

extern "C" void CombustionKernel(HNS::GridIndexedData& data, const HandleT& handle, const float dt, const float voxelSize,
                                 const cudaStream_t& stream) {
	const auto vec = data.getBlocksOfType<openvdb::Vec3f>();
	if (vec.size() != 1) throw std::runtime_error("Expected exactly one Vec3f block (velocity)");
	hns_grid* g = lookup(handle.deviceData(), data.size());
	hns_grid* temp = nullptr;
	if (!g) {
		check(hns_grid_create_from_coords(coords_of(data), data.size(), voxelSize, 2, &temp));
		g = temp;
	}
	const int rc = hns_combustion_kernel(g, reinterpret_cast<float*>(data.pValues<openvdb::Vec3f>(vec[0])), data.size(), dt, voxelSize, stream);
	if (temp) hns_grid_destroy(temp);
	check(rc);
}
