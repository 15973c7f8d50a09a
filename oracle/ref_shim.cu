// TEST INFRASTRUCTURE ONLY -- not part of the product; nothing under hnanosolver_b200/ links this.
//
// Plain-C handles around the UNMODIFIED reference launchers so that Python (ctypes) can drive them.
// This file is compiled together with the reference's own translation units, taken where they lie
// under /root/reference/src/Cuda (see oracle/Makefile); outputs go to oracle/_ref/ only.
//
// What it wraps (reference file:line):
//   CreateIndexGrid            src/Cuda/HNanoSolver.cu:387-390
//   Compute_Sim                src/Cuda/HNanoSolver.cu:393-396
//   AdvectIndexGrid            src/Cuda/Advection.cu:169-171
//   AdvectIndexGridVelocity    src/Cuda/Advection.cu:173-175
//   ProjectNonDivergent        src/Cuda/PressureProjection.cu:132-135
//   Divergence                 src/Cuda/PressureProjection.cu:127-129
// plus direct launches of the reference __global__ kernels (src/Cuda/Kernels.cuh:15-104) on device
// arrays, used for stage-by-stage parity and for kernel-only timing of the reference on the same GPU.
#include <openvdb/Types.h>

#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../Utils/GridData.hpp"
#ifndef HNS_SHIM_LAUNCHERS_ONLY
#include "Kernels.cuh"
#else
struct CombustionParams {
	float expansionRate, temperatureRelease, buoyancyStrength, ambientTemp, vorticityScale, factorScale;
};
#endif
#include "nanovdb/GridHandle.h"
#include "nanovdb/NanoVDB.h"
#include "nanovdb/cuda/DeviceBuffer.h"

using HandleT = nanovdb::GridHandle<nanovdb::cuda::DeviceBuffer>;

extern "C" void CreateIndexGrid(HNS::GridIndexedData&, HandleT&, float);
extern "C" void Compute_Sim(HNS::GridIndexedData&, const HandleT&, int, float, float, const CombustionParams&, bool, const cudaStream_t&);
extern "C" void AdvectIndexGrid(HNS::GridIndexedData&, float, float, const cudaStream_t&);
extern "C" void AdvectIndexGridVelocity(HNS::GridIndexedData&, float, float, const cudaStream_t&);
extern "C" void ProjectNonDivergent(HNS::GridIndexedData&, size_t, float, const cudaStream_t&);
extern "C" void Divergence(HNS::GridIndexedData&, float, const cudaStream_t&);

static thread_local std::string g_err;

#define REF_TRY try {
#define REF_CATCH                                          \
	}                                                      \
	catch (const std::exception& e) {                      \
		g_err = e.what();                                  \
		return 1;                                          \
	}                                                      \
	catch (...) {                                          \
		g_err = "unknown exception";                       \
		return 1;                                          \
	}                                                      \
	{                                                      \
		cudaError_t e_ = cudaGetLastError();               \
		if (e_ != cudaSuccess) {                           \
			g_err = std::string("cuda: ") + cudaGetErrorString(e_); \
			return 2;                                      \
		}                                                  \
	}                                                      \
	return 0;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// ---- HNS::GridIndexedData as an opaque handle (reference src/Utils/GridData.hpp:16-166) ----
void* ref_data_create(uint64_t n, int allocType) {
	auto* d = new HNS::GridIndexedData();
	d->setAllocationType(static_cast<AllocationType>(allocType));
	if (!d->allocateCoords(n)) {
		delete d;
		return nullptr;
	}
	return d;
}
void ref_data_destroy(void* d) { delete static_cast<HNS::GridIndexedData*>(d); }
int32_t* ref_data_coords(void* d) { return reinterpret_cast<int32_t*>(static_cast<HNS::GridIndexedData*>(d)->pCoords()); }
uint64_t ref_data_size(void* d) { return static_cast<HNS::GridIndexedData*>(d)->size(); }
float* ref_data_add_float(void* d_, const char* name) {
	auto* d = static_cast<HNS::GridIndexedData*>(d_);
	if (!d->addValueBlock<float>(name, d->size())) return nullptr;
	return d->pValues<float>(name);
}
float* ref_data_add_vec3(void* d_, const char* name) {
	auto* d = static_cast<HNS::GridIndexedData*>(d_);
	if (!d->addValueBlock<openvdb::Vec3f>(name, d->size())) return nullptr;
	return reinterpret_cast<float*>(d->pValues<openvdb::Vec3f>(name));
}
float* ref_data_float(void* d_, const char* name) { return static_cast<HNS::GridIndexedData*>(d_)->pValues<float>(name); }

// ---- index grid ----
int ref_create_index_grid(void* data, float voxelSize, void** outHandle) {
	REF_TRY
	auto* h = new HandleT();
	CreateIndexGrid(*static_cast<HNS::GridIndexedData*>(data), *h, voxelSize);
	cudaDeviceSynchronize();
	*outHandle = h;
	REF_CATCH
}
void ref_grid_destroy(void* h) { delete static_cast<HandleT*>(h); }
uint64_t ref_grid_bytes(void* h) { return static_cast<HandleT*>(h)->buffer().size(); }
const void* ref_grid_device_ptr(void* h) { return static_cast<HandleT*>(h)->deviceData(); }
int ref_grid_download(void* h, void* dst) {
	auto* hh = static_cast<HandleT*>(h);
	return cudaMemcpy(dst, hh->deviceData(), hh->buffer().size(), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

// getValue(ijk) through the real nanovdb::ReadAccessor for a list of coordinates (device side).
__global__ void ref_get_values_kernel(const nanovdb::NanoGrid<nanovdb::ValueOnIndex>* grid, const nanovdb::Coord* ijk, uint64_t* out,
                                      uint64_t n) {
	const uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
	if (t >= n) return;
	auto acc = grid->getAccessor();
	out[t] = acc.getValue(ijk[t]);
}
int ref_grid_get_values(void* h, const int32_t* ijkHost, uint64_t n, uint64_t* outHost) {
	REF_TRY
	auto* grid = static_cast<HandleT*>(h)->deviceGrid<nanovdb::ValueOnIndex>();
	if (!grid) throw std::runtime_error("no device grid");
	nanovdb::Coord* d_ijk;
	uint64_t* d_out;
	cudaMalloc(&d_ijk, n * 12);
	cudaMalloc(&d_out, n * 8);
	cudaMemcpy(d_ijk, ijkHost, n * 12, cudaMemcpyHostToDevice);
	ref_get_values_kernel<<<(n + 255) / 256, 256>>>(grid, d_ijk, d_out, n);
	cudaMemcpy(outHost, d_out, n * 8, cudaMemcpyDeviceToHost);
	cudaFree(d_ijk);
	cudaFree(d_out);
	REF_CATCH
}

// ---- the launchers, in place on the host sidecar exactly like the SOP nodes call them ----
// elapsedMs (optional) = host wall clock around the reference call(s) only.
int ref_compute_sim(void* data, void* handle, int iterations, float dt, float voxelSize, const float* params6, int hasCollision,
                    double* elapsedMs) {
	REF_TRY
	CombustionParams p{params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	const auto t0 = std::chrono::steady_clock::now();
	Compute_Sim(*static_cast<HNS::GridIndexedData*>(data), *static_cast<HandleT*>(handle), iterations, dt, voxelSize, p, hasCollision != 0,
	            stream);
	const auto t1 = std::chrono::steady_clock::now();
	cudaStreamDestroy(stream);
	if (elapsedMs) *elapsedMs = std::chrono::duration<double, std::milli>(t1 - t0).count();
	REF_CATCH
}

// CreateIndexGrid + Compute_Sim back to back, as SOP_HNanoSolverVerb::cook does (src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:231-253).
int ref_cook_frame(void* data, int iterations, float dt, float voxelSize, const float* params6, int hasCollision, double* elapsedMs) {
	REF_TRY
	CombustionParams p{params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
	const auto t0 = std::chrono::steady_clock::now();
	HandleT h;
	CreateIndexGrid(*static_cast<HNS::GridIndexedData*>(data), h, voxelSize);
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	Compute_Sim(*static_cast<HNS::GridIndexedData*>(data), h, iterations, dt, voxelSize, p, hasCollision != 0, stream);
	cudaStreamDestroy(stream);
	const auto t1 = std::chrono::steady_clock::now();
	if (elapsedMs) *elapsedMs = std::chrono::duration<double, std::milli>(t1 - t0).count();
	REF_CATCH
}

int ref_advect_index_grid(void* data, float dt, float voxelSize) {
	REF_TRY
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	AdvectIndexGrid(*static_cast<HNS::GridIndexedData*>(data), dt, voxelSize, stream);
	cudaStreamDestroy(stream);
	REF_CATCH
}
int ref_advect_index_grid_velocity(void* data, float dt, float voxelSize) {
	REF_TRY
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	AdvectIndexGridVelocity(*static_cast<HNS::GridIndexedData*>(data), dt, voxelSize, stream);
	cudaStreamDestroy(stream);
	REF_CATCH
}
int ref_project_non_divergent(void* data, uint64_t iterations, float voxelSize) {
	REF_TRY
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	ProjectNonDivergent(*static_cast<HNS::GridIndexedData*>(data), iterations, voxelSize, stream);
	cudaStreamDestroy(stream);
	REF_CATCH
}
int ref_divergence(void* data, float voxelSize) {
	REF_TRY
	cudaStream_t stream;
	cudaStreamCreate(&stream);
	Divergence(*static_cast<HNS::GridIndexedData*>(data), voxelSize, stream);
	cudaStreamDestroy(stream);
	REF_CATCH
}

#ifndef HNS_SHIM_LAUNCHERS_ONLY
// ---- device-resident "north-star frame" with the reference's own kernels ----
// Same step list as the product's resident frame (advect_vector -> divergence -> I x (red, black) ->
// subtractPressureGradient -> advect_scalars; combustion/buoyancy/vorticity off), launched exactly as
// Compute() launches them (1-D, 256 threads; src/Cuda/HNanoSolver.cu:137-148,164,184,263-268,282,346).
// Buffers live on the device across calls so kernel-only time can be taken with CUDA events.
struct RefFrame {
	uint64_t n = 0;
	int S = 0;
	nanovdb::Coord* coords = nullptr;
	nanovdb::Vec3f *vel = nullptr, *adv = nullptr, *proj = nullptr;
	float *div = nullptr, *p = nullptr;
	float *in[16] = {}, *out[16] = {};
	float **dIn = nullptr, **dOut = nullptr;
	float* sdf = nullptr;  // collision SDF of the collision stages below (allocated on first use)
};

void* ref_frame_create(void* data_, int S, const char** names) {
	auto* data = static_cast<HNS::GridIndexedData*>(data_);
	auto* f = new RefFrame();
	f->n = data->size();
	f->S = S;
	const auto velNames = data->getBlocksOfType<openvdb::Vec3f>();
	cudaMalloc(&f->coords, f->n * 12);
	cudaMalloc(&f->vel, f->n * 12);
	cudaMalloc(&f->adv, f->n * 12);
	cudaMalloc(&f->proj, f->n * 12);
	cudaMalloc(&f->div, f->n * 4);
	cudaMalloc(&f->p, f->n * 4);
	cudaMemcpy(f->coords, data->pCoords(), f->n * 12, cudaMemcpyHostToDevice);
	cudaMemcpy(f->vel, data->pValues<openvdb::Vec3f>(velNames.at(0)), f->n * 12, cudaMemcpyHostToDevice);
	for (int s = 0; s < S; ++s) {
		cudaMalloc(&f->in[s], f->n * 4);
		cudaMalloc(&f->out[s], f->n * 4);
		cudaMemcpy(f->in[s], data->pValues<float>(names[s]), f->n * 4, cudaMemcpyHostToDevice);
	}
	cudaMalloc(&f->dIn, 16 * sizeof(float*));
	cudaMalloc(&f->dOut, 16 * sizeof(float*));
	cudaMemcpy(f->dIn, f->in, 16 * sizeof(float*), cudaMemcpyHostToDevice);
	cudaMemcpy(f->dOut, f->out, 16 * sizeof(float*), cudaMemcpyHostToDevice);
	return f;
}
void ref_frame_destroy(void* f_) {
	auto* f = static_cast<RefFrame*>(f_);
	cudaFree(f->coords), cudaFree(f->vel), cudaFree(f->adv), cudaFree(f->proj), cudaFree(f->div), cudaFree(f->p);
	for (int s = 0; s < f->S; ++s) cudaFree(f->in[s]), cudaFree(f->out[s]);
	cudaFree(f->dIn), cudaFree(f->dOut), cudaFree(f->sdf);
	delete f;
}
// Runs `frames` frames; returns total device ms (CUDA events on the launch stream). State is NOT advanced
// between frames (inputs are re-used) so every frame does identical work.
int ref_frame_run(void* f_, void* handle, int iterations, float dt, float voxelSize, int frames, float* msOut) {
	REF_TRY
	auto* f = static_cast<RefFrame*>(f_);
	auto* grid = static_cast<HandleT*>(handle)->deviceGrid<nanovdb::ValueOnIndex>();
	const float inv = 1.0f / voxelSize;
	const int bs = 256;
	const int gs = int((f->n + bs - 1) / bs);
	const float omega = 2.0f / (1.0f + sinf(static_cast<float>(3.14159) * voxelSize));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	cudaEventRecord(e0, 0);
	for (int it = 0; it < frames; ++it) {
		cudaMemsetAsync(f->p, 0, f->n * 4, 0);
		advect_vector<<<gs, bs>>>(grid, f->coords, f->vel, f->adv, nullptr, false, f->n, dt, inv);
		divergence<<<gs, bs>>>(grid, f->coords, f->adv, f->div, inv, f->n);
		for (int i = 0; i < iterations; ++i) {
			redBlackGaussSeidelUpdate<<<gs, bs>>>(grid, f->coords, f->div, f->p, voxelSize, f->n, 0, omega);
			redBlackGaussSeidelUpdate<<<gs, bs>>>(grid, f->coords, f->div, f->p, voxelSize, f->n, 1, omega);
		}
		subtractPressureGradient<<<gs, bs>>>(grid, f->coords, f->n, f->adv, f->p, f->proj, nullptr, false, inv);
		advect_scalars<<<gs, bs>>>(grid, f->coords, f->proj, f->dIn, f->dOut, f->S, nullptr, false, f->n, dt, inv);
	}
	cudaEventRecord(e1, 0);
	cudaEventSynchronize(e1);
	cudaEventElapsedTime(msOut, e0, e1);
	cudaEventDestroy(e0), cudaEventDestroy(e1);
	REF_CATCH
}
// Download results of the last frame: projected velocity, pressure, divergence, advected velocity, scalars.
int ref_frame_download(void* f_, float* velProj, float* pressure, float* div, float* velAdv, float** scalars) {
	auto* f = static_cast<RefFrame*>(f_);
	if (velProj) cudaMemcpy(velProj, f->proj, f->n * 12, cudaMemcpyDeviceToHost);
	if (pressure) cudaMemcpy(pressure, f->p, f->n * 4, cudaMemcpyDeviceToHost);
	if (div) cudaMemcpy(div, f->div, f->n * 4, cudaMemcpyDeviceToHost);
	if (velAdv) cudaMemcpy(velAdv, f->adv, f->n * 12, cudaMemcpyDeviceToHost);
	if (scalars)
		for (int s = 0; s < f->S; ++s)
			if (scalars[s]) cudaMemcpy(scalars[s], f->out[s], f->n * 4, cudaMemcpyDeviceToHost);
	return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

// The reference's vorticityConfinement kernel (Kernel.cu:969-1025) launched OUT OF PLACE on the frame's velocity (input f->vel, output
// a scratch buffer, downloaded to outHost): Compute() passes one buffer as input and output (HNanoSolver.cu:174), which races; with two buffers
// the same kernel is deterministic, and that is what the oracle's restatement is pinned against.
int ref_frame_vorticity(void* f_, void* handle, float dt, float voxelSize, float scale, float factorScale, float* outHost) {
	REF_TRY
	auto* f = static_cast<RefFrame*>(f_);
	auto* grid = static_cast<HandleT*>(handle)->deviceGrid<nanovdb::ValueOnIndex>();
	const int bs = 256;
	const int gs = int((f->n + bs - 1) / bs);
	nanovdb::Vec3f* tmp = nullptr;  // a scratch output: the frame's own buffers keep the results of the last ref_frame_run
	cudaMalloc(&tmp, f->n * 12);
	vorticityConfinement<<<gs, bs>>>(grid, f->coords, f->vel, tmp, dt, 1.0f / voxelSize, scale, factorScale, f->n);
	const cudaError_t e = cudaDeviceSynchronize();
	cudaMemcpy(outHost, tmp, f->n * 12, cudaMemcpyDeviceToHost);
	cudaFree(tmp);
	if (e != cudaSuccess) return 2;
	REF_CATCH
}

// One reference kernel of the collision path (hasCollision = true) on the frame's buffers, launched as Compute() launches it
// (HNanoSolver.cu:153-157, 164-169, 282-296, 346-348). stage 0: enforceCollisionBoundaries on a copy of the input velocity -> outVec;
// 1: advect_vector(velocity, sdf) -> outVec; 2: subtractPressureGradient(advected velocity and pressure of the last ref_frame_run,
// sdf) -> outVec; 3: advect_scalars(projected velocity of the last ref_frame_run, sdf) -> fetch with ref_frame_download.
int ref_frame_collision_stage(void* f_, void* handle, int stage, const float* sdfHost, float dt, float voxelSize, float* outVec) {
	REF_TRY
	auto* f = static_cast<RefFrame*>(f_);
	auto* grid = static_cast<HandleT*>(handle)->deviceGrid<nanovdb::ValueOnIndex>();
	const float inv = 1.0f / voxelSize;
	const int bs = 256;
	const int gs = int((f->n + bs - 1) / bs);
	if (!f->sdf) cudaMalloc(&f->sdf, f->n * 4);
	cudaMemcpy(f->sdf, sdfHost, f->n * 4, cudaMemcpyHostToDevice);
	nanovdb::Vec3f* tmp = nullptr;
	cudaMalloc(&tmp, f->n * 12);
	switch (stage) {
		case 0:
			cudaMemcpy(tmp, f->vel, f->n * 12, cudaMemcpyDeviceToDevice);
			enforceCollisionBoundaries<<<gs, bs>>>(grid, f->coords, tmp, f->sdf, voxelSize, f->n);
			break;
		case 1: advect_vector<<<gs, bs>>>(grid, f->coords, f->vel, tmp, f->sdf, true, f->n, dt, inv); break;
		case 2: subtractPressureGradient<<<gs, bs>>>(grid, f->coords, f->n, f->adv, f->p, tmp, f->sdf, true, inv); break;
		case 3: advect_scalars<<<gs, bs>>>(grid, f->coords, f->proj, f->dIn, f->dOut, f->S, f->sdf, true, f->n, dt, inv); break;
		default: cudaFree(tmp); return 1;
	}
	const cudaError_t e = cudaDeviceSynchronize();
	if (outVec && stage != 3) cudaMemcpy(outVec, tmp, f->n * 12, cudaMemcpyDeviceToHost);
	cudaFree(tmp);
	if (e != cudaSuccess) return 2;
	REF_CATCH
}

#endif  // HNS_SHIM_LAUNCHERS_ONLY

}  // extern "C"
