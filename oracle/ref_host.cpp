// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// The reference's own __hostdev__ code running on the CPU, used to pin oracle/hns_oracle.c without a GPU:
//   * a REAL NanoVDB 32.7 host ValueOnIndex grid (vendored externals/nanovdb: tools/GridBuilder.h +
//     tools/CreateNanoGrid.h) built from a voxel list, queried through nanovdb::ReadAccessor, and
//   * the reference's samplers, IndexOffsetSampler<0> / IndexSampler<T,0|1> / TrilinearSampler
//     (src/Utils/Stencils.hpp:51-173), compiled by g++ (not nvcc) with -D__fmaf_rn=fmaf.
// Built by oracle/Makefile into oracle/_ref/libref_host.so from the headers where they lie under /root/reference.
//
// Note the host instantiation of TrilinearSampler<Vec3f> takes the non-fused branch of lerp_dispatch
// (Stencils.hpp:135-137: a + (b-a)*w), so Vec3f samples can differ from the device path in the last ulp;
// the float path is a + w*(b-a) on both.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../Utils/Stencils.hpp"
#include "nanovdb/NanoVDB.h"
#include "nanovdb/tools/CreateNanoGrid.h"
#include "nanovdb/tools/GridBuilder.h"

struct RefHostGrid {
	nanovdb::GridHandle<nanovdb::HostBuffer> handle;
	const nanovdb::NanoGrid<nanovdb::ValueOnIndex>* grid = nullptr;
};

extern "C" {

void* refhost_create(const int32_t* coords, uint64_t n) {
	using SrcT = nanovdb::tools::build::Grid<float>;
	SrcT src(0.0f);
	auto acc = src.getAccessor();
	for (uint64_t i = 0; i < n; ++i) acc.setValue(nanovdb::Coord(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]), 1.0f);
	auto* g = new RefHostGrid();
	// same call shape as Tests/IndexGrid.cpp:125 (channels = 0 here: no blind data, stats off, tiles off)
	g->handle = nanovdb::tools::createNanoGrid<SrcT, nanovdb::ValueOnIndex, nanovdb::HostBuffer>(src, 0u, false, false, 0);
	g->grid = g->handle.grid<nanovdb::ValueOnIndex>();
	if (!g->grid) {
		delete g;
		return nullptr;
	}
	return g;
}
void refhost_destroy(void* g) { delete static_cast<RefHostGrid*>(g); }
uint64_t refhost_value_count(void* g) { return static_cast<RefHostGrid*>(g)->grid->valueCount(); }
uint64_t refhost_leaf_count(void* g) { return static_cast<RefHostGrid*>(g)->grid->tree().nodeCount(0); }
uint64_t refhost_bytes(void* g) { return static_cast<RefHostGrid*>(g)->handle.buffer().size(); }
const void* refhost_data(void* g) { return static_cast<RefHostGrid*>(g)->handle.data(); }

void refhost_get_values(void* g_, const int32_t* ijk, uint64_t n, uint64_t* out) {
	const IndexOffsetSampler<0> s(static_cast<RefHostGrid*>(g_)->grid);
	for (uint64_t i = 0; i < n; ++i) out[i] = s.offset(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
}
void refhost_nearest_f(void* g_, const float* data, const int32_t* ijk, uint64_t n, float* out) {
	const IndexOffsetSampler<0> s(static_cast<RefHostGrid*>(g_)->grid);
	const IndexSampler<float, 0> f(s, data);
	for (uint64_t i = 0; i < n; ++i) out[i] = f(nanovdb::Coord(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]));
}
void refhost_trilinear_f(void* g_, const float* data, const float* xyz, uint64_t n, float* out) {
	const IndexOffsetSampler<0> s(static_cast<RefHostGrid*>(g_)->grid);
	const IndexSampler<float, 1> f(s, data);
	for (uint64_t i = 0; i < n; ++i) out[i] = f(nanovdb::Vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}
void refhost_trilinear_v(void* g_, const float* data, const float* xyz, uint64_t n, float* out) {
	const IndexOffsetSampler<0> s(static_cast<RefHostGrid*>(g_)->grid);
	const IndexSampler<nanovdb::Vec3f, 1> f(s, reinterpret_cast<const nanovdb::Vec3f*>(data));
	for (uint64_t i = 0; i < n; ++i) {
		const nanovdb::Vec3f v = f(nanovdb::Vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
		out[3 * i] = v[0], out[3 * i + 1] = v[1], out[3 * i + 2] = v[2];
	}
}
// ---- NanoVDB's own file IO on the CPU (pins hnanosolver_b200/csrc/nvdb_io.cu) ----
// writeUncompressedGrid / readUncompressedGrids are the dependency-free reference implementations inside NanoVDB.h (:6316-6422).
struct FileOut {
	FILE* f;
	void write(const char* data, size_t n) { fwrite(data, 1, n, f); }
};
// buffer (any valid NanoVDB grid) -> file, segment layout (raw = 0) or raw dump (raw = 1)
int refhost_write_nvdb(const char* path, const void* buffer, int raw) {
	FileOut os{fopen(path, "wb")};
	if (!os.f) return 1;
	nanovdb::io::writeUncompressedGrid(os, static_cast<const nanovdb::GridData*>(buffer), raw != 0);
	fclose(os.f);
	return 0;
}
// file -> first grid's bytes (returns its size; copies at most `capacity` bytes into dst when dst != null)
uint64_t refhost_read_nvdb(const char* path, void* dst, uint64_t capacity) {
	auto handles = nanovdb::io::readUncompressedGrids<nanovdb::GridHandle<nanovdb::HostBuffer>, std::vector>(path);
	if (handles.empty()) return 0;
	const uint64_t n = handles[0].buffer().size();
	if (dst) memcpy(dst, handles[0].data(), n < capacity ? n : capacity);
	return n;
}
}

// ---- NanoVDB's own float / Vec3f grids on the CPU (pins the product's value-grid IO, hnanosolver_b200/csrc/nvdb_io.cu) ----
// a grid built by NanoVDB's host builder from (coords, values): components 1 -> float, 3 -> Vec3f
struct RefValueGrid {
	nanovdb::GridHandle<nanovdb::HostBuffer> handle;
};
extern "C" {
void* refhost_value_grid_create(const int32_t* coords, const float* values, uint64_t n, int components, double voxel_size, const char* name, int grid_class) {
	auto* out = new RefValueGrid();
	const auto cls = static_cast<nanovdb::GridClass>(grid_class);
	if (components == 1) {
		nanovdb::tools::build::Grid<float> src(0.0f, name, cls);
		src.setTransform(voxel_size);
		auto acc = src.getAccessor();
		for (uint64_t i = 0; i < n; ++i) acc.setValue(nanovdb::Coord(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]), values[i]);
		out->handle = nanovdb::tools::createNanoGrid(src);
	} else {
		nanovdb::tools::build::Grid<nanovdb::Vec3f> src(nanovdb::Vec3f(0.0f), name, cls);
		src.setTransform(voxel_size);
		auto acc = src.getAccessor();
		for (uint64_t i = 0; i < n; ++i)
			acc.setValue(nanovdb::Coord(coords[3 * i], coords[3 * i + 1], coords[3 * i + 2]), nanovdb::Vec3f(values[3 * i], values[3 * i + 1], values[3 * i + 2]));
		out->handle = nanovdb::tools::createNanoGrid(src);
	}
	return out;
}
void refhost_value_grid_destroy(void* g) { delete static_cast<RefValueGrid*>(g); }
uint64_t refhost_value_grid_bytes(void* g) { return static_cast<RefValueGrid*>(g)->handle.buffer().size(); }
const void* refhost_value_grid_data(void* g) { return static_cast<RefValueGrid*>(g)->handle.data(); }

// a grid buffer (e.g. written by the product) read through NanoVDB's own classes: values + active states at n coordinates
// returns 0 on success, 1 when the buffer is not a valid grid of the expected type
int refhost_query_grid(const void* buffer, const int32_t* ijk, uint64_t n, int components, float* values_out, uint8_t* active_out) {
	const auto* gd = static_cast<const nanovdb::GridData*>(buffer);
	if (!gd->isValid()) return 1;
	if (components == 1) {
		if (gd->mGridType != nanovdb::GridType::Float) return 1;
		const auto* grid = static_cast<const nanovdb::NanoGrid<float>*>(buffer);
		auto acc = grid->getAccessor();
		for (uint64_t i = 0; i < n; ++i) {
			const nanovdb::Coord c(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
			values_out[i] = acc.getValue(c), active_out[i] = acc.isActive(c);
		}
	} else {
		if (gd->mGridType != nanovdb::GridType::Vec3f) return 1;
		const auto* grid = static_cast<const nanovdb::NanoGrid<nanovdb::Vec3f>*>(buffer);
		auto acc = grid->getAccessor();
		for (uint64_t i = 0; i < n; ++i) {
			const nanovdb::Coord c(ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
			const nanovdb::Vec3f v = acc.getValue(c);
			values_out[3 * i] = v[0], values_out[3 * i + 1] = v[1], values_out[3 * i + 2] = v[2], active_out[i] = acc.isActive(c);
		}
	}
	return 0;
}
// what NanoVDB reports about a grid buffer: out[0] grid class, [1] grid type, [2] leaf count, [3] lower count, [4] upper count, [5] active voxels
int refhost_grid_meta(const void* buffer, uint64_t* out6, char* name256, double* voxel_size, int32_t* index_bbox6) {
	const auto* gd = static_cast<const nanovdb::GridData*>(buffer);
	if (!gd->isValid()) return 1;
	const auto* grid = static_cast<const nanovdb::NanoGrid<float>*>(buffer);  // GridData / TreeData accessors do not depend on the value type
	out6[0] = uint64_t(gd->mGridClass), out6[1] = uint64_t(gd->mGridType);
	const auto* tree = reinterpret_cast<const nanovdb::TreeData*>(static_cast<const uint8_t*>(buffer) + sizeof(nanovdb::GridData));
	out6[2] = tree->mNodeCount[0], out6[3] = tree->mNodeCount[1], out6[4] = tree->mNodeCount[2], out6[5] = tree->mVoxelCount;
	strncpy(name256, gd->mGridName, 255);
	*voxel_size = gd->mVoxelSize[0];
	const auto bbox = grid->indexBBox();
	for (int a = 0; a < 3; ++a) index_bbox6[a] = bbox.min()[a], index_bbox6[3 + a] = bbox.max()[a];
	return 0;
}
}
