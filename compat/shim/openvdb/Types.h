// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product.
//
// Minimal stand-in for <openvdb/Types.h> so that the reference launchers
// (/root/reference/src/Cuda/*.cu) compile without OpenVDB/Houdini. The reference
// uses openvdb::Coord / openvdb::Vec3f purely as 12-byte POD type names for the
// host sidecar (reference src/Cuda/HNanoSolver.cu:48,54, src/Utils/GridData.hpp:96);
// no OpenVDB function is ever called on the launcher path.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace openvdb {
struct Coord {
	int32_t v[3];
	int32_t x() const { return v[0]; }
	int32_t y() const { return v[1]; }
	int32_t z() const { return v[2]; }
};
struct Vec3f {
	float v[3];
};
}  // namespace openvdb
static_assert(sizeof(openvdb::Coord) == 12 && sizeof(openvdb::Vec3f) == 12, "POD layout");
