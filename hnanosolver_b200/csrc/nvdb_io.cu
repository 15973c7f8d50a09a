// Host-only NanoVDB file IO for index grids (round trip of the topology without Houdini / OpenVDB).
//
// Format: NanoVDB's uncompressed single-grid segment (reference externals/nanovdb/NanoVDB.h:6252-6299, writer :6316-6341, reader
// :6369-6422): FileHeader (16 B: magic, version, gridCount = 1, codec NONE = 0) | FileMetaData (176 B) | grid name incl. '\0' | the
// raw grid buffer. Files written here are read by stock NanoVDB tools (io::readGrid, nanovdb_print), and the reader accepts both that
// layout and a raw buffer dump (which starts with GridData itself, :6375-6386). Everything is restated from the documented layout:
// no NanoVDB header is included. Nothing here touches the GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

constexpr uint64_t kMagicNumb = 0x304244566f6e614eull, kMagicGrid = 0x314244566f6e614eull, kMagicFile = 0x324244566f6e614eull;
// GridData (672 B): magic 0, checksum 8, version 16, flags 20, gridIndex 24, gridCount 28, gridSize 32, gridName[256] 40, map 296,
// worldBBox 560 (6 doubles), voxelSize 608 (3 doubles), gridClass 632, gridType 636, blind metadata 640...
constexpr size_t kOffVersion = 16, kOffGridIndex = 24, kOffGridCount = 28, kOffGridSize = 32, kOffName = 40, kNameMax = 256, kOffWorldBBox = 560,
                 kOffVoxelSize = 608, kOffGridClass = 632, kOffGridType = 636;
// TreeData (64 B) behind it: nodeOffset[4] (leaf, lower, upper, root; bytes from the tree) 0, nodeCount[3] 32, tileCount[3] 44, voxelCount 56
constexpr size_t kTree = hns::nvdb::kGrid, kOffNodeOffset = 0, kOffNodeCount = 32, kOffTileCount = 44, kOffVoxelCount = 56;
constexpr uint32_t kGridTypeOnIndex = 20;  // GridType::OnIndex

#pragma pack(push, 1)
struct FileHeader {
	uint64_t magic;
	uint32_t version;
	uint16_t gridCount;
	uint16_t codec;
};
struct FileMetaData {
	uint64_t gridSize, fileSize, nameKey, voxelCount;
	uint32_t gridType, gridClass;
	double worldBBox[6];
	int32_t indexBBox[6];
	double voxelSize[3];
	uint32_t nameSize;
	uint32_t nodeCount[4];
	uint32_t tileCount[3];
	uint16_t codec, padding;
	uint32_t version;
};
#pragma pack(pop)
static_assert(sizeof(FileHeader) == 16 && sizeof(FileMetaData) == 176, "NanoVDB file structures");

template <typename T>
T rd(const uint8_t* p) {
	T v;
	std::memcpy(&v, p, sizeof(T));
	return v;
}

// a buffer that can be a single ValueOnIndex grid: long enough, known magic, sizes consistent
int check_grid(const uint8_t* g, uint64_t bytes, bool need_index_grid) {
	using namespace hns;
	if (!g || bytes < nvdb::kGrid + nvdb::kTree + nvdb::kRoot) return fail(HNS_ERR_INVALID_ARGUMENT, "buffer is too small for a NanoVDB grid");
	const uint64_t magic = rd<uint64_t>(g);
	if (magic != kMagicNumb && magic != kMagicGrid) return fail(HNS_ERR_INVALID_ARGUMENT, "buffer does not start with a NanoVDB grid magic number");
	if (rd<uint64_t>(g + kOffGridSize) > bytes) return fail(HNS_ERR_INVALID_ARGUMENT, "grid size in the header exceeds the buffer");
	if (need_index_grid && rd<uint32_t>(g + kOffGridType) != kGridTypeOnIndex)
		return fail(HNS_ERR_UNSUPPORTED, "not a ValueOnIndex grid (grid type " + std::to_string(rd<uint32_t>(g + kOffGridType)) + ")");
	return HNS_OK;
}

// Does a file start with GridData itself (a raw buffer dump) rather than with a FileHeader? NanoVDB's own test (GridData::isValid,
// NanoVDB.h:1864-1875): grid magic, or the 32.6 marker in mData2, or the shared magic with gridIndex < gridCount and sane class / type
// -- in a segment those bytes hold FileMetaData::fileSize, whose upper half (read as gridCount) is 0 for files below 4 GB.
bool starts_with_grid(const uint8_t* first, size_t got) {
	if (got < hns::nvdb::kGrid) return false;
	const uint64_t magic = rd<uint64_t>(first);
	if (magic == kMagicGrid || rd<uint64_t>(first + 664) == kMagicGrid) return true;
	if (magic != kMagicNumb) return false;
	const uint32_t major = rd<uint32_t>(first + kOffVersion) >> 21, index = rd<uint32_t>(first + kOffGridIndex), count = rd<uint32_t>(first + kOffGridCount);
	return major == 32 && count > 0 && index < count && rd<uint32_t>(first + kOffGridClass) < 10 && rd<uint32_t>(first + kOffGridType) < 27;
}

struct File {
	FILE* f = nullptr;
	explicit File(const char* path, const char* mode) : f(std::fopen(path, mode)) {}
	~File() {
		if (f) std::fclose(f);
	}
};

}  // namespace

extern "C" {

// Writes one grid (host copy of a NanoVDB buffer, e.g. from hns_grid_nanovdb_download) as an uncompressed .nvdb file.
int hns_nvdb_write(const char* path, const void* nanovdb_buffer, uint64_t bytes) {
	using namespace hns;
	if (!path) return fail(HNS_ERR_INVALID_ARGUMENT, "path is null");
	const uint8_t* g = static_cast<const uint8_t*>(nanovdb_buffer);
	int rc = check_grid(g, bytes, false);
	if (rc) return rc;
	const uint8_t* tree = g + kTree;
	const uint64_t grid_size = rd<uint64_t>(g + kOffGridSize);
	const char* name = reinterpret_cast<const char*>(g + kOffName);
	const uint32_t name_size = uint32_t(strnlen(name, kNameMax - 1)) + 1;  // including the terminator
	FileHeader head{kMagicNumb, rd<uint32_t>(g + kOffVersion), 1, 0};       // NANOVDB_USE_NEW_MAGIC_NUMBERS is off in 32.7 (NanoVDB.h:142)
	FileMetaData meta{};
	meta.gridSize = meta.fileSize = grid_size;
	meta.nameKey = 0;  // as writeUncompressedGrid does
	meta.voxelCount = rd<uint64_t>(tree + kOffVoxelCount);
	meta.gridType = rd<uint32_t>(g + kOffGridType), meta.gridClass = rd<uint32_t>(g + kOffGridClass);
	std::memcpy(meta.worldBBox, g + kOffWorldBBox, sizeof(meta.worldBBox));
	const uint64_t root_off = rd<uint64_t>(tree + kOffNodeOffset + 3 * 8);  // TreeData::bbox() = the root's bbox, its first member (:2281)
	if (root_off) {
		if (kTree + root_off + sizeof(meta.indexBBox) > bytes) return fail(HNS_ERR_INVALID_ARGUMENT, "root offset points outside the buffer");
		std::memcpy(meta.indexBBox, tree + root_off, sizeof(meta.indexBBox));
	} else {  // CoordBBox(): an empty box
		for (int i = 0; i < 3; ++i) meta.indexBBox[i] = INT32_MAX, meta.indexBBox[3 + i] = INT32_MIN;
	}
	std::memcpy(meta.voxelSize, g + kOffVoxelSize, sizeof(meta.voxelSize));
	meta.nameSize = name_size;
	for (int i = 0; i < 3; ++i) meta.nodeCount[i] = rd<uint32_t>(tree + kOffNodeCount + 4 * i), meta.tileCount[i] = rd<uint32_t>(tree + kOffTileCount + 4 * i);
	meta.nodeCount[3] = 1;
	meta.codec = 0, meta.padding = 0, meta.version = head.version;
	File out(path, "wb");
	if (!out.f) return fail(HNS_ERR_RUNTIME, std::string("cannot open for writing: ") + path);
	std::string name_bytes(name, name_size - 1);
	name_bytes.push_back('\0');
	if (std::fwrite(&head, sizeof(head), 1, out.f) != 1 || std::fwrite(&meta, sizeof(meta), 1, out.f) != 1 ||
	    std::fwrite(name_bytes.data(), 1, name_size, out.f) != name_size || std::fwrite(g, 1, grid_size, out.f) != grid_size)
		return fail(HNS_ERR_RUNTIME, std::string("short write: ") + path);
	return HNS_OK;
}

// Size of the first grid stored in `path` (segment layout or raw buffer dump), so that the caller can allocate.
int hns_nvdb_file_grid_bytes(const char* path, uint64_t* bytes_out) {
	using namespace hns;
	if (!path || !bytes_out) return fail(HNS_ERR_INVALID_ARGUMENT, "null argument");
	*bytes_out = 0;
	File in(path, "rb");
	if (!in.f) return fail(HNS_ERR_RUNTIME, std::string("cannot open: ") + path);
	uint8_t first[nvdb::kGrid] = {};
	const size_t got = std::fread(first, 1, sizeof(first), in.f);
	if (starts_with_grid(first, got)) {
		*bytes_out = rd<uint64_t>(first + kOffGridSize);
		return HNS_OK;
	}
	if (got < sizeof(FileHeader) + sizeof(FileMetaData)) return fail(HNS_ERR_INVALID_ARGUMENT, "file is too short for a NanoVDB grid");
	const FileHeader head = rd<FileHeader>(first);
	const FileMetaData meta = rd<FileMetaData>(first + sizeof(FileHeader));
	if (head.magic != kMagicNumb && head.magic != kMagicFile) return fail(HNS_ERR_INVALID_ARGUMENT, "not a NanoVDB file (magic number)");
	if (head.gridCount < 1) return fail(HNS_ERR_INVALID_ARGUMENT, "NanoVDB file segment holds no grid");
	if (head.codec != 0 || meta.codec != 0) return fail(HNS_ERR_UNSUPPORTED, "compressed NanoVDB files (ZIP / BLOSC) are not supported");
	*bytes_out = meta.gridSize;
	return HNS_OK;
}

// Reads the first grid of `path` into `dst` (capacity `capacity` bytes, see hns_nvdb_file_grid_bytes).
int hns_nvdb_read(const char* path, void* dst, uint64_t capacity) {
	using namespace hns;
	uint64_t need = 0;
	int rc = hns_nvdb_file_grid_bytes(path, &need);
	if (rc) return rc;
	if (!dst || capacity < need) return fail(HNS_ERR_INVALID_ARGUMENT, "destination buffer is too small");
	File in(path, "rb");
	if (!in.f) return fail(HNS_ERR_RUNTIME, std::string("cannot open: ") + path);
	uint8_t first[nvdb::kGrid] = {};
	const size_t got = std::fread(first, 1, sizeof(first), in.f);
	long skip = 0;
	if (!starts_with_grid(first, got)) skip = long(sizeof(FileHeader) + sizeof(FileMetaData) + rd<FileMetaData>(first + sizeof(FileHeader)).nameSize);
	if (std::fseek(in.f, skip, SEEK_SET) != 0) return fail(HNS_ERR_RUNTIME, "seek failed");
	if (std::fread(dst, 1, need, in.f) != need) return fail(HNS_ERR_RUNTIME, "short read: file ends inside the grid");
	return check_grid(static_cast<const uint8_t*>(dst), need, false);
}

// Leaf origins (int32[L][3], in the buffer's leaf order = NanoVDB order) and voxel size of a ValueOnIndex grid buffer on the host:
// what hns_grid_create_from_origins needs to rebuild the device grid. origins_out may be null to query the count.
int hns_nvdb_leaf_origins(const void* nanovdb_buffer, uint64_t bytes, int32_t* origins_out, uint64_t* num_leaves_out, float* voxel_size_out) {
	using namespace hns;
	const uint8_t* g = static_cast<const uint8_t*>(nanovdb_buffer);
	int rc = check_grid(g, bytes, true);
	if (rc) return rc;
	const uint8_t* tree = g + kTree;
	const uint64_t leaf_off = rd<uint64_t>(tree + kOffNodeOffset), L = rd<uint32_t>(tree + kOffNodeCount);
	if (L && kTree + leaf_off + L * nvdb::kLeaf > bytes) return fail(HNS_ERR_INVALID_ARGUMENT, "leaf nodes lie outside the buffer");
	if (num_leaves_out) *num_leaves_out = L;
	if (voxel_size_out) *voxel_size_out = float(rd<double>(g + kOffVoxelSize));
	if (origins_out)
		for (uint64_t l = 0; l < L; ++l) std::memcpy(origins_out + 3 * l, tree + leaf_off + l * nvdb::kLeaf, 12);  // LeafData starts with mBBoxMin
	return HNS_OK;
}

}  // extern "C"
