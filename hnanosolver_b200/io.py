"""Topology + sidecar round trip through files, without Houdini or OpenVDB.

The index grid travels as a standard uncompressed NanoVDB file (libhns_b200: hns_nvdb_write / hns_nvdb_read, host only), readable
by stock NanoVDB tools; the sidecar blocks (reference src/Utils/GridData.hpp: named float / Vec3f arrays in leaf order) as .npy files
next to it, listed in a small JSON manifest in insertion order -- the order that fixes the scalar order inside advect_scalars
(GridData.hpp:136-145).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _lib
from .grid_data import FLOAT, VEC3F, GridIndexedData


def write_nvdb(path: str, nanovdb_buffer: np.ndarray) -> None:
    """nanovdb_buffer: uint8 array holding one NanoVDB grid (IndexGridHandle.nanovdb_buffer(), or the oracle's)."""
    buf = np.ascontiguousarray(nanovdb_buffer, np.uint8)
    _lib.check(_lib.lib().hns_nvdb_write(os.fsencode(path), buf.ctypes.data_as(C.c_void_p), buf.size))


def read_nvdb(path: str) -> np.ndarray:
    """The first grid of an uncompressed .nvdb file (or of a raw buffer dump) as a uint8 array."""
    n = C.c_uint64()
    _lib.check(_lib.lib().hns_nvdb_file_grid_bytes(os.fsencode(path), C.byref(n)))
    buf = np.empty(n.value, np.uint8)
    _lib.check(_lib.lib().hns_nvdb_read(os.fsencode(path), buf.ctypes.data_as(C.c_void_p), buf.size))
    return buf


def leaf_origins(nanovdb_buffer: np.ndarray):
    """(origins int32 (L, 3) in NanoVDB order, voxel size) of a ValueOnIndex grid buffer."""
    buf = np.ascontiguousarray(nanovdb_buffer, np.uint8)
    n, h = C.c_uint64(), C.c_float()
    L = _lib.lib()
    _lib.check(L.hns_nvdb_leaf_origins(buf.ctypes.data_as(C.c_void_p), buf.size, None, C.byref(n), C.byref(h)))
    origins = np.empty((n.value, 3), np.int32)
    _lib.check(L.hns_nvdb_leaf_origins(buf.ctypes.data_as(C.c_void_p), buf.size, origins.ctypes.data_as(_lib.c_i32p), C.byref(n), C.byref(h)))
    return origins, float(h.value)


def save_cache(directory: str, nanovdb_buffer: np.ndarray, data: GridIndexedData) -> None:
    """grid.nvdb + one .npy per sidecar block + manifest.json (block names and types in insertion order)."""
    os.makedirs(directory, exist_ok=True)
    write_nvdb(os.path.join(directory, "grid.nvdb"), nanovdb_buffer)
    blocks = []
    for name, kind, arr in data._blocks:  # insertion order
        np.save(os.path.join(directory, f"{len(blocks):02d}_{name}.npy"), arr)
        blocks.append({"name": name, "type": kind, "file": f"{len(blocks):02d}_{name}.npy"})
    with open(os.path.join(directory, "manifest.json"), "w") as f:
        json.dump({"voxels": int(data.size()), "blocks": blocks}, f, indent=1)


def load_cache(directory: str):
    """-> (origins (L, 3), voxel size, GridIndexedData with coords rebuilt from the leaf origins and the stored blocks)."""
    from . import synth

    origins, h = leaf_origins(read_nvdb(os.path.join(directory, "grid.nvdb")))
    with open(os.path.join(directory, "manifest.json")) as f:
        man = json.load(f)
    data = GridIndexedData()
    data.allocateCoords(origins.shape[0] * 512)
    if origins.shape[0]:
        data.pCoords()[:] = synth.dense_coords(origins)
    if man["voxels"] != data.size():
        raise ValueError(f"manifest lists {man['voxels']} voxels, the grid has {data.size()}")
    for b in man["blocks"]:
        arr = np.load(os.path.join(directory, b["file"]))
        kind = VEC3F if b["type"] == VEC3F else FLOAT
        data.addValueBlock(kind, b["name"])
        data.pValues(kind, b["name"])[:] = arr
    return origins, h, data


# ------------------------------------------------------------------------------------------------------------------
# value grids <-> sidecar blocks: IndexGridBuilder::build / writeIndexGrid (reference src/Utils/GridBuilder.hpp:87-216) over NanoVDB
# float / Vec3f grids (OpenVDB is not vendored with the reference; NanoVDB grids carry the same 512-value leaf buffers)
# ------------------------------------------------------------------------------------------------------------------
GRID_TYPE_FLOAT, GRID_TYPE_VEC3F, GRID_TYPE_ONINDEX = 1, 6, 20
GRID_CLASS_FOG_VOLUME, GRID_CLASS_STAGGERED = 2, 3


def _buf(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def grid_info(nanovdb_buffer: np.ndarray) -> dict:
    buf, p = _buf(nanovdb_buffer)
    t, c, n, h = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_float()
    name = C.create_string_buffer(256)
    _lib.check(_lib.lib().hns_nvdb_grid_info(p, buf.size, C.byref(t), C.byref(c), C.byref(n), C.byref(h), name))
    return {"grid_type": t.value, "grid_class": c.value, "num_leaves": n.value, "voxel_size": float(h.value), "name": name.value.decode()}


def leaf_topology(nanovdb_buffer: np.ndarray):
    """(leaf origins int32 (L, 3), active-voxel masks uint64 (L, 8)) of a float / Vec3f / index grid: the inputs of build_domain"""
    buf, p = _buf(nanovdb_buffer)
    n = C.c_uint64()
    L = _lib.lib()
    _lib.check(L.hns_nvdb_leaf_topology(p, buf.size, None, None, C.byref(n)))
    origins, masks = np.empty((n.value, 3), np.int32), np.empty((n.value, 8), np.uint64)
    _lib.check(L.hns_nvdb_leaf_topology(p, buf.size, origins.ctypes.data_as(_lib.c_i32p), masks.ctypes.data_as(_lib.c_u64p), C.byref(n)))
    return origins, masks


def sidecar_from_grid(nanovdb_buffer: np.ndarray, domain_origins: np.ndarray, fill_byte: int = 0) -> np.ndarray:
    """One block of IndexGridBuilder::build: float32 (N,) for a float grid, (N, 3) for a Vec3f grid, N = 512 per domain leaf; leaves
    the grid does not have are filled with bytes `fill_byte` (1 for the collision SDF, GridBuilder.hpp:108)."""
    buf, p = _buf(nanovdb_buffer)
    o = np.ascontiguousarray(np.asarray(domain_origins, np.int32).reshape(-1, 3))
    comps = 3 if grid_info(buf)["grid_type"] == GRID_TYPE_VEC3F else 1
    out = np.empty((o.shape[0] * 512, 3) if comps == 3 else o.shape[0] * 512, np.float32)
    _lib.check(_lib.lib().hns_sidecar_from_nanovdb(p, buf.size, o.ctypes.data_as(_lib.c_i32p), o.shape[0], int(fill_byte), out.ctypes.data_as(C.c_void_p)))
    return out


def grid_from_sidecar(domain_origins: np.ndarray, values: np.ndarray, voxel_size: float, name: str, masks=None) -> np.ndarray:
    """writeIndexGrid: the block as a NanoVDB float (FogVolume) or Vec3f (Staggered) grid over the domain's leaves -> uint8 buffer"""
    o = np.ascontiguousarray(np.asarray(domain_origins, np.int32).reshape(-1, 3))
    v = np.ascontiguousarray(values, np.float32)
    comps = 3 if v.ndim == 2 else 1
    assert v.shape[0] == o.shape[0] * 512
    m = None if masks is None else np.ascontiguousarray(np.asarray(masks, np.uint64).reshape(-1, 8))
    L = _lib.lib()
    n = L.hns_sidecar_nanovdb_bytes(o.ctypes.data_as(_lib.c_i32p), o.shape[0], comps)
    out = np.empty(n, np.uint8)
    _lib.check(L.hns_sidecar_to_nanovdb(o.ctypes.data_as(_lib.c_i32p), o.shape[0], m.ctypes.data_as(_lib.c_u64p) if m is not None else None,
                                        v.ctypes.data_as(C.c_void_p), comps, voxel_size, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
    return out


def build_sidecar(domain_origins: np.ndarray, grids) -> GridIndexedData:
    """IndexGridBuilder over a domain: grids = [(name, nanovdb buffer, is_sdf)], blocks added in that order (GridBuilder.hpp:58-166)"""
    from . import synth

    o = np.ascontiguousarray(np.asarray(domain_origins, np.int32).reshape(-1, 3))
    data = GridIndexedData()
    data.allocateCoords(o.shape[0] * 512)
    if o.shape[0]:
        data.pCoords()[:] = synth.dense_coords(o)
    for name, buf, is_sdf in grids:
        block = sidecar_from_grid(buf, o, 1 if is_sdf else 0)
        kind = VEC3F if block.ndim == 2 else FLOAT
        data.addValueBlock(kind, name)
        data.pValues(kind, name)[:] = block
    return data


def write_index_grids(directory: str, domain_origins: np.ndarray, data: GridIndexedData, voxel_size: float) -> list:
    """writeIndexGrid for every block of the sidecar: one <name>.nvdb per block; returns the paths"""
    os.makedirs(directory, exist_ok=True)
    paths = []
    for name, kind, arr in data._blocks:
        path = os.path.join(directory, f"{name}.nvdb")
        write_nvdb(path, grid_from_sidecar(domain_origins, arr, voxel_size, name))
        paths.append(path)
    return paths


# ------------------------------------------------------------------------------------------------------------------
# sourcing: SOP_HNanoSolverVerb::cook adds its source grids to the fed-back state with openvdb::tools::compSum before it defines the
# domain (reference src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:159-181). compSum(A, B) leaves A with the union of both topologies and
# A + B in every voxel (a voxel a grid does not hold counts as its background, 0). Over NanoVDB value grids and sidecar blocks that is:
# merge the topologies in front of the domain construction, then add the source's block to the state's block over the domain.
# ------------------------------------------------------------------------------------------------------------------
def union_topology(*topologies):
    """(origins (L, 3), masks (L, 8)) of the union of several (origins, masks) pairs as leaf_topology returns them: a leaf is present
    when any grid has it, a voxel active when it is active in any of them. Sorted by (x, y, z) of the origin; build_domain orders its
    result itself."""
    origins = np.concatenate([np.asarray(o, np.int32).reshape(-1, 3) for o, _ in topologies] or [np.zeros((0, 3), np.int32)])
    masks = np.concatenate([np.asarray(m, np.uint64).reshape(-1, 8) for _, m in topologies] or [np.zeros((0, 8), np.uint64)])
    if not len(origins):
        return origins, masks
    uniq, inverse = np.unique(origins, axis=0, return_inverse=True)
    out = np.zeros((uniq.shape[0], 8), np.uint64)
    np.bitwise_or.at(out, inverse.reshape(-1), masks)
    return np.ascontiguousarray(uniq, np.int32), out


def comp_sum(domain_origins: np.ndarray, data: GridIndexedData, sources) -> None:
    """sources = [(block name, NanoVDB float / Vec3f grid buffer)]: adds every source grid to the sidecar block of that name, in place,
    over the domain's leaves (leaves the source does not have add nothing). The domain has to cover the sources' topology -- pass
    union_topology(state, sources...) to build_domain -- or the part outside it is lost, as it would be in the reference if the source
    were added after the domain had been defined."""
    o = np.ascontiguousarray(np.asarray(domain_origins, np.int32).reshape(-1, 3))
    for name, buf in sources:
        add = sidecar_from_grid(buf, o, 0)
        kind = VEC3F if add.ndim == 2 else FLOAT
        block = data.pValues(kind, name)
        block += add.reshape(block.shape)
