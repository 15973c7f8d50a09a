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
