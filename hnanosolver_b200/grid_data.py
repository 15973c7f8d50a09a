"""Host sidecar container: the Python mirror of HNS::GridIndexedData (reference src/Utils/GridData.hpp:16-166,
src/Utils/Memory.hpp:22-160). Same method names, same semantics: a coords block plus named, typed value blocks kept in
insertion order (getBlocksOfType returns names in insertion order, GridData.hpp:136-145 -- this fixes the scalar order
of the all-in-one frame)."""
from __future__ import annotations

import enum

import numpy as np

FLOAT = "float"          # TypedValueBlock<float>
VEC3F = "openvdb::Vec3f"  # TypedValueBlock<openvdb::Vec3f>, stored AoS float[N][3]


class AllocationType(enum.Enum):
    """reference: enum class AllocationType (src/Utils/Memory.hpp:20)"""
    Standard = 0
    Aligned = 1
    CudaPinned = 2


def _alloc(shape, dtype, mode: AllocationType) -> np.ndarray:
    if mode == AllocationType.CudaPinned:
        import torch

        if torch.cuda.is_available():
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return t.numpy()  # shares the pinned storage; the tensor stays alive through the array's base
    if mode == AllocationType.Aligned:
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        raw = np.empty(nbytes + 64, np.uint8)
        off = (-raw.ctypes.data) % 64
        return raw[off:off + nbytes].view(dtype).reshape(shape)
    return np.empty(shape, dtype)


class GridIndexedData:
    def __init__(self):
        self._coords: np.ndarray | None = None
        self._blocks: list[tuple[str, str, np.ndarray]] = []  # (name, type, array) in insertion order
        self._size = 0
        self._alloc = AllocationType.Standard

    # --- allocation -------------------------------------------------------------------------------------
    def setAllocationType(self, mode: AllocationType) -> None:
        self._alloc = mode

    def allocateCoords(self, numElements: int) -> bool:
        self.clearCoords()
        self._coords = _alloc((numElements, 3), np.int32, self._alloc)
        self._size = numElements
        return True

    def addValueBlock(self, type_: str, name: str, numElements: int | None = None) -> bool:
        """addValueBlock<T>(name, numElements): False if the name already exists (GridData.hpp:62-76)."""
        if any(n == name for n, _, _ in self._blocks):
            return False
        n = self._size if numElements is None else numElements
        if type_ == FLOAT:
            arr = _alloc((n,), np.float32, self._alloc)
        elif type_ == VEC3F:
            arr = _alloc((n, 3), np.float32, self._alloc)
        else:
            raise TypeError(f"unsupported block type {type_!r}")
        self._blocks.append((name, type_, arr))
        return True

    # --- access -----------------------------------------------------------------------------------------
    def pCoords(self) -> np.ndarray | None:
        return self._coords

    def pValues(self, type_: str, name: str) -> np.ndarray | None:
        """pValues<T>(name): None if the block does not exist or has another type (GridData.hpp:79-104)."""
        for n, t, a in self._blocks:
            if n == name:
                return a if t == type_ else None
        return None

    def getBlocksOfType(self, type_: str) -> list[str]:
        return [n for n, t, _ in self._blocks if t == type_]

    def size(self) -> int:
        return self._size

    def numValueBlocks(self) -> int:
        return len(self._blocks)

    # --- deallocation -----------------------------------------------------------------------------------
    def clear(self) -> None:
        self.clearValues()
        self.clearCoords()
        self._size = 0

    def clearValues(self) -> None:
        self._blocks.clear()

    def clearCoords(self) -> None:
        self._coords = None

    # --- convenience (not in the reference) ---------------------------------------------------------------
    @classmethod
    def from_arrays(cls, coords: np.ndarray, velocity: np.ndarray | None = None, velocity_name: str = "vel",
                    alloc: AllocationType = AllocationType.Standard, **floats: np.ndarray) -> "GridIndexedData":
        d = cls()
        d.setAllocationType(alloc)
        coords = np.asarray(coords, np.int32).reshape(-1, 3)
        d.allocateCoords(coords.shape[0])
        d.pCoords()[:] = coords
        if velocity is not None:
            d.addValueBlock(VEC3F, velocity_name)
            d.pValues(VEC3F, velocity_name)[:] = np.asarray(velocity, np.float32).reshape(-1, 3)
        for name, arr in floats.items():
            d.addValueBlock(FLOAT, name)
            d.pValues(FLOAT, name)[:] = np.asarray(arr, np.float32).reshape(-1)
        return d
