"""Host-side logic that needs no GPU: the GridIndexedData mirror, workload ordering, launcher argument validation."""
import numpy as np
import pytest

import hnanosolver_b200 as H
from hnanosolver_b200 import synth
from hnanosolver_b200.grid_data import FLOAT, VEC3F, AllocationType, GridIndexedData


@pytest.mark.parametrize("mode", list(AllocationType))
def test_allocate_coords_and_add_blocks(mode):
    # Tests/IndexGrid.cpp:473-516 (AllocateCoordsAndAddBlocks) for Standard / Aligned / CudaPinned
    d = GridIndexedData()
    d.setAllocationType(mode)
    assert d.allocateCoords(1000) and d.size() == 1000 and d.pCoords() is not None
    assert d.addValueBlock(FLOAT, "density", 1000)
    assert d.addValueBlock(VEC3F, "velocity", 1000)
    assert not d.addValueBlock(FLOAT, "density", 1000)          # duplicate name
    assert d.numValueBlocks() == 2
    assert d.pValues(FLOAT, "density").shape == (1000,)
    assert d.pValues(VEC3F, "velocity").shape == (1000, 3)
    assert d.pValues(VEC3F, "density") is None                  # type mismatch
    assert d.pValues(FLOAT, "nope") is None
    if mode == AllocationType.Aligned:
        assert d.pValues(FLOAT, "density").ctypes.data % 64 == 0


def test_clear_blocks():
    # Tests/IndexGrid.cpp:518-539 (ClearBlocks)
    d = GridIndexedData()
    d.allocateCoords(10)
    d.addValueBlock(FLOAT, "a")
    d.clearValues()
    assert d.numValueBlocks() == 0 and d.size() == 10
    d.clear()
    assert d.size() == 0 and d.pCoords() is None


def test_blocks_of_type_keep_insertion_order():
    d = GridIndexedData()
    d.allocateCoords(4)
    for n in ("density", "fuel", "waste"):
        d.addValueBlock(FLOAT, n)
    d.addValueBlock(VEC3F, "vel")
    d.addValueBlock(FLOAT, "temperature")
    assert d.getBlocksOfType(FLOAT) == ["density", "fuel", "waste", "temperature"]   # GridData.hpp:136-145
    assert d.getBlocksOfType(VEC3F) == ["vel"]


def test_dense_coords_follow_leaf_offset_order():
    o = np.array([[8, -16, 24]], np.int32)
    c = synth.dense_coords(o)
    j = 5 << 6 | 3 << 3 | 6
    assert c.shape == (512, 3) and c[j].tolist() == [8 + 5, -16 + 3, 24 + 6]       # GridBuilder.hpp:156-166


def test_nanovdb_order_is_hierarchical_not_plain_lexicographic():
    o = np.array([[128, 0, 0], [0, 0, 8], [0, 120, 0], [-8, 0, 0], [0, 0, 0], [4096, 0, 0]], np.int32)
    s = o[synth.nanovdb_order(o)]
    assert s.tolist() == [[-8, 0, 0], [0, 0, 0], [0, 0, 8], [0, 120, 0], [128, 0, 0], [4096, 0, 0]]


def test_compute_sim_argument_validation_order():
    # Compute() input validation (HNanoSolver.cu:12-28): invalid_argument in this order, then silent return on empty data
    d = GridIndexedData()
    p = H.CombustionParams()
    with pytest.raises(ValueError, match="voxelSize must be positive"):
        H.Compute_Sim(d, None, 0, -1.0, 0.0, p, False)
    with pytest.raises(ValueError, match="dt"):
        H.Compute_Sim(d, None, 0, -1.0, 0.1, p, False)
    with pytest.raises(ValueError, match="iterations"):
        H.Compute_Sim(d, None, 0, 0.1, 0.1, p, False)
    with pytest.raises(ValueError, match="GridHandle"):
        H.Compute_Sim(d, None, 1, 0.1, 0.1, p, False)


def test_launchers_require_exactly_one_velocity_block():
    d = GridIndexedData()
    d.allocateCoords(512)
    d.addValueBlock(FLOAT, "density")
    for fn, args in ((H.AdvectIndexGrid, (0.1, 0.1)), (H.AdvectIndexGridVelocity, (0.1, 0.1)), (H.ProjectNonDivergent, (4, 0.1)),
                     (H.Divergence, (0.1,))):
        with pytest.raises(RuntimeError, match="exactly one Vec3f block"):
            fn(d, *args)


def test_workload_shapes_are_dense_leaves():
    w = synth.smoke_sphere(32, 1)
    assert w.num_voxels == w.num_leaves * 512 == w.coords.shape[0] == w.velocity.shape[0]
    assert np.array_equal(w.coords[::512], w.origins)
    assert np.abs(w.velocity).max() * w.dt / w.voxel_size <= w.meta["cfl_max"] + 1e-6
