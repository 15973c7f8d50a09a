"""Device-side domain construction against the host set-based restatement of
SOP_HNanoSolverVerb::cook's OpenVDB pass (reference src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:188-199): leaf lists must be identical,
element for element, in NanoVDB order. OpenVDB itself is not vendored: parity unpinned at that boundary."""
import numpy as np
import pytest

from hnanosolver_b200 import synth


def _random_topology(seed, n_leaves, extent, offset, density):
    rng = np.random.default_rng(seed)
    cells = np.unique(rng.integers(0, extent, size=(n_leaves * 2, 3)), axis=0)[:n_leaves]
    origins = (cells * 8 + np.asarray(offset)).astype(np.int32)
    rng.shuffle(origins)                                            # any input order
    masks = np.zeros((origins.shape[0], 8), np.uint64)
    bits = rng.random((origins.shape[0], 512)) < density
    for w in range(8):
        masks[:, w] = (bits[:, 64 * w:64 * w + 64].astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(1, dtype=np.uint64)
    return origins, masks


# ---- CPU: the restatement itself -----------------------------------------------------------------------------------------------
def test_oracle_domain_known_answers(oracle_mod):
    O = oracle_mod
    one = np.array([[0, 0, 0]], np.int32)
    full = np.full((1, 8), np.uint64(0xFFFFFFFFFFFFFFFF))
    assert O.domain_leaves(one, full, 0).tolist() == [[0, 0, 0]]
    assert O.domain_leaves(one, full, 1).shape[0] == 27                       # every voxel active: one voxel of padding reaches all 26 neighbours
    assert O.domain_leaves(one, full, 9).shape[0] == 125
    centre = np.zeros((1, 8), np.uint64)
    centre[0, 3] = np.uint64(1) << np.uint64(3 * 8 + 3)                        # voxel (3, 3, 3) only
    assert O.domain_leaves(one, centre, 3).tolist() == [[0, 0, 0]]             # reaches [0, 6]: stays inside
    assert O.domain_leaves(one, centre, 4).shape[0] == 8                       # reaches -1 on each axis, not 8
    assert O.domain_leaves(one, centre, 5).shape[0] == 27
    empty = np.zeros((1, 8), np.uint64)
    assert O.domain_leaves(one, empty, 8).tolist() == [[0, 0, 0]]              # the leaf node stays (topologyUnion copies nodes), nothing dilates
    got = O.domain_leaves(one, empty, 2, np.array([[64, 0, -8]], np.int32))
    assert sorted(got.tolist()) == [[0, 0, 0], [64, 0, -8]]
    order = synth.nanovdb_order(got)
    assert np.array_equal(order, np.arange(got.shape[0]))                      # already in NanoVDB order


# ---- GPU -----------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(seed=1, n=40, extent=5, offset=(0, 0, 0), density=0.02, padding=0, sdf=False),
                                  dict(seed=2, n=60, extent=6, offset=(-24, 4096 - 16, -40), density=0.01, padding=1, sdf=True),
                                  dict(seed=3, n=60, extent=6, offset=(-24, -8, 120), density=0.004, padding=4, sdf=True),
                                  dict(seed=4, n=30, extent=9, offset=(8, 8, 8), density=0.003, padding=8, sdf=False),
                                  dict(seed=5, n=20, extent=9, offset=(-4096, 0, 4096), density=0.002, padding=13, sdf=True),
                                  dict(seed=6, n=50, extent=4, offset=(0, 0, 0), density=1.0, padding=2, sdf=False)])
def test_device_domain_equals_host_restatement(oracle_mod, case):
    import hnanosolver_b200 as H

    O = oracle_mod
    vo, vm = _random_topology(case["seed"], case["n"], case["extent"], case["offset"], case["density"])
    so = _random_topology(case["seed"] + 100, 9, 12, case["offset"], 0.5)[0] if case["sdf"] else None
    want = O.domain_leaves(vo, vm, case["padding"], so)
    got = H.build_domain(vo, vm, case["padding"], so)
    assert np.array_equal(got, want)
    if case["density"] == 1.0:
        assert np.array_equal(H.build_domain(vo, None, case["padding"], so), want)      # masks omitted = all voxels active
    H.create_index_grid_from_origins(got, 0.1)                                          # accepted as is: strictly increasing NanoVDB order


@pytest.mark.gpu
def test_device_domain_on_the_config4_topology_and_errors():
    import hnanosolver_b200 as H

    w = synth.WORKLOADS["c2"](with_coords=False)
    got = H.build_domain(w.origins, None, 1)
    # every voxel active: one voxel of padding = the 26-neighbour dilation of the leaf set
    cells = {tuple(c) for c in (w.origins // 8).tolist()}
    want = {(c[0] + a, c[1] + b, c[2] + d) for c in cells for a in (-1, 0, 1) for b in (-1, 0, 1) for d in (-1, 0, 1)}
    assert {tuple(c) for c in (got // 8).tolist()} == want and got.shape[0] == len(want)
    assert np.array_equal(H.build_domain(w.origins, None, 0), w.origins)
    assert H.build_domain(np.zeros((0, 3), np.int32)).shape == (0, 3)
    with pytest.raises(H.HnsError):
        H.build_domain(np.array([[3, 0, 0]], np.int32))
    with pytest.raises(H.HnsError):
        H.build_domain(np.array([[1 << 23, 0, 0]], np.int32))
    with pytest.raises(ValueError):
        H.build_domain(np.array([[0, 0, 0]], np.int32), None, -1)
