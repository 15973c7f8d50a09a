"""Sampler semantics: the reference's own known answers (Tests/IndexGrid.cpp) + the oracle against the reference's
__hostdev__ samplers running on the CPU over a real NanoVDB host grid."""
import numpy as np
import pytest

from oracle import oracle as O

needs_refhost = pytest.mark.skipif(not O.ref_host_available(), reason="oracle/_ref/libref_host.so not built (needs /root/reference)")


def _line(axis, n=15):
    c = np.zeros((n, 3), np.int32)
    c[:, axis] = np.arange(n)
    return c


def test_trilinear_known_answers():
    # Tests/IndexGrid.cpp:157-224: density(i,0,0) = i -> f(0.5,0,0) = 0.5, f(1.25,0,0) = 1.25, inactive (1,1,0) -> 0
    pts = _line(0)
    ix = O.OracleIndex(pts)
    data = ix.get_values(pts).astype(np.float32) * 0
    data[ix.get_values(pts).astype(np.int64) - 1] = np.arange(15, dtype=np.float32)
    got = ix.trilinear_f(data, [[0.5, 0, 0], [1.25, 0, 0], [5.0, 0, 0]])
    assert got.tolist() == [0.5, 1.25, 5.0]
    assert ix.nearest_f(data, [[1, 1, 0]])[0] == 0.0
    # Tests/IndexGrid.cpp:226-282: nearest (i,0,0) -> i ; Vec3f at (0,i,0) -> (i,i,i)
    assert ix.nearest_f(data, pts).tolist() == list(range(15))
    pv = _line(1)
    iv = O.OracleIndex(pv)
    vdata = np.zeros((15, 3), np.float32)
    vdata[iv.get_values(pv).astype(np.int64) - 1] = np.arange(15, dtype=np.float32)[:, None]
    assert iv.trilinear_v(vdata, [[0, 5.0, 0]])[0].tolist() == [5.0, 5.0, 5.0]
    # Tests/IndexGrid.cpp:84-155: values scaled x2 -> sample(5.5) == 11
    assert ix.trilinear_f(data * 2, [[5.5, 0, 0]])[0] == 11.0
    assert iv.trilinear_v(vdata * 2, [[0, 5.5, 0]])[0][0] == 11.0


def test_floor_rounds_down_for_negative_positions():
    # Floor() uses __float2int_rd (Stencils.hpp:25-43): -0.25 -> cell -1, fraction 0.75
    pts = np.array([[-1, 0, 0], [0, 0, 0]], np.int32)
    ix = O.OracleIndex(pts)
    data = np.zeros(2, np.float32)
    data[ix.get_values(pts).astype(np.int64) - 1] = [10.0, 20.0]
    assert ix.trilinear_f(data, [[-0.25, 0, 0]])[0] == np.float32(10.0) + np.float32(0.75) * np.float32(10.0)


@needs_refhost
def test_samplers_match_reference_host_code():
    rng = np.random.default_rng(11)
    pts = np.unique(rng.integers(-12, 30, size=(6000, 3)).astype(np.int32), axis=0)
    ix, rh = O.OracleIndex(pts), O.RefHostGrid(pts)
    n = ix.num_active
    f = rng.standard_normal(n).astype(np.float32)
    v = rng.standard_normal((n, 3)).astype(np.float32)
    q = pts[:400] + rng.integers(-1, 2, size=(400, 3)).astype(np.int32)
    assert np.array_equal(ix.nearest_f(f, q), rh.nearest_f(f, q))          # integer lookups + loads: bit exact
    xyz = (pts[:400] + rng.uniform(-1.5, 1.5, size=(400, 3))).astype(np.float32)
    # The oracle follows the DEVICE arithmetic (fused a + w*(b-a), see hns_oracle.c header); the host build of the
    # reference samplers is unfused, so agreement is to rounding, not bitwise: tolerance 4 ulp of the data range.
    tol = 4 * np.finfo(np.float32).eps * 4.0
    assert np.abs(ix.trilinear_f(f, xyz) - rh.trilinear_f(f, xyz)).max() <= tol
    assert np.abs(ix.trilinear_v(v, xyz) - rh.trilinear_v(v, xyz)).max() <= tol
