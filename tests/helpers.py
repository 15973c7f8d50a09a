"""Shared helpers of the parity tests."""
import numpy as np

# Tolerance of the north star: advected / projected fields within 1e-5 relative of the reference kernels.
# "Relative" is measured against the field's max magnitude (a voxel-wise relative error is meaningless at zero crossings).
REL_TOL = 1e-5


def rel_err(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def assert_close(a, b, what, tol=REL_TOL, exact=True):
    """The north star's tolerance (1e-5 of the field's magnitude) is the contract; what the product actually delivers -- and what
    DESIGN.md claims -- is equality float for float with the oracle and with the reference's own kernels, so that is asserted too
    (`exact`; +0.0 == -0.0). A one-ulp regression in a kernel rewrite fails here, not three orders of magnitude later."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    if exact:
        bad = a != b
        if bad.any():
            ulp = np.abs(a.astype(np.float32).view(np.int32).astype(np.int64) - b.astype(np.float32).view(np.int32).astype(np.int64))[bad]
            raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} values differ bitwise (max {int(ulp.max())} ulp, relative error {e:.3e})")


# Byte ranges of the NanoVDB buffer that voxelsToGrid leaves uninitialised: grid name bytes 1..255,
# RootData tail padding, root-tile tail (state padding + value). Everything else must be bit-identical.
def nanovdb_compare_mask(nbytes: int, num_tiles: int) -> np.ndarray:
    m = np.ones(nbytes, bool)
    m[41:296] = False                       # GridData::mGridName[1..255]
    root = 672 + 64
    m[root + 28:root + 32] = False          # padding after mTableSize
    m[root + 72:root + 96] = False          # RootData tail padding
    for t in range(num_tiles):
        tile = root + 96 + 32 * t
        m[tile + 20:tile + 32] = False      # padding after state + the untouched tile value
    return m
