"""NanoVDB file round trip of the index grid (host only): the product's writer / reader against NanoVDB's own
dependency-free file IO (writeUncompressedGrid / readUncompressedGrids in the vendored NanoVDB.h, compiled into
oracle/_ref/libref_host.so) and against itself. No GPU: the grids come from the oracle's CPU emitter and from NanoVDB's host builder."""
import ctypes as C
import os

import numpy as np
import pytest

from hnanosolver_b200 import io as hio
from hnanosolver_b200 import synth
from hnanosolver_b200.grid_data import FLOAT, VEC3F, GridIndexedData
from oracle import oracle as O

needs_refhost = pytest.mark.skipif(not O.ref_host_available(), reason="oracle/_ref/libref_host.so not built (needs /root/reference)")


def _cases():
    return {"soup": synth.random_leaves(28, 4, 7, offset=(-24, 4096 - 16, -16), cfl=1.8, S=2), "sphere": synth.smoke_sphere(32, 1),
            "single": synth.random_leaves(1, 1, 1, S=1)}


@pytest.fixture(scope="module", params=list(_cases()))
def w(request):
    return _cases()[request.param]


def _refhost():
    L = O.refhost()
    L.refhost_write_nvdb.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    L.refhost_write_nvdb.restype = C.c_int
    L.refhost_read_nvdb.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    L.refhost_read_nvdb.restype = C.c_uint64
    return L


def test_round_trip_and_leaf_origins(w, tmp_path):
    buf = O.OracleIndex(w.coords).nanovdb_buffer(w.voxel_size)
    path = str(tmp_path / "grid.nvdb")
    hio.write_nvdb(path, buf)
    back = hio.read_nvdb(path)
    assert np.array_equal(back, buf)
    assert os.path.getsize(path) == 16 + 176 + 1 + buf.size          # FileHeader + FileMetaData + empty name + grid
    origins, h = hio.leaf_origins(back)
    assert np.array_equal(origins, w.origins) and h == np.float32(w.voxel_size)
    # a raw dump of the buffer (what GridHandle::write produces since 32.6) is accepted too
    raw = str(tmp_path / "raw.nvdb")
    buf.tofile(raw)
    assert np.array_equal(hio.read_nvdb(raw), buf)


@needs_refhost
def test_writer_matches_nanovdb_byte_for_byte(w, tmp_path):
    buf = O.OracleIndex(w.coords).nanovdb_buffer(w.voxel_size)
    mine, theirs = str(tmp_path / "mine.nvdb"), str(tmp_path / "theirs.nvdb")
    hio.write_nvdb(mine, buf)
    assert _refhost().refhost_write_nvdb(theirs.encode(), buf.ctypes.data_as(C.c_void_p), 0) == 0
    assert open(mine, "rb").read() == open(theirs, "rb").read()


@needs_refhost
def test_nanovdb_reads_what_the_product_writes_and_vice_versa(w, tmp_path):
    R = _refhost()
    # a grid made by NanoVDB's own host builder, written by NanoVDB, read by the product
    g = O.RefHostGrid(w.coords)
    theirs = str(tmp_path / "theirs.nvdb")
    native = g.buffer()
    assert R.refhost_write_nvdb(theirs.encode(), native.ctypes.data_as(C.c_void_p), 0) == 0
    got = hio.read_nvdb(theirs)
    assert np.array_equal(got, native)
    origins, _ = hio.leaf_origins(got)
    assert np.array_equal(origins, w.origins)
    # ... and as a raw dump
    raw = str(tmp_path / "theirs_raw.nvdb")
    assert R.refhost_write_nvdb(raw.encode(), native.ctypes.data_as(C.c_void_p), 1) == 0
    assert np.array_equal(hio.read_nvdb(raw), native)
    # written by the product, read by NanoVDB
    mine = str(tmp_path / "mine.nvdb")
    hio.write_nvdb(mine, native)
    out = np.empty_like(native)
    assert R.refhost_read_nvdb(mine.encode(), out.ctypes.data_as(C.c_void_p), out.size) == native.size
    assert np.array_equal(out, native)


def test_errors(tmp_path):
    from hnanosolver_b200._lib import HnsError

    junk = str(tmp_path / "junk.nvdb")
    open(junk, "wb").write(b"not a nanovdb file" * 50)
    with pytest.raises(HnsError):
        hio.read_nvdb(junk)
    with pytest.raises(HnsError):
        hio.read_nvdb(str(tmp_path / "missing.nvdb"))
    with pytest.raises(HnsError):
        hio.write_nvdb(str(tmp_path / "x.nvdb"), np.zeros(100, np.uint8))
    w = synth.random_leaves(1, 1, 1, S=1)
    buf = O.OracleIndex(w.coords).nanovdb_buffer(w.voxel_size)
    short = str(tmp_path / "short.nvdb")
    hio.write_nvdb(short, buf)
    data = open(short, "rb").read()
    open(short, "wb").write(data[:-100])
    with pytest.raises(HnsError):
        hio.read_nvdb(short)
    wrong = buf.copy()
    wrong[636] = 1                                                   # GridType::Float
    with pytest.raises(HnsError):
        hio.leaf_origins(wrong)


def test_cache_round_trip(w, tmp_path):
    d = GridIndexedData()
    d.allocateCoords(w.num_voxels)
    d.pCoords()[:] = w.coords
    d.addValueBlock(VEC3F, "vel")
    d.pValues(VEC3F, "vel")[:] = w.velocity
    for n, s in zip(w.scalar_names, w.scalars):
        d.addValueBlock(FLOAT, n)
        d.pValues(FLOAT, n)[:] = s
    hio.save_cache(str(tmp_path / "frame0001"), O.OracleIndex(w.coords).nanovdb_buffer(w.voxel_size), d)
    origins, h, back = hio.load_cache(str(tmp_path / "frame0001"))
    assert np.array_equal(origins, w.origins) and h == np.float32(w.voxel_size)
    assert np.array_equal(back.pCoords(), w.coords)
    assert back.getBlocksOfType(FLOAT) == list(w.scalar_names) and back.getBlocksOfType(VEC3F) == ["vel"]
    assert np.array_equal(back.pValues(VEC3F, "vel"), w.velocity)
    for n, s in zip(w.scalar_names, w.scalars):
        assert np.array_equal(back.pValues(FLOAT, n), s)
