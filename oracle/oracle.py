"""ctypes bindings for the CPU oracle (oracle/hns_oracle.c) and for the reference builds under oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package (hnanosolver_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libhns_ref.so")
_REFHOST_SO = os.path.join(_HERE, "_ref", "libref_host.so")

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)
c_u8p = C.POINTER(C.c_uint8)


def build(ref: bool | None = None) -> None:
    """Compile liboracle.so (always) and, when /root/reference is present, the oracle/_ref/ libraries."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if ref is None:
        ref = os.path.isdir("/root/reference/src/Cuda")
    if ref:
        subprocess.check_call(["make", "-s", "-j4", "-C", _HERE, "ref"])
        if os.path.exists(os.path.join(os.path.dirname(_HERE), "compat", "libhns_compat.so")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/libcompat_driver.so"])


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(c_f32p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        L = C.CDLL(_ORACLE_SO)
        L.ora_index_create.restype = C.c_void_p
        L.ora_index_create.argtypes = [c_i32p, C.c_uint64]
        L.ora_index_destroy.argtypes = [C.c_void_p]
        L.ora_index_num_leaves.restype = C.c_int64
        L.ora_index_num_leaves.argtypes = [C.c_void_p]
        L.ora_index_num_active.restype = C.c_uint64
        L.ora_index_num_active.argtypes = [C.c_void_p]
        L.ora_index_leaf_origins.argtypes = [C.c_void_p, c_i32p]
        L.ora_get_values.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_u64p]
        L.ora_nanovdb_size.restype = C.c_uint64
        L.ora_nanovdb_size.argtypes = [C.c_void_p]
        L.ora_nanovdb_emit.argtypes = [C.c_void_p, C.c_float, c_u8p]
        L.ora_trilinear_f.restype = C.c_float
        L.ora_trilinear_f.argtypes = [C.c_void_p, c_f32p, C.c_float, C.c_float, C.c_float]
        L.ora_trilinear_v.argtypes = [C.c_void_p, c_f32p, C.c_float, C.c_float, C.c_float, c_f32p]
        L.ora_nearest_float.restype = C.c_float
        L.ora_nearest_float.argtypes = [C.c_void_p, c_f32p, C.c_int32, C.c_int32, C.c_int32]
        L.ora_advect_vector.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, C.c_uint64, C.c_float, C.c_float]
        L.ora_advect_scalar.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, c_f32p, C.c_uint64, C.c_float, C.c_float]
        L.ora_advect_scalars.argtypes = [C.c_void_p, c_i32p, c_f32p, C.POINTER(c_f32p), C.POINTER(c_f32p), C.c_int, C.c_uint64,
                                         C.c_float, C.c_float]
        L.ora_divergence.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, C.c_float, C.c_uint64]
        L.ora_vorticity_confinement.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ora_rbgs.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, C.c_float, C.c_uint64, C.c_int, C.c_float]
        L.ora_subtract_gradient.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, c_f32p, c_f32p, C.c_float]
        L.ora_omega_compute.restype = C.c_float
        L.ora_omega_compute.argtypes = [C.c_float]
        L.ora_omega_project.restype = C.c_float
        L.ora_omega_project.argtypes = [C.c_float]
        L.ora_frame.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, C.POINTER(c_f32p), C.c_int, C.c_int, C.c_float, C.c_float,
                                c_f32p, c_f32p, c_f32p]
        L.ora_compute_sim_collision.restype = C.c_int
        L.ora_compute_sim_collision.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, C.POINTER(c_f32p), C.POINTER(C.c_char_p), C.c_int,
                                                C.c_int, C.c_float, C.c_float, c_f32p, C.c_int]
        L.ora_collision_boundary.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, c_f32p, C.c_float, C.c_float, C.c_int, C.c_uint64]
        L.ora_advect_vector_sdf.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, C.c_uint64, C.c_float, C.c_float, c_f32p]
        L.ora_advect_scalars_sdf.argtypes = [C.c_void_p, c_i32p, c_f32p, C.POINTER(c_f32p), C.POINTER(c_f32p), C.c_int, C.c_uint64,
                                             C.c_float, C.c_float, c_f32p]
        L.ora_compute_sim.restype = C.c_int
        L.ora_compute_sim.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, C.POINTER(c_f32p), C.POINTER(C.c_char_p), C.c_int,
                                      C.c_int, C.c_float, C.c_float, c_f32p]
        L.ora_project_non_divergent.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, C.c_int, C.c_float, c_f32p, c_f32p]
        L.ora_num_threads.restype = C.c_int
        c_f64p = C.POINTER(C.c_double)
        L.ora_mg_residual.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, c_f32p, c_f32p, C.c_float, c_f32p]
        L.ora_mg_diag.argtypes = [C.c_void_p, c_i32p, C.c_uint64, C.c_float, c_f32p]
        L.ora_mg_rbgs.argtypes = [C.c_void_p, c_i32p, c_f32p, c_f32p, c_f32p, C.c_float, C.c_uint64, C.c_int, C.c_float]
        L.ora_residual_sums_f64.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_f32p, c_f32p, C.c_double, c_f64p]
        L.ora_mg_restrict.argtypes = [C.c_void_p, c_f32p, c_i32p, C.c_uint64, c_f32p]
        L.ora_mg_prolong_add.argtypes = [C.c_void_p, c_f32p, c_i32p, C.c_uint64, c_f32p]
        L.ora_sum_squares_f64.restype = C.c_double
        L.ora_sum_squares_f64.argtypes = [c_f32p, C.c_uint64]
        _lib = L
    return _lib


def _ptr_array(arrs):
    P = (c_f32p * len(arrs))()
    for i, a in enumerate(arrs):
        P[i] = a.ctypes.data_as(c_f32p)
    return P


class OracleIndex:
    """voxelsToGrid<ValueOnIndex> restated on the CPU (any voxel list: unsorted / partial leaves allowed)."""

    def __init__(self, coords):
        self.coords, p = _i32(np.asarray(coords).reshape(-1, 3))
        self.n = self.coords.shape[0]
        self._h = lib().ora_index_create(p, self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_index_destroy(self._h)
            self._h = None

    @property
    def num_leaves(self) -> int:
        return lib().ora_index_num_leaves(self._h)

    @property
    def num_active(self) -> int:
        return lib().ora_index_num_active(self._h)

    def leaf_origins(self) -> np.ndarray:
        out = np.empty((self.num_leaves, 3), np.int32)
        lib().ora_index_leaf_origins(self._h, out.ctypes.data_as(c_i32p))
        return out

    def get_values(self, ijk) -> np.ndarray:
        ijk, p = _i32(np.asarray(ijk).reshape(-1, 3))
        out = np.empty(ijk.shape[0], np.uint64)
        lib().ora_get_values(self._h, p, ijk.shape[0], out.ctypes.data_as(c_u64p))
        return out

    def nanovdb_buffer(self, voxel_size: float) -> np.ndarray:
        buf = np.empty(lib().ora_nanovdb_size(self._h), np.uint8)
        lib().ora_nanovdb_emit(self._h, C.c_float(voxel_size), buf.ctypes.data_as(c_u8p))
        return buf

    # --- samplers -------------------------------------------------------------------------------
    def trilinear_f(self, data, xyz) -> np.ndarray:
        data, dp = _f32(data)
        xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
        return np.array([lib().ora_trilinear_f(self._h, dp, *map(float, q)) for q in xyz], np.float32)

    def trilinear_v(self, data, xyz) -> np.ndarray:
        data, dp = _f32(data)
        xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
        out = np.empty((xyz.shape[0], 3), np.float32)
        for i, q in enumerate(xyz):
            lib().ora_trilinear_v(self._h, dp, float(q[0]), float(q[1]), float(q[2]), out[i].ctypes.data_as(c_f32p))
        return out

    def nearest_f(self, data, ijk) -> np.ndarray:
        data, dp = _f32(data)
        ijk = np.asarray(ijk, np.int32).reshape(-1, 3)
        return np.array([lib().ora_nearest_float(self._h, dp, int(q[0]), int(q[1]), int(q[2])) for q in ijk], np.float32)

    # --- kernels (out-of-place, return new arrays) ------------------------------------------------
    def advect_vector(self, vel, dt, voxel_size):
        vel, vp = _f32(vel)
        out = np.empty_like(vel)
        lib().ora_advect_vector(self._h, self.coords.ctypes.data_as(c_i32p), vp, out.ctypes.data_as(c_f32p), self.n, dt,
                                np.float32(1.0) / np.float32(voxel_size))
        return out

    def advect_scalar(self, vel, phi, dt, voxel_size):
        vel, vp = _f32(vel)
        phi, pp = _f32(phi)
        out = np.empty_like(phi)
        lib().ora_advect_scalar(self._h, self.coords.ctypes.data_as(c_i32p), vp, pp, out.ctypes.data_as(c_f32p), self.n, dt,
                                np.float32(1.0) / np.float32(voxel_size))
        return out

    def advect_scalars(self, vel, phis, dt, voxel_size):
        vel, vp = _f32(vel)
        ins = [np.ascontiguousarray(a, np.float32) for a in phis]
        outs = [np.empty_like(a) for a in ins]
        lib().ora_advect_scalars(self._h, self.coords.ctypes.data_as(c_i32p), vp, _ptr_array(ins), _ptr_array(outs), len(ins), self.n,
                                 dt, np.float32(1.0) / np.float32(voxel_size))
        return outs

    def divergence(self, vel, voxel_size):
        vel, vp = _f32(vel)
        out = np.empty(self.n, np.float32)
        lib().ora_divergence(self._h, self.coords.ctypes.data_as(c_i32p), vp, out.ctypes.data_as(c_f32p),
                             np.float32(1.0) / np.float32(voxel_size), self.n)
        return out

    def vorticity_confinement(self, vel, dt, voxel_size, scale, factor_scale):
        """out-of-place vorticityConfinement (Kernel.cu:969-1025) of an (N, 3) velocity"""
        vel, vp = _f32(vel)
        out = np.empty((self.n, 3), np.float32)
        lib().ora_vorticity_confinement(self._h, self.coords.ctypes.data_as(c_i32p), vp, out.ctypes.data_as(c_f32p), self.n, dt,
                                        np.float32(1.0) / np.float32(voxel_size), scale, factor_scale)
        return out

    def rbgs(self, div, p, voxel_size, iterations, omega):
        """iterations x (red, black) in place on a copy of p."""
        div, dp = _f32(div)
        p = np.array(p, np.float32, copy=True)
        for _ in range(iterations):
            for color in (0, 1):
                lib().ora_rbgs(self._h, self.coords.ctypes.data_as(c_i32p), dp, p.ctypes.data_as(c_f32p), voxel_size, self.n, color,
                               omega)
        return p

    def rbgs_color(self, div, p, voxel_size, color, omega):
        """one half-sweep in place on p (a float32 array)"""
        div, dp = _f32(div)
        assert p.dtype == np.float32 and p.flags.c_contiguous
        lib().ora_rbgs(self._h, self.coords.ctypes.data_as(c_i32p), dp, p.ctypes.data_as(c_f32p), voxel_size, self.n, color, omega)
        return p

    def subtract_gradient(self, vel, p, voxel_size):
        vel, vp = _f32(vel)
        p, pp = _f32(p)
        out = np.empty_like(vel)
        lib().ora_subtract_gradient(self._h, self.coords.ctypes.data_as(c_i32p), self.n, vp, pp, out.ctypes.data_as(c_f32p),
                                    np.float32(1.0) / np.float32(voxel_size))
        return out

    def frame(self, vel, scalars, iterations, dt, voxel_size):
        """The north-star frame. Returns dict(vel, scalars, div, p, adv)."""
        vel = np.array(vel, np.float32, copy=True)
        sc = [np.array(a, np.float32, copy=True) for a in scalars]
        div = np.empty(self.n, np.float32)
        p = np.empty(self.n, np.float32)
        adv = np.empty((self.n, 3), np.float32)
        lib().ora_frame(self._h, self.coords.ctypes.data_as(c_i32p), self.n, vel.ctypes.data_as(c_f32p), _ptr_array(sc), len(sc),
                        iterations, dt, voxel_size, div.ctypes.data_as(c_f32p), p.ctypes.data_as(c_f32p), adv.ctypes.data_as(c_f32p))
        return dict(vel=vel, scalars=sc, div=div, p=p, adv=adv)

    # ---- the hasCollision path ----
    SITE_ENFORCE, SITE_ADVECT, SITE_GRADIENT = 0, 1, 2

    def collision_boundary(self, vel, sdf, voxel_size, site):
        """enforceCollisionBoundaries (site 0), or the boundary tail of advect_vector (1) / subtractPressureGradient (2) applied to `vel`"""
        vel, vp = _f32(vel)
        sdf, sp = _f32(sdf)
        out = np.empty((self.n, 3), np.float32)
        lib().ora_collision_boundary(self._h, self.coords.ctypes.data_as(c_i32p), vp, out.ctypes.data_as(c_f32p), sp,
                                     np.float32(1.0) / np.float32(voxel_size), 1.5 if site == 1 else 0.1, site, self.n)
        return out

    def advect_vector_sdf(self, vel, sdf, dt, voxel_size):
        vel, vp = _f32(vel)
        sdf, sp = _f32(sdf)
        out = np.empty((self.n, 3), np.float32)
        lib().ora_advect_vector_sdf(self._h, self.coords.ctypes.data_as(c_i32p), vp, out.ctypes.data_as(c_f32p), self.n, dt,
                                    np.float32(1.0) / np.float32(voxel_size), sp)
        return out

    def advect_scalars_sdf(self, vel, scalars, sdf, dt, voxel_size):
        vel, vp = _f32(vel)
        sdf, sp = _f32(sdf)
        ins = [np.ascontiguousarray(a, np.float32) for a in scalars]
        outs = [np.empty(self.n, np.float32) for _ in ins]
        lib().ora_advect_scalars_sdf(self._h, self.coords.ctypes.data_as(c_i32p), vp, _ptr_array(ins), _ptr_array(outs), len(ins), self.n, dt,
                                     np.float32(1.0) / np.float32(voxel_size), sp)
        return outs

    def compute_sim(self, vel, fields: dict, iterations, dt, voxel_size, params6, has_collision=False):
        vel = np.array(vel, np.float32, copy=True)
        names = list(fields.keys())
        sc = [np.array(fields[k], np.float32, copy=True) for k in names]
        cn = (C.c_char_p * len(names))(*[s.encode() for s in names])
        pr = np.asarray(params6, np.float32)
        rc = lib().ora_compute_sim_collision(self._h, self.coords.ctypes.data_as(c_i32p), self.n, vel.ctypes.data_as(c_f32p), _ptr_array(sc), cn,
                                             len(sc), iterations, dt, voxel_size, pr.ctypes.data_as(c_f32p), int(has_collision))
        if rc:
            raise RuntimeError("Missing required input field for combustion")
        return vel, dict(zip(names, sc))

    def project_non_divergent(self, vel, iterations, voxel_size):
        vel = np.array(vel, np.float32, copy=True)
        div = np.empty(self.n, np.float32)
        p = np.empty(self.n, np.float32)
        lib().ora_project_non_divergent(self._h, self.coords.ctypes.data_as(c_i32p), self.n, vel.ctypes.data_as(c_f32p), iterations,
                                        voxel_size, div.ctypes.data_as(c_f32p), p.ctypes.data_as(c_f32p))
        return vel, div, p


def domain_leaves(vel_origins, vel_masks, padding: int, sdf_origins=None) -> np.ndarray:
    """Host set-based restatement of the domain construction of SOP_HNanoSolverVerb::cook (reference
    src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:188-199): topologyUnion(velocity tree) -> every leaf node of the velocity grid;
    dilateVoxels(padding, NN_FACE_EDGE_VERTEX) -> every leaf an active voxel reaches within `padding` voxels in the Chebyshev metric
    (`padding` rounds of 26-neighbour dilation); topologyUnion(sdf tree) -> every leaf node of the SDF grid. Returns the leaf origins in
    NanoVDB order. PARITY UNPINNED against OpenVDB itself (an un-vendored dependency): this restates its documented
    semantics voxel by voxel.  vel_masks: uint64 (n, 8), word x, bit y*8+z; None = all active."""
    vo = np.asarray(vel_origins, np.int64).reshape(-1, 3)
    leaves = {tuple(o) for o in vo.tolist()}
    if sdf_origins is not None:
        leaves |= {tuple(o) for o in np.asarray(sdf_origins, np.int64).reshape(-1, 3).tolist()}
    p = int(padding)
    for l in range(vo.shape[0]):
        if vel_masks is None:
            bits = np.ones(512, bool)
        else:
            words = np.asarray(vel_masks[l], np.uint64)
            bits = ((words[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool).reshape(512)
        n = np.nonzero(bits)[0]
        if n.size == 0:
            continue
        v = vo[l][None, :] + np.stack([n >> 6, (n >> 3) & 7, n & 7], 1)          # active voxels, global coordinates
        lo, hi = (v - p) >> 3, (v + p) >> 3                                       # leaf range the dilated voxel touches, per axis
        for a, b in {(tuple(x), tuple(y)) for x, y in zip(lo.tolist(), hi.tolist())}:
            for X in range(a[0], b[0] + 1):
                for Y in range(a[1], b[1] + 1):
                    for Z in range(a[2], b[2] + 1):
                        leaves.add((8 * X, 8 * Y, 8 * Z))
    out = np.array(sorted(leaves), np.int32).reshape(-1, 3)
    c = out.astype(np.int64)
    b = c + (1 << 31)
    tile = ((b[:, 0] >> 12) << 42) | ((b[:, 1] >> 12) << 21) | (b[:, 2] >> 12)
    up = (((c[:, 0] & 4095) >> 7) << 10) | (((c[:, 1] & 4095) >> 7) << 5) | ((c[:, 2] & 4095) >> 7)
    lo_ = (((c[:, 0] & 127) >> 3) << 8) | (((c[:, 1] & 127) >> 3) << 4) | ((c[:, 2] & 127) >> 3)
    return np.ascontiguousarray(out[np.lexsort((lo_, up, tile))])


def sum_squares(a) -> float:
    a, ap = _f32(np.asarray(a).reshape(-1))
    return float(lib().ora_sum_squares_f64(ap, a.size))


def nanovdb_value_order(coords: np.ndarray) -> np.ndarray:
    """permutation that puts a voxel list into NanoVDB ValueOnIndex order (root tile, upper offset, lower offset, voxel offset;
    externals/nanovdb/tools/cuda/PointsToGrid.cuh:596-645): the order in which the oracle's samplers index a sidecar array"""
    c = np.asarray(coords, np.int64).reshape(-1, 3)
    b = c + (1 << 31)
    tile = ((b[:, 0] >> 12) << 42) | ((b[:, 1] >> 12) << 21) | (b[:, 2] >> 12)
    up = (((c[:, 0] & 4095) >> 7) << 10) | (((c[:, 1] & 4095) >> 7) << 5) | ((c[:, 2] & 4095) >> 7)
    lo = (((c[:, 0] & 127) >> 3) << 8) | (((c[:, 1] & 127) >> 3) << 4) | ((c[:, 2] & 127) >> 3)
    vox = ((c[:, 0] & 7) << 6) | ((c[:, 1] & 7) << 3) | (c[:, 2] & 7)
    return np.lexsort((vox, lo, up, tile))


class OracleMultigrid:
    """CPU restatement of the product's multigrid pressure solve (hnanosolver_b200/csrc/multigrid.cu), level by level on plain voxel
    lists. PARITY UNPINNED against the reference (its v_cycle is dead code, src/Cuda/HNanoSolver.cu:399-507); the checks are the fp64
    Poisson residual and the divergence of the projected velocity."""

    def __init__(self, coords, voxel_size: float, max_levels: int = 16, coarsest_iterations: int = 32, coarsest_omega: float = 1.5,
                 rule: str = "all"):
        self.levels = []            # (OracleIndex, dx)
        self.diag = []              # per level: None (fine level: 6 everywhere) or the operator's diagonal
        c = np.ascontiguousarray(np.asarray(coords, np.int32).reshape(-1, 3))
        dx = np.float32(voxel_size)
        while True:
            ix = OracleIndex(c)
            k = len(self.levels)
            self.levels.append((ix, dx))
            if k == 0:
                self.diag.append(None)
            else:
                d = np.empty(ix.n, np.float32)
                theta = 0.5 + 2.0 ** (-(k + 1))
                lib().ora_mg_diag(ix._h, ix.coords.ctypes.data_as(c_i32p), ix.n, np.float32(1.0 / theta - 1.0), d.ctypes.data_as(c_f32p))
                self.diag.append(d)
            leaves = np.unique(c >> 3, axis=0).shape[0]
            if leaves <= 1 or len(self.levels) >= max_levels:
                break
            # a cell belongs to the next level iff all 8 of its children belong to this one ("any": iff one of them does -- kept for
            # the record: it makes coarse domains too large around thin features and the cycle overshoots there)
            c, cnt = np.unique(c >> 1, axis=0, return_counts=True)
            c = c[cnt == 8] if rule == "all" else c
            if c.shape[0] == 0:
                break
            c = c.astype(np.int32)
            c = np.ascontiguousarray(c[nanovdb_value_order(c)])  # sidecar order == index order, as every kernel assumes
            dx = np.float32(dx * np.float32(2.0))
        self.coarsest_iterations, self.coarsest_omega = coarsest_iterations, np.float32(coarsest_omega)

    def _smooth(self, k, rhs, p, iters, omega):
        ix, dx = self.levels[k]
        for _ in range(iters):
            for color in (0, 1):
                if k == 0:
                    ix.rbgs_color(rhs, p, dx, color, omega)      # the reference's update (src/Cuda/Kernel.cu:591-623)
                else:
                    lib().ora_mg_rbgs(ix._h, ix.coords.ctypes.data_as(c_i32p), rhs.ctypes.data_as(c_f32p), p.ctypes.data_as(c_f32p),
                                      self.diag[k].ctypes.data_as(c_f32p), dx, ix.n, color, omega)

    def residual(self, k, p, rhs):
        ix, dx = self.levels[k]
        r = np.empty(ix.n, np.float32)
        lib().ora_mg_residual(ix._h, ix.coords.ctypes.data_as(c_i32p), ix.n, p.ctypes.data_as(c_f32p), rhs.ctypes.data_as(c_f32p),
                              self.diag[k].ctypes.data_as(c_f32p) if k else None, dx, r.ctypes.data_as(c_f32p))
        return r

    def v_cycle(self, p, rhs, nu_pre, nu_post, omega):
        n = len(self.levels)
        P, F = [p] + [None] * (n - 1), [rhs] + [None] * (n - 1)
        for k in range(n - 1):
            self._smooth(k, F[k], P[k], nu_pre, omega)
            r = self.residual(k, P[k], F[k])
            ixc = self.levels[k + 1][0]
            F[k + 1] = np.empty(ixc.n, np.float32)
            lib().ora_mg_restrict(self.levels[k][0]._h, r.ctypes.data_as(c_f32p), ixc.coords.ctypes.data_as(c_i32p), ixc.n,
                                  F[k + 1].ctypes.data_as(c_f32p))
            P[k + 1] = np.zeros(ixc.n, np.float32)
        if n > 1:
            one_leaf = np.unique(self.levels[-1][0].coords >> 3, axis=0).shape[0] == 1
            self._smooth(n - 1, F[-1], P[-1], self.coarsest_iterations, self.coarsest_omega if one_leaf else omega)
        else:
            self._smooth(0, F[0], P[0], nu_pre + nu_post, omega)
        for k in range(n - 2, -1, -1):
            ixf = self.levels[k][0]
            lib().ora_mg_prolong_add(self.levels[k + 1][0]._h, P[k + 1].ctypes.data_as(c_f32p), ixf.coords.ctypes.data_as(c_i32p), ixf.n,
                                     P[k].ctypes.data_as(c_f32p))
            self._smooth(k, F[k], P[k], nu_post, omega)

    def residual_sums(self, p, rhs):
        """{sum (rhs - L p)^2, sum rhs^2} of the fine level, fp64 arithmetic on the fp32 fields"""
        ix, dx = self.levels[0]
        p, pp = _f32(p)
        rhs, rp = _f32(rhs)
        out = np.zeros(2, np.float64)
        lib().ora_residual_sums_f64(ix._h, ix.coords.ctypes.data_as(c_i32p), ix.n, pp, rp, float(dx), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def solve(self, div, cycles, nu_pre=2, nu_post=2, omega=1.0, rel_tol=0.0):
        div = np.ascontiguousarray(div, np.float32)
        p = np.zeros_like(div)
        rel = None
        done = 0
        for _ in range(cycles):
            self.v_cycle(p, div, nu_pre, nu_post, np.float32(omega))
            done += 1
            if rel_tol > 0:
                a, b = self.residual_sums(p, div)
                rel = float(np.sqrt(a / b)) if b > 0 else 0.0
                if rel <= rel_tol:
                    break
        return p, done, rel


def omega_compute(voxel_size: float) -> float:
    return float(lib().ora_omega_compute(voxel_size))


def omega_project(voxel_size: float) -> float:
    return float(lib().ora_omega_project(voxel_size))


def num_threads() -> int:
    return lib().ora_num_threads()


# --------------------------------------------------------------------------------------------------
# oracle/_ref: the real reference
# --------------------------------------------------------------------------------------------------
def ref_host_available() -> bool:
    return os.path.exists(_REFHOST_SO)


def ref_gpu_available() -> bool:
    return os.path.exists(_REF_SO)


_refhost = None


def refhost() -> C.CDLL:
    global _refhost
    if _refhost is None:
        L = C.CDLL(_REFHOST_SO)
        L.refhost_create.restype = C.c_void_p
        L.refhost_create.argtypes = [c_i32p, C.c_uint64]
        L.refhost_destroy.argtypes = [C.c_void_p]
        for f in ("refhost_value_count", "refhost_leaf_count", "refhost_bytes"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.refhost_data.restype = C.c_void_p
        L.refhost_data.argtypes = [C.c_void_p]
        L.refhost_get_values.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_u64p]
        L.refhost_nearest_f.argtypes = [C.c_void_p, c_f32p, c_i32p, C.c_uint64, c_f32p]
        L.refhost_trilinear_f.argtypes = [C.c_void_p, c_f32p, c_f32p, C.c_uint64, c_f32p]
        L.refhost_trilinear_v.argtypes = [C.c_void_p, c_f32p, c_f32p, C.c_uint64, c_f32p]
        _refhost = L
    return _refhost


class RefHostGrid:
    """Real NanoVDB host ValueOnIndex grid + the reference's Stencils.hpp samplers (CPU)."""

    def __init__(self, coords):
        c, p = _i32(np.asarray(coords).reshape(-1, 3))
        self._h = refhost().refhost_create(p, c.shape[0])
        if not self._h:
            raise RuntimeError("refhost_create failed")

    def __del__(self):
        if getattr(self, "_h", None):
            refhost().refhost_destroy(self._h)
            self._h = None

    @property
    def value_count(self):
        return refhost().refhost_value_count(self._h)

    @property
    def leaf_count(self):
        return refhost().refhost_leaf_count(self._h)

    def buffer(self) -> np.ndarray:
        n = refhost().refhost_bytes(self._h)
        return np.ctypeslib.as_array(C.cast(refhost().refhost_data(self._h), c_u8p), shape=(n,)).copy()

    def get_values(self, ijk):
        ijk, p = _i32(np.asarray(ijk).reshape(-1, 3))
        out = np.empty(ijk.shape[0], np.uint64)
        refhost().refhost_get_values(self._h, p, ijk.shape[0], out.ctypes.data_as(c_u64p))
        return out

    def nearest_f(self, data, ijk):
        data, dp = _f32(data)
        ijk, p = _i32(np.asarray(ijk).reshape(-1, 3))
        out = np.empty(ijk.shape[0], np.float32)
        refhost().refhost_nearest_f(self._h, dp, p, ijk.shape[0], out.ctypes.data_as(c_f32p))
        return out

    def trilinear_f(self, data, xyz):
        data, dp = _f32(data)
        xyz, xp = _f32(np.asarray(xyz).reshape(-1, 3))
        out = np.empty(xyz.shape[0], np.float32)
        refhost().refhost_trilinear_f(self._h, dp, xp, xyz.shape[0], out.ctypes.data_as(c_f32p))
        return out

    def trilinear_v(self, data, xyz):
        data, dp = _f32(data)
        xyz, xp = _f32(np.asarray(xyz).reshape(-1, 3))
        out = np.empty((xyz.shape[0], 3), np.float32)
        refhost().refhost_trilinear_v(self._h, dp, xp, xyz.shape[0], out.ctypes.data_as(c_f32p))
        return out


# --------------------------------------------------------------------------------------------------
# oracle/_ref/libhns_ref.so: the unmodified reference launchers + kernels (needs a GPU)
# --------------------------------------------------------------------------------------------------
_ref = None
_ref_libs: dict = {}
_COMPAT_DRIVER_SO = os.path.join(_HERE, "_ref", "libcompat_driver.so")


def compat_driver_available() -> bool:
    return os.path.exists(_COMPAT_DRIVER_SO)


class reference_library:
    """Context manager selecting which implementation of the seven launcher symbols the Ref* helpers drive:
    "reference" = oracle/_ref/libhns_ref.so (the unmodified reference), "compat" = oracle/_ref/libcompat_driver.so
    (the same plain-C driver, oracle/ref_shim.cu, linked against compat/libhns_compat.so, i.e. the product)."""

    def __init__(self, which: str):
        self.path = {"reference": _REF_SO, "compat": _COMPAT_DRIVER_SO}[which]

    def __enter__(self):
        global _ref
        self._saved = _ref
        _ref = _load_ref(self.path)
        return self

    def __exit__(self, *a):
        global _ref
        _ref = self._saved


def ref() -> C.CDLL:
    global _ref
    if _ref is None:
        _ref = _load_ref(_REF_SO)
    return _ref


def _load_ref(path: str) -> C.CDLL:
    if path in _ref_libs:
        return _ref_libs[path]
    if True:
        L = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_data_create.restype = C.c_void_p
        L.ref_data_create.argtypes = [C.c_uint64, C.c_int]
        L.ref_data_destroy.argtypes = [C.c_void_p]
        L.ref_data_coords.restype = c_i32p
        L.ref_data_coords.argtypes = [C.c_void_p]
        L.ref_data_add_float.restype = c_f32p
        L.ref_data_add_float.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_data_add_vec3.restype = c_f32p
        L.ref_data_add_vec3.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_create_index_grid.argtypes = [C.c_void_p, C.c_float, C.POINTER(C.c_void_p)]
        L.ref_grid_destroy.argtypes = [C.c_void_p]
        L.ref_grid_bytes.restype = C.c_uint64
        L.ref_grid_bytes.argtypes = [C.c_void_p]
        L.ref_grid_download.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_grid_get_values.argtypes = [C.c_void_p, c_i32p, C.c_uint64, c_u64p]
        L.ref_compute_sim.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, c_f32p, C.c_int, C.POINTER(C.c_double)]
        L.ref_cook_frame.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, c_f32p, C.c_int, C.POINTER(C.c_double)]
        L.ref_advect_index_grid.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_advect_index_grid_velocity.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_project_non_divergent.argtypes = [C.c_void_p, C.c_uint64, C.c_float]
        L.ref_divergence.argtypes = [C.c_void_p, C.c_float]
        if hasattr(L, "ref_frame_create"):
            L.ref_frame_create.restype = C.c_void_p
            L.ref_frame_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
            L.ref_frame_destroy.argtypes = [C.c_void_p]
            L.ref_frame_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, c_f32p]
            L.ref_frame_download.argtypes = [C.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, C.POINTER(c_f32p)]
            if hasattr(L, "ref_frame_collision_stage"):
                L.ref_frame_collision_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_int, c_f32p, C.c_float, C.c_float, c_f32p]
            if hasattr(L, "ref_frame_vorticity"):
                L.ref_frame_vorticity.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, c_f32p]
        _ref_libs[path] = L
    return L


class _Quiet:
    """The reference prints a timing banner per step (src/Cuda/Utils.cuh:260-269): silence fd 1 around calls."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._saved, 1)
        os.close(self._null)
        os.close(self._saved)


def _rcheck(rc):
    if rc:
        raise RuntimeError("reference: " + ref().ref_last_error().decode(errors="replace"))


class RefData:
    """An HNS::GridIndexedData owned by the reference build; blocks are exposed as numpy views (zero copy)."""

    def __init__(self, coords, alloc_type: int = 0):
        coords = np.ascontiguousarray(np.asarray(coords, np.int32).reshape(-1, 3))
        self.n = coords.shape[0]
        self._lib = ref()
        self._h = ref().ref_data_create(self.n, alloc_type)
        if not self._h:
            raise RuntimeError("ref_data_create failed")
        self.coords = np.ctypeslib.as_array(ref().ref_data_coords(self._h), shape=(self.n, 3))
        self.coords[:] = coords
        self.blocks: dict[str, np.ndarray] = {}

    def __del__(self):
        if getattr(self, "_h", None):
            self.blocks = {}
            self._lib.ref_data_destroy(self._h)
            self._h = None

    def add_float(self, name: str, values) -> np.ndarray:
        p = ref().ref_data_add_float(self._h, name.encode())
        a = np.ctypeslib.as_array(p, shape=(self.n,))
        a[:] = values
        self.blocks[name] = a
        return a

    def add_vec3(self, name: str, values) -> np.ndarray:
        p = ref().ref_data_add_vec3(self._h, name.encode())
        a = np.ctypeslib.as_array(p, shape=(self.n, 3))
        a[:] = values
        self.blocks[name] = a
        return a


class RefGrid:
    def __init__(self, data: RefData, voxel_size: float):
        h = C.c_void_p()
        self._lib = ref()
        with _Quiet():
            rc = ref().ref_create_index_grid(data._h, voxel_size, C.byref(h))
        _rcheck(rc)
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_grid_destroy(self._h)
            self._h = None

    def buffer(self) -> np.ndarray:
        buf = np.empty(self._lib.ref_grid_bytes(self._h), np.uint8)
        _rcheck(self._lib.ref_grid_download(self._h, buf.ctypes.data_as(C.c_void_p)))
        return buf

    def get_values(self, ijk) -> np.ndarray:
        ijk, p = _i32(np.asarray(ijk).reshape(-1, 3))
        out = np.empty(ijk.shape[0], np.uint64)
        _rcheck(self._lib.ref_grid_get_values(self._h, p, ijk.shape[0], out.ctypes.data_as(c_u64p)))
        return out


def ref_compute_sim(data: RefData, grid: RefGrid, iterations, dt, voxel_size, params6, has_collision=False) -> float:
    pr = np.asarray(params6, np.float32)
    ms = C.c_double()
    with _Quiet():
        rc = ref().ref_compute_sim(data._h, grid._h, iterations, dt, voxel_size, pr.ctypes.data_as(c_f32p), int(has_collision), C.byref(ms))
    _rcheck(rc)
    return ms.value


def ref_cook_frame(data: RefData, iterations, dt, voxel_size, params6, has_collision=False) -> float:
    pr = np.asarray(params6, np.float32)
    ms = C.c_double()
    with _Quiet():
        rc = ref().ref_cook_frame(data._h, iterations, dt, voxel_size, pr.ctypes.data_as(c_f32p), int(has_collision), C.byref(ms))
    _rcheck(rc)
    return ms.value


def ref_advect_index_grid(data: RefData, dt, voxel_size):
    with _Quiet():
        rc = ref().ref_advect_index_grid(data._h, dt, voxel_size)
    _rcheck(rc)


def ref_advect_index_grid_velocity(data: RefData, dt, voxel_size):
    with _Quiet():
        rc = ref().ref_advect_index_grid_velocity(data._h, dt, voxel_size)
    _rcheck(rc)


def ref_project_non_divergent(data: RefData, iterations, voxel_size):
    with _Quiet():
        rc = ref().ref_project_non_divergent(data._h, iterations, voxel_size)
    _rcheck(rc)


def ref_divergence(data: RefData, voxel_size):
    with _Quiet():
        rc = ref().ref_divergence(data._h, voxel_size)
    _rcheck(rc)


class RefFrame:
    """The north-star frame with the reference's own kernels on device-resident buffers (oracle/ref_shim.cu)."""

    def __init__(self, data: RefData, grid: RefGrid, scalar_names):
        self.data, self.grid, self.names = data, grid, list(scalar_names)
        cn = (C.c_char_p * max(1, len(self.names)))(*[s.encode() for s in self.names])
        self._h = ref().ref_frame_create(data._h, len(self.names), cn)

    def __del__(self):
        if getattr(self, "_h", None):
            ref().ref_frame_destroy(self._h)
            self._h = None

    def run(self, iterations, dt, voxel_size, frames=1) -> float:
        ms = C.c_float()
        with _Quiet():
            rc = ref().ref_frame_run(self._h, self.grid._h, iterations, dt, voxel_size, frames, C.byref(ms))
        _rcheck(rc)
        return ms.value

    def vorticity(self, dt, voxel_size, scale, factor_scale):
        """the reference's vorticityConfinement kernel, out of place, on the frame's input velocity -> (N, 3)"""
        out = np.empty((self.data.n, 3), np.float32)
        with _Quiet():
            rc = ref().ref_frame_vorticity(self._h, self.grid._h, dt, voxel_size, scale, factor_scale, out.ctypes.data_as(c_f32p))
        _rcheck(rc)
        return out

    def collision_stage(self, stage, sdf, dt, voxel_size):
        """one reference kernel of the hasCollision path (see ref_frame_collision_stage); stages 0-2 return (N, 3), stage 3 the
        advected scalars"""
        sdf = np.ascontiguousarray(sdf, np.float32)
        out = np.empty((self.data.n, 3), np.float32)
        with _Quiet():
            rc = ref().ref_frame_collision_stage(self._h, self.grid._h, stage, sdf.ctypes.data_as(c_f32p), dt, voxel_size,
                                                 out.ctypes.data_as(c_f32p))
        _rcheck(rc)
        return self.download()["scalars"] if stage == 3 else out

    def download(self):
        n = self.data.n
        vel = np.empty((n, 3), np.float32)
        p = np.empty(n, np.float32)
        div = np.empty(n, np.float32)
        adv = np.empty((n, 3), np.float32)
        sc = [np.empty(n, np.float32) for _ in self.names]
        _rcheck(ref().ref_frame_download(self._h, vel.ctypes.data_as(c_f32p), p.ctypes.data_as(c_f32p), div.ctypes.data_as(c_f32p),
                                         adv.ctypes.data_as(c_f32p), _ptr_array(sc) if sc else None))
        return dict(vel=vel, p=p, div=div, adv=adv, scalars=sc)
