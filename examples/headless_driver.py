"""Headless driver: what the Houdini SOP does around the hot path, without Houdini (reference src/SOP/HNanoSolver/SOP_HNanoSolver.cpp:
build the sidecar, CreateIndexGrid once per topology, then one Compute_Sim per frame), on a 128^3-bounded smoke plume.

    python examples/headless_driver.py [frames=8] [out_dir]

Two ways to run a frame are shown side by side and must agree bit for bit:
  * the drop-in launchers on HOST buffers (CreateIndexGrid + Compute_Sim), every frame crossing PCIe like the reference does, and
  * the resident state (Simulation.step), which keeps the fields in HBM between frames.
With an output directory the last frame is written as a NanoVDB file + sidecar blocks (hnanosolver_b200.io) and read back.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import hnanosolver_b200 as H
from hnanosolver_b200 import io as hio
from hnanosolver_b200 import synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
out_dir = sys.argv[2] if len(sys.argv) > 2 else None
iterations = 40
w = synth.smoke_plume(128, 2)
fields = dict(density=w.scalars[0], **synth.combustion_fields(w))
# a sphere collider in the plume's way (third SOP input), 0.05 per voxel so the 0.1-wide boundary band is two voxels thick
c = w.coords.astype(np.float32)
fields["collision_sdf"] = (0.05 * (np.sqrt(((c - np.array([64.0, 72.0, 64.0], np.float32)) ** 2).sum(1)) - 10.0)).astype(np.float32)
params = H.CombustionParams(0.5, 2.0, 1.5, 0.1, 1.0, 0.5)  # expansion, temperature gain, buoyancy, ambient, vorticity 1, factor_scale 0.5 (SOP defaults)
names = list(fields)

# ---- drop-in launchers on host buffers ----
data = H.GridIndexedData()
data.setAllocationType(H.AllocationType.CudaPinned)
data.allocateCoords(w.num_voxels)
data.pCoords()[:] = w.coords
data.addValueBlock(H.VEC3F, "vel")
data.pValues(H.VEC3F, "vel")[:] = w.velocity
for k, v in fields.items():
    data.addValueBlock(H.FLOAT, k)
    data.pValues(H.FLOAT, k)[:] = v
grid = H.CreateIndexGrid(data, w.voxel_size)

# ---- resident state ----
sim = H.Simulation(grid, len(fields))
sim.upload(w.velocity, list(fields.values()))
sim.set_combustion(True, names.index("fuel"), names.index("waste"), names.index("temperature"), names.index("flame"), params)
sim.set_collision(names.index("collision_sdf"))

t_host = t_dev = 0.0
for f in range(frames):
    sdf = fields["collision_sdf"]
    data.pValues(H.FLOAT, "collision_sdf")[:] = sdf            # the launcher hands the SDF block back zeroed, like the reference
    t0 = time.perf_counter()
    H.Compute_Sim(data, grid, iterations, w.dt, w.voxel_size, params, True)
    t1 = time.perf_counter()
    sim.step(iterations, w.dt)
    sim.sync()
    t2 = time.perf_counter()
    t_host += t1 - t0
    t_dev += t2 - t1
    vel = data.pValues(H.VEC3F, "vel")
    same = np.array_equal(sim.velocity(), vel) and all(np.array_equal(sim.scalar(i), data.pValues(H.FLOAT, k)) for i, k in enumerate(names) if k != "collision_sdf")
    print(f"frame {f + 1}: max|u| {np.abs(vel).max():8.4f}  mean density {data.pValues(H.FLOAT, 'density').mean():.6f}  "
          f"max temperature {data.pValues(H.FLOAT, 'temperature').max():.4f}  host-buffer path == resident path: {same}", flush=True)
    assert same
print(f"{w.num_leaves} leaves, {w.num_voxels} voxels, {iterations} pressure iterations: {t_host / frames * 1e3:.2f} ms per frame through the "
      f"launchers on host buffers, {t_dev / frames * 1e3:.2f} ms per frame on resident state")

if out_dir:
    hio.save_cache(out_dir, grid.nanovdb_buffer(), data)
    origins, h, back = hio.load_cache(out_dir)
    g2 = H.create_index_grid_from_origins(origins, h)
    assert np.array_equal(g2.nanovdb_buffer()[672:], grid.nanovdb_buffer()[672:]) and np.array_equal(back.pValues(H.VEC3F, "vel"), vel)
    print(f"wrote and re-read {out_dir}/grid.nvdb ({os.path.getsize(os.path.join(out_dir, 'grid.nvdb'))} bytes) + {back.numValueBlocks()} sidecar blocks")
