"""Minimal resident-frame driver for ncu: builds a workload, runs `frames` full frames. Usage: profile_frame.py [workload] [frames] [forward]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hnanosolver_b200 as H
from hnanosolver_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = 2 if len(sys.argv) > 3 and sys.argv[3] == "forward" else 0
w = synth.WORKLOADS[name](with_coords=False)
fields = dict(density=w.scalars[0], **synth.combustion_fields(w))
names = list(fields)
g = H.create_index_grid_from_origins(w.origins, w.voxel_size)
sim = H.Simulation(g, len(fields))
sim.upload(w.velocity, list(fields.values()))
sim.set_combustion(True, names.index("fuel"), names.index("waste"), names.index("temperature"), names.index("flame"),
                   H.CombustionParams(0.5, 2.0, 1.5, 0.1, 0.0, 1.0))
for _ in range(frames):
    sim.step(40, w.dt, flags)
sim.sync()
print("done", w.num_voxels)
