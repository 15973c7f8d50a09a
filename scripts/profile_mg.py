"""Minimal multigrid driver for ncu / timing: python scripts/profile_mg.py [workload] [cycles] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hnanosolver_b200 as H
from hnanosolver_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = synth.WORKLOADS[name](with_coords=False)
grid = H.create_index_grid_from_origins(w.origins, w.voxel_size)
sim = H.Simulation(grid, 0)
sim.upload(w.velocity)
sim.advect_velocity(w.dt)
sim.divergence(True)
mg = H.Multigrid(grid)
sim.pressure_solve_mg(mg, cycles, 0.0, 2, 2, 1.15)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sim.pressure_solve_mg(mg, cycles, 0.0, 2, 2, 1.15)
e1.record()
torch.cuda.synchronize()
print(f"{w.name}: {cycles} V(2,2) cycles: {e0.elapsed_time(e1) / reps:.3f} ms per solve, relative residual {sim.relative_residual():.3e}, "
      f"graph={'off' if os.environ.get('HNS_MG_GRAPH') == '0' else 'on'}")
