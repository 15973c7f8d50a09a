"""python scripts/mg_gpu.py [workload] : the reference's fixed-count red-black solve vs multigrid V-cycles on one GPU -- time (CUDA
events), relative Poisson residual, divergence of the projected velocity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hnanosolver_b200 as H
from hnanosolver_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = synth.WORKLOADS[name](with_coords=False)
grid = H.create_index_grid_from_origins(w.origins, w.voxel_size)
sim = H.Simulation(grid, 0)
sim.upload(w.velocity)
sim.advect_velocity(w.dt)
sim.divergence(True)
t = time.time()
mg = H.Multigrid(grid)
torch.cuda.synchronize()
print(f"{w.name}: {w.num_leaves} leaves; hierarchy built in {time.time()-t:.3f}s:", [(mg.level_leaves(k), mg.level_cells(k)) for k in range(mg.num_levels)])
div0 = sim.divergence_sum_squares(True)
omega = H.launchers.omega_compute(w.voxel_size)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(tag, solve):
    sim.divergence(True)
    ms = timed(solve)
    rel = sim.relative_residual()
    sim.subtract_gradient(True)
    d = sim.divergence_sum_squares(False)
    print(f"{tag:34s} {ms:8.3f} ms   rel residual {rel:.3e}   ||div u_new|| / ||div u*|| {np.sqrt(d / div0):.4f}", flush=True)


for I in (40, 100):
    report(f"red-black SOR I={I}", lambda: sim.pressure_solve(I, omega))
for nu in ((1, 1), (2, 2), (2, 1)):
    for om in (1.0, 1.15):
        for cyc in (1, 2, 3):
            report(f"V({nu[0]},{nu[1]}) x{cyc} omega {om}", lambda: sim.pressure_solve_mg(mg, cyc, 0.0, nu[0], nu[1], om))
c, rel = sim.pressure_solve_mg(mg, 30, 1e-4, 2, 2, 1.15)
print("to 1e-4:", c, "cycles, rel", rel, f"{timed(lambda: sim.pressure_solve_mg(mg, 30, 1e-4, 2, 2, 1.15), 3):.3f} ms")
