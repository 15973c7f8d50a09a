"""How fast is a pressure half-sweep when its working set fits the L2? Sparse boxes of growing size (30 % fill), 40 iterations each:
ns per leaf per half-sweep against the bytes the sweep touches (3 KB per leaf: this colour's p and div, the other colour's p)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hnanosolver_b200 as H
from hnanosolver_b200 import dist as hdist, launchers as HL

st = torch.cuda.current_stream().cuda_stream
for box in ((128, 128, 128), (192, 192, 192), (256, 256, 192), (256, 256, 256), (320, 320, 256), (384, 320, 320), (384, 384, 384), (448, 448, 384), (512, 512, 512)):
    go = hdist.global_sparse_origins(box)
    g = H.create_index_grid_from_origins(go, 0.1)
    sim = H.Simulation(g, 0)
    omega = HL.omega_compute(0.1)
    sim.pressure_solve(40, omega, 0, st); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.pressure_solve(40, omega, 0, st); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    L = go.shape[0]
    us = best * 1e3 / 80
    print(f"box {box}: {L:6d} leaves, sweep touches {3 * L / 1024:7.1f} MB (fields total {4 * L / 1024:7.1f} MB): {us:6.2f} us per half-sweep = "
          f"{us * 1e3 / L:5.3f} ns/leaf = {8 * 512 * L / (us * 1e-6) / 1e12:5.2f} TB/s algorithmic", flush=True)
    del sim, g
