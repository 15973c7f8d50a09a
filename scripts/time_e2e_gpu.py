"""Where the end-to-end time of one cook goes (one GPU): CreateIndexGrid, Compute_Sim, and the raw PCIe rates of this box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hnanosolver_b200 as H
from hnanosolver_b200 import synth
from bench import PARAMS6

w = synth.WORKLOADS["c4"](with_coords=False)
fields = dict(density=w.scalars[0], **synth.combustion_fields(w))
N = w.num_voxels
data = H.GridIndexedData()
data.setAllocationType(H.AllocationType.CudaPinned)
data.allocateCoords(N)
for a in range(0, w.num_leaves, 8192):
    b = min(w.num_leaves, a + 8192)
    data.pCoords()[a * 512:b * 512] = synth.dense_coords(w.origins[a:b])
data.addValueBlock(H.VEC3F, "vel"); data.pValues(H.VEC3F, "vel")[:] = w.velocity
for nm, a in fields.items():
    data.addValueBlock(H.FLOAT, nm); data.pValues(H.FLOAT, nm)[:] = a
params = H.CombustionParams(*PARAMS6)
# raw PCIe
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory(); d = torch.empty_like(x, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(f"pinned {name}: {3 * x.numel() * 4 / (time.perf_counter() - t) / 1e9:.1f} GB/s")
y = torch.empty_like(x).pin_memory(); d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); print(f"pinned H2D + D2H concurrently: {3 * x.numel() * 4 / (time.perf_counter() - t) / 1e9:.1f} GB/s each way")
del x, y, d, d2
tg, tc = [], []
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = H.CreateIndexGrid(data, w.voxel_size)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    H.Compute_Sim(data, g, 40, w.dt, w.voxel_size, params, False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    g.reset()
    if i >= 2: tg.append((t1 - t0) * 1e3); tc.append((t2 - t1) * 1e3)
print(f"CreateIndexGrid {np.median(tg):.2f} ms   Compute_Sim {np.median(tc):.2f} ms   (H2D {N * 32 / 1e9:.2f} GB, D2H {N * 32 / 1e9:.2f} GB per cook)")
