"""First-contact GPU check: product vs oracle vs the unmodified reference on small inputs, then a timing on a big one."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hnanosolver_b200 import synth, launchers as H
from oracle import oracle as O


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def run_case(w, iters):
    print(f"== {w.name}: L={w.num_leaves} N={w.num_voxels} I={iters}", flush=True)
    g = H.create_index_grid_from_origins(w.origins, w.voxel_size)
    ix = O.OracleIndex(w.coords)
    # topology
    bo = ix.nanovdb_buffer(w.voxel_size); bp = g.nanovdb_buffer()
    print("  nanovdb buffer: product == oracle emitter:", bool(np.array_equal(bo, bp)), bo.size)
    q = np.concatenate([w.coords[::7], w.coords[::13] + np.array([0, 0, 9], np.int32), w.coords[::11] - np.array([9, 0, 0], np.int32)])
    print("  getValue product vs oracle mismatches:", int((g.get_values(q) != ix.get_values(q)).sum()))
    have_ref = O.ref_gpu_available()
    if have_ref:
        rd = O.RefData(w.coords); rd.add_vec3("vel", w.velocity)
        for n, a in zip(w.scalar_names, w.scalars): rd.add_float(n, a)
        rg = O.RefGrid(rd, w.voxel_size)
        br = rg.buffer()
        print("  ref buffer size", br.size, "getValue ref vs oracle mismatches:", int((rg.get_values(q) != ix.get_values(q)).sum()))
        diff = np.nonzero(br != bp)[0]
        print("  bytes differing product vs reference voxelsToGrid:", diff.size, diff[:20])
    sim = H.Simulation(g, len(w.scalars)); sim.upload(w.velocity, w.scalars)
    t = time.time(); orc = ix.frame(w.velocity, w.scalars, iters, w.dt, w.voxel_size); t_or = time.time() - t
    for flags, nm in ((0, "alternating"), (2, "forward")):
        sim.upload(w.velocity, w.scalars)
        sim.step(iters, w.dt, flags); sim.sync()
        out = dict(vel=sim.velocity(), div=sim.aux(0), p=sim.aux(1), adv=sim.aux(2), scalars=[sim.scalar(i) for i in range(len(w.scalars))])
        print(f"  [{nm}] vs oracle: adv {rel(out['adv'], orc['adv']):.2e} div {rel(out['div'], orc['div']):.2e} p {rel(out['p'], orc['p']):.2e} "
              f"vel {rel(out['vel'], orc['vel']):.2e} scal {[f'{rel(a, b):.2e}' for a, b in zip(out['scalars'], orc['scalars'])]}")
        if flags == 0: fused = out
        else: print("  alternating == forward bitwise:", all(np.array_equal(fused[k], out[k]) for k in ("vel", "div", "p", "adv")))
    if have_ref:
        rf = O.RefFrame(rd, rg, w.scalar_names); rf.run(iters, w.dt, w.voxel_size, 1); r = rf.download()
        print(f"  product vs REFERENCE kernels: adv {rel(fused['adv'], r['adv']):.2e} div {rel(fused['div'], r['div']):.2e} p {rel(fused['p'], r['p']):.2e} "
              f"vel {rel(fused['vel'], r['vel']):.2e} scal {[f'{rel(a, b):.2e}' for a, b in zip(fused['scalars'], r['scalars'])]}")
        print(f"  oracle  vs REFERENCE kernels: adv {rel(orc['adv'], r['adv']):.2e} div {rel(orc['div'], r['div']):.2e} p {rel(orc['p'], r['p']):.2e} "
              f"vel {rel(orc['vel'], r['vel']):.2e} scal {[f'{rel(a, b):.2e}' for a, b in zip(orc['scalars'], r['scalars'])]}")
        print("  bitwise product==ref:", {k: bool(np.array_equal(fused[k], r[k])) for k in ("adv", "div", "p", "vel")},
              [bool(np.array_equal(a, b)) for a, b in zip(fused['scalars'], r['scalars'])])
    print(f"  oracle frame on {O.num_threads()} threads: {t_or*1e3:.1f} ms")


def timing(name, iters=40, frames=5):
    t = time.time(); w = synth.WORKLOADS[name](with_coords=False); print(f"== timing {w.name}: L={w.num_leaves} N={w.num_voxels} gen {time.time()-t:.1f}s", flush=True)
    g = H.create_index_grid_from_origins(w.origins, w.voxel_size)
    sim = H.Simulation(g, len(w.scalars)); sim.upload(w.velocity, w.scalars)
    for flags, nm in ((0, "alternating"), (2, "forward")):
        sim.time_frames(2, iters, w.dt, flags)
        tot, pr = sim.time_frames(frames, iters, w.dt, flags)
        ms = tot / frames
        B = (80 + 16 * iters + 8 * len(w.scalars)) * w.num_voxels
        print(f"  [{nm}] {ms:.3f} ms/frame (pressure {pr/frames:.3f} ms)  {w.num_voxels/ms*1e3/1e9:.2f} Gvox/s  algorithmic {B/ms*1e3/1e12:.2f} TB/s", flush=True)
    if O.ref_gpu_available() and w.num_voxels <= 5e7:
        coords = synth.dense_coords(w.origins)
        rd = O.RefData(coords); rd.add_vec3("vel", w.velocity)
        for n, a in zip(w.scalar_names, w.scalars): rd.add_float(n, a)
        rg = O.RefGrid(rd, w.voxel_size); rf = O.RefFrame(rd, rg, w.scalar_names)
        rf.run(iters, w.dt, w.voxel_size, 1)
        ms = rf.run(iters, w.dt, w.voxel_size, 3) / 3
        print(f"  [reference kernels, resident] {ms:.3f} ms/frame", flush=True)


if __name__ == "__main__":
    import torch
    print(torch.cuda.get_device_name(0))
    run_case(synth.random_leaves(40, 5, 0, offset=(-24, -4096 - 16, 4080)), 8)
    run_case(synth.smoke_sphere(64, 1), 20)
    run_case(synth.smoke_plume(128, 2), 40)
    for nm in sys.argv[1:]:
        timing(nm)
