"""torchrun --nproc-per-node N scripts/check_sharded_gpu.py : the sharded GPU frame (NCCL ghost exchange) against the single-GPU
frame of the same global domain, bit for bit, on a 128^3-bounded sparse box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import hnanosolver_b200 as H
from hnanosolver_b200 import dist as hdist, synth, _lib

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
_lib.lib().hns_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
go = hdist.global_sparse_origins((256, 128, 128), 0.35, 7)
vel, den, tem = synth._swirl_fields(256)
wg = synth._finish("global", go, vel, [den, tem], ["density", "temperature"], 40, 7, with_coords=False)
plan = hdist.make_plan(go, world, rank)
lo = (np.repeat(plan.local_ids, 512) * 512 + np.tile(np.arange(512), plan.n_local))
comb = synth.combustion_fields(wg)
names = ["density", "fuel", "waste", "temperature", "flame"]
gfields = [wg.scalars[0]] + [comb[k] for k in names[1:]]
COLL = int(os.environ.get("COLLISION", "0"))  # COLLISION=1: a sphere collider SDF as the last scalar, hasCollision path on
if COLL:
    gc = (np.repeat(go, 512, axis=0) + np.stack(np.unravel_index(np.arange(512), (8, 8, 8)), 1)[np.tile(np.arange(512), go.shape[0])]).astype(np.float32)
    sdf = (0.05 * (np.sqrt(((gc - np.array([128.0, 64.0, 64.0], np.float32)) ** 2).sum(1)) - 20.0)).astype(np.float32)
    sdf[0] = 0.0
    names = names + ["collision_sdf"]
    gfields = gfields + [sdf]
NS = len(gfields)
sh = hdist.ShardedSimulation(plan, np.ascontiguousarray(go[plan.local_ids]), wg.voxel_size, NS, torch.device("cuda", lr),
                             native=not int(os.environ.get("PY_EXCHANGE", "0")))
if COLL: sh.sim.set_collision(NS - 1)
VS, VF = (float(x) for x in os.environ.get("VORT", "0,1").split(","))  # vorticityScale, factorScale (VORT="0.8,2" turns the pass on)
P = H.CombustionParams(0.5, 2.0, 1.5, 0.1, VS, VF)
if not int(os.environ.get("NO_COMB", "0")): sh.set_combustion(names, P)
sh.upload(wg.velocity[lo], [f[lo] for f in gfields])
I = 12
NF = int(os.environ.get("NFRAMES", "2"))
COOK = int(os.environ.get("COOK", "0"))  # COOK=1: drive the frames through the host-buffer entry point (hns_dist_cook), in place
m = np.repeat(plan.owned_local, 512)
if COOK:
    lv = np.ascontiguousarray(wg.velocity[lo]); lf = [np.ascontiguousarray(f[lo]) for f in gfields]
    for _ in range(NF):
        sh.cook(lv, lf, I, wg.dt)
    mine = [lv[m]] + [a[m] for a in lf] + [sh.sim.aux(1)[m]]
else:
    for _ in range(NF):
        sh.frame(I, wg.dt)
    torch.cuda.synchronize()
    mine = [sh.sim.velocity()[m]] + [sh.sim.scalar(i)[m] for i in range(NS)] + [sh.sim.aux(1)[m]]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
if rank == 0:
    g = H.create_index_grid_from_origins(go, wg.voxel_size)
    sim = H.Simulation(g, NS)
    sim.upload(wg.velocity, gfields)
    if COLL: sim.set_collision(NS - 1)
    if not int(os.environ.get("NO_COMB", "0")): sim.set_combustion(True, 1, 2, 3, 4, P)
    for _ in range(NF):
        sim.step(I, wg.dt)
    sim.sync()
    ref = [sim.velocity()] + [sim.scalar(i) for i in range(NS)] + [sim.aux(1)]
    ok = True
    for k, nm in enumerate(["velocity"] + names + ["pressure"]):
        got = np.concatenate([gathered[r][k] for r in range(world)])
        same = np.array_equal(got, ref[k])
        ok &= same
        bad = np.nonzero((got != ref[k]).reshape(got.shape[0], -1).any(1))[0]
        print(f"  {nm:12s} sharded({world}) == single GPU bitwise: {same}  mismatching voxels {bad.size}/{got.shape[0]} "
              f"max|diff| {np.abs(got.astype(np.float64) - ref[k]).max():.3e} first leaves {np.unique(bad // 512)[:8]}")
    print("SHARDED PARITY", "OK" if ok else "FAILED", f"leaves={go.shape[0]} exchanges/frame={sh.exchanges // NF} native={sh.native} p2p={getattr(sh, 'p2p', False)} vorticity=({VS},{VF}) collision={COLL} cook={COOK}", flush=True)
sh.check_errors()
dist.barrier()
sh.close()
dist.destroy_process_group()
