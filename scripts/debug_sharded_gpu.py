"""Stage-by-stage comparison of the sharded frame against a global single-GPU frame computed redundantly on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import hnanosolver_b200 as H
from hnanosolver_b200 import dist as hdist, synth, _lib

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); _lib.lib().hns_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
go = hdist.global_sparse_origins((256, 128, 128), 0.35, 7)
vel, den, tem = synth._swirl_fields(256)
wg = synth._finish("global", go, vel, [den, tem], ["density", "temperature"], 40, 7, with_coords=False)
plan = hdist.make_plan(go, world, rank)
lo = (np.repeat(plan.local_ids, 512) * 512 + np.tile(np.arange(512), plan.n_local))
own = np.repeat(plan.owned_local, 512)
sh = hdist.ShardedSimulation(plan, np.ascontiguousarray(go[plan.local_ids]), wg.voxel_size, 2, torch.device("cuda", lr))
g = H.create_index_grid_from_origins(go, wg.voxel_size)
ref = H.Simulation(g, 2)
ref.upload(wg.velocity, wg.scalars)
# ghost leaves start with garbage so that a missing exchange shows
v0 = wg.velocity[lo].copy(); v0[~own] = 777.0
sh.upload(v0, [s[lo] for s in wg.scalars])
def cmp(name, local, glob, mask=None):
    a, b = local, glob[lo]
    if mask is not None: a, b = a[mask], b[mask]
    bad = (a != b).reshape(a.shape[0], -1).any(1)
    print(f"[rank {rank}] {name:34s} mismatching voxels {int(bad.sum())}/{a.shape[0]}", flush=True)
st = torch.cuda.current_stream().cuda_stream
sh.ex.exchange(hdist.F_VEL); torch.cuda.synchronize()
cmp("velocity after exchange (all)", sh.sim.velocity(), wg.velocity)
sh.sim.advect_velocity(wg.dt, st); ref.advect_velocity(wg.dt); torch.cuda.synchronize()
cmp("advected velocity (owned)", sh.sim.aux(2), ref.aux(2), own)
sh.ex.exchange(hdist.F_ADV); torch.cuda.synchronize()
cmp("advected velocity after exch (all)", sh.sim.aux(2), ref.aux(2))
sh.sim.divergence(True, st); ref.divergence(True); torch.cuda.synchronize()
cmp("divergence (owned)", sh.sim.aux(0), ref.aux(0), own)
om = H.launchers.omega_compute(wg.voxel_size)
sh.sim.pressure_init(st); ref.pressure_init()
for it in range(3):
    for c in (0, 1):
        sh.sim.pressure_half_sweep(c, om, bool(c), st); ref.pressure_half_sweep(c, om, False); torch.cuda.synchronize()
        cmp(f"p after sweep it{it} c{c} (owned)", sh.sim.aux(1), ref.aux(1), own)
        sh.ex.exchange([hdist.F_P_RED if c == 0 else hdist.F_P_BLK]); torch.cuda.synchronize()
        cmp(f"p after exchange it{it} c{c} (all)", sh.sim.aux(1), ref.aux(1))
dist.barrier(); dist.destroy_process_group()
