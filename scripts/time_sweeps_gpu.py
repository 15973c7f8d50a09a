"""torchrun: where does a sharded pressure half-sweep spend its time? Exchange-free sweeps over the different work lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hnanosolver_b200 import dist as hdist, _lib

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); _lib.lib().hns_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
go = hdist.global_sparse_origins(hdist.WEAK_BOX[world])
plan = hdist.make_plan(go, world, rank)
sh = hdist.ShardedSimulation(plan, np.ascontiguousarray(go[plan.local_ids]), 0.1, 2, torch.device("cuda", lr))
names = ["all local leaves, no list", "owned list", "interior list", "boundary list", "interior || boundary (2 streams)"]
out = {}
for mode, name in enumerate(names):
    sh.time_sweeps(mode, 20)
    dist.barrier()
    out[name] = round(sh.time_sweeps(mode, 200), 1)
nb = int(plan.owned_local.sum())
print(f"[rank {rank}] local leaves {len(plan.local_ids)} owned {nb} us/half-sweep {out}", flush=True)
dist.barrier(); sh.close(); dist.destroy_process_group()
