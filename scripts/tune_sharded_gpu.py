"""torchrun: A/B variants of the sharded frame on the weak-scaling workload, built once.

    VARIANTS="HNS_SIGNAL_IN_KERNEL=0;HNS_SIGNAL_IN_KERNEL=1" torchrun ... scripts/tune_sharded_gpu.py

Each variant is a ';'-separated list of comma-separated KEY=VALUE environment settings applied before the ShardedSimulation is made.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hnanosolver_b200 import dist as hdist, launchers as H, _lib
from bench import PARAMS6

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); _lib.lib().hns_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
w, names, fields, kind = hdist.build_sharded_workload("c4", rank, world)
plan = w.meta["plan"]
steps = int(os.environ.get("STEPS", "10"))
for variant in os.environ.get("VARIANTS", "").split(";"):
    for kv in filter(None, variant.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    sh = hdist.ShardedSimulation(plan, w.origins, w.voxel_size, len(fields), dev)
    sh.set_combustion(names, H.CombustionParams(*PARAMS6))
    sh.upload(w.velocity, fields)
    for _ in range(3):
        sh.frame(40, w.dt)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sh.frame(40, w.dt)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ph = sh.frame_timed(40, w.dt)
    sh.check_errors()
    sweeps = {m: round(sh.time_sweeps(m, 100), 1) for m in (0, 2, 3, 4)}
    step = ph.pop("step20_us(B_end,push_end,wait_end,unpack_end,I_start,I_end)", None)
    if rank in (0, 1):
        print(f"[{variant}] rank {rank} ghost {plan.n_local - plan.n_owned}: frame {ms.item():.3f} ms (max over ranks); pressure {ph['pressure']:.3f} "
              f"div+comb {ph['div+comb']:.3f} exch {ph['exch_vel'] + ph['exch_adv'] + ph['exch_final']:.3f} adv_s {ph['advect_scalars']:.3f} step20 {step} exchange-free sweeps us {sweeps}", flush=True)
    dist.barrier(); sh.close()
dist.destroy_process_group()
