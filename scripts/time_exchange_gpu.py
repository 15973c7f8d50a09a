"""torchrun: latency of one ghost exchange (pack + grouped ncclSend/ncclRecv + unpack) of a pressure half-field, back to back."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hnanosolver_b200 import dist as hdist, synth, _lib

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); _lib.lib().hns_set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
go = hdist.global_sparse_origins(hdist.WEAK_BOX[world])
plan = hdist.make_plan(go, world, rank)
sh = hdist.ShardedSimulation(plan, np.ascontiguousarray(go[plan.local_ids]), 0.1, 2, torch.device("cuda", lr))
L = _lib.lib()
for fields, name in (([6], "p_red half (1 KB/leaf)"), ([0, 1, 2], "velocity (6 KB/leaf)")):
    arr = (C.c_int * len(fields))(*fields)
    for _ in range(20):
        _lib.check(L.hns_dist_exchange(sh._dist, sh.sim._h, len(fields), arr, None))
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 200
    for _ in range(n):
        _lib.check(L.hns_dist_exchange(sh._dist, sh.sim._h, len(fields), arr, None))
    e1.record(); torch.cuda.synchronize()
    print(f"[rank {rank}] {name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us per exchange; send leaves { {p: len(v) for p, v in plan.send.items()} }", flush=True)
dist.barrier(); sh.close(); dist.destroy_process_group()
