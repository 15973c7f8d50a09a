"""Worker of tests/test_sharded_gpu.py, launched with torch.distributed.run: every rank runs the sharded frame of a small sparse box
in the modes listed in HNS_TEST_MODES (JSON) and rank 0 compares the owned voxels bit for bit with the single-GPU frame
(hnanosolver_b200.dist.sharded_parity_check). HNS_TEST_SHARE_GPU=1: all ranks use cuda:0 (gloo process group, no NCCL -- the ghost
exchange is the CUDA-IPC peer-memory path, which works between processes on one device), so the whole flag / peer-store protocol of
csrc/dist.cu is exercised on a single-GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hnanosolver_b200 import _lib  # noqa: E402
from hnanosolver_b200 import dist as hdist  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    share = bool(int(os.environ.get("HNS_TEST_SHARE_GPU", "0")))
    dev_index = 0 if share else lr
    torch.cuda.set_device(dev_index)
    _lib.lib().hns_set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if share:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    modes = json.loads(os.environ["HNS_TEST_MODES"])
    failed = []
    for mode in modes:
        env = {k: str(v) for k, v in mode.pop("env", {}).items()}
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ok, report = hdist.sharded_parity_check(rank, world, dev, **mode)
        finally:
            for k, v in saved.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
        if rank == 0:
            print("SHARDED_PARITY", json.dumps(dict(report, env=env)), flush=True)
            if not ok:
                failed.append((mode, env))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_PARITY_SUMMARY", "FAILED" if failed else "OK", len(modes), "modes", flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
