"""BASELINE configs[3] shape only (the block-cyclic stage 1 of bench.py, without the rest of the bench):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_only.py [n]
prints the same record bench.py stores under config.north_star.config3_block_cyclic_stage1."""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svdsolver_b200 import capi  # noqa: E402

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


args = types.SimpleNamespace(dist_n=int(sys.argv[1]) if len(sys.argv) > 1 else 65536)
stream = torch.cuda.Stream()
res = bench.dist_stage1_config(args, capi, torch, dist, stream, dev, lr, rank, world, barrier)
if rank == 0:
    print(json.dumps(res, indent=1), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
