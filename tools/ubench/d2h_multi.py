"""Aggregate device->pinned-host copy rate when k of the node's GPUs copy at once (torchrun, one rank per GPU):
does the host side sustain eight 56 GB/s streams, and does the aggregate DROP beyond some k (contention)?
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ubench/d2h_multi.py"""
import json, os, time
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = (2 << 30) // 4
d = torch.empty(n, dtype=torch.int32, device="cuda")
h = torch.empty(n, dtype=torch.int32).pin_memory()
h.copy_(d); torch.cuda.synchronize()
res = {}
for k in (1, 2, 3, 4, 5, 6, 8):
    if k > world: continue
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank < k:
        for _ in range(3): h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0 if rank < k else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    res[k] = round(k * 3 * n * 4 / float(dt.item()) / 1e9, 1)
if rank == 0: print(json.dumps({"aggregate_d2h_GBps_by_concurrent_gpus": res}))
dist.destroy_process_group()
