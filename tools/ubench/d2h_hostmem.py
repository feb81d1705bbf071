"""Aggregate device->host copy rate of the node for different kinds of host memory: when all N
ranks copy at once the pool's boxes top out near 90 GB/s with ordinary pinned buffers.  Tried here, same copies:
  pinned      cudaHostAlloc(default)                 (what torch's pin_memory() and bench.py's e2e use)
  wc          cudaHostAlloc(cudaHostAllocWriteCombined)
  thp         anonymous mmap, 2 MiB aligned, madvise(MADV_HUGEPAGE), touched, then cudaHostRegister
  hugetlb     mmap(MAP_HUGETLB) + cudaHostRegister    (only if the box has reserved huge pages)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/ubench/d2h_hostmem.py"""
import ctypes, json, mmap, os, time
import torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
libc = ctypes.CDLL(None, use_errno=True)
libc.mmap.restype = ctypes.c_void_p
libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_long]
libc.munmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]

BYTES = 2 << 30
d = torch.empty(BYTES // 4, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def alloc(kind):
    p = ctypes.c_void_p()
    if kind in ("pinned", "wc"):
        rc = rt.cudaHostAlloc(ctypes.byref(p), BYTES, 4 if kind == "wc" else 0)
        return (p.value, lambda: rt.cudaFreeHost(p)) if rc == 0 else (None, None)
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | (0x40000 if kind == "hugetlb" else 0)  # MAP_HUGETLB
    size = BYTES + (2 << 20)
    base = libc.mmap(None, size, mmap.PROT_READ | mmap.PROT_WRITE, flags, -1, 0)
    if base in (None, ctypes.c_void_p(-1).value):
        return None, None
    addr = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
    if kind == "thp":
        libc.madvise(addr, BYTES, 14)  # MADV_HUGEPAGE
    ctypes.memset(addr, 0, BYTES)  # touch
    if rt.cudaHostRegister(addr, BYTES, 0) != 0:
        libc.munmap(base, size)
        return None, None
    def free():
        rt.cudaHostUnregister(addr)
        libc.munmap(base, size)
    return addr, free


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


res = {}
for kind in ("pinned", "wc", "thp", "hugetlb"):
    h, free = alloc(kind)
    ok = torch.tensor([1 if h else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if not int(ok.item()):
        res[kind] = None
        if h:
            free()
        continue
    rt.cudaMemcpyAsync(h, d.data_ptr(), BYTES, 2, stream)
    for k in sorted({1, world}):
        barrier()
        t0 = time.perf_counter()
        if rank < k:
            for _ in range(3):
                rt.cudaMemcpyAsync(h, d.data_ptr(), BYTES, 2, stream)
            torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0 if rank < k else 0.0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res.setdefault(kind, {})[f"{k}_gpus_GBps"] = round(k * 3 * BYTES / float(dt.item()) / 1e9, 1)
    barrier()
    free()
if rank == 0:
    thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists("/sys/kernel/mm/transparent_hugepage/enabled") else "n/a"
    print(json.dumps({"aggregate_d2h_by_host_memory_kind": res, "world": world, "transparent_hugepage": thp,
                      "cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
