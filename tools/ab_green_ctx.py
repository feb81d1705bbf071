#!/usr/bin/env python
"""Feasibility of SM partitioning (CUDA green contexts) for the symmetric path at configuration D's shard: the uniform
sampler's sequential chain keeps one warp per SM sub-partition busy on ~128 of the 148 SMs for ~21 ms; can the encode and the
CBD sampler of (part of) the batch run on the other SMs meanwhile?   python tools/ab_green_ctx.py [sms_for_side_work]"""
import importlib, os, sys
import numpy as np, torch
from cuda.bindings import driver as cu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")


def ck(r):
    err = r[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return r[1] if len(r) == 2 else r[1:]


side = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n, np_, batch = 16384, 6, 16384
torch.cuda.init(); torch.zeros(1, device="cuda")
dev = ck(cu.cuDeviceGet(0))
res = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print("SMs:", res.sm.smCount)
groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, res, 0, side))
print("side group SMs:", groups[0].sm.smCount, "remaining:", rem.sm.smCount)
streams = []
for r in (rem, groups[0]):
    desc = ck(cu.cuDevResourceGenerateDesc([r], 1))
    g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
    s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
    streams.append(torch.cuda.ExternalStream(int(s)))
sa, sb = streams  # sa: the big partition (sampler chain), sb: the side partition
main = torch.cuda.Stream()

ctxs = [seb.Context(n, np_, asym=False, device=0) for _ in range(3)]  # main / A / B
for c, s in zip(ctxs, (main, sa, sb)):
    c.set_stream(s.cuda_stream)
gen = torch.Generator(device="cuda").manual_seed(3)
d_vals = torch.rand((batch, n // 2), generator=gen, device="cuda") * 32 - 16
d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
d_pt = torch.empty((batch, n), dtype=torch.int64, device="cuda")
d_e = torch.empty((batch, n), dtype=torch.int8, device="cuda")
d_a = torch.empty((batch, np_, n), dtype=torch.int32, device="cuda")
d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")


def enc_cbd(c, lo, hi):
    if hi <= lo:
        return
    c.encode_device(d_vals[lo:hi], n // 2, hi - lo, d_pt[lo:hi])
    c.sample_cbd_device(d_seeds[lo:hi], None, 1, hi - lo, d_e[lo:hi])


def chain(c):
    d_ctr.zero_()
    for p in range(np_):
        c.sample_uniform_device(d_ss, d_ctr, p, batch, d_a.data_ptr() + 4 * p * n, np_ * n)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _ in range(reps):
        fn()
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def serial():
    with torch.cuda.stream(main):
        enc_cbd(ctxs[0], 0, batch)
        chain(ctxs[0])


def overlapped(b1):
    def run():
        with torch.cuda.stream(main):
            enc_cbd(ctxs[0], b1, batch)  # the part the whole machine does first
            ev = torch.cuda.Event(); ev.record(main)
        sa.wait_event(ev); sb.wait_event(ev)
        with torch.cuda.stream(sa):
            chain(ctxs[1])
            ea = torch.cuda.Event(); ea.record(sa)
        with torch.cuda.stream(sb):
            enc_cbd(ctxs[2], 0, b1)
            eb = torch.cuda.Event(); eb.record(sb)
        main.wait_event(ea); main.wait_event(eb)
    return run


def chain_only_on(c, s):
    def run():
        ev = torch.cuda.Event(); ev.record(main); s.wait_event(ev)
        with torch.cuda.stream(s):
            chain(c)
            e = torch.cuda.Event(); e.record(s)
        main.wait_event(e)
    return run


print(f"serial (one stream, whole device): {timed(serial):.3f} ms")
print(f"sampler chain alone, whole device: {timed(chain_only_on(ctxs[0], main)):.3f} ms")
print(f"sampler chain alone, big partition: {timed(chain_only_on(ctxs[1], sa)):.3f} ms")
for frac in (0.25, 0.30, 0.33, 0.36, 0.40, 0.5):
    b1 = int(batch * frac) // 8 * 8
    print(f"overlapped, {b1} of {batch} items' encode+CBD on the side partition: {timed(overlapped(b1)):.3f} ms")
