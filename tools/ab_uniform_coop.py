#!/usr/bin/env python
"""Crossover between the two bulk kernels of the uniform sampler (SEB_UNIFORM_COOP is read per launch):
time of the whole `a` chain (all primes) per batch size.   python tools/ab_uniform_coop.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

def run(ctx, stream, n, np_, batch):
    gen = torch.Generator(device="cuda").manual_seed(3)
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, n), dtype=torch.int32, device="cuda")
    d_ctr = torch.zeros(batch, dtype=torch.int32, device="cuda")
    def step():
        d_ctr.zero_()
        for p in range(np_):
            ctx.sample_uniform_device(d_ss, d_ctr, p, batch, d_out.data_ptr() + 4 * p * n, np_ * n)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3): step()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3, int(d_out.view(-1)[::1031].to(torch.int64).sum().item())

for n, np_ in ((1024, 1), (4096, 3), (16384, 6)):
    ctx = seb.Context(n, np_, asym=False, device=0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    for batch in (1, 256, 1024, 2048, 4096, 8192, 16384):
        os.environ["SEB_UNIFORM_COOP"] = "0"; t0, c0 = run(ctx, stream, n, np_, batch)
        os.environ["SEB_UNIFORM_COOP"] = "1"; t1, c1 = run(ctx, stream, n, np_, batch)
        print(json.dumps({"n": n, "nprimes": np_, "batch": batch, "thread_per_ct_ms": round(t0, 3), "warp_per_ct_ms": round(t1, 3),
                          "ratio": round(t0 / t1, 2), "same": c0 == c1}), flush=True)
    ctx.close()
