#!/usr/bin/env python
"""Symmetric full path (seb_encrypt_sym_device) per configuration with the per-prime pipeline (prime p's encrypt kernel
on a second stream under prime p+1's sampler) and the two-lane sampler switched off / on:
  python tools/ab_sym_pipeline.py"""
import importlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

def run(ctx, stream, n, np_, batch, reps=3):
    gen = torch.Generator(device="cuda").manual_seed(5)
    d_vals = torch.rand((batch, n // 2), generator=gen, device="cuda", dtype=torch.float32) * 32 - 16
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_sd = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
    def step(): ctx.encrypt_sym_device(d_vals, n // 2, d_ss, d_sd, batch, d_out, False)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): step()
    e1.record(stream); torch.cuda.synchronize()
    d_dig = torch.empty(batch, dtype=torch.int64, device="cuda")
    ctx.digest_device(d_out, 2 * np_ * n, batch, d_dig); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, int(d_dig.sum().item())

rng = np.random.default_rng(7)
for name, n, np_, batch in (("B-sym", 4096, 3, 65536), ("B-sym/4", 4096, 3, 16384), ("C-sym", 8192, 4, 32768),
                            ("D", 16384, 6, 16384), ("D/2", 16384, 6, 8192), ("D/4", 16384, 6, 4096)):
    ctx = seb.Context(n, np_, asym=False, device=0)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    line = {"config": name, "n": n, "nprimes": np_, "batch": batch}
    sums = set()
    for pipe in (0, 1):
        for pair in (0, 1):
            ctx.set_option("sym_pipeline", pipe); ctx.set_option("uniform_pair", pair)
            ms, dg = run(ctx, stream, n, np_, batch)
            line[f"pipeline={pipe},pair={pair}_ms"] = round(ms, 3); sums.add(dg)
    ctx.set_option("sym_pipeline", -1); ctx.set_option("uniform_pair", -1)
    ms, dg = run(ctx, stream, n, np_, batch); sums.add(dg)
    line["auto_ms"] = round(ms, 3); line["same_digests"] = len(sums) == 1
    print(json.dumps(line), flush=True)
    ctx.close()
