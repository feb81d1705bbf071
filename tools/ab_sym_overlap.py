#!/usr/bin/env python
"""A/B of the symmetric path's two-stream overlap (encode + CBD beside the uniform sampler) per batch size.
  python tools/ab_sym_overlap.py   (prints one JSON line per case; SEB_SYM_OVERLAP is read at context creation)"""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

def run(n, np_, batch, overlap, seedct=False, reps=3):
    os.environ["SEB_SYM_OVERLAP"] = str(overlap)
    ctx = seb.Context(n, np_, asym=False, device=0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
    gen = torch.Generator(device="cuda").manual_seed(3)
    vlen = n // 2
    d_vals = torch.rand((batch, vlen), generator=gen, device="cuda") * 32 - 16
    d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, 1 if seedct else 2, n), dtype=torch.int32, device="cuda")
    def step():
        if seedct: ctx.encrypt_sym_seedct_device(d_vals, vlen, d_ss, d_seeds, batch, d_out)
        else: ctx.encrypt_sym_device(d_vals, vlen, d_ss, d_seeds, batch, d_out, False)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): step()
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    chk = int(d_out.view(-1)[::4099].to(torch.int64).sum().item())
    ctx.close(); del d_out; torch.cuda.empty_cache()
    return ms, chk

for n, np_, batch in ((1024, 1, 65536), (4096, 3, 65536), (4096, 3, 16384), (4096, 3, 4096), (8192, 4, 16384),
                      (16384, 6, 16384), (16384, 6, 4096), (16384, 6, 32768)):
    a, ca = run(n, np_, batch, 0)
    b, cb = run(n, np_, batch, 1)
    print(json.dumps({"n": n, "nprimes": np_, "batch": batch, "serial_ms": round(a, 3), "overlap_ms": round(b, 3),
                      "speedup": round(a / b, 3), "same_output": ca == cb}), flush=True)
