#!/usr/bin/env python
"""The symmetric device path with and without the SM partition (option "sym_partition"), per shape:
  python tools/ab_sym_partition.py"""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

for n, np_, batches in ((16384, 6, (2048, 4096, 8192, 12288, 16384)), (16384, 2, (16384,)), (8192, 4, (4096, 8192, 16384)),
                        (8192, 1, (16384,))):
    ctx = seb.Context(n, np_, asym=False, device=0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
    gen = torch.Generator(device="cuda").manual_seed(3)
    for batch in batches:
        d_vals = torch.rand((batch, n // 2), generator=gen, device="cuda") * 32 - 16
        d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
        d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
        d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
        line = {"n": n, "nprimes": np_, "batch": batch}
        sums = []
        for name, mode in (("serial", 0), ("auto", -1), ("forced", 1)):
            ctx.set_option("sym_partition", mode)
            ctx.encrypt_sym_device(d_vals, n // 2, d_ss, d_seeds, batch, d_out, False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                ctx.encrypt_sym_device(d_vals, n // 2, d_ss, d_seeds, batch, d_out, False)
            e1.record(stream); torch.cuda.synchronize()
            line[name + "_ms"] = round(e0.elapsed_time(e1) / 3, 3)
            sums.append(int(d_out.view(-1)[::1031].to(torch.int64).sum().item()))
        line["same"] = len(set(sums)) == 1
        print(json.dumps(line), flush=True)
        del d_vals, d_out
    ctx.close()
