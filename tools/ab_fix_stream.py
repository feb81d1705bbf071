#!/usr/bin/env python
"""The uniform sampler's full-warp fix-up as a warp per ciphertext (K = 1) against the streamed form with K = 2 / 4 / 8
ciphertexts per warp, on the whole `a` chain (all primes, bulk squeeze + fix-up):   python tools/ab_fix_stream.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")
from tools.ab_uniform_pair_run import run  # noqa: E402

for n, np_ in ((4096, 3), (8192, 4), (16384, 6)):
    ctx = seb.Context(n, np_, asym=False, device=0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    ctx.set_option("uniform_fix_wide", 0)
    for batch in (8192, 16384, 32768, 65536):
        if batch * np_ * n * 4 > (40 << 30): continue
        res = {}
        for k in (0, 2, 4, 8):
            ctx.set_option("uniform_fix_stream", k)
            res[k] = run(ctx, stream, n, np_, batch)
        ctx.set_option("uniform_fix_stream", -1)
        auto = run(ctx, stream, n, np_, batch)
        line = {"n": n, "nprimes": np_, "batch": batch}
        for k, (t, c) in res.items(): line["K=%d_ms" % max(k, 1)] = round(t, 3)
        line["auto_ms"] = round(auto[0], 3)
        line["same"] = len({c for _, c in res.values()} | {auto[1]}) == 1
        print(json.dumps(line), flush=True)
    ctx.close()
