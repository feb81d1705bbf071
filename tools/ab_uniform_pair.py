#!/usr/bin/env python
"""The uniform sampler's three bulk kernels per batch size — one sponge per thread (k_uniform_bulk), two lanes per
sponge with bit-interleaved halves (k_uniform_bulk_pair), 25 lanes per sponge (k_uniform_bulk_coop) — on the whole `a`
chain (all primes, bulk squeeze + fix-up):   python tools/ab_uniform_pair.py [--quick]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")

from tools.ab_uniform_pair_run import run  # noqa: E402

quick = "--quick" in sys.argv
for n, np_ in ((4096, 3), (16384, 6)) if quick else ((1024, 1), (4096, 3), (8192, 4), (16384, 6)):
    ctx = seb.Context(n, np_, asym=False, device=0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    for batch in (256, 512, 1024, 1536, 2048, 4096, 8192, 12288, 16384, 20480, 24576, 28672, 32768, 40960, 49152, 65536):
        if batch * np_ * n * 4 > (40 << 30): continue
        res = {}
        for name, coop, pair in (("thread", 0, 0), ("pair", 0, 1), ("warp25", 1, 0)):
            if name == "warp25" and batch > 8192: continue
            ctx.set_option("uniform_coop", coop); ctx.set_option("uniform_pair", pair); ctx.set_option("uniform_mix", 0)
            res[name] = run(ctx, stream, n, np_, batch)
        ctx.set_option("uniform_coop", -1); ctx.set_option("uniform_pair", -1); ctx.set_option("uniform_mix", -1)
        res["auto"] = run(ctx, stream, n, np_, batch)  # the library's own choice (incl. the mixed squeeze)
        line = {"n": n, "nprimes": np_, "batch": batch}
        for k, (t, c) in res.items(): line[k + "_ms"] = round(t, 3)
        line["thread/pair"] = round(res["thread"][0] / res["pair"][0], 3)
        line["same"] = len({c for _, c in res.values()}) == 1
        print(json.dumps(line), flush=True)
    ctx.close()
