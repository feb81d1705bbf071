#!/usr/bin/env python
"""Host-pointer asymmetric path (config B, pinned buffers) vs chunk size (SEB_HOST_CHUNK is read per call)."""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")
n, np_, batch = 4096, 3, 65536
ctx = seb.Context(n, np_, asym=True, device=0)
rng = np.random.default_rng(1)
ctx.set_public_key(*[np.stack([rng.integers(0, q, n, dtype=np.uint32) for q in ctx.primes]) for _ in range(2)])
vals = torch.empty((batch, n // 2), dtype=torch.float32).uniform_(-16, 16).pin_memory()
seeds = torch.randint(0, 256, (batch, 64), dtype=torch.uint8).pin_memory()
out = torch.empty((batch, np_, 2, n), dtype=torch.int32).pin_memory()
for chunk in (sys.argv[1:] or ["682", "1024", "2048", "4096", "8192", "0"]):
    if chunk == "0": os.environ.pop("SEB_HOST_CHUNK", None)
    else: os.environ["SEB_HOST_CHUNK"] = chunk
    fn = lambda: ctx.lib.seb_encrypt_asym_host(ctx.h, vals.data_ptr(), n // 2, seeds.data_ptr(), batch, out.data_ptr())
    fn()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(json.dumps({"chunk": chunk, "ms_best": round(dt * 1e3, 2), "ms_all": [round(t * 1e3, 1) for t in ts],
                      "ct_per_s": round(batch / dt), "d2h_GBps": round(batch * np_ * 2 * n * 4 / dt / 1e9, 1)}), flush=True)
