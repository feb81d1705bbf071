#!/usr/bin/env python
"""NTT-only kernel (seb_ntt_device) at one degree: ms per launch and fraction of the HBM peak given on the command line.
  python tools/ntt_rate.py <n> <primes> <log2 batch> [peak GB/s]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
seb = importlib.import_module("seal-embedded_b200")
n, np_, lb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
peak = float(sys.argv[4]) if len(sys.argv) > 4 else 6535.4
batch = 1 << lb
ctx = seb.Context(n, np_, asym=False, device=0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
gen = torch.Generator(device="cuda").manual_seed(1)
q = min(ctx.primes)
d = torch.randint(0, int(q), (batch, np_, n), generator=gen, device="cuda", dtype=torch.int32)
for _ in range(3): ctx.ntt_device(d, batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record(stream)
for _ in range(reps): ctx.ntt_device(d, batch)
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
gbs = 8.0 * n * np_ * batch / (ms * 1e-3) / 1e9
print(f"n={n} primes={np_} batch=2^{lb}: {ms:.3f} ms  {gbs:.0f} GB/s  {100 * gbs / peak:.1f} % of {peak:.0f}")
