"""Configuration D's per-GPU shard (n = 16384, 6 primes, symmetric, 16384 items), two calls: the workload of the ncu
captures of the symmetric path.  argv[1] = 1 forces the two-lane uniform sampler (k_uniform_bulk_pair), -1 leaves the
choice to the library; argv[2..4] = n, primes, batch of another symmetric configuration (B-sym: 4096 3 65536)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
seb = importlib.import_module("seal-embedded_b200")
n, np_, batch = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (16384, 6, 16384)
ctx = seb.Context(n, np_, asym=False, device=0)
if len(sys.argv) > 1 and int(sys.argv[1]) >= 0:
    ctx.set_option("uniform_pair", int(sys.argv[1]))
rng = np.random.default_rng(1)
t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
gen = torch.Generator(device="cuda").manual_seed(3)
d_vals = torch.rand((batch, n // 2), generator=gen, device="cuda") * 32 - 16
d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
d_ss = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.encrypt_sym_device(d_vals, n // 2, d_ss, d_seeds, batch, d_out, False)
torch.cuda.synchronize()
