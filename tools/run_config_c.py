"""Configuration C's per-GPU shard (n = 8192, 4 primes, asymmetric, 32768 items), two calls: the workload of the ncu
captures of the asymmetric path at n = 8192.  argv[1..3] = n, primes, batch of another asymmetric configuration."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
seb = importlib.import_module("seal-embedded_b200")
n, np_, batch = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8192, 4, 32768)
ctx = seb.Context(n, np_, asym=True, device=0)
rng = np.random.default_rng(1)
t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
ctx.gen_public_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
gen = torch.Generator(device="cuda").manual_seed(3)
d_vals = torch.rand((batch, n // 2), generator=gen, device="cuda") * 32 - 16
d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.encrypt_asym_device(d_vals, n // 2, d_seeds, batch, d_out)
torch.cuda.synchronize()
