import importlib, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
seb = importlib.import_module("seal-embedded_b200")
for n, np_ in ((4096, 3), (8192, 4), (16384, 6)):
    for spec in ("0", "auto"):
        if spec == "auto": os.environ.pop("SEB_UNIFORM_SPEC", None)
        else: os.environ["SEB_UNIFORM_SPEC"] = spec
        ctx = seb.Context(n, np_, asym=False, device=0)
        rng = np.random.default_rng(1)
        t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
        ctx.set_secret_key(((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8))
        res = {}
        for batch in (1, 2, 4, 8, 16, 32):
            d_vals = torch.rand((batch, n // 2), device="cuda") * 32 - 16
            d_seeds = torch.randint(0, 256, (batch, 64), device="cuda", dtype=torch.uint8)
            d_ss = torch.randint(0, 256, (batch, 64), device="cuda", dtype=torch.uint8)
            d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
            f = lambda: (ctx.encrypt_sym_device(d_vals, n // 2, d_ss, d_seeds, batch, d_out, False), ctx.encode_failures())
            f(); t0 = time.perf_counter()
            for _ in range(5): f()
            res[batch] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
        print(json.dumps({"n": n, "nprimes": np_, "spec": spec, "ms_per_call_by_batch": res}), flush=True)
        ctx.close()
