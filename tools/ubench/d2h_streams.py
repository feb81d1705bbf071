import torch, time, json
n = 6 * (1 << 30) // 4
d = torch.empty(n, dtype=torch.int32, device="cuda")
h = torch.empty(n, dtype=torch.int32).pin_memory()
def run(nstreams, pieces):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // pieces
    for i in range(pieces):
        with torch.cuda.stream(streams[i % nstreams]):
            h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return round(n * 4 / dt / 1e9, 2)
res = {}
for ns, pc in ((1, 1), (1, 32), (2, 32), (4, 32), (2, 2), (4, 64)):
    run(ns, pc)
    res[f"{ns} streams x {pc} pieces"] = max(run(ns, pc) for _ in range(3))
print(json.dumps(res))
# H2D concurrent effect
h2 = torch.empty(512 << 18, dtype=torch.int32).pin_memory(); d2 = torch.empty_like(h2, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
with torch.cuda.stream(s2):
    for _ in range(1): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("D2H 6 GiB with concurrent 512 MiB H2D:", round(n * 4 / dt / 1e9, 2), "GB/s")
