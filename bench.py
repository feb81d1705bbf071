#!/usr/bin/env python
"""bench.py — CKKS encryptions/s at n=4096, 3-prime RNS (BASELINE.json configs[1]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (N>1: under torchrun)
  python bench.py --impl reference [--gpus N] --steps K --warmup W the reference's CPU path on this host

One step = one pass of the hot path (encode -> sample u,e0,e1 -> 3 primes x [3 NTT + pointwise]) over
one batch of 65536 synthetic fp32 messages per GPU, inputs resident in HBM.  Ciphertexts are
independent, so N GPUs shard the batch with no data-path collective ("weak" scaling: the per-GPU
batch is fixed).  Prints ONE JSON line (rank 0).

Beside the headline (config B) the same line carries `other_configs`: BASELINE.json's configs C and D as ONE GLOBAL
BATCH sharded over the N ranks with shard_range ("strong" scaling), B in symmetric mode, and the NTT-only sweep
(config E: n in {1024, 4096, 16384} x primes 1..8 x batch 1..2^20, plus n = 2048 / 8192 at their largest batch) per rank — each timed like the headline (barrier, CUDA events, max over
ranks) — and, at N = 1, the reference's CPU path on bounded samples of A, C and D.
"""
from __future__ import annotations

import argparse
import importlib
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DEG, N_PRIMES, BATCH = 4096, 3, 65536
WORKLOAD = f"n={N_DEG}, {N_PRIMES}-prime RNS, batch={BATCH}/GPU, asymmetric encrypt (BASELINE.json configs[1])"
METRIC = "ckks_encryptions_per_sec"
UNIT = "ciphertexts/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
PRIMES30 = [1053818881, 1054015489, 1054212097, 1055260673, 1056178177, 1056440321, 1058209793, 1060175873,
            1060700161, 1060765697, 1061093377, 1062469633, 1062535169]  # device/lib/parameters.c:129-174


def bench_config(batch: int, world: int) -> dict:
    """The `config` object: identical for both arms (the reference arm runs a bounded SAMPLE of this workload and
    says so in cpu_baseline.sample / sample_items, not here)."""
    in_out_mib = (batch * (N_DEG // 2) * 4 + batch * N_PRIMES * 2 * N_DEG * 4) >> 20
    return {"workload": WORKLOAD.replace(f"batch={BATCH}/GPU", f"batch={batch}/GPU"), "n": N_DEG, "nprimes": N_PRIMES,
            "batch_per_gpu": batch, "parallelism": f"batch-sharded x{world}, no data-path collective",
            "l2": f"inputs+outputs per step ({in_out_mib} MiB) exceed the 126 MB L2"}


# ----------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = the unmodified reference compiled here; else the oracle port)
# ----------------------------------------------------------------------------------------------
def _cpu_worker(conn, idx: int, n: int, np_: int, asym: bool, items: int):
    """One process per core: the reference keeps static state (seal_embedded.c:18-22)."""
    from oracle import oracle as O

    orc = O.Oracle()
    sk = O.make_sk(n)
    pk0 = pk1 = None
    if asym:
        pk0, pk1 = orc.gen_pk(n, np_, sk)
    vals = O.make_values(items, n // 2, seed=1000 + idx)
    seeds = O.make_seeds(items, b"cpu-%d" % idx)
    sseeds = None if asym else O.make_seeds(items, b"cpu-share-%d" % idx)
    ref = None
    if O.have_reference():
        ref = O.ReferenceLib()
        ref.setup(n, np_, asym, sk=sk, pk0=pk0, pk1=pk1, primes=orc.primes(n, np_))
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        t0 = time.perf_counter()
        if ref is not None:
            ref.encrypt_loop(sseeds, seeds, vals)
        elif asym:
            orc.encrypt_asym_batch(n, np_, vals, seeds, pk0, pk1)
        else:
            for b in range(items):
                orc.encrypt_sym(n, np_, vals[b], sseeds[b], seeds[b], sk, ref_quirk=True)
        conn.send(time.perf_counter() - t0)
    if ref is not None:
        ref.close()
    conn.close()


class CpuArm:
    def __init__(self, items_per_worker: int, cores: int | None = None, n: int = N_DEG, np_: int = N_PRIMES,
                 asym: bool = True):
        from oracle import oracle as O

        O.build()
        self.kind = "reference" if O.have_reference() else "port"
        self.cores = cores or (os.cpu_count() or 1)
        self.items, self.n, self.np_, self.asym = items_per_worker, n, np_, asym
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for i in range(self.cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, i, n, np_, asym, items_per_worker), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        for c in self.conns:
            assert c.recv() == "ready"

    def step(self) -> float:
        """All workers encrypt their `items` once; returns the slowest worker's seconds."""
        for c in self.conns:
            c.send("go")
        return max(c.recv() for c in self.conns)

    def close(self):
        for c in self.conns:
            c.send("stop")
        for p in self.procs:
            p.join(timeout=10)

    def describe(self, value: float) -> dict:
        return {"value": value, "unit": UNIT, "cores": self.cores, "kind": self.kind,
                "sample": f"{self.items} items/process x {self.cores} processes per step, se_encrypt_seeded "
                          f"n={self.n} {self.np_} primes {'asym' if self.asym else 'sym'} (one process per core: the "
                          "reference is non-reentrant)"}


def cpu_baseline_for(n: int, np_: int, asym: bool, items: int, reps: int = 2) -> dict:
    arm = CpuArm(items_per_worker=items, n=n, np_=np_, asym=asym)
    arm.step()
    secs = [arm.step() for _ in range(reps)]
    arm.close()
    return arm.describe(arm.cores * arm.items * len(secs) / sum(secs))


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = CpuArm(items_per_worker=96)
    for _ in range(args.warmup):
        arm.step()
    times = [arm.step() for _ in range(args.steps)]
    arm.close()
    total = sum(times)
    value = arm.cores * arm.items * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": bench_config(args.batch, args.gpus),
            "sample_items": arm.cores * arm.items, "host_threads": arm.cores,
            "cpu_baseline": arm.describe(value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pcie_link(gpu_index: int) -> dict | None:
    """Current / maximum PCIe link of the GPU (the e2e number is bound by it; boxes of one pool differ)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        return {"gen": pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width": pynvml.nvmlDeviceGetCurrPcieLinkWidth(h),
                "max_gen": pynvml.nvmlDeviceGetMaxPcieLinkGeneration(h), "max_width": pynvml.nvmlDeviceGetMaxPcieLinkWidth(h)}
    except Exception:  # noqa: BLE001 - informational only
        return None


def hbm_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbps", "hbm_gb_s"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str, batch: int):
    """dram read+write bytes per launch from the committed `ncu --set full` summary (profiles/traffic.json, written by
    tools/summarize_profile.py from a capture at THIS batch size; a capture at another batch is not scaled — a
    small-batch capture under-counts write-backs that were still in L2 when it ended — and reads as null)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p)).get(kernel)
            if t and int(t["batch"]) == int(batch):
                return float(t["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing: barrier + max-over-ranks of device-timed milliseconds."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor(ms if isinstance(ms, (list, tuple)) else [ms], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        v = t.cpu().tolist()
        return v if isinstance(ms, (list, tuple)) else v[0]


def random_sk(n: int, seed: int = 7) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    return ((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8)


def run_full_config(seb, torch, D: Dist, stream, local: int, rank: int, world: int, name: str, n: int, np_: int,
                    asym: bool, global_batch: int, max_sub: int, steps: int, sharded: bool) -> dict:
    """One BASELINE configuration.  sharded: `global_batch` is ONE batch cut over the ranks by shard_range (strong
    scaling; value = global items / max-over-ranks time); otherwise every rank runs `global_batch` items (weak)."""
    sh = seb.shard_range(global_batch, rank, world) if sharded else seb.Shard(rank, world, 0, global_batch)
    count = sh.count
    vlen = n // 2
    ctx = seb.Context(n, np_, asym=asym, device=local)
    ctx.set_stream(stream.cuda_stream)
    sk = random_sk(n)
    if asym:
        ctx.gen_public_key(sk)
    else:
        ctx.set_secret_key(sk)
    sub = min(count, max_sub)
    nsub = (count + sub - 1) // sub if count else 0
    gen = torch.Generator(device="cuda").manual_seed(77 + sh.first)
    d_vals = torch.rand((max(sub, 1), vlen), generator=gen, device="cuda", dtype=torch.float32) * 32 - 16
    d_seeds = torch.randint(0, 256, (max(sub, 1), 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_ss = torch.randint(0, 256, (max(sub, 1), 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((max(sub, 1), np_, 2, n), dtype=torch.int32, device="cuda")

    def step():
        # the rank's shard, in sub-batches that bound the resident output (same inputs re-used per sub-batch: the
        # work per item does not depend on the data, and the outputs of a sub-batch are a complete result)
        done = 0
        while done < count:
            m = min(sub, count - done)
            if asym:
                ctx.encrypt_asym_device(d_vals, vlen, d_seeds, m, d_out)
            else:
                ctx.encrypt_sym_device(d_vals, vlen, d_ss, d_seeds, m, d_out, False)
            done += m

    step()
    assert ctx.encode_failures() == 0
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_ms(e0.elapsed_time(e1)) / steps
    ctx.profile_begin(1)
    if count:
        m = min(sub, count)
        if asym:
            ctx.encrypt_asym_device(d_vals, vlen, d_seeds, m, d_out)
        else:
            ctx.encrypt_sym_device(d_vals, vlen, d_ss, d_seeds, m, d_out, False)
    k = ctx.profile_end()
    kern = {nm: float(v) for nm, v in zip(ctx.PROFILE_SEGMENTS[asym], k[0])} if len(k) else {}
    err = 0.0
    if count:
        m = min(sub, count)
        d_dec = torch.empty((m, vlen), dtype=torch.float32, device="cuda")
        ctx.decrypt_decode_device(d_out, m, 0, vlen, d_dec)
        torch.cuda.synchronize()
        err = float((d_dec - d_vals[:m]).abs().max())
        del d_dec
    err = D.max_ms(err)  # max over ranks (same reduction)
    items = global_batch if sharded else global_batch * world
    res = {"config": name, "n": n, "nprimes": np_, "mode": "asym" if asym else "sym",
           "global_batch": items, "items_this_rank": count, "sub_batch": sub, "sub_batches_per_step": nsub,
           "scaling": "strong (one global batch, shard_range over the ranks)" if sharded else "weak (per-GPU batch fixed)",
           "steps": steps, "ms_per_step": ms, "ciphertexts_per_s": items / (ms * 1e-3) if ms > 0 else None,
           "kernels_ms_one_sub_batch": kern,
           "verify": {"items": min(sub, count), "max_abs_err": err, "tolerance": 0.1, "ok": err < 0.1}}
    ctx.close()
    del d_out, d_vals, d_seeds, d_ss
    torch.cuda.empty_cache()
    return res


def run_ntt_sweep(seb, torch, D: Dist, stream, local: int, peak: float, cap_bytes: int = 6 << 30) -> dict:
    """Config E: k_ntt_forward alone, n x primes 1..8 x batch 2^0..2^20 (points whose polynomials exceed cap_bytes are
    skipped), per rank; the time of a point is the max over ranks, GB/s = 8n bytes x polynomials / that time."""
    points = []
    gen = torch.Generator(device="cuda").manual_seed(99)
    buf = torch.randint(0, 1 << 27, (cap_bytes // 4,), generator=gen, device="cuda", dtype=torch.int32)
    for n in (1024, 4096, 16384):
        for np_ in range(1, 9):
            ctx = seb.Context(n, np_, asym=False, device=local, primes=PRIMES30[:np_])
            ctx.set_stream(stream.cuda_stream)
            for lb in range(0, 21):
                batch = 1 << lb
                if batch * np_ * n * 4 > cap_bytes:
                    break
                reps = 20 if batch * np_ * n < (1 << 24) else 5
                ctx.ntt_device(buf, batch)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    ctx.ntt_device(buf, batch)
                e1.record(stream)
                e1.synchronize()
                points.append((n, np_, lb, e0.elapsed_time(e1) / reps))
            ctx.close()
    # the two degrees between them, at their largest batch under the cap (the plan changes at n = 8192)
    for n in (2048, 8192):
        for np_ in (1, 4):
            ctx = seb.Context(n, np_, asym=False, device=local, primes=PRIMES30[:np_])
            ctx.set_stream(stream.cuda_stream)
            lb = 0
            while (2 << lb) * np_ * n * 4 <= cap_bytes and lb < 20:
                lb += 1
            batch = 1 << lb
            ctx.ntt_device(buf, batch)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                ctx.ntt_device(buf, batch)
            e1.record(stream)
            e1.synchronize()
            points.append((n, np_, lb, e0.elapsed_time(e1) / 5))
            ctx.close()
    ms = D.max_ms([p[3] for p in points])
    out = {}
    for (n, np_, lb, _), m in zip(points, ms):
        gbs = 8 * n * np_ * (1 << lb) / (m * 1e-3) / 1e9
        e = out.setdefault(f"n={n}", {}).setdefault(f"primes={np_}", {"log2_batch": [], "us": [], "GBps": [], "frac_of_hbm_peak": []})
        e["log2_batch"].append(lb)
        e["us"].append(round(m * 1e3, 2))
        e["GBps"].append(round(gbs, 1))
        e["frac_of_hbm_peak"].append(round(gbs / peak, 4))
    best = {}
    for nk, d in out.items():
        best[nk] = max(max(v["frac_of_hbm_peak"]) for v in d.values())
    del buf
    torch.cuda.empty_cache()
    return {"what": "k_ntt_forward only (seb_ntt_device), in place over [batch][primes][n] u32, per rank; time = max over ranks",
            "primes": "the first k of the reference's 30-bit primes (parameters.c:129-174) at every degree; roots = "
                      "seb_minimal_psi = the reference's table where it has an entry",
            "cap": f"points above {cap_bytes >> 30} GiB of polynomials are skipped", "peak_GBps": peak,
            "best_frac_of_hbm_peak": best, "points": out}


def run_b200_arm(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    seb = importlib.import_module("seal-embedded_b200")
    # before any pinned allocation: run next to this rank's GPU (matters for e2e at N > 1)
    # (at N = 1 too: a pinned buffer on the far socket costs the D2H copy 10-15 % on a two-socket host; the original
    # mask is restored before the CPU baselines, which want every core)
    all_cpus = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = seb.bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = Dist(torch, dist, world)
    n, np_, batch, vlen = N_DEG, N_PRIMES, args.batch, N_DEG // 2
    ctx = seb.Context(n, np_, asym=True, device=local)
    # a real key pair: random ternary secret key, public key generated on the GPU (seb_gen_public_key =
    # the reference's gen_pk, ckks_asym.c:159-171), so the ciphertexts of the timed steps can be decrypted
    sk = random_sk(n)
    ctx.gen_public_key(sk)
    ctx.reserve(batch)
    # a real (non-default) stream shared with torch, so torch.cuda.Event times the kernels' own stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    d_vals = torch.rand((batch, vlen), generator=gen, device="cuda", dtype=torch.float32) * 32 - 16
    d_seeds = torch.randint(0, 256, (batch, 64), generator=gen, device="cuda", dtype=torch.uint8)
    d_out = torch.empty((batch, np_, 2, n), dtype=torch.int32, device="cuda")
    barrier = D.barrier

    def step():
        ctx.encrypt_asym_device(d_vals, vlen, d_seeds, batch, d_out)

    for _ in range(args.warmup):
        step()
    assert ctx.encode_failures() == 0
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    ctx.profile_begin(args.steps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    kern_ms = ctx.profile_end()
    launches = ctx.launch_count - launches0
    clock_info = clocks.stop() if clocks else None
    ms_max = D.max_ms(ms)
    value = world * batch * args.steps / (ms_max * 1e-3)

    # ---- every ciphertext of the last timed step decrypts and decodes back to its message (GPU verifier,
    # the reference's acceptance criterion: within 0.1, device/test/ckks_tests_common.c:228)
    d_dec = torch.empty((batch, vlen), dtype=torch.float32, device="cuda")
    ctx.decrypt_decode_device(d_out, batch, 0, vlen, d_dec)
    torch.cuda.synchronize()
    verify_err = float((d_dec - d_vals).abs().max())
    del d_dec

    # ---- how many PRNG counters the ternary sampler CONSUMES per ciphertext (data dependent: 43 blocks + ~32 redraws at
    # n = 4096): the useful permutations its rate is counted in (it launches whole waves of 32, i.e. a few more)
    tb = min(batch, 4096)
    d_u = torch.empty(tb * n // 4, dtype=torch.uint8, device="cuda")
    d_e = torch.empty(tb * 2 * n, dtype=torch.int8, device="cuda")
    d_c = torch.zeros(tb, dtype=torch.int32, device="cuda")
    ctx.sample_asym_device(d_seeds[:tb].contiguous(), tb, d_u, d_e, d_c)
    torch.cuda.synchronize()
    ternary_counters = float(d_c.to(torch.float64).mean().item())
    del d_u, d_e, d_c

    # ---- the integer-issue ceilings of THIS device, now: register-only Keccak-f and lazy-butterfly loops inside the
    # library (seb_measure_ceilings): the denominators for the kernels that are not HBM-bound
    keccak_peak, bfly_peak = ctx.measure_ceilings()

    # ---- NTT-only micro-benchmark (config E shape: same n, primes; polys >> L2), rank-local
    polys = torch.randint(0, 1 << 30, (batch, np_, n), generator=gen, device="cuda", dtype=torch.int32)
    for _ in range(2):
        ctx.ntt_device(polys, batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ntt_iters = 5
    for _ in range(ntt_iters):
        ctx.ntt_device(polys, batch)
    e1.record(stream)
    torch.cuda.synchronize()
    ntt_ms = e0.elapsed_time(e1) / ntt_iters
    del polys

    # ---- e2e: host buffers (pinned) through the host-pointer C ABI, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        eb = args.e2e_batch or batch
        h_vals = torch.empty((eb, vlen), dtype=torch.float32).pin_memory()
        h_vals.copy_(d_vals[:eb])
        h_seeds = torch.empty((eb, 64), dtype=torch.uint8).pin_memory()
        h_seeds.copy_(d_seeds[:eb])
        h_out = torch.empty((eb, np_, 2, n), dtype=torch.int32).pin_memory()

        def e2e_step():
            rc = ctx.lib.seb_encrypt_asym_host(ctx.h, h_vals.data_ptr(), vlen, h_seeds.data_ptr(), eb,
                                               h_out.data_ptr())
            assert rc == 0, ctx.lib.seb_last_error()

        e2e_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(1, min(args.steps, 5))
        s0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        s1.record(stream)
        torch.cuda.synchronize()
        barrier()
        e2e_ms = D.max_ms(s0.elapsed_time(s1)) / e2e_steps
        same = bool(torch.equal(h_out[:64].cuda(), d_out[:64]))
        e2e = {"value": world * eb / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "batch_per_gpu": eb, "h2d_bytes_per_step": eb * (vlen * 4 + 64),
               "d2h_bytes_per_step": eb * (np_ * 2 * n * 4 + 4), "matches_device_path": same,
               "d2h_GBps": eb * (np_ * 2 * n * 4 + 4) * world / (e2e_ms * 1e-3) / 1e9 / world,
               "pcie": pcie_link(local),
               "api": "seb_encrypt_asym_host (pinned host buffers, 2 chunks in flight)", "cpu_binding": numa}
        # the optional packed wire form: 30-bit residues, 15 words per 16 coefficients (6.25 % fewer bytes on the link)
        if hasattr(ctx, "encrypt_asym_host_packed30"):
            pw = ctx.packed30_words()
            h_pk = torch.empty((eb, pw), dtype=torch.int32).pin_memory()

            def e2e_packed():
                ctx.encrypt_asym_host_packed30_raw(h_vals.data_ptr(), vlen, h_seeds.data_ptr(), eb, h_pk.data_ptr())

            e2e_packed()
            barrier()
            s0.record(stream)
            for _ in range(e2e_steps):
                e2e_packed()
            s1.record(stream)
            torch.cuda.synchronize()
            barrier()
            pk_ms = D.max_ms(s0.elapsed_time(s1)) / e2e_steps
            unpacked = ctx.unpack30_host(h_pk[:64].numpy().view(np.uint32))
            e2e["packed30"] = {"value": world * eb / (pk_ms * 1e-3), "unit": UNIT, "ms_per_step": pk_ms,
                               "d2h_bytes_per_step": eb * (pw * 4 + 4),
                               "unpacks_to_the_full_form": bool(np.array_equal(unpacked.view(np.int32), h_out[:64].numpy())),
                               "api": "seb_encrypt_asym_host_packed30 + seb_unpack30 (optional wire form)"}
            del h_pk
        del h_vals, h_seeds, h_out

    # ---- optional collation (BASELINE north_star: "a single NCCL all-gather only to collate outputs"): every
    # rank receives every rank's ciphertexts.  Off the throughput path; timed on a bounded slice so that the
    # gathered tensor stays small (world x 4096 items x 96 KiB = 3 GiB at N = 8).
    collate = None
    if world > 1 and not args.no_collate:
        cb = min(batch, 4096)
        local_ct = d_out[:cb].contiguous()
        g = seb.all_gather_ciphertexts(local_ct, cb * world)  # warm-up (NCCL channel setup)
        ok = bool(torch.equal(g[rank * cb:(rank + 1) * cb], local_ct))
        del g
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(3):
            g = seb.all_gather_ciphertexts(local_ct, cb * world)
        c1.record(stream)
        torch.cuda.synchronize()
        barrier()
        cms = D.max_ms(c0.elapsed_time(c1) / 3)
        recv = (world - 1) * cb * np_ * 2 * n * 4
        collate = {"what": "one NCCL all-gather of finished ciphertexts to every rank (off the throughput path)",
                   "items_per_gpu": cb, "ms": cms, "received_GB_per_gpu": recv / 1e9,
                   "receive_GBps_per_gpu": recv / 1e9 / (cms * 1e-3), "own_shard_intact": ok,
                   "ciphertexts_per_s_if_collated": world * cb / (cms * 1e-3)}
        del g, local_ct

    # ---- the other BASELINE configurations, same timing discipline (every rank takes part)
    ctx_launches_other = 0
    other = None
    peak, peak_src = hbm_peak()
    if not args.no_other:
        del d_out
        torch.cuda.empty_cache()
        osteps = max(1, min(args.steps, 3))
        other = {}
        other["C"] = run_full_config(seb, torch, D, stream, local, rank, world, "C: n=8192, 4-prime RNS, batch=262144, asymmetric "
                                     "(BASELINE.json configs[2])", 8192, 4, True, 262144, 32768, osteps, True)
        other["D"] = run_full_config(seb, torch, D, stream, local, rank, world, "D: n=16384, 6-prime RNS, batch=131072, symmetric "
                                     "(BASELINE.json configs[3])", 16384, 6, False, 131072, 16384, osteps, True)
        other["B_sym"] = run_full_config(seb, torch, D, stream, local, rank, world, "B in symmetric mode: n=4096, 3-prime RNS, "
                                         f"batch={batch}/GPU", 4096, 3, False, batch, batch, osteps, False)
        other["A"] = run_full_config(seb, torch, D, stream, local, rank, world, "A batched: n=1024, 1-prime RNS, symmetric, "
                                     f"batch={batch}/GPU (BASELINE.json configs[0] is the single-message CPU case)", 1024, 1,
                                     False, batch, batch, osteps, False)
        other["E"] = run_ntt_sweep(seb, torch, D, stream, local, peak)
        d_out = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (by measured share of the step)
    names = ctx.PROFILE_SEGMENTS[True]
    avg = kern_ms.mean(axis=0) if len(kern_ms) else np.zeros(4)
    alg_bytes = {  # algorithmic bytes per ciphertext for each kernel (DESIGN.md "Kernels")
        "encode": vlen * 4 + 8 * n,
        "sample_ternary": 64 + n // 4 + 4,
        "sample_cbd": 64 + 4 + 2 * n,
        "encrypt": 8 * n + 2 * n + n // 4 + 8 * n * np_,
    }
    sym_of = {"encode": "k_encode", "sample_ternary": "k_sample_ternary_pair" if n == 4096 else "k_sample_ternary",
              "sample_cbd": "k_sample_cbd", "encrypt": "k_encrypt_asym"}
    # what binds each kernel and the unit its ceiling is measured in: Keccak-f/s for the samplers (ALU pipe), lazy
    # butterflies/s for the fused encrypt (FMA pipe); the encode is FP64/shared-memory bound and is only given its
    # HBM view.  Units of work per ciphertext:
    bfly_per_ct = 3 * np_ * (n // 2) * (n.bit_length() - 1)
    work = {"sample_cbd": ("alu", "Keccak-f/s", 2 * (n // 16), keccak_peak),
            "sample_ternary": ("alu", "Keccak-f/s", ternary_counters, keccak_peak),  # USEFUL permutations (counters consumed)
            "encrypt": ("fma", "butterflies/s", bfly_per_ct, bfly_peak)}
    per_kernel = {}
    for nm, ms_k in zip(names, avg):
        if ms_k <= 0:
            continue
        gbs = alg_bytes[nm] * batch / (ms_k * 1e-3) / 1e9
        ent = {"ms": float(ms_k), "share_of_step": float(ms_k / max(avg.sum(), 1e-9)),
               "hbm": {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                       "algorithmic_bytes_per_launch": alg_bytes[nm] * batch, "traffic": ncu_traffic(sym_of[nm], batch)}}
        if nm in work:
            pipe, unit, per_ct, pk = work[nm]
            rate = per_ct * batch / (ms_k * 1e-3)
            ent.update({"bound": pipe, "achieved": rate, "peak": pk, "unit": unit, "frac": rate / pk if pk else None,
                        "work_per_ciphertext": per_ct})
        else:
            ent.update({"bound": "fp64/shared-memory (see DESIGN.md)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                        "frac": gbs / peak})
        per_kernel[sym_of[nm]] = ent
    top_name = names[int(np.argmax(avg))]
    top = per_kernel[sym_of[top_name]]
    roofline = {"kernel": sym_of[top_name], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
                "unit": top["unit"], "frac": top["frac"], "traffic": top["hbm"]["traffic"],
                "peak_source": "measured in this run: register-only loop of the same inner operation inside the library "
                               "(seb_measure_ceilings)" if top_name in work else peak_src,
                "avg_launch_ms": top["ms"], "share_of_step": top["share_of_step"], "hbm": top["hbm"],
                "note": "the step-dominant kernel is bound by integer issue (the pipe named in `bound`), not by HBM; its HBM "
                        "view is kept under `hbm`; the HBM-rooflined kernel BASELINE.json names is in ntt_microbench",
                "all_kernels": per_kernel}
    ntt_gbs = 8 * n * np_ * batch / (ntt_ms * 1e-3) / 1e9
    ntt_bf = batch * np_ * (n // 2) * (n.bit_length() - 1) / (ntt_ms * 1e-3)
    ntt_micro = {"kernel": "k_ntt_forward", "bound": "hbm", "achieved": ntt_gbs, "peak": peak, "unit": "GB/s",
                 "frac": ntt_gbs / peak, "traffic": ncu_traffic("k_ntt_forward", batch), "peak_source": peak_src,
                 "ms_per_launch": ntt_ms,
                 "polys_per_launch": batch * np_, "algorithmic_bytes_per_launch": 8 * n * np_ * batch,
                 "ntt_per_sec": batch * np_ / (ntt_ms * 1e-3),
                 "issue_bound": {"pipe": "fma", "achieved": ntt_bf, "peak": bfly_peak, "unit": "butterflies/s",
                                 "frac": ntt_bf / bfly_peak if bfly_peak else None}}
    ceilings = {"keccak_f_per_s": keccak_peak, "butterflies_per_s": bfly_peak, "hbm_GBps": peak, "hbm_source": peak_src,
                "how": "seb_measure_ceilings: register-only Keccak-f[1600] (24 full rounds x 174 ALU-pipe operations, the "
                       "bit-interleaved form) and Harvey/Shoup lazy-butterfly loops, best of 3 launches, CUDA events, in "
                       "this process.  A sampler call runs a SHORTER permutation (round 0 folded on the sponge's constant "
                       "lanes, last round pruned to the 96 bytes read: ~3930 operations instead of 4176) plus seed split "
                       "and extraction, so a fraction near 1.0 means the kernel's whole instruction stream runs at the "
                       "rate of the bare permutation"}

    # ---- CPU baseline: the reference's own path on this host's cores, bounded sample
    cpu = None
    cpu_others = None
    if all_cpus is not None and world == 1:
        os.sched_setaffinity(0, all_cpus)
    if world == 1 and not args.no_cpu:
        arm = CpuArm(items_per_worker=256)
        arm.step()
        secs = [arm.step() for _ in range(3)]
        arm.close()
        cpu = arm.describe(arm.cores * arm.items * len(secs) / sum(secs))
        if not args.no_other:
            cpu_others = {}
            for key, (cn, cnp, casym, citems) in {"A": (1024, 1, False, 4096), "B_sym": (4096, 3, False, 256),
                                                  "C": (8192, 4, True, 96), "D": (16384, 6, False, 32)}.items():
                try:
                    cpu_others[key] = cpu_baseline_for(cn, cnp, casym, citems)
                except Exception as e:  # noqa: BLE001 - reported baselines, never fatal
                    cpu_others[key] = {"unavailable": f"{type(e).__name__}: {e}"}

    # NTT-only CPU baseline: the reference's ntt_inpl, roots prebuilt — one core and all cores
    if cpu is not None:
        try:
            from oracle import oracle as O

            if O.have_reference():
                ref = O.ReferenceLib()
                reps = 2000
                ref.ntt_loop_seconds(n, np_, 0, 50)
                secs_ntt = ref.ntt_loop_seconds(n, np_, 0, reps)
                ntt_micro["cpu_reference"] = {"ntt_per_sec": reps / secs_ntt, "cores": 1, "kind": "reference",
                                              "sample": f"{reps} x ntt_inpl n={n} on one core, root table built once"}
                ncores = os.cpu_count() or 1
                with mp.get_context("spawn").Pool(ncores) as pool:
                    rates = pool.map(_ntt_loop_worker, [(n, np_, reps)] * ncores)
                ntt_micro["cpu_reference_all_cores"] = {"ntt_per_sec": float(sum(rates)), "cores": ncores, "kind": "reference",
                                                        "sample": f"{reps} x ntt_inpl n={n} in each of {ncores} processes"}
        except Exception as e:  # noqa: BLE001 - a reported baseline, never fatal
            ntt_micro["cpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": bench_config(batch, world),
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "verify": {"what": "decrypt+decode of every ciphertext of the last timed step on the GPU verifier",
                       "items": batch, "max_abs_err": verify_err, "tolerance": 0.1, "ok": verify_err < 0.1},
            "kernels_ms": {nm: float(v) for nm, v in zip(names, avg)},
            "roofline": roofline, "ntt_microbench": ntt_micro, "ceilings": ceilings, "cpu_baseline": cpu}
    if collate is not None:
        line["collate"] = collate
    if other is not None:
        if cpu_others:
            for key, v in cpu_others.items():
                if key in other:
                    other[key]["cpu_baseline"] = v
        line["other_configs"] = other
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _ntt_loop_worker(a):
    n, np_, reps = a
    from oracle import oracle as O

    ref = O.ReferenceLib()
    ref.ntt_loop_seconds(n, np_, 0, 50)
    return reps / ref.ntt_loop_seconds(n, np_, 0, reps)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of this process, on the process's original stdout."""
    data = (json.dumps(line, default=float) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner,
    # NCCL_DEBUG output, torchrun notices) is sent to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="ciphertexts per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=0, help="ciphertexts per GPU per e2e step (default: --batch)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-collate", action="store_true", help="skip the N>1 all-gather collation measurement")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs (C, D, B-sym, A, the NTT sweep)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
