#!/usr/bin/env python
"""bench.py — CKKS encryptions/s at n=4096, 3-prime RNS (BASELINE.json configs[1]) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (N>1: under torchrun)
  python bench.py --impl reference [--gpus N] --steps K --warmup W the reference's CPU path on this host

One step = one pass of the hot path (encode -> sample u,e0,e1 -> 3 primes x [3 NTT + pointwise]) over
one batch of 65536 synthetic fp32 messages per GPU, inputs resident in HBM.  Ciphertexts are
independent, so N GPUs shard the batch with no data-path collective ("weak" scaling: the per-GPU
batch is fixed).  Prints ONE JSON line (rank 0).
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
# integer-issue ceilings measured on this pool's B200 with tools/ubench (profiles/r01_ubench_int_pipes.txt,
# profiles/r01_ubench_bfly.txt): Keccak-f[1600] is ALU-pipe bound (LOP3/SHF at 64 lanes/clk/SM), the lazy
# butterfly FMA-pipe bound (IMAD.HI + 2 IMAD)
KECCAK_PEAK_PER_S = 4.26e9
BUTTERFLY_PEAK_PER_S = 3.77e12
METRIC = "ckks_encryptions_per_sec"
UNIT = "ciphertexts/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = the unmodified reference compiled here; else the oracle port)
# ----------------------------------------------------------------------------------------------
def _cpu_worker(conn, idx: int, n: int, np_: int, items: int):
    """One process per core: the reference keeps static state (seal_embedded.c:18-22)."""
    from oracle import oracle as O

    orc = O.Oracle()
    sk = O.make_sk(n)
    pk0, pk1 = orc.gen_pk(n, np_, sk)
    vals = O.make_values(items, n // 2, seed=1000 + idx)
    seeds = O.make_seeds(items, b"cpu-%d" % idx)
    ref = None
    if O.have_reference():
        ref = O.ReferenceLib()
        ref.setup(n, np_, True, sk=sk, pk0=pk0, pk1=pk1, primes=orc.primes(n, np_))
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        t0 = time.perf_counter()
        if ref is not None:
            ref.encrypt_loop(None, seeds, vals)
        else:
            orc.encrypt_asym_batch(n, np_, vals, seeds, pk0, pk1)
        conn.send(time.perf_counter() - t0)
    if ref is not None:
        ref.close()
    conn.close()


class CpuArm:
    def __init__(self, items_per_worker: int, cores: int | None = None):
        from oracle import oracle as O

        O.build()
        self.kind = "reference" if O.have_reference() else "port"
        self.cores = cores or (os.cpu_count() or 1)
        self.items = items_per_worker
        ctx = mp.get_context("spawn")
        self.procs, self.conns = [], []
        for i in range(self.cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, i, N_DEG, N_PRIMES, items_per_worker), daemon=True)
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
                          f"n={N_DEG} {N_PRIMES} primes asym (one process per core: the reference is non-reentrant)"}


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
            "config": {"workload": WORKLOAD, "n": N_DEG, "nprimes": N_PRIMES,
                       "batch_per_step": arm.cores * arm.items, "host_threads": arm.cores,
                       "sample": f"{arm.items} of the workload's items per process per step"},
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
    """dram read+write bytes per launch from the committed `ncu --set full` summary (profiles/traffic.json,
    written by tools/summarize_profile.py), scaled from the batch it was captured at to this run's."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p)).get(kernel)
            return float(t["dram_bytes_per_launch"]) * batch / float(t["batch"]) if t else None
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_b200_arm(args) -> None:
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    seb = importlib.import_module("seal-embedded_b200")
    # before any pinned allocation: run next to this rank's GPU (matters for e2e at N > 1)
    numa = seb.bind_to_gpu_numa(local) if world > 1 else "not bound (single rank)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, np_, batch, vlen = N_DEG, N_PRIMES, args.batch, N_DEG // 2
    ctx = seb.Context(n, np_, asym=True, device=local)
    # a real key pair: random ternary secret key, public key generated on the GPU (seb_gen_public_key =
    # the reference's gen_pk, ckks_asym.c:159-171), so the ciphertexts of the timed steps can be decrypted
    rng = np.random.default_rng(7)
    t = rng.integers(0, 3, (n // 4, 4), dtype=np.uint8)
    sk = ((t[:, 0] << 6) | (t[:, 1] << 4) | (t[:, 2] << 2) | t[:, 3]).astype(np.uint8)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * batch * args.steps / (ms_max * 1e-3)

    # ---- every ciphertext of the last timed step decrypts and decodes back to its message (GPU verifier,
    # the reference's acceptance criterion: within 0.1, device/test/ckks_tests_common.c:228)
    d_dec = torch.empty((batch, vlen), dtype=torch.float32, device="cuda")
    ctx.decrypt_decode_device(d_out, batch, 0, vlen, d_dec)
    torch.cuda.synchronize()
    verify_err = float((d_dec - d_vals).abs().max())
    del d_dec

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
        t2 = torch.tensor([s0.elapsed_time(s1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_ms = float(t2.item()) / e2e_steps
        same = bool(torch.equal(h_out[:64].cuda(), d_out[:64]))
        e2e = {"value": world * eb / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "batch_per_gpu": eb, "h2d_bytes_per_step": eb * (vlen * 4 + 64),
               "d2h_bytes_per_step": eb * (np_ * 2 * n * 4 + 4), "matches_device_path": same,
               "d2h_GBps": eb * (np_ * 2 * n * 4 + 4) * world / (e2e_ms * 1e-3) / 1e9 / world,
               "pcie": pcie_link(local),
               "api": "seb_encrypt_asym_host (pinned host buffers, 2 chunks in flight)", "cpu_binding": numa}

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
        t3 = torch.tensor([c0.elapsed_time(c1) / 3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        cms = float(t3.item())
        recv = (world - 1) * cb * np_ * 2 * n * 4
        collate = {"what": "one NCCL all-gather of finished ciphertexts to every rank (off the throughput path)",
                   "items_per_gpu": cb, "ms": cms, "received_GB_per_gpu": recv / 1e9,
                   "receive_GBps_per_gpu": recv / 1e9 / (cms * 1e-3), "own_shard_intact": ok,
                   "ciphertexts_per_s_if_collated": world * cb / (cms * 1e-3)}
        del g

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
    top = int(np.argmax(avg))
    peak, peak_src = hbm_peak()
    top_name = names[top]
    achieved = alg_bytes[top_name] * batch / (avg[top] * 1e-3) / 1e9 if avg[top] > 0 else 0.0
    kernel_sym = {"encode": "k_encode", "sample_ternary": "k_sample_ternary", "sample_cbd": "k_sample_cbd",
                  "encrypt": "k_encrypt_asym"}[top_name]
    roofline = {"kernel": kernel_sym, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kernel_sym, batch), "peak_source": peak_src,
                "avg_launch_ms": float(avg[top]), "share_of_step": float(avg[top] / max(avg.sum(), 1e-9)),
                "algorithmic_bytes_per_launch": alg_bytes[top_name] * batch,
                "note": "integer-issue bound kernel (Keccak / modular butterflies), not an HBM-bound one: "
                        "see issue_bound here and ntt_microbench for the HBM-rooflined NTT-only kernel"}
    # what actually bounds the Keccak / butterfly kernels: integer issue, against the measured pipe ceilings
    keccak_per_ct = {"sample_cbd": 2 * (n // 16), "sample_ternary": 96}  # permutations launched per ciphertext
    if top_name in keccak_per_ct and avg[top] > 0:
        rate = keccak_per_ct[top_name] * batch / (avg[top] * 1e-3)
        roofline["issue_bound"] = {"pipe": "alu", "achieved": rate, "peak": KECCAK_PEAK_PER_S, "unit": "Keccak-f/s",
                                   "frac": rate / KECCAK_PEAK_PER_S,
                                   "peak_source": "tools/ubench best Keccak-f variant on this pool (profiles/)"}
    elif top_name == "encrypt" and avg[top] > 0:
        rate = 3 * np_ * (n // 2) * (n.bit_length() - 1) * batch / (avg[top] * 1e-3)
        roofline["issue_bound"] = {"pipe": "fma", "achieved": rate, "peak": BUTTERFLY_PEAK_PER_S, "unit": "butterflies/s",
                                   "frac": rate / BUTTERFLY_PEAK_PER_S,
                                   "peak_source": "tools/ubench_bfly register-only butterflies (profiles/)"}
    # every kernel of the step against the HBM peak (algorithmic bytes / measured duration), beside the dominant one above
    sym_of = {"encode": "k_encode", "sample_ternary": "k_sample_ternary", "sample_cbd": "k_sample_cbd", "encrypt": "k_encrypt_asym"}
    per_kernel = {}
    for nm, ms_k in zip(names, avg):
        if ms_k > 0:
            gbs = alg_bytes[nm] * batch / (ms_k * 1e-3) / 1e9
            per_kernel[sym_of[nm]] = {"ms": float(ms_k), "share_of_step": float(ms_k / max(avg.sum(), 1e-9)),
                                      "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
    roofline["all_kernels"] = per_kernel
    ntt_gbs = 8 * n * np_ * batch / (ntt_ms * 1e-3) / 1e9
    ntt_micro = {"kernel": "k_ntt_forward", "bound": "hbm", "achieved": ntt_gbs, "peak": peak, "unit": "GB/s",
                 "frac": ntt_gbs / peak, "traffic": ncu_traffic("k_ntt_forward", batch), "ms_per_launch": ntt_ms,
                 "polys_per_launch": batch * np_, "algorithmic_bytes_per_launch": 8 * n * np_ * batch,
                 "ntt_per_sec": batch * np_ / (ntt_ms * 1e-3),
                 "issue_bound": {"pipe": "fma", "achieved": batch * np_ * (n // 2) * (n.bit_length() - 1) / (ntt_ms * 1e-3),
                                 "peak": BUTTERFLY_PEAK_PER_S, "unit": "butterflies/s",
                                 "frac": batch * np_ * (n // 2) * (n.bit_length() - 1) / (ntt_ms * 1e-3) / BUTTERFLY_PEAK_PER_S}}

    # ---- CPU baseline: the reference's own path on this host's cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        arm = CpuArm(items_per_worker=256)
        arm.step()
        secs = [arm.step() for _ in range(3)]
        arm.close()
        cpu = arm.describe(arm.cores * arm.items * len(secs) / sum(secs))

    # NTT-only CPU baseline: the reference's ntt_inpl, one host core, roots prebuilt
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
        except Exception as e:  # noqa: BLE001 - a reported baseline, never fatal
            ntt_micro["cpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD.replace(f"batch={BATCH}/GPU", f"batch={batch}/GPU"), "n": n, "nprimes": np_,
                       "batch_per_gpu": batch,
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "l2": f"inputs+outputs per step ({(d_vals.numel() * 4 + d_out.numel() * 4) >> 20} MiB) "
                             "exceed the 126 MB L2"},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "verify": {"what": "decrypt+decode of every ciphertext of the last timed step on the GPU verifier",
                       "items": batch, "max_abs_err": verify_err, "tolerance": 0.1, "ok": verify_err < 0.1},
            "kernels_ms": {nm: float(v) for nm, v in zip(names, avg)},
            "roofline": roofline, "ntt_microbench": ntt_micro, "cpu_baseline": cpu}
    if collate is not None:
        line["collate"] = collate
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
