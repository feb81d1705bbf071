#!/usr/bin/env python
"""Turn gpurun_out/<tag>/{launches.csv,hotpath.ncu-rep} into committed summaries under profiles/.

  python tools/summarize_profile.py <launch_tag> <ncu_tag> <round>     e.g.  r01a r01b r01

Writes profiles/<round>_launches.csv (per-kernel launch list, ncu gpu__time_duration pass),
profiles/<round>_ncu_summary.csv (selected `ncu --set full` metrics per kernel) and updates
profiles/traffic.json (dram bytes per launch + the batch it was captured at) which bench.py reads
for `roofline.traffic`.
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NCU_BATCH = int(os.environ.get("NCU_BATCH", "65536"))  # ciphertexts per launch in the captured run (tools/gpu_profile.sh)
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_selected",
]


def short(name: str) -> str:
    name = name.replace("void ", "")
    return name.split("(")[0]


def launches(tag: str, rnd: str) -> None:
    src = os.path.join(ROOT, "gpurun_out", tag, "launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size"))
    out = os.path.join(ROOT, "profiles", f"{rnd}_launches.csv")
    agg = collections.OrderedDict()
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu_time_ns"])
        for r in rows[1:]:
            k = short(r[ki])
            ns = float(r[vi].replace(",", ""))
            w.writerow([r[0], k, r[gi], r[bi], int(ns)])
            if k.startswith("k_"):
                agg.setdefault(k, []).append(ns)
    # the step = the four kernels of the asymmetric full path, at the bench's grid (setup-time launches of
    # the same kernels with tiny grids — key generation, ntt(s) — and the verifier are listed but not counted)
    hot = ("k_encode", "k_sample_ternary", "k_sample_ternary_pair", "k_sample_cbd", "k_encrypt_asym")
    means = {}
    for k, v in agg.items():
        big = [x for x in v if x >= 0.5 * max(v)]
        means[k] = (sum(big) / len(big), len(big), len(v))
    tot = sum(m for k, (m, _, _) in means.items() if k.split("<")[0] in hot)
    print(f"wrote {out}")
    for k, (m, nbig, nall) in means.items():
        base = k.split("<")[0]
        share = f"{100 * m / tot:5.1f}% of step" if base in hot else "(not part of the step)"
        print(f"  {k:28s} launches {nall:3d} ({nbig} at full grid)  mean {m / 1e3:9.1f} us  {share}")


def ncu_full(tag: str, rnd: str) -> None:
    import glob

    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag, "*.ncu-rep")))
    out = os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.csv")
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = {}
    seen = set()
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for rep in reps:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                                 check=True).stdout
            rows = list(csv.reader(raw.splitlines()))
            hdr, units = rows[0], rows[1]
            idx = {h: i for i, h in enumerate(hdr)}
            for r in rows[2:]:
                k = short(r[idx["Kernel Name"]])
                if k in seen:
                    continue
                seen.add(k)
                for m in KEEP:
                    if m in idx:
                        w.writerow([k, m, units[idx[m]], r[idx[m]]])

                def val(m):
                    v = float(r[idx[m]].replace(",", ""))
                    u = units[idx[m]].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

                base = k.split("<")[0]
                traffic[base] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                                 "grid": r[idx["launch__grid_size"]], "batch": NCU_BATCH,
                                 "source": f"profiles/{rnd}_ncu_summary.csv"}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print(f"wrote {out} and {traffic_path}: {sorted(seen)}")


if __name__ == "__main__":
    ltag, ntag, rnd = sys.argv[1:4]
    if ltag != "-":
        launches(ltag, rnd)
    if ntag != "-":
        ncu_full(ntag, rnd)
