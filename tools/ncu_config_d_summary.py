#!/usr/bin/env python
"""Selected metrics of the captures made by tools/ncu_config_d.sh -> one JSON file under profiles/.
  python tools/ncu_config_d_summary.py <tag> <out.json>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def main() -> None:
    tag, out = sys.argv[1], sys.argv[2]
    res = {}
    for name in ("k_uniform_bulk", "k_uniform_fix"):
        rep = os.path.join(ROOT, "gpurun_out", tag, f"d_{name}.ncu-rep")
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        ent = {"kernel": vals[hdr.index("Kernel Name")][:70]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                ent[k] = f"{vals[i]} {units[i]}".strip()
        res[name] = ent
    res["_what"] = ("ncu --set full, one launch each, configuration D's shard (n=16384, 6 primes, 16384 items): "
                    "bash tools/ncu_config_d.sh; the library's own choice of kernels")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
