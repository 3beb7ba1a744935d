"""Condenses Nsight Compute output into the small text summaries kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv>        per-kernel totals / shares of a launch list
  python tools/ncu_summary.py report <file.ncu-rep>          key metrics of every captured launch (`--set full`)
  python tools/ncu_summary.py traffic <file.ncu-rep> <key> <summary file> [json]
                                                             dram bytes per launch of the FIRST captured launch -> entry `key`
                                                             (e.g. config1) of profiles/traffic.json, which bench.py reads for
                                                             `roofline.traffic`
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

KEYS = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
)


def launches(path: str) -> None:
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {100 * v[1] / tot:6.1f}%")
    print(f"{'total':44s} {sum(v[0] for v in agg.values()):8d} {tot:12.1f}")


def report(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for row in data:
        print(f"== {row[ki].split('(')[0]}  grid {row[hdr.index('Grid Size')]}  block {row[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:88s} {row[i]:>16s} {units[i]}")


def traffic(path: str, key: str, summary: str, out: str = "profiles/traffic.json") -> None:
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, row = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        v = float(row[i].replace(",", ""))
        u = units[i].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
                 "ms": 1.0, "msecond": 1.0, "%": 1.0}.get(u, 1.0)
        return v * scale

    entry = {"kernel": row[hdr.index("Kernel Name")].split("(")[0], "grid": row[hdr.index("Grid Size")],
             "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
             "duration_ms_under_ncu": val("gpu__time_duration.sum"),
             "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "l2_hit_pct": val("lts__t_sector_hit_rate.pct"), "source": summary,
             "how": "ncu --set full --clock-control none, first captured launch; per launch like roofline.achieved"}
    data = {}
    if os.path.isfile(out):
        with open(out) as f:
            data = json.load(f)
    data[key] = entry
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(entry))


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
