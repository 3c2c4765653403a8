#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches.csv            > profiles/rNN_launches_*.txt
  python profiles/summarize_ncu.py full     gpurun_out/prof_step.ncu-rep       > profiles/rNN_ncu_full_*.txt

`launches`: per-kernel count / mean / min of gpu__time_duration.sum (the `--metrics gpu__time_duration.sum
--clock-control none` pass of B200_PROFILING.md) and each kernel's share of the summed device time.
`full`: the roofline-relevant raw metrics of every launch in an `ncu --set full` report (needs ncu on PATH).
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        agg.setdefault(row["Kernel Name"].split("(")[0][-48:], []).append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, gpu__time_duration.sum (ns); cold-cache, serialised")
    print(f"{'kernel':50s} {'n':>4s} {'mean_us':>10s} {'min_us':>10s} {'share':>7s}")
    for k, v in agg.items():
        print(f"{k:50s} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {min(v) / 1e3:10.1f} {sum(v) / total:7.3f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full --clock-control none, one launch per kernel")
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]][:100])
        for m in METRICS:
            if m in idx:
                print(f"  {m:66s} {r[idx[m]]:>16s} {units[idx[m]]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
