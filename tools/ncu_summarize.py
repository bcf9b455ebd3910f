"""Turn ncu outputs into the small text summaries committed under profiles/.

  python tools/ncu_summarize.py launches <launches.csv>         -> per-kernel count / total / share table
  python tools/ncu_summarize.py kernel <report.ncu-rep> [idx]   -> key metrics of one captured launch
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((r["Kernel Name"].split("(")[0], v * scale))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"{'kernel':60s} {'launches':>9s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:9d} {us:12.1f} {us / n:10.2f} {100 * us / total:6.1f}%")
    print(f"{'TOTAL':60s} {len(rows):9d} {total:12.1f}")


def kernel(rep, idx=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    row = data[int(idx)]
    col = {n: i for i, n in enumerate(hdr)}
    print("kernel:", row[col["Kernel Name"]])
    for k in KEYS:
        if k in col:
            print(f"  {k:70s} {row[col[k]]:>16s} {units[col[k]]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(*sys.argv[2:])
