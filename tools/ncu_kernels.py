"""Per-kernel table from an `ncu --set full` report with many captured launches:

  python tools/ncu_kernels.py <report.ncu-rep> [<report2.ncu-rep> ...] --json out.json --txt out.txt [--tag NAME]

For every kernel name: launches captured, mean duration, DRAM bytes read / written per launch, achieved DRAM GB/s,
DRAM throughput (% of peak), issue-active %, active warps %, registers, block / grid size, dynamic shared memory.
The numbers come from a profiler run (kernels serialised, caches cold, ~40 replays per launch): use them for traffic and
shares, not as bench values.
"""
import collections
import csv
import io
import json
import subprocess
import sys

M = {
    "dur_ns": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
    "block": "launch__block_size", "grid": "launch__grid_size", "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "l2_hit": "lts__t_sector_hit_rate.pct", "inst": "smsp__inst_executed.sum",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9,
              "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6, "second": 1e9}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    col = {n: i for i, n in enumerate(hdr)}
    rows = []
    for r in data:
        d = {"name": r[col["Kernel Name"]]}
        for k, m in M.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                d[k] = v * UNIT_SCALE.get(units[col[m]], 1.0) if k in ("dur_ns", "dram_rd", "dram_wr", "smem_dyn") else v
        rows.append(d)
    return rows


def main():
    args = sys.argv[1:]
    reps, jpath, tpath, tag = [], None, None, ""
    while args:
        a = args.pop(0)
        if a == "--json":
            jpath = args.pop(0)
        elif a == "--txt":
            tpath = args.pop(0)
        elif a == "--tag":
            tag = args.pop(0)
        else:
            reps.append(a)
    agg = collections.OrderedDict()
    for rep in reps:
        for d in load(rep):
            agg.setdefault(d["name"].split("(")[0].replace("void ", "").replace("b200c::", ""), []).append(d)
    kernels = {}
    lines = [f"{'kernel':44s} {'n':>4s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'dram%':>6s} {'issue%':>6s} "
             f"{'warps%':>6s} {'regs':>5s} {'block':>6s} {'grid':>8s} {'smem KB':>8s}"]
    for name, ds in agg.items():
        def mean(k):
            v = [d[k] for d in ds if k in d]
            return sum(v) / len(v) if v else None
        us = mean("dur_ns") / 1e3
        rd, wr = mean("dram_rd") or 0.0, mean("dram_wr") or 0.0
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us else 0.0
        base = name.split("<")[0]
        rec = {"launches_captured": len(ds), "us": us, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_gbs": gbs,
               "dram_pct_of_peak": mean("dram_pct"), "issue_active_pct": mean("issue_pct"), "warps_active_pct": mean("warps_pct"),
               "registers": mean("regs"), "block": mean("block"), "grid": mean("grid"), "smem_dyn_bytes": mean("smem_dyn"),
               "l2_hit_pct": mean("l2_hit")}
        kernels[name] = rec
        if base not in kernels or kernels[base]["us"] < us:
            kernels[base] = rec  # the template-free name maps to its heaviest instance
        lines.append(f"{name[:44]:44s} {len(ds):4d} {us:9.2f} {rd / 1e6:9.3f} {wr / 1e6:9.3f} {gbs:8.1f} "
                     f"{(rec['dram_pct_of_peak'] or 0):6.1f} {(rec['issue_active_pct'] or 0):6.1f} {(rec['warps_active_pct'] or 0):6.1f} "
                     f"{int(rec['registers'] or 0):5d} {int(rec['block'] or 0):6d} {int(rec['grid'] or 0):8d} {(rec['smem_dyn_bytes'] or 0) / 1e3:8.1f}")
    txt = "\n".join(lines)
    print(txt)
    if tpath:
        with open(tpath, "w") as f:
            f.write((f"# {tag}\n" if tag else "") + txt + "\n")
    if jpath:
        try:
            old = json.load(open(jpath))
        except Exception:
            old = {"kernels": {}}
        old.setdefault("captures", {})[tag or "default"] = {k: v for k, v in kernels.items()}
        if not tag or tag.startswith("single"):
            old["kernels"].update(kernels)
        json.dump(old, open(jpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
