#!/usr/bin/env python
"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.
  python tools/summarize_ncu.py launches gpurun_out/launches.csv
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        key = (name, row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0, 1e30, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = min(a[2], v)
        a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time over the captured launch window (ncu gpu__time_duration.sum; cold-cache,")
    print("# serialised launches: compare SHARES, not absolutes)")
    print("%-28s %-16s %-14s %6s %12s %10s %10s %10s %7s" % ("kernel", "grid", "block", "n", "total_us", "avg_us",
                                                            "min_us", "max_us", "share"))
    for (name, grid, block), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %-16s %-14s %6d %12.1f %10.2f %10.2f %10.2f %6.1f%%" % (name[:28], grid, block, a[0], a[1],
                                                                           a[1] / a[0], a[2], a[3], 100 * a[1] / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                print("  %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        print()


def traffic(path, key, note=""):
    """DRAM bytes per launch of the first captured kernel -> profiles/ncu_traffic.json[key] (read by bench.py)."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name):
        v = float(r[hdr.index(name)].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(name)]]
    total = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    rec = json.load(open(dst)) if os.path.exists(dst) else {}
    rec[key] = {"dram_bytes_per_launch": int(total), "kernel": r[hdr.index("Kernel Name")][:120],
                "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), %s%s" % (os.path.basename(path),
                                                                                               (": " + note) if note else "")}
    json.dump(rec, open(dst, "w"), indent=1)
    print(json.dumps(rec[key]))


if __name__ == "__main__":
    cmd = {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]]
    cmd(*sys.argv[2:])
