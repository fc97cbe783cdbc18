"""Pivot an `ncu --csv --metrics ...` launch list into one line per launch.
Usage: python profiles/launch_table.py gpurun_out/launches.csv [last N launches per kernel]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def main():
    path = sys.argv[1]
    rows = OrderedDict()
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        key = int(r["ID"])
        d = rows.setdefault(key, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        val = r["Metric Value"].replace(",", "")
        try:
            val = float(val)
        except ValueError:
            pass
        d[r["Metric Name"]] = (val, r["Metric Unit"])
    short = lambda s: re.sub(r"\(.*", "", s)[:70]
    seen = defaultdict(int)
    keep_last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    items = list(rows.items())
    if keep_last:
        total = defaultdict(int)
        for _, d in items:
            total[d["name"]] += 1
    for key, d in items:
        seen[d["name"]] += 1
        if keep_last and seen[d["name"]] <= total[d["name"]] - keep_last:
            continue
        g = lambda m: d.get(m, (float("nan"), ""))
        dur, du = g("gpu__time_duration.sum")
        if du == "ns":
            dur /= 1e3
        elif du == "ms":
            dur *= 1e3
        rd, ru = g("dram__bytes_read.sum")
        wr, wu = g("dram__bytes_write.sum")
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        rd *= scale.get(ru, 1e-6)
        wr *= scale.get(wu, 1e-6)
        inst = g("smsp__inst_executed.sum")[0]
        print("%5d %-70s grid %-10s %8.1f us  rd %8.1f MB wr %8.1f MB  %6.2f TB/s  inst %8.2fM issue %5.1f%% warps %5.1f%% regs %3.0f" % (
            key, short(d["name"]), d["grid"].strip("()").split(",")[0], dur, rd, wr, (rd + wr) / dur / 1e6 * 1e6 / 1e6 if dur else 0, inst / 1e6,
            g("smsp__issue_active.avg.pct_of_peak_sustained_active")[0], g("sm__warps_active.avg.pct_of_peak_sustained_active")[0],
            g("launch__registers_per_thread")[0]))


if __name__ == "__main__":
    main()
