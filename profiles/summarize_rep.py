"""Text summary of an ncu report (--set full --import-source on): per launch the headline metrics and stall
mix, then per kernel the source lines that carry the most executed instructions / stall samples.
Usage: python profiles/summarize_rep.py report.ncu-rep [top_lines] > summary.txt"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("time us", "gpu__time_duration.sum"), ("dram rd MB", "dram__bytes_read.sum"), ("dram wr MB", "dram__bytes_write.sum"),
    ("dram % peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warp inst", "smsp__inst_executed.sum"), ("issue act %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("smem/CTA B", "launch__shared_mem_per_block_static"),
    ("dyn smem B", "launch__shared_mem_per_block_dynamic"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % rep.split("/")[-1])
    print("## launches")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("\n%s" % name[:150])
        parts = []
        for label, m in METRICS:
            if m in hdr:
                v, u = r[hdr.index(m)], units[hdr.index(m)]
                try:
                    f = float(v)
                    if u == "byte":
                        f /= 1e6
                    elif u in ("ns", "nsecond"):
                        f /= 1e3
                    v = ("%.1f" % f) if f < 1e6 else ("%.3g" % f)
                except ValueError:
                    pass
                parts.append("%s %s" % (label, v))
        print("  " + " | ".join(parts))
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  stalls per issue: " + ", ".join("%s %.1f" % (n, v) for v, n in stalls[:6]))
    print("\n## hot source lines (share of executed warp instructions / of stall samples, per kernel)")
    out = ncu(rep, "--page", "source", "--print-source", "sass,cuda", "--csv")
    fname = fn = hdr2 = None
    agg = collections.OrderedDict()
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] == "Line No":
            hdr2 = r
            continue
        if hdr2 is None or len(r) < len(hdr2) or r[0] in ("", "-"):
            continue
        try:
            ins = int(r[hdr2.index("Instructions Executed")] or 0)
            st = int(r[hdr2.index("Warp Stall Sampling (All Samples)")] or 0)
        except ValueError:
            continue
        a = agg.setdefault(fn, collections.OrderedDict()).setdefault((fname, r[0]), [r[1].strip()[:110], 0, 0])
        a[1] += ins
        a[2] += st
    for fn, lines in agg.items():
        ti = sum(a[1] for a in lines.values()) or 1
        ts = sum(a[2] for a in lines.values()) or 1
        print("\n%s   (warp inst %d, stall samples %d; launches of this kernel in the report are summed)" % (fn[:150], ti, ts))
        ranked = sorted(lines.items(), key=lambda kv: -(kv[1][1] / ti + kv[1][2] / ts))[:top]
        for (f, l), (src, ins, st) in ranked:
            print("  %-14s %5s ex=%5.1f%% st=%5.1f%%  %s" % (f, l, 100.0 * ins / ti, 100.0 * st / ts, src))


if __name__ == "__main__":
    main()
