"""Source-correlated hot lines of an ncu report: share of executed warp instructions and of stall
samples per CUDA source line.  Usage: python profiles/hot_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr = None, None
    agg = collections.OrderedDict()
    tot_s = tot_i = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or fname is None or len(r) < len(hdr):
            continue
        try:
            st = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
            ins = int(r[hdr.index("Instructions Executed")] or 0)
        except ValueError:
            continue
        if r[0] in ("", "-"):          # SASS rows under a source line: already counted in the line's row
            continue
        key = (fname, r[0])
        if key not in agg:
            agg[key] = [r[1].strip()[:100], 0, 0]
        agg[key][1] += st
        agg[key][2] += ins
        tot_s += st
        tot_i += ins
    print("tot warp-inst %d stall samples %d" % (tot_i, tot_s))
    for (f, l), (src, st, ins) in sorted(agg.items(), key=lambda kv: -(kv[1][1] / max(tot_s, 1) + kv[1][2] / max(tot_i, 1)))[:top]:
        print("%-12s %5s ex=%5.1f%% st=%5.1f%%  %s" % (f, l, 100.0 * ins / max(tot_i, 1), 100.0 * st / max(tot_s, 1), src))


if __name__ == "__main__":
    main()
