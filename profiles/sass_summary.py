"""Instruction histogram per cubin of libsparray_b200.so (cuobjdump -sass): the mnemonics that show what the
kernels are made of -- TMA bulk copies (UBLKCP), mbarrier traffic (SYNCS), 128-bit global accesses, atomics,
votes / matches / shuffles.  Usage: python profiles/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OBJ = os.path.join(HERE, "..", "sparray_b200", "build")
WATCH = ["UBLKCP", "SYNCS", "LDG.E.128", "STG.E.128", "LDG.E.64", "STG.E.64", "LDGSTS", "ATOMS", "ATOMG", "RED",
         "ATOM", "MATCH", "VOTE", "SHFL", "BAR", "LDS", "STS", "NANOSLEEP", "ACQBULK", "UTMA", "HMMA", "UTC"]


def main():
    print("# SASS summary of libsparray_b200.so (sm_100a), one block per translation unit; counts are static")
    print("# instruction counts over all kernels of the unit (cuobjdump -sass)\n")
    for obj in sorted(f for f in os.listdir(OBJ) if f.endswith(".o")):
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
        kernels = re.findall(r"Function : (\S+)", out)
        ops = collections.Counter()
        total = 0
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            ops[m.group(1)] += 1
        print("== %s: %d kernels, %d instructions" % (obj, len(kernels), total))
        for w in WATCH:
            c = sum(v for k, v in ops.items() if k == w or k.startswith(w + ".") or (w.count(".") and k.startswith(w)))
            if c:
                print("   %-10s %6d" % (w, c))
        print()


if __name__ == "__main__":
    main()
