#!/usr/bin/env python3
"""Development helper: SASS instruction counts per source line for one kernel of a cubin.

    cuobjdump -xelf all lib/composite.o && nvdisasm -g composite.sm_100a.cubin > sass.txt
    python tools/sass_lines.py sass.txt ncr_compositeILb1ELb0 [lo hi]

Prints, per composite.cu line (innermost line of the inlining chain), the number of SASS instructions and the
opcode mix; with lo/hi, only lines in that range and a total.  Static counts — use with ncu's source page for weights.
"""
import collections
import re
import sys


def main():
    path, kern = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10**9)
    cur, inside = None, False
    per = collections.defaultdict(collections.Counter)
    for ln in open(path):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "[^"]*composite\.cu", line (\d+)', ln)
        if m:
            if "inlined at" not in ln or cur is None or True:
                cur = int(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur is not None:
            per[cur][m.group(1)] += 1
    tot = collections.Counter()
    for line in sorted(per):
        if lo <= line <= hi:
            c = per[line]
            tot.update(c)
            print(f"{line:5d} {sum(c.values()):5d}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(8)))
    print("TOTAL", sum(tot.values()), " ".join(f"{k}:{v}" for k, v in tot.most_common(25)))


if __name__ == "__main__":
    main()
