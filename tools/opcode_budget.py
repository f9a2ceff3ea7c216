#!/usr/bin/env python3
"""Development helper: where a kernel's executed warp instructions go, basic block by basic block.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/opcode_budget.py sass.csv [min_share_pct] [units]

Consecutive SASS instructions with the same executed count are taken as one straight-line block.  For every block above
`min_share_pct` of all executed warp instructions the tool prints its position, length, executions, share, its stall-sample
share and its opcode mix; `units` (e.g. the number of list entries the launch walked) adds "executions per unit".  The totals
by opcode class (f64 / integer+logic / shared memory / global memory / conversions / control) close the table."""
import collections
import csv
import sys

CLASSES = {
    "f64": ("DMUL", "DADD", "DFMA", "DSETP", "DMNMX"),
    "shared memory": ("LDS", "STS", "LDSM"),
    "global/local memory": ("LDG", "STG", "LDL", "STL", "ATOMG", "RED", "LD", "ST", "LDC", "LDCU", "ULDC"),
    "conversion (XU)": ("F2I", "I2F", "F2F", "MUFU", "I2FP"),
    "control": ("BRA", "BSSY", "BSYNC", "BREAK", "CALL", "RET", "EXIT", "WARPSYNC", "NOP", "BAR", "YIELD", "BPT"),
    "warp (shuffle/vote)": ("SHFL", "VOTE", "MATCH", "REDUX"),
}


def opcode(src: str) -> str:
    parts = src.split()
    if not parts:
        return "?"
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    return op.split(".")[0]


def klass(op: str) -> str:
    for name, ops in CLASSES.items():
        if op in ops:
            return name
    return "integer / logic / select / move"


def main():
    path = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    inst = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            inst.append((int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0), opcode(r[col["Source"]].strip())))
        except ValueError:
            continue
    total = sum(i[0] for i in inst)
    samples = sum(i[1] for i in inst)
    blocks = []   # [first index, executions, [opcodes], samples]
    for k, (ex, sm, op) in enumerate(inst):
        if blocks and blocks[-1][1] == ex:
            blocks[-1][2].append(op)
            blocks[-1][3] += sm
        else:
            blocks.append([k, ex, [op], sm])
    print(f"executed warp instructions {total:,}; stall samples {samples:,}; static instructions {len(inst):,}" +
          (f"; units {units:,.0f} -> {total / units:.1f} instructions per unit" if units else ""))
    print()
    print("| first SASS index | instructions | executions | per unit | share of executed | share of stall samples | opcode mix |")
    print("|---|---|---|---|---|---|---|")
    shown = 0.0
    for b in blocks:
        share = 100.0 * b[1] * len(b[2]) / max(1, total)
        if share < min_share:
            continue
        shown += share
        mix = collections.Counter(b[2]).most_common(7)
        per = f"{b[1] / units:.3f}" if units else "—"
        print(f"| {b[0]} | {len(b[2])} | {b[1]:,} | {per} | {share:.1f} % | {100.0 * b[3] / max(1, samples):.1f} % | " +
              ", ".join(f"{n} {o}" for o, n in mix) + " |")
    print(f"\nblocks listed: {shown:.1f} % of all executed instructions\n")
    by_class = collections.Counter()
    by_op = collections.Counter()
    for ex, _, op in inst:
        by_class[klass(op)] += ex
        by_op[op] += ex
    print("| class | executed | share" + (" | per unit |" if units else " |"))
    print("|---|---|---|" + ("---|" if units else ""))
    for name, ex in by_class.most_common():
        print(f"| {name} | {ex:,} | {100.0 * ex / max(1, total):.1f} %" + (f" | {ex / units:.1f} |" if units else " |"))
    print("\ntop opcodes: " + ", ".join(f"{o} {100.0 * n / max(1, total):.1f}%" for o, n in by_op.most_common(20)))


if __name__ == "__main__":
    main()
