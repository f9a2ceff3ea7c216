#!/usr/bin/env python3
"""Development helper: hottest SASS instructions of a kernel from an ncu report (stall samples + executed counts).

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_hot.py sass.csv [top_n]
Prints total samples, the top-N instructions by stall samples with their dominant stall reasons, and executed-instruction totals."""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    data = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            samples = int(r[col["# Samples"]] or 0)
            ex = int(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        stalls = {n: int(r[col[n]] or 0) for n in stall_cols}
        data.append((samples, ex, r[col["Source"]].strip(), stalls, len(data)))
    total = sum(d[0] for d in data)
    total_ex = sum(d[1] for d in data)
    print(f"total samples {total}, executed warp instructions {total_ex}, static instructions {len(data)}")
    agg = {}
    for d in data:
        for n, v in d[3].items():
            agg[n] = agg.get(n, 0) + v
    print("stall mix:", ", ".join(f"{n[6:]} {100 * v / max(1, sum(agg.values())):.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print("\nrank  idx  samples%  executed  instruction  | top stalls")
    for rank, d in enumerate(sorted(data, key=lambda d: -d[0])[:top]):
        st = sorted(d[3].items(), key=lambda kv: -kv[1])[:2]
        print(f"{rank:3d} {d[4]:5d} {100 * d[0] / max(1, total):6.2f}%  {d[1]:8d}  {d[2][:70]:70s} | " + ", ".join(f"{n[6:]}:{v}" for n, v in st if v))


if __name__ == "__main__":
    main()
