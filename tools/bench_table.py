#!/usr/bin/env python3
"""Markdown tables from a bench.py JSON line (DESIGN.md §6 is generated with this):  python tools/bench_table.py gpurun_out/x.json"""
import json
import sys


def main():
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    rows = [(d["metric"].split("(")[1].rstrip(")"), d)] + list(d.get("workloads", {}).items())
    print("| workload | resident frames/s | e2e frames/s (8 host threads) | step ms | bin_coarse / bin_fine / composite ms | composite, canvas-writing flush ms | HBM frac (present-only / canvas-writing) | FP64 frac | list entries (interior) | parity |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for name, r in rows:
        k = r["kernel_ms"]
        det = r.get("details", r)
        par = d["parity"]["by_leg"].get(name, {})
        ok = all(v.get("match", True) for v in par.values() if isinstance(v, dict) and v.get("checked"))
        n_checked = sum(1 for v in par.values() if isinstance(v, dict) and v.get("checked"))
        print(f"| {name} | {r['value']:,.0f} | {r['e2e']['value']:,.0f} | {k['step']:.3f} | {k['ncr_bin_coarse']:.3f} / {k['ncr_bin_fine']:.3f} / {k['ncr_composite']:.3f} | "
              f"{r['roofline_canvas_flush']['kernel_ms']:.3f} | {r['roofline']['frac']:.3f} / {r['roofline_canvas_flush']['frac']:.3f} | "
              f"{r['roofline_fp64']['frac']:.3f} | {det['region_list_entries']:,} ({det.get('interior_entries', 0):,}) | {'match' if ok else 'MISMATCH'} ({n_checked} outputs) |")
    if "video" in d:
        print()
        print("| video leg | frames | frames/s | D2H GB/s | of the measured D2H ceiling | parity (frames checked) |")
        print("|---|---|---|---|---|---|")
        for name, v in d["video"].items():
            print(f"| {name} | {v['frames']:,} | {v['value']:,.0f} | {v['d2h_gb_per_s']:.1f} | {v.get('fraction_of_d2h_ceiling') or 0:.2f} | "
                  f"{'match' if v['parity']['match'] else 'MISMATCH'} ({v['parity']['frames_checked']}) |")
    print()
    print("d2h_ceiling", d.get("d2h_ceiling"))
    print("e2e_python", d.get("e2e_python"))
    print("cpu_baseline", d.get("cpu_baseline"))
    print("parity", {k: v for k, v in d["parity"].items() if k != "by_leg"})


if __name__ == "__main__":
    main()
