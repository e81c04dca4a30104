"""Summarise an `ncu --csv` launch list (one row per kernel x metric) into a per-kernel table.

    python tools/ncu_launch_summary.py gpurun_out/launches.csv [--skip N] [--count M]

Columns: launches, total us, us/launch, share of the summed time and, when the list carries it,
`sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active` time-weighted per kernel and over all conv kernels.
"""
import argparse
import collections
import csv
import re


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("rfv::", "").replace("(anonymous namespace)::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0, help="ignore the first N launches (set-up, warm-up)")
    ap.add_argument("--count", type=int, default=0)
    a = ap.parse_args()
    rows = collections.OrderedDict()
    with open(a.csv, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"]), "grid": r["Grid Size"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    ids = sorted(rows)[a.skip:]
    if a.count:
        ids = ids[:a.count]
    T, P = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    agg = collections.OrderedDict()
    for i in ids:
        d = rows[i]
        k = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "pw": 0.0})
        k["n"] += 1
        k["ns"] += d.get(T, 0.0)
        k["pw"] += d.get(P, 0.0) * d.get(T, 0.0)
    tot = sum(k["ns"] for k in agg.values())
    has_p = any(P in rows[i] for i in ids)
    print(f"{len(ids)} launches, {tot / 1e3:.0f} us summed")
    print("| kernel | launches | total us | us / launch | share |" + (" tensor pipe % (time-weighted) |" if has_p else ""))
    print("|---|---|---|---|---|" + ("---|" if has_p else ""))
    for name, k in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        line = f"| `{name}` | {k['n']} | {k['ns'] / 1e3:.0f} | {k['ns'] / 1e3 / k['n']:.1f} | {100 * k['ns'] / tot:.1f} % |"
        if has_p:
            line += f" {k['pw'] / k['ns']:.1f} |" if k["ns"] else " |"
        print(line)
    if has_p:
        conv = [k for n, k in agg.items() if n.startswith("conv_")]
        cns = sum(k["ns"] for k in conv)
        if cns:
            print(f"\nAll tcgen05 conv launches ({sum(k['n'] for k in conv)}): tensor pipe {sum(k['pw'] for k in conv) / cns:.1f} % "
                  f"time-weighted, {100 * cns / tot:.1f} % of the summed time")


if __name__ == "__main__":
    main()
