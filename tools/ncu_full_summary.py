"""Summarise an `ncu --set full` capture of one velocity evaluation (raw page exported as CSV) per kernel class.

    ncu -i forward_full.ncu-rep --page raw --csv > forward_full_raw.csv
    python tools/ncu_full_summary.py forward_full_raw.csv --micro-batch 512 [--json profiles/r2_ncu_traffic.json] > profiles/r2_ncu_forward_full.md

Per launch: duration, tensor-pipe activity, DRAM bytes read + written, issue-slot utilisation, L2 hit rate; per kernel class
(the engine's profile classes: conv_wa, conv_umma, gn_apply, ...) the time-weighted means and the DRAM bytes per image that
`bench.py` reports as `roofline.traffic`.
"""
import argparse
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLS = {
    "t": "gpu__time_duration.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "issue": "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "l2hit": "lts__t_sector_hit_rate.pct",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_clk": "sm__cycles_elapsed.avg.per_second",
}
UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6, "s": 1e6,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def klass(name: str) -> str:
    for k in ("input_conv", "output_conv", "conv_wa", "conv_umma", "conv_mma", "gn_apply", "gn_coef", "attn", "temb"):
        if k in name:
            return "attention" if k == "attn" else k
    return "other"


def short(name: str) -> str:
    return re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", name)).replace("rfv::", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--micro-batch", type=int, required=True)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    with open(a.csv, newline="") as f:
        rows = list(csv.reader(ln for ln in f if ln.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {c: hdr.index(c) for c in hdr}

    def val(r, key):
        c = COLS[key]
        if c not in ix:
            return None
        try:
            v = float(r[ix[c]].replace(",", ""))
        except ValueError:
            return None
        return v * UNIT.get(units[ix[c]], 1.0) if key in ("t", "rd", "wr") else v

    per = collections.OrderedDict()
    print(f"# ncu --set full, one velocity evaluation at micro-batch {a.micro_batch} ({len(data)} launches)\n")
    print("| # | kernel | grid | us | tensor pipe % | DRAM MB (r+w) | DRAM % of peak | issue slots % | L2 hit % |")
    print("|---|---|---|---|---|---|---|---|---|")
    for i, r in enumerate(data):
        name = short(r[ix["Kernel Name"]])
        t, tp, rd, wr = val(r, "t"), val(r, "tensor") or 0.0, val(r, "rd") or 0.0, val(r, "wr") or 0.0
        k = per.setdefault(klass(name), {"n": 0, "us": 0.0, "tp": 0.0, "bytes": 0.0, "issue": 0.0})
        k["n"] += 1; k["us"] += t; k["tp"] += tp * t; k["bytes"] += rd + wr; k["issue"] += (val(r, "issue") or 0.0) * t
        fmt = lambda v, p=1: "" if v is None else f"{v:.{p}f}"
        print(f"| {i} | `{name}` | {r[ix['Grid Size']]} | {t:.1f} | {tp:.1f} | {(rd + wr) / 1e6:.1f} | {fmt(val(r, 'dram_pct'))} | "
              f"{fmt(val(r, 'issue'))} | {fmt(val(r, 'l2hit'))} |")
    tot = sum(k["us"] for k in per.values())
    print("\n| class | launches | us | share | tensor pipe % (time-weighted) | issue slots % | DRAM MB (r+w) | DRAM bytes / image |")
    print("|---|---|---|---|---|---|---|---|")
    for name, k in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| {name} | {k['n']} | {k['us']:.0f} | {100 * k['us'] / tot:.1f} % | {k['tp'] / k['us']:.1f} | {k['issue'] / k['us']:.1f} | "
              f"{k['bytes'] / 1e6:.0f} | {k['bytes'] / a.micro_batch:.0f} |")
    conv = [k for n, k in per.items() if n.startswith("conv_")]
    cus = sum(k["us"] for k in conv)
    if cus:
        print(f"\nAll tcgen05 conv launches ({sum(k['n'] for k in conv)}): `sm__pipe_tensor_cycles_active` "
              f"{sum(k['tp'] for k in conv) / cus:.1f} % time-weighted.")
    if a.json:
        from rectified_flow_vision_b200 import _build
        out = {"micro_batch": a.micro_batch, "library_digest": _build._digest(), "source": os.path.basename(a.csv),
               "dram_bytes_per_image": {n: k["bytes"] / a.micro_batch for n, k in per.items()},
               "tensor_pipe_pct_time_weighted": {n: k["tp"] / k["us"] for n, k in per.items() if k["us"]},
               "us_per_forward_under_ncu": {n: k["us"] for n, k in per.items()}}
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
