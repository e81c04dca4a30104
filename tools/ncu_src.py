"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --launch-skip K --launch-count 1 --print-source sass,cuda > src.csv
    python tools/ncu_src.py src.csv [top]
"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            sections.append(cur)
        elif r and r[0] == "Line No" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for sec in sections:
        h = sec["hdr"]
        iex, ismp = h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(num(r[iex]) for r in sec["rows"])
        tots = sum(num(r[ismp]) for r in sec["rows"])
        print(f"{sec['file'].split('/')[-1]}: instructions {tot}, samples {tots}, lines {len(sec['rows'])}")
        for r in sorted(sec["rows"], key=lambda r: -num(r[iex]))[:top]:
            print(f"  L{r[0]:>4} ex={num(r[iex]):>9} smp={num(r[ismp]):>6} | {r[1][:120]}")


if __name__ == "__main__":
    main()
