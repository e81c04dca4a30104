"""SASS evidence: which kernels of librfv_b200.so carry tcgen05 / TMEM / TMA instructions.

    python tools/sass_summary.py > profiles/r2_sass_summary.md

Disassembles the shipped library with `cuobjdump -sass` and counts, per kernel, the mnemonics that prove the Blackwell path
(/opt/skills/guides/B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF (TMA
tensor loads / stores / prefetch), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rectified_flow_vision_b200", "librfv_b200.so")
MNEMONICS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.TANH"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*$", "", name).replace("rfv::", "").replace("void ", "")
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    cur[mn] += 1
    arch = re.search(r"arch = (sm_\w+)", out)
    print(f"# SASS summary of `rectified_flow_vision_b200/librfv_b200.so` ({arch.group(1) if arch else '?'}; `cuobjdump -sass`, {len(counts)} kernels)\n")
    print("| kernel | SASS instructions | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    tot = collections.Counter()
    for name, c in sorted(counts.items(), key=lambda kv: (-kv[1]["UTCHMMA"], -kv[1]["HMMA"], kv[0])):
        if not any(c[m] for m in MNEMONICS if m not in ("SYNCS", "MUFU.TANH")):
            continue
        print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
        tot.update(c)
    print(f"| **all kernels listed** | {tot['_total']} | " + " | ".join(str(tot[m]) for m in MNEMONICS) + " |")
    plain = [n for n, c in counts.items() if not any(c[m] for m in MNEMONICS if m not in ("SYNCS", "MUFU.TANH"))]
    print(f"\n{len(plain)} further kernels carry none of these (elementwise / reduction / packing kernels): "
          + ", ".join(f"`{n}`" for n in sorted(plain)))


if __name__ == "__main__":
    sys.exit(main())
