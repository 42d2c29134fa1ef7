"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_summary.txt "<command>"
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    n = 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[r[ui]]
        name = re.sub(r"^void |nmi::<unnamed>::|nmi::\(anonymous namespace\)::", "", r[ki].split("(")[0])[:48]
        c, t = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, t + us)
        n += 1
    total = sum(t for _, t in tot.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : {cmd}",
           f"# (cold-cache, serialised replays: compare SHARES, not absolutes). {n} launches captured", "",
           f"{'kernel':48s} {'launches':>8s} {'total us':>12s} {'share':>7s}"]
    for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{name:48s} {c:8d} {t:12.1f} {100 * t / total:6.1f}%")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
