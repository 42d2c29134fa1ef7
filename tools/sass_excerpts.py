"""profiles/r02_sass_excerpts.txt: per kernel, the SASS mnemonics that show which hardware paths it uses.

    python tools/sass_excerpts.py [lib.so] [out.txt]
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "orbslam2_nmi_b200" / "_lib" / "libnmi_b200.so")
dst = sys.argv[2] if len(sys.argv) > 2 else str(ROOT / "profiles" / "r02_sass_excerpts.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
out = ["# cuobjdump -sass orbslam2_nmi_b200/_lib/libnmi_b200.so (sm_100a) -- the instructions that show which hardware paths the kernels use.",
       "# Per kernel: count of each mnemonic of interest, then the first occurrence of each with its address.",
       "# UBLKCP = cp.async.bulk (TMA bulk copy global -> shared), SYNCS.* = mbarrier, ATOMS = shared-memory atomics,",
       "# UBLKRED = cp.reduce.async.bulk (bulk reduce-add into another CTA's shared memory over the cluster network),",
       "# UCGABAR = barrier.cluster, UTCCP / LDTM = tcgen05.cp / tcgen05.ld (tensor-memory staging, variant 10),",
       "# TLD4 = texture gather (warp kernel), MATCH = match.any, FENCE.VIEW.ASYNC = fence.proxy.async.", ""]
pat = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P[0-9T]\s+)?)([A-Z0-9_.]+)(.*?);")
want = re.compile(r"^(UBLKCP|UBLKRED|SYNCS|ATOMS|UTCCP|LDTM|UTCBAR|UCGABAR|TLD4|MATCH|FENCE\.VIEW|CCTL\.IVALL|MAPA)")
cur, counts, firsts, order = None, {}, {}, []
for ln in sass:
    if "Function :" in ln:
        cur = ln.split("Function :")[1].strip()
        counts[cur], firsts[cur] = {}, {}
        order.append(cur)
        continue
    m = pat.match(ln)
    if m and cur:
        mn = m.group(3)
        if want.match(mn):
            counts[cur][mn] = counts[cur].get(mn, 0) + 1
            firsts[cur].setdefault(mn, f"/*{m.group(1)}*/ {m.group(2)}{mn}{m.group(4)} ;")


def short(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"nmi::\(anonymous namespace\)::", "", r).split("(")[0][:110]


for k in order:
    if not counts[k] or not any(x in k for x in ("hist", "warp_kernel", "bin_kernel", "tile_resolve", "mesh_raster", "image_mode")):
        continue
    out.append(short(k))
    out.append("  counts: " + ", ".join(f"{a} x{b}" for a, b in sorted(counts[k].items(), key=lambda t: -t[1])))
    out += ["    " + b for b in firsts[k].values()] + [""]
Path(dst).write_text("\n".join(out))
print(dst, len(out), "lines")
