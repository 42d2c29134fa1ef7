"""Build libnmi_b200 with extra -D flags into orbslam2_nmi_b200/_lib/variants/<name>.so (A/B experiments on the
GPU box: NMI_B200_LIB=<path> python bench.py ...).   python tools/build_variant.py NAME -DFOO=1 -DBAR=2"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from orbslam2_nmi_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = b.LIBDIR / "variants"
out.mkdir(parents=True, exist_ok=True)
lib = out / f"{name}.so"
srcs = [b.CSRC / s for s in b.CU_SOURCES]
cmd = [b._nvcc(), "-ccbin", b._host_cxx(), *[f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-shared", "-o", str(lib),
       *map(str, srcs)]
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    sys.exit(1)
print(lib)
