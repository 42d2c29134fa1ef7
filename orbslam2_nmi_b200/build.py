"""In-tree build of the sm_100a library and (for tests only) the CPU oracle.

`python -m orbslam2_nmi_b200.build` or `__graft_entry__.build()`.
Outputs (git-ignored, but shipped to the GPU box by gpurun):
  orbslam2_nmi_b200/_lib/libnmi_b200.so     the product (C ABI, include/nmi_b200.h)
  oracle/_build/libnmi_oracle.so            test infrastructure (never loaded by the product)
  oracle/_ref/libnmi_ref.so                 the reference's own NMI.cu + kernel.cu for sm_100a (checker;
                                            built only where /root/reference exists)
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "orbslam2_nmi_b200" / "csrc"
LIBDIR = ROOT / "orbslam2_nmi_b200" / "_lib"
LIB = LIBDIR / "libnmi_b200.so"
ORACLE_LIB = ROOT / "oracle" / "_build" / "libnmi_oracle.so"

CU_SOURCES = ["capi.cu", "project.cu", "mesh.cu", "warp.cu", "hist.cu", "argmax.cu", "host_math.cpp", "driver.cpp", "compat.cpp", "compat_nmi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--compiler-options", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-fopenmp",
    "-lgomp",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _host_cxx() -> str:
    # the image exports CC/CXX=/opt/gcc/bin/*, a wrapper without libgomp; use the system compiler
    return "/usr/bin/g++" if Path("/usr/bin/g++").exists() else (shutil.which("g++") or "g++")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in CU_SOURCES]
    deps = srcs + [CSRC / "nmi_internal.h", ROOT / "include" / "nmi_b200.h"] + sorted((ROOT / "include" / "compat").glob("*"))
    stamp = LIBDIR / "build.sha256"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cmd = [_nvcc(), "-ccbin", _host_cxx(), *NVCC_FLAGS, "-shared", "-o", str(LIB), *map(str, srcs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (LIBDIR / "build.log").write_text(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see orbslam2_nmi_b200/_lib/build.log")
    stamp.write_text(dig)
    return LIB


def build_oracle(force: bool = False) -> Path:
    src = ROOT / "oracle" / "nmi_oracle.c"
    hdr = ROOT / "oracle" / "nmi_oracle.h"
    if not force and ORACLE_LIB.exists() and ORACLE_LIB.stat().st_mtime >= max(
        src.stat().st_mtime, hdr.stat().st_mtime
    ):
        return ORACLE_LIB
    res = subprocess.run(["make", "-C", str(ROOT / "oracle"), "-B"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("oracle build failed")
    return ORACLE_LIB


UBENCH = LIBDIR / "ubench_atoms"


def build_ubench(force: bool = False) -> Path:
    """tools/ubench_atoms.cu -> _lib/ubench_atoms: the shared-memory atomic microbenchmark bench.py runs
    for the peak of its smem_atomic roofline (a measurement tool, not part of the search path)."""
    src = ROOT / "tools" / "ubench_atoms.cu"
    if not force and UBENCH.exists() and UBENCH.stat().st_mtime >= src.stat().st_mtime:
        return UBENCH
    LIBDIR.mkdir(parents=True, exist_ok=True)
    res = subprocess.run([_nvcc(), "-ccbin", _host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                          "-lineinfo", "-o", str(UBENCH), str(src)], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("ubench_atoms build failed")
    return UBENCH


NPP_LIB = ROOT / "oracle" / "_build" / "libnmi_nppcheck.so"


def build_nppcheck(force: bool = False):
    """NPP's nppiWarpPerspective_8u_C1R behind a C ABI (oracle/npp_check.cu): a checker for the warp
    stage in the GPU tests, never loaded by the product.  None when libnppig is not installed."""
    src = ROOT / "oracle" / "npp_check.cu"
    if not force and NPP_LIB.exists() and NPP_LIB.stat().st_mtime >= src.stat().st_mtime:
        return NPP_LIB
    if not Path("/usr/local/cuda/include/nppi_geometry_transforms.h").exists():
        return NPP_LIB if NPP_LIB.exists() else None
    res = subprocess.run(["make", "-C", str(ROOT / "oracle"), "npp", f"NVCC={_nvcc()}"] + (["-B"] if force else []),
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("NPP checker build failed")
    return NPP_LIB


REF_LIB = ROOT / "oracle" / "_ref" / "libnmi_ref.so"
REF_CUF = Path("/root/reference/Thirdparty/CUDA_Functions")


def build_reference(force: bool = False):
    """The reference's own NMI.cu + kernel.cu, compiled unmodified from /root/reference for sm_100a
    (oracle/Makefile.ref) -- a checker for the GPU tests, never loaded by the product.  Only
    possible where /root/reference exists (the build container); the GPU box uses the prebuilt
    oracle/_ref/libnmi_ref.so.  Returns the path, or None when it can neither be built nor found."""
    if not (REF_CUF / "NMI.cu").exists():
        return REF_LIB if REF_LIB.exists() else None
    args = ["make", "-C", str(ROOT / "oracle"), "-f", "Makefile.ref", f"NVCC={_nvcc()}"]
    if force:
        args.append("-B")
    res = subprocess.run(args, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("reference (oracle/_ref) build failed")
    return REF_LIB


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose=True))
    print(build_oracle(force="--force" in sys.argv))
    print(build_reference(force="--force" in sys.argv))
    print(build_nppcheck(force="--force" in sys.argv))
    print(build_ubench(force="--force" in sys.argv))
