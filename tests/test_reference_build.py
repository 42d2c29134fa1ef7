"""CPU-side checks of oracle/_ref: the reference's own NMI.cu + kernel.cu compile unmodified
from /root/reference (oracle/Makefile.ref) and the library exports the reference's entry points
next to the harness's.  No compute here (the reference has no CPU path); the comparison itself
is tests/test_gpu_reference_kernels.py on the GPU box."""
import ctypes
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference/Thirdparty/CUDA_Functions")

REFERENCE_SYMBOLS = [
    "_ZN5CUDAF18NMIWithCuda_noMaskEPN2cv4cuda7PtrStepIhEEiiiiPfj",  # kernel.cuh:35-38
    "initHistogram256all", "closeHistogram256all", "histogram256all",   # NMI.cuh:60-71
    "_Z20ComputeEntropyKernelPjS_S_iPfS0_S0_",                         # NMI.cuh:74
    "_Z25AddvectorParwiseMidKernelPfS_", "_Z23AddVectorPairwiseKernelPfS_S_",  # NMI.cuh:76-78
]
HARNESS_SYMBOLS = ["nmiref_score", "nmiref_stages", "nmiref_time", "nmiref_describe", "refshim_register_gl_texture"]


@pytest.fixture(scope="module")
def ref_lib():
    from orbslam2_nmi_b200 import build

    lib = build.build_reference()
    if lib is None:
        pytest.skip("no /root/reference here and no prebuilt oracle/_ref/libnmi_ref.so")
    return lib


def test_reference_sources_compile_and_export(ref_lib):
    lib = ctypes.CDLL(str(ref_lib))
    for s in REFERENCE_SYMBOLS + HARNESS_SYMBOLS:
        assert getattr(lib, s) is not None, s


def test_reference_kernels_are_in_the_fatbin(ref_lib):
    out = subprocess.run(["cuobjdump", "-elf", str(ref_lib)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    for k in ("histogram256Kernel", "mergeHistogram256Kernel", "mergeJointHistogram256Kernel",
              "ComputeEntropyKernel", "AddvectorParwiseMidKernel", "AddVectorPairwiseKernel"):
        assert k in out.stdout, f"{k} missing from the sm_100a image"
    assert "sm_100a" in out.stdout or "sm_100" in out.stdout


def test_nothing_from_the_reference_is_in_the_repo():
    """The recipe compiles the sources where they lie; only build outputs land in oracle/_ref/ (git-ignored)."""
    res = subprocess.run(["git", "ls-files"], cwd=ROOT, capture_output=True, text=True)
    if res.returncode == 0:   # a snapshot without .git (the GPU box) has nothing to list
        assert not [f for f in res.stdout.split() if f.startswith("oracle/_ref/")]
    mk = (ROOT / "oracle" / "Makefile.ref").read_text()
    assert "$(CUF)/NMI.cu" in mk and "$(CUF)/kernel.cu" in mk
    assert "oracle/_ref/" in (ROOT / ".gitignore").read_text()
    gi = ROOT / ".gpurunignore"
    assert not gi.exists() or "oracle/_ref" not in gi.read_text()


def test_product_does_not_touch_the_reference_library():
    for f in (ROOT / "orbslam2_nmi_b200").rglob("*"):
        if f.is_file() and f.suffix in (".py", ".cu", ".cpp", ".h") and f.name != "build.py":
            assert "libnmi_ref" not in f.read_text(errors="ignore") and "ref_py" not in f.read_text(errors="ignore"), f
