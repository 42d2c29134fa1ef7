"""C++ host + NCCL: tests/cpp/test_nccl_host.cpp (one process, one thread and one nmi_ctx per GPU,
ncclCommInitAll, nmi_relocalize_sharded with ncclAllReduce(ncclUint64, ncclMax) as the exchange
callback of INTEGRATION.md section 5).  Every rank must return the single-GPU driver's result.

CPU suite: the program compiles and links against libnmi_b200.so + libnccl (no GPU call).
GPU suite: it runs on every visible GPU (world = 1 on a one-GPU box still goes through NCCL;
the multi-rank run needs gpurun --gpus N)."""
import shutil
import subprocess
from pathlib import Path

import pytest

from orbslam2_nmi_b200 import build

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    if not Path("/usr/include/nccl.h").exists():
        pytest.skip("nccl.h not installed")
    lib = build.build_cuda()
    out = tmp_path_factory.mktemp("bin") / "test_nccl_host"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else (shutil.which("g++") or "g++")
    cmd = [cxx, "-std=c++17", "-O1", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           str(ROOT / "tests" / "cpp" / "test_nccl_host.cpp"), "-L", str(lib.parent), "-lnmi_b200",
           f"-Wl,-rpath,{lib.parent}", "-L", "/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lpthread", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_nccl_host_builds(exe):
    assert exe.exists()


@pytest.mark.gpu
def test_nccl_host_all_ranks_agree_with_one_gpu(exe):
    res = subprocess.run([str(exe), "0"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "NCCL HOST OK" in res.stdout, res.stdout + res.stderr
