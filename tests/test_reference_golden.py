"""The CPU oracle against golden vectors produced by the REFERENCE'S OWN kernels.

tests/golden/reference_kernels.json was written on a B200 by tests/golden/make_reference_golden.py
from oracle/_ref/libnmi_ref.so -- orbslam2_NMI's NMI.cu + kernel.cu compiled unmodified for sm_100a --
for the seeded image pairs of tests/golden/reference_pairs.py: histogram256all's three histograms,
ComputeEntropyKernel's terms, AddvectorParwiseMidKernel's row sums, AddVectorPairwiseKernel's totals
and the score.  Here, on any CPU, the oracle has to reproduce all of it bit for bit (rows a7-a10 of
SURVEY 8a).  The GPU-side comparison of the CUDA path is tests/test_gpu_reference_kernels.py.
"""
import json
import struct
import sys
import zlib
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLDEN))
import reference_pairs  # noqa: E402

REF = json.loads((GOLDEN / "reference_kernels.json").read_text())


def f32(bits):
    return np.float32(struct.unpack("<f", struct.pack("<I", bits))[0])


@pytest.mark.parametrize("name", sorted(REF["pairs"]))
def test_oracle_reproduces_reference_kernel_outputs(oracle, name):
    want = REF["pairs"][name]
    render, warped = reference_pairs.pairs()[name]
    assert list(render.shape) == want["shape"]
    assert zlib.crc32(render.tobytes()) == want["render_crc32"] and zlib.crc32(warped.tobytes()) == want["warped_crc32"], \
        "the seeded inputs are not the ones the reference was run on"
    J, HA, HB = oracle.joint_hist(render, warped)
    assert zlib.crc32(J.tobytes()) == want["joint_crc32"]          # histogram256all + merges (NMI.cu:52-226)
    assert zlib.crc32(HA.tobytes()) == want["hist1_crc32"] and zlib.crc32(HB.tobytes()) == want["hist2_crc32"]
    assert int((J != 0).sum()) == want["joint_nonzero"] and int(J.max()) == want["joint_max"]
    st = oracle.score_stages_f32(J, HA, HB, render.size)
    assert zlib.crc32(st["ea"].tobytes()) == want["entropy1_crc32"]  # ComputeEntropyKernel (NMI.cu:230-267)
    assert zlib.crc32(st["eb"].tobytes()) == want["entropy2_crc32"]
    assert zlib.crc32(st["ej"].tobytes()) == want["joint_entropy_crc32"]
    assert zlib.crc32(st["mid"].tobytes()) == want["row_sums_crc32"]  # AddvectorParwiseMidKernel (NMI.cu:270-287)
    for got, bits in zip(st["sums"], want["totals_bits"]):            # AddVectorPairwiseKernel (NMI.cu:290-338)
        assert np.float32(got) == f32(bits)
    assert np.float32(st["score"]) == f32(want["score_bits"])         # NMI.cu:342-362
    assert np.float32(oracle.eval_one(render, warped)) == f32(want["score_bits"])
    # what CUDAF::NMIWithCuda_noMask itself returned on the box (its last kernel is race-prone; on these
    # runs it did not lose the race)
    assert all(f32(b) == f32(want["score_bits"]) for b in want["entry_point_outputs_bits"])


def test_golden_file_says_where_it_came_from():
    assert "NMI.cu" in REF["_source"] and "unmodified" in REF["_source"]
    assert len(REF["pairs"]) >= 9
