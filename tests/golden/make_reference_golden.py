"""Writes tests/golden/reference_kernels.json from the REFERENCE'S OWN kernels.

Runs on a GPU box (gpurun): oracle/_ref/libnmi_ref.so is orbslam2_NMI's NMI.cu + kernel.cu
compiled unmodified for sm_100a (oracle/Makefile.ref).  For every seeded pair of
tests/golden/reference_pairs.py it records what histogram256all, ComputeEntropyKernel,
AddvectorParwiseMidKernel and AddVectorPairwiseKernel produced (CRCs of the integer and float
arrays, the three totals and the score as exact bit patterns).  tests/test_reference_golden.py
then holds the CPU oracle to these numbers on any machine.
    gpurun -- 'python tests/golden/make_reference_golden.py'   # writes gpurun_out/reference_kernels_golden.json
"""
import json
import struct
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(Path(__file__).parent))

from oracle import ref_py  # noqa: E402
import reference_pairs  # noqa: E402


def bits(x):
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def main():
    out = {"_source": ref_py.describe(), "_generator": "tests/golden/make_reference_golden.py", "pairs": {}}
    for name, (render, warped) in reference_pairs.pairs().items():
        st = ref_py.stages(render, warped)
        sa, sb, sab = (np.float32(x) for x in st["sums"])
        if sa == 0 and sb == 0 and sab == 0:      # NMI.cu:352
            score = np.float32(0)
        else:                                     # NMI.cu:357, fp32, from the race-free totals
            score = np.float32(2) * (np.float32(1) - ((-sab) / ((-sa) + (-sb))))
        raw = [ref_py.score(render, warped) for _ in range(4)]
        out["pairs"][name] = dict(
            shape=list(render.shape),
            render_crc32=zlib.crc32(render.tobytes()), warped_crc32=zlib.crc32(warped.tobytes()),
            joint_crc32=zlib.crc32(st["J"].tobytes()), hist1_crc32=zlib.crc32(st["HA"].tobytes()),
            hist2_crc32=zlib.crc32(st["HB"].tobytes()),
            joint_nonzero=int((st["J"] != 0).sum()), joint_max=int(st["J"].max()),
            entropy1_crc32=zlib.crc32(st["ea"].tobytes()), entropy2_crc32=zlib.crc32(st["eb"].tobytes()),
            joint_entropy_crc32=zlib.crc32(st["ej"].tobytes()), row_sums_crc32=zlib.crc32(st["mid"].tobytes()),
            totals_bits=[bits(sa), bits(sb), bits(sab)], totals=[float(sa), float(sb), float(sab)],
            score_bits=bits(score), score=float(score),
            entry_point_outputs_bits=[bits(x) for x in raw],   # NMIWithCuda_noMask itself (race-prone)
        )
        print(name, float(score), raw)
    dst = ROOT / "gpurun_out" / "reference_kernels_golden.json"
    dst.parent.mkdir(exist_ok=True)
    dst.write_text(json.dumps(out, indent=1))
    print("wrote", dst)


if __name__ == "__main__":
    main()
