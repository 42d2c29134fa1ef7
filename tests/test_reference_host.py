"""CPU tests against the reference's OWN host-side search bookkeeping.

oracle/_ref/libnmi_ref_host.so is Thirdparty/Localization/nmiSearchKernel.cpp + helperFunctions.cpp
compiled unmodified with g++ from /root/reference (oracle/Makefile.ref, oracle/ref_host_harness.cpp).
It pins, on the reference itself:
  a12  helperFunctions::find_max_elements + "element [0]" (src/Tracking.cc:1952): max from 0, strict >,
       ties -> lowest index in wz,wy,wx,sz,sy,sx order, all-negative -> empty vector
  a15  NmiSearchKernel::isMiddle / resizeKernel (periphery rule, 0.5 factor, 0.005 m / 0.001 rad floors)
  f4   operator<<(ostream&, NmiSearchKernel): the `_log.txt` line
for the CPU oracle, for the product's host code (nmi_grid_is_middle / nmi_grid_resize /
nmi_decode_key, no GPU needed) and for the compat C++ operator<<.
"""
import ctypes as C
import struct
import subprocess
from pathlib import Path

import numpy as np
import pytest

from orbslam2_nmi_b200.capi import Grid

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def refh():
    from oracle import ref_py
    from orbslam2_nmi_b200 import build

    build.build_reference()
    if not ref_py.host_available():
        pytest.skip("no /root/reference here and no prebuilt oracle/_ref/libnmi_ref_host.so")
    ref_py.load_host()
    return ref_py


def _grids(rng, n):
    for _ in range(n):
        nS = tuple(int(x) for x in rng.integers(1, 5, 3))
        nW = tuple(int(x) for x in rng.integers(1, 5, 3))
        yield nS, nW


def _key(score, index):
    """the packed winner key of the argmax kernel."""
    bits = struct.unpack("<I", struct.pack("<f", max(float(score), 0.0)))[0]
    return (bits << 32) | (0xFFFFFFFF - index)


def test_argmax_rule_matches_find_max_elements(refh, oracle, nmi_lib):
    from orbslam2_nmi_b200 import search

    rng = np.random.default_rng(7)
    cases = 0
    for nS, nW in _grids(rng, 60):
        n = int(np.prod(nS) * np.prod(nW))
        g = Grid.make(nS, nW, (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
        for kind in ("random", "ties", "zeros_and_negatives", "all_zero", "all_negative", "quantised"):
            if kind == "random":
                r = rng.random(n).astype(np.float32)
            elif kind == "ties":
                r = rng.random(n).astype(np.float32)
                r[rng.integers(0, n, max(1, n // 3))] = r.max()
            elif kind == "zeros_and_negatives":
                r = -rng.random(n).astype(np.float32)
                r[rng.integers(0, n, max(1, n // 4))] = 0.0
            elif kind == "all_zero":
                r = np.zeros(n, np.float32)
            elif kind == "all_negative":
                r = -rng.random(n).astype(np.float32) - 0.01
            else:
                r = (rng.integers(0, 4, n) / 4).astype(np.float32)
            count, bs, bw, sc = refh.find_max(r, nS, nW)
            want, wmax = oracle.argmax(r)
            if count == 0:                      # reference: empty vector, [0] would be UB
                assert want == -1
                continue
            assert count == int((r == r[want]).sum())
            assert oracle.unravel(g, want) == (bs, bw)
            assert np.float32(wmax) == np.float32(sc) == r[want]
            # product host code: the key the argmax kernel publishes for that winner decodes to it
            dec = search.decode_key(g, _key(r[want], want))
            assert (dec.best_s, dec.best_w) == (bs, bw) and dec.best_index == want
            assert np.float32(dec.best_score) == np.float32(sc)
            cases += 1
    assert cases > 200


def test_resize_and_is_middle_match_reference(refh, oracle, nmi_lib):
    from orbslam2_nmi_b200 import search

    rng = np.random.default_rng(11)
    n_mid = 0
    for i in range(3000):
        nS = tuple(int(x) for x in rng.integers(1, 6, 3))
        nW = tuple(int(x) for x in rng.integers(1, 6, 3))
        # steps around the 0.005 m / 0.001 rad floors as well as ordinary ones
        stepT = tuple(float(x) for x in np.float32(rng.choice([0.2, 0.5, 0.02, 0.011, 0.01, 0.0099, 0.006, 0.004], 3)))
        stepR = tuple(float(x) for x in np.float32(rng.choice([0.05, 0.02, 0.004, 0.0021, 0.002, 0.0019, 0.0012, 0.0008], 3)))
        if i % 3 == 0:   # the middle cell, or near it
            bs = tuple(n // 2 for n in nS); bw = tuple(n // 2 for n in nW)
        else:
            bs = tuple(int(rng.integers(0, n)) for n in nS); bw = tuple(int(rng.integers(0, n)) for n in nW)
        mid, rS, rW, rT, rR = refh.resize(nS, nW, stepT, stepR, bs, bw)
        g = Grid.make(nS, nW, stepT, stepR)
        n_mid += mid
        # oracle
        assert oracle.is_middle(g, bs, bw) == mid
        og = oracle.resize_grid(g, bs, bw)
        assert tuple(og.nS) == rS and tuple(og.nW) == rW
        assert np.array_equal(np.array(list(og.stepT), np.float32), rT)
        assert np.array_equal(np.array(list(og.stepR), np.float32), rR)
        # product host code
        assert search.grid_is_middle(g, bs, bw) == mid
        pg = search.grid_resize(g, bs, bw)
        assert tuple(pg.nS) == rS and tuple(pg.nW) == rW
        assert np.array_equal(np.array(list(pg.stepT), np.float32), rT)
        assert np.array_equal(np.array(list(pg.stepR), np.float32), rR)
    assert n_mid > 100


def test_log_line_format_matches_reference(refh, tmp_path):
    """compat's operator<<(NmiSearchKernel) prints the reference's line byte for byte."""
    src = tmp_path / "fmt.cpp"
    src.write_text(r'''
#include "nmiSearchKernel.hpp"
#include <cstdio>
#include <cstdlib>
#include <sstream>
int main(int argc, char** argv) {
  if (argc != 20) return 2;
  int v[12]; float f[7];
  for (int i = 0; i < 6; i++) v[i] = atoi(argv[1 + i]);
  for (int i = 0; i < 6; i++) f[i] = strtof(argv[7 + i], nullptr);
  for (int i = 0; i < 6; i++) v[6 + i] = atoi(argv[13 + i]);
  f[6] = strtof(argv[19], nullptr);
  NmiSearchKernel k(v[0], v[1], v[2], v[3], v[4], v[5], f[0], f[1], f[2], f[3], f[4], f[5]);
  k.setBest(v[6], v[7], v[8], v[9], v[10], v[11], f[6]);
  std::stringstream ss; ss << k;
  fputs(ss.str().c_str(), stdout);
  return 0;
}
''')
    exe = tmp_path / "fmt"
    lib = ROOT / "orbslam2_nmi_b200" / "_lib"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    res = subprocess.run([cxx, "-std=c++17", "-O1", f"-I{ROOT / 'include' / 'compat'}", f"-I{ROOT / 'include'}",
                          str(src), f"-L{lib}", "-lnmi_b200", f"-Wl,-rpath,{lib}", "-o", str(exe)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    rng = np.random.default_rng(3)
    for _ in range(40):
        nS = [int(x) for x in rng.integers(1, 12, 3)]; nW = [int(x) for x in rng.integers(1, 12, 3)]
        stepT = [float(np.float32(x)) for x in rng.random(3) * rng.choice([1, 0.01, 30])]
        stepR = [float(np.float32(x)) for x in rng.random(3) * 0.1]
        bs = [int(rng.integers(-1, n)) for n in nS]; bw = [int(rng.integers(-1, n)) for n in nW]
        nmi = float(np.float32(rng.random() * rng.choice([1, 1e-3, 2])))
        want = refh.format_kernel(nS, nW, stepT, stepR, bs, bw, nmi)
        args = [str(x) for x in nS + nW] + [repr(x) for x in stepT + stepR] + [str(x) for x in bs + bw] + [repr(nmi)]
        got = subprocess.run([str(exe)] + args, capture_output=True, text=True)
        assert got.returncode == 0
        assert got.stdout == want
