"""The reference-named C++ drop-ins (include/compat/) compiled and driven like Tracking.cc.

CPU part: host logic (NmiSearchKernel, find_max_elements, setupCam, YAML, loadXYZ).
GPU part: the reference's own call sequence (renderToTextureOnGPU -> NMIWithCuda_noMask per
pair -> find_max_elements) and the batched NmiObjects::searchGrid, both against the oracle.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from orbslam2_nmi_b200 import build, synth
from orbslam2_nmi_b200.capi import Grid

ROOT = Path(__file__).resolve().parent.parent

YAML = """%YAML:1.0
# reference-style settings (Examples/Monocular/ETH_small.yaml:62-96 keys)
Camera.fx: {fx}
Camera.fy: {fy}
Camera.cx: {cx}
Camera.cy: {cy}
Camera.Width: {W}
Camera.Height: {H}
NMI.Init1: !!opencv-matrix
    rows: 4
    cols: 4
    dt: f
    data: [1.0, 0.0, 0.0, 0.5, 0.0, 1.0, 0.0, 0.25,
           0.0, 0.0, 1.0, 2.0, 0.0000, 0.0000, 0.0000, 1.0000]
NMI.Offset: 10
NMI.Treshold: 0.05
NMI.SynthNumX:2
NMI.SynthNumY:1
NMI.SynthNumZ:2
NMI.WarpNumX: 1
NMI.WarpNumY: 2
NMI.WarpNumZ: 1
NMI.SynthStepX: 0.2
NMI.SynthStepY: 0.2
NMI.SynthStepZ: 0.5
NMI.WarpStepX: 0.02
NMI.WarpStepY: 0.03
NMI.WarpStepZ: 0.05
NMI.Render.PointSize: 3.0
NMI.Render.NearPlane: 5.0
NMI.Render.FarPlane: 30.0
NMI.Render.Object: "unused.obj"
NMI.Render.Texture: "unused.bmp"
NMI.Render.Cloud: "{cloud}"
NMI.Render.Offset: "{offset}"
"""
OFFSET = (1000.0, -2000.0, 50.0)


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("compat")
    sc = synth.make_scene("tiny", n_points=6000)
    # the .xyz format stores integer colours and absolute coordinates (objloader.cpp:253-261)
    rgb = np.rint(sc.xyzi[:, 3] * 256.0)
    lines = ["%.6f %.6f %.6f %d %d %d" % (x + OFFSET[0], y + OFFSET[1], z + OFFSET[2], c, c // 2, 0)
             for (x, y, z), c in zip(sc.xyzi[:, :3].astype(np.float64), rgb.astype(int))]
    (d / "cloud.xyz").write_text("\n".join(lines) + "\n")  # trailing newline: last point duplicates
    (d / "offset.xyz").write_text("%.1f %.1f %.1f\n" % OFFSET)
    (d / "settings.yaml").write_text(YAML.format(fx=sc.fx, fy=sc.fy, cx=sc.cx, cy=sc.cy, W=sc.W, H=sc.H,
                                                 cloud=d / "cloud.xyz", offset=d / "offset.xyz"))
    frame = synth.frame_textured(sc.W, sc.H, seed=21)
    frame.tofile(d / "frame.raw")
    np.savetxt(d / "twc.txt", sc.Twc.reshape(1, 16), fmt="%.9g")
    return d, sc, frame


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    lib = build.build_cuda()
    out = tmp_path_factory.mktemp("bin") / "test_compat"
    cmd = ["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-std=c++17", "-O1",
           "-ffp-contract=off", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_compat.cpp"),
           "-L", str(lib.parent), "-lnmi_b200", f"-Wl,-rpath,{lib.parent}", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def test_compat_host_logic(exe, workdir):
    d, _, _ = workdir
    res = subprocess.run([str(exe), "host", str(d / "settings.yaml"), str(d / "cloud.xyz"),
                          str(d / "offset.xyz")], capture_output=True, text=True)
    assert res.returncode == 0 and "HOST OK" in res.stdout, res.stdout + res.stderr


def test_compat_obj_bmp_loaders(exe, tmp_path):
    """loadOBJ / loadBMP_custom formats (objloader.cpp:140-223, texture.cpp:31-107)."""
    import struct

    obj = tmp_path / "m.obj"
    obj.write_text("# comment\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\n"
                   "vt 0.25 0.25\nvt 0.75 0.25\nvt 0.25 0.75\nvt 0.75 0.75\n"
                   "f 1/1 2/2 3/3\nf 2/2 4/4 3/3\n")
    px = bytes([10, 20, 30, 40, 50, 60, 0, 0,      # bottom row (2 texels + 2 pad bytes)
                70, 80, 90, 100, 110, 120, 0, 0])  # top row
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(px), 0, 0, 54) + struct.pack(
        "<IiiHHIIiiII", 40, 2, 2, 1, 24, 0, 0, 2835, 2835, 0, 0)
    bmp = tmp_path / "t.bmp"
    # the reference reads tightly packed rows (GL_UNPACK_ALIGNMENT 1): write an unpadded image
    bmp.write_bytes(hdr[:34] + struct.pack("<I", 12) + hdr[38:] + bytes([10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 110, 120]))
    res = subprocess.run([str(exe), "loaders", str(obj), str(bmp)], capture_output=True, text=True)
    assert res.returncode == 0 and "LOADERS OK" in res.stdout, res.stdout + res.stderr


def test_compat_output_files(exe, tmp_path):
    """FrameTrajectory*.txt tags / layout (System.cc:514-599) and the red/green overlay."""
    import struct

    res = subprocess.run([str(exe), "outputs", str(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0 and "OUTPUTS OK" in res.stdout, res.stdout + res.stderr
    lines = (tmp_path / "FrameTrajectory.txt").read_text().splitlines()
    assert lines == [
        "7 12.500000 KF 1.500000000 -2.250000000 3.000000000 0.000000000 0.000000000 0.000000000 1.000000000",
        "8 12.600000 KF, NMI 1.750000000 -2.250000000 3.000000000 0.000000000 0.000000000 0.000000000 1.000000000",
        "9 12.700000 KF, FAILED 1.750000000 -2.250000000 3.000000000 0.000000000 0.000000000 0.000000000 1.000000000",
        "10 12.800000 1.500000000 -2.250000000 3.000000000 0.000000000 0.000000000 0.000000000 1.000000000",
    ]
    twc = (tmp_path / "FrameTrajectory_twc.txt").read_text()
    assert twc.startswith("7 12.500000 KF\n[1, 0, 0, 1.5;\n 0, 1, 0, -2.25;\n 0, 0, 1, 3;\n 0, 0, 0, 1]\n")
    assert "8 12.600000 KF, NMI\n//////////Previous Poses" + "\\" * 10 + "\n[1, 0, 0, 1.5;" in twc
    assert "//////////Previous Poses End" + "\\" * 10 + "\n[1, 0, 0, 1.75;" in twc
    assert "\n11 " not in twc
    bmp = (tmp_path / "overlay.bmp").read_bytes()
    assert bmp[:2] == b"BM" and struct.unpack_from("<I", bmp, 2)[0] == len(bmp) == 54 + 16 * 3
    assert struct.unpack_from("<iiHH", bmp, 18) == (5, 3, 1, 24)
    # first stored row is the image's bottom row (y = 2): B 0, G render, R camera
    assert bmp[54:60] == bytes([0, 110, 20, 0, 111, 21]) and bmp[54 + 15] == 0
    assert bmp[54 + 32:54 + 35] == bytes([0, 100, 10])  # top row stored last
    name = [l for l in res.stdout.splitlines() if l.startswith("NAME ")][0][5:]
    assert name == "res/0012_NMI_[0.25]_WzyxSzyx_[0,1,2,2,1,0]_grid_[3x3x3_3x3x3].bmp"
    jname = [l for l in res.stdout.splitlines() if l.startswith("JNAME ")][0][6:]
    assert jname == name[:-4] + ".jpg"
    # the JPEG writer (cv::imwrite's container for the reference's .jpg overlays): any decoder reads it, and the
    # picture is the BMP's within quantisation error at quality 95
    jpg = (tmp_path / "overlay.jpg").read_bytes()
    assert jpg[:4] == b"\xff\xd8\xff\xe0" and jpg[6:11] == b"JFIF\0" and jpg[-2:] == b"\xff\xd9"
    cv2 = pytest.importorskip("cv2")
    got = cv2.imread(str(tmp_path / "overlay.jpg"), cv2.IMREAD_COLOR)
    want = cv2.imread(str(tmp_path / "overlay_ref.bmp"), cv2.IMREAD_COLOR)
    assert got is not None and got.shape == want.shape == (37, 83, 3)
    err = got.astype(np.float64) - want.astype(np.float64)
    psnr = 10 * np.log10(255.0 ** 2 / np.mean(err ** 2))
    assert psnr > 32.0 and np.abs(err).max() < 60, (psnr, np.abs(err).max())


def _loaded_cloud(d):
    """What loadXYZ produces (float64 parse, offset subtraction, float cast, /256, duplicate)."""
    raw = np.loadtxt(d / "cloud.xyz")
    xyz = (raw[:, :3] - np.array(OFFSET)).astype(np.float32)
    inten = (np.float32(1.0 / 256.0) * raw[:, 3].astype(np.float32)).astype(np.float32)
    pts = np.concatenate([xyz, inten[:, None]], axis=1).astype(np.float32)
    return np.vstack([pts, pts[-1:]])


@pytest.mark.gpu
def test_compat_reference_loop_on_gpu(exe, workdir, oracle):
    d, sc, frame = workdir
    import os

    res = subprocess.run([str(exe), "gpu", str(d / "settings.yaml"), str(d / "frame.raw"), str(d / "twc.txt")],
                         capture_output=True, text=True, cwd=d, timeout=600,
                         env=dict(os.environ, NMI_OUTPUT_LOC=str(d / "results")))
    assert res.returncode == 0 and "GPU OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    out = {l.split(" ", 1)[0]: l.split()[1:] for l in res.stdout.splitlines() if l and l.split()[0].isupper()}
    g = Grid.make((2, 1, 2), (1, 2, 1), (0.2, 0.2, 0.5), (0.02, 0.03, 0.05))
    sc.xyzi = _loaded_cloud(d)
    scores, _, _ = oracle.search_points(sc, sc.Twc, g, sc.xyzi, frame)
    batch = np.array(out["BATCH"], dtype=np.float64)
    loop = np.array(out["LOOP"], dtype=np.float64)
    assert np.allclose(batch, scores, rtol=1e-5, atol=0)
    # the zero-change call sequence (setupCam -> setCamera -> renderToTextureOnGPU -> NMIWithCuda_noMask):
    # setCamera gets the exact pose back from setupCam's memo, so it scores what the batched search scores
    assert np.array_equal(loop.astype(np.float32).view(np.uint32), batch.astype(np.float32).view(np.uint32))
    assert np.allclose(loop, scores, rtol=1e-5, atol=0)
    want, wmax = oracle.argmax(scores)
    s, w = oracle.unravel(g, want)
    assert [int(v) for v in out["BATCHBEST"][:6]] == list(s) + list(w)
    assert [int(v) for v in out["LOOPBEST"][:6]] == list(s) + list(w)
    new = np.array(out["NEWTWC"], dtype=np.float32).reshape(4, 4)
    assert np.array_equal(new, oracle.apply_winner(sc.Twc, g, s, w))
    reloc = out["RELOC"]
    assert int(reloc[2]) >= 2  # at least two levels always run (Tracking.cc:2108 needs i > 1)
    # _log.txt: one "NmiKernel / LastNmiKernel / Kernel rate" block per level (Tracking.cc:2103-2106)
    log = Path(out["LOGPATH"][0]).read_text()
    assert str(d / "results") in out["LOGPATH"][0]
    # two drivers ran over the same frame -- relocalize, then relocalizeSharded as rank 0 of 1 with
    # the same number of levels (checked inside the program) -- and both log every level
    n_blocks = 2 * int(reloc[2])
    assert log.count("\nNmiKernel:\tsX:") == n_blocks and log.count("\nLastNmiKernel:\tsX:") == n_blocks
    assert log.count("Kernel rate:\t") == n_blocks
    assert "Kernel rate:\tinf" in log  # first level: LastNmiKernel->NMI is 0 after reset()


@pytest.mark.gpu
def test_nmi_cuh_secondary_exports(tmp_path, oracle):
    """NMI.cuh:60-78 -- histogram256all + the three kernels, called like kernel.cu:63-100 does,
    with the render in a cudaArray (bottom-up rows, flipped like NMI.cu:82)."""
    import shutil

    lib = build.build_cuda()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = tmp_path / "test_nmi_cuh"
    cmd = [nvcc, "-ccbin", "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-std=c++17",
           "-gencode", "arch=compute_100a,code=sm_100a", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "cpp" / "test_nmi_cuh.cu"), "-L", str(lib.parent), "-lnmi_b200",
           f"-Xlinker=-rpath,{lib.parent}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    sc = synth.make_scene("tiny", n_points=5000)
    sc.W, sc.H = 150, 70  # not a multiple of 16 pixels per row
    sc.cx, sc.cy = 75.0, 35.0
    g = Grid.make((1, 1, 1), (1, 1, 1), (0.2, 0.2, 0.5), (0.02, 0.02, 0.05))
    _, render = oracle.render_points(sc, sc.Twc, oracle.cell_translation(sc.Twc, g, 0, 0, 0), sc.xyzi)
    warped = synth.frame_textured(sc.W, sc.H, seed=9)
    render[::-1].copy().tofile(tmp_path / "render_bottom_up.raw")  # GL row order
    warped.tofile(tmp_path / "warped.raw")
    res = subprocess.run([str(exe), str(sc.W), str(sc.H), str(tmp_path / "render_bottom_up.raw"),
                          str(tmp_path / "warped.raw"), str(tmp_path / "hist.bin")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "NMICUH OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    out = np.fromfile(tmp_path / "hist.bin", dtype=np.uint32)
    J, HA, HB = oracle.joint_hist(render, warped)
    assert np.array_equal(out[:65536].reshape(256, 256), J.reshape(256, 256))
    assert np.array_equal(out[65536:65536 + 256], HA) and np.array_equal(out[65536 + 256:], HB)
    want = oracle.score_f32(J, HA, HB, sc.W * sc.H)
    got = [float(v) for v in res.stdout.split("SCORE")[1].split()[:2]]
    assert got[0] == got[1]
    assert abs(got[0] - want) <= 1e-5 * abs(want)


# ---- Rendering<1>: the reference's default render mode (allProperties.hpp:42), OBJ + BMP -> GPU ----
@pytest.fixture(scope="module")
def exe_mesh(tmp_path_factory):
    """The same drop-in program compiled the way a RENDER_TEXTURE build of the reference compiles it:
    -Dnmi_prop_RENDER=1 -> NmiObjects::myRenderer is a Rendering<1> (loadOBJ + loadBMP_custom)."""
    lib = build.build_cuda()
    out = tmp_path_factory.mktemp("bin") / "test_compat_mesh"
    cmd = ["/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++", "-std=c++17", "-O1", "-Dnmi_prop_RENDER=1",
           "-ffp-contract=off", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cpp" / "test_compat.cpp"),
           "-L", str(lib.parent), "-lnmi_b200", f"-Wl,-rpath,{lib.parent}", "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return out


def _write_obj_bmp(d, verts, tris, uv, tex):
    """OBJ in the one dialect loadOBJ reads (v / vt / f a/b c/d e/f, 1-based, objloader.cpp:167-196) and
    a 24-bit BMP whose payload is `tex` as it lies (rows bottom-up, B,G,R)."""
    import struct

    lines = ["# synthetic terrain"]
    lines += ["v %.9g %.9g %.9g" % tuple(v[:3]) for v in verts]
    flat_uv = uv.reshape(-1, 2)
    lines += ["vt %.9g %.9g" % tuple(t) for t in flat_uv]
    for i, t in enumerate(tris):
        lines.append("f %d/%d %d/%d %d/%d" % (t[0] + 1, 3 * i + 1, t[1] + 1, 3 * i + 2, t[2] + 1, 3 * i + 3))
    (d / "mesh.obj").write_text("\n".join(lines) + "\n")
    th, tw, _ = tex.shape
    payload = tex.tobytes()  # width * 3 is a multiple of 4 here: no row padding
    assert (tw * 3) % 4 == 0
    hdr = b"BM" + struct.pack("<IHHI", 54 + len(payload), 0, 0, 54) + struct.pack(
        "<IiiHHIIiiII", 40, tw, th, 1, 24, 0, len(payload), 2835, 2835, 0, 0)
    (d / "tex.bmp").write_bytes(hdr + payload)


@pytest.fixture(scope="module")
def workdir_mesh(tmp_path_factory):
    d = tmp_path_factory.mktemp("compat_mesh")
    sc = synth.make_scene("tiny", n_points=10)
    verts, tris = synth.make_mesh(40, 40, extent=24.0)
    uv = synth.make_mesh_uv(verts, tris, extent=24.0, repeats=2.0)
    tex = synth.make_texture(64, 48)
    _write_obj_bmp(d, verts, tris, uv, tex)
    (d / "settings.yaml").write_text(YAML.format(fx=sc.fx, fy=sc.fy, cx=sc.cx, cy=sc.cy, W=sc.W, H=sc.H,
                                                 cloud="unused.xyz", offset="unused.xyz")
                                     .replace('"unused.obj"', '"%s"' % (d / "mesh.obj"))
                                     .replace('"unused.bmp"', '"%s"' % (d / "tex.bmp")))
    frame = synth.frame_textured(sc.W, sc.H, seed=21)
    frame.tofile(d / "frame.raw")
    np.savetxt(d / "twc.txt", sc.Twc.reshape(1, 16), fmt="%.9g")
    return d, sc, frame, verts, tris, uv, tex


def test_compat_mesh_build_compiles(exe_mesh):
    assert exe_mesh.exists()


@pytest.mark.gpu
def test_compat_textured_mesh_on_gpu(exe_mesh, workdir_mesh, oracle):
    """OBJ + BMP files -> Rendering<1> -> nmi_set_mesh_textured -> the reference's loop and the batched
    search, against the oracle's textured renderer fed with what the loaders produce."""
    import os

    d, sc, frame, verts, tris, uv, tex = workdir_mesh
    res = subprocess.run([str(exe_mesh), "gpu", str(d / "settings.yaml"), str(d / "frame.raw"), str(d / "twc.txt")],
                         capture_output=True, text=True, cwd=d, timeout=600,
                         env=dict(os.environ, NMI_OUTPUT_LOC=str(d / "results")))
    assert res.returncode == 0 and "GPU OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    out = {l.split(" ", 1)[0]: l.split()[1:] for l in res.stdout.splitlines() if l and l.split()[0].isupper()}
    g = Grid.make((2, 1, 2), (1, 2, 1), (0.2, 0.2, 0.5), (0.02, 0.03, 0.05))
    # what loadOBJ builds: one vertex per face corner (objloader.cpp:206-220), parsed from "%.9g" text
    corners = np.array([[float("%.9g" % c) for c in verts[k, :3]] for k in tris.reshape(-1)], np.float32)
    v4 = np.concatenate([corners, np.zeros((corners.shape[0], 1), np.float32)], axis=1)
    t3 = np.arange(corners.shape[0], dtype=np.uint32).reshape(-1, 3)
    uvp = np.array([[float("%.9g" % c) for c in t] for t in uv.reshape(-1, 2)], np.float32).reshape(-1, 3, 2)
    scores, renders, _ = oracle.search_mesh_tex(sc, sc.Twc, g, v4, t3, uvp, tex, frame, keep_images=True)
    assert (renders[0] != 255).mean() > 0.9
    batch = np.array(out["BATCH"], dtype=np.float64)
    loop = np.array(out["LOOP"], dtype=np.float64)
    assert np.allclose(batch, scores, rtol=1e-5, atol=0)
    assert np.allclose(loop, scores, rtol=1e-5, atol=0)  # per-call path: exact pose through setupCam's memo
    want, _ = oracle.argmax(scores)
    s, w = oracle.unravel(g, want)
    assert [int(v) for v in out["BATCHBEST"][:6]] == list(s) + list(w)
