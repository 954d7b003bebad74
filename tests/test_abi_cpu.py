"""CPU-side checks of the drop-in boundary: libbevk.so loads, exports every symbol that
include/bevk.h declares, the ctypes table matches the header, and the product refuses to
run without a GPU instead of falling back to anything."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bevk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bevk_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from cameracalibration_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bevk.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.load().bevk_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cameracalibration_b200 import _lib
    with pytest.raises(_lib.BevkError, match="no CPU fallback"):
        _lib.Context(0)
    from cameracalibration_b200 import ops
    import numpy as np
    with pytest.raises(_lib.BevkError):
        ops.remap(np.zeros((4, 4, 3), np.uint8), np.zeros((2, 2, 2), np.int16), np.zeros((2, 2), np.uint16))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cameracalibration_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions oracle/"


def test_shim_import_leaves_argv_alone_and_keeps_reference_names(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["prog", "--some-foreign-flag", "-x"])
    from cameracalibration_b200.SurroundBirdEyeView import surroundBEV as S
    from cameracalibration_b200.SurroundBirdEyeView import BevGenerator
    a = BevGenerator.get_args()
    for k, v in dict(FRAME_WIDTH=1280, FRAME_HEIGHT=1024, BEV_WIDTH=1000, BEV_HEIGHT=1000, CAR_WIDTH=250,
                     CAR_HEIGHT=400, FOCAL_SCALE=1, SIZE_SCALE=2, BLEND_FLAG=False, BALANCE_FLAG=False).items():
        assert getattr(a, k) == v
    for name in ("padding", "color_balance", "luminance_balance", "Camera", "Mask", "BlendMask", "BevGenerator"):
        assert hasattr(S, name)
    from cameracalibration_b200.IntrinsicCalibration import InCalibrator
    from cameracalibration_b200.ExtrinsicCalibration import ExCalibrator
    assert hasattr(InCalibrator, "undistort") and hasattr(ExCalibrator, "warp")
    with pytest.raises(Exception, match="camera should be fisheye/normal"):
        InCalibrator("wide")


def test_mask_polygons_match_reference_geometry(fx):
    """Host-side polygon tables of the shim == the oracle's (which == the reference's)."""
    from cameracalibration_b200.SurroundBirdEyeView import surroundBEV as S
    from oracle import restate as R
    import numpy as np
    for (BW, BH, CW, CH) in [(1000, 1000, 250, 400), (1200, 1200, 300, 480), (777, 900, 194, 360)]:
        g = S._Geo()
        g.BW, g.BH, g.CW, g.CH = BW, BH, CW, CH
        for n in S.NAMES:
            assert (S._plain_points(n, g) == R.plain_polygon(n, BW, BH, CW, CH)).all()
            assert (S._blend_points(n, g) == R.blend_polygon(n, BW, BH, CW, CH)).all()
        L = R.blend_lines(BW, BH, CW, CH)
        got = S._seam_lines(g)
        for i, k in enumerate(["FL", "FR", "BL", "BR", "LF", "LB", "RF", "RB"]):
            assert (got[i] == L[k]).all()
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    from oracle.cv2_path import padding
    assert (S.padding(img, 12, 10) == padding(img, 12, 10)).all()
    assert (S.padding(img, 11, 9) == padding(img, 11, 9)).all()


def test_docs_name_only_real_entry_points():
    """Every bevk_* function named in INTEGRATION.md / DESIGN.md / README.md is declared in include/bevk.h."""
    declared = set(_declared())
    not_functions = {"bevk_ctx", "bevk_status", "bevk_bev", "bevk_api", "bevk_kernels", "bevk_device", "bevk_gather4", "bevk_plan",
                     "bevk_bev_tma", "bevk_plan_tma", "bevk_shard", "bevk_und_src", "bevk_und_dst", "bevk_und_ref"}   # source files, temp dirs
    wildcards = {"bevk_shard", "bevk_graph", "bevk_bev_run"}   # "bevk_shard_*" style family names
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        txt = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(bevk_[a-z0-9_]*[a-z0-9])\b", txt)) - not_functions - wildcards:
            assert name in declared, f"{doc} mentions {name}, which include/bevk.h does not declare"


def test_tools_undistort_directory_pipeline(tmp_path):
    """Host logic of the overlapped Tools/undistort pipeline (reference Tools/undistort.py:59-77): listing order,
    -name numbering, format dispatch -- with a stand-in for the GPU call (no device here)."""
    import cv2
    import numpy as np
    from cameracalibration_b200.Tools import undistort as T
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir(); dst.mkdir()
    rng = np.random.default_rng(5)
    imgs = {f"f{i:02d}.png": rng.integers(0, 256, (24, 32, 3), dtype=np.uint8) for i in range(11)}
    for n, im in imgs.items():
        cv2.imwrite(str(src / n), im)
    (src / "skip.jpg").write_bytes(b"not read")
    listing = [e for e in os.listdir(src) if e.endswith(".png")]
    seen = []
    def fake(img):
        seen.append(img.copy())
        return 255 - img
    for workers in (1, 3):
        seen.clear()
        opts = T.make_parser().parse_args(["-path_read", str(src) + "/", "-path_save", str(dst) + "/", "-srcformat", "png",
                                           "-dstformat", "png", "-quality", "1", "-workers", str(workers)])
        assert T.run_directory(opts, fake, cv2) == listing
        assert all((a == imgs[n]).all() for a, n in zip(seen, listing))          # GPU calls happen in listing order
        for n in listing:
            assert (cv2.imread(str(dst / n)) == 255 - imgs[n]).all()
    opts = T.make_parser().parse_args(["-path_read", str(src) + "/", "-path_save", str(dst) + "/", "-srcformat", "png",
                                       "-dstformat", "png", "-quality", "1", "-name", "cam", "-workers", "2"])
    assert T.run_directory(opts, fake, cv2) == [f"cam_{i:04d}.png" for i in range(1, 12)]
    assert (cv2.imread(str(dst / "cam_0003.png")) == 255 - imgs[listing[2]]).all()
    def boom(img):
        raise RuntimeError("device lost")
    with pytest.raises(RuntimeError, match="device lost"):
        T.run_directory(opts, boom, cv2)


def test_cuda_array_interface_validation():
    """Host-side checks of run_cuda's inputs (no device needed: a stand-in object carries the interface)."""
    from cameracalibration_b200 import _lib as L
    from cameracalibration_b200 import ops

    class Arr:
        def __init__(self, shape, strides=None, typestr="|u1", ptr=4096):
            self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr, False), strides=strides, version=3)

    assert ops._cuda_ptr(Arr((2, 3, 4)), (2, 3, 4)) == (4096, (2, 3, 4))
    assert ops._cuda_ptr(Arr((2, 3, 4), (12, 4, 1)), None) == (4096, (2, 3, 4))
    assert ops._cuda_ptr(Arr((1, 3, 4), (999, 4, 1)), None)[0] == 4096          # stride of a length-1 axis is free
    for bad, msg in ((Arr((2, 3, 4), (24, 4, 1)), "C-contiguous"), (Arr((2, 3, 4), typestr="<f4"), "uint8"),
                     (Arr((2, 3, 4), ptr=0), "null"), (object(), "__cuda_array_interface__")):
        with pytest.raises(L.BevkError, match=msg):
            ops._cuda_ptr(bad, None)
    with pytest.raises(L.BevkError, match="shape"):
        ops._cuda_ptr(Arr((2, 3, 4)), (2, 3, 5))
