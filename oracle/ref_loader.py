"""Import the UNMODIFIED reference from /root/reference (this container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; callers
must check ``available()`` first.  Non-invasive shims:
  * sys.argv neutralised while importing (argparse runs at import,
    surroundBEV.py:17, intrinsicCalib.py:29, extrinsicCalib.py:20);
  * cv2.pointPolygonTest wrapped to cast the point to Python floats (cv2 4.13 +
    numpy 2 reject numpy.int64 tuples at surroundBEV.py:274);
  * numpy.load optionally wrapped to rescale K/H for non-default geometry
    (SURVEY 8d recipe) -- the reference hard-wires the fixture paths
    (surroundBEV.py:83-85).
Nothing is copied out of the reference tree.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys

import numpy as np

REF = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "SurroundBirdEyeView"))


def _import(modname: str):
    import cv2
    if not getattr(cv2.pointPolygonTest, "_bevk_shim", False):
        orig = cv2.pointPolygonTest

        def ppt(contour, pt, measure):
            return orig(contour, (float(pt[0]), float(pt[1])), measure)
        ppt._bevk_shim = True
        cv2.pointPolygonTest = ppt
    argv, sys.argv = sys.argv, ["x"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        return importlib.import_module(modname)
    finally:
        sys.argv = argv


def surround():
    return _import("SurroundBirdEyeView.surroundBEV")


def intrinsic():
    return _import("IntrinsicCalibration.intrinsicCalib")


def extrinsic():
    return _import("ExtrinsicCalibration.extrinsicCalib")


@contextlib.contextmanager
def rescaled_calibration(FW, FH, BW, BH):
    """np.load shim: camera_*_K.npy / camera_*_H.npy come back rescaled."""
    from .cv2_path import Geometry, rescale_calib
    g = Geometry(FW=FW, FH=FH, BW=BW, BH=BH)
    orig = np.load

    def load(path, *a, **k):
        arr = orig(path, *a, **k)
        p = str(path)
        if p.endswith("_K.npy"):
            Hm = orig(p[:-6] + "_H.npy")
            return rescale_calib(arr, Hm, g)[0]
        if p.endswith("_H.npy"):
            K = orig(p[:-6] + "_K.npy")
            return rescale_calib(K, arr, g)[1]
        return arr
    np.load = load
    try:
        yield
    finally:
        np.load = orig


def make_bev(FW=1280, FH=1024, BW=1000, BH=1000, CW=250, CH=400, FS=1, SS=2, blend=False, balance=False):
    """Construct the reference's BevGenerator at a given geometry."""
    m = surround()
    a = m.BevGenerator.get_args()
    a.FRAME_WIDTH, a.FRAME_HEIGHT, a.BEV_WIDTH, a.BEV_HEIGHT = FW, FH, BW, BH
    a.CAR_WIDTH, a.CAR_HEIGHT, a.FOCAL_SCALE, a.SIZE_SCALE = CW, CH, FS, SS
    with rescaled_calibration(FW, FH, BW, BH):
        return m.BevGenerator(blend=blend, balance=balance)
