"""Batch fisheye undistortion (the reference's Tools/undistort.py:25-77) on the GPU: the
map is built once on the device and stays there; every image is one H2D copy, one
gather kernel and one D2H copy.  File decode / encode stays on the host with cv2, as in
the reference.  Same command-line flags; ``-load`` additionally accepts 0/1/true/false
(the reference's ``type=bool`` makes every non-empty string True)."""
from __future__ import annotations

import argparse
import os

import numpy as np

from .. import ops


def _flag(s):
    return str(s).lower() not in ("0", "false", "no", "")


def make_parser():
    p = argparse.ArgumentParser(description="Fisheye Camera Undistortion (B200)")
    p.add_argument("-width", default=1280, type=int)
    p.add_argument("-height", default=1024, type=int)
    p.add_argument("-load", default=True, type=_flag)
    p.add_argument("-path_read", default="./data/", type=str)
    p.add_argument("-path_save", default="./", type=str)
    p.add_argument("-path_k", default="./data/camera_0_K.npy", type=str)
    p.add_argument("-path_d", default="./data/camera_0_D.npy", type=str)
    p.add_argument("-focalscale", default=1, type=float)
    p.add_argument("-sizescale", default=1, type=float)
    p.add_argument("-offset_h", default=0, type=float)
    p.add_argument("-offset_v", default=0, type=float)
    p.add_argument("-srcformat", default="jpg", type=str)
    p.add_argument("-dstformat", default="jpg", type=str)
    p.add_argument("-quality", default=100, type=int)
    p.add_argument("-name", default=None, type=str)
    p.add_argument("-fused", default=False, type=_flag, help="evaluate the camera model in-kernel (no map in HBM)")
    return p


DEFAULT_K = [[350.4931893001142, 0.0, 647.6297467576265], [0.0, 352.43072872484805, 513.5196785119657], [0.0, 0.0, 1.0]]
DEFAULT_D = [[-0.03367245449576437], [0.015380779195912842], [-0.018654590946883556], [0.0058128945633924185]]


def build_undistorter(a) -> ops.Undistorter:
    if not a.load:
        K, D = np.array(DEFAULT_K), np.array(DEFAULT_D)
    else:
        if not os.path.exists(a.path_k):
            raise Exception("Camera K File Path not exist")
        if not os.path.exists(a.path_d):
            raise Exception("Camera D File Path not exist")
        K, D = np.load(a.path_k), np.load(a.path_d)
    P = K.copy()
    P[0, 0] *= a.focalscale
    P[1, 1] *= a.focalscale
    P[0, 2] = a.width / 2 * a.sizescale + a.offset_h
    P[1, 2] = a.height / 2 * a.sizescale + a.offset_v
    return ops.Undistorter(K, D, P, (int(a.width * a.sizescale), int(a.height * a.sizescale)), fused=a.fused)


def main(argv=None):
    import cv2
    a = make_parser().parse_args(argv)
    und = build_undistorter(a)
    if not os.path.exists(a.path_read):
        raise Exception("Original Image Read Path not exist")
    if not os.path.exists(a.path_save):
        raise Exception("Undistortion Image Save Path not exist")
    index, done = 1, []
    for filename in os.listdir(a.path_read):
        if filename[-4:] != "." + a.srcformat:
            continue
        img = und(cv2.imread(os.path.join(a.path_read, filename)))
        if a.name is not None:
            filename = a.name + "_{:04d}.".format(index) + a.srcformat
            index += 1
        stem = os.path.join(a.path_save, filename[:-4])
        if a.dstformat == "jpg":
            cv2.imwrite(stem + ".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, a.quality])
        elif a.dstformat == "png":
            cv2.imwrite(stem + ".png", img, [cv2.IMWRITE_PNG_COMPRESSION, a.quality])
        else:
            cv2.imwrite(filename[:-4] + "." + a.dstformat, img)
        done.append(filename)
    return done


if __name__ == "__main__":
    main()
