"""The reference's CPU path, restated as explicit-parameter functions over cv2.

TEST INFRASTRUCTURE ONLY (see oracle/restate.py header).  The reference
(dyfcalid/CameraCalibration) is pure Python over OpenCV; its classes read module
globals set by argparse at import.  This module issues the *same cv2 calls in the
same order* with the geometry passed explicitly, so that it can (a) run on the
GPU box where /root/reference does not exist, (b) be timed as the CPU baseline
(``bench.py`` ``cpu_baseline`` / ``--impl reference``), and (c) be checked
against the unmodified reference classes where /root/reference exists
(tests/test_oracle.py).  Each function cites the lines it follows.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np

from . import restate

NAMES = restate.NAMES


@dataclass
class Geometry:
    """surroundBEV.py:7-14 argparse defaults."""
    FW: int = 1280
    FH: int = 1024
    BW: int = 1000
    BH: int = 1000
    CW: int = 250
    CH: int = 400
    FS: float = 1.0
    SS: float = 2.0


def rescale_calib(K, Hm, g: Geometry, base=(1280, 1024, 1000, 1000)):
    """SURVEY 8(d) recipe: K' = diag(sx,sy,1) K ; H' = diag(bx,by,1) H diag(1/sx,1/sy,1)."""
    sx, sy = g.FW / base[0], g.FH / base[1]
    bx, by = g.BW / base[2], g.BH / base[3]
    S = np.diag([sx, sy, 1.0])
    B = np.diag([bx, by, 1.0])
    return S @ np.asarray(K, np.float64), B @ np.asarray(Hm, np.float64) @ np.linalg.inv(S)


def dst_camera_matrix(K, FW, FH, FS, SS, off_h=0.0, off_v=0.0):
    """surroundBEV.py:90-96 == intrinsicCalib.py:90-96 == undistort.py:42-46."""
    P = np.array(K, np.float64)
    P[0][0] *= FS
    P[1][1] *= FS
    P[0][2] = FW / 2 * SS + off_h
    P[1][2] = FH / 2 * SS + off_v
    return P


def undistort_maps(K, D, P, W, H):
    """surroundBEV.py:98-103."""
    return cv2.fisheye.initUndistortRectifyMap(
        np.asarray(K, np.float64), np.asarray(D, np.float64), np.eye(3), P, (int(W), int(H)), cv2.CV_16SC2)


def pinhole_maps(K, D5, P, W, H):
    """intrinsicCalib.py:158-163."""
    return cv2.initUndistortRectifyMap(np.asarray(K, np.float64), np.asarray(D5, np.float64),
                                       np.eye(3), P, (int(W), int(H)), cv2.CV_16SC2)


class RefCamera:
    """Camera (surroundBEV.py:81-117) with explicit K, D, H and geometry."""

    def __init__(self, K, D, Hm, g: Geometry):
        self.K, self.D, self.H, self.g = K, D, Hm, g
        self.P = dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS)
        self.undistort_maps = undistort_maps(K, D, self.P, int(g.FW * g.SS), int(g.FH * g.SS))
        self.bev_maps = (self.warp_homography(self.undistort_maps[0]),
                         self.warp_homography(self.undistort_maps[1]))

    def undistort(self, img):
        return cv2.remap(img, *self.undistort_maps, interpolation=cv2.INTER_LINEAR)

    def warp_homography(self, img):
        return cv2.warpPerspective(img, self.H, (self.g.BW, self.g.BH))

    def raw2bev(self, img):
        return cv2.remap(img, *self.bev_maps, interpolation=cv2.INTER_LINEAR)


def plain_mask(name, g: Geometry):
    """Mask.get_mask (surroundBEV.py:156-159)."""
    return restate.fill_poly(g.BW, g.BH, restate.plain_polygon(name, g.BW, g.BH, g.CW, g.CH))


def blend_mask_loop(name, g: Geometry):
    """BlendMask.__init__ incl. the per-pixel pointPolygonTest loop (surroundBEV.py:165-188,
    270-277).  Slow (seconds); used to pin restate.blend_mask."""
    polys = {n: restate.fill_poly(g.BW, g.BH, restate.blend_polygon(n, g.BW, g.BH, g.CW, g.CH)) for n in NAMES}
    L = restate.blend_lines(g.BW, g.BH, g.CW, g.CH)
    order = {"front": [("left", "FL", "LF"), ("right", "FR", "RF")],
             "back": [("left", "BL", "LB"), ("right", "BR", "RB")],
             "left": [("front", "LF", "FL"), ("back", "LB", "BL")],
             "right": [("front", "RF", "FR"), ("back", "RB", "BR")]}[name]
    m = polys[name]
    for other, la, lb in order:
        ov = cv2.bitwise_and(m, polys[other])
        for y, x in zip(*np.where(ov != 0)):
            dA = cv2.pointPolygonTest(L[la], (float(x), float(y)), True)
            dB = cv2.pointPolygonTest(L[lb], (float(x), float(y)), True)
            m[y, x] = dA ** 2 / (dA ** 2 + dB ** 2 + 1e-6) * 255
    return m


def luminance_balance(images):
    """surroundBEV.py:57-79."""
    hsv = [cv2.cvtColor(im, cv2.COLOR_BGR2HSV) for im in images]
    planes = [list(cv2.split(x)) for x in hsv]
    means = [np.mean(p[2]) for p in planes]
    v_mean = (means[0] + means[1] + means[2] + means[3]) / 4
    out = []
    for p, m in zip(planes, means):
        v = cv2.add(p[2], (v_mean - m))
        out.append(cv2.cvtColor(cv2.merge([p[0], p[1], v]), cv2.COLOR_HSV2BGR))
    return out


def color_balance(image):
    """surroundBEV.py:43-55."""
    b, g, r = cv2.split(image)
    B, G, R = np.mean(b), np.mean(g), np.mean(r)
    K = (R + G + B) / 3
    cv2.addWeighted(b, K / B, 0, 0, 0, b)
    cv2.addWeighted(g, K / G, 0, 0, 0, g)
    cv2.addWeighted(r, K / R, 0, 0, 0, r)
    return cv2.merge([b, g, r])


def padding(img, width, height):
    """surroundBEV.py:28-41."""
    h, w = img.shape[:2]
    top = bottom = (height - h) // 2
    if top + bottom + h < height:
        bottom += 1
    left = right = (width - w) // 2
    if left + right + w < width:
        right += 1
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(0, 0, 0))


class RefBev:
    """BevGenerator (surroundBEV.py:282-325) with explicit calibration and geometry.
    ``masks`` may be injected (e.g. restate.blend_mask output) to skip the slow loop."""

    def __init__(self, calib, g: Geometry, blend=False, balance=False, masks=None):
        self.g, self.blend, self.balance = g, blend, balance
        self.cameras = [RefCamera(*calib[n], g) for n in NAMES]
        if masks is None:
            masks = [blend_mask_loop(n, g) if blend else plain_mask(n, g) for n in NAMES]
        self.masks = masks
        if blend:
            self.weights = [(np.repeat(m[:, :, None], 3, axis=2) / 255.0).astype(np.float32) for m in masks]

    def __call__(self, front, back, left, right, car=None):
        images = [front, back, left, right]
        if self.balance:
            images = luminance_balance(images)
        tiles = []
        for i, (img, cam) in enumerate(zip(images, self.cameras)):
            w = cam.raw2bev(img)
            if self.blend:
                tiles.append((w * self.weights[i]).astype(np.uint8))      # surroundBEV.py:279-280
            else:
                tiles.append(cv2.bitwise_and(w, w, mask=self.masks[i]))   # surroundBEV.py:161-162
        s = cv2.add(tiles[0], tiles[1])
        s = cv2.add(s, tiles[2])
        s = cv2.add(s, tiles[3])
        if self.balance:
            s = color_balance(s)
        if car is not None:
            s = cv2.add(s, car)
        return s
