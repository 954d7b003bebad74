"""Drop-in for the hot-path part of the reference's ExtrinsicCalibration/extrinsicCalib.py:
``ExCalibrator.warp()`` = cv2.warpPerspective(src_img, homography, dst size) on the GPU
(reference extrinsicCalib.py:166-169).

Estimating the homography (chessboard corners in two views + cv2.findHomography(RANSAC),
reference :155-183) is offline, irregular work outside the hot path: set
``src_img``, ``dst_img`` (only its shape is used) and ``homography`` directly, or use
``set_views(src_img, dst_img, homography)``.
"""
from __future__ import annotations

import argparse

import numpy as np

from .. import ops

parser = argparse.ArgumentParser(description="Homography from Source to Destination Image (B200 warp path)")
parser.add_argument("-id", "--CAMERA_ID", default=1, type=int)
args = parser.parse_known_args([])[0]


class ExCalibrator:
    def __init__(self):
        self.src_corners_total = np.empty([0, 1, 2])
        self.dst_corners_total = np.empty([0, 1, 2])
        self.src_img = None
        self.dst_img = None
        self.homography = None

    @staticmethod
    def get_args():
        return args

    def set_views(self, src_img, dst_img, homography):
        self.src_img, self.dst_img = src_img, dst_img
        self.homography = np.asarray(homography, np.float64)
        return self.homography

    def warp(self):
        if self.src_img is None or self.dst_img is None or self.homography is None:
            raise Exception("src_img, dst_img and homography must be set before warp()")
        return ops.warp_perspective(self.src_img, self.homography, (self.dst_img.shape[1], self.dst_img.shape[0]))

    def __call__(self, src_img, dst_img):
        raise Exception("corner detection / cv2.findHomography is outside the B200 hot path (see module docstring)")
