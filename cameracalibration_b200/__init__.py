"""cameracalibration_b200 -- B200 (sm_100a) engine for the surround-BEV warping hot path of
dyfcalid/CameraCalibration.  Sub-packages mirror the reference's layout:

    cameracalibration_b200.SurroundBirdEyeView.BevGenerator
    cameracalibration_b200.IntrinsicCalibration.InCalibrator      (.undistort)
    cameracalibration_b200.ExtrinsicCalibration.ExCalibrator      (.warp)
    cameracalibration_b200.Tools.undistort                        (batch CLI)

All per-pixel work runs in cameracalibration_b200/libbevk.so (C ABI: include/bevk.h).
"""
from ._lib import BevkError, Context, default_context, pinned_empty  # noqa: F401

__all__ = ["BevkError", "Context", "default_context", "pinned_empty"]
