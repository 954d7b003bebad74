from .intrinsicCalib import InCalibrator  # noqa: F401
