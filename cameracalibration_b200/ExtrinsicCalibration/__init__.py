from .extrinsicCalib import ExCalibrator  # noqa: F401
