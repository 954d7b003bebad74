"""Surround Camera Bird Eye View Generator on the B200 engine.

    from cameracalibration_b200.SurroundBirdEyeView import BevGenerator
    bev = BevGenerator()                       # real-time path
    surround = bev(front, back, left, right)
    bev = BevGenerator(blend=True, balance=True)
    surround = bev(front, back, left, right, car)

    args = BevGenerator.get_args(); args.CAR_WIDTH = 200; args.CAR_HEIGHT = 350
"""
from .surroundBEV import BevGenerator  # noqa: F401
