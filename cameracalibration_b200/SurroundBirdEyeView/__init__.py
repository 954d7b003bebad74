"""Bird's-eye-view stitching of four fisheye cameras on the B200 engine (libbevk.so).

``BevGenerator`` keeps the reference's constructor and call signature; geometry comes from the
namespace returned by ``BevGenerator.get_args()`` and is read when a generator is constructed.
See surroundBEV.py in this package for the differences that are deliberate.
"""
from .surroundBEV import BevGenerator  # noqa: F401
