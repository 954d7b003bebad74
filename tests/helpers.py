"""Shared test inputs: the reference's own data/ images and calibration, carried in
tests/golden/fixtures.npz (made by oracle/gen_golden.py), plus the SURVEY 8(d)
rescaling recipe for the BASELINE.json configs."""
import hashlib
import json
import os

import cv2
import numpy as np

from oracle.cv2_path import Geometry, padding, rescale_calib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("front", "back", "left", "right")


def h16(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


class Fixtures:
    def __init__(self):
        z = np.load(os.path.join(GOLD, "fixtures.npz"))
        self.gold = json.load(open(os.path.join(GOLD, "golden.json")))
        self.calib = {n: (z[f"K_{n}"], z[f"D_{n}"], z[f"H_{n}"]) for n in NAMES}
        self.D5 = z["D5_synth"]
        self._jpg = {k[4:]: z[k] for k in z.files if k.startswith("jpg_")}
        self._dec = {}

    def img(self, key):
        if key not in self._dec:
            self._dec[key] = cv2.imdecode(self._jpg[key], cv2.IMREAD_COLOR)
        return self._dec[key]

    def frames(self, FW=1280, FH=1024):
        f = [self.img(n) for n in NAMES]
        if (FW, FH) != (1280, 1024):
            f = [cv2.resize(x, (FW, FH), interpolation=cv2.INTER_LINEAR) for x in f]
        return f

    def car(self, BW=1000, BH=1000):
        CW, CH = int(250 * BW / 1000), int(400 * BH / 1000)
        c = self.img("car")
        if (CW, CH) != (250, 400):
            c = cv2.resize(c, (CW, CH))
        return padding(c, BW, BH)

    def geometry(self, FW=1280, FH=1024, BW=1000, BH=1000, CW=None, CH=None):
        return Geometry(FW=FW, FH=FH, BW=BW, BH=BH,
                        CW=int(250 * BW / 1000) if CW is None else CW,
                        CH=int(400 * BH / 1000) if CH is None else CH)

    def scaled_calib(self, g: Geometry):
        out = {}
        for n in NAMES:
            K, D, H = self.calib[n]
            K2, H2 = rescale_calib(K, H, g)
            out[n] = (K2, D, H2)
        return out

    def perturbed_frames(self, FW, FH, i):
        """cfg4 frame-set i: fixture frames blended 50/50 with seeded noise."""
        base = self.frames(FW, FH)
        out = []
        for c, f in enumerate(base):
            r = np.random.default_rng(1234 + 4 * i + c).integers(0, 256, f.shape, dtype=np.uint8)
            out.append(((f.astype(np.uint16) + r) >> 1).astype(np.uint8))
        return out
