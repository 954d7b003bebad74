"""Small batched run of the fused BEV path for compute-sanitizer: k_bev_tma<false,4>, <true,4>, <false,1> (plain and
BALANCE, a ragged batch of 6, car overlay) on a 640x512 -> 500x500 rig, checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cameracalibration_b200 import ops  # noqa: E402
from oracle import cv2_path as C  # noqa: E402
from oracle import restate as R  # noqa: E402
from tests.helpers import NAMES, Fixtures  # noqa: E402

fx = Fixtures()
g = fx.geometry(640, 512, 500, 500)
calib = fx.scaled_calib(g)
eng = ops.BevEngine(4, (g.FW, g.FH), (g.BW, g.BH))
masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) for n in NAMES]
for i, n in enumerate(NAMES):
    K, D, H = calib[n]
    eng.set_camera(i, K, D, C.dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS), (int(g.FW * g.SS), int(g.FH * g.SS)), H)
    eng.set_mask(i, masks[i])
eng.finalize()
F = fx.frames(g.FW, g.FH)
sets = [[np.ascontiguousarray(np.roll(f, 9 * i, axis=1)) for f in F] for i in range(6)]
car = fx.car(g.BW, g.BH)
ref = C.RefBev(calib, g, True, False, masks=masks)
for balance in (False, True):
    ref.balance = balance
    got = eng.run(sets, car, balance)
    assert eng.last_path() == "tma"
    for i in (0, 5):
        assert (got[i] == ref(*sets[i], car)).all(), (balance, i)
    one = eng.run(sets[:1], car, balance)
    assert (one[0] == got[0]).all()
print("sanitize_small ok, launches", eng.ctx.launches)
