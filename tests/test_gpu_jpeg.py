"""GPU test of the JPEG ingest (bevk_jpeg_decode / bevk_bev_run_jpeg): the streams are decoded on the device by nvJPEG
straight into the frame stack the fused kernel reads.

Parity statement: nvJPEG's decoded pixels are not libjpeg-turbo's (cv2.imread: different IDCT / chroma-upsampling
rounding), so the check has two parts -- (a) the decode is a faithful JPEG decode: within a few LSB of cv2.imdecode on
the reference's own data/ JPEGs; (b) everything downstream is bit-exact: the canvases of bevk_bev_run_jpeg equal the
oracle (the reference's cv2 call sequence) applied to the very frames nvJPEG produced."""
import numpy as np
import pytest

from oracle import cv2_path as C
from oracle import restate as R
from tests.helpers import NAMES

pytestmark = pytest.mark.gpu


def test_jpeg_streams_to_bev(fx):
    import torch
    from cameracalibration_b200 import ops
    g = fx.geometry()
    e = ops.BevEngine(4, (g.FW, g.FH), (g.BW, g.BH))
    masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) for n in NAMES]
    for i, n in enumerate(NAMES):
        K, D, H = fx.calib[n]
        e.set_camera(i, K, D, C.dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS), (int(g.FW * g.SS), int(g.FH * g.SS)), H)
        e.set_mask(i, masks[i])
    e.finalize()
    streams = [bytes(fx._jpg[n]) for n in NAMES]               # the reference's data/<name>/<name>.jpg files, byte for byte
    dec = ops.jpeg_decode(streams, g.FW, g.FH, ctx=e.ctx)
    e.ctx.sync()
    dec = dec.cpu().numpy()
    for k, n in enumerate(NAMES):                               # (a) a faithful decode
        diff = np.abs(dec[k].astype(np.int16) - fx.img(n).astype(np.int16))
        psnr = 10 * np.log10(255.0 ** 2 / max(1e-9, float((diff.astype(np.float64) ** 2).mean())))
        # a handful of pixels at sharp chroma edges differ by tens of levels (the two decoders upsample 4:2:0 chroma
        # differently); the image as a whole must be the same picture
        assert diff.mean() < 0.8 and psnr > 42 and (diff > 8).mean() < 2e-3, (n, int(diff.max()), float(diff.mean()), psnr)
    car = fx.car()
    ref = C.RefBev(fx.calib, g, True, False, masks=masks)
    for balance in (False, True):                               # (b) bit-exact downstream
        ref.balance = balance
        got = e.run_jpeg([streams, streams[::-1]], car, balance)
        assert e.last_path() == "tma"
        assert (got[0] == ref(*dec, car)).all(), balance
        assert (got[1] == ref(*dec[::-1], car)).all(), balance
    with pytest.raises(Exception, match="not a JPEG|holds"):
        e.run_jpeg([[streams[0][:100]] + streams[1:]])
