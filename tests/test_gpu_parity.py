"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every check goes through the
C ABI (ctypes -> libbevk.so) and compares bit-for-bit with (a) the golden hashes made from
the unmodified reference, (b) the oracle (cv2 call sequence / NumPy restatement) run live
on the same inputs.  Bar: bit-exact for all of it (integer / fixed-point arithmetic; the
FP64/FP32 steps are reproduced operation for operation)."""
import os

import cv2
import numpy as np
import pytest

from oracle import cv2_path as C
from oracle import restate as R
from tests.helpers import NAMES, h16

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from cameracalibration_b200 import ops as o
    return o


def _engine(ops, fx, g, blend, calib=None, masks=None):
    calib = calib or fx.scaled_calib(g)
    e = ops.BevEngine(4, (g.FW, g.FH), (g.BW, g.BH))
    if masks is None:
        masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) if blend else C.plain_mask(n, g) for n in NAMES]
    for i, n in enumerate(NAMES):
        K, D, H = calib[n]
        e.set_camera(i, K, D, C.dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS), (int(g.FW * g.SS), int(g.FH * g.SS)), H)
        e.set_mask(i, masks[i])
    e.finalize()
    return e, masks


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", NAMES)
def test_fisheye_map_golden(ops, fx, name):
    K, D, _ = fx.calib[name]
    P = C.dst_camera_matrix(K, 1280, 1024, 1, 2)
    m1, m2 = ops.fisheye_init_undistort_rectify_map(K, D, P, (2560, 2048))
    g = fx.gold["camera"][name]
    assert (h16(m1), h16(m2)) == (g["und_map1"], g["und_map2"])


def test_fisheye_and_pinhole_maps_vs_cv2_random_intrinsics(ops):
    """Closed-form 3x3 inverse on the host vs OpenCV's SVD inverse, many intrinsics."""
    rng = np.random.default_rng(11)
    for it in range(12):
        W, H = int(rng.integers(300, 900)), int(rng.integers(200, 700))
        K = np.array([[rng.uniform(200, 600), 0, W / 2 + rng.uniform(-30, 30)],
                      [0, rng.uniform(200, 600), H / 2 + rng.uniform(-30, 30)], [0, 0, 1.0]])
        D = rng.uniform(-0.03, 0.03, (4, 1))
        P = C.dst_camera_matrix(K, W, H, rng.uniform(0.4, 1.2), 1, rng.uniform(-20, 20), rng.uniform(-20, 20))
        a = ops.fisheye_init_undistort_rectify_map(K, D, P, (W, H))
        b = C.undistort_maps(K, D, P, W, H)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), it
        D5 = np.array([[rng.uniform(-0.3, 0.1), rng.uniform(-0.05, 0.1), rng.uniform(-1e-3, 1e-3),
                        rng.uniform(-1e-3, 1e-3), rng.uniform(-0.02, 0.02)]])
        a = ops.init_undistort_rectify_map(K, D5, P, (W, H))
        b = C.pinhole_maps(K, D5, P, W, H)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), it


def test_incalib_maps_golden(ops, fx):
    K, D, _ = fx.calib["front"]
    m1, m2 = ops.fisheye_init_undistort_rectify_map(K, D, C.dst_camera_matrix(K, 1280, 1024, 0.5, 1), (1280, 1024))
    gi = fx.gold["incalib_fisheye_raw0"]
    assert (h16(m1), h16(m2)) == (gi["map1"], gi["map2"])
    Kn = K * np.array([[2.0], [2.0], [1.0]])
    m1, m2 = ops.init_undistort_rectify_map(Kn, fx.D5, C.dst_camera_matrix(Kn, 1280, 1024, 0.5, 1), (1280, 1024))
    gn = fx.gold["incalib_normal_raw0"]
    assert (h16(m1), h16(m2)) == (gn["map1"], gn["map2"])


# ------------------------------------------------------------------ K3
def test_remap_fixture_maps(ops, fx):
    g = fx.geometry()
    for n in ("front", "right"):
        rc = C.RefCamera(*fx.calib[n], g)
        img = fx.img(n)
        assert h16(ops.remap(img, *rc.bev_maps)) == fx.gold["camera"][n]["raw2bev"]
        assert h16(ops.remap(img, *rc.undistort_maps)) == fx.gold["camera"][n]["undistort"]
        assert (ops.remap(img, *rc.bev_maps, interpolation=ops.INTER_NEAREST)
                == cv2.remap(img, *rc.bev_maps, cv2.INTER_NEAREST)).all()
        assert (ops.remap(img, rc.bev_maps[0], None, interpolation=ops.INTER_NEAREST)
                == cv2.remap(img, rc.bev_maps[0], None, cv2.INTER_NEAREST)).all()


def test_remap_random_maps_border_channels_strides(ops):
    rng = np.random.default_rng(5)
    for ch in (1, 3, 4):
        shape = (97, 131) if ch == 1 else (97, 131, ch)
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        m1 = rng.integers(-5, 140, (60, 75, 2)).astype(np.int16)
        m2 = rng.integers(0, 1024, (60, 75)).astype(np.uint16)
        assert (ops.remap(src, m1, m2) == cv2.remap(src, m1, m2, cv2.INTER_LINEAR)).all()
        assert (ops.remap(src, m1, m2, ops.INTER_NEAREST) == cv2.remap(src, m1, m2, cv2.INTER_NEAREST)).all()
    big = rng.integers(0, 256, (50, 200, 3), dtype=np.uint8)
    view = big[:, 10:90]                      # row stride > row bytes
    m1 = rng.integers(-2, 82, (33, 41, 2)).astype(np.int16)
    m2 = rng.integers(0, 1024, (33, 41)).astype(np.uint16)
    assert (ops.remap(view, m1, m2) == cv2.remap(np.ascontiguousarray(view), m1, m2, cv2.INTER_LINEAR)).all()
    # extreme map values (int16 range) sample the zero border
    m1 = np.array([[[-32768, -32768], [32767, 32767], [0, 49], [79, 0]]], np.int16)
    m2 = np.array([[1023, 0, 33, 1000]], np.uint16)
    assert (ops.remap(view, m1, m2) == cv2.remap(np.ascontiguousarray(view), m1, m2, cv2.INTER_LINEAR)).all()


# ------------------------------------------------------------------ undistortion (a6)
@pytest.mark.parametrize("fused", [False, True])
def test_undistort_golden(ops, fx, fused):
    for n in NAMES:
        K, D, _ = fx.calib[n]
        u = ops.Undistorter(K, D, C.dst_camera_matrix(K, 1280, 1024, 1, 2), (2560, 2048), fused=fused)
        assert h16(u(fx.img(n))) == fx.gold["camera"][n]["undistort"]
    K, D, _ = fx.calib["front"]
    u = ops.Undistorter(K, D, C.dst_camera_matrix(K, 1280, 1024, 1, 1), (1280, 1024), fused=fused)
    assert h16(u(fx.img("front"))) == fx.gold["tools_undistort_front"]
    m1, m2 = u.maps()
    ref = C.undistort_maps(K, D, C.dst_camera_matrix(K, 1280, 1024, 1, 1), 1280, 1024)
    assert (m1 == ref[0]).all() and (m2 == ref[1]).all()
    assert (u(fx.img("front"), ops.INTER_NEAREST) == cv2.remap(fx.img("front"), *ref, cv2.INTER_NEAREST)).all()
    gray = cv2.cvtColor(fx.img("front"), cv2.COLOR_BGR2GRAY)
    assert (u(gray) == cv2.remap(gray, *ref, cv2.INTER_LINEAR)).all()


def test_incalibrator_dropin(fx):
    """configs[0] of BASELINE.json: single front camera undistort, 1280x1024 and 640x480."""
    from cameracalibration_b200.IntrinsicCalibration import InCalibrator
    K, D, _ = fx.calib["front"]
    a = InCalibrator.get_args()
    a.FRAME_WIDTH, a.FRAME_HEIGHT, a.FOCAL_SCALE, a.SIZE_SCALE = 1280, 1024, 0.5, 1
    cal = InCalibrator("fisheye")
    cal.camera.data.camera_mat, cal.camera.data.dist_coeff = K, D
    cal.camera._get_undistort_maps()
    assert h16(cal.undistort(fx.img("raw0"))) == fx.gold["incalib_fisheye_raw0"]["undistort"]
    assert h16(cal.camera.data.map1) == fx.gold["incalib_fisheye_raw0"]["map1"]
    a.FRAME_WIDTH, a.FRAME_HEIGHT = 640, 480
    cal2 = InCalibrator("fisheye")
    cal2.set_calibration(np.diag([0.5, 480 / 1024, 1.0]) @ K, D)
    small = cv2.resize(fx.img("raw0"), (640, 480), interpolation=cv2.INTER_LINEAR)
    assert h16(cal2.undistort(small)) == fx.gold["incalib_fisheye_raw0_640x480"]["undistort"]
    a.FRAME_WIDTH, a.FRAME_HEIGHT = 1280, 1024
    caln = InCalibrator("normal")
    caln.set_calibration(K * np.array([[2.0], [2.0], [1.0]]), fx.D5)
    assert h16(caln.undistort(fx.img("raw0"))) == fx.gold["incalib_normal_raw0"]["undistort"]
    a.FOCAL_SCALE = 0.5


def test_tools_undistort_cli(fx, tmp_path):
    from cameracalibration_b200.Tools import undistort as T
    (tmp_path / "in").mkdir()
    (tmp_path / "out").mkdir()
    K, D, _ = fx.calib["front"]
    np.save(tmp_path / "K.npy", K)
    np.save(tmp_path / "D.npy", D)
    cv2.imwrite(str(tmp_path / "in" / "front.png"), fx.img("front"))
    for fused in ("0", "1"):
        T.main(["-path_read", str(tmp_path / "in") + "/", "-path_save", str(tmp_path / "out") + "/", "-path_k",
                str(tmp_path / "K.npy"), "-path_d", str(tmp_path / "D.npy"), "-srcformat", "png", "-dstformat", "png",
                "-quality", "1", "-fused", fused])
        assert h16(cv2.imread(str(tmp_path / "out" / "front.png"))) == fx.gold["tools_undistort_front"]
    with pytest.raises(Exception, match="Camera K File Path not exist"):
        T.main(["-path_k", str(tmp_path / "nope.npy")])


# ------------------------------------------------------------------ K4 / K2
def test_warp_perspective(ops, fx):
    H = fx.calib["back"][2]
    src = fx.img("src_back")
    assert h16(ops.warp_perspective(src, H, (1000, 1000))) == fx.gold["excalib_warp_back"]
    assert (ops.warp_perspective(src, H, (1000, 1000), ops.INTER_NEAREST)
            == cv2.warpPerspective(src, H, (1000, 1000), flags=cv2.INTER_NEAREST)).all()
    gray = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
    assert (ops.warp_perspective(gray, H, (777, 333)) == cv2.warpPerspective(gray, H, (777, 333))).all()
    rng = np.random.default_rng(3)
    for _ in range(6):   # random homographies incl. strong perspective; dst wider than one 64-px block
        Hr = np.eye(3) + rng.normal(0, [[0.3, 0.3, 60], [0.3, 0.3, 60], [4e-4, 4e-4, 0]])
        img = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
        assert (ops.warp_perspective(img, Hr, (300, 200)) == cv2.warpPerspective(img, Hr, (300, 200))).all()
    from cameracalibration_b200.ExtrinsicCalibration import ExCalibrator
    ex = ExCalibrator()
    ex.src_img, ex.homography, ex.dst_img = src, H, np.zeros((1000, 1000, 3), np.uint8)
    assert h16(ex.warp()) == fx.gold["excalib_warp_back"]
    with pytest.raises(Exception):
        ExCalibrator().warp()


def test_warp_maps_and_fused_lut(ops, fx):
    g = fx.geometry()
    e, _ = _engine(ops, fx, g, blend=False, calib=fx.calib)
    for i, n in enumerate(NAMES):
        b1, b2 = e.get_maps(i)                       # fused K1+K2 (no 2560x2048 intermediate)
        gold = fx.gold["camera"][n]
        assert (h16(b1), h16(b2)) == (gold["bev_map1"], gold["bev_map2"]), n
    rc = C.RefCamera(*fx.calib["left"], g)
    o1, o2 = ops.warp_perspective_maps(*rc.undistort_maps, fx.calib["left"][2], (1000, 1000))   # K2 on planes
    assert (o1 == rc.bev_maps[0]).all() and (o2 == rc.bev_maps[1]).all()


# ------------------------------------------------------------------ masks
def test_blend_masks_gpu(ops, fx):
    for (BW, BH, CW, CH) in [(1000, 1000, 250, 400), (777, 900, 194, 360), (203, 177, 51, 77)]:
        e = ops.BevEngine(1, (64, 64), (BW, BH))
        polys = np.stack([R.fill_poly(BW, BH, R.blend_polygon(n, BW, BH, CW, CH)) for n in NAMES])
        L = R.blend_lines(BW, BH, CW, CH)
        lines = np.stack([L[k] for k in ["FL", "FR", "BL", "BR", "LF", "LB", "RF", "RB"]])
        out = e.blend_masks(polys, lines)
        for i, n in enumerate(NAMES):
            assert (out[i] == R.blend_mask(n, BW, BH, CW, CH)).all(), (BW, n)
        if BW == 1000:
            assert [h16(out[i]) for i in range(4)] == [fx.gold["mask_blend"][n] for n in NAMES]


# ------------------------------------------------------------------ the fused BEV path
@pytest.mark.parametrize("blend", [False, True])
@pytest.mark.parametrize("balance", [False, True])
def test_bev_native_golden(ops, fx, blend, balance):
    g = fx.geometry()
    e, _ = _engine(ops, fx, g, blend, calib=fx.calib)
    gold = fx.gold["native"][f"blend{int(blend)}_balance{int(balance)}"]
    F = fx.frames()
    assert h16(e.run([F], None, balance)[0]) == gold["nocar"]
    assert h16(e.run([F], fx.car(), balance)[0]) == gold["car"]


@pytest.mark.parametrize("key,FW,FH,BW,BH,blend,balance,car", [
    ("cfg2_1280x960_1000_plain", 1280, 960, 1000, 1000, False, False, False),
    ("cfg3_1920x1080_1200_blend_balance", 1920, 1080, 1200, 1200, True, True, False),
    ("cfg3_1920x1080_1200_blend_balance_car", 1920, 1080, 1200, 1200, True, True, True),
    ("1920x1080_1200_plain", 1920, 1080, 1200, 1200, False, False, False),
    ("cfg4_1920x1080_1000_blend", 1920, 1080, 1000, 1000, True, False, False),
    ("odd_1000x750_777x900_blend_balance_car", 1000, 750, 777, 900, True, True, True),
    ("cfg5size_3840x2160_2000_blend_balance", 3840, 2160, 2000, 2000, True, True, False),
])
def test_bev_configs_golden(ops, fx, key, FW, FH, BW, BH, blend, balance, car):
    g = fx.geometry(FW, FH, BW, BH)
    e, _ = _engine(ops, fx, g, blend)
    out = e.run([fx.frames(FW, FH)], fx.car(BW, BH) if car else None, balance)[0]
    assert h16(out) == fx.gold["cfg"][key]


def test_bev_batch_vs_live_oracle(ops, fx):
    """cfg4 shape: a batch of distinct frame-sets (fixture + seeded noise, all-random, all-255)
    through one call, against the cv2 call sequence on each set."""
    FW, FH, BW, BH = 1920, 1080, 1000, 1000
    g = fx.geometry(FW, FH, BW, BH)
    for blend, balance in ((True, False), (True, True), (False, True)):
        e, masks = _engine(ops, fx, g, blend)
        ref = C.RefBev(fx.scaled_calib(g), g, blend, balance, masks=masks)
        sets = [fx.perturbed_frames(FW, FH, i) for i in range(3)]
        rng = np.random.default_rng(99)
        sets.append([rng.integers(0, 256, (FH, FW, 3), dtype=np.uint8) for _ in range(4)])
        sets.append([np.full((FH, FW, 3), 255, np.uint8) for _ in range(4)])
        sets.append([np.zeros((FH, FW, 3), np.uint8) for _ in range(4)])
        car = fx.car(BW, BH)
        out = e.run(sets, car, balance)
        for i, s in enumerate(sets):
            if balance and i == 5:
                continue   # all-zero frames: the reference divides by a zero channel mean (NaN gains)
            assert (out[i] == ref(*s, car)).all(), (blend, balance, i)


def test_bev_border_and_wrong_size_frames(ops, fx):
    """Injected LUT with out-of-frame taps (BORDER_CONSTANT) and frames of another size."""
    g = fx.geometry(320, 240, 200, 160, CW=50, CH=64)
    rng = np.random.default_rng(8)
    e = ops.BevEngine(4, (g.FW, g.FH), (g.BW, g.BH))
    maps, masks = [], []
    for i, n in enumerate(NAMES):
        m1 = np.stack([rng.integers(-3, g.FW + 3, (g.BH, g.BW)), rng.integers(-3, g.FH + 3, (g.BH, g.BW))], -1).astype(np.int16)
        m2 = rng.integers(0, 1024, (g.BH, g.BW)).astype(np.uint16)
        mk = R.blend_mask(n, g.BW, g.BH, g.CW, g.CH)
        e.set_maps(i, m1, m2)
        e.set_mask(i, mk)
        maps.append((m1, m2))
        masks.append(mk)
    F = [rng.integers(0, 256, (g.FH, g.FW, 3), dtype=np.uint8) for _ in range(4)]
    for balance in (False, True):
        want = R.bev_generate(F, maps, masks, True, balance, None)
        assert (e.run([F], None, balance)[0] == want).all(), balance
    small = [f[:200, :300] for f in F]                       # reference: cv2.remap samples zeros outside
    want = R.bev_generate([np.ascontiguousarray(s) for s in small], maps, masks, True, False, None)
    assert (e.run([small])[0] == want).all()


def test_standalone_helpers(fx):
    from cameracalibration_b200.SurroundBirdEyeView import surroundBEV as S
    F = fx.frames()
    for a, b in zip(S.luminance_balance(F), C.luminance_balance(F)):
        assert (a == b).all()
    odd = [np.ascontiguousarray(f[:37, :1279]) for f in F]   # rows with a < 32 px tail
    for a, b in zip(S.luminance_balance(odd), C.luminance_balance(odd)):
        assert (a == b).all()
    cv = np.random.default_rng(0).integers(0, 256, (300, 201, 3), dtype=np.uint8)
    assert (S.color_balance(cv) == C.color_balance(cv.copy())).all()


def test_dropin_bevgenerator_classes(fx, tmp_path):
    """The reference's public classes, data-dir convention included (K/D/H .npy per camera)."""
    from cameracalibration_b200.SurroundBirdEyeView import surroundBEV as S
    for n in NAMES:
        (tmp_path / n).mkdir()
        for k, m in zip("KDH", fx.calib[n]):
            np.save(tmp_path / n / f"camera_{n}_{k}.npy", m)
    a = S.BevGenerator.get_args()
    a.DATA_DIR = str(tmp_path)
    a.CAR_WIDTH, a.CAR_HEIGHT = 250, 400
    F = fx.frames()
    bev = S.BevGenerator()
    assert h16(bev(*F)) == fx.gold["native"]["blend0_balance0"]["nocar"]
    assert h16(bev(*F, fx.car())) == fx.gold["native"]["blend0_balance0"]["car"]
    bev = S.BevGenerator(blend=True, balance=True)
    assert h16(bev(*F, fx.car())) == fx.gold["native"]["blend1_balance1"]["car"]
    for n, mk in zip(NAMES, bev.masks):
        assert h16(mk.mask) == fx.gold["mask_blend"][n]
    assert bev.masks[0].weight.dtype == np.float32 and bev.masks[0].weight.shape == (1000, 1000, 3)
    cam = bev.cameras[2]
    gold = fx.gold["camera"]["left"]
    assert h16(cam.raw2bev(F[2])) == gold["raw2bev"]
    u = cam.undistort(F[2])
    assert h16(u) == gold["undistort"]
    assert h16(cam.warp_homography(u)) == gold["warp_undistort"]
    assert (h16(cam.bev_maps[0]), h16(cam.undistort_maps[1])) == (gold["bev_map1"], gold["und_map2"])
    # the reference's own way to the BEV maps: warp_homography applied to each undistort-map plane
    assert h16(cam.warp_homography(cam.undistort_maps[0])) == gold["bev_map1"]
    assert h16(cam.warp_homography(cam.undistort_maps[1])) == gold["bev_map2"]
    plain = S.Mask("front")
    assert h16(plain.mask) == fx.gold["mask_plain"]["front"]
    w = cam.raw2bev(F[2])
    assert (plain(w) == cv2.bitwise_and(w, w, mask=plain.mask)).all()
    bm = bev.masks[2]
    assert (bm(w) == (w * bm.weight).astype(np.uint8)).all()
    # main.py:79-84 variant
    a.CAR_WIDTH, a.CAR_HEIGHT = 200, 350
    assert h16(S.BevGenerator(blend=True, balance=True)(*F)) == fx.gold["main_py_variant"]
    a.CAR_WIDTH, a.CAR_HEIGHT = 250, 400
    with pytest.raises(Exception, match="name should be front/back/left/right"):
        S.Mask("top")
    out = S.BevGenerator().run_batch([F, F[::-1]], fx.car())
    assert h16(out[0]) == fx.gold["native"]["blend0_balance0"]["car"] and out.shape == (2, 1000, 1000, 3)


def test_device_resident_and_camera_sharded_compose(ops, fx):
    """Device pointers in / out (torch tensors as the allocator), and the camera-sharded
    decomposition used for multi-GPU: sat-sum of per-camera partial canvases == full canvas."""
    import torch
    g = fx.geometry()
    e, _ = _engine(ops, fx, g, blend=True, calib=fx.calib)
    dev = torch.device("cuda", e.ctx.device)
    stream = torch.cuda.Stream(device=dev)
    e.ctx.set_stream(stream.cuda_stream)
    F = fx.frames()
    d_frames = [torch.from_numpy(f).to(dev) for f in F] * 2
    ptrs = torch.tensor([t.data_ptr() for t in d_frames], dtype=torch.int64, device=dev)
    car = torch.from_numpy(fx.car()).to(dev)
    out = torch.empty((2, 1000, 1000, 3), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    e.run_device(ptrs.data_ptr(), 2, out.data_ptr(), car.data_ptr(), balance=True)
    torch.cuda.synchronize()
    assert h16(out[1].cpu().numpy()) == fx.gold["native"]["blend1_balance1"]["car"]
    assert e.last_kernel_ms() > 0
    parts = []
    for lo in range(4):
        p = torch.empty((1, 1000, 1000, 3), dtype=torch.uint8, device=dev)
        e.run_device_cams(ptrs.data_ptr(), 1, lo, lo + 1, p.data_ptr())
        parts.append(p)
    full = torch.empty((1000, 1000, 3), dtype=torch.uint8, device=dev)
    e.sat_sum_device([p.data_ptr() for p in parts], 3_000_000, full.data_ptr(), car.data_ptr())
    torch.cuda.synchronize()
    assert h16(full.cpu().numpy()) == fx.gold["native"]["blend1_balance0"]["car"]
    e.ctx.set_stream(None)
    assert e.ctx.launches > 0


def test_undistorter_slots_are_owned(ops, fx):
    """A bevk_ctx has 8 undistorter slots: each live Undistorter owns one, a 9th on an explicit ctx raises, closing one
    frees its slot, and on the shared default ctx the 9th gets a context of its own; a call whose destination was sized
    for another map is refused by the library instead of writing past it."""
    import ctypes as Ct
    from cameracalibration_b200 import _lib as L
    K, D, _ = fx.calib["front"]
    P = C.dst_camera_matrix(K, 1280, 1024, 1, 1)
    ctx = L.Context(0)
    img = fx.img("front")
    us = [ops.Undistorter(K, D, P, (320 + 16 * i, 256), ctx=ctx) for i in range(8)]
    assert sorted(u.slot for u in us) == list(range(8))
    with pytest.raises(L.BevkError, match="8 undistorter slots"):
        ops.Undistorter(K, D, P, (320, 256), ctx=ctx)
    want0 = us[0](img)
    us[3].close()
    with pytest.raises(L.BevkError, match="closed"):
        us[3](img)
    u9 = ops.Undistorter(K, D, P, (640, 512), ctx=ctx)
    assert u9.slot == 3 and u9(img).shape == (512, 640, 3)
    assert (us[0](img) == want0).all()                       # nobody else's map was touched
    out = np.empty((256, 320, 3), np.uint8)
    rc = ctx.lib.bevk_undistort(ctx.h, 3, L.vptr(img), 1280, 1024, 3840, 3, L.vptr(out), 320, 256, 960, 1)
    assert rc != 0 and b"holds a 640x512 map" in ctx.lib.bevk_last_error()
    d = L.default_context()
    keep = [ops.Undistorter(K, D, P, (320, 256)) for _ in range(9)]
    assert sum(u.ctx is d for u in keep) <= 8 and any(u.ctx is not d for u in keep)
    assert all((u(img) == keep[0](img)).all() for u in keep[1:])


def test_errors_are_loud(ops, fx):
    from cameracalibration_b200 import BevkError
    e = ops.BevEngine(4, (64, 48), (40, 40))
    with pytest.raises(BevkError, match="no maps"):
        e.finalize()
    with pytest.raises(BevkError):
        ops.remap(np.zeros((4, 4, 3), np.uint8), np.zeros((2, 2, 2), np.int16), None, ops.INTER_LINEAR)
    with pytest.raises(BevkError):
        ops.remap(np.zeros((4, 4, 2), np.uint8), np.zeros((2, 2, 2), np.int16), np.zeros((2, 2), np.uint16))
    with pytest.raises(BevkError, match="out of range"):
        ops.BevEngine(9, (64, 48), (40, 40))


def test_blend_weight_all_pixel_mask_pairs(ops):
    """Every (pixel value, mask value) pair through the fused kernel: identity LUT, image
    value = column, mask = row."""
    n = 256
    e = ops.BevEngine(1, (n + 2, n + 2), (n, n))
    yy, xx = np.mgrid[0:n, 0:n]
    e.set_maps(0, np.stack([xx, yy], -1).astype(np.int16), np.zeros((n, n), np.uint16))
    mask = yy.astype(np.uint8)
    e.set_mask(0, mask)
    img = np.zeros((n + 2, n + 2, 3), np.uint8)
    img[:n, :n, 0] = xx
    img[:n, :n, 1] = 255 - xx
    img[:n, :n, 2] = (xx * 7) & 255
    got = e.run([[img]])[0]
    want = R.apply_blend(img[:n, :n], mask)
    assert (got == want).all()
    assert (ops.apply_mask(img[:n, :n], mask, blend=True) == want).all()


def test_bev_eight_cameras_wedge_masks(ops, fx):
    """BASELINE configs[4] semantics at a small size: cameras 0-3 = the scaled
    fixtures, 4-7 = the same four with H post-multiplied by a 45-degree rotation about the canvas
    centre; 8 angular wedge masks; oracle = the reference's Camera.raw2bev per camera + the N-way
    saturating compose.  Also exercises camera ranges (camera-per-GPU sharding) on 8 cameras."""
    g = fx.geometry(640, 512, 480, 480)
    calib4 = fx.scaled_calib(g)
    c, s_ = np.cos(np.pi / 4), np.sin(np.pi / 4)
    cx, cy = g.BW / 2, g.BH / 2
    Rot = np.array([[c, -s_, cx - c * cx + s_ * cy], [s_, c, cy - s_ * cx - c * cy], [0, 0, 1.0]])
    cams = [calib4[n] for n in NAMES] + [(calib4[n][0], calib4[n][1], Rot @ calib4[n][2]) for n in NAMES]
    ang = np.linspace(0, 2 * np.pi, 9)
    rad = g.BW
    masks = []
    for i in range(8):
        tri = np.array([[cx, cy], [cx + rad * np.cos(ang[i]), cy + rad * np.sin(ang[i])],
                        [cx + rad * np.cos(ang[i + 1]), cy + rad * np.sin(ang[i + 1])]]).astype(np.int32)
        masks.append(cv2.fillPoly(np.zeros((g.BH, g.BW), np.uint8), [tri], 255))
    e = ops.BevEngine(8, (g.FW, g.FH), (g.BW, g.BH))
    frames = fx.frames(g.FW, g.FH)
    frames8 = frames + [np.ascontiguousarray(f[:, ::-1]) for f in frames]
    want = np.zeros((g.BH, g.BW, 3), np.uint8)
    for i, (K, D, H) in enumerate(cams):
        e.set_camera(i, K, D, C.dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS), (int(g.FW * g.SS), int(g.FH * g.SS)), H)
        e.set_mask(i, masks[i])
        rc = C.RefCamera(K, D, H, g)
        want = R.sat_add(want, R.apply_plain(rc.raw2bev(frames8[i]), masks[i]))
    got = e.run([frames8])[0]
    assert (got == want).all()
    import torch
    dev = torch.device("cuda", e.ctx.device)
    d = [torch.from_numpy(f).to(dev) for f in frames8]
    ptrs = torch.tensor([t.data_ptr() for t in d], dtype=torch.int64, device=dev)
    parts = [torch.empty((g.BH, g.BW, 3), dtype=torch.uint8, device=dev) for _ in range(3)]
    torch.cuda.synchronize()
    for p, (lo, hi) in zip(parts, [(0, 3), (3, 6), (6, 8)]):       # ragged camera ranges
        e.run_device_cams(ptrs.data_ptr(), 1, lo, hi, p.data_ptr())
    full = torch.empty_like(parts[0])
    e.sat_sum_device([p.data_ptr() for p in parts], full.numel(), full.data_ptr())
    e.ctx.sync()
    assert (full.cpu().numpy() == want).all()


def test_bev_batch_sizes_and_host_pipeline(ops, fx):
    """Batches that exercise every grouping: 1 and 3 (single-frame units), 5 and 9 (groups of 4 with
    a ragged tail), 19 (three chunks of the two-stream host pipeline, ragged), each frame-set
    different; every canvas must equal the single-frame-set call."""
    g = fx.geometry(480, 384, 330, 350)
    e, _ = _engine(ops, fx, g, blend=True)
    base = fx.frames(g.FW, g.FH)
    rng = np.random.default_rng(3)
    sets = []
    for i in range(19):
        sets.append([np.ascontiguousarray(np.roll(f, 7 * i + 3 * c, axis=1) ^ rng.integers(0, 8, f.shape, dtype=np.uint8))
                     for c, f in enumerate(base)])
    car = fx.car(g.BW, g.BH)
    singles = [e.run([s], car)[0].copy() for s in sets]
    for nb in (1, 3, 5, 9, 19):
        for balance in (False, True):
            out = e.run(sets[:nb], car, balance)
            if not balance:
                for i in range(nb):
                    assert (out[i] == singles[i]).all(), (nb, i)
            else:
                ref = e.run([sets[nb - 1]], car, True)[0]
                assert (out[nb - 1] == ref).all(), nb


def test_bev_pinned_host_frames_zero_copy_ingest(ops, fx):
    """Page-locked frames take the span-by-span zero-copy ingest (k_fetch_spans) instead of DMA
    rectangles; results and the reported PCIe bytes must agree with the pageable path / be smaller."""
    from cameracalibration_b200 import pinned_empty
    g = fx.geometry(1280, 1024, 1000, 1000)        # pitch 3840 is a multiple of 16
    e, masks = _engine(ops, fx, g, blend=True, calib=fx.calib)
    F = fx.frames()
    sets_pageable = [[np.ascontiguousarray(np.roll(f, 11 * i, axis=0)) for f in F] for i in range(11)]
    want = e.run(sets_pageable, fx.car())
    dma_bytes = e.last_h2d_bytes()
    sets_pinned = []
    for s in sets_pageable:
        row = []
        for f in s:
            p = pinned_empty(f.shape)
            p[...] = f
            row.append(p)
        sets_pinned.append(row)
    got = e.run(sets_pinned, fx.car())
    assert (got == want).all()
    assert h16(got[0]) == fx.gold["native"]["blend1_balance0"]["car"]
    assert 0 < e.last_h2d_bytes() < dma_bytes < 11 * 4 * 1280 * 1024 * 3
    bal = e.run(sets_pinned[:2], fx.car(), balance=True)      # BALANCE needs whole frames: DMA path
    assert h16(bal[0]) == fx.gold["native"]["blend1_balance1"]["car"]
    assert e.last_h2d_bytes() == 2 * 4 * 1280 * 1024 * 3


def test_bev_odd_frame_width_and_strided_frames(ops, fx):
    """Frame widths whose row pitch is not a multiple of 4 (every LUT entry takes the per-tap checked
    path, k_vsum its unaligned path) and frames that are strided views of larger host arrays."""
    g = fx.geometry(333, 250, 203, 177)
    e, masks = _engine(ops, fx, g, blend=True)
    ref = C.RefBev(fx.scaled_calib(g), g, True, False, masks=masks)
    F = fx.frames(g.FW, g.FH)
    car = fx.car(g.BW, g.BH)
    for balance in (False, True):
        ref.balance = balance
        assert (e.run([F], car, balance)[0] == ref(*F, car)).all(), balance
    g2 = fx.geometry(640, 480, 300, 300)
    e2, masks2 = _engine(ops, fx, g2, blend=False)
    ref2 = C.RefBev(fx.scaled_calib(g2), g2, False, False, masks=masks2)
    F2 = fx.frames(g2.FW, g2.FH)
    big = [np.zeros((500, 700, 3), np.uint8) for _ in F2]
    views = []
    for b, f in zip(big, F2):
        b[10:490, 30:670] = f
        views.append(b[10:490, 30:670])          # row stride 2100 B > 1920 B row
    assert (e2.run([views])[0] == ref2(*F2)).all()
    assert (e2.run([views, F2])[1] == ref2(*F2)).all() if views[0].strides[0] == F2[0].strides[0] else True


def test_bev_nearest_neighbour_mode(ops, fx):
    """INTER_NEAREST through the fused engine (LUT compiled with OpenCV's fixed-point NN rule):
    bit-exact against cv2.remap(..., INTER_NEAREST) per camera + the reference's mask / compose."""
    g = fx.geometry()
    for blend in (False, True):
        e, masks = _engine(ops, fx, g, blend, calib=fx.calib)
        e.set_interpolation(ops.INTER_NEAREST)
        F = fx.frames()
        want = np.zeros((g.BH, g.BW, 3), np.uint8)
        for i, n in enumerate(NAMES):
            rc = C.RefCamera(*fx.calib[n], g)
            w = cv2.remap(F[i], *rc.bev_maps, interpolation=cv2.INTER_NEAREST)
            want = R.sat_add(want, R.apply_blend(w, masks[i]) if blend else R.apply_plain(w, masks[i]))
        want = R.sat_add(want, fx.car())
        assert (e.run([F], fx.car())[0] == want).all(), blend
        e.set_interpolation(ops.INTER_LINEAR)
        gold = fx.gold["native"][f"blend{int(blend)}_balance0"]["car"]
        assert h16(e.run([F], fx.car())[0]) == gold


def test_bev_frames_already_on_the_gpu(ops, fx):
    """frame-sets that already live in HBM (torch tensors here, any __cuda_array_interface__
    object in general) go through bevk_bev_run_frames: same bytes as the host path, table re-uploaded only
    when it changes, wrong shapes / layouts rejected."""
    import torch
    from cameracalibration_b200 import _lib as L
    g = fx.geometry(1280, 1024, 1000, 1000)
    e, masks = _engine(ops, fx, g, blend=True, calib=fx.calib)
    F = fx.frames()
    host_sets = [[np.ascontiguousarray(np.roll(f, 7 * i, axis=1)) for f in F] for i in range(5)]
    want = e.run(host_sets, fx.car())
    want_bal = e.run(host_sets, fx.car(), balance=True)
    dev = torch.device("cuda", e.ctx.device)
    stream = torch.cuda.Stream(device=dev)
    d_all = torch.from_numpy(np.stack([np.stack(s) for s in host_sets])).to(dev)          # [5][4][FH][FW][3]
    d_car = torch.from_numpy(fx.car()).to(dev)
    torch.cuda.synchronize()
    out = e.run_cuda(d_all, d_car, stream=stream.cuda_stream)
    stream.synchronize()
    assert isinstance(out, torch.Tensor) and out.shape == (5, 1000, 1000, 3)
    assert (out.cpu().numpy() == want).all()
    assert h16(out[0].cpu().numpy()) == fx.gold["native"]["blend1_balance0"]["car"]
    # same buffers, new contents: the cached table is reused and must still be right
    d_all.copy_(d_all.flip(0)); torch.cuda.synchronize()
    out2 = torch.empty_like(out)
    assert e.run_cuda(d_all, d_car, out=out2) is out2
    stream.synchronize()
    assert (out2.cpu().numpy() == want[::-1]).all()
    # nested lists of separately allocated frames (a different table), balance, no car
    nested = [[torch.from_numpy(f).to(dev) for f in s] for s in host_sets[:3]]
    torch.cuda.synchronize()
    out3 = e.run_cuda(nested, d_car, balance=True)
    stream.synchronize()
    assert (out3.cpu().numpy() == want_bal[:3]).all()
    e.ctx.set_stream(None)        # back to the ctx's own stream: the cached table is dropped and re-uploaded
    out4 = e.run_cuda(nested)
    e.ctx.sync()
    assert (out4.cpu().numpy() == e.run(host_sets[:3])).all()
    with pytest.raises(L.BevkError, match="must be uint8"):
        e.run_cuda(d_all.to(torch.float32))
    skewed = nested[0][0].permute(1, 0, 2).contiguous().permute(1, 0, 2)      # right shape, wrong strides
    with pytest.raises(L.BevkError, match="C-contiguous"):
        e.run_cuda([[skewed] + nested[0][1:]])
    with pytest.raises(L.BevkError, match="got"):
        e.run_cuda(d_all[:, :3].contiguous())
    with pytest.raises(L.BevkError, match="expected 4"):
        e.run_cuda([nested[0][:3]])
    with pytest.raises(L.BevkError, match="__cuda_array_interface__"):
        e.run_cuda([[f for f in host_sets[0]]])
