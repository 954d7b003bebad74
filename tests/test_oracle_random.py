"""Seeded random cases that pin the NumPy restatement (oracle/restate.py) to live cv2 beyond the
reference's fixtures: random intrinsics / distortion, random homographies with strong perspective,
random maps with out-of-range taps, every channel count.  CPU only, a few seconds."""
import cv2
import numpy as np
import pytest

from oracle import cv2_path as C
from oracle import restate as R


@pytest.mark.parametrize("seed", range(6))
def test_random_intrinsics_maps(seed):
    rng = np.random.default_rng(100 + seed)
    W, H = int(rng.integers(64, 400)), int(rng.integers(48, 300))
    K = np.array([[rng.uniform(80, 500), 0, W / 2 + rng.uniform(-20, 20)],
                  [0, rng.uniform(80, 500), H / 2 + rng.uniform(-20, 20)], [0, 0, 1.0]])
    D = rng.uniform(-0.05, 0.05, (4, 1))
    P = C.dst_camera_matrix(K, W, H, rng.uniform(0.3, 1.5), rng.choice([1, 2]), rng.uniform(-9, 9), rng.uniform(-9, 9))
    w2, h2 = int(W * rng.choice([1, 2])), int(H * rng.choice([1, 2]))
    a, b = R.fisheye_map(K, D, P, w2, h2), C.undistort_maps(K, D, P, w2, h2)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    d = R.fisheye_map(K, D, P, w2, h2, running=False)     # closed form (what the GPU evaluates)
    assert (d[0] == b[0]).all() and (d[1] == b[1]).all()
    D5 = np.array([[rng.uniform(-0.3, 0.1), rng.uniform(-0.05, 0.1), rng.uniform(-1e-3, 1e-3), rng.uniform(-1e-3, 1e-3),
                    rng.uniform(-0.02, 0.02)]])
    a, b = R.pinhole_map(K, D5, P, w2, h2), C.pinhole_maps(K, D5, P, w2, h2)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


@pytest.mark.parametrize("seed", range(6))
def test_random_homographies(seed):
    rng = np.random.default_rng(200 + seed)
    Hm = np.eye(3) + rng.normal(0, [[0.3, 0.3, 40], [0.3, 0.3, 40], [6e-4, 6e-4, 0]])
    for ch in (1, 3, 4):
        shape = (150, 210) if ch == 1 else (150, 210, ch)
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        dw, dh = int(rng.integers(70, 260)), int(rng.integers(50, 200))     # wider than one 64-px block
        assert (R.warp_perspective_u8(img, Hm, dw, dh) == cv2.warpPerspective(img, Hm, (dw, dh))).all()
        assert (R.warp_perspective_u8(img, Hm, dw, dh, nearest=True)
                == cv2.warpPerspective(img, Hm, (dw, dh), flags=cv2.INTER_NEAREST)).all()
    # the map planes themselves through warpPerspective (Camera.get_bev_maps)
    m1 = rng.integers(-3, 400, (90, 120, 2)).astype(np.int16)
    m2 = rng.integers(0, 1024, (90, 120)).astype(np.uint16)
    o1, o2 = R.warp_perspective_maps(m1, m2, Hm, 100, 80)
    assert (o1 == cv2.warpPerspective(m1, Hm, (100, 80))).all()
    assert (o2 == cv2.warpPerspective(m2, Hm, (100, 80))).all()


@pytest.mark.parametrize("seed", range(4))
def test_random_remap_and_compose(seed):
    rng = np.random.default_rng(300 + seed)
    src = rng.integers(0, 256, (int(rng.integers(20, 90)), int(rng.integers(20, 120)), 3), dtype=np.uint8)
    dh, dw = int(rng.integers(10, 80)), int(rng.integers(10, 100))
    m1 = np.stack([rng.integers(-4, src.shape[1] + 4, (dh, dw)), rng.integers(-4, src.shape[0] + 4, (dh, dw))], -1).astype(np.int16)
    m2 = rng.integers(0, 1024, (dh, dw)).astype(np.uint16)
    assert (R.remap_linear(src, m1, m2) == cv2.remap(src, m1, m2, cv2.INTER_LINEAR)).all()
    assert (R.remap_nearest(src, m1, m2) == cv2.remap(src, m1, m2, cv2.INTER_NEAREST)).all()
    a = rng.integers(0, 256, (dh, dw, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (dh, dw, 3), dtype=np.uint8)
    mask = rng.integers(0, 256, (dh, dw), dtype=np.uint8)
    assert (R.sat_add(a, b) == cv2.add(a, b)).all()
    assert (R.apply_plain(a, mask) == cv2.bitwise_and(a, a, mask=mask)).all()
    w = (np.repeat(mask[:, :, None], 3, axis=2) / 255.0).astype(np.float32)
    assert (R.apply_blend(a, mask) == (a * w).astype(np.uint8)).all()
    assert (R.color_balance(a) == C.color_balance(a.copy())).all()


def test_invert3_is_cv2_invert():
    """cv2.warpPerspective inverts H with cv::invert's closed 3x3 form; LAPACK's inverse differs in the last bits
    (found by tests/test_host_math.py: about one 1/32-px coordinate per million moved)."""
    rng = np.random.default_rng(400)
    differs = 0
    for _ in range(300):
        Hm = np.eye(3) + rng.normal(0, [[0.3, 0.3, 40], [0.3, 0.3, 40], [6e-4, 6e-4, 0]])
        assert (R.invert3(Hm) == cv2.invert(Hm)[1]).all()
        differs += int((np.linalg.inv(Hm) != cv2.invert(Hm)[1]).any())
    assert differs > 0                                   # the distinction is real
    assert (R.invert3(np.zeros((3, 3))) == 0).all()      # singular: zeros, as cv::invert leaves them
