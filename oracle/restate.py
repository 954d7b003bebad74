"""CPU restatement oracle (NumPy) of the surround-BEV hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``cameracalibration_b200`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may.  It restates, in plain NumPy, the arithmetic that the
reference (dyfcalid/CameraCalibration, pure Python) obtains from its
third-party dependency OpenCV (unpinned by the reference: ``README.md:14``
"opencv(>=3.4.2)"; pinned here to opencv-python-headless 4.13.0.92, the build in
this image).  Every function cites the reference call site (file:line relative
to /root/reference) it follows and the SURVEY.md appendix item that specifies the
arithmetic.

Parity status: PINNED.  ``tests/test_oracle.py`` and ``tests/test_oracle_random.py`` check every
function here bit-for-bit against (a) live ``cv2`` calls, (b) the unmodified
reference classes imported from /root/reference (when that tree exists), and (c)
the committed golden hashes in ``tests/golden/golden.json`` produced by
``oracle/gen_golden.py`` from the unmodified reference.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS  # 32


# ----------------------------------------------------------------------------
# A1  cv2.fisheye.initUndistortRectifyMap(K, D, I, P, (W, H), CV_16SC2)
#     reference call sites: SurroundBirdEyeView/surroundBEV.py:98-103,
#     Tools/undistort.py:50-52, IntrinsicCalibration/intrinsicCalib.py:98-103
# ----------------------------------------------------------------------------
def _ray_terms(iR: np.ndarray, W: int, H: int, running: bool):
    """_x,_y,_w per destination pixel.  ``running`` reproduces OpenCV's loop
    carried sums (_x += iR00 per column); otherwise the closed form j*iR00+..."""
    i = np.arange(H, dtype=np.float64)[:, None]
    if running:
        def chain(a, b, c):  # start = i*b + c, then += a, sequentially
            start = i * b + c
            steps = np.full((H, W), a, dtype=np.float64)
            steps[:, 0] = start[:, 0]
            return np.cumsum(steps, axis=1)  # numpy cumsum adds left-to-right
        return (chain(iR[0, 0], iR[0, 1], iR[0, 2]),
                chain(iR[1, 0], iR[1, 1], iR[1, 2]),
                chain(iR[2, 0], iR[2, 1], iR[2, 2]))
    j = np.arange(W, dtype=np.float64)[None, :]
    return (j * iR[0, 0] + (i * iR[0, 1] + iR[0, 2]),
            j * iR[1, 0] + (i * iR[1, 1] + iR[1, 2]),
            j * iR[2, 0] + (i * iR[2, 1] + iR[2, 2]))


def _quantise_maps(u: np.ndarray, v: np.ndarray, simd_saturate: bool = False):
    """iu=cvRound(u*32) saturated to int32, then split (SURVEY A1 last 2 lines).
    ``simd_saturate``: cv2.initUndistortRectifyMap's (pinhole) vector body packs map1 with
    signed saturation for columns j < W - W%8 while its scalar row tail wraps like a C cast
    (measured against cv2 4.13.0; only visible when |u| or |v| >= 32768 px)."""
    def sat_i32(a):
        a = np.rint(a * TAB)
        a = np.where(np.isnan(a), 0.0, a)
        return np.clip(a, -2147483648.0, 2147483647.0).astype(np.int64)
    iu, iv = sat_i32(u), sat_i32(v)
    mx, my = iu >> INTER_BITS, iv >> INTER_BITS
    if simd_saturate:
        W = u.shape[1]
        body = np.arange(W) < W - W % 8
        mx = np.where(body, np.clip(mx, -32768, 32767), mx)
        my = np.where(body, np.clip(my, -32768, 32767), my)
    map1 = np.stack([mx, my], axis=-1).astype(np.int16)
    map2 = ((iv & (TAB - 1)) * TAB + (iu & (TAB - 1))).astype(np.uint16)
    return map1, map2


def fisheye_map(K, D, P, W: int, H: int, running: bool = True):
    """Equidistant fisheye inverse map, 1/32-px fixed point."""
    K = np.asarray(K, np.float64)
    k1, k2, k3, k4 = (float(t) for t in np.asarray(D, np.float64).ravel()[:4])
    iR = np.linalg.inv(np.asarray(P, np.float64))
    _x, _y, _w = _ray_terms(iR, W, H, running)
    with np.errstate(all="ignore"):
        x = _x / _w
        y = _y / _w
        r = np.sqrt(x * x + y * y)
        th = np.arctan(r)
        t2 = th * th
        t4 = t2 * t2
        t6 = t4 * t2
        t8 = t4 * t4
        thd = th * (1 + k1 * t2 + k2 * t4 + k3 * t6 + k4 * t8)
        s = np.where(r == 0, 1.0, thd / r)
        u = K[0, 0] * x * s + K[0, 2]
        v = K[1, 1] * y * s + K[1, 2]
        neg = _w <= 0  # OpenCV 4.x branch: rays behind the camera go to +-inf
        u = np.where(neg, np.where(_x > 0, -np.inf, np.inf), u)
        v = np.where(neg, np.where(_y > 0, -np.inf, np.inf), v)
    return _quantise_maps(u, v)


# A11  cv2.initUndistortRectifyMap (pinhole, 5 coeffs) -- intrinsicCalib.py:158-163
def pinhole_map(K, D5, P, W: int, H: int):
    K = np.asarray(K, np.float64)
    d = np.zeros(5)
    dd = np.asarray(D5, np.float64).ravel()
    d[:min(5, dd.size)] = dd[:5]
    k1, k2, p1, p2, k3 = d
    ir = invert3(P)     # cv::Matx33d::inv(DECOMP_LU): the closed 3x3 form
    _x, _y, _w = _ray_terms(ir, W, H, running=True)
    w = 1.0 / _w
    x = _x * w
    y = _y * w
    x2 = x * x
    y2 = y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    u = K[0, 0] * xd + K[0, 2]
    v = K[1, 1] * yd + K[1, 2]
    return _quantise_maps(u, v, simd_saturate=True)


# ----------------------------------------------------------------------------
# A2  cv2.remap(src, map1:16SC2, map2:16UC1, INTER_LINEAR | INTER_NEAREST),
#     BORDER_CONSTANT 0 -- surroundBEV.py:111,117; undistort.py:66;
#     intrinsicCalib.py:195
# ----------------------------------------------------------------------------
def _taps(src: np.ndarray, sy: np.ndarray, sx: np.ndarray):
    Hs, Ws = src.shape[:2]
    ok = (sx >= 0) & (sx < Ws) & (sy >= 0) & (sy < Hs)
    px = src[np.clip(sy, 0, Hs - 1), np.clip(sx, 0, Ws - 1)].astype(np.int64)
    if px.ndim == ok.ndim + 1:
        ok = ok[..., None]
    return np.where(ok, px, 0)


def remap_linear(src: np.ndarray, map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    sx = map1[..., 0].astype(np.int64)
    sy = map1[..., 1].astype(np.int64)
    fx = (map2 & (TAB - 1)).astype(np.int64)
    fy = ((map2 >> INTER_BITS) & (TAB - 1)).astype(np.int64)
    if src.ndim == 3:
        fx = fx[..., None]
        fy = fy[..., None]
    acc = ((TAB - fx) * (TAB - fy) * _taps(src, sy, sx)
           + fx * (TAB - fy) * _taps(src, sy, sx + 1)
           + (TAB - fx) * fy * _taps(src, sy + 1, sx)
           + fx * fy * _taps(src, sy + 1, sx + 1) + 512) >> 10
    return acc.astype(np.uint8)


def remap_nearest(src: np.ndarray, map1: np.ndarray, map2: np.ndarray | None) -> np.ndarray:
    """INTER_NEAREST with fixed-point maps.  OpenCV's NNDeltaTab is inverted: a
    fraction < 16 picks the +1 neighbour."""
    sx = map1[..., 0].astype(np.int64)
    sy = map1[..., 1].astype(np.int64)
    if map2 is not None:
        fx = (map2 & (TAB - 1)).astype(np.int64)
        fy = ((map2 >> INTER_BITS) & (TAB - 1)).astype(np.int64)
        sx = sx + (fx < 16)
        sy = sy + (fy < 16)
    return _taps(src, sy, sx).astype(np.uint8)


# ----------------------------------------------------------------------------
# A3  cv2.warpPerspective(src, H, (DW, DH))  (INTER_LINEAR, BORDER_CONSTANT 0)
#     surroundBEV.py:113-114, extrinsicCalib.py:166-169
# ----------------------------------------------------------------------------
def invert3(S) -> np.ndarray:
    """cv::invert of a 3x3 CV_64F matrix (DECOMP_LU takes the closed adjugate form for n <= 3; what
    cv2.warpPerspective applies to H).  Bit-identical to ``cv2.invert(S)[1]``; ``np.linalg.inv`` (LAPACK) differs
    in the last bits, enough to move about one 1/32-px coordinate per million.  Singular -> zeros."""
    S = np.asarray(S, np.float64)
    d = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0])
         + S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    T = np.zeros((3, 3))
    if d == 0.0:
        return T
    d = 1.0 / d
    T[0, 0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    T[0, 1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    T[0, 2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    T[1, 0] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    T[1, 1] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    T[1, 2] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    T[2, 0] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    T[2, 1] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    T[2, 2] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return T


def warp_coords(Hm, DW: int, DH: int, unit: float = TAB):
    """Fixed-point pre-image of every dst pixel (block form, 64-px blocks).  ``unit``
    is 32 for INTER_LINEAR (1/32 px) and 1 for INTER_NEAREST (whole pixels)."""
    M = invert3(Hm).ravel()
    x = np.arange(DW, dtype=np.int64)[None, :]
    y = np.arange(DH, dtype=np.float64)[:, None]
    bx = ((x // 64) * 64).astype(np.float64)
    x1 = (x - (x // 64) * 64).astype(np.float64)
    X0 = M[0] * bx + M[1] * y + M[2]
    Y0 = M[3] * bx + M[4] * y + M[5]
    W0 = M[6] * bx + M[7] * y + M[8]
    W = W0 + M[6] * x1
    with np.errstate(all="ignore"):
        W = np.where(W != 0, unit / W, 0.0)
        fX = np.clip((X0 + M[0] * x1) * W, -2147483648.0, 2147483647.0)
        fY = np.clip((Y0 + M[3] * x1) * W, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    return X, Y


def _sat_i16(a):
    return np.clip(a, -32768, 32767)


def warp_perspective_u8(src: np.ndarray, Hm, DW: int, DH: int, nearest: bool = False) -> np.ndarray:
    X, Y = warp_coords(Hm, DW, DH)
    map1 = np.stack([_sat_i16(X >> INTER_BITS), _sat_i16(Y >> INTER_BITS)], -1).astype(np.int16)
    map2 = ((Y & (TAB - 1)) * TAB + (X & (TAB - 1))).astype(np.uint16)
    if nearest:
        Xn, Yn = warp_coords(Hm, DW, DH, unit=1.0)  # rint of the whole-pixel pre-image
        m = np.stack([_sat_i16(Xn), _sat_i16(Yn)], -1).astype(np.int16)
        return remap_nearest(src, m, None)
    return remap_linear(src, map1, map2)


def warp_perspective_maps(map1: np.ndarray, map2: np.ndarray, Hm, DW: int, DH: int):
    """A4: Camera.get_bev_maps (surroundBEV.py:105-108): warpPerspective applied to
    the 16SC2 and 16UC1 undistort-map planes themselves.  FP32 interpolation with
    the float bilinear table, rounded (half-even) and saturated to the plane type."""
    X, Y = warp_coords(Hm, DW, DH)
    sx = _sat_i16(X >> INTER_BITS)
    sy = _sat_i16(Y >> INTER_BITS)
    fx = (X & (TAB - 1)).astype(np.float32)
    fy = (Y & (TAB - 1)).astype(np.float32)
    s = np.float32(1.0 / TAB)
    ax, ay = fx * s, fy * s
    w00 = (np.float32(1) - ay) * (np.float32(1) - ax)
    w01 = (np.float32(1) - ay) * ax
    w10 = ay * (np.float32(1) - ax)
    w11 = ay * ax

    def interp(plane, lo, hi, dt):
        p = plane if plane.ndim == 3 else plane[..., None]
        t = [(_taps(p, sy + dy, sx + dx)).astype(np.float32) for dy in (0, 1) for dx in (0, 1)]
        acc = (t[0] * w00[..., None] + t[1] * w01[..., None]
               + t[2] * w10[..., None] + t[3] * w11[..., None])  # fp32, left-to-right
        out = np.clip(np.rint(acc), lo, hi).astype(dt)
        return out if plane.ndim == 3 else out[..., 0]

    return (interp(map1, -32768, 32767, np.int16), interp(map2, 0, 65535, np.uint16))


# ----------------------------------------------------------------------------
# A6/A7  masks and blend weights -- surroundBEV.py:119-159, 164-277
# ----------------------------------------------------------------------------
def _i32(pts):
    return np.array(pts).astype(np.int32)  # float expressions truncated toward 0


def plain_polygon(name: str, BW, BH, CW, CH):
    if name == "front":
        return _i32([[0, 0], [BW, 0], [(BW + CW) / 2, (BH - CH) / 2], [(BW - CW) / 2, (BH - CH) / 2]])
    if name == "back":
        return _i32([[0, BH], [BW, BH], [(BW + CW) / 2, (BH + CH) / 2], [(BW - CW) / 2, (BH + CH) / 2]])
    if name == "left":
        return _i32([[0, 0], [0, BH], [(BW - CW) / 2, (BH + CH) / 2], [(BW - CW) / 2, (BH - CH) / 2]])
    if name == "right":
        return _i32([[BW, 0], [BW, BH], [(BW + CW) / 2, (BH + CH) / 2], [(BW + CW) / 2, (BH - CH) / 2]])
    raise Exception("name should be front/back/left/right")


def blend_polygon(name: str, BW, BH, CW, CH):
    if name == "front":
        return _i32([[0, 0], [BW, 0], [BW, BH / 5], [(BW + CW) / 2, (BH - CH) / 2],
                     [(BW - CW) / 2, (BH - CH) / 2], [0, BH / 5]])
    if name == "back":
        return _i32([[0, BH], [BW, BH], [BW, BH - BH / 5], [(BW + CW) / 2, (BH + CH) / 2],
                     [(BW - CW) / 2, (BH + CH) / 2], [0, BH - BH / 5]])
    if name == "left":
        return _i32([[0, 0], [0, BH], [BW / 5, BH], [(BW - CW) / 2, (BH + CH) / 2],
                     [(BW - CW) / 2, (BH - CH) / 2], [BW / 5, 0]])
    if name == "right":
        return _i32([[BW, 0], [BW, BH], [BW - BW / 5, BH], [(BW + CW) / 2, (BH + CH) / 2],
                     [(BW + CW) / 2, (BH - CH) / 2], [BW - BW / 5, 0]])
    raise Exception("name should be front/back/left/right")


def blend_lines(BW, BH, CW, CH):
    """The 8 seam segments of BlendMask.get_lines (surroundBEV.py:236-268)."""
    fl = ((BW - CW) / 2, (BH - CH) / 2)
    fr = ((BW + CW) / 2, (BH - CH) / 2)
    bl = ((BW - CW) / 2, (BH + CH) / 2)
    br = ((BW + CW) / 2, (BH + CH) / 2)
    return {
        "FL": _i32([[0, BH / 5], fl]), "FR": _i32([[BW, BH / 5], fr]),
        "BL": _i32([[0, BH - BH / 5], bl]), "BR": _i32([[BW, BH - BH / 5], br]),
        "LF": _i32([[BW / 5, 0], fl]), "LB": _i32([[BW / 5, BH], bl]),
        "RF": _i32([[BW - BW / 5, 0], fr]), "RB": _i32([[BW - BW / 5, BH], br]),
    }


def fill_poly(BW: int, BH: int, pts: np.ndarray) -> np.ndarray:
    """cv2.fillPoly rasterisation (surroundBEV.py:156-159, 231-234).  Polygon scan
    conversion is init-time host work: the oracle defers to cv2."""
    import cv2
    return cv2.fillPoly(np.zeros((BH, BW), np.uint8), [pts], 255)


def seg_dist(px, py, a, b):
    """|cv2.pointPolygonTest([a,b], p, True)| -- Euclid distance to the closed segment."""
    ax, ay = float(a[0]), float(a[1])
    bx, by = float(b[0]), float(b[1])
    dx, dy = bx - ax, by - ay
    d1x, d1y = px - ax, py - ay
    d2x, d2y = px - bx, py - by
    dot1 = d1x * dx + d1y * dy
    dot2 = d2x * dx + d2y * dy
    cross = d1y * dx - d1x * dy
    with np.errstate(all="ignore"):
        perp2 = cross * cross / (dx * dx + dy * dy)
    sq = np.where(dot1 <= 0, d1x * d1x + d1y * d1y,
                  np.where(dot2 >= 0, d2x * d2x + d2y * d2y, perp2))
    return np.sqrt(sq)


def blend_mask(name: str, BW: int, BH: int, CW, CH) -> np.ndarray:
    """BlendMask.__init__ (surroundBEV.py:165-188) without the Python pixel loop."""
    polys = {n: fill_poly(BW, BH, blend_polygon(n, BW, BH, CW, CH))
             for n in ("front", "back", "left", "right")}
    L = blend_lines(BW, BH, CW, CH)
    order = {"front": [("left", "FL", "LF"), ("right", "FR", "RF")],
             "back": [("left", "BL", "LB"), ("right", "BR", "RB")],
             "left": [("front", "LF", "FL"), ("back", "LB", "BL")],
             "right": [("front", "RF", "FR"), ("back", "RB", "BR")]}[name]
    m = polys[name].copy()
    for other, la, lb in order:
        ys, xs = np.nonzero(m & polys[other])  # overlap uses the running maskA
        x = xs.astype(np.float64)
        y = ys.astype(np.float64)
        dA = seg_dist(x, y, L[la][0], L[la][1])
        dB = seg_dist(x, y, L[lb][0], L[lb][1])
        val = dA ** 2 / (dA ** 2 + dB ** 2 + 1e-6) * 255
        m[ys, xs] = val.astype(np.uint8)  # truncation, as NumPy's item assignment does
    return m


# ----------------------------------------------------------------------------
# A8  per-camera mask / weight and compose -- surroundBEV.py:161-162, 279-280, 316-324
# ----------------------------------------------------------------------------
def apply_plain(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    return np.where(mask[..., None] != 0, img, 0).astype(np.uint8)


def apply_blend(img: np.ndarray, mask: np.ndarray) -> np.ndarray:
    w = (mask.astype(np.float64) / 255.0).astype(np.float32)
    return (img.astype(np.float32) * w[..., None]).astype(np.uint8)


def sat_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.minimum(a.astype(np.int32) + b.astype(np.int32), 255).astype(np.uint8)


# ----------------------------------------------------------------------------
# A9  luminance_balance -- surroundBEV.py:57-79
# ----------------------------------------------------------------------------
def _div_table(scale: int, per: float):
    t = np.zeros(256, np.int64)
    i = np.arange(1, 256, dtype=np.float64)
    t[1:] = np.rint((scale << 12) / (per * i)).astype(np.int64)
    return t


_SDIV = _div_table(255, 1.0)
_HDIV = _div_table(180, 6.0)


def bgr2hsv(img: np.ndarray):
    b, g, r = (img[..., c].astype(np.int64) for c in range(3))
    v = np.maximum(np.maximum(b, g), r)
    mn = np.minimum(np.minimum(b, g), r)
    d = v - mn
    s = (d * _SDIV[v] + 2048) >> 12
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h * _HDIV[d] + 2048) >> 12
    h = np.where(h < 0, h + 180, h)
    return h.astype(np.uint8), s.astype(np.uint8), v.astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def hsv2bgr(h: np.ndarray, s: np.ndarray, v: np.ndarray, simd_body: np.ndarray | bool = True) -> np.ndarray:
    """8-bit HSV2BGR.  OpenCV's vector body truncates x*255 while its scalar row
    tail rounds; ``simd_body`` selects per pixel (see ``hsv_tail_mask``)."""
    f32 = np.float32
    sf = s.astype(f32) * f32(1.0 / 255.0)
    vf = v.astype(f32) * f32(1.0 / 255.0)
    hx = h.astype(f32) * f32(6.0 / 180.0)
    sec = np.trunc(hx)
    f = hx - sec
    sec = sec.astype(np.int64) % 6
    # fmaf(-s, f, 1) and fmaf(-s, 1-f, 1): emulate the single rounding in float64
    # (exact product of two float32 fits in float64, so one final rounding).
    t2m = (f32(1) + (-sf.astype(np.float64) * f.astype(np.float64))).astype(f32)
    omf = f32(1) - f
    t3m = (f32(1) + (-sf.astype(np.float64) * omf.astype(np.float64))).astype(f32)
    tab = np.stack([vf, vf * (f32(1) - sf), vf * t2m, vf * t3m], axis=-1)
    idx = _SECTOR[sec]
    bgr = np.take_along_axis(tab, idx, axis=-1) * f32(255.0)
    trunc = np.clip(np.trunc(bgr), 0, 255)
    rnd = np.clip(np.rint(bgr), 0, 255)
    if isinstance(simd_body, bool):
        out = trunc if simd_body else rnd
    else:
        out = np.where(simd_body[..., None], trunc, rnd)
    return out.astype(np.uint8)


def luminance_offsets(frames):
    """rint(V_mean - V_i) for each frame: the scalar cv2.add rounds half-even."""
    means = [float(np.maximum(np.maximum(f[..., 0], f[..., 1]), f[..., 2]).astype(np.uint64).sum())
             / (f.shape[0] * f.shape[1]) for f in frames]
    vm = (means[0] + means[1] + means[2] + means[3]) / 4 if len(means) == 4 else sum(means) / len(means)
    return [int(np.rint(vm - m)) for m in means], means


def luminance_apply(frame: np.ndarray, delta: int, simd_body=True) -> np.ndarray:
    h, s, v = bgr2hsv(frame)
    v2 = np.clip(v.astype(np.int64) + delta, 0, 255).astype(np.uint8)
    return hsv2bgr(h, s, v2, simd_body)


def luminance_balance(frames, hsv_vec_width: int = 32):
    deltas, _ = luminance_offsets(frames)
    out = []
    for f, d in zip(frames, deltas):
        W = f.shape[1]
        body = np.zeros(f.shape[:2], bool)
        body[:, : W - (W % hsv_vec_width)] = True
        out.append(luminance_apply(f, d, body))
    return out


# ----------------------------------------------------------------------------
# A10  color_balance -- surroundBEV.py:43-55
# ----------------------------------------------------------------------------
def color_gains(canvas: np.ndarray):
    n = canvas.shape[0] * canvas.shape[1]
    B, G, R = (float(canvas[..., c].astype(np.uint64).sum()) / n for c in range(3))
    K = (R + G + B) / 3
    return K / B, K / G, K / R


def color_balance(canvas: np.ndarray) -> np.ndarray:
    gains = color_gains(canvas)
    out = np.empty_like(canvas)
    for c in range(3):
        out[..., c] = np.clip(np.rint(canvas[..., c].astype(np.float64) * gains[c]), 0, 255).astype(np.uint8)
    return out


# ----------------------------------------------------------------------------
# End-to-end restatement of BevGenerator (surroundBEV.py:282-325)
# ----------------------------------------------------------------------------
NAMES = ("front", "back", "left", "right")


def camera_tables(K, D, Hm, FW, FH, BW, BH, FS=1.0, SS=2.0):
    """Camera.__init__ (surroundBEV.py:82-108): P, undistort maps, BEV maps."""
    P = np.array(K, np.float64)
    P[0, 0] *= FS
    P[1, 1] *= FS
    P[0, 2] = FW / 2 * SS
    P[1, 2] = FH / 2 * SS
    um1, um2 = fisheye_map(K, D, P, int(FW * SS), int(FH * SS))
    bm1, bm2 = warp_perspective_maps(um1, um2, Hm, BW, BH)
    return P, (um1, um2), (bm1, bm2)


def bev_generate(frames, bev_maps, masks, blend: bool, balance: bool, car=None):
    """BevGenerator.__call__ (surroundBEV.py:312-325)."""
    if balance:
        frames = luminance_balance(frames)
    tiles = []
    for f, (m1, m2), mk in zip(frames, bev_maps, masks):
        w = remap_linear(f, m1, m2)
        tiles.append(apply_blend(w, mk) if blend else apply_plain(w, mk))
    out = tiles[0]
    for t in tiles[1:]:
        out = sat_add(out, t)
    if balance:
        out = color_balance(out)
    if car is not None:
        out = sat_add(out, car)
    return out
