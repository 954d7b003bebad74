#!/usr/bin/env python
"""Secondary measurements on the GPU box: every BASELINE.json config next to the reference's
cv2 path on the same host (oracle/cv2_path.py; the reference is Python over OpenCV).

    python tools/bench_configs.py > gpurun_out/configs.json

For each config: parity (GPU == cv2 on the timed inputs), device-resident time per frame-set
(batch 1 and batch 32, CUDA events), end-to-end time through the public API from pinned host
frames, and the cv2 time (default threads).  One JSON object per line.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import torch  # noqa: E402

from cameracalibration_b200 import ops, pinned_empty  # noqa: E402
from oracle import cv2_path as C  # noqa: E402
from oracle import restate as R  # noqa: E402
from tests.helpers import Fixtures, NAMES  # noqa: E402


def med_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def bev_config(fx, name, FW, FH, BW, BH, blend, balance, batches=(1, 32)):
    g = fx.geometry(FW, FH, BW, BH)
    calib = fx.scaled_calib(g)
    masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) if blend else C.plain_mask(n, g) for n in NAMES]
    eng = ops.BevEngine(4, (FW, FH), (BW, BH))
    for i, n in enumerate(NAMES):
        K, D, H = calib[n]
        eng.set_camera(i, K, D, C.dst_camera_matrix(K, FW, FH, 1, 2), (FW * 2, FH * 2), H)
        eng.set_mask(i, masks[i])
    eng.finalize()
    frames = fx.frames(FW, FH)
    car = fx.car(BW, BH)
    ref = C.RefBev(calib, g, blend, balance, masks=masks)
    want = ref(*frames, car)
    got = eng.run([frames], car, balance)[0]
    out = {"config": name, "geometry": f"4x{FW}x{FH}->{BW}x{BH}", "blend": blend, "balance": balance,
           "parity_bit_exact": bool((got == want).all())}
    out["cv2_ms_per_frame_set"] = med_ms(lambda: ref(*frames, car), n=15)
    out["cv2_threads"] = cv2.getNumThreads()
    dev = torch.device("cuda", eng.ctx.device)
    stream = torch.cuda.Stream(device=dev)
    eng.ctx.set_stream(stream.cuda_stream)
    for nb in batches:
        d = torch.from_numpy(np.stack([np.stack(frames)] * nb)).to(dev)
        fb = FW * FH * 3
        dcar = torch.from_numpy(car).to(dev)
        dout = torch.empty((nb, BH, BW, 3), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        for _ in range(5):
            eng.run_stack(d.data_ptr(), fb, nb, dout.data_ptr(), dcar.data_ptr(), balance)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record(stream)
        for _ in range(reps):
            eng.run_stack(d.data_ptr(), fb, nb, dout.data_ptr(), dcar.data_ptr(), balance)
        e1.record(stream)
        torch.cuda.synchronize()
        out[f"gpu_device_us_per_frame_set_batch{nb}"] = e0.elapsed_time(e1) / reps / nb * 1e3
        assert bool((dout[0].cpu().numpy() == want).all())
        out["fused_kernel"] = eng.last_path()
    eng.ctx.set_stream(None)
    pin = [pinned_empty(f.shape) for f in frames]
    for p, f in zip(pin, frames):
        p[...] = f
    pout = pinned_empty((1, BH, BW, 3))
    out["gpu_e2e_ms_per_frame_set_batch1"] = med_ms(lambda: eng.run([pin], car, balance, out=pout), n=30)
    out["speedup_e2e_vs_cv2"] = out["cv2_ms_per_frame_set"] / out["gpu_e2e_ms_per_frame_set_batch1"]
    return out


def undistort_config(fx, name, W, H, FS, SS):
    K, D, _ = fx.calib["front"]
    img = fx.img("raw0")
    if (W, H) != (1280, 1024):
        K = np.diag([W / 1280, H / 1024, 1.0]) @ K
        img = cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)
    P = C.dst_camera_matrix(K, W, H, FS, SS)
    size = (int(W * SS), int(H * SS))
    m1, m2 = C.undistort_maps(K, D, P, *size)
    want = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
    out = {"config": name, "geometry": f"{W}x{H}->{size[0]}x{size[1]}"}
    out["cv2_remap_ms"] = med_ms(lambda: cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
    out["cv2_map_build_ms"] = med_ms(lambda: C.undistort_maps(K, D, P, *size), n=5, warm=1)
    pin = pinned_empty(img.shape)
    pin[...] = img
    for fused in (False, True):
        t0 = time.perf_counter()
        u = ops.Undistorter(K, D, P, size, fused=fused)
        u.ctx.sync()
        out[f"gpu_setup_ms_fused{int(fused)}"] = (time.perf_counter() - t0) * 1e3
        out[f"parity_bit_exact_fused{int(fused)}"] = bool((u(img) == want).all())
        pout = pinned_empty(want.shape)
        out[f"gpu_e2e_ms_fused{int(fused)}"] = med_ms(lambda: u(pin, out=pout), n=30)
        assert bool((pout == want).all())
    return out


def warp_config(fx):
    src = fx.img("src_back")
    H = fx.calib["back"][2]
    want = cv2.warpPerspective(src, H, (1000, 1000))
    pin = pinned_empty(src.shape)
    pin[...] = src
    pout = pinned_empty(want.shape)
    return {"config": "ExCalibrator.warp 2560x2048->1000x1000",
            "parity_bit_exact": bool((ops.warp_perspective(src, H, (1000, 1000)) == want).all()),
            "cv2_ms": med_ms(lambda: cv2.warpPerspective(src, H, (1000, 1000))),
            "gpu_e2e_ms": med_ms(lambda: ops.warp_perspective(pin, H, (1000, 1000), out=pout), n=30)}


def main():
    fx = Fixtures()
    print(json.dumps({"host": {"cpu_count": os.cpu_count(), "cv2": cv2.__version__, "cv2_threads": cv2.getNumThreads(),
                               "gpu": torch.cuda.get_device_name(0)}}), flush=True)
    print(json.dumps(undistort_config(fx, "cfg1a InCalibrator.undistort 1280x1024 (FS=0.5)", 1280, 1024, 0.5, 1)), flush=True)
    print(json.dumps(undistort_config(fx, "cfg1b InCalibrator.undistort 640x480 (FS=0.5)", 640, 480, 0.5, 1)), flush=True)
    print(json.dumps(undistort_config(fx, "Camera.undistort 1280x1024->2560x2048 (SS=2)", 1280, 1024, 1, 2)), flush=True)
    print(json.dumps(warp_config(fx)), flush=True)
    print(json.dumps(bev_config(fx, "native", 1280, 1024, 1000, 1000, False, False)), flush=True)
    print(json.dumps(bev_config(fx, "cfg2", 1280, 960, 1000, 1000, False, False)), flush=True)
    print(json.dumps(bev_config(fx, "cfg3", 1920, 1080, 1200, 1200, True, True)), flush=True)
    print(json.dumps(bev_config(fx, "cfg4-shape", 1920, 1080, 1000, 1000, True, False)), flush=True)
    print(json.dumps(bev_config(fx, "cfg5-size (4 of 8 cams)", 3840, 2160, 2000, 2000, True, False, batches=(1, 4))), flush=True)
    print(json.dumps(bev_config(fx, "cfg5-size blend+balance", 3840, 2160, 2000, 2000, True, True, batches=(1,))), flush=True)


if __name__ == "__main__":
    main()
