"""GPU tests of the multi-GPU sharding in libbevk.so (bevk_shard_* / bevk_bev_run_sharded) and its thin Python caller
ShardedBev.  Everything goes through the C ABI.

  * one GPU is enough for the decomposition itself: bevk_shard_render renders the slabs of ANY rank, bevk_shard_compose
    composes them; for world = 2, 3, 4, 8 the composed canvases must equal the plain render and the golden of the
    unmodified reference, and the library's partition / slab geometry must equal the Python statement the CPU gloo tests use;
  * with >= 2 GPUs the real thing: two processes, NCCL, ShardedBev with both policies (run under `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

from oracle import cv2_path as C
from oracle import restate as R
from tests.helpers import NAMES, h16

pytestmark = pytest.mark.gpu


def _engine(fx, g, blend, calib=None, device=0):
    from cameracalibration_b200 import _lib as L
    from cameracalibration_b200 import ops
    calib = calib or fx.scaled_calib(g)
    e = ops.BevEngine(4, (g.FW, g.FH), (g.BW, g.BH), ctx=L.Context(device))
    masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) if blend else C.plain_mask(n, g) for n in NAMES]
    for i, n in enumerate(NAMES):
        K, D, H = calib[n]
        e.set_camera(i, K, D, C.dst_camera_matrix(K, g.FW, g.FH, g.FS, g.SS), (int(g.FW * g.SS), int(g.FH * g.SS)), H)
        e.set_mask(i, masks[i])
    e.finalize()
    return e, masks


@pytest.mark.parametrize("blend", [False, True])
def test_camera_sharding_on_one_gpu_every_world_size(fx, blend):
    import torch
    from cameracalibration_b200.sharding import ShardedBev, camera_range, slab_rect
    g = fx.geometry()
    e, masks = _engine(fx, g, blend, calib=fx.calib)
    dev = torch.device("cuda", e.ctx.device)
    F = fx.frames()
    sets = [F, [np.ascontiguousarray(f[::-1]) for f in F], [np.ascontiguousarray(np.roll(f, 31, axis=1)) for f in F]]
    d_all = torch.from_numpy(np.stack([np.stack(s) for s in sets])).to(dev)
    car = torch.from_numpy(fx.car()).to(dev)
    full = torch.empty((3, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
    e.run_stack(d_all.data_ptr(), g.FH * g.FW * 3, 3, full.data_ptr(), car.data_ptr())
    e.ctx.sync()
    full = full.cpu().numpy()
    assert h16(full[0]) == fx.gold["native"][f"blend{int(blend)}_balance0"]["car"]
    for world in (2, 3, 4, 8):
        sh = ShardedBev(e, "cameras", rank=0, world=world, connect=False)
        infos = [sh.info(r) for r in range(world)]
        slab_bytes = infos[0][3]
        for r, (lo, hi, rect, sb) in enumerate(infos):
            assert (lo, hi) == camera_range(4, r, world) and tuple(rect) == slab_rect(masks, lo, hi) and sb == slab_bytes
        assert slab_bytes % 256 == 0 and slab_bytes >= max((x1 - x0) * (y1 - y0) * 3 for _, _, (x0, y0, x1, y1), _ in infos)
        if world == 4:
            assert slab_bytes < 0.4 * g.BW * g.BH * 3          # 1.18 MB slabs, not 3 MB canvases
        slabs = sh.slab_buffer(3)
        for r in range(world):                                 # every rank's render, on this one GPU
            sh.render_slabs(d_all, r, slabs)
        out = torch.empty((3, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
        sh.compose(slabs, out, car)
        torch.cuda.synchronize()
        assert (out.cpu().numpy() == full).all(), world
    # policy 'frames' and a world of one go straight to the plain render
    out = torch.empty((3, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
    ShardedBev(e, "frames", rank=1, world=4).render(d_all, out, car)
    torch.cuda.synchronize()
    assert (out.cpu().numpy() == full).all()
    with pytest.raises(Exception, match="bevk_shard_connect"):
        ShardedBev(e, "cameras", rank=0, world=2, connect=False).render(d_all, out, car)


def test_peer_store_kernel_variant_on_one_gpu(fx):
    """bevk_bev_run_scattered with a world of one: the render kernel's peer-store instantiation (k_bev_tma<.., SCATTER>)
    writes the slabs through the peer table -- here its own receive buffer -- and the rank composes them; no second GPU
    and no NCCL are needed for that.  Batch 6 (ragged tail of the 4-frame-set units), with and without the car."""
    import ctypes as C
    import torch
    from cameracalibration_b200 import _lib as L
    g = fx.geometry()
    e, _ = _engine(fx, g, True, calib=fx.calib)
    dev = torch.device("cuda", e.ctx.device)
    F = fx.frames()
    sets = [[np.ascontiguousarray(np.roll(f, 17 * i + 3 * c, axis=1)) for c, f in enumerate(F)] for i in range(6)]
    d_all = torch.from_numpy(np.stack([np.stack(s) for s in sets])).to(dev)
    car = torch.from_numpy(fx.car()).to(dev)
    lib, h = e.ctx.lib, e.ctx.h
    L.check(lib.bevk_shard_configure(h, L.SHARD_CAMERAS, 0, 1))
    handle = (C.c_uint8 * 64)()
    L.check(lib.bevk_shard_prepare(h, 6, handle))
    L.check(lib.bevk_shard_attach(h, bytes(handle)))
    for c in (None, car):
        full = torch.empty((6, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
        e.run_stack(d_all.data_ptr(), g.FH * g.FW * 3, 6, full.data_ptr(), 0 if c is None else c.data_ptr())
        own = torch.zeros((6, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
        n_own = C.c_int()
        L.check(lib.bevk_bev_run_scattered(h, C.c_void_p(d_all.data_ptr()), g.FH * g.FW * 3, 6, C.c_void_p(0 if c is None else c.data_ptr()), 0,
                                           C.c_void_p(own.data_ptr()), C.byref(n_own)))
        e.ctx.sync()
        assert n_own.value == 6 and e.last_path() == "tma"
        assert (own.cpu().numpy() == full.cpu().numpy()).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cameracalibration_b200.sharding import ShardedBev
        from tests.helpers import Fixtures
        fx = Fixtures()
        g = fx.geometry()
        res = {}
        for blend in (False, True):
            e, masks = _engine(fx, g, blend, calib=fx.calib, device=rank)
            dev = torch.device("cuda", rank)
            F = fx.frames()
            sets = [F, [np.ascontiguousarray(f[::-1]) for f in F], F, F, [np.ascontiguousarray(np.roll(f, 31, axis=1)) for f in F]]
            host = np.stack([np.stack(s) for s in sets])
            lo, hi = ShardedBev(e, "cameras", connect=False).my_cameras()
            mine = host.copy()
            mine[:, :lo] = 0xAB                      # a rank only holds its own cameras' frames: poison the others
            mine[:, hi:] = 0xAB
            d_mine = torch.from_numpy(mine).to(dev)
            car = torch.from_numpy(fx.car()).to(dev)
            sh = ShardedBev(e, "cameras")            # NCCL id over torch.distributed, bevk_shard_connect
            out = torch.empty((5, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):
                d_in = d_mine.clone()                # produced on `stream`: render() must order itself after it
                sh.render(d_in, out, car)
                got = out.clone()
            stream.synchronize()
            got = got.cpu().numpy()
            ref = C.RefBev(fx.calib, g, blend, False, masks=masks)
            ok = all((got[i] == ref(*sets[i], fx.car())).all() for i in (0, 1, 4))
            res[f"cameras_blend{int(blend)}"] = (ok, h16(got[0]) == fx.gold["native"][f"blend{int(blend)}_balance0"]["car"], sh.link_bytes())
            # the fused form: slabs stored straight into the owning rank over NVLink, canvases stay sharded; three steps so
            # that both halves of the double-buffered receive area are used and reused
            own = sh.own_frame_sets(5)
            out_own = torch.zeros((3, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
            okp = True
            for step in range(3):
                d_in = d_mine.roll(step, 0).contiguous()
                torch.cuda.synchronize()
                n_own = sh.render_scattered(d_in, out_own, car)
                torch.cuda.synchronize()
                want = [ref(*sets[(b - step) % 5], fx.car()) for b in own]
                okp &= n_own == len(own) and all((out_own[i].cpu().numpy() == want[i]).all() for i in range(n_own))
            res[f"p2p_blend{int(blend)}"] = (bool(okp), sh.link_bytes() > 0, sh.link_bytes())
            # frames policy: each rank its own block of the batch, no collective
            shf = ShardedBev(e, "frames")
            a, b = shf.my_frame_sets(5)
            d_own = torch.from_numpy(host[a:b]).to(dev)
            out_f = torch.empty((b - a, g.BH, g.BW, 3), dtype=torch.uint8, device=dev)
            shf.render(d_own, out_f, car, balance=True)
            torch.cuda.synchronize()
            refb = C.RefBev(fx.calib, g, blend, True, masks=masks)
            res[f"frames_blend{int(blend)}"] = (bool((out_f.cpu().numpy()[0] == refb(*sets[a], fx.car())).all()), shf.link_bytes() == 0, (a, b))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_sharded_bev_world2_nccl(fx):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(60)
    for r in (0, 1):
        for key, val in res[r].items():
            assert val[0] and val[1], (r, key, val)
        assert res[r]["cameras_blend1"][2] > 0          # bytes did cross NVLink
    assert res[0]["frames_blend0"][2] == (0, 3) and res[1]["frames_blend0"][2] == (3, 5)
