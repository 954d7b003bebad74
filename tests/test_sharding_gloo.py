"""World-size-2/3 CPU (gloo) tests of the multi-GPU host logic: the partition functions, the slab geometry and the
camera-sharded decomposition (per-rank slabs -> ONE all-gather -> saturating compose), checked against the oracle's full
render.  The GPU kernels are not involved here; on the GPU box tests/test_gpu_shard.py drives the same decomposition
through libbevk.so (bevk_shard_* / bevk_bev_run_sharded) and compares its partition with the functions tested here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cameracalibration_b200.sharding import block_range, camera_range, slab_rect
from oracle import cv2_path as C
from oracle import restate as R
from tests.helpers import NAMES, Fixtures


def test_block_range_partitions_exactly():
    for n in (0, 1, 4, 7, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [block_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert camera_range(4, 5, 8) == (4, 4)          # more ranks than cameras: empty range
    with pytest.raises(ValueError):
        block_range(4, 2, 2)


def test_slab_rects_cover_their_masks_and_are_tile_aligned():
    """at 1000x1000 the per-camera boxes are front 1000x301, back 1000x300, left 376x1000, right 375x1000
    for plain and blend masks; the slab is that box grown to 32-px tile boundaries and clipped to the canvas."""
    fx = Fixtures()
    g = fx.geometry()
    for blend in (False, True):
        masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) if blend else C.plain_mask(n, g) for n in NAMES]
        rects = [slab_rect(masks, k, k + 1) for k in range(4)]
        assert rects == [(0, 0, 1000, 320), (0, 672, 1000, 1000), (0, 0, 384, 1000), (608, 0, 1000, 1000)], rects
        for k, (x0, y0, x1, y1) in enumerate(rects):
            outside = masks[k].copy()
            outside[y0:y1, x0:x1] = 0
            assert not outside.any() and x0 % 32 == 0 and y0 % 32 == 0
        assert slab_rect(masks, 0, 2) == (0, 0, 1000, 1000) and slab_rect(masks, 2, 2) == (0, 0, 0, 0)
    assert max((x1 - x0) * (y1 - y0) * 3 for x0, y0, x1, y1 in rects) == 392 * 1000 * 3       # 1.18 MB vs the 3 MB canvas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, blend, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = Fixtures()
        g = fx.geometry(320, 256, 250, 250)
        calib = fx.scaled_calib(g)
        masks = [R.blend_mask(n, g.BW, g.BH, g.CW, g.CH) if blend else C.plain_mask(n, g) for n in NAMES]
        frames = fx.frames(g.FW, g.FH)
        rects = [slab_rect(masks, *camera_range(4, r, world)) for r in range(world)]
        slab_bytes = max((x1 - x0) * (y1 - y0) * 3 for x0, y0, x1, y1 in rects)
        lo, hi = camera_range(4, rank, world)
        part = np.zeros((g.BH, g.BW, 3), np.uint8)
        for c in range(lo, hi):   # this rank's cameras only (cv2.add order inside the rank)
            cam = C.RefCamera(*calib[NAMES[c]], g)
            w = cam.raw2bev(frames[c])
            t = R.apply_blend(w, masks[c]) if blend else R.apply_plain(w, masks[c])
            part = R.sat_add(part, t)
        x0, y0, x1, y1 = rects[rank]
        mine = torch.zeros(slab_bytes, dtype=torch.uint8)
        mine[:(x1 - x0) * (y1 - y0) * 3] = torch.from_numpy(np.ascontiguousarray(part[y0:y1, x0:x1]).reshape(-1))
        assert not np.delete(part.reshape(-1, 3), np.ravel_multi_index(np.mgrid[y0:y1, x0:x1].reshape(2, -1), (g.BH, g.BW)),
                             axis=0).any() if x1 > x0 else not part.any()          # nothing of this rank lies outside its slab
        gathered = torch.empty(world * slab_bytes, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, mine)          # the single collective of this policy
        out = np.zeros((g.BH, g.BW, 3), np.uint8)
        for r, (a0, b0, a1, b1) in enumerate(rects):
            if a1 > a0:
                slab = gathered[r * slab_bytes:r * slab_bytes + (a1 - a0) * (b1 - b0) * 3].numpy().reshape(b1 - b0, a1 - a0, 3)
                out[b0:b1, a0:a1] = R.sat_add(out[b0:b1, a0:a1], slab)
        full = C.RefBev(calib, g, blend, False, masks=masks)(*frames)
        q.put((rank, bool((out == full).all()), slab_bytes <= g.BW * g.BH * 3))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,blend", [(2, False), (2, True), (3, True)])
def test_camera_sharded_slab_compose(world, blend):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, blend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(r, True, True) for r in range(world)]
