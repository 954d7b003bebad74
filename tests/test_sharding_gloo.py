"""World-size-2 CPU (gloo) tests of the multi-GPU host logic: the partition functions and the
camera-sharded decomposition (per-rank partial canvases -> ONE all-gather -> saturating sum),
checked against the oracle's full render.  The GPU kernels are not involved here; on the GPU box
tests/test_gpu_parity.py::test_device_resident_and_camera_sharded_compose covers the same
decomposition through libbevk.so."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cameracalibration_b200.sharding import block_range, camera_range
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
        lo, hi = camera_range(4, rank, world)
        part = np.zeros((g.BH, g.BW, 3), np.uint8)
        for c in range(lo, hi):   # this rank's cameras only (cv2.add order inside the rank)
            cam = C.RefCamera(*calib[NAMES[c]], g)
            w = cam.raw2bev(frames[c])
            t = R.apply_blend(w, masks[c]) if blend else R.apply_plain(w, masks[c])
            part = R.sat_add(part, t)
        mine = torch.from_numpy(part.reshape(-1))
        gathered = torch.empty(world * mine.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, mine)          # the single collective of this policy
        parts = gathered.view(world, -1).numpy()
        out = parts[0]
        for r in range(1, world):
            out = R.sat_add(out, parts[r])
        full = C.RefBev(calib, g, blend, False, masks=masks)(*frames)
        ok = bool((out.reshape(full.shape) == full).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("blend", [False, True])
def test_camera_sharded_compose_world2(blend):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, blend, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
