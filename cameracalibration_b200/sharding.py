"""Multi-GPU sharding of the BEV path: one process per GPU, torch.distributed for the plumbing.

The reference is single-process (no collective anywhere); the path shards two ways:

  * frame-sets per GPU ("frames"): every rank renders its own block of frame-sets with a full
    replica of the (8 MB) LUT.  No data-path collective; an optional all-gather only if every
    rank must end up with every canvas.
  * cameras per GPU ("cameras"): rank r renders the masked, weighted partial canvas of its
    cameras only (bevk_bev_run_device_cams), ONE all-gather moves the partial canvases over
    NVLink, and each rank composes them with the saturating sum (bevk_sat_sum_device).  The
    compose is exact because cv2.add's saturation is order-independent on this path (blend
    weights sum to <= 255; plain seams overlap at most pairwise -- SURVEY 8a row a10), which
    tests/test_sharding_gloo.py checks against the reference's result.  balance=True is not available in
    this mode (it needs the per-camera V sums before the warp).

The pure partition functions below are what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations


def block_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of n_items over world ranks; the first n_items % world ranks
    get one extra item.  Returns [lo, hi)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def camera_range(n_cam: int, rank: int, world: int) -> tuple[int, int]:
    """Cameras of one rank.  With more ranks than cameras the trailing ranks get an empty range
    (they still join the all-gather with an all-zero partial canvas)."""
    return block_range(n_cam, rank, world)


class ShardedBev:
    """Drives a BevEngine under torch.distributed.  Tensors are torch CUDA tensors; the engine
    only sees their device pointers."""

    def __init__(self, engine, policy: str = "frames", group=None):
        import torch.distributed as dist
        if policy not in ("frames", "cameras"):
            raise ValueError("policy must be 'frames' or 'cameras'")
        self.e, self.policy, self.group = engine, policy, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._parts = None
        self._mine = None

    def my_frame_sets(self, n_sets: int):
        return block_range(n_sets, self.rank, self.world)

    def my_cameras(self):
        return camera_range(self.e.n_cam, self.rank, self.world)

    def render(self, frame_ptrs, batch: int, out, car=None, balance: bool = False):
        """frame_ptrs: int64 CUDA tensor [batch * n_cam] of device pointers (for policy
        'cameras' every rank passes the full table but only its cameras' entries are read).
        out: uint8 CUDA tensor [batch, BH, BW, 3].  Only enqueues work on the engine's stream
        (plus the collective on torch's current stream for 'cameras')."""
        import torch
        import torch.distributed as dist
        e = self.e
        carp = 0 if car is None else car.data_ptr()
        if self.policy == "frames" or self.world == 1:
            e.run_device(frame_ptrs.data_ptr(), batch, out.data_ptr(), carp, balance)
            return out
        if balance:
            raise ValueError("balance=True is not supported with camera sharding")
        lo, hi = self.my_cameras()
        n = out.numel()
        if self._parts is None or self._parts.numel() != n * self.world:
            self._parts = torch.empty((self.world, n), dtype=torch.uint8, device=out.device)
            self._mine = torch.empty((n,), dtype=torch.uint8, device=out.device)
        e.run_device_cams(frame_ptrs.data_ptr(), batch, lo, hi, self._mine.data_ptr())
        dist.all_gather_into_tensor(self._parts.view(-1), self._mine, group=self.group)
        # the all-gather carries every rank's partial canvas; compose locally (8 partials per call)
        ptrs = [self._parts[r].data_ptr() for r in range(self.world)]
        while len(ptrs) > 8:   # bevk_sat_sum_device takes up to 8 inputs: fold the rest pairwise
            head, ptrs = ptrs[:8], ptrs[8:]
            e.sat_sum_device(head, n, self._parts[0].data_ptr(), 0)
            ptrs = [self._parts[0].data_ptr()] + ptrs
        e.sat_sum_device(ptrs, n, out.data_ptr(), 0)
        if car is not None:   # car overlay, tiled over the batch
            per = n // batch
            for b in range(batch):
                e.sat_sum_device([out.data_ptr() + b * per], per, out.data_ptr() + b * per, carp)
        return out
