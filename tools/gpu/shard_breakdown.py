"""One-GPU timing of the two halves of the camera-sharded step (world = 4, bench workload): the slab render of every
rank (k_bev_tma restricted to the rank's cameras, writing its slab) and the compose, CUDA events on the ctx stream.
What remains of a multi-GPU step beyond these is the exchange (all-gather, or peer stores + barrier) and launch gaps."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cameracalibration_b200.sharding import ShardedBev  # noqa: E402

w = dict(bench.WORKLOAD)
eng, calib, masks, g = bench.build_engine(w, 0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
eng.ctx.set_stream(stream.cuda_stream)
nb, nc = w["batch"], w["n_cam"]
host = bench.synthetic_frames(w["FW"], w["FH"], nc, nb, seed=1000)
d_frames = torch.from_numpy(host).to(dev)
out = {"workload": "32 frame-sets x 4 x 1920x1080 -> 1000x1000 blend, world 4 emulated on one GPU"}
for world in (4,):
    sh = ShardedBev(eng, "cameras", rank=0, world=world, connect=False)
    slabs = sh.slab_buffer(nb)
    d_out = torch.empty((nb, w["BH"], w["BW"], 3), dtype=torch.uint8, device=dev)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    with torch.cuda.stream(stream):
        for r in range(world):
            out[f"render_rank{r}_ms"] = timed(lambda r=r: sh.render_slabs(d_frames, r, slabs, stream=stream.cuda_stream))
        out["compose_all_32_ms"] = timed(lambda: sh.compose(slabs, d_out, None, stream=stream.cuda_stream))
        out["full_render_one_gpu_ms"] = timed(lambda: eng.run_stack(d_frames.data_ptr(), w["FW"] * w["FH"] * 3, nb, d_out.data_ptr()))
    out["slab_bytes"] = sh.info(0)[3]
print(json.dumps(out))
