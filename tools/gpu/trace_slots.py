#!/usr/bin/env python
"""Slot timeline of k_bev_tma from a -DBEVK_TRACE build (bevk_api.cu: launch_bev_tma, bevk_bev_tma.cuh).

    BEVK_NVCC_FLAGS=-DBEVK_TRACE python -m cameracalibration_b200.build --force     (measurement build only)
    BEVK_TRACE_FILE=trace.bin python bench.py --no-cpu-baseline --e2e-steps 1
    python tools/gpu/trace_slots.py trace.bin

The 12th launch records, for the first 8 CTAs and their first 512 ring slots, clock64 stamps of the producer
(0: starts waiting for the slot, 1: has it, 2: descriptor / entries / boxes posted; 6: bytes expected, 7: descriptor flags)
and of consumer thread 0 (3: starts waiting for the slot, 4: slot complete, 5: done, before the release, 8: released;
slots that end a unit: 9: write-out starts (after the barrier, if any), 10: written).  Printed per
slot class: share of the time, mean period (done -> done), consumer compute, consumer wait, gap between the previous
slot's done and this wait (release, unit end: barrier + write-out), bytes per slot, cycles per 1000 (sample, frame-set)."""
import sys
from collections import defaultdict

import numpy as np


def main():
    t = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(8, 512, 16).astype(np.int64)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    tot = 0
    spans = []
    for cta in range(8):
        T = t[cta]
        n = int((T[:, 4] > 0).sum())
        if n < 3:
            continue
        T = T[:n]
        spans.append((T[n - 2, 5] - T[0, 3]) / 1965.0)
        for i in range(1, n):
            fl = int(T[i, 7])
            if fl & 1:      # D_END
                continue
            kind, nk, gather = (fl >> 20) & 3, (fl >> 16) & 15, bool(fl & 2)
            key = ("gather" if gather else ("4 boxes", "2 boxes", "1 box")[min(kind, 2)], nk)
            a = agg[key]
            per = T[i, 5] - T[i - 1, 5]
            a[0] += 1; a[1] += per; a[2] += T[i, 5] - T[i, 4]; a[3] += T[i, 4] - T[i, 3]; a[4] += T[i, 6]; a[5] += T[i, 3] - T[i - 1, 5]
            tot += per
    print("CTA spans (us at 1965 MHz):", np.round(spans, 1))
    print(f"{'slot class':22s} {'n':>5s} {'share':>6s} {'period':>7s} {'compute':>8s} {'wait':>6s} {'gap':>6s} {'bytes':>7s} {'cyc/1k sample-fs':>17s}")
    S = 0
    for k, a in sorted(agg.items()):
        fs = {"4 boxes": 4, "2 boxes": 2, "1 box": 1, "gather": 4}[k[0]]
        sfs = k[1] * 256 * fs
        S += sfs * a[0]
        print(f"{k[0] + ', %d groups' % k[1]:22s} {a[0]:5d} {100 * a[1] / tot:5.1f}% {a[1] // a[0]:7d} {a[2] // a[0]:8d} {a[3] // a[0]:6d} {a[5] // a[0]:6d} "
              f"{a[4] // a[0]:7d} {int(a[1] / a[0] / max(1, sfs) * 1000):17d}")
    # unit ends: release, barrier, write-out
    rel, bar, wo, n_last, n_rows = [], [], [], 0, 0
    for cta in range(8):
        T = t[cta]
        n = int((T[:, 4] > 0).sum())
        for i in range(n):
            if T[i, 8] > 0:
                rel.append(T[i, 8] - T[i, 5])
            if T[i, 10] > 0 and T[i, 9] > 0:
                n_last += 1
                n_rows += bool(int(T[i, 7]) & 512)
                bar.append(T[i, 9] - T[i, 8]); wo.append(T[i, 10] - T[i, 9])
    if rel:
        print(f"release (done -> released): mean {np.mean(rel):.0f} cycles")
    if wo:
        print(f"unit ends traced: {n_last} ({n_rows} without barrier): released -> write-out starts mean {np.mean(bar):.0f} (median {np.median(bar):.0f}), "
              f"write-out mean {np.mean(wo):.0f} (median {np.median(wo):.0f}) cycles")
    tc, tw, tg = (sum(a[i] for a in agg.values()) for i in (2, 3, 5))
    print(f"all: {tot / max(1, S) * 1000:.0f} cycles per 1000 (sample, frame-set); compute {100 * tc / tot:.1f}%  wait {100 * tw / tot:.1f}%  gap {100 * tg / tot:.1f}%")


if __name__ == "__main__":
    main()
