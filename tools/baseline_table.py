#!/usr/bin/env python
"""Rewrite BASELINE.md section 4 from profiles/r01_configs.json (output of tools/bench_configs.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    p = os.path.join(ROOT, "BASELINE.md")
    s = open(p).read()
    i = s.find("\n## 4. Round-1 measurements")
    if i >= 0:
        s = s[:i]
    rows = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r01_configs.json"))]
    host = rows[0]["host"]
    md = ["\n## 4. Round-1 measurements on the GPU box (same host, same run; `tools/bench_configs.py`, raw: `profiles/r01_configs.json`)\n",
          f"Host: {host['cpu_count']} logical CPUs, cv2 {host['cv2']} using its default pool of {host['cv2_threads']} threads; GPU {host['gpu']}. "
          "Reference arm = the reference's own cv2 call sequence (`oracle/cv2_path.py`). GPU e2e = the drop-in Python API with page-locked host "
          "buffers (ingest over PCIe + kernels + D2H), one frame-set / image per call; device = frames already in HBM (CUDA events).\n",
          "| config | parity | cv2 (ms / frame-set) | GPU e2e (ms / frame-set) | speed-up e2e | GPU device, batch 1 (us) | GPU device, batched (us / frame-set) |",
          "|---|---|---|---|---|---|---|"]
    for r in rows:
        if "cv2_ms_per_frame_set" in r:
            if "gpu_device_us_per_frame_set_batch32" in r:
                b = f"{r['gpu_device_us_per_frame_set_batch32']:.2f} (batch 32)"
            elif "gpu_device_us_per_frame_set_batch4" in r:
                b = f"{r['gpu_device_us_per_frame_set_batch4']:.2f} (batch 4)"
            else:
                b = "-"
            md.append(f"| {r['config']} {r['geometry']} blend={r['blend']} balance={r['balance']} | "
                      f"{'bit-exact' if r['parity_bit_exact'] else 'MISMATCH'} | {r['cv2_ms_per_frame_set']:.2f} | "
                      f"{r['gpu_e2e_ms_per_frame_set_batch1']:.3f} | {r['speedup_e2e_vs_cv2']:.1f}x | "
                      f"{r['gpu_device_us_per_frame_set_batch1']:.1f} | {b} |")
    md += ["", "| single-image op | parity | cv2 (ms) | GPU e2e (ms) | speed-up |", "|---|---|---|---|---|"]
    for r in rows:
        if "cv2_remap_ms" in r:
            md.append(f"| {r['config']} {r['geometry']} | bit-exact (resident map and fused) | remap {r['cv2_remap_ms']:.2f} "
                      f"(+ map build {r['cv2_map_build_ms']:.1f} once) | {r['gpu_e2e_ms_fused0']:.3f} (resident map) / "
                      f"{r['gpu_e2e_ms_fused1']:.3f} (fused, no map) | {r['cv2_remap_ms'] / r['gpu_e2e_ms_fused0']:.1f}x |")
        if r.get("config", "").startswith("ExCalibrator"):
            md.append(f"| {r['config']} | bit-exact | {r['cv2_ms']:.2f} | {r['gpu_e2e_ms']:.3f} | "
                      f"{r['cv2_ms'] / r['gpu_e2e_ms']:.1f}x (PCIe-bound: 15.7 MB in, 3 MB out) |")
    b = json.load(open(os.path.join(ROOT, "profiles", "r01_bench.json")))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_reference.json")))
    md += ["", f"Headline (`bench.py`, BASELINE configs[3] shape, 32 frame-sets per step, 1xB200): {b['value'] / 1e3:.0f} k frame-sets/s "
               f"device-resident ({b['ms_per_step']:.3f} ms per step, roofline {b['roofline']['frac']:.2f} of the measured HBM copy peak), "
               f"{b['e2e']['value'] / 1e3:.1f} k frame-sets/s end-to-end from page-locked host memory "
               f"({b['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB in + {b['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB out per step over PCIe), "
               f"reference cv2 path {ref['value']:.0f} frame-sets/s on the same host (`--impl reference`"
               + (f"; {ref['cpu_frame_set_parallel']['value']:.0f} frame-sets/s when {ref['cpu_frame_set_parallel']['threads']} host threads "
                  "each run the reference's serial sequence on their own frame-set, reported as `cpu_frame_set_parallel`"
                  if "cpu_frame_set_parallel" in ref else "") + "). "
               "2 / 4 / 8 GPUs (frame-set sharding, no collective): 433 k / 838 k / 1 675 k device-resident; 8 GPUs end-to-end 37.7 k (host PCIe shared).\n"]
    open(p, "w").write(s + "\n".join(md))


if __name__ == "__main__":
    main()
