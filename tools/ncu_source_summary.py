#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv` per SASS opcode / stall reason / code region.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_source_summary.py src.csv [--regions]
"""
import csv
import sys
from collections import Counter, defaultdict


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    ins = rows[2:]
    def f(r, name):
        try:
            return float(r[col[name]] or 0)
        except (KeyError, ValueError):
            return 0.0
    tot_inst = sum(f(r, "Instructions Executed") for r in ins)
    tot_samp = sum(f(r, "# Samples") for r in ins)
    print(f"kernel: {rows[0][1]}")
    print(f"warp instructions executed: {tot_inst:,.0f}   stall samples: {tot_samp:,.0f}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    sc = Counter({s: sum(f(r, s) for r in ins) for s in stalls})
    print("\nstall reason        samples  share")
    for s, v in sc.most_common(10):
        print(f"{s:18s} {v:9.0f}  {100 * v / max(1, tot_samp):5.1f}%")
    ops = defaultdict(lambda: [0.0, 0.0])
    for r in ins:
        op = r[col["Source"]].split()[0] if not r[col["Source"]].startswith("@") else r[col["Source"]].split()[1]
        op = op.split(".")[0]
        ops[op][0] += f(r, "Instructions Executed")
        ops[op][1] += f(r, "# Samples")
    print("\nopcode      warp-instr     share   stall samples")
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"{op:10s} {n:12,.0f}  {100 * n / tot_inst:6.1f}%   {100 * s / max(1, tot_samp):5.1f}%")
    print("\nmemory instructions: warp-instr, L1 tag requests (global) / shared wavefronts (ideal)")
    mem = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for r in ins:
        src = r[col["Source"]]
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        if op.split(".")[0] in ("LDG", "STG", "LDS", "STS", "LD", "ST", "LDL", "STL", "UTMALDG", "SYNCS", "ATOMS", "RED", "ATOMG"):
            m = mem[op]
            m[0] += f(r, "Instructions Executed"); m[1] += f(r, "L1 Tag Requests Global")
            m[2] += f(r, "L1 Wavefronts Shared"); m[3] += f(r, "L1 Wavefronts Shared Ideal")
    for op, m in sorted(mem.items(), key=lambda kv: -kv[1][0]):
        extra = f"tags/instr {m[1] / m[0]:.2f}" if m[1] else (f"wavefronts/instr {m[2] / m[0]:.2f} (ideal {m[3] / m[0]:.2f})" if m[2] else "")
        print(f"{op:28s} {m[0]:12,.0f}  {m[1]:12,.0f} {m[2]:12,.0f}  {extra}")
    if "--regions" in sys.argv:
        # hottest 40 instructions by stall samples
        print("\nhottest instructions by stall samples")
        for r in sorted(ins, key=lambda r: -f(r, "# Samples"))[:40]:
            top = max(stalls, key=lambda s: f(r, s))
            print(f"{r[col['Address']][-5:]} {f(r, '# Samples'):6.0f} {f(r, 'Instructions Executed'):10,.0f}  {top:16s} {r[col['Source']][:70]}")


if __name__ == "__main__":
    main()
