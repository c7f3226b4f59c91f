#!/usr/bin/env python
"""Instruction / stall distribution by source line for one kernel.  usage: tools/ncu_lines.py report.ncu-rep kernel_substr [min_pct]"""
import csv, collections, io, subprocess, sys
rep, kname = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kern = hdr = fpath = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ''])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": kern = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or kname not in (kern or '') or len(r) < len(hdr) - 5 or r[2] != '-': continue
    try:
        inst = float(r[hdr.index("Instructions Executed")]); samp = float(r[hdr.index("Warp Stall Sampling (All Samples)")])
        th = float(r[hdr.index("Thread Instructions Executed")])
    except Exception: continue
    d = agg[(fpath, int(r[0]))]; d[0] += inst; d[1] += samp; d[2] += th; d[3] = r[1].strip()
ti = sum(d[0] for d in agg.values()); ts = sum(d[1] for d in agg.values())
print("total warp-inst %.3e samples %d" % (ti, ts))
for k in sorted(agg):
    d = agg[k]
    if d[0] / ti * 100 > thr or d[1] / ts * 100 > thr:
        print(f"{k[0][:10]:10s} {k[1]:4d} {d[0]/ti*100:5.1f}%i {d[1]/ts*100:5.1f}%s act={d[2]/max(d[0],1):4.1f}  {d[3][:110]}")
