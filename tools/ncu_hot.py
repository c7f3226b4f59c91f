#!/usr/bin/env python
"""Top source lines of a kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_hot.py report.ncu-rep [file-substring] [top-n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; agg = collections.OrderedDict(); seen_kernel = 0; kern = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        if kern is None: kern = r[1]
        elif r[1] != kern or cur_file is None: pass
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[2] != "-": continue            # keep only the per-source-line aggregate rows (Address == '-')
    try:
        inst = float(r[hdr.index("Instructions Executed")]); samp = float(r[hdr.index("Warp Stall Sampling (All Samples)")])
        thr = float(r[hdr.index("Avg. Threads Executed")])
    except ValueError: continue
    key = (cur_file.split("/")[-1], r[0])
    d = agg.setdefault(key, [0.0, 0.0, 0.0, r[1].strip(), {}])
    d[0] += inst; d[1] += samp; d[2] = max(d[2], thr)
    for name in ("stall_long_sb", "stall_lg", "stall_mio", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math", "stall_not_selected", "stall_branch_resolving", "stall_membar"):
        try: d[4][name] = d[4].get(name, 0) + float(r[hdr.index(name)])
        except (ValueError, IndexError): pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"kernel: {kern[:100] if kern else None}\ntotal warp-inst {ti:.3e}, samples {ts:.0f} (multiple kernel instances in the report are summed)")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    if want and want not in f: continue
    st = sorted(v[4].items(), key=lambda kv: -kv[1])[:2]
    print(f"{v[0]/ti*100:5.1f}%i {v[1]/ts*100:5.1f}%s thr{v[2]:4.0f} {f}:{ln:>4} {v[3][:95]}  [{', '.join(f'{k[6:]}={int(x)}' for k, x in st)}]")
if len(sys.argv) > 4 and sys.argv[4] == "bylines":
    print("---- by line (file filter applied), warp-inst % and stall-sample % ----")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: (kv[0][0], int(kv[0][1]))):
        if want and want not in f: continue
        if v[0] / ti < 0.002 and v[1] / ts < 0.002: continue
        print(f"{v[0]/ti*100:5.1f}%i {v[1]/ts*100:5.1f}%s {f}:{ln:>4} {v[3][:110]}")
