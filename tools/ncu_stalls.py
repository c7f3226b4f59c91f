#!/usr/bin/env python
"""Per-kernel top stall lines.  usage: tools/ncu_stalls.py source_page.csv [topn]"""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1]))); topn=int(sys.argv[2]) if len(sys.argv)>2 else 12
kern=None; hdr=None; agg=collections.defaultdict(lambda: collections.defaultdict(lambda:[0,0,'',{}]))
for r in rows:
    if not r: continue
    if r[0]=="Function Name": kern=r[1][30:75]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0]=="File Path": continue
    if hdr is None or len(r)<len(hdr)-5 or r[2]!='-': continue
    try: inst=float(r[hdr.index("Instructions Executed")]); samp=float(r[hdr.index("Warp Stall Sampling (All Samples)")])
    except: continue
    d=agg[kern][r[0]]; d[0]+=inst; d[1]+=samp; d[2]=r[1].strip()
    for nm in ("stall_long_sb","stall_mio","stall_barrier","stall_wait","stall_short_sb","stall_math","stall_not_selected","stall_membar","stall_branch_resolving","stall_lg","stall_selected","stall_sleep"):
        try: d[3][nm]=d[3].get(nm,0)+float(r[hdr.index(nm)])
        except: pass
for k,v in agg.items():
    ts=sum(x[1] for x in v.values()); ti=sum(x[0] for x in v.values())
    tot=collections.Counter()
    for x in v.values():
        for a,b in x[3].items(): tot[a]+=b
    print("==",k,"samples",ts,"warp-inst %.3e"%ti, "| by reason:", ", ".join(f"{a[6:]}={b/ts*100:.0f}%" for a,b in tot.most_common(7)))
    for ln,x in sorted(v.items(), key=lambda kv:-kv[1][1])[:topn]:
        st=sorted(x[3].items(), key=lambda kv:-kv[1])[:2]
        print(f"  {x[1]/ts*100:5.1f}%s {x[0]/ti*100:5.1f}%i line {ln:>4} {x[2][:88]} [{', '.join(a[6:]+'='+str(int(b)) for a,b in st)}]")
