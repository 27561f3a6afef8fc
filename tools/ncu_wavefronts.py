#!/usr/bin/env python3
"""Shared-memory wavefronts of the detection kernel per SASS instruction / code region (ncu source page)."""
import csv, subprocess, sys, collections
rep=sys.argv[1]
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--kernel-name","regex:fdf_detect"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
ia=hdr.index("Address"); isrc=hdr.index("Source"); iw=hdr.index("L1 Wavefronts Shared"); ii=hdr.index("L1 Wavefronts Shared Ideal"); ie=hdr.index("Instructions Executed")
base=int(rows[2][ia],16)
tot=0; items=[]
for r in rows[2:]:
    try: w=int(r[iw]); wi=int(r[ii]); n=int(r[ie])
    except: continue
    if w==0: continue
    tot+=w; items.append((int(r[ia],16)-base,w,wi,n,r[isrc].strip()))
print("total wavefronts",tot)
reg=collections.OrderedDict()
for a,w,wi,n,s in items:
    k=a//0x400
    d=reg.setdefault(k,[0,0,0]); d[0]+=w; d[1]+=wi; d[2]+=n
for k,d in reg.items(): print(f"{k*0x400:6x} wavefronts {100*d[0]/tot:5.1f}% ideal {100*d[1]/tot:5.1f}%  per-inst {d[0]/max(1,d[2]):.2f}")
for a,w,wi,n,s in sorted(items,key=lambda x:-x[1])[:45]: print(f"{a:6x} {100*w/tot:5.2f}% ideal {100*wi/tot:5.2f}% per-inst {w/max(n,1):5.2f} n {n:8d}  {s[:60]}")
