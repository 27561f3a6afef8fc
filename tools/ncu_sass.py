#!/usr/bin/env python3
"""SASS-level view of one kernel of an ncu report: per instruction executed count, lanes, stall samples and the top
stall reason, with the CUDA source line it belongs to.  usage: ncu_sass.py report.ncu-rep [kernel regex] [min samples]"""
import csv, subprocess, sys
rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "fdf_detect"
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, line, src_text = None, None, None, ""
sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        ith = hdr.index("Thread Instructions Executed")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and r[0].isdigit(): line = int(r[0])
    elif hdr and r[0] == "" and len(r) > isamp and r[ia].startswith("0x"):
        n, s, th = int(r[ii] or 0), int(r[isamp] or 0), int(r[ith] or 0)
        st = sorted(((int(r[i] or 0), nm) for i, nm in stall_cols), reverse=True)[:2]
        sass.append((int(r[ia], 16), cur_file, line, n, th, s, st, r[ia + 1].strip()))
sass.sort()
tot_i = sum(x[3] for x in sass); tot_s = sum(x[5] for x in sass)
print(f"total warp-inst {tot_i}  samples {tot_s}")
base = sass[0][0]
for addr, fn, ln, n, th, s, st, txt in sass:
    if s < mins: continue
    lanes = th / n if n else 0
    print(f"{addr - base:6x} {fn or '?':16.16s}:{ln or 0:4d} inst {100 * n / tot_i:5.2f}% lanes {lanes:4.1f} samp {100 * s / tot_s:5.2f}%  "
          f"{st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  {txt[:70]}")
