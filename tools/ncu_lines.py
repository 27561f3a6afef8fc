#!/usr/bin/env python3
"""Per-CUDA-source-line instruction / stall-sample totals from `ncu -i rep --page source --csv --print-source cuda,sass`."""
import csv
import subprocess
import sys


def main(rep, top=60, kernel=None):
    kern = ["--kernel-name", "regex:" + kernel] if kernel else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    lines = []  # (file, line, src, inst, thread_inst, samples)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            ii, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and r[0] not in ("", "Function Name") and r[0].isdigit():
            try:
                lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ii]), int(r[it]), int(r[isamp])))
            except ValueError:
                pass
    tot_i = sum(x[3] for x in lines)
    tot_s = sum(x[5] for x in lines)
    print(f"total warp-inst {tot_i}  samples {tot_s}")
    byfile = {}
    for f, ln, s, i, t, sm in lines:
        a = byfile.setdefault(f, [0, 0, 0])
        a[0] += i; a[1] += t; a[2] += sm
    for f, a in byfile.items():
        print(f"  {f:24s} inst {100*a[0]/tot_i:5.1f}%  lane-eff {a[1]/max(a[0],1)/32:.2f}  samples {100*a[2]/tot_s:5.1f}%")
    for f, ln, s, i, t, sm in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{f:18s}:{ln:4d} inst {100*i/tot_i:5.2f}% lanes {t/max(i,1):5.1f} samp {100*sm/tot_s:5.2f}%  {s[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60, sys.argv[3] if len(sys.argv) > 3 else None)
