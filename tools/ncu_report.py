#!/usr/bin/env python3
"""Text summary of one kernel of an .ncu-rep (read here, no GPU needed):

    python tools/ncu_report.py gpurun_out/prof_v19.ncu-rep [kernel-regex] > profiles/r02_..._ncu_summary.txt

Sections: headline raw metrics, warp-state distribution, the SASS regions that execute the most warp-instructions
(a region = consecutive instructions with the same execution count), the instructions with the most stall samples and
the shared-memory instructions with the most wavefronts (with the conflict-free ideal)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "fdf_detect_kernel"


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def I(x):
    try:
        return int(float(x.replace(",", "")))
    except ValueError:
        return 0


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = dict(zip(hdr, zip(units, vals)))
print("== %s  (%s)" % (m.get("Kernel Name", ("", "?"))[1], rep.split("/")[-1]))
WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__grid_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
lts__t_bytes.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed""".split()
for k in WANT:
    if k in m:
        print("%-86s %-16s %s" % (k, m[k][0], m[k][1]))
print("\n== warp states (average warps per scheduler in that state, per issued instruction)")
st = sorted(((float(v[1]), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_")
             and k.endswith("_per_issue_active.ratio")), reverse=True)
for v, k in st[:12]:
    print("  %-28s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))

src = page("source")
h = src[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in src[2:] if len(r) >= len(h)]
ex = [I(r[ix["Instructions Executed"]]) for r in data]
sm = [I(r[ix["# Samples"]]) for r in data]
tot, tots = sum(ex), sum(sm)
print("\n== SASS regions by executed warp-instructions (total %d, %d stall samples)" % (tot, tots))
regions, start = [], 0
for n in range(1, len(data) + 1):
    if n == len(data) or abs(ex[n] - ex[start]) > 0.02 * max(ex[start], 1):
        regions.append((start, n - 1, sum(ex[start:n]), sum(sm[start:n])))
        start = n
print("  sass lines     n  executions/inst   share of inst   share of samples   first instruction")
for a, b, e, s in regions:
    if e > 0.008 * tot or s > 0.01 * tots:
        print("  %5d-%5d %4d  %12d   %6.1f %%        %6.1f %%          %s" %
              (a, b, b - a + 1, ex[a], 100.0 * e / tot, 100.0 * s / max(tots, 1), data[a][ix["Source"]].strip()[:48]))
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print("\n== stall samples by reason")
agg = sorted(((sum(I(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)
print("  " + ", ".join("%s %.1f %%" % (s[6:], 100.0 * v / max(tots, 1)) for v, s in agg[:10]))
print("\n== instructions with the most stall samples")
for n in sorted(range(len(data)), key=lambda n: -sm[n])[:16]:
    top = sorted(((I(data[n][ix[s]]), s[6:]) for s in stalls), reverse=True)[0]
    print("  %5d  %5.1f %%  executed %9d  %-46s mostly %s" %
          (n, 100.0 * sm[n] / max(tots, 1), ex[n], data[n][ix["Source"]].strip()[:46], top[1]))
if "L1 Wavefronts Shared" in ix:
    wf = [I(r[ix["L1 Wavefronts Shared"]]) for r in data]
    ideal = [I(r[ix["L1 Wavefronts Shared Ideal"]]) for r in data]
    print("\n== shared-memory instructions by wavefronts (total %d, conflict-free ideal %d)" % (sum(wf), sum(ideal)))
    for n in sorted(range(len(data)), key=lambda n: -wf[n])[:28]:
        print("  %5d  %9d wavefronts (ideal %9d)  %5.2f per instruction   %s" %
              (n, wf[n], ideal[n], wf[n] / max(ex[n], 1), data[n][ix["Source"]].strip()[:50]))
