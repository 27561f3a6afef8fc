#!/usr/bin/env python3
"""Attributes executed warp-instructions / stall samples of the detection kernel to its phases.
SASS instructions are walked in address order; helper lines (one-instruction wrappers in fdf_core.cuh)
inherit the phase of the nearest preceding instruction that belongs to a phase-specific source line."""
import csv, subprocess, sys, collections, re

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "fdf_detect"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))

def I(x):
    try: return int(x)
    except ValueError: return 0

src_lines = {}
for fn in ("fdf_core.cuh", "fdf_strip.cuh", "fdf_kernels.cu"):
    src_lines[fn] = open("feature_detector_fast_b200/csrc/" + fn).read().splitlines()

def func_of(fn, ln):
    """name of the enclosing function (nearest preceding line that looks like a definition)"""
    L = src_lines.get(fn)
    if not L: return None
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = re.match(r'^(?:template.*>\s*)?(?:FDF_HD|__global__|__device__|static|inline)[^;(]*?\b(\w+)\s*\(', L[i])
        if m: return m.group(1)
        m = re.match(r'^fdf_detect_kernel\(', L[i])
        if m: return "fdf_detect_kernel"
    return None

HELPERS = {"max3u", "mad32", "absdiff4", "byte_perm", "popc32", "highest_set_bit", "swap16", "min_u16x2", "max_u16x2",
           "min3_u16x2", "max3_u16x2", "addrelu_s16x2", "exceeds4", "dual_word", "dual_bias", "best_of_lanes", "load16", "atomic_add_u32", "atomic_or_u32",
           "lowest_set_bit", "live_score", "filter_kbias", "mask_bit_to_px", "smem_u32"}
PHASE = {"vertical_any": "A1 stage1", "stage1_lane": "A1 stage1", "candidate_mask16": "A2 stage2", "stage2_entry": "A2 stage2", "stage2_mask": "A2 stage2", "push_candidates": "A2 push", "best_window": "B test", "best_window_k": "B test", "score_sum_abs_dual": "B score",
         "phase_a_warp": "A1 stage1", "ring_masks": "B test", "has_arc": "B test", "phase_b": "B test",
         "score_max_threshold": "B score", "max_of_extended": "B score", "score_sum_abs": "B score",
         "nms_emits": "NMS+stage", "nms_is_max": "NMS+stage", "emit_list": "NMS+stage", "nms_dense": "NMS+stage",
         "staged_entry": "NMS+stage", "stage1_band": "A1 stage1", "live_mask": "A1 stage1", "live_rows": "A1 stage1",
         "close_run": "NMS+stage", "open_run": "NMS+stage", "reserve_staging": "NMS+stage",
         "fdf_detect_kernel": "main loop", "mbar_wait": "main loop", "mbar_try_wait": "main loop", "tma_load_3d": "main loop",
         "make_geo": "main loop", "valid_word": "main loop", "vtab_variant": "main loop"}
cur_file, hdr, line = None, None, None
sass = []  # (addr, file, line, inst, samples, src)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        isrc = ia + 1
    elif hdr and r[0].isdigit(): line = int(r[0])
    elif hdr and r[0] == "" and len(r) > isamp and r[ia].startswith("0x"):
        sass.append((int(r[ia], 16), cur_file, line, I(r[ii]), I(r[isamp]), r[isrc].strip()))
sass.sort()
inst, samp = collections.Counter(), collections.Counter()
phase = "main loop"
for addr, fn, ln, n, s, src in sass:
    f = func_of(fn, ln) if fn in src_lines else None
    if f and f not in HELPERS and f in PHASE:
        phase = PHASE[f]
        if phase == "A1 stage1" and f == "phase_a_warp" and "stage2_entry" in src_lines[fn][ln - 1]: phase = "A2 stage2"
    inst[phase] += n
    samp[phase] += s
ti, ts = sum(inst.values()), sum(samp.values())
print(f"warp-instructions {ti}, samples {ts}")
for ph in sorted(inst, key=lambda k: -inst[k]):
    print(f"  {ph:12s} inst {100 * inst[ph] / ti:5.1f}%   samples {100 * samp[ph] / ts:5.1f}%")
