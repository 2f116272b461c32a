#!/usr/bin/env python
"""Aggregate an .ncu-rep's per-source-line instruction / stall-sample counts into the kernel's stages.
    python tools/ncu_regions.py gpurun_out/prof_hopper.ncu-rep"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fp, fn, j, body = rows[i][1], rows[i + 1][1], i + 3, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        secs.append((fp, fn, body)); i = j
    else:
        i += 1
lines = open(os.path.join(ROOT, "dart_env_b200", "csrc", "planar_kernels.cuh")).read().split("\n")


def find(pat):
    for k, l in enumerate(lines):
        if pat in l:
            return k + 1
    return None


marks = [(0, "static_for glue (inlined K1-K5 lambdas)"), (find("DEVI void closest_segment_box2"), "K4 segment-box walk"),
         (find("DEVI bool chol_solve_sub"), "K6 Dantzig fallback"), (find("DEVI bool lcp_small"), "K6 lcp_small (registers)"),
         (find("DEVI bool lcp_bpp_local"), "K6 lcp_bpp_local"), (find("DEVI void lcp_exact"), "K6 dispatch"),
         (find("DEVI void lcp_pgs"), "K6 PGS"), (find("struct ContactSink"), "fk_positions (obs/height)"),
         (find("// ---------------- K1"), "K1 FK"), (find("// ---------------- K2"), "K2 bias + implicit inertia"),
         (find("// ---------------- K3"), "K3 forward pass"), (find("// ---------------- K4/K5"), "K4/K5 collide + contact rows"),
         (find("// joint-limit rows"), "K5 limit rows"), (find("// plain (non-implicit)"), "K5 plain inertia"),
         (find("// M^-1 J^T, one impulse pass"), "K5 impulse passes + A rows"), (find("if (lcp_mode == 1)"), "K6 call / hints"),
         (find("// ---------------- K7"), "K7 apply + contact read-back"), (find("// ---------------- integrate positions"), "integrate")]
marks = [m for m in marks if m[0] is not None]
tot_i = tot_s = 0
agg, other = {}, {}
for fp, fn, body in secs:
    for r in body:
        try:
            ln, ins, sm, ti = int(r[0]), int(r[7]), int(r[6]), int(r[8])
        except (ValueError, IndexError):
            continue
        tot_i += ins; tot_s += sm
        if "planar_kernels" in fp:
            k = [m[1] for m in marks if m[0] <= ln][-1]
        else:
            k = os.path.basename(fp)
        a = agg.setdefault(k, [0, 0, 0]); a[0] += ins; a[1] += sm; a[2] += ti
print("| stage | warp instructions | % | stall samples % | active lanes |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("| %s | %d | %.1f | %.1f | %.1f |" % (k, a[0], 100.0 * a[0] / tot_i, 100.0 * a[1] / max(tot_s, 1), a[2] / max(a[0], 1)))
print("\ntotal warp instructions %d, samples %d" % (tot_i, tot_s))
