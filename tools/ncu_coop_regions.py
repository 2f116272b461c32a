#!/usr/bin/env python
"""Per-stage instruction / stall-sample attribution of the lane-cooperative kernel from an .ncu-rep
(source page through -lineinfo), plus the opcode mix.    python tools/ncu_coop_regions.py REP"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]


def page(args):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + args, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page(["--print-source", "sass,cuda"])
secs, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fp, hdr, j, body = rows[i][1], rows[i + 2], i + 3, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        secs.append((fp, hdr, body)); i = j
    else:
        i += 1
src = open(os.path.join(ROOT, "dart_env_b200", "csrc", "planar_coop.cuh")).read().split("\n")
PATS = [("inline void coop_lane_init", "lane table (host)"), ("DEVI V gshfl", "collectives"), ("DEVI void anc_prefix", "anc_prefix"),
        ("DEVI RM mass_rsqrt", "mass_rsqrt"), ("DEVI void ltl_factor", "ltl_factor"), ("DEVI void ltl_solve_t", "ltl_solve_t"),
        ("DEVI void ltl_solve(", "ltl_solve"), ("static DEVI void gather_rows", "K6 gather_rows"),
        ("static DEVI bool exchange", "K6 exchange"), ("static DEVI bool solve", "K6 solve"), ("DEVI void coop_pgs", "K6 pgs"),
        ("DEVI void coop_fk_positions", "fk_positions (height)"), ("DEVI void coop_constraints", "K5 head"),
        ("DEVI void coop_substep", "K1 kinematics"), ("// ---------------- K2: per-body", "K2 wrench"),
        ("// ---------------- K2/K2'", "K2' subtree sums"), ("// bias force and F_i", "K2' M rows + gather"),
        ("// ---------------- K3", "K3 FD solve"), ("// ---------------- K4", "K4 collide"),
        ("// joint limits: q BEFORE", "limits + class dispatch"), ("// ---------------- integrate", "integrate"),
        ("// ---------------- K5: compact", "K5 rows / J / Y"), ("// A[r][s] = Y_r", "K5 A"), ("// ---------------- K6", "K6 call"),
        ("// ---------------- K7", "K7 apply"), ("// stick/slide sets", "hint + contact read-back"),
        ("COOP_GLOBAL void k_substep_coop", "k_substep_coop"), ("COOP_GLOBAL void k_env_step_coop", "env-step prologue / task layer")]
marks = sorted((k + 1, name) for k, l in enumerate(src) for pat, name in PATS if pat in l)
agg, tot_i, tot_s = collections.OrderedDict(), 0, 0
for fp, hdr, body in secs:
    ix = {h: k for k, h in enumerate(hdr)}
    for r in body:
        try:
            ln, ins, sm = int(r[0]), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except (ValueError, IndexError, KeyError):
            continue
        base = os.path.basename(fp)
        if base == "planar_coop.cuh":
            name = ([m[1] for m in marks if m[0] <= ln] or ["head"])[-1]
        elif base == "planar_kernels.cuh":
            name = "planar_kernels.cuh: " + ("static_for glue" if ln < 70 else ("Num<> math" if ln < 160 else ("Philox reset" if ln < 180 else ("segment-box walk" if ln < 232 else "other"))))
        else:
            name = base
        a = agg.setdefault(name, [0, 0]); a[0] += ins; a[1] += sm
        tot_i += ins; tot_s += sm
print("| stage | warp instructions % | stall samples % |\n|---|---|---|")
for k, (a, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %.1f | %.1f |" % (k, 100.0 * a / max(tot_i, 1), 100.0 * b / max(tot_s, 1)))
rows = page(["--print-source", "sass"])
hdr = rows[1]; ix = {h: k for k, h in enumerate(hdr)}
data = rows[2:]
ex, ns = ix["Instructions Executed"], ix["# Samples"]
tot = sum(int(r[ex]) for r in data); tots = sum(int(r[ns]) for r in data)
op, ops = collections.Counter(), collections.Counter()
for r in data:
    t = r[1].strip().split()
    if not t:
        continue
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    op[o] += int(r[ex]); ops[o] += int(r[ns])
print("\nexecuted warp instructions %d (static SASS %d, distinct executed %d)\n" % (tot, len(data), sum(1 for r in data if int(r[ex]) > 0)))
print("| opcode | executed % | samples % |\n|---|---|---|")
for o, c in op.most_common(16):
    print("| %s | %.1f | %.1f |" % (o, 100.0 * c / tot, 100.0 * ops[o] / max(tots, 1)))
st = collections.Counter()
for r in data:
    for h, k in ix.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            st[h[6:]] += int(r[k] or 0)
tt = sum(st.values())
print("\nstalls: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tt) for k, v in st.most_common(8)))
