#!/usr/bin/env python
"""Stage-level time attribution of the lane-cooperative kernel: joins ncu's per-SASS-instruction stall
samples (.ncu-rep, source page) with the inline call chains nvdisasm reads from -lineinfo, so that
instructions of inlined helpers (shuffles, static_for lambdas, LTL solves) are charged to the stage
that called them.   python tools/ncu_stage_time.py REP OBJ KERNEL_SUBSTRING"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(ROOT, "dart_env_b200", "csrc", "planar_coop.cuh")).read().split("\n")
PATS = [("DEVI void coop_fk_positions", "task: height FK"), ("DEVI void coop_constraints", "K5 rows / J / Y"),
        ("DEVI void coop_substep", "K1 kinematics"), ("// ---------------- K2: per-body", "K2 wrench"),
        ("// ---------------- K2/K2'", "K2' subtree sums"), ("// prismatic axes of the ancestors", "K2' M rows + gather"),
        ("// ---------------- K3", "K3 factor + FD solve"), ("// ---------------- K4", "K4 collide"),
        ("// joint limits: q BEFORE", "limits / ballots / class dispatch"), ("// ---------------- integrate", "integrate"),
        ("// ---------------- K5: compact", "K5 rows / J / Y"), ("// A[r][s] = Y_r", "K5 A"), ("// ---------------- K6", "K6 LCP"),
        ("// ---------------- K7", "K7 apply"), ("// stick/slide sets", "hint + contact read-back"),
        ("COOP_GLOBAL void k_substep_coop", "k_substep_coop"), ("COOP_GLOBAL void k_env_step_coop", "task: prologue / reward / obs")]
marks = sorted((k + 1, name) for k, l in enumerate(src) for pat, name in PATS if pat in l)
first_stage_line = min(m[0] for m in marks)


def region(line):
    if line < first_stage_line:
        return None      # helper (collectives, LTL, LCP struct ...): charge the caller
    return [m[1] for m in marks if m[0] <= line][-1]


lcp_lo = next(k + 1 for k, l in enumerate(src) if "struct CoopLcp" in l)
lcp_hi = next(k + 1 for k, l in enumerate(src) if "DEVI void coop_pgs" in l)
LCP_PATS = [("static DEVI void gather_rows", "gather_rows"), ("static DEVI bool exchange", "exchange"), ("static DEVI bool solve", "init"),
            ("if (stage == 1) {", "stage-2 setup"), ("int best = NC + 1, tries = 3;", "iteration control"), ("unsigned flip = 0;", "flip ballots"),
            ("while (__any_sync(COOP_FULL, flip != 0))", "exchange loop"), ("R z[RPL], zg[NC], y[RPL];", "z gather + matvec"),
            ("const R xs = group_max<G>(axm)", "xs / S reductions"), ("unsigned badm = 0, nst[RPL];", "feasibility check + ballots"),
            ("const int nbad = __popc(badm);", "set update"), ("// one round of iterative refinement", "refinement")]
lcp_marks = sorted((k + 1, name) for k, l in enumerate(src) for pat, name in LCP_PATS if pat in l and lcp_lo <= k + 1 < lcp_hi)
stage_of = {}
infun, chain, pending = False, [], []
ann = re.compile(r'//## File "([^"]+)", line (\d+)')
ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
for l in sass:
    if l.startswith("//----") or ".section" in l:
        if ".text." in l:
            infun = kern in l
        continue
    if not infun:
        continue
    m = ann.search(l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = ins.search(l)
    if m:
        off = int(m.group(1), 16)
        if pending:            # a new annotation group replaces the call chain; none = same as the previous instruction
            chain, pending = pending, []
        st = None
        in_lcp = any(f == "planar_coop.cuh" and lcp_lo <= ln < lcp_hi for f, ln in chain)
        for f, ln in chain:                    # innermost first: the first line that belongs to a stage (helpers have none)
            if f == "planar_coop.cuh" and region(ln):
                st = region(ln)
                break
        if in_lcp:
            st = "K6 LCP"
            # finer: which part of CoopLcp (innermost line inside the struct)
            for f, ln in chain:
                if f == "planar_coop.cuh" and lcp_lo <= ln < lcp_hi:
                    sub = [m[1] for m in lcp_marks if m[0] <= ln]
                    st = "K6 LCP: " + (sub[-1] if sub else "gather_rows / exchange")
                    break
        stage_of[off] = (st or "other", m.group(2).split()[0] if not m.group(2).startswith("@") else m.group(2).split()[1])
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                                                 capture_output=True, text=True).stdout)))
hdr = rows[1]; ix = {h: k for k, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
agg = collections.OrderedDict()
tot_i = tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in data:
    off = int(r[0], 16) - base
    st, op = stage_of.get(off, ("unmapped", "?"))
    a = agg.setdefault(st, [0, 0, collections.Counter(), collections.Counter()])
    e, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    a[0] += e; a[1] += s; tot_i += e; tot_s += s
    a[3][op.split(".")[0]] += e
    for h in stall_cols:
        a[2][h[6:]] += int(r[ix[h]] or 0)
print("| stage | warp instructions % | time (stall samples) % | top stall reasons | top opcodes |\n|---|---|---|---|---|")
for k, (e, s, stc, opc) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tt = sum(stc.values()) or 1
    print("| %s | %.1f | %.1f | %s | %s |" % (k, 100.0 * e / tot_i, 100.0 * s / tot_s,
                                             ", ".join("%s %.0f%%" % (n, 100.0 * v / tt) for n, v in stc.most_common(3)),
                                             ", ".join("%s %.0f%%" % (n, 100.0 * v / max(e, 1)) for n, v in opc.most_common(4))))
print("\ntotal executed warp instructions %d, samples %d" % (tot_i, tot_s))
