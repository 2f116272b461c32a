#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md + a JSON entry.

    python tools/ncu_summary.py gpurun_out/prof_hopper.ncu-rep profiles/r1_hopper DartHopper-v1
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out, env_id = sys.argv[1], sys.argv[2], sys.argv[3]
json_key = sys.argv[4] if len(sys.argv) > 4 else env_id   # e.g. "DartHopper-v1/static" keeps an older capture beside the current one
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def tonum(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit)
    return x * mult if mult else x


launches = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in keys:
        if k in hdr:
            d[k] = tonum(r[hdr.index(k)], units[hdr.index(k)])
    launches.append(d)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
stall_tot, per_file = {}, {}
i = 0
shdr = None
first_kernel_done = False
while i < len(srows):
    if srows[i] and srows[i][0] == "File Path":
        fp = srows[i][1]
        shdr = srows[i + 2]
        st = [c for c, h in enumerate(shdr) if h.startswith("stall_") and "Not Issued" not in h]
        j = i + 3
        inst = samp = 0
        while j < len(srows) and not (srows[j] and srows[j][0] == "File Path"):
            r = srows[j]
            try:
                inst += int(r[7]); samp += int(r[6])
                for c in st:
                    stall_tot[shdr[c]] = stall_tot.get(shdr[c], 0) + int(r[c])
            except (ValueError, IndexError):
                pass
            j += 1
        per_file.setdefault(os.path.basename(fp), [0, 0])
        per_file[os.path.basename(fp)][0] += inst
        per_file[os.path.basename(fp)][1] += samp
        i = j
    else:
        i += 1
T = sum(stall_tot.values()) or 1
stalls = {k.replace("stall_", ""): round(100.0 * v / T, 1) for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]) if v > 0}

L = launches[0]
cyc = L.get("sm__cycles_elapsed.max", 0)
flops = cyc * (2 * L.get("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", 0)
               + L.get("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", 0)
               + L.get("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", 0))
dur_s = L["gpu__time_duration.sum"] * 1e-6
summary = {"kernel": L["kernel"][:80], "duration_us": L["gpu__time_duration.sum"],
           "dram_bytes_per_launch": L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0),
           "warp_inst_per_launch": L.get("smsp__inst_executed.sum"), "avg_active_lanes": L.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "ipc_per_sm": L.get("sm__inst_executed.avg.per_cycle_elapsed"), "registers": L.get("launch__registers_per_thread"),
           "grid": L.get("launch__grid_size"), "block": L.get("launch__block_size"),
           "issue_active_pct": L.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "warps_active_pct": L.get("sm__warps_active.avg.pct_of_peak_sustained_active"), "stall_pct": stalls,
           "fp32": {"flop_per_launch": flops, "achieved_tflops_under_ncu": flops / dur_s / 1e12 if dur_s else None,
                    "note": "FADD+FMUL+2*FFMA thread instructions; B200 non-tensor fp32 peak ~ 74 TFLOP/s (148 SM x 128 lanes x 2 x 1.965 GHz)"}}
# fp64 work (the cooperative kernels carry the mass-matrix core in fp64): thread-level DFMA / DMUL / DADD from the SASS page
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(sass)))
if len(sr) > 2:
    h2 = {h: k for k, h in enumerate(sr[1])}
    f64 = 0
    for r in sr[2:]:
        t = r[1].strip().split()
        if not t:
            continue
        o = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        if o in ("DFMA", "DMUL", "DADD"):
            f64 += (2 if o == "DFMA" else 1) * int(r[h2["Predicated-On Thread Instructions Executed"]])
    summary["fp64"] = {"flop_per_launch": f64, "note": "DADD+DMUL+2*DFMA thread instructions (fp64 mass-matrix core of the cooperative kernel)"}
jpath = os.path.join(os.path.dirname(out), os.path.basename(out).split("_")[0] + "_ncu_summary.json")   # r1_..., r2_...
allj = json.load(open(jpath)) if os.path.exists(jpath) else {}
allj[json_key] = summary
json.dump(allj, open(jpath, "w"), indent=1)
with open(out + "_ncu_summary.md", "w") as fh:
    fh.write("# ncu summary: %s\n\nsource: `%s` (`ncu --set full --clock-control none --import-source on`), read with `ncu -i`.\n\n" % (env_id, os.path.basename(rep)))
    fh.write("| metric | " + " | ".join("launch %d" % k for k in range(len(launches))) + " | unit |\n|---|" + "---|" * (len(launches) + 1) + "\n")
    for k in keys:
        if k in hdr:
            u = units[hdr.index(k)]
            u = "byte" if u in ("Kbyte", "Mbyte", "Gbyte") else u     # tonum() already converted these to bytes
            fh.write("| %s | %s | %s |\n" % (k, " | ".join(str(l.get(k)) for l in launches), u))
    fh.write("\nWarp stall sampling (all kernel instances, %% of samples): %s\n" % json.dumps(stalls))
    fh.write("\nInstructions / samples per source file: %s\n" % json.dumps(per_file))
    fh.write("\nDerived: %s\n" % json.dumps(summary, indent=1))
print(json.dumps(summary, indent=1))
