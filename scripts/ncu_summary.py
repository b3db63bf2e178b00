"""Turns one `ncu --set full --import-source on` capture (.ncu-rep) of the two tensor kernels into the tracked summaries under
profiles/:  <tag>_raw.txt (key counters per kernel), source_sampling_<tag>.md (warp-state samples per warp role: the code
between two setmaxnreg instructions) and the DRAM traffic entries of profiles/ncu_traffic.json.

  python scripts/ncu_summary.py gpurun_out/prof_c3_r2.ncu-rep r2 c3 [scale]

`scale` multiplies the per-launch byte counts (a capture at N = 4M scaled to the 10M-event workload would pass 2.5; default 1).
Runs here (no GPU needed): ncu only reads the report."""
import csv
import io
import json
import os
import subprocess
import sys

rep, tag, workload = sys.argv[1], sys.argv[2], sys.argv[3]
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
ix = {h: i for i, h in enumerate(hdr)}
out = [f"# key counters of {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n"]
traffic = {}
for r in raw[2:]:
    name = r[ix["Kernel Name"]]
    short = "estep" if "estep_tc" in name else ("mstep" if "mstep_tc_kernel" in name else None)
    out.append(f"\n## {name.split('(')[0]}\n")
    for w in WANT:
        if w in ix:
            out.append(f"{w:78s} {r[ix[w]]:>18s} {units[ix[w]]}\n")
    if short:
        b = sum(float(r[ix[k]]) * UNIT.get(units[ix[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[f"{workload}:tensor:{short}"] = round(b * scale)
        out.append(f"DRAM read + write per launch: {b / 1e9:.3f} GB" + (f" (x{scale} = {b * scale / 1e9:.3f} GB at the bench workload)\n" if scale != 1 else "\n"))
open(os.path.join(ROOT, "profiles", f"ncu_{workload}_{tag}_raw.txt"), "w").write("".join(out))

tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
tj = json.load(open(tp)) if os.path.exists(tp) else {}
tj.update(traffic)
tj["_source"] = f"profiles/ncu_{workload}_{tag}_raw.txt ({os.path.basename(rep)})"
json.dump(tj, open(tp, "w"), indent=1)

md = [f"# warp-state samples per warp role — {os.path.basename(rep)}\n\n",
      "Roles = code sections between two `setmaxnreg` instructions (in address order); the last section of each kernel also holds the\n",
      "out-of-line parked mbarrier waits and the exit barrier.  Columns: share of the kernel's samples, warp instructions executed, top stall reasons.\n"]
for kern in ("estep_tc_kernel", "mstep_tc_kernel"):
    rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--kernel-name", f"regex:{kern}"))))
    if len(rows) < 3:
        continue
    h = rows[1]
    data = [r for r in rows[2:] if len(r) > 10]
    jx = {k: i for i, k in enumerate(h)}
    n = len(data)
    for i in range(1, n):
        if data[i][jx["Address"]] == data[0][jx["Address"]]:
            n = i
            break
    data = data[:n]
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]

    def gi(r, k):
        try:
            return int(float(r[jx[k]] or 0))
        except ValueError:
            return 0
    tot = sum(gi(r, "# Samples") for r in data)
    segs, cur = [], None

    def new(i, name):
        global cur
        cur = dict(start=i, name=name, n=0, inst=0, st={s: 0 for s in stalls}, marks=set())
        segs.append(cur)
    new(0, "prologue")
    for i, r in enumerate(data):
        src = r[jx["Source"]]
        if "USETMAXREG" in src:
            new(i, src.strip().split()[0] + " " + src.strip().split()[-1])
        cur["n"] += gi(r, "# Samples")
        cur["inst"] += gi(r, "Instructions Executed")
        for s in stalls:
            cur["st"][s] += gi(r, s)
        for k in ("UTMALDG", "UTCHMMA", "LDTM", "STTM", "STS.128", "F2FP", "STG", "MUFU.EX2", "FFMA2", "LDG"):
            if k in src:
                cur["marks"].add(k)
    md.append(f"\n## {kern} ({n} SASS instructions, {tot} samples)\n\n| section | marker instructions | samples | warp-inst | top stall reasons |\n|---|---|---|---|---|\n")
    for s in segs:
        top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:6]
        md.append(f"| {s['name']} | {' '.join(sorted(s['marks']))} | {100 * s['n'] / max(1, tot):.1f} % | {s['inst']} | "
                  + ", ".join(f"{k[6:]} {100 * v / max(1, s['n']):.0f} %" for k, v in top) + " |\n")
    hot = sorted(((gi(r, "# Samples"), i, r) for i, r in enumerate(data)), reverse=True)[:12]
    md.append("\nHottest instructions:\n\n| samples | executed | top stall | SASS |\n|---|---|---|---|\n")
    for nn, i, r in hot:
        st = max(stalls, key=lambda s: gi(r, s))
        md.append(f"| {nn} | {gi(r, 'Instructions Executed')} | {st[6:]} {gi(r, st)} | `{r[jx['Source']].strip()[:90]}` |\n")
open(os.path.join(ROOT, "profiles", f"source_sampling_{tag}.md"), "w").write("".join(md))
print("written", f"profiles/ncu_{workload}_{tag}_raw.txt", f"profiles/source_sampling_{tag}.md", traffic)
