"""Markdown summary of an `ncu --set full` report (one row per profiled launch): time, tensor / XU / issue utilisation, DRAM bytes
and throughput, occupancy limits, top stall reasons.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--title "..."] > profiles/r02_x_ncu_summary.md
"""
import csv
import subprocess
import sys

path = sys.argv[1]
title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else path
M = ["gpu__time_duration.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
     "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "math_pipe_throttle", "mio_throttle", "not_selected", "branch_resolving", "no_instruction",
         "dispatch_stall", "lg_throttle", "sleeping", "membar", "tex_throttle"]
metrics = M + [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in STALL]
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for i, h in enumerate(hdr):                    # some columns carry a section prefix ("TPC.TriageCompute.sm__pipe_tensor_..."): index by the bare name too
    ix.setdefault(h.split(".", 2)[-1] if h.count(".") >= 3 and not h.startswith(("sm", "dram", "gpu", "launch", "l1tex", "lts")) else h, i)


def val(r, m):
    if m not in ix or r[ix[m]] in ("", "n/a", "no data"):
        return None
    v = float(r[ix[m]].replace(",", ""))
    u = units[ix[m]]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0) if "byte" in u else v * {"ms": 1e3, "ns": 1e-3, "us": 1.0, "s": 1e6}.get(u, 1.0) if m.startswith("gpu__time") else v


print(f"# {title}\n\n`ncu --set full --clock-control none` (each launch replayed in isolation, cold caches: compare SHARES and ratios, not absolute times)\n")
print("| # | kernel | grid x block | regs | time us | tcgen05 tensor active % (sm__mem_tensor_cycles_active) | HMMA (mma.sync) issue % | XU (MUFU) % | issue active % | warps active % | DRAM read MB | DRAM write MB | DRAM % of peak | warp instr (M) | top stalls (per issued instr) |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for n, r in enumerate(rows[2:]):
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    st = sorted(((val(r, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio") or 0.0, s) for s in STALL), reverse=True)[:3]
    f = lambda m, d=1: "" if val(r, m) is None else f"{val(r, m):.{d}f}"
    print(f"| {n} | `{name[:60]}` | {f('launch__grid_size', 0)} x {f('launch__block_size', 0)} | {f('launch__registers_per_thread', 0)} | {f('gpu__time_duration.sum')} | "
          f"{f(M[1])} | {f(M[2])} | {f(M[3])} | {f(M[4])} | {f(M[5])} | {(val(r, M[6]) or 0) / 1e6:.1f} | {(val(r, M[7]) or 0) / 1e6:.1f} | {f(M[8])} | "
          f"{(val(r, M[12]) or 0) / 1e6:.1f} | " + ", ".join(f"{s} {v:.2f}" for v, s in st) + " |")
