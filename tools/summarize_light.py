"""ncu launch list of tools/perf_light.py -> profiles/<tag>_light_kernels.md: per-kernel medians of duration, DRAM
bytes, achieved DRAM GB/s against the measured HBM peak, FP64-pipe and issue utilisation, registers, occupancy.
    python tools/summarize_light.py gpurun_out/r02_light_ncu.csv [more.csv ...] profiles/r02_light_kernels.md"""
import collections, csv, json, os, statistics, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcs, dst = sys.argv[1:-1], sys.argv[-1]
peak = 6558.7
try: peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception: pass
per = collections.OrderedDict()
for src in srcs:
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; kn, mn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    bs = hdr.index("Block Size")
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        name = r[kn].split("(")[0].split("::")[-1]
        grid = int(r[gs].strip("()").split(",")[0]); block = int(r[bs].strip("()").split(",")[0])
        per.setdefault((name, grid, block), {}).setdefault(r[mn], []).append(float(r[mv].replace(",", "")))
# algorithmic bytes per env and threads per env of each kernel
ALG = {"state_from_sim_kernel": (196, 4), "hybrid_motor_kernel": (384, 12), "step_prologue_kernel": (None, 4), "step_prologue_env_kernel": (None, 1),
       "step_epilogue_kernel": (None, 4)}
with open(dst, "w") as fh:
    fh.write("# Light kernels (either side of the MPC solve) against the HBM roofline\n\n"
             f"`ncu --metrics ... --clock-control none` launch list of `python tools/perf_light.py` (medians over the launches of each kernel; "
             f"cold-cache, serialised launches).  HBM peak {peak:.1f} GB/s (MEASURED_PEAKS.json).  `DRAM GB/s` uses the bytes ncu counted "
             "(`dram__bytes_read.sum + dram__bytes_write.sum`); a kernel whose inputs were just written by its predecessor reads them from the 126 MB L2, "
             "and outputs still in L2 when the kernel ends are not counted either.\n\n"
             "| kernel | envs | threads/env | regs | duration us | DRAM MB (read + write) | DRAM GB/s | frac of HBM peak | FP64 pipe % | issue slots % | warps active % | bound by |\n"
             "|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for (name, grid, block), m in per.items():
        med = lambda k: statistics.median(m[k]) if k in m else float("nan")
        tpe = ALG.get(name, (None, 1))[1]
        envs = grid * block // tpe
        dur = med("gpu__time_duration.sum") / 1e3
        rd, wr = med("dram__bytes_read.sum") / 1e6, med("dram__bytes_write.sum") / 1e6
        gbs = (rd + wr) / dur * 1e3      # MB / us = TB/s
        fp64, issue = med("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"), med("smsp__issue_active.avg.pct_of_peak_sustained_active")
        bound = "HBM" if gbs / peak > 0.6 else ("FP64 instruction issue (double-precision trigonometry of the leg chain / IK)" if fp64 > 25 else "launch / memory latency (less than one wave of work)")
        fh.write(f"| `{name}` | ~{envs} | {tpe} | {med('launch__registers_per_thread'):.0f} | {dur:.1f} | {rd:.1f} + {wr:.1f} | {gbs:.0f} | {gbs / peak:.2f} | "
                 f"{fp64:.0f} | {issue:.0f} | {med('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {bound} |\n")
print(open(dst).read())
