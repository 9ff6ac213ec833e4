"""Turn the scratch outputs of tools/final_artifacts.sh (gpurun_out/<tag>_*) into the committed profiles/ files.

    python tools/collect_profiles.py r01 "<note for the ncu summaries>"
Needs the ncu CLI (no GPU): the .ncu-rep files are summarised, not copied.
"""
import collections, csv, json, os, shutil, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
note = sys.argv[2] if len(sys.argv) > 2 else ""
src, dst = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")

def run_summary(rep, out, text):
    subprocess.run([sys.executable, os.path.join(REPO, "tools", "summarize_ncu.py"), os.path.join(src, rep), os.path.join(dst, out), text], check=True,
                   stdout=subprocess.DEVNULL)

run_summary(f"{tag}_prof_bench_mpc.ncu-rep", f"{tag}_mpc_ncu_summary",
            f"{note} mpc_solve_kernel<10> inside `python bench.py` (4096 envs, BASELINE config[1]); ncu --set full --clock-control none. "
            "dram_bytes_per_launch feeds bench.py roofline.traffic.")
run_summary(f"{tag}_prof_mpc_65536.ncu-rep", f"{tag}_mpc_ncu_summary_65536", f"{note} same kernel at 65536 envs (tools/perf_mpc.py): steady state, 55 waves.")
if os.path.exists(os.path.join(src, f"{tag}_prof_mpc_h20.ncu-rep")):
    run_summary(f"{tag}_prof_mpc_h20.ncu-rep", f"{tag}_mpc_ncu_summary_h20", f"{note} lean kernel at h = 20 (Riccati sweep), 16384 envs (tools/perf_mpc.py).")

# launch list -> per-kernel shares
rows = list(csv.reader(open(os.path.join(src, f"{tag}_launches.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    us = v / 1e3 if r[mu] in ("ns", "nsecond") else (v if r[mu] in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
own = [k for k in agg if "mpc_solve_kernel" in k or "mpc_fallback_kernel" in k or "step_prologue" in k or "step_epilogue" in k]
step_kernels = {k: v for k, v in agg.items() if "mpc_solve_kernel" in k or "mpc_fallback_kernel" in k}
with open(os.path.join(dst, f"{tag}_launch_list_summary.md"), "w") as fh:
    fh.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 1` ({tag})\n\n"
             "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` -- cold-cache, serialised launches: compare SHARES, not absolute times.\n"
             f"Raw CSV: profiles/{tag}_launches.csv.  A timed bench step is one launch of the lean `mpc_solve_kernel<10, 1>` plus one of "
             "`mpc_fallback_kernel<10>` (empty queue on the trot batch: a wave of immediate exits); "
             "the other launches below belong to the un-timed parts of bench.py (FMA-peak probes, the latency / config-4 / config-5 extras, L2 flush memsets, torch fills).\n\n"
             "| kernel | launches | total us | share of all launches % |\n|---|---|---|---|\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"| `{k[:110]}` | {c} | {t:.1f} | {100 * t / tot:.1f} |\n")
shutil.copy(os.path.join(src, f"{tag}_launches.csv"), os.path.join(dst, f"{tag}_launches.csv"))
for name in ("bench_n1.json", "bench_reference_n1.json", "configs.json", "timeline_4096.log", "pytest_gpu.log", "config1_substitute.json",
             "chol_bench.log", "smoke.log"):
    if os.path.exists(os.path.join(src, f"{tag}_{name}")):
        shutil.copy(os.path.join(src, f"{tag}_{name}"), os.path.join(dst, f"{tag}_{name}"))
shutil.copy(os.path.join(src, f"{tag}_racecheck.log"), os.path.join(dst, f"{tag}_compute_sanitizer_racecheck.log"))
shutil.copy(os.path.join(src, f"{tag}_memcheck.log"), os.path.join(dst, f"{tag}_compute_sanitizer_memcheck.log"))
for tool in ("synccheck", "initcheck"):
    if os.path.exists(os.path.join(src, f"{tag}_{tool}.log")):
        shutil.copy(os.path.join(src, f"{tag}_{tool}.log"), os.path.join(dst, f"{tag}_compute_sanitizer_{tool}.log"))
print("profiles/ updated:", sorted(f for f in os.listdir(dst) if f.startswith(tag + "_")))
