"""Summarise an `ncu --set full` report (one kernel launch) into profiles/<name>.json + .md.

    python tools/summarize_ncu.py gpurun_out/r01_prof_mpc.ncu-rep profiles/r01_mpc_ncu_summary "<note>"
Needs the ncu CLI (no GPU).  Keeps: duration, DRAM bytes (-> bench.py roofline.traffic), occupancy limits,
pipe utilisation, executed FP64 instruction counts, stall-reason shares, and the hottest source lines.
"""
import collections, csv, io, json, subprocess, sys

def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep, out_base = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = ncu_csv(rep, "--page", "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    def num(k):
        try: return float(m[k][0].replace(",", ""))
        except Exception: return None
    def scaled_bytes(k):
        v, u = num(k), m.get(k, ("", ""))[1]
        if v is None: return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second"]
    summary = {"report": rep, "note": note, "kernel": rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else None,
               "metrics": {k: {"value": num(k), "unit": m[k][1]} for k in keys if k in m}}
    rd, wr = scaled_bytes("dram__bytes_read.sum"), scaled_bytes("dram__bytes_write.sum")
    summary["dram_bytes_read"], summary["dram_bytes_write"] = rd, wr
    summary["dram_bytes_per_launch"] = (rd or 0) + (wr or 0)
    dur_ms = num("gpu__time_duration.sum")
    if m.get("gpu__time_duration.sum", ("", ""))[1] == "us": dur_ms = dur_ms / 1e3
    grid = num("launch__grid_size")
    # ncu's full set records these per elapsed cycle: multiply back by the elapsed SM cycles
    cycles = num("sm__cycles_elapsed.max") or num("sm__cycles_elapsed.avg") or 0
    dfma, dmul, dadd = ((num("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k) or 0) * cycles
                        for k in ("dfma", "dmul", "dadd"))
    summary["duration_ms"] = dur_ms
    summary["executed_fp64_flops_per_launch"] = 2 * dfma + dmul + dadd
    if grid: summary["executed_fp64_flops_per_env"] = (2 * dfma + dmul + dadd) / grid
    if dur_ms: summary["executed_fp64_tflops"] = (2 * dfma + dmul + dadd) / (dur_ms * 1e-3) / 1e12
    # stall reasons over all SASS instructions
    src = ncu_csv(rep, "--page", "source")
    h2 = src[1]
    stall_cols = [i for i, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in src[2:]:
        if len(r) < len(h2): continue
        for i in stall_cols:
            try: tot[h2[i]] += int(r[i])
            except ValueError: pass
    s = sum(tot.values()) or 1
    summary["sass_instructions"] = len(src) - 2
    summary["stall_share_pct"] = {k: round(100 * v / s, 2) for k, v in tot.most_common(8)}
    # hottest source lines
    cs = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    samples, inst, text = collections.Counter(), collections.Counter(), {}
    for r in cs:
        if len(r) >= 8 and r[0].isdigit():
            ln = int(r[0]); text[ln] = r[1].strip()
            try: samples[ln] += int(r[4]); inst[ln] += int(r[7])
            except ValueError: pass
    ts, ti = sum(samples.values()) or 1, sum(inst.values()) or 1
    summary["hot_lines"] = [{"line": ln, "samples_pct": round(100 * c / ts, 1), "inst_pct": round(100 * inst[ln] / ti, 1), "source": text[ln][:120]}
                            for ln, c in samples.most_common(15)]
    with open(out_base + ".json", "w") as fh: json.dump(summary, fh, indent=1)
    with open(out_base + ".md", "w") as fh:
        fh.write(f"# ncu summary: {summary['kernel']}\n\n{note}\n\nreport: `{rep}` (scratch, not committed)\n\n")
        fh.write(f"* duration {dur_ms:.3f} ms for grid {int(grid or 0)} CTAs; DRAM read {rd/1e6:.2f} MB + write {wr/1e6:.2f} MB per launch\n")
        fh.write(f"* executed FP64: {summary['executed_fp64_flops_per_launch']:.3e} flop per launch "
                 f"({summary.get('executed_fp64_flops_per_env', 0):.3e} per env) = {summary.get('executed_fp64_tflops', 0):.2f} TFLOP/s\n\n")
        fh.write("| metric | value | unit |\n|---|---|---|\n")
        for k, v in summary["metrics"].items(): fh.write(f"| {k} | {v['value']} | {v['unit']} |\n")
        fh.write("\n| stall reason | share of samples % |\n|---|---|\n")
        for k, v in summary["stall_share_pct"].items(): fh.write(f"| {k} | {v} |\n")
        fh.write("\n| line | samples % | instructions % | source |\n|---|---|---|---|\n")
        for h in summary["hot_lines"]: fh.write(f"| {h['line']} | {h['samples_pct']} | {h['inst_pct']} | `{h['source']}` |\n")
    print(json.dumps({k: summary[k] for k in ("duration_ms", "dram_bytes_per_launch", "executed_fp64_tflops", "stall_share_pct")}))

if __name__ == "__main__":
    main()
