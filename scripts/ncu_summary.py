"""Turns a .ncu-rep into the small text summary committed under profiles/ (raw metrics + hot lines).

    python scripts/ncu_summary.py gpurun_out/prof_path.ncu-rep profiles/r01_path_mega.md "title"
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = list(csv.reader(ncu(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    hdr, units = raw[0], raw[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on), read with `ncu -i ... --page raw/source --csv`", ""]
    for k, row in enumerate(raw[2:]):
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines += [f"## launch {k}: `{name}`", "", "| metric | unit | value |", "|---|---|---|"]
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"| {w} | {units[i]} | {row[i]} |")
        lines.append("")
    src = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv"]).splitlines()))
    # the source page holds one table per kernel launch; take the first
    try:
        h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
        hdr2 = src[h]
        data = []
        for r in src[h + 1:]:
            if len(r) != len(hdr2):
                break
            data.append(r)
        ix = {n: i for i, n in enumerate(hdr2)}
        f = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ("", "-") else 0.0
        tot_s = sum(f(r, "# Samples") for r in data) or 1.0
        tot_i = sum(f(r, "Instructions Executed") for r in data) or 1.0
        lines += ["## hottest SASS lines (warp-state samples)", "", f"total samples {int(tot_s)}, warp instructions {int(tot_i)}", "",
                  "| line | SASS | samples % | warp instr | avg threads | top stall reasons |", "|---|---|---|---|---|---|"]
        stall = [k for k in hdr2 if k.startswith("stall_") and "Not Issued" not in k]
        top = sorted(range(len(data)), key=lambda n: -f(data[n], "# Samples"))[:40]
        for n in sorted(top):
            r = data[n]
            st = sorted(((k[6:], f(r, k)) for k in stall), key=lambda kv: -kv[1])[:3]
            lines.append(f"| {n} | `{r[ix['Source']].strip()[:70]}` | {100 * f(r, '# Samples') / tot_s:.2f} | {int(f(r, 'Instructions Executed'))} | "
                         f"{f(r, 'Avg. Threads Executed'):.1f} | {', '.join(f'{a} {int(b)}' for a, b in st)} |")
        lines.append("")
    except StopIteration:
        pass
    # the same samples grouped by CUDA source line (needs -lineinfo + --import-source on)
    try:
        cs = list(csv.reader(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]).splitlines()))
        cur, agg = "?", []
        for r in cs:
            if len(r) == 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif len(r) > 10 and r[0].isdigit() and r[6].isdigit() and r[7].isdigit():
                agg.append((cur, int(r[0]), r[1].strip()[:100].replace("|", "/"), int(r[6]), int(r[7])))
        ts = sum(a[3] for a in agg) or 1
        tn = sum(a[4] for a in agg) or 1
        lines += ["## hottest CUDA source lines (all launches of the report)", "", "| file:line | samples % | warp instr % | source |", "|---|---|---|---|"]
        for a in sorted(agg, key=lambda a: -a[3])[:30]:
            lines.append(f"| {a[0]}:{a[1]} | {100 * a[3] / ts:.1f} | {100 * a[4] / tn:.1f} | `{a[2]}` |")
        lines.append("")
    except Exception as e:   # older captures without source import
        lines += [f"(no CUDA source correlation: {e})", ""]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
