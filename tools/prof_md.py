#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep) into the tracked evidence under profiles/: one raw metric page (csv) per report, a
markdown summary with the metrics the design argues from, the hottest source lines, a launch table, and SASS listings.

    tools/prof_md.py <tag> [--rep label=path.ncu-rep ...] [--launches launches.csv] [--sass object.o:kernel_substring ...]
                           [--note "text"]
"""
import argparse
import collections
import csv
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--launches")
    ap.add_argument("--sass", action="append", default=[])
    ap.add_argument("--note", action="append", default=[])
    a = ap.parse_args()
    md = ["# ncu evidence `%s`\n" % a.tag,
          "Every capture below was taken AFTER the same command had exited 0 without ncu (`--set full --clock-control none --import-source on`, one",
          "launch per kernel; launch lists with `--metrics gpu__time_duration.sum --clock-control none`).  Times under ncu are cold-cache and",
          "serialised: they give each kernel's SHARE, the bench's CUDA-event times give the absolute numbers.\n"]
    md += [n + "\n" for n in a.note]
    for spec in a.rep:
        label, rep = spec.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(P, "%s_%s_raw.csv" % (a.tag, label)), "w").write(raw)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            if len(vals) != len(hdr):
                continue
            d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
            md.append("## %s  (`%s_%s_raw.csv`)\n" % (d.get("Kernel Name", (label,))[0], a.tag, label))
            md.append("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in d:
                    md.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
            md.append("")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
        cur = None
        agg = []
        for r in csv.reader(src.splitlines()):
            if len(r) == 2 and r[0] == "File Path":
                cur = r[1].split("/")[-1]
                continue
            if len(r) > 10 and r[0] not in ("", "Line No") and r[2] == "-":
                try:
                    agg.append((cur, int(r[0]), r[1].strip()[:110], int(r[4]), int(r[7])))
                except ValueError:
                    pass
        if agg:
            ts = sum(x[3] for x in agg) or 1
            ti = sum(x[4] for x in agg) or 1
            md.append("Hottest source lines (warp-stall samples / executed warp instructions):\n")
            md.append("| file:line | samples % | inst % | source |\n|---|---|---|---|")
            for x in sorted(agg, key=lambda x: -x[3])[:18]:
                md.append("| %s:%d | %.2f | %.2f | `%s` |" % (x[0], x[1], 100 * x[3] / ts, 100 * x[4] / ti, x[2].replace("|", "\\|")))
            md.append("")
    if a.launches and os.path.exists(a.launches):
        rows = [r for r in csv.reader(l for l in open(a.launches) if l.startswith('"'))]
        h = rows[0]
        ik = h.index("Kernel Name"); iv = h.index("Metric Value")
        t = collections.OrderedDict()
        for r in rows[1:]:
            t.setdefault(r[ik].split("(")[0], []).append(float(r[iv].replace(",", "")))
        tot = sum(sum(v) for v in t.values())
        md.append("## launch list (`%s_launches.csv`)\n" % a.tag)
        md.append("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
        for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
            md.append("| %s | %d | %.3f | %.1f%% | %.4f |" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot, sum(v) / len(v) / 1e6))
        md.append("")
        open(os.path.join(P, "%s_launches.csv" % a.tag), "w").write(open(a.launches).read())
    for spec in a.sass:
        obj, kern = spec.split(":", 1)
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        out = []
        on = False
        for l in sass.splitlines():
            if "Function :" in l:
                on = kern in l
                if on and out:
                    break                      # first instantiation only
            if on and ("/*" in l or "Function" in l) and not l.strip().startswith("/* 0x"):
                out.append(l.rstrip())
        fn = "%s_%s.sass" % (a.tag, kern)
        open(os.path.join(P, fn), "w").write("\n".join(out) + "\n")
        ops = collections.Counter()
        for l in out:
            parts = l.split("*/")
            if len(parts) > 1 and parts[1].strip():
                tok = parts[1].strip().split()
                op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
                ops[op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("LDG", "STG", "LDS", "UBLKCP", "SYNCS", "ATOM", "RED")) else "")] += 1
        md.append("## SASS `%s` (%d instructions; first instantiation in `%s`)\n" % (fn, len(out) - 1, os.path.basename(obj)))
        md.append(", ".join("%s x%d" % kv for kv in ops.most_common(24)) + "\n")
    open(os.path.join(P, "%s_summary.md" % a.tag), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    main()
