#!/usr/bin/env python
"""Turn gpurun_out/<tag>_{gather,march}.ncu-rep + <tag>_launches.csv into the tracked evidence under profiles/:
raw metric pages (csv), a per-kernel launch table, a short markdown summary, the DRAM traffic json bench.py reads,
and the SASS listing of the gather kernel.  usage: tools/profile_summary.py <tag>"""
import csv, collections, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]
md = ["# ncu evidence `%s`\n" % tag, "Commands: `tools/run_full.sh` (plain run first, then `ncu --metrics gpu__time_duration.sum --clock-control none` for the launch list,",
      "then one `ncu --set full --clock-control none --import-source on` capture per kernel).  Times under ncu are cold-cache / serialised.\n"]
for kern in ("gather", "march"):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, kern))
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, "%s_%s_raw.csv" % (tag, kern)), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    md.append("## %s\n" % d.get("Kernel Name", (kern,))[0])
    md.append("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            md.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
    md.append("")
    if kern == "gather":
        def b(x):
            v, u = d[x]; v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        tr = {"dram_bytes_per_launch": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"), "dram_read": b("dram__bytes_read.sum"),
              "dram_write": b("dram__bytes_write.sum"), "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"][0]),
              "l1_hit_pct": float(d["l1tex__t_sector_hit_rate.pct"][0]), "source": "profiles/%s_gather_raw.csv" % tag}
        json.dump(tr, open(os.path.join(P, "r01_gather_traffic.json"), "w"), indent=1)
    # hottest source lines
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur = None; agg = []
    for r in csv.reader(src.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if len(r) > 10 and r[0] not in ("", "Line No") and r[2] == "-":
            try:
                agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[4]), int(r[7])))
            except ValueError:
                pass
    ts = sum(a[3] for a in agg) or 1; ti = sum(a[4] for a in agg) or 1
    md.append("Hottest source lines (warp-stall samples / executed instructions):\n")
    md.append("| file:line | samples % | inst % | source |\n|---|---|---|---|")
    for a in sorted(agg, key=lambda a: -a[3])[:25]:
        md.append("| %s:%d | %.2f | %.2f | `%s` |" % (a[0], a[1], 100 * a[3] / ts, 100 * a[4] / ti, a[2].replace("|", "\\|")))
    md.append("")
# the shooter: one row per captured launch (volume-only waves first, then the all-maps waves of pv_shoot_maps)
rep = os.path.join(G, "%s_shoot.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(P, "%s_shoot_raw.csv" % tag), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    cols = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
    md.append("## shoot_kernel launches (`%s_shoot_raw.csv`; `<false>` = volume photons only, `<true>` = all photon maps)\n" % tag)
    md.append("| kernel | " + " | ".join(c.split(".")[0].replace("smsp__", "").replace("sm__", "") for c in cols) + " |\n|---|" + "---|" * len(cols))
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        md.append("| %s | " % r[ik][:40] + " | ".join(r[hdr.index(c)] if c in hdr else "-" for c in cols) + " |")
    md.append("")
lc = os.path.join(G, "%s_launches.csv" % tag)
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if l.startswith('"'))]
    h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value")
    t = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ik].split("(")[0]; t.setdefault(k, []).append(float(r[iv].replace(",", "")))
    tot = sum(sum(v) for v in t.values())
    md.append("## launch list (`%s_launches.csv`, whole bench process incl. shooting sample, map build, warm-up)\n" % tag)
    md.append("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
    for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        md.append("| %s | %d | %.3f | %.1f%% | %.3f |" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot, sum(v) / len(v) / 1e6))
    open(os.path.join(P, "%s_launches.csv" % tag), "w").write(open(lc).read())
open(os.path.join(P, "%s_summary.md" % tag), "w").write("\n".join(md) + "\n")
# SASS of the gather kernel (proves UBLKCP / SYNCS / LDG.128 etc.)
so = os.path.join(ROOT, "cs348b-pbrt_b200", "csrc", "pv_gather.o")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
out = []; on = False
for l in sass.splitlines():
    if "Function :" in l:
        on = "gather_kernel" in l
    if on and ("/*" in l or "Function" in l) and not l.strip().startswith("/* 0x"):
        out.append(l.rstrip())
open(os.path.join(P, "%s_gather_kernel.sass" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(md[:60]))
