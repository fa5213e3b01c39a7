#!/usr/bin/env python
"""Where do a kernel's instructions go?  Reads the source page of an `ncu --set full --import-source on` report and sums the
executed warp instructions and warp-stall samples per FUNCTION of the CUDA source (inlined device functions keep their own
line numbers, so `__forceinline__` helpers show up as themselves), optionally divided by a unit count.

    git show eeaf664:cs348b-pbrt_b200/csrc/pv_gather.cu > /tmp/pv_gather.cu        # the source as profiled
    tools/kernel_regions.py gpurun_out/r01_v5_gather.ncu-rep --source pv_gather.cu=/tmp/pv_gather.cu --units 54822484 --unit lookup
"""
import argparse
import collections
import csv
import re
import subprocess

FUNC = re.compile(r"^(?:static\s+)?(?:__device__|__global__)[^;]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--units", type=float, default=0.0, help="units of work of the launch (lookups, steps, paths ...)")
    ap.add_argument("--unit", default="unit")
    ap.add_argument("--source", action="append", default=[], help="source file AS PROFILED (e.g. from `git show <commit>:<path>`), "
                    "optionally name=path; function starts are read from it (the report lists only lines that own instructions)")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    cur = None
    starts = collections.defaultdict(list)      # file -> [(line, function)]
    for spec in args.source:
        name, _, path = spec.rpartition("=")
        name = name or path.split("/")[-1]
        for ln, text in enumerate(open(path, errors="ignore"), 1):
            if text.startswith((" ", "\t")):
                continue
            m = FUNC.match(text)
            if m and m.group(1) not in ("__launch_bounds__", "__maxnreg__"):
                starts[name].append((ln, m.group(1)))
            elif "__global__" in text:
                k = re.search(r"\b([A-Za-z_][A-Za-z_0-9]*)\s*\([A-Za-z_]+ ", text.split(")", 1)[-1])
                if k:
                    starts[name].append((ln, k.group(1)))
    rows = []
    kernel = "?"
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "Kernel Name":
            kernel = r[1]
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] not in ("", "Line No"):
            try:
                ln = int(r[0])
            except ValueError:
                continue
            if len(r) > 10 and r[2] == "-":
                try:
                    rows.append((cur, ln, int(r[4]), int(r[7])))
                except ValueError:
                    pass
    inst = collections.Counter(); samp = collections.Counter()
    for f, ln, s, i in rows:
        name = "(file scope)"
        for a, n in starts[f]:
            if a <= ln:
                name = n
        inst[(f, name)] += i; samp[(f, name)] += s
    ti = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
    print("# instruction breakdown of `%s` (`%s`)\n" % (kernel, args.report.split("/")[-1]))
    hdr = "| file | function | warp instructions | share | stall samples |"
    if args.units:
        hdr = "| file | function | warp instructions / %s | share | stall samples |" % args.unit
    print(hdr); print("|---|---|---|---|---|")
    for k, v in inst.most_common():
        if v / ti < 0.002:
            continue
        print("| %s | `%s` | %s | %.1f %% | %.1f %% |" % (k[0], k[1], ("%.0f" % (v / args.units)) if args.units else "%d" % v, 100 * v / ti, 100 * samp[k] / ts))
    print("| | total | %s | | |" % (("%.0f" % (ti / args.units)) if args.units else "%d" % ti))


if __name__ == "__main__":
    main()
