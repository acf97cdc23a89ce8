#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): samples and executed
warp instructions per CUDA source line of the first (or --launch k) kernel in the report.

    python profiles/ncu_lines.py gpurun_out/prof_x.ncu-rep [--top 40] [--launch 0]
"""
import csv, io, subprocess, sys

def main():
    args = sys.argv[1:]
    top, launch = 40, 0
    if "--top" in args:
        i = args.index("--top"); top = int(args[i + 1]); del args[i:i + 2]
    if "--launch" in args:
        i = args.index("--launch"); launch = int(args[i + 1]); del args[i:i + 2]
    out = subprocess.run(["ncu", "-i", args[0], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    # the report is a sequence of blocks: "File Path", "Function Name", header row, rows
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur = {"file": row[1], "rows": [], "header": None, "func": None}
            blocks.append(cur)
        elif row[0] == "Function Name" and cur is not None:
            cur["func"] = row[1]
        elif row[0] == "Line No" and cur is not None:
            cur["header"] = row
        elif cur is not None and cur["header"] is not None:
            cur["rows"].append(row)
    # group blocks by function occurrence: a new launch starts when the same file repeats
    launches, seen = [], set()
    for b in blocks:
        key = b["file"]
        if not launches or key in seen:
            launches.append([]); seen = set()
        seen.add(key)
        launches[-1].append(b)
    sel = launches[min(launch, len(launches) - 1)]
    lines = []
    tot_s = tot_i = 0
    for b in sel:
        h = b["header"]
        i_s = h.index("# Samples"); i_i = h.index("Instructions Executed")
        for r in b["rows"]:
            if r[0] == "":
                continue  # SASS row
            try:
                smp, ins = int(r[i_s]), int(r[i_i])
            except ValueError:
                continue
            tot_s += smp; tot_i += ins
            lines.append((smp, ins, b["file"].split("/")[-1], r[0], r[1].strip()[:110]))
    print("function: %s" % (sel[0]["func"] or "")[:100])
    print("total samples %d, warp instructions %d" % (tot_s, tot_i))
    print("%7s %7s  %s" % ("smp%", "inst%", "line"))
    for smp, ins, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
        print("%6.2f%% %6.2f%%  %s:%s  %s" % (100.0 * smp / max(tot_s, 1), 100.0 * ins / max(tot_i, 1), f, ln, src))

if __name__ == "__main__":
    main()
