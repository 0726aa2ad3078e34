#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an ncu report: joins the SASS page of
the report (executed instructions per SASS instruction) with the line table of the matching
cubin (nvdisasm -g), so that the time of a kernel can be attributed to phases of its source.

  python tools/ncu_lines.py gpurun_out/x.ncu-rep psxavenc_b200/libpsxav_b200.so 'bs_pack_kernel<(bool)0, (bool)1, (bool)0, (int)320, (int)4>' [lo:hi:label ...]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1]
    head = rows[1]
    i_src, i_exec, i_thr = head.index("Source"), head.index("Instructions Executed"), head.index("Thread Instructions Executed")
    i_samp = head.index("# Samples")
    return name, [(r[i_src].strip(), int(r[i_exec] or 0), int(r[i_thr] or 0), int(r[i_samp] or 0)) for r in rows[2:] if len(r) > i_thr]


def line_table(so, kernel_pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    for cubin in sorted(os.listdir(tmp)):
        text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        # split into functions
        parts = re.split(r"\n\s*\.text\.", text)
        for part in parts[1:]:
            header = part.split("\n", 1)[0]
            demangled = subprocess.run(["c++filt", header.split(":")[0].strip()], capture_output=True, text=True).stdout.strip()
            if kernel_pattern in demangled.replace("psxb200::", ""):
                lines, cur = [], 0
                for ln in part.split("\n"):
                    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
                    if m:
                        cur = (m.group(1), int(m.group(2)))
                        continue
                    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
                        lines.append(cur)
                return demangled, lines
    raise SystemExit("kernel %r not found in %s" % (kernel_pattern, so))


def main():
    rep, so, pattern = sys.argv[1:4]
    ranges = []
    for spec in sys.argv[4:]:
        lo, hi, label = spec.split(":", 2)
        ranges.append((int(lo), int(hi), label))
    name, rows = sass_rows(rep)
    demangled, lines = line_table(so, pattern)
    if len(lines) != len(rows):
        print("# warning: %d SASS instructions in the report, %d in the cubin (different build?)" % (len(rows), len(lines)))
    per_line = {}
    for (src, ex, thr, samp), where in zip(rows, lines):
        key = where if where else ("?", 0)
        a = per_line.setdefault(key, [0, 0, 0])
        a[0] += ex
        a[1] += thr
        a[2] += samp
    total = sum(v[0] for v in per_line.values()) or 1
    tsamp = sum(v[2] for v in per_line.values()) or 1
    print("# %s" % name)
    print("# total warp instructions executed: %d" % total)
    if ranges:
        for lo, hi, label in ranges:
            ex = sum(v[0] for (f, l), v in per_line.items() if lo <= l <= hi and f.endswith(".cu"))
            sm = sum(v[2] for (f, l), v in per_line.items() if lo <= l <= hi and f.endswith(".cu"))
            print("%-40s lines %4d-%4d  %6.2f %% of instructions  %6.2f %% of samples" % (label, lo, hi, 100.0 * ex / total, 100.0 * sm / tsamp))
    else:
        for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:60]:
            print("%-24s %5d  %6.2f %% instr  %6.2f %% samples" % (f, l, 100.0 * v[0] / total, 100.0 * v[2] / tsamp))


if __name__ == "__main__":
    main()
