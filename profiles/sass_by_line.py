#!/usr/bin/env python
"""Attribute executed SASS instructions of one profiled kernel to CUDA source lines.

    python profiles/sass_by_line.py REPORT.ncu-rep LIB.so KERNEL_SUBSTR [TOP]

Joins the per-instruction "Instructions Executed" / stall samples of the ncu source page (SASS
view) with `nvdisasm -g` line info of the cubin inside LIB.so (offset within the function).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main(rep, lib, kernel, top=60):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    line_of = {}
    fn = None
    for cub in os.listdir(tmp):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kernel in cur_fn:
                fn = fn or cur_fn
                if cur_fn == fn:
                    line_of[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, base = None, None
    by_line = collections.Counter()
    smp_line = collections.Counter()
    ops_line = collections.defaultdict(collections.Counter)
    total = 0
    seen_kernel = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            seen_kernel += 1
            if seen_kernel > 1:
                break
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        addr = int(d["Address"], 16)
        base = addr if base is None else base
        off = addr - base
        n = int(d["Instructions Executed"] or 0)
        smp = int(d["# Samples"] or 0)
        line, sass = line_of.get(off, (None, d["Source"].strip()))
        op = re.sub(r"^@!?U?P\d+\s+", "", d["Source"].strip()).split()[0].split(".")[0]
        by_line[line] += n
        smp_line[line] += smp
        ops_line[line][op] += n
        total += n
    print(f"# {fn}: {total} executed warp instructions (first profiled launch of {rep})")
    print("#   instr      %   samples  file:line   top opcodes")
    for line, n in by_line.most_common(top):
        ops = ", ".join(f"{o} {c * 100 // max(n, 1)}%" for o, c in ops_line[line].most_common(5))
        tag = f"{line[0]}:{line[1]}" if line else "?"
        print(f"{n:>11d} {100.0 * n / total:5.2f}% {smp_line[line]:>8d}  {tag:28s} {ops}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 60)
