#!/usr/bin/env python
"""Attribute an ncu source-page capture to CUDA source lines.

    tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top_n] [--by-wavefronts]

`ncu --page source --csv` lists per-SASS-instruction counters; `nvdisasm -g` on the cubin extracted
from the library gives the source line of every SASS instruction.  The two listings have the same
instruction order, so they are joined by index and summed per (file, line)."""

import csv
import glob
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = []
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
        cur, active, loc = None, False, ("?", 0)
        for line in txt.splitlines():
            if line.startswith(".text."):
                active = kernel in line
                if active and out:
                    return out
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), loc, m.group(2).strip()))
        if out:
            return out
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    cw, cx = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")

    def num(v):
        try:
            return int(float(v or 0))
        except ValueError:
            return 0

    prof = [(r[cn].strip(), num(r[ci]), num(r[cs]), num(r[cw]), num(r[cx])) for r in rows[hdr_i + 1:] if len(r) > cx]
    sass = sass_lines(lib, kernel)
    if len(sass) != len(prof):
        print(f"warning: {len(sass)} SASS instructions in the cubin vs {len(prof)} in the report", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0, 0, 0])
    for (off, loc, text), (src, inst, samp, wav, exc) in zip(sass, prof):
        a = agg[loc]
        a[0] += inst
        a[1] += samp
        a[2] += 1
        a[3] += wav
        a[4] += exc
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[1] for a in agg.values()) or 1
    tot_w = sum(a[3] for a in agg.values()) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}, shared-memory wavefronts {tot_w} "
          f"(excessive {sum(a[4] for a in agg.values())})")
    key = 3 if "--by-wavefronts" in sys.argv else 1
    print(f"{'file:line':40s} {'inst%':>7s} {'samples%':>9s} {'#sass':>6s} {'smem wf%':>9s} {'excess%':>8s}")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{loc[0] + ':' + str(loc[1]):40s} {100 * a[0] / tot_i:7.2f} {100 * a[1] / tot_s:9.2f} {a[2]:6d} "
              f"{100 * a[3] / tot_w:9.2f} {100 * a[4] / tot_w:8.2f}")


if __name__ == "__main__":
    main()
