#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libcumind_b200.so -> profiles/<tag>_sass_opcodes.csv.

    tools/sass_opcodes.py <tag>

Static evidence for what each kernel is made of: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG
(TMA tensor load), UBLKCP (TMA bulk copy), SYNCS (mbarrier), FFMA2 (fma.rn.f32x2), DFMA, USETMAXREG (setmaxnreg), MATCH,
loads / stores with .SYS scope (peer-memory flags), ACQBULK / PREEXIT (programmatic dependent launch)."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "cumind_b200", "_lib", "libcumind_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "HFMA2", "DFMA", "DADD", "DMUL",
         "USETMAXREG", "MATCH", "ACQBULK", "PREEXIT", "LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "RED", "BAR", "SHFL", "LDGSTS", "MUFU"]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True, check=True)
rows = []
for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
    if os.path.basename(cubin).count("-") > 2:   # the merged module repeats every kernel
        continue
    txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    name, counts, total, sys_scope = None, None, 0, 0
    for line in txt.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                rows.append((os.path.basename(cubin).split(".")[0], name, total, sys_scope, counts))
            name, counts, total, sys_scope = m.group(1), collections.Counter(), 0, 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            total += 1
            op = m.group(1)
            if op in WATCH:
                counts[op] += 1
            if ".SYS" in m.group(2) and op in ("LD", "ST", "LDG", "STG", "ATOM", "RED", "ATOMG"):
                sys_scope += 1
    if name:
        rows.append((os.path.basename(cubin).split(".")[0], name, total, sys_scope, counts))
demangle = subprocess.run(["cu++filt"] + [r[1] for r in rows], capture_output=True, text=True).stdout.splitlines()
out = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.csv")
with open(out, "w") as f:
    f.write("module,kernel,sass_instructions,sys_scope_mem_ops," + ",".join(WATCH) + "\n")
    for (mod, name, total, sys_scope, counts), dm in zip(rows, demangle):
        dm = re.sub(r"^void ", "", dm).replace('"', "'")
        f.write(f"{mod},\"{dm}\",{total},{sys_scope}," + ",".join(str(counts.get(w, 0)) for w in WATCH) + "\n")
print(out, len(rows), "kernels")
