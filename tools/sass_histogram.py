#!/usr/bin/env python
"""SASS opcode histogram of the product library's kernels (cuobjdump -sass), so that "Blackwell-native, hand-written"
need not be re-derived: packed fp32 (FFMA2 / FADD2 / FMUL2), MUFU, TMA bulk copies (UBLKCP), mbarrier transactions
(SYNCS), cluster barriers (UCGABAR), cp.async (LDGSTS), tensor-core opcodes (none expected: not a contraction).
usage: tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nbodysim_b200", "libnbody_gpu.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
WATCH = ("FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "DFMA", "UBLKCP", "SYNCS", "UCGABAR", "LDGSTS", "ATOM", "ATOMS", "RED", "REDUX",
         "MATCH", "LDS", "STS", "LDG", "STG", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "BAR", "CCTL", "MEMBAR")
print(f"library: {os.path.relpath(lib, ROOT)}   cubin architectures: {', '.join(arch)}")
print("per kernel: instruction count, then the watched opcode families (base mnemonic before the first dot)\n")
total = collections.Counter()
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter(m.group(1).split(".")[0].split("_")[0] for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", f))
    total.update(ops)
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    dem = dem.replace("(int)", "").replace("(bool)", "").replace("nb::", "")
    dem = re.sub(r">\(.*$", ">", dem) if ">(" in dem else re.sub(r"\(.*$", "", dem)
    dem = dem[:120]
    shown = "  ".join(f"{k}:{ops[k]}" for k in WATCH if ops.get(k))
    print(f"{dem}\n    {sum(ops.values()):6d} instr   {shown}")
print("\nwhole library: " + "  ".join(f"{k}:{total[k]}" for k in WATCH if total.get(k)))
print("tensor-core opcodes (HMMA / UTC*MMA / LDTM): " + str(sum(total[k] for k in ("HMMA", "UTCHMMA", "UTCQMMA", "LDTM"))) + " -- the path is not a dense contraction")
