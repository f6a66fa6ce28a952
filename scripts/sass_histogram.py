"""cuobjdump -sass of libvae_b200.so -> per-kernel histogram of the opcodes that prove what each kernel runs on:
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
FFMA / FFMA2 (FP32 FMA, packed), LDS / STS, LDG / STG, BAR, UCGABAR (cluster barrier).
Usage: python scripts/sass_histogram.py [lib.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "vae_training_b200", "libvae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR", "FFMA2", "FFMA",
         "FMUL", "FADD", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "BAR", "SHFL", "LOP3", "IADD3", "HMMA", "IMAD"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op, mods = m.group(1), m.group(2)
        total[kern] += 1
        if op in WATCH:
            hist[kern][op] += 1
            if op in ("UTMALDG", "UTCHMMA", "LDTM", "LDG", "STG", "LDS", "STS") and mods:
                hist[kern][op + mods] += 1
print(f"# SASS opcode histogram of {os.path.basename(lib)} (cuobjdump -sass, sm_100a); static instruction counts per kernel")
arch = set(re.findall(r"arch = (sm_\w+)", out))
print(f"# cubin architectures: {sorted(arch)}")
for k, h in hist.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    if len(name) > 150:
        name = name[:150] + "..."
    print(f"\n{name}\n  instructions: {total[k]}")
    for op in sorted(h, key=lambda o: (WATCH.index(o.split('.')[0]), o)):
        print(f"    {op:28s} {h[op]}")
