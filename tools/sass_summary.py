#!/usr/bin/env python3
"""SASS mnemonic counts per kernel of libnsb200.so (cuobjdump -sass): the instructions that prove tcgen05 / TMEM / TMA / mma.sync use.
   python tools/sass_summary.py > profiles/r02_sass_mnemonics.txt        (runs without a GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nemotron-speech.cpp_b200", "libnsb200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
WANT = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "LDTM", "STTM", "HMMA", "IMMA", "LDSM", "LDGSTS", "SYNCS", "REDG", "ATOMG", "UCGABAR"]
tot, per = collections.Counter(), collections.OrderedDict()
for f in funcs:
    mangled = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+_kernel(?:<[^(]*>)?)", dem)
    name = m.group(1) if m else dem[:60]
    c = collections.Counter({w: len(re.findall(r"\b" + w + r"\b", f)) for w in WANT})
    c = collections.Counter({k: v for k, v in c.items() if v})
    if c:
        per.setdefault(name, collections.Counter()).update(c)
        tot.update(c)
print("# SASS mnemonic counts in libnsb200.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma = UTCHMMA (kind::f16 / kind::tf32), tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, TMA load = UTMALDG, TMA L2 prefetch = UTMAPF,")
print("# mma.sync f16 / bf16 = HMMA, mma.sync s8 = IMMA (strict Q8_0), ldmatrix = LDSM, cp.async = LDGSTS, mbarrier ops = SYNCS, cluster barrier = UCGABAR.")
print("# UTMASTG (TMA store) and STTM (tcgen05.st) do not occur: every epilogue stores from registers (fp32 tiles through a shared-memory transpose).")
print()
print("total: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())))
print()
for name, c in sorted(per.items(), key=lambda kv: (-kv[1].get("UTCHMMA", 0), -sum(kv[1].values()))):
    print(f"{name[:64]:64s} " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())))
