#!/usr/bin/env python
"""SASS instruction count per source-line range of agar_step_kernel<128,6> (code size, not execution):
    cuobjdump -xelf all agarai_b200/libagar_b200.so; nvdisasm --print-line-info *.cubin > dis.txt
    python profiles/codesize.py dis.txt [bucket]"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("\n")
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur = None; infn = False; counts = collections.Counter(); inl = collections.Counter()
for l in txt:
    if l.startswith(".text."):
        infn = "agar_step_kernelILi128ELi7ENS_5StCfgILi1ELi25" in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        counts[cur] += 1
tot = sum(counts.values())
print("total instructions", tot, f"{tot*16/1024:.1f} KB")
b = collections.Counter()
for (f, ln), c in counts.items():
    b[(f, ln // bucket * bucket)] += c
for (f, ln), c in sorted(b.items()):
    if c >= 40: print(f"{f}:{ln:5d}  {c:6d}  {100*c/tot:5.1f}%")
