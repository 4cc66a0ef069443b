#!/usr/bin/env python3
"""Dynamic SASS opcode mix of a kernel from an ncu report's source page:
   ncu -i rep.ncu-rep --page source --csv > src.csv ; python lab/sass_mix.py src.csv [perms]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, isrc, iex = h.index("Address"), h.index("Source"), h.index("Thread Instructions Executed")
mix = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= iex: continue
    toks = r[isrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op] += float(r[iex] or 0)
tot = sum(mix.values())
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
fma = sum(v for k, v in mix.items() if k.startswith(FMA))
wide = sum(v for k, v in mix.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
print(f"total thread-instr {tot:.4g}" + (f" = {tot/units:.0f} per unit" if units else ""))
print(f"fma-pipe instr {fma/tot:.1%} (of which IMAD.WIDE/HI {wide/tot:.1%}); fmaheavy slots/instr-total = {(fma+wide)/tot:.2f}")
for k, v in mix.most_common(28):
    print(f"  {k:28s} {v/tot:6.1%}" + (f"  {v/units:8.0f}/unit" if units else ""))
