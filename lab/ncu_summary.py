#!/usr/bin/env python3
"""Markdown summary of an `ncu --set full` report: python lab/ncu_summary.py rep.ncu-rep [units] [launch index, default 0] > profiles/x.md
`units` = work units the launch processed (permutations, field elements ...) for the per-unit instruction count."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h, unit_row, d = rows[0], dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2 + which]))
def g(k):
    v = d.get(k, "")
    try: return float(v.replace(",", ""))
    except Exception: return None
keys = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipes (heavy+lite) % of peak"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy (IMAD) pipe cycles active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("sm__icc_request_hit_rate.pct", "instruction cache hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
]
print(f"kernel: `{d.get('Kernel Name','?')}`  (report {rep.split('/')[-1]}, ncu --set full --clock-control none)\n")
print("| metric | value |\n|---|---|")
for k, name in keys:
    if k in d and d[k] != "":
        print(f"| {name} (`{k}`) | {d[k]} {unit_row.get(k,'')} |")
inst = g("smsp__inst_executed.sum")
if inst and units:
    print(f"| thread instructions per unit ({units:.0f} units) | {inst*32/units:.0f} |")
print("\nwarp stall reasons (cycles per issued instruction):\n")
st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), g(k)) for k in h
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
print(", ".join(f"{k} {v:.2f}" for k, v in sorted(st, key=lambda kv: -(kv[1] or 0)) if v and v >= 0.02))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    lo = hi[min(which, len(hi) - 1)]                       # one "Address" header per captured launch
    end = hi[which + 1] if which + 1 < len(hi) else len(rows)
    hh = rows[lo]
    isrc, iex = hh.index("Source"), hh.index("Thread Instructions Executed")
    mix = collections.Counter()
    for r in rows[lo + 1:end]:
        if len(r) <= iex or not r[isrc].split(): continue
        t = r[isrc].split()
        mix[t[1] if t[0].startswith("@") else t[0]] += float(r[iex] or 0)
    tot = sum(mix.values())
    print("\ndynamic SASS mix (thread instructions, top 14):\n")
    print(", ".join(f"{k} {v/tot:.1%}" for k, v in mix.most_common(14)))
    wide = sum(v for k, v in mix.items() if k.startswith("IMAD.WIDE") or k.startswith("IMAD.HI"))
    fma = sum(v for k, v in mix.items() if k.startswith("IMAD"))
    print(f"\nIMAD-family share {fma/tot:.1%}, of which IMAD.WIDE/HI (2 fmaheavy slots each) {wide/tot:.1%}")
