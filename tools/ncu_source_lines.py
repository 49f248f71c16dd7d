"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: python tools/ncu_source_lines.py dump.csv [top_n] [kernel-substring]
Prints, per kernel, the lines with the most stall samples and executed warp instructions."""
import csv, sys
from collections import defaultdict

def main(path, top=40, only=None):
    rows = list(csv.reader(open(path)))
    fname, func, hdr = None, 'kernel', None
    per = defaultdict(list)
    for r in rows:
        if not r:
            continue
        if r[0] in ('File Name', 'File Path'):
            fname = r[1].split('/')[-1]
            continue
        if r[0] in ('Function Name', 'Kernel Name'):
            func = r[1].split('(')[0]
            continue
        if r[0] == 'Line No':
            hdr = r
            iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed')
            stall = {h: i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
            continue
        if hdr is None or not r[0].isdigit() or len(r) <= iI:
            continue
        try:
            smp = int(r[iS]); ins = int(r[iI])
        except ValueError:
            continue
        st = {k[6:]: int(r[i]) for k, i in stall.items() if i < len(r) and r[i].isdigit() and int(r[i]) > 0}
        per[func].append((fname, int(r[0]), r[1].strip()[:100], smp, ins, st))
    for func, out in per.items():
        if only and only not in func:
            continue
        tot_s = sum(o[3] for o in out) or 1; tot_i = sum(o[4] for o in out) or 1
        print(f"== {func}: samples {tot_s}, warp instructions {tot_i}")
        out.sort(key=lambda o: -o[3])
        for f, ln, src, smp, ins, st in out[:top]:
            top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"{f}:{ln:<4d} {100*smp/tot_s:5.1f}% smp {100*ins/tot_i:5.1f}% ins  {top3}  | {src}")

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else None)
