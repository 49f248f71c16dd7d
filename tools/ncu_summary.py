#!/usr/bin/env python
"""Summarise ncu outputs into markdown for profiles/.

  python tools/ncu_summary.py launches <launches.csv> [--step-marker march_ray_count]   -> per-kernel time of one step
  python tools/ncu_summary.py full <file.ncu-rep>                                      -> key metrics per captured kernel
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def launches(path, marker="march_ray_count"):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    names = [r[ki] for r in data]
    vals = [float(r[vi].replace(',', '')) for r in data]
    idx = [i for i, n in enumerate(names) if marker in n]
    a, b = idx[-3], idx[-2]
    agg = collections.OrderedDict()
    for n, v in zip(names[a:b], vals[a:b]):
        k = n.split('(')[0][-80:]
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += v
    tot = sum(e[1] for e in agg.values())
    print(f"one training step = launches [{a},{b}) of {len(names)}: {b - a} kernels, {tot / 1e3:.1f} us of GPU time (ncu, serialised, cold cache)\n")
    print("| us | launches | share | kernel |\n|---:|---:|---:|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {t / 1e3:.1f} | {n} | {100 * t / tot:.1f}% | `{k}` |")


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(lbl for _, lbl in KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split('(')[0][-48:]
        cells = []
        for k, _ in KEYS:
            if k in idx:
                cells.append(f"{r[idx[k]]} {units[idx[k]]}".strip())
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], *(sys.argv[4:5] if len(sys.argv) > 4 else []))
    else:
        full(sys.argv[2])
