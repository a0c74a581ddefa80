#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into small tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof_r01.ncu-rep  profiles/r01_kernels.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[i_name].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[i_val].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
        f.write("source: `%s`, %d launches, %.3f ms of kernel time\n\n" % (src, len(rows) - 1, total / 1e6))
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (name[:90], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / total))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        seen.setdefault(name, []).append(r)
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: per-kernel raw metrics (last captured launch of each kernel)\n\n")
        f.write("source: `%s`\n\n" % src)
        for name, rs in seen.items():
            r = rs[-1]
            f.write("## `%s` (%d launches captured)\n\n| metric | value | unit |\n|---|---|---|\n" % (name, len(rs)))
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
