#!/usr/bin/env python
"""Condense ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python scripts/ncu_summary.py launches gpurun_out/X_launches.csv profiles/X_launches.csv
    python scripts/ncu_summary.py full gpurun_out/X_kernel.ncu-rep profiles/X_kernel_ncu.txt
"""
import csv
import re
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed")


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name[:110]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    h = rows[0]
    ik, im, iv, iu, ig, ib = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    out = []
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        us = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        out.append((int(r[0]), short(r[ik]), r[ig], r[ib], us))
    tot = sum(o[4] for o in out)
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("id,kernel,grid,block,duration_us,share_of_listed\n")
        for o in out:
            f.write(f"{o[0]},\"{o[1]}\",\"{o[2]}\",\"{o[3]}\",{o[4]:.2f},{o[4] / tot:.4f}\n")
    print(f"{len(out)} launches, {tot / 1e3:.3f} ms listed -> {dst}")


def full(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in txt.splitlines() if l.startswith('"')))
    h, u = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {src}\n")
        for v in rows[2:]:
            f.write(f"kernel: {short(v[h.index('Kernel Name')])}\n")
            rd = wr = None
            for i, n in enumerate(h):
                if n in KEEP:
                    f.write(f"  {n} [{u[i]}] = {v[i]}\n")
                    if n == "dram__bytes_read.sum":
                        rd = (float(v[i].replace(",", "")), u[i])
                    if n == "dram__bytes_write.sum":
                        wr = (float(v[i].replace(",", "")), u[i])
            if rd and wr:
                sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                f.write(f"  traffic_bytes (dram read + write) = {rd[0] * sc[rd[1]] + wr[0] * sc[wr[1]]:.0f}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
