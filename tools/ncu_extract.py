#!/usr/bin/env python
"""Compact per-launch table from an `ncu --set full` report (run where ncu is installed; no GPU needed).
usage: python tools/ncu_extract.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "warp_inst")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size")
    print(f"# {rep} (ncu --set full --clock-control none; per-launch, cold cache, serialised)")
    print("kernel | grid | " + " | ".join(f"{n} [{units[i]}]" for i, n in idx))
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace("void ", "").replace("cfb::<unnamed>::", "")
        print(f"{name} | {r[ig]} | " + " | ".join(r[i] for i, _ in idx))


if __name__ == "__main__":
    main()
