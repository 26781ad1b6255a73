#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small text table."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wave%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
        ("smsp__inst_executed.sum", "warp_insts"), ("sm__cycles_elapsed.avg", "cycles")]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("== " + r[ki].split("(")[0])
    for key, name in KEYS:
        hits = [i for i, h in enumerate(hdr) if h == key]
        if hits and r[hits[0]] not in ("", "n/a"):
            print("   %-14s %s %s" % (name, r[hits[0]], units[hits[0]]))
    if len(sys.argv) > 2:   # extra substring filters
        for i, h in enumerate(hdr):
            if any(s in h for s in sys.argv[2:]) and r[i] not in ("", "n/a", "0"):
                print("   %-70s %s %s" % (h, r[i], units[i]))
