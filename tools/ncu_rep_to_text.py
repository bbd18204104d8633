#!/usr/bin/env python3
"""One compact text summary per kernel out of a multi-launch `ncu --set full` report (first launch of each kernel name):
duration, DRAM bytes and throughput, tensor / FP64 / ALU / FMA / LSU pipe utilisation, issue slots, occupancy, registers,
shared-memory bank conflicts.  usage: python tools/ncu_rep_to_text.py report.ncu-rep > profiles/rNN_kernels_full.txt"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/LSU wavefronts %"),
        ("sm__inst_executed.sum", "warp instructions"), ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA (FP64 tensor) %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]
seen = set()
print(f"# per-kernel summary of {rep} (ncu --set full --clock-control none; first profiled launch of every kernel)")
for r in data:
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    if name in seen:
        continue
    seen.add(name)
    print(f"\n== {name}")
    for key, label in WANT:
        if key in ix and r[ix[key]] not in ("", "n/a"):
            print(f"   {label:34s} {r[ix[key]]} {units[ix[key]]}")
