#!/usr/bin/env python
"""Condense .ncu-rep files (ncu --set full) into a small text table under profiles/: per captured launch the
duration, DRAM bytes and throughput, tensor-pipe activity, issue rate, registers and occupancy.
usage: python scripts/ncu_summary.py out.txt rep1.ncu-rep [rep2 ...]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma sub-pipe active %"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (TPC triage, elapsed)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
        ("smsp__issue_active.avg.pct", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
        ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
        ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait"),
        ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
        ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard")]
out = ["# ncu --set full --clock-control none summaries (one row block per captured launch; cold-cache, serialised:",
       "# use for shares and pipe utilisation, not for absolute step time) — scripts/ncu_summary.py", ""]
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        out.append(f"{rep}: no data")
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out.append(f"## {rep.split('/')[-1]}")
    for r in rows[2:]:
        out.append("kernel: " + r[idx["Kernel Name"]][:140])
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                out.append(f"    {label:38s} {r[idx[k]]} {units[idx[k]]}")
        out.append("")
open(sys.argv[1], "w").write("\n".join(out))
print(sys.argv[1], len(out), "lines")
