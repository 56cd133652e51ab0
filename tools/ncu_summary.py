"""Compact per-pipe summary of an `ncu --set full` report (run here, no GPU needed):

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_pipes.csv

One CSV row per selected raw metric (name, unit, value) for every profiled launch: duration, registers, issue-slot and
per-pipe utilisation (XU/MUFU, FMA, ALU, tensor, LSU, shared), warp-stall breakdown per issued instruction, DRAM bytes."""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+)|"
    r"launch__shared_mem_per_block_dynamic|sm__cycles_elapsed\.avg|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__inst_executed\.sum|"
    r"sm__inst_executed_pipe_(xu|fma|alu|lsu|tmem|uniform|adu|cbu|tc|fma_type_fp16)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(tensor|fma|alu|shared|fmaheavy|tc)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"dram__bytes_(read|write)\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"l1tex__t_requests_pipe_lsu_mem_local_op_(ld|st)\.sum|smsp__warps_active\.avg\.per_cycle_active)$")

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name")
w = csv.writer(sys.stdout)
w.writerow(["kernel", "metric", "unit", "value"])
for r in rows[2:]:
    kern = re.sub(r"\(.*", "", r[name_col])
    for i, h in enumerate(hdr):
        if KEEP.match(h):
            w.writerow([kern, h, units[i], r[i]])
