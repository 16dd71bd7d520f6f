"""Extract the metrics DESIGN.md / bench.py quote from an ncu report (first profiled launch).
usage: python tools/ncu_summary.py <report.ncu-rep> <out.json> [note]"""
import csv
import json
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
pat = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"launch__(grid_size|block_size|registers_per_thread|occupancy_limit_\w+|waves_per_multiprocessor)|lts__t_sector_hit_rate\.pct|"
                 r"sm__cycles_elapsed\.avg|sm__inst_executed_pipe_(alu|fma|lsu|xu)\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                 r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                 r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                 r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|Kernel Name)$")
m = {}
for i, n in enumerate(h):
    if pat.match(n):
        m[n] = {"value": v[i], "unit": u[i]} if u[i] else v[i]
json.dump({"report": rep, "note": note, "metrics": m}, open(out, "w"), indent=1)
print(json.dumps(m, indent=1)[:3000])
