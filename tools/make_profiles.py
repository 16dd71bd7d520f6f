#!/usr/bin/env python
"""Turns the scratch ncu artefacts of a gpurun call into the tracked evidence under profiles/.
usage: python tools/make_profiles.py TAG REPORT.ncu-rep LAUNCHES.csv "command that was profiled"
writes profiles/TAG_warp_ncu_summary.json, profiles/TAG_warp_by_line.txt, profiles/TAG_launches_bench.csv and
profiles/traffic.json (what bench.py's roofline block quotes as ncu_*)."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, launches, cmd = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
P = os.path.join(ROOT, "profiles")
summ = os.path.join(P, f"{tag}_warp_ncu_summary.json")
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, summ,
                       f"ncu --set full --clock-control none, one launch of filter_warp_kernel inside: {cmd}"], stdout=subprocess.DEVNULL)
with open(os.path.join(P, f"{tag}_warp_by_line.txt"), "w") as f:
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), rep, "filter_warp_kernelILb0ELb0", "70"], stdout=f)
shutil.copy(launches, os.path.join(P, f"{tag}_launches_bench.csv"))
m = json.load(open(summ))["metrics"]


def val(k):
    v = m[k]
    return float((v["value"] if isinstance(v, dict) else v).replace(",", ""))


def scaled(k):   # ncu prints bytes with a unit prefix
    v = m[k]
    x = float(v["value"].replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[v["unit"]]


rd, wr = scaled("dram__bytes_read.sum"), scaled("dram__bytes_write.sum")
ms = val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}.get(m["gpu__time_duration.sum"]["unit"], 1.0)
traffic = {
    "kernel": "filter_warp_kernel<ASCII>",
    "report": f"profiles/{tag}_warp_ncu_summary.json",
    "command": cmd,
    "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
    "kernel_ms": round(ms, 4),
    "alu_pipe_pct_of_peak": round(val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
    "fma_pipe_pct_of_peak": round(val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
    "issue_active_pct": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
    "warp_instructions_per_launch": int(val("smsp__inst_executed.sum")),
    "l2_hit_rate_pct": round(val("lts__t_sector_hit_rate.pct"), 2),
    "note": "a probe of a random 32-byte bucket costs one L2 request and ~126 bytes of DRAM traffic (profiles/r2_bucket_ab.json); "
            "the ASCII bases are read once (TMA bulk copies); per-launch numbers of the 5 M-pair (1.5 Gbp) step",
}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
# launch-list shares of the step
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
tot = {}
body = rows[hi + 1:]
# the device-resident leg comes first; the end-to-end leg behind it launches the packed instantiation: stop there
first_packed = next((i for i, r in enumerate(body) if len(r) > kn and "filter_warp_kernel<1" in r[kn].replace("(bool)", "")), len(body))
for r in body[:first_packed]:
    if len(r) > mv:
        name = r[kn].split("(")[0].replace("void ", "").replace("dcn::", "")
        tot[name] = tot.get(name, 0) + float(r[mv].replace(",", ""))
STEP = ("wplan_kernel", "filter_warp_kernel", "filter_tail_kernel", "commit_counters_kernel")   # one device-resident step
ours = {k: v for k, v in tot.items() if any(k.startswith(n) for n in STEP)}
s = sum(ours.values())
shares = {k: round(v / s, 4) for k, v in sorted(ours.items(), key=lambda kv: -kv[1])}
json.dump({"launch_list": f"profiles/{tag}_launches_bench.csv", "command": cmd,
           "share_of_a_device_resident_step": shares,
           "note": "ncu per-launch times are cold-cache and serialised: the SHARES are what compares with bench.py's kernel_share_of_step"},
          open(os.path.join(P, f"{tag}_launch_shares.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
print(json.dumps(shares, indent=1))
