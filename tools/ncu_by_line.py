"""Join an ncu --page source SASS dump with nvdisasm line info and aggregate per source line.

usage: python tools/ncu_by_line.py <report.ncu-rep> <kernel-substring> [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "deacon_server_b200", "libdeacon_cuda.so")
tmp = "/tmp/ncu_by_line"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=tmp, capture_output=True, text=True).stdout

# offset -> (file, line) for the requested kernel
line_of = {}
in_k = False
cur = ("?", 0)
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        in_k = kname in ln
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')
agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
tot_inst = tot_samp = 0
for blk in blocks[1:2]:  # first launch only
    rows = list(csv.reader(io.StringIO(blk.split("\n", 1)[1])))
    hdr = rows[0]
    ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        (key, _txt) = line_of.get(a - base, (("?", 0), ""))
        e = agg[key]
        e[0] += int(r[ii]); e[1] += int(r[it]); e[2] += int(r[isamp])
        for i, h in stall_cols:
            e[3][h] += int(r[i] or 0)
        tot_inst += int(r[ii]); tot_samp += int(r[isamp])

print(f"total warp-instructions {tot_inst:,}  samples {tot_samp:,}")
src_cache = {}
def src(f, l):
    for d in ("deacon_server_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[l - 1].strip()[:90] if 0 < l <= len(L) else ""
    return ""
print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s}  top stalls / source")
for key, e in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    st = sorted(e[3].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{h[6:]}={v}" for h, v in st if v)
    print(f"{key[0] + ':' + str(key[1]):28s} {100*e[0]/tot_inst:6.2f} {100*e[2]/max(tot_samp,1):6.2f}  [{sts}] {src(*key)}")

# ---- per-phase roll-up (line ranges of dcn_tile.cuh)
if os.environ.get("NCU_PHASES", "1") == "1":
    import bisect
    tile = os.path.join(root, "deacon_server_b200", "csrc", "dcn_tile.cuh")
    marks = []
    for i, ln in enumerate(open(tile).read().splitlines(), 1):
        m = re.match(r"DCN_HD\s+\S+\s+\*?(\w+)\(", ln)
        if m:
            marks.append((i, m.group(1)))
    starts = [m[0] for m in marks]
    ph = defaultdict(lambda: [0, 0, defaultdict(int)])
    for key, e in agg.items():
        if key[0] == "dcn_tile.cuh":
            j = bisect.bisect_right(starts, key[1]) - 1
            name = marks[j][1] if j >= 0 else "?"
        else:
            name = key[0]
        ph[name][0] += e[0]; ph[name][1] += e[2]
        for h, v in e[3].items():
            ph[name][2][h] += v
    print("\nper function: inst%  samp%  top stalls")
    for name, e in sorted(ph.items(), key=lambda kv: -kv[1][0]):
        st = sorted(e[2].items(), key=lambda kv: -kv[1])[:3]
        print(f"{name:28s} {100*e[0]/tot_inst:6.2f} {100*e[1]/max(tot_samp,1):6.2f}  " + " ".join(f"{h[6:]}={100*v/max(tot_samp,1):.1f}%" for h, v in st if v))
