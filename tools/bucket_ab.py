#!/usr/bin/env python
"""Bucket-width A/B: what one probe costs the memory system when a bucket is 32 bytes (4 keys: the
table's layout), 64 bytes (8 keys) or 128 bytes (16 keys).  Random buckets of a table of the headline size; four probes
in flight per thread, 2368 CTAs.  Prints probes/s per width; run the same command under
  ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum -k regex:random_access
for the DRAM bytes and L2 sectors per probe (profiles/r2_bucket_ab.json holds both)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deacon_server_b200 as d  # noqa: E402

n_keys = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 387_000_000
dev = torch.device("cuda", 0)
keys = torch.randint(-2**63, 2**63 - 1, (n_keys,), dtype=torch.int64, device=dev)
gpu = d.DeaconGpu(0)
gpu.index_upload_device(keys, d.IndexHeader(2, 31, 15))
torch.cuda.synchronize()
out = {"table_bytes": gpu.index_info()["table_bytes"], "n_keys": n_keys, "widths": {}}
for sectors in (1, 2, 4, -4):
    gpu.measure_random_access_wide(sectors, 1 << 27)
    n, ms = gpu.measure_random_access_wide(sectors, 1 << 28)
    if sectors == -4:   # four lanes share a probe: one 128-byte line request per probe
        out["widths"]["128_one_line_request_per_probe"] = {"probes": n, "ms": round(ms, 3), "gprobes_per_s": round(n / ms / 1e6, 2),
                                                           "requested_gb_per_s": round(n * 128 / ms / 1e6, 1)}
        continue
    out["widths"][str(32 * sectors)] = {"probes": n, "ms": round(ms, 3), "gprobes_per_s": round(n / ms / 1e6, 2),
                                        "requested_gb_per_s": round(n * 32 * sectors / ms / 1e6, 1)}
print(json.dumps(out))
