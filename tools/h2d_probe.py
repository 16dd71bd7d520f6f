#!/usr/bin/env python
"""Host-ingest ceiling of a box: pinned host -> device bandwidth per rank, alone and with all ranks at once.
  python tools/h2d_probe.py                                                       (one GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py   (N ranks of one host)
Prints one JSON line per rank 0; the same probe runs inside bench.py (key "e2e.ingest_ceiling")."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deacon_server_b200 import parallel as par  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
p = par.h2d_probe(dev)
p["pack_threads_policy"] = par.pack_threads_for_rank(int(os.environ.get("LOCAL_WORLD_SIZE", world)), p)
p["cpus"] = os.cpu_count()
if int(os.environ.get("RANK", "0")) == 0:
    print(json.dumps(p))
if world > 1:
    dist.destroy_process_group()
