"""Feasibility probe for a hybrid ingest: can the box keep a pinned H2D copy at full rate while N host threads
pack other reads (dcn_pack_ascii: reads 1 B/bp, writes 0.375 B/bp to pinned memory)?"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deacon_server_b200 as d

lib = d.load()
n = 1_500_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.random_(65, 85)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
per = 64 << 20
src = torch.empty(16 * per, dtype=torch.uint8).pin_memory()
src.random_(65, 85)
codes = torch.empty(16 * per // 4, dtype=torch.uint8).pin_memory()
inv = torch.empty(16 * per // 8, dtype=torch.uint8).pin_memory()


def h2d_rate(reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


print(f"H2D alone: {h2d_rate():.1f} GB/s")
for threads in (2, 4, 6, 8, 12):
    stop = False
    packed = [0] * threads

    def worker(i):
        b = src.data_ptr() + i * per
        c = codes.data_ptr() + i * per // 4
        v = inv.data_ptr() + i * per // 8
        while not stop:
            lib.dcn_pack_ascii(C.c_void_p(b), C.c_uint64(per), C.c_void_p(c), C.c_void_p(v))
            packed[i] += per

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    time.sleep(0.2)
    p0, t0 = sum(packed), time.perf_counter()
    rate = h2d_rate()
    dt = time.perf_counter() - t0
    prate = (sum(packed) - p0) / dt / 1e9
    stop = True
    for t in ts:
        t.join()
    total = rate + prate * 1.0
    print(f"{threads:2d} packers: H2D {rate:.1f} GB/s beside {prate:.1f} GB/s of packing -> {rate + prate:.1f} Gbp/s of reads ingested")
