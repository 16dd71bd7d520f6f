import sys, time, numpy as np, torch
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, 'tests'))
import deacon_server_b200 as d, helpers as H
gpu = d.DeaconGpu(0)
g = H.random_genome(20_000_000, 1)
rng = np.random.default_rng(2)
N = 4_000_000
pos = rng.integers(0, len(g) - 150, N)
bases = g[(pos[:, None] + np.arange(150)[None, :])].reshape(-1).copy()
off = np.arange(N + 1, dtype=np.uint64) * np.uint64(150)
import ctypes as C
import os
cap = int(0.11 * len(bases)) + 1024
hb = torch.from_numpy(bases).pin_memory(); ho = torch.from_numpy(off.view(np.int64)).pin_memory()
oh = torch.empty(cap, dtype=torch.int64).pin_memory(); op = torch.empty(cap, dtype=torch.int32).pin_memory()
oo = torch.empty(N + 1, dtype=torch.int64).pin_memory()
for mode in ("tiles", "generic"):
    if mode == "generic":
        os.environ["DCN_EXTRACT_GENERIC"] = "1"
    for it in range(3):
        t0 = time.perf_counter()
        rc = gpu._lib.dcn_extract(gpu._ctx, 0, hb.data_ptr(), ho.data_ptr(), N, 31, 15, 0, 0.0, oh.data_ptr(), op.data_ptr(), oo.data_ptr(), cap)
        dt = time.perf_counter() - t0
        assert rc == 0
    print(f"B3 {mode} (pinned in/out, host-pointer call): {len(bases)/dt/1e9:.2f} Gbp/s, {int(oo[-1])} minimizers, {dt*1e3:.2f} ms")
