import torch, time
n = 1_580_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = torch.empty(45_000_000, dtype=torch.uint8, device="cuda"); ho = torch.empty(45_000_000, dtype=torch.uint8).pin_memory()
for chunk in (n, 32 << 20, 8 << 20):
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for a in range(0, n, chunk):
            d[a:a + chunk].copy_(h[a:a + chunk], non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H2D pinned, chunks of {chunk >> 20} MB: {n / ms / 1e6:.1f} GB/s ({ms:.2f} ms)")
# H2D with concurrent D2H on another stream
s2 = torch.cuda.Stream()
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
with torch.cuda.stream(s2):
    for _ in range(8): ho.copy_(out, non_blocking=True)
d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D beside D2H: {n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
