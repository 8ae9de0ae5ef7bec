"""H2D rate of 148 pinned 1241x376 fp32 images (the e2e floor of nalo_track_frames)."""
import time, torch
n, sz = 148, 1241 * 376
hs = [torch.empty(sz, dtype=torch.float32).pin_memory() for _ in range(n)]
d = torch.empty(n * sz, dtype=torch.float32, device="cuda")
big = torch.empty(n * sz, dtype=torch.float32).pin_memory()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i, h in enumerate(hs):
        d[i * sz:(i + 1) * sz].copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d.copy_(big, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"148 separate copies: {1e3*(t1-t0):.3f} ms = {n*sz*4/(t1-t0)/1e9:.1f} GB/s ; one 276 MB copy: {1e3*(t2-t1):.3f} ms = {n*sz*4/(t2-t1)/1e9:.1f} GB/s")
