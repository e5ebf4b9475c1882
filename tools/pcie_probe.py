import torch, time
for nbytes in [1<<10, 64<<10, 418<<10, 4<<20, 64<<20]:
    d = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt2 = (time.perf_counter() - t0) / n
    print(f"{nbytes:>10d} B  D2H {dt*1e6:8.1f} us ({nbytes/dt/1e9:6.2f} GB/s)   H2D {dt2*1e6:8.1f} us ({nbytes/dt2/1e9:6.2f} GB/s)")
