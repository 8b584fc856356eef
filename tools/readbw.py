import torch, time
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda").normal_()   # 4 GiB
for fn, name, nbytes in [(lambda: x.sum(), "sum fp32 (read-only)", x.numel() * 4),
                         (lambda: x.add_(1.0), "add_ (read+write)", x.numel() * 8),
                         (lambda: x.clone(), "clone (read+write)", x.numel() * 8)]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(10):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%-24s %.3f ms  %.0f GB/s" % (name, best, nbytes / best / 1e6))
