import torch, time
x = torch.empty(520093696//4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for n in (1, 8):
    for it in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        chunks = x.chunk(n); dch = d.chunk(n)
        for a, b in zip(chunks, dch): b.copy_(a, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print(n, "chunks:", e0.elapsed_time(e1), "ms", 520.09/e0.elapsed_time(e1), "GB/s")
