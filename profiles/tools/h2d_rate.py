import torch, time
n = 611095552
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
s = torch.cuda.Stream()
for chunk in (n, n // 8, n // 64):
    with torch.cuda.stream(s):
        for _ in range(2):
            d.copy_(h, non_blocking=True)
        s.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            for o in range(0, n, chunk):
                d[o:o+chunk].copy_(h[o:o+chunk], non_blocking=True)
        e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("chunk", chunk, "H2D GB/s", n / ms / 1e6)
