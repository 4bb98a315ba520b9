"""Host -> device copy time of page-locked buffers by size (CUDA events; what a queued chain's copy costs)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
bind = bench.bind_near_gpu(0) if len(sys.argv) > 1 and sys.argv[1] == "bind" else None
torch.cuda.init()
dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
res = {"bind": bind, "sizes": {}}
for kb in (256, 1024, 2048, 4096, 8192, 32768, 65536):
    n = kb << 10
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    ts = []
    with torch.cuda.stream(st):
        for rep in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            dev[:n].copy_(h, non_blocking=True)
            b.record(st)
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res["sizes"][f"{kb}KB"] = {"us_median": round(ts[len(ts) // 2], 1), "GB_per_s": round(n / ts[len(ts) // 2] / 1e3, 1)}
print(json.dumps(res))
