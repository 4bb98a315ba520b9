import time, ctypes as C, numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from sdrdaemon_b200 import capi
lib = capi.load()
nf, F = 4096, 32
rng = np.random.default_rng(1)
x = rng.integers(-32768, 32768, size=(1, nf * 16129, 2), dtype=np.int16)
sk = capi.Sink(max_samples=nf * 16129, n_fec=F)
frames = sk.write(x)[0]
sb = np.zeros((nf, 128, 512), np.uint8)
for f in range(nf):
    keep = np.ones(128, bool); keep[rng.permutation(128)[:20]] = False
    sb[f, :108] = frames[f, :128][keep]; sb[f, 108:] = frames[f, 128:148]
d_sb = torch.from_numpy(sb).cuda()
d_nb = torch.full((nf,), 128, dtype=torch.int32, device="cuda")
d_pay = torch.empty((nf, 127, 508), dtype=torch.uint8, device="cuda")
d_b0 = torch.empty((nf, 508), dtype=torch.uint8, device="cuda")
d_st = torch.empty((nf,), dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
def step():
    lib.check(lib.sdrd_fec_decode_dev(d_sb.data_ptr(), 128, d_nb.data_ptr(), nf, d_pay.data_ptr(), d_b0.data_ptr(), d_st.data_ptr(), C.c_void_p(stream.cuda_stream)))
for _ in range(3): step()
stream.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); stream.synchronize(); t2 = time.perf_counter()
    print(f"host call {1e3*(t1-t0):.3f} ms, until done {1e3*(t2-t0):.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20): step()
e1.record(stream); stream.synchronize()
print("20 steps: per step", e0.elapsed_time(e1)/20, "ms")
