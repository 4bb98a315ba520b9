#!/usr/bin/env python
"""bench.py -- throughput of sdrdaemon's Rx hot path (decimate + superframe FEC encode) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N = 1 (BASELINE.json configs[1]): TestSource-shaped 10 Msps int16 I/Q, decimate-by-16 (4 half-band
stages, centred), 128 data + 16 FEC blocks per superframe, one stream.  A step is one pass of the hot path over
one batch of FRAMES superframes (FRAMES * 258064 input samples, ~611 MB: larger than the 126 MB L2, so no flush
is needed between steps).
Workload at N > 1 (BASELINE.json configs[4], the configuration its scaling claim is stated on): 2048 streams
@ 61.44 Msps, decimate-by-64, 128 + 32 FEC, 2 superframes per stream per step, the streams sharded in contiguous
ranges over the N ranks -- STRONG scaling, no data-path collective; NCCL carries the barrier, the digest gather
and, in the drop-in figure, the scatter of the streams from rank 0 (overlapped with the shards' processing).
`--config 2|3|5` selects a workload explicitly (3 / 5 at N = 1: one GPU's shard of 256 streams).

One JSON line on rank 0:
  value        input Msamples/s, whole job, inputs resident in HBM, CUDA events, max over ranks
  e2e          the same metric through the C ABI with HOST buffers (sdrd_rx_process: H2D + kernels + D2H)
  roofline     the dominant kernel (K1, the half-band cascade) against the measured HBM copy bandwidth
  cpu_baseline the reference's CPU path on this box's host cores, bounded sample (rank 0, N=1 only)
--impl reference prints the CPU arm as its own line (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

FRAME_SAMPLES = 127 * 127
# BASELINE.json configs; 2 is the headline (default), the others are optional extra measurements
WORKLOADS = {
    2: dict(M=4, F=16, S=1, frames=592, rate=10_000_000,
            name="config2: 10 Msps int16 IQ stream, decimate-by-16 centred (4 half-band stages), 128+16 FEC"),
    3: dict(M=5, F=32, S=256, frames=8, rate=20_000_000,
            name="config3: 256 streams @ 20 Msps, decimate-by-32, 128+32 FEC"),
    5: dict(M=6, F=32, S=256, frames=2, rate=61_440_000,
            name="config5 (per-GPU shard): 256 streams @ 61.44 Msps, decimate-by-64, 128+32 FEC"),
}
M_LOG2, N_FEC, N_STREAMS = 4, 16, 1
FRAME_IN = FRAME_SAMPLES << M_LOG2  # input samples per superframe (258064 at decimate-by-16)


def select_workload(cfg: int, frames=None):
    global M_LOG2, N_FEC, N_STREAMS, FRAME_IN
    w = dict(WORKLOADS[cfg])
    if frames:
        w["frames"] = frames
    M_LOG2, N_FEC, N_STREAMS = w["M"], w["F"], w["S"]
    FRAME_IN = FRAME_SAMPLES << M_LOG2
    return w
METRIC = "Msamples/s IQ through decimate+FEC"
UNIT = "Msamples/s"


def algorithmic_bytes_per_sample(m: int, f: int) -> float:
    # read 4 B per input sample, write (128+F) datagrams of 512 B per 16129 * 2^M samples
    return 4.0 + (128 + f) * 512.0 / (FRAME_SAMPLES * (1 << m))


def k1_bytes_per_sample(m: int) -> float:
    # decimator alone: 4 B in, 4 B out per 2^M
    return 4.0 + 4.0 / (1 << m)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "50", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, name in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # the sampler also sees the idle moments around the region: the upper half is "under load"
        top = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(seconds: float, cores: int, rng):
    """A bounded sample of the workload for the CPU arm: `cores` streams of whole superframes sized so that one
    pass takes about `seconds`.  Returns (x, frames)."""
    from oracle import bindings as ob

    frames_cal = 2
    x = rng.integers(-32768, 32768, size=(cores, frames_cal * FRAME_IN, 2), dtype=np.int16)
    t0 = time.perf_counter()
    ob.cpu_rx_streams(x, M_LOG2, N_FEC, cores)
    dt = time.perf_counter() - t0
    rate = x.shape[0] * x.shape[1] / dt
    frames = max(frames_cal, min(64, int(rate * seconds / (cores * FRAME_IN))))
    if frames != frames_cal:
        x = rng.integers(-32768, 32768, size=(cores, frames * FRAME_IN, 2), dtype=np.int16)
    return x, frames


def cpu_pass(x, frames: int, cores: int):
    """One timed pass of the reference CPU path over the sample: `cores` concurrent streams (the reference runs
    one stream per thread; Decimators state is sequential)."""
    from oracle import bindings as ob

    t0 = time.perf_counter()
    fr, dig, kind = ob.cpu_rx_streams(x, M_LOG2, N_FEC, cores)
    dt = time.perf_counter() - t0
    assert fr == cores * frames, (fr, cores, frames)
    msps = x.shape[0] * x.shape[1] / dt / 1e6
    fec = ("restated CM256 with the SSSE3 byte-shuffle block multiply (what cm256cc runs on x86)" if ob.simd() else
           "restated CM256, scalar 256-entry tables (no SSSE3 on this host)")
    sample = (f"{cores} streams x {frames} superframes ({cores * frames * FRAME_IN} samples) of the workload, one stream per "
              f"thread, 65536-sample blocks; decimator = reference Decimators.cpp (EO1/SSE4.1 build), FEC = {fec}; the GPU arm "
              f"runs ONE stream of {WORKLOADS[2]['frames']} superframes where the workload is config 2: a single stream cannot use "
              f"more than one host core, so the CPU arm is given one stream per core" if kind == "reference" else
              f"{cores} streams x {frames} superframes of the workload; C restatement (oracle port), FEC = {fec}")
    return {"value": round(msps, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, dt


def cpu_leg(seconds: float, cores: int):
    x, frames = cpu_sample(seconds, cores, np.random.default_rng(1234))
    return cpu_pass(x, frames, cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    try:
        os.sched_getaffinity
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # every step is one pass over the same bounded sample; the sample is sized so that the whole run
    # (warm-up + steps) stays near one minute whatever --steps says
    n_pass = max(args.warmup + args.steps, 1)
    per_step = min(2.0, 60.0 / n_pass)
    x, frames = cpu_sample(per_step, cores, np.random.default_rng(1234))
    vals, times = [], []
    base = None
    for i in range(n_pass):
        base, dt = cpu_pass(x, frames, cores)
        if i >= args.warmup:
            vals.append(base["value"])
            times.append(dt)
    v = statistics.mean(vals)
    base["value"] = round(v, 3)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * statistics.mean(times), 3), "higher_is_better": True,
        "scaling": "strong" if getattr(args, "strong", False) else "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config]["name"] + " (reference CPU path, bounded sample per step)",
                   "log2_decim": M_LOG2, "n_fec": N_FEC},
        "cpu_baseline": base,
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def decode_cpu_leg(sb, seconds):
    """config 4 on the host, all cores: the reference's own SDRdaemonFECBuffer (oracle/_ref) when its build is present,
    else the oracle port; bounded to about `seconds`."""
    from oracle import bindings as ob

    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kind = "reference" if ob.ref_available(0) else "port"
    run = (lambda a: ob.ref_decode_frames(a, 128, cores)) if kind == "reference" else (lambda a: ob.decode_frames(a, 128, cores))
    run(sb[: min(len(sb), 4 * cores)])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        run(sb)
        n += len(sb)
    dt = time.perf_counter() - t0
    fec = "SSSE3 byte-shuffle block multiply" if ob.simd() else "scalar tables"
    return {"value": round(n / dt / 1e6, 6), "unit": "Msuperframes/s", "cores": cores, "kind": kind,
            "sample": f"{n} superframes ({len(sb)} distinct) of the workload on {cores} threads: " +
                      ("SDRdaemonFECBuffer.cpp of the reference (oracle/_ref)" if kind == "reference" else "SDRdaemonFECBuffer logic of the oracle") +
                      f" + restated CM256 ({fec})"}


def run_decode(args):
    """BASELINE config 4: 4096 superframes, 20 of 128 blocks erased (F = 32), recover on the GPU."""
    import ctypes as C

    import torch

    from sdrdaemon_b200 import capi

    if args.impl == "reference":
        from oracle import bindings as ob

        rng = np.random.default_rng(0xFEC0)
        nf, F = 64, 32
        x = rng.integers(-32768, 32768, size=(nf * FRAME_SAMPLES, 2), dtype=np.int16)
        sk = ob.Sink(n_fec=F)
        sk.write(x)
        frames = np.stack(sk.frames)
        sbs = []
        for f in range(nf):
            keep = np.ones(128, bool)
            keep[rng.permutation(128)[:20]] = False
            sbs.append(np.concatenate([frames[f, :128][keep], frames[f, 128:148]]))
        base = decode_cpu_leg(np.stack(sbs), 10.0)
        v = base["value"]
        print(json.dumps({"impl": "reference", "metric": "Msuperframes/s recovered (20/128 erasures)", "value": v,
                          "unit": "Msuperframes/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "higher_is_better": True,
                          "config": {"workload": "config4: superframes with 20/128 random erasures, F=32 (CPU decode, bounded sample)"},
                          "cpu_baseline": base,
                          "e2e": {"value": v, "unit": "Msuperframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    lib = capi.load()
    torch.cuda.set_device(0)
    rng = np.random.default_rng(0xFEC0)
    nf, F = 4096, 32
    x = rng.integers(-32768, 32768, size=(1, nf * FRAME_SAMPLES, 2), dtype=np.int16)
    sk = capi.Sink(max_samples=nf * FRAME_SAMPLES, n_fec=F)
    frames = sk.write(x)[0]
    sb = np.zeros((nf, 128, 512), np.uint8)
    for f in range(nf):
        keep = np.ones(128, bool)
        keep[rng.permutation(128)[:20]] = False
        sb[f, :108] = frames[f, :128][keep]
        sb[f, 108:] = frames[f, 128:148]
    d_sb = torch.from_numpy(sb).cuda()
    d_nb = torch.full((nf,), 128, dtype=torch.int32, device="cuda")
    d_pay = torch.empty((nf, 127, 508), dtype=torch.uint8, device="cuda")
    d_b0 = torch.empty((nf, 508), dtype=torch.uint8, device="cuda")
    d_st = torch.empty((nf,), dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()

    def step():
        lib.check(lib.sdrd_fec_decode_dev(d_sb.data_ptr(), 128, d_nb.data_ptr(), nf, d_pay.data_ptr(), d_b0.data_ptr(),
                                          d_st.data_ptr(), C.c_void_p(stream.cuda_stream)))

    for _ in range(max(args.warmup, 3)):
        step()
    stream.synchronize()
    ok = bool((d_st == 2).all()) and bool((d_pay.cpu().numpy() == frames[:, 1:128, 4:]).all())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = min(args.steps, 1000)
    sampler = ClockSampler(0)
    sampler.start()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    stream.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    peak, peak_src = measured_peaks()
    alg = nf * (128 * 512 + 127 * 508)
    k3_traffic, k3_traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k3_traffic.json")))
        if t.get("frames_per_launch") == nf:
            k3_traffic, k3_traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
    except Exception:
        k3_traffic = None
    # end to end through the host-pointer C ABI (what UDPSourceFEC would call): pinned host buffers, H2D of the
    # received datagrams + kernels + D2H of the payload inside the timed region
    e2e = None
    if not args.no_e2e:
        h_sb = torch.from_numpy(sb).pin_memory()
        h_nb = torch.full((nf,), 128, dtype=torch.int32).pin_memory()
        h_pay = torch.empty((nf, 127, 508), dtype=torch.uint8).pin_memory()
        h_b0 = torch.empty((nf, 508), dtype=torch.uint8).pin_memory()
        h_st = torch.empty((nf,), dtype=torch.int32).pin_memory()

        def e2e_step():
            lib.check(lib.sdrd_fec_decode(h_sb.data_ptr(), 128, h_nb.data_ptr(), nf, h_pay.data_ptr(), h_b0.data_ptr(), h_st.data_ptr()))

        e2e_step()
        e2e_ok = bool((h_st == 2).all()) and bool((h_pay.numpy() == frames[:, 1:128, 4:]).all())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": round(nf / dt / 1e6, 3), "unit": "Msuperframes/s", "h2d_bytes_per_step": int(sb.nbytes + 4 * nf),
               "d2h_bytes_per_step": int(h_pay.numel() + h_b0.numel() + 4 * nf), "steps": args.e2e_steps,
               "api": "sdrd_fec_decode (host pointers, pinned)", "parity": "ok" if e2e_ok else "MISMATCH"}
    cpu = None
    if not args.no_cpu:
        cpu = decode_cpu_leg(sb[:512], 10.0)
    print(json.dumps({"metric": "Msuperframes/s recovered (20/128 erasures)", "value": round(nf / ms / 1e3, 3),
                      "unit": "Msuperframes/s", "n_gpus": 1, "steps": steps, "warmup": args.warmup, "ms_per_step": round(ms, 4),
                      "higher_is_better": True, "dtype": "u8 (GF(2^8))", "data": "synthetic",
                      "config": {"workload": "config4: 4096 superframes, 20/128 random erasures, F=32, bit-exact recover",
                                 "parity": "all frames recovered == transmitted" if ok else "MISMATCH"},
                      "gpu_launches": 2 * steps, "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
                      "roofline": {"bound": "hbm", "kernel": "fec::decode_stream_kernel (K3)", "achieved": round(alg / ms / 1e6, 1),
                                   "peak": peak, "unit": "GB/s", "frac": round(alg / ms / 1e6 / peak, 4), "traffic": k3_traffic,
                                   "traffic_source": k3_traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                                   "note": "K3 is ALU-pipe bound (PRMT/LOP3 table arithmetic), see DESIGN.md"}}), flush=True)


def run_interp(args):
    """SURVEY 8(f)-1, the Tx side: 592 superframes of 16129 samples interpolated by 16 (Upsampler)."""
    import ctypes as C

    M = args.log2_decim or 4
    nfr = args.frames or 592
    n_in = nfr * FRAME_SAMPLES
    metric = "Msamples/s IQ out of the interpolation cascade"
    if args.impl == "reference":
        from oracle import bindings as ob

        rng = np.random.default_rng(0x7A)
        n = 16 * FRAME_SAMPLES
        x = rng.integers(-32768, 32768, size=(n, 2), dtype=np.int16)
        kind = "reference" if ob.ref_available(0) else "port"
        u = ob.RefUpsampler(M, 0) if kind == "reference" else ob.Interpolator(M)
        t0 = time.perf_counter()
        y = u.process(x)
        dt = time.perf_counter() - t0
        v = len(y) / dt / 1e6
        base = {"value": round(v, 3), "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"{n} input samples, one Upsampler (reference Interpolators.cpp, one thread)"}
        print(json.dumps({"impl": "reference", "metric": metric, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": 1,
                          "warmup": 0, "higher_is_better": True, "config": {"workload": f"tx: interpolate by {1 << M}, CPU, 1 thread"},
                          "cpu_baseline": base, "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    import torch

    from sdrdaemon_b200 import capi

    lib = capi.load()
    torch.cuda.set_device(0)
    u = capi.Interpolator(M, 1, max_in=n_in)
    in_ptr, _ = u.dev_input()
    out_ptr, _ = u.dev_output()
    g = torch.Generator(device="cuda")
    g.manual_seed(0x7A)
    dev_in = torch.as_tensor(DevView(in_ptr, n_in * 4), device="cuda").view(torch.int16)
    dev_in.copy_(torch.randint(-32768, 32768, (n_in * 2,), dtype=torch.int16, device="cuda", generator=g))
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    # parity of the first 2048 inputs against the oracle (outside the timed region)
    parity = "unchecked"
    u.process_dev(n_in, stream.cuda_stream)
    stream.synchronize()
    try:
        from oracle import bindings as ob

        x0 = dev_in[: 2 * 2048].cpu().numpy().reshape(-1, 2)
        y0 = torch.as_tensor(DevView(out_ptr, (2048 << M) * 4), device="cuda").cpu().numpy().view(np.int16).reshape(-1, 2)
        parity = "first 2048 inputs bit-exact vs oracle" if np.array_equal(y0, ob.Interpolator(M).process(x0)) else "MISMATCH vs oracle"
    except Exception as e:
        parity = f"unchecked ({type(e).__name__})"
    if parity.startswith("MISMATCH"):
        raise SystemExit("bench: GPU result differs from the oracle; refusing to report a number")
    for _ in range(max(args.warmup, 3)):
        u.process_dev(n_in, stream.cuda_stream)
    stream.synchronize()
    steps = min(args.steps, 1000)
    l0 = u.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(0)
    sampler.start()
    e0.record(stream)
    for _ in range(steps):
        u.process_dev(n_in, stream.cuda_stream)
    e1.record(stream)
    stream.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    # end to end through Upsampler's host entry point (page-locked buffers, H2D + kernel + D2H per step) and the
    # reference's own Upsampler on one host core (the cascade is sequential per stream)
    e2e, cpu = None, None
    if not args.no_e2e:
        n_e = 64 * FRAME_SAMPLES  # a bounded piece of the stream: the output is 2^M times larger
        hi = torch.empty((n_e, 2), dtype=torch.int16).pin_memory()
        hi.copy_(dev_in[: 2 * n_e].reshape(n_e, 2))
        ho = torch.empty((n_e << M, 2), dtype=torch.int16).pin_memory()
        ue = capi.Interpolator(M, 1, max_in=n_e)
        no = C.c_size_t(0)
        for warm in (True, False):
            t0 = time.perf_counter()
            for _ in range(2 if warm else args.e2e_steps):
                lib.check(lib.sdrd_int_process(ue._h, hi.data_ptr(), n_e, n_e, ho.data_ptr(), n_e << M, C.byref(no)))
            dt = (time.perf_counter() - t0) / (2 if warm else args.e2e_steps)
        e2e = {"value": round((n_e << M) / dt / 1e6, 1), "unit": UNIT, "h2d_bytes_per_step": n_e * 4, "d2h_bytes_per_step": (n_e << M) * 4,
               "steps": args.e2e_steps, "api": "sdrd_int_process (host pointers, pinned)"}
        ue.close()
    if not args.no_cpu:
        try:
            from oracle import bindings as ob

            xs = dev_in[: 2 * 16 * FRAME_SAMPLES].cpu().numpy().reshape(-1, 2)
            kind = "reference" if ob.ref_available(0) else "port"
            uu = ob.RefUpsampler(M, 0) if kind == "reference" else ob.Interpolator(M)
            t0 = time.perf_counter()
            yy = uu.process(xs)
            dtc = time.perf_counter() - t0
            cpu = {"value": round(len(yy) / dtc / 1e6, 3), "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{len(xs)} input samples through one Upsampler ({'reference Interpolators.cpp' if kind == 'reference' else 'oracle port'}), one thread"}
        except Exception:
            cpu = None
    n_out = n_in << M
    peak, peak_src = measured_peaks()
    alg = 4 * n_in + 4 * n_out
    k4_traffic, k4_traffic_src = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "k4_traffic.json")))
        if t.get("log2_interp") == M and t.get("samples_in_per_launch") == n_in:
            k4_traffic, k4_traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
    except Exception:
        k4_traffic = None
    print(json.dumps({"metric": metric, "value": round(n_out / ms / 1e3, 1), "unit": UNIT, "n_gpus": 1, "steps": steps,
                      "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "dtype": "int32", "data": "synthetic",
                      "config": {"workload": f"tx (SURVEY 8f-1): {nfr} superframes ({n_in} samples) interpolated by {1 << M}, "
                                             f"{n_out * 4 / 1e6:.0f} MB out per step", "parity": parity,
                                 "l2": "outputs larger than L2 (no flush needed)"},
                      "gpu_launches": int(u.launches - l0), "e2e": e2e, "cpu_baseline": cpu, "clocks": clocks,
                      "roofline": {"bound": "hbm", "kernel": f"hbi::interpolate_warp_kernel<{min(M, 5)}> (K4)", "achieved": round(alg / ms / 1e6, 1),
                                   "peak": peak, "unit": "GB/s", "frac": round(alg / ms / 1e6 / peak, 4), "traffic": k4_traffic,
                                   "traffic_source": k4_traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg}}), flush=True)


def bind_near_gpu(local: int):
    """Run this rank (and first-touch its page-locked buffers) on the CPUs next to its GPU: the CPUs the PCI device
    reports as local, when the container shows them.  Returns what was done, for the JSON line."""
    try:
        import torch

        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bdf is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip()
            bdf = out[4:] if out.startswith("0000") and len(out) > 12 else out
        bdf = str(bdf).lower()
        if len(bdf.split(":")) == 2:
            bdf = "0000:" + bdf
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(base + "/numa_node").read().strip()
        cpus = open(base + "/local_cpulist").read().strip()
        want = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            want.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(want & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            return {"numa_node": node, "cpus": cpus, "bound": True}
        return {"numa_node": node, "cpus": cpus, "bound": False, "why": "the local CPUs are all the rank is allowed anyway" if use else "no local CPU in the allowed set"}
    except Exception as e:
        return {"bound": False, "why": f"{type(e).__name__}"}


def small_block_leg(lib, capi, n_blocks=4000, blk=65536):
    """The reference's call granularity: ONE stream fed in TestSource-sized blocks of 65536 samples
    (include/TestSource.h:33) from pageable host memory, the way sdrdaemonrx's main loop feeds Downsampler::process +
    UDPSinkFEC::write (sdrdaemonrx.cpp:579-663).  Three forms, same config as the headline:
      sync    sdrd_rx_process once per block (copy in, kernels, copy out, synchronise -- every block)
      queued  sdrd_rx_submit per block + sdrd_rx_collect: no synchronisation per block, blocks batched on the way
      classes the C++ mirror classes (Downsampler::process then UDPSinkFEC::write, host vectors in between)"""
    import ctypes as C

    rng = np.random.default_rng(77)
    src = [rng.integers(-32768, 32768, size=(blk, 2), dtype=np.int16) for _ in range(8)]  # pageable
    ptrs = [a.ctypes.data for a in src]
    bpf = 128 + N_FEC
    out = np.zeros((64, bpf, 512), np.uint8)
    res = {"block_samples": blk, "blocks": n_blocks, "memory": "pageable host vectors, one stream"}
    # what any entry point that accepts pageable memory has to pay first: one host copy of the block into page-locked
    # staging memory (measured with the same blocks and one thread; the queued form is bounded by it)
    try:
        import torch

        stage = torch.empty((64, blk, 2), dtype=torch.int16).pin_memory()   # as large as the queue's two buffers
        sp = stage.data_ptr()
        for warm in (True, False):
            t0 = time.perf_counter()
            for b in range(200 if warm else n_blocks):
                C.memmove(sp + (b & 63) * blk * 4, ptrs[b & 7], blk * 4)
            dt = time.perf_counter() - t0
        res["host_copy_ceiling"] = {"value": round(n_blocks * blk / dt / 1e6, 1), "unit": UNIT, "GB_per_s": round(n_blocks * blk * 4 / dt / 1e9, 2),
                                    "what": "memmove of each block into 16 MB of page-locked memory (not cache resident), one thread"}
    except Exception:
        pass
    rx = capi.Rx(M_LOG2, n_streams=1, max_in=32 * blk, n_fec=N_FEC)
    nfr = C.c_size_t(0)
    for warm in (True, False):
        n = 500 if warm else n_blocks
        t0 = time.perf_counter()
        for b in range(n):
            lib.check(lib.sdrd_rx_process(rx._h, ptrs[b & 7], blk, blk, out.ctypes.data, 64, C.byref(nfr), None))
        dt = time.perf_counter() - t0
    res["sync"] = {"value": round(n_blocks * blk / dt / 1e6, 1), "unit": UNIT, "us_per_block": round(dt / n_blocks * 1e6, 2),
                   "api": "sdrd_rx_process per block"}
    bp = C.c_int(0)
    for key, min_chain in (("queued", 0), ("queued_batched", 16 * blk)):
        rx.reset()
        rx.set_min_chain(min_chain)
        frames = 0
        for warm in (True, False):
            n = 512 if warm else n_blocks
            c0 = rx.chains
            t0 = time.perf_counter()
            for b in range(n):
                lib.check(lib.sdrd_rx_submit(rx._h, ptrs[b & 7], blk, blk, None))
                lib.check(lib.sdrd_rx_collect(rx._h, out.ctypes.data, 64, C.byref(nfr), C.byref(bp), 0))
                frames += nfr.value
            while True:
                lib.check(lib.sdrd_rx_collect(rx._h, out.ctypes.data, 64, C.byref(nfr), C.byref(bp), 1))
                frames += nfr.value
                if not nfr.value:
                    break
            dt = time.perf_counter() - t0
            chains = rx.chains - c0
        res[key] = {"value": round(n_blocks * blk / dt / 1e6, 1), "unit": UNIT, "us_per_block": round(dt / n_blocks * 1e6, 2),
                    "blocks_per_chain": round(n_blocks / max(chains, 1), 2),
                    "api": "sdrd_rx_submit + sdrd_rx_collect" + (f", sdrd_rx_set_min_chain({min_chain})" if min_chain else " (an idle device starts at once)")}
    rx.close()
    # the C++ mirror classes, through tests/host/host_pipeline.cpp (built here if a compiler is present)
    try:
        exe = os.path.join(ROOT, "tests", "host", "host_pipeline_gpu")
        srcf = os.path.join(ROOT, "tests", "host", "host_pipeline.cpp")
        libdir = os.path.join(ROOT, "sdrdaemon_b200")
        if (not os.path.exists(exe)) or os.path.getmtime(exe) < max(os.path.getmtime(srcf), os.path.getmtime(os.path.join(libdir, "libsdrd_b200.so"))):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-std=c++17", "-O2", "-pthread", "-o", exe, srcf, f"-L{libdir}", "-lsdrd_b200", f"-Wl,-rpath,{libdir}"],
                           check=True, capture_output=True)
        r = subprocess.run([exe, "blocks", str(n_blocks), str(M_LOG2), str(N_FEC), str(blk), "500"], capture_output=True, text=True, timeout=300)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        res["classes"] = {"value": round(j["msamples_per_s"], 1), "unit": UNIT, "us_per_block": j["us_per_block"],
                          "api": "Downsampler::process + UDPSinkFEC::write (sdrd_host.hpp), native caller"}
        # the queued entry points from a native caller (the ctypes legs above pay ~5 us of interpreter per call)
        # (`_staged`: sdrd_rx_set_staging_threads -- helper threads of the handle share the copy into page-locked memory,
        # the way the reference spreads this work over its source / downsampler / sink threads)
        for key, mc, helpers in (("queued_native", 0, 0), ("queued_batched_native", 16, 0), ("queued_native_staged", 0, 2),
                                 ("queued_batched_native_staged", 16, 2)):
            r = subprocess.run([exe, "blocksq", str(5 * n_blocks), str(M_LOG2), str(N_FEC), str(blk), str(mc), "2000", str(helpers)],
                               capture_output=True, text=True, timeout=300)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            res[key] = {"value": round(j["msamples_per_s"], 1), "unit": UNIT, "us_per_block": j["us_per_block"],
                        "blocks_per_chain": j["blocks_per_chain"], "staging_threads": helpers,
                        "api": "sdrd_rx_submit + sdrd_rx_collect from C++" + (f", sdrd_rx_set_min_chain({mc} blocks)" if mc else "")
                               + (f", sdrd_rx_set_staging_threads({helpers})" if helpers else "")}
    except Exception as e:
        res["classes"] = {"value": None, "note": f"not measured ({type(e).__name__})"}
    return res


class DevView:
    """Zero-copy torch view of a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5, 6],
                    help="BASELINE.json config (default: 2, the headline, on one GPU; 5 as stated -- 2048 streams over all "
                         "GPUs -- under torchrun); 6 = the Tx interpolation cascade (SURVEY 8f-1)")
    ap.add_argument("--frames", type=int, default=0, help="superframes per stream per step (default: per config)")
    ap.add_argument("--log2-decim", type=int, default=0, help="experiments: override the workload's log2 decimation")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    strong = False
    if args.config == 0:
        args.config = 2 if world_env == 1 else 5
        strong = world_env > 1
    if args.config == 4:
        return run_decode(args)
    if args.config == 6:
        return run_interp(args)
    if args.log2_decim:
        WORKLOADS[args.config] = dict(WORKLOADS[args.config], M=args.log2_decim,
                                      name=WORKLOADS[args.config]["name"] + f" [log2_decim overridden to {args.log2_decim}]")
    if strong:
        if 2048 % world_env:
            raise SystemExit("config 5 shards 2048 streams: the number of GPUs has to divide it")
        WORKLOADS[5] = dict(WORKLOADS[5], S=2048 // world_env,
                            name=f"config5: 2048 streams @ 61.44 Msps, decimate-by-64, 128+32 FEC, sharded {2048 // world_env} per GPU")
    work = select_workload(args.config, args.frames or None)
    args.frames = work["frames"]
    args.strong = strong
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from sdrdaemon_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the library has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load()
    lib.check(lib.sdrd_set_device(local))
    numa = bind_near_gpu(local) if world > 1 else None

    S = N_STREAMS
    n_in = args.frames * FRAME_IN
    out_rate = work["rate"] >> M_LOG2
    rx = capi.Rx(M_LOG2, n_streams=S, max_in=n_in, n_fec=N_FEC, sample_rate=out_rate)
    dec_h, sink = rx.dec_handle, rx.sink
    in_ptr, in_stride = rx.dev_input()
    # zero-copy view of the handle's input buffer: row s = stream s (pitch in_stride samples)
    dev_all = torch.as_tensor(DevView(in_ptr, ((S - 1) * in_stride + n_in) * 4), device="cuda").view(torch.int16)
    dev_in = torch.as_strided(dev_all, (S, n_in * 2), (in_stride * 2, 1))
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5D12DAE0 + rank)
    for s0 in range(S):  # per stream: keeps the temporary small
        dev_in[s0].copy_(torch.randint(-32768, 32768, (n_in * 2,), dtype=torch.int16, device="cuda", generator=g))
    torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    sptr = stream.cuda_stream
    out_ptr, out_stride = None, None
    import ctypes as C

    def step(ev_a=None, ev_b=None):
        n_out = C.c_size_t(0)
        ss = C.c_uint(16)
        if ev_a is not None:
            ev_a.record(stream)
        lib.check(lib.sdrd_dec_process_dev(dec_h, n_in, C.byref(n_out), C.byref(ss), C.c_void_p(sptr)))
        if ev_b is not None:
            ev_b.record(stream)
        st = C.c_size_t(0)
        op = lib.sdrd_dec_dev_output(dec_h, C.byref(st))
        return sink.write_dev(op, n_out.value, st.value, sptr)

    # ---- first step from reset state: check frame 0 against the oracle (outside the timed region) ----
    parity = "unchecked"
    nfr = step()
    stream.synchronize()
    assert nfr == args.frames, (nfr, args.frames)
    try:
        from oracle import bindings as ob

        dg_ptr, _ = sink.dev_datagrams()
        bpf = 128 + N_FEC
        dg0 = torch.as_tensor(DevView(dg_ptr, 2 * bpf * 512), device="cuda").cpu().numpy().reshape(2, bpf, 512)
        x0 = dev_in[0, : 2 * 2 * FRAME_IN].cpu().numpy().reshape(-1, 2)
        y0, _ = ob.Decimator(M_LOG2).process(x0)
        osk = ob.Sink(n_fec=N_FEC, sample_rate=out_rate)
        osk.write(y0)
        parity = "frames 0-1 bit-exact vs oracle" if np.array_equal(dg0, np.stack(osk.frames)) else "MISMATCH vs oracle"
    except Exception as e:  # the oracle is only the checker; its absence must not stop the measurement
        parity = f"unchecked ({type(e).__name__})"
    if parity.startswith("MISMATCH") and not os.environ.get("SDRD_BENCH_ALLOW_MISMATCH"):
        raise SystemExit("bench: GPU result differs from the oracle; refusing to report a number")

    for _ in range(max(args.warmup - 1, 0)):
        step()
    stream.synchronize()

    launches0 = rx.launches
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ka = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kb = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0.record(stream)
    for i in range(args.steps):
        step(ka[i], kb[i])
    ev1.record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
    launches = rx.launches - launches0
    ms = ev0.elapsed_time(ev1)
    k1_ms = sum(a.elapsed_time(b) for a, b in zip(ka, kb)) / args.steps
    t = torch.tensor([ms, k1_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, k1_ms = float(t[0]), float(t[1])
    value = world * S * n_in * args.steps / (ms * 1e-3) / 1e6

    # ---- end to end through the host-pointer C ABI --------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_in = torch.empty((S, n_in, 2), dtype=torch.int16).pin_memory()
        for s0 in range(S):  # per stream: no second host copy of the whole shard
            host_in[s0].copy_(dev_in[s0].reshape(n_in, 2))
        torch.cuda.synchronize()
        bpf = 128 + N_FEC
        host_out = torch.empty((S, args.frames + 1, bpf, 512), dtype=torch.uint8).pin_memory()
        nfr = C.c_size_t(0)

        def e2e_step():
            lib.check(lib.sdrd_rx_process(rx._h, host_in.data_ptr(), n_in, n_in, host_out.data_ptr(), args.frames + 1,
                                          C.byref(nfr), None))

        e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {"value": round(world * S * n_in * args.e2e_steps / dt / 1e6, 1), "unit": UNIT,
               "h2d_bytes_per_step": S * n_in * 4, "d2h_bytes_per_step": S * int(nfr.value) * bpf * 512,
               "steps": args.e2e_steps, "api": "sdrd_rx_process (host pointers, pinned)"}
        # the ceiling of any end-to-end figure: a bare host -> device copy of the same page-locked bytes, every rank at
        # once (one PCIe link per GPU, one host memory system for all of them)
        flat_h = host_in.view(torch.uint8).reshape(-1)
        flat_d = dev_all.view(torch.uint8)[: flat_h.numel()] if S == 1 else torch.empty(flat_h.numel(), dtype=torch.uint8, device="cuda")
        cs = torch.cuda.Stream()
        with torch.cuda.stream(cs):
            flat_d.copy_(flat_h, non_blocking=True)
            cs.synchronize()
            if world > 1:
                dist.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(cs)
            for _ in range(3):
                flat_d.copy_(flat_h, non_blocking=True)
            c1.record(cs)
            cs.synchronize()
        tc = torch.tensor([c0.elapsed_time(c1) / 3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        ceil_v = world * S * n_in / (float(tc[0]) * 1e-3) / 1e6
        e2e["host_link_ceiling"] = {"value": round(ceil_v, 1), "unit": UNIT, "GB_per_s_all_gpus": round(world * flat_h.numel() / float(tc[0]) / 1e6, 1),
                                    "what": "bare cudaMemcpyAsync of the same page-locked input bytes, all ranks at once"}
        e2e["frac_of_host_link_ceiling"] = round(e2e["value"] / ceil_v, 3)
        if numa is not None:
            e2e["numa"] = numa
        del flat_d
        if world == 1 and args.config == 2:
            e2e["small_block"] = small_block_leg(lib, capi)

    # the only exchange of the job: one 8-byte digest per stream (sdrdaemon_b200.multi)
    digests = None
    if world > 1:
        from sdrdaemon_b200 import multi

        first, count = multi.stream_range(world * S, world, rank)
        assert count == S
        dg_ptr, frame_pitch = sink.dev_datagrams()
        fb = (128 + N_FEC) * 512  # first frame of every stream of this rank (stream pitch = frame_pitch frames)
        allv = torch.as_tensor(DevView(dg_ptr, ((S - 1) * frame_pitch + 1) * fb), device="cuda")
        head = torch.as_strided(allv, (S, fb), (frame_pitch * fb, 1)).cpu().numpy()
        digests = multi.gather_digests(multi.datagram_digest(head), world * S, world, rank, device=torch.device("cuda", local))

    # drop-in mode: one process holds every stream and scatters contiguous ranges to the ranks over
    # NVLink (one NCCL group of sends); reported next to the headline, not part of `value` (there each rank
    # generates its own streams on the device)
    scatter = None
    dropin = None
    if world > 1:
        x_all = torch.zeros((world * S if rank == 0 else 0, n_in * 2), dtype=torch.int16, device="cuda")
        if rank == 0:  # distinct streams: the first ones are this rank's own (already in its input buffer)
            gg = torch.Generator(device="cuda")
            gg.manual_seed(0x5D12DAE0)
            for s0 in range(world * S):
                x_all[s0].copy_(torch.randint(-32768, 32768, (n_in * 2,), dtype=torch.int16, device="cuda", generator=gg))
        multi.scatter_streams(x_all, world * S, world, rank)  # warm-up: NCCL P2P channels
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        sa.record()
        for _ in range(3):
            mine = multi.scatter_streams(x_all, world * S, world, rank)
        sb.record()
        torch.cuda.synchronize()
        ts = torch.tensor([sa.elapsed_time(sb) / 3], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sc_bytes = (world - 1) * S * n_in * 4
        scatter = {"ms": round(float(ts[0]), 3), "bytes_from_rank0": sc_bytes, "GB_per_s": round(sc_bytes / float(ts[0]) / 1e6, 1),
                   "api": "sdrdaemon_b200.multi.scatter_streams (NCCL send/recv, device buffers), alone"}
        del mine
        # ---- drop-in mode, the scatter INSIDE the timed region and overlapped with the work: rank 0 holds every
        # stream and sends each rank's range in K pieces (piece 0 to everybody, then piece 1, ...); a rank decimates /
        # frames / encodes piece k while piece k + 1 is on the wire.  One Rx handle per piece (S / K streams each).
        K = 4 if S % 4 == 0 and S >= 4 else 1
        rx.close()
        Sk = S // K
        rxk = [capi.Rx(M_LOG2, n_streams=Sk, max_in=n_in, n_fec=N_FEC, sample_rate=out_rate) for _ in range(K)]
        views = []
        for h in rxk:
            ip, istr = h.dev_input()
            va = torch.as_tensor(DevView(ip, ((Sk - 1) * istr + n_in) * 4), device="cuda").view(torch.int16)
            views.append(torch.as_strided(va, (Sk, n_in * 2), (istr * 2, 1)))
        cur = torch.cuda.current_stream()

        def dropin_step():
            pieces = multi.scatter_streams_sliced(x_all, world * S, world, rank, K)
            for k, (works, t) in enumerate(pieces):
                if rank != 0:
                    for w in works:
                        w.wait()  # this stream waits for piece k only
                views[k].copy_(t, non_blocking=True)
                rxk[k].process_dev(n_in, cur.cuda_stream)
            if rank == 0:
                for works, _ in pieces:
                    for w in works:
                        w.wait()

        dropin_step()
        torch.cuda.synchronize()
        dist.barrier()
        da, db = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        da.record()
        n_it = 3
        for _ in range(n_it):
            dropin_step()
        db.record()
        torch.cuda.synchronize()
        td = torch.tensor([da.elapsed_time(db) / n_it], device="cuda", dtype=torch.float64)
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
        td = float(td[0])
        dev_ms = ms / args.steps
        dropin = {"value": round(world * S * n_in / (td * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(td, 3), "pieces": K,
                  "scatter_bytes_from_rank0": sc_bytes, "scatter_GB_per_s_inside": round(sc_bytes / td / 1e6, 1),
                  "device_resident_ms_per_step": round(dev_ms, 3), "scatter_alone_ms": scatter["ms"],
                  "limiter": ("the scatter: rank 0 sends (N-1)/N of all samples over its own NVLink ports (900 GB/s per direction nominal), "
                              "the shards' work hides behind it" if scatter["ms"] > dev_ms else "the shards' work: the scatter hides behind it"),
                  "what": "rank 0 holds all streams in HBM; NCCL send/recv of every other rank's range in pieces, each rank works on "
                          "piece k while piece k+1 arrives; timed with the scatter inside, max over ranks"}
        # ---- the same drop-in step with the scatter done by rank 0's COPY ENGINES (sdrd_dec_ipc_export / sdrd_ipc_copy_rows:
        # CUDA IPC + cudaMemcpy2DAsync device-to-device over NVLink) straight into each rank's decimator input buffer.
        # No SM is involved, so the pushes run under the kernels of every rank, and the receivers need no staging copy.
        dropin_dma = None
        NCE = 6  # copy streams per peer: one copy engine moves ~150 GB/s over NVLink, several run side by side
        try:
            mine_h = []
            for h in rxk:
                hb = (C.c_ubyte * 64)()
                off, strd = C.c_size_t(0), C.c_size_t(0)
                lib.check(lib.sdrd_dec_ipc_export(h.dec_handle, hb, C.byref(off), C.byref(strd)))
                mine_h.append((bytes(hb), off.value, strd.value))
            all_h = [None] * world
            dist.all_gather_object(all_h, mine_h)
            gloo = dist.new_group(backend="gloo")
            dev = torch.device("cuda", local)
            peer = {}
            cs = {}
            ev = {}
            if rank == 0:
                for r in range(1, world):
                    cs[r] = [torch.cuda.Stream() for _ in range(NCE)]
                    for k in range(K):
                        pp = C.c_void_p()
                        lib.check(lib.sdrd_ipc_open((C.c_ubyte * 64).from_buffer_copy(all_h[r][k][0]), C.byref(pp)))
                        peer[(r, k)] = (pp.value, all_h[r][k][1], all_h[r][k][2])
                        e = torch.cuda.Event(enable_timing=False, interprocess=True)
                        e.record(cs[r][0])
                        ev[(r, k)] = e
                ev_h = [{key: e.ipc_handle() for key, e in ev.items()}]
            else:
                ev_h = [None]
            dist.broadcast_object_list(ev_h, src=0)
            if rank != 0:
                ev = {k: torch.cuda.Event.from_ipc_handle(dev, ev_h[0][(rank, k)]) for k in range(K)}
            row_bytes = n_in * 4

            def dma_step():
                if rank == 0:
                    for k in range(K):
                        for r in range(1, world):
                            a, c = multi.stream_range(world * S, world, r)
                            f0, n = multi.slice_range(c, K, k)
                            pp, off, strd = peer[(r, k)]
                            for q in range(NCE):  # rows of the piece spread over the peer's copy streams
                                q0, qn = multi.slice_range(n, NCE, q)
                                if qn:
                                    lib.check(lib.sdrd_ipc_copy_rows(pp + off + q0 * strd * 4, strd * 4,
                                                                     x_all.data_ptr() + (a + f0 + q0) * row_bytes, row_bytes, row_bytes, qn,
                                                                     C.c_void_p(cs[r][q].cuda_stream)))
                                if q:
                                    cs[r][0].wait_stream(cs[r][q])
                            ev[(r, k)].record(cs[r][0])
                dist.barrier(group=gloo)  # this step's copies and event records are in rank 0's streams
                for k in range(K):
                    if rank == 0:
                        f0, n = multi.slice_range(S, K, k)
                        views[k].copy_(x_all[f0:f0 + n], non_blocking=True)
                    else:
                        cur.wait_event(ev[k])
                    rxk[k].process_dev(n_in, cur.cuda_stream)
                if rank == 0:
                    for r in cs:
                        cur.wait_stream(cs[r][0])
                torch.cuda.synchronize()
                dist.barrier(group=gloo)  # nobody's input buffer is overwritten while it is still being read

            dma_step()
            # what arrived is what rank 0 holds: one checksum per rank
            a0, c0 = multi.stream_range(world * S, world, rank)
            got_sum = sum(int(v.to(torch.int64).sum()) for v in views)
            sums = [None] * world
            dist.all_gather_object(sums, got_sum)
            verified = None
            if rank == 0:
                verified = all(int(x_all[multi.stream_range(world * S, world, r)[0]:multi.stream_range(world * S, world, r)[0] + S].to(torch.int64).sum()) == sums[r]
                               for r in range(world))
            qa, qb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            qa.record()
            for _ in range(n_it):
                dma_step()
            qb.record()
            torch.cuda.synchronize()
            tq = torch.tensor([qa.elapsed_time(qb) / n_it], device="cuda", dtype=torch.float64)
            dist.all_reduce(tq, op=dist.ReduceOp.MAX)
            tq = float(tq[0])
            dropin_dma = {"value": round(world * S * n_in / (tq * 1e-3) / 1e6, 1), "unit": UNIT, "ms_per_step": round(tq, 3), "pieces": K,
                          "scatter_GB_per_s_inside": round(sc_bytes / tq / 1e6, 1), "scatter_verified": verified,
                          "what": "as `dropin`, the scatter pushed by rank 0's copy engines through CUDA IPC straight into each rank's "
                                  "decimator input buffer (sdrd_dec_ipc_export + sdrd_ipc_copy_rows); host-side handshakes (gloo) inside the timed region"}
            if rank == 0:
                for (r, k), (pp, _, _) in peer.items():
                    lib.sdrd_ipc_close(C.c_void_p(pp))
        except Exception as e:  # the figure is an extra: never lose the line over it
            dropin_dma = {"value": None, "note": f"not measured ({type(e).__name__}: {e})"[:300]}
        if dropin is not None:
            dropin["copy_engine_form"] = dropin_dma
        for h in rxk:
            h.close()
        del x_all, views

    if rank == 0:
        peak, peak_src = measured_peaks()
        k1_bytes = S * n_in * k1_bytes_per_sample(M_LOG2)
        achieved = k1_bytes / (k1_ms * 1e-3) / 1e9
        step_bytes = S * n_in * algorithmic_bytes_per_sample(M_LOG2, N_FEC)
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
        if os.path.exists(tp):
            try:
                t = json.load(open(tp))
                # the capture is of one launch of a given shape: only quote it for that shape
                if t.get("log2_decim", 4) == M_LOG2 and t.get("samples_per_launch", 592 * 16129 * 16) == S * n_in:
                    traffic, traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {
                "workload": f"{work['name']}, {args.frames} superframes per stream "
                            f"({S * n_in} samples, {S * n_in * 4 / 1e6:.0f} MB) per step per GPU"
                            + (f"; {world * S} streams in all, the total fixed as GPUs are added" if args.strong else ""),
                "log2_decim": M_LOG2, "n_fec": N_FEC, "streams_per_gpu": S, "frames_per_step": args.frames,
                "l2": "inputs larger than L2 (no flush needed)", "parity": parity,
            },
            "e2e": e2e,
            "gpu_launches": int(launches),
            # one 8-byte digest per stream gathered from all ranks; the line carries the first of each rank and a fold of all
            "stream_digests": ([hex(int(v)) for v in digests[::S]] if digests is not None else None),
            "stream_digests_fold": (hex(int(np.bitwise_xor.reduce(digests))) if digests is not None else None),
            "n_stream_digests": (int(len(digests)) if digests is not None else None),
            "scatter": scatter,
            "dropin": dropin,
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": f"hb::decimate_warp_kernel<{M_LOG2}> (K1)", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "k1_ms_per_launch": round(k1_ms, 4), "algorithmic_bytes_per_launch": int(k1_bytes),
                "whole_step": {"algorithmic_bytes": int(step_bytes), "achieved": round(step_bytes / (ms / args.steps * 1e-3) / 1e9, 1),
                               "frac": round(step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, 4)},
                "note": "K1 is instruction-issue bound (32*(1-2^-M) IMAD + as many IADD3 per input sample on CUDA cores; ceiling ~41 % of the HBM roofline at M=4), see DESIGN.md",
            },
        }
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            line["cpu_baseline"], _ = cpu_leg(10.0, cores)
            one, _ = cpu_leg(3.0, 1)  # the single-thread figure next to the all-core one
            line["cpu_baseline"]["one_core"] = one["value"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _keep_stdout_for_the_json_line():
    """Rank 0's stdout must carry ONE JSON line.  Native libraries write to file descriptor 1 directly (NCCL prints
    "NCCL version ..." there when NCCL_DEBUG is VERSION or higher): point fd 1 at stderr and keep the original
    stdout for Python's own prints, which are the JSON lines only."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _keep_stdout_for_the_json_line()
    main()
