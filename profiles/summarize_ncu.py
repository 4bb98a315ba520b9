#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full --import-source on) into the text summaries kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [kernel-name-substring] > profiles/rN/<name>.txt

Per kernel: the raw-page metrics the roofline discussion in DESIGN.md uses, the executed-instruction
histogram by opcode (source page), shared-memory accesses with bank conflicts, and where the warp
stall samples fall.  Runs on the CPU box (ncu -i needs no GPU)."""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
       "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__warps_eligible.avg.per_cycle_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = page(rep, "raw")
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if want and want not in d.get("Kernel Name", ""):
            continue
        print("== kernel:", d.get("Kernel Name"), " (launch id", d.get("ID"), ")")
        for k in RAW:
            if k in d:
                print(f"{k} = {d[k]}")
        for k in sorted(d):
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and float(d[k] or 0) > 0.005:
                print(f"{k} = {d[k]}")
        print()
    # source page: one section per kernel
    src = page(rep, "source")
    i = 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            kname = src[i][1]
            h = src[i + 1]
            ix = {c: n for n, c in enumerate(h)}
            j = i + 2
            body = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                if len(src[j]) >= len(h):
                    body.append(src[j])
                j += 1
            i = j
            if want and want not in kname:
                continue
            print("== source page:", kname)
            ops = collections.Counter()
            samp = collections.Counter()
            stalls = collections.Counter()
            for r in body:
                t = r[ix["Source"]].split()
                op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else ""))
                ops[op] += float(r[ix["Instructions Executed"]] or 0)
                samp[op] += float(r[ix["# Samples"]] or 0)
                for c in h:
                    if c.startswith("stall_") and "Not Issued" not in c:
                        stalls[c] += float(r[ix[c]] or 0)
            tot = sum(ops.values())
            print(f"warp-instructions executed {tot:.0f}, stall samples {sum(samp.values()):.0f}")
            for op, n in ops.most_common(24):
                print(f"  {op:22s} {n:12.0f} {100 * n / tot:5.1f}%   samples {samp[op]:7.0f}")
            ts = sum(stalls.values()) or 1
            print("  stall samples:", ", ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in stalls.most_common() if v > 0.005 * ts))
            print("  shared-memory accesses with excess wavefronts (instruction, executed, wavefronts, ideal):")
            for r in body:
                w = float(r[ix["L1 Wavefronts Shared"]] or 0)
                wi = float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
                if w > wi * 1.05 and w > 0:
                    print(f"    {r[ix['Source']].strip()[:60]:60s} {float(r[ix['Instructions Executed']] or 0):10.0f} {w:10.0f} {wi:10.0f}")
            print()
        else:
            i += 1


if __name__ == "__main__":
    main()
