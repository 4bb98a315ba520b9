/*
 * hb_decimate.cuh -- K1: fused multi-stage half-band decimator for interleaved int16 I/Q.
 *
 * Replaces, for the centred (fcpos=2) case, Decimators::decimate{2,4,8,16,32,64}_cen
 * (reference sdmnbase/Decimators.cpp:94,173,270,403,595,902) built from
 * IntHalfbandFilterEO1<64>::myDecimate (include/IntHalfbandFilterEO1.h:34-42,100-147) or, with
 * round_add = 1, IntHalfbandFilterDB<64> (include/IntHalfbandFilterDB.h:80-107).  Per stage
 *
 *     y[n] = ( sum_{i<16} H[i] * ( x[2n+1-2i] + x[2n-61+2i] ) + ((x[2n-30] + round_add) << 13) ) >> 13
 *
 * in wrapping int32 arithmetic, no truncation between stages, and after the last stage
 * out = (int16)((y << norm_shift) >> trunk_shift)  (Decimators.cpp:408-409,510-511, SDRDaemon.h:59).
 * With prologue = 1/2 the cascade is fed by the infra/supra-dyne rotate-and-sum-of-four
 * (Decimators.cpp:337-367 and siblings): log2 decimation = M + 2.
 *
 * Design (B200; an instruction-issue / shared-memory bound streaming kernel, no tensor cores):
 * one WARP (a 32-thread CTA) walks one segment of one stream through all M stages.
 *
 *   - a step consumes WC0 = 512 cascade-input samples (2 KB, one TMA bulk copy requested a step ahead) and
 *     every lane owns one FIR task of 32 outputs x one component for the life of the warp:
 *         lanes  0-15  stage 1 (8 tasks per component)      lanes 24-27  stage 3
 *         lanes 16-23  stage 2                              lanes 28-29  stage 4
 *         lanes 30-31  stage 5 on every 2nd step (32 outputs need two chunks) and, on the steps in
 *                      between, stage 6 on every 4th
 *     so 30 of 32 lanes (M = 4) .. 31.5 of 32 (M = 6) do FIR work in every step and there is no
 *     CTA-wide barrier anywhere: the stages are software-pipelined across steps (stage j works on the
 *     chunk unpacked j steps earlier) and a step is  window loads + arithmetic | __syncwarp | stores |
 *     __syncwarp : every read of a stage region precedes every write of the step, so a region is a
 *     single linear buffer [32 history entries | the step's new entries] per (stage, component, parity)
 *     -- no double buffer; a few lanes shift the last 32 entries down after the consumers have read
 *     them.  Warps drift apart freely, which is what keeps the four schedulers of an SM fed (12 resident
 *     warps per SM: 3 per scheduler, <= 168 registers, ~17 KB of shared memory each at M = 4);
 *   - a task of 32 outputs reads a window of 64 odd-phase + 33 even-phase inputs (3 words per output;
 *     a 16-output task needs 4): shared-memory bandwidth (128 B/clk/SM) is the second limit of this
 *     kernel right behind instruction issue, see DESIGN.md.  The window streams through registers in
 *     four groups of 8 outputs (fir32_stream);
 *   - regions are padded 32 -> 36 words and laid out in an order that puts their start residues mod 32
 *     words where the eight lanes of every quarter-warp hit eight different 16-byte bank groups on every
 *     LDS.128 / STS.128, with no padding between regions (wregion_unit_off);
 *   - the last stage leaves int32 results in a small staging area that the next step packs to int16
 *     pairs, one output per lane, coalesced.
 *
 * This file is single-source: nvcc builds the product kernel, tests/emu builds the same code for
 * the host (see sdrd_platform.cuh).
 */
#pragma once
#include "sdrd_platform.cuh"

namespace sdrd {
namespace hb {

/* HBFIRFilterTraits<64>::hbCoeffs as integers (sdmnbase/HBFilterTraits.cpp:210-228, Q14, truncated
 * toward zero), outermost tap first; centre tap is 1 << 13. */
#define SDRD_HB64_TAPS {-7, 11, -20, 32, -49, 71, -101, 140, -190, 256, -345, 469, -656, 978, -1698, 5201}
constexpr int HB_SHIFT = 13; /* hbShift - 1, EO1.h:145 */

struct Params {
    const uint32_t* in;    /* stream s, raw sample i (may be negative: history): in[s * in_stride + i] */
    long long in_stride;   /* words */
    uint32_t* out;         /* out[s * out_stride + n] */
    long long out_stride;  /* words */
    long long n_out;       /* outputs per stream */
    long long ev_stream;   /* pack events (wfin_n(M) outputs each) per stream: ceil(n_out / wfin_n(M)) */
    long long ev_total;    /* n_streams * ev_stream: the global event axis, stream after stream */
    long long ev_warp;     /* events per warp: warp w takes global events [w * ev_warp, (w + 1) * ev_warp) */
    int warm_chunks;       /* informational: wwarm_chunks(M), the chunks processed ahead of every segment */
    int round_add;         /* 0: EO1, 1: DB */
    int norm_shift, trunk_shift;
    int prologue;          /* 0: centred; 1: infradyne /4; 2: supradyne /4 in front of the cascade */
    long long origin;      /* cascade-input samples consumed since reset (only consulted when round_add) */
    uint32_t steer_zero, steer_one, steer_k32, steer_k256, steer_k8192; /* 0, 1, 32, 256, 8192: see Steer */
};

/* Pipe steering.  The FIR body is issue-bound: per output and component 16 pre-adds + 16
 * multiply-accumulates, on two pipes that each take one warp instruction every 2 clocks.  Left alone,
 * ptxas turns about half of the pre-adds into IMAD.IADD, which piles them onto the FMA pipe next to the
 * IMADs.  `Steer` carries run-time constants the compiler cannot fold:
 *   zero   added as the THIRD operand of a pre-add  -> IADD3 with three sources, ALU pipe only;
 *   one    multiplier of an add written as IMAD      -> FMA pipe (SDRD_HB_FMA_ADD_TAPS taps per output);
 *   k32, k256, k8192  multipliers of the power-of-two taps / centre tap -> IMAD instead of a shift-add
 *          (SDRD_HB_POW2_IMAD; off: measured on the warp-private kernel, FMA is the busier pipe there).
 * Measured (config 2, K1 ms): FMA_ADD_TAPS 0/1/2/3 = 0.422/0.416/0.425/0.430 with POW2_IMAD = 0;
 * POW2_IMAD 7 -> 0.422. */
struct Steer {
    uint32_t zero, one, k32, k256, k8192;
};

#ifndef SDRD_HB_FMA_ADD_TAPS
#define SDRD_HB_FMA_ADD_TAPS 1 /* taps (outermost first) whose pre-add runs on the FMA pipe */
#endif

SDRD_DEVICE int s16lo(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
SDRD_DEVICE int s16hi(uint32_t v) { return ((int)v) >> 16; }

/* Infra/supra-dyne divide-by-4 prologue (Decimators.cpp:337-367 for M=3 and the same pattern in
 * decimate16/32/64_{inf,sup}): four consecutive raw samples -> one cascade input. */
SDRD_DEVICE int2 rot4(uint4 v, int prologue)
{
    int r0 = s16lo(v.x), i0 = s16hi(v.x), r1 = s16lo(v.y), i1 = s16hi(v.y);
    int r2 = s16lo(v.z), i2 = s16hi(v.z), r3 = s16lo(v.w), i3 = s16hi(v.w);
    if (prologue == 1) return make_int2(r0 - i1 + i3 - r2, i0 - i2 + r1 - r3);
    return make_int2(i0 - r1 - i2 + r3, -r0 - i1 + r2 + i3);
}

constexpr int WC0 = 512; /* cascade-input samples per step */
constexpr int BLK = 36;  /* 32 entries + 4 words of padding */

/* logical entry -> word offset inside a region */
SDRD_HD constexpr int wphys(int k) { return k + 4 * (k >> 5); }
/* logical entries of the region holding stage m's output (m = 0: the unpacked input), one parity, one
 * component: 32 history + what one consumer step reads as new (m >= 4: two producer steps' worth) */
SDRD_HD constexpr int wregion_entries(int m) { return m <= 3 ? 32 + (256 >> m) : 64; }
SDRD_HD constexpr int wregion_words(int m) { return wregion_entries(m) / 32 * BLK; }
/* Placement of the regions.  All E regions, then all O regions, each set in the order
 *   m1.I  m2.Q  m0.I  m0.Q  m3.I  m1.Q  m3.Q  m2.I  (m4.I  m4.Q  m5.I  m5.Q)
 * which makes the start residues modulo 32 words (8 bank groups of 16 bytes) come out right WITHOUT any
 * padding between the first eight: the mixed quarter-warps 16-23 (stage 2: tasks I0-3, Q0-3) and 24-31
 * (stages 3..6) then hit eight different bank groups both when they load their windows and when they
 * store their results (m1: Q - I = 4 groups; m2: Q - I = 2; lanes 24-31 read groups {3,4},{5,6},2,1,0,7).
 * Found by exhaustive search over the 8! orders; offsets in units of 16 bytes. */
SDRD_HD constexpr int wregion_unit_off(int m, int comp)
{
    return m == 0 ? (comp ? 153 : 72) : m == 1 ? (comp ? 252 : 0) : m == 2 ? (comp ? 45 : 315) : m == 3 ? (comp ? 297 : 234)
         : m == 4 ? (comp ? 367 : 344) : (comp ? 415 : 392);
}
SDRD_HD constexpr int wset_units(int M) { return M <= 4 ? 342 : M == 5 ? 385 : 433; } /* one parity's regions */
/* word offset of region (m, parity, comp) from the start of the plane area, for an M-stage cascade */
SDRD_HD constexpr int wplane_off(int M, int m, int parity, int comp) { return 4 * (parity * wset_units(M) + wregion_unit_off(m, comp)); }
SDRD_HD constexpr int wplanes_words(int M) { return 8 * wset_units(M); }
SDRD_HD constexpr int wfin_n(int M) { return (WC0 >> M) > 32 ? (WC0 >> M) : 32; } /* outputs per pack event */
SDRD_HD constexpr int wfin_stride(int M) { return wfin_n(M) + 4; }                  /* words between the I and Q results */
SDRD_HD constexpr int wmacro(int M) { return M <= 4 ? 1 : 1 << (M - 4); }           /* steps per pack event */
/* steps between unpacking the first chunk of a pack event and packing its outputs */
SDRD_HD constexpr int wdelay(int M) { return M <= 4 ? M + 1 : (M == 5 ? 7 : 10); }
SDRD_HD constexpr int wraw_words(int PRO) { return PRO ? 4 * WC0 : WC0; }
/* raw chunk | mbarrier | regions of stages 0..M-1 | 8 words | last stage's results */
SDRD_HD constexpr size_t wsmem_bytes(int M, int PRO)
{
    return (size_t)wraw_words(PRO) * 4 + 128 + (size_t)(wplanes_words(M) + 8) * 4 + (size_t)2 * wfin_stride(M) * 4;
}
/* warm-up chunks in front of a segment: >= 61 * (2^M - 1) samples, whole pack events */
SDRD_HD constexpr int wwarm_chunks(int M)
{
    return ((61 * ((1 << M) - 1) + WC0 - 1) / WC0 + wmacro(M) - 1) / wmacro(M) * wmacro(M);
}

#ifndef SDRD_HB_POW2_IMAD
#define SDRD_HB_POW2_IMAD 0 /* bit 0/1/2: tap 32 / tap 256 / centre tap multiply as IMAD with a run-time factor (FMA pipe) instead of a shift-add (ALU pipe) */
#endif
#ifndef SDRD_K1_WARPS_PER_SM
#define SDRD_K1_WARPS_PER_SM 12 /* resident warps per SM the register budget is set for */
#endif
/* 32 consecutive outputs n0 .. n0+31 of one stage and component, from the regions the task's pointers
 * address: w[j] = O[n0 - 32 + j], e[j] = E[n0 - 16 + j].  Four groups of 8 outputs; group g needs
 * w[1 + 8g .. 39 + 8g] and e[1 + 8g .. 8 + 8g], so the operands of group g + 1 (2 + 2 vector loads) are
 * requested while group g is being computed and at most ~50 window registers are live.  Inside a group
 * the order is explicit: tap by tap, the multiply-accumulate of tap t for output j followed by the
 * pre-add of tap t + 1 for the same output, so that FMA- and ALU-pipe instructions alternate and every
 * IMAD reads operands produced ~16 instructions earlier. */
SDRD_DEVICE void fir32_ld4(const int* p, uint32_t* d)
{
    const int4 v = *reinterpret_cast<const int4*>(p);
    d[0] = (uint32_t)v.x; d[1] = (uint32_t)v.y; d[2] = (uint32_t)v.z; d[3] = (uint32_t)v.w;
}
/* 16-byte unit k of the window: w[4k .. 4k+3] resp. e[4k .. 4k+3] (e[0] is entry 16 of the first block) */
SDRD_DEVICE const int* fir32_w_unit(const int* srcO, int k) { return srcO + (k < 8 ? 4 * k : BLK + 4 * (k - 8)); }
SDRD_DEVICE const int* fir32_e_unit(const int* srcE, int k) { return srcE + (k < 4 ? 16 + 4 * k : BLK + 4 * (k - 4)); }

template <int DB>
SDRD_DEVICE void fir32_stream(const int* SDRD_RESTRICT srcE, const int* SDRD_RESTRICT srcO, const Steer st, long long a0, int (&y)[32])
{
    constexpr int H[16] = SDRD_HB64_TAPS;
    constexpr uint32_t acc0 = (uint32_t)DB << HB_SHIFT;
    uint32_t w[64], e[36], tmp[8], acc[8];
#pragma unroll
    for (int k = 0; k < 10; k++) fir32_ld4(fir32_w_unit(srcO, k), &w[4 * k]);
#pragma unroll
    for (int k = 0; k < 3; k++) fir32_ld4(fir32_e_unit(srcE, k), &e[4 * k]);
#pragma unroll
    for (int j = 0; j < 8; j++) tmp[j] = SDRD_HB_FMA_ADD_TAPS > 0 ? mad_lo(w[32 + j], st.one, w[1 + j]) : add3(w[32 + j], w[1 + j], st.zero);
#pragma unroll
    for (int g = 0; g < 4; g++) {
        if (g < 3) { /* what group g + 1 adds to the window */
            fir32_ld4(fir32_w_unit(srcO, 10 + 2 * g), &w[40 + 8 * g]);
            fir32_ld4(fir32_w_unit(srcO, 11 + 2 * g), &w[44 + 8 * g]);
            fir32_ld4(fir32_e_unit(srcE, 3 + 2 * g), &e[12 + 8 * g]);
            if (g < 2) fir32_ld4(fir32_e_unit(srcE, 4 + 2 * g), &e[16 + 8 * g]);
            else e[32] = (uint32_t)srcE[BLK + 16];
        }
#pragma unroll
        for (int t = 0; t < 16; t++) {
            const uint32_t h = (H[t] == 32 && (SDRD_HB_POW2_IMAD & 1)) ? st.k32 : ((H[t] == 256 && (SDRD_HB_POW2_IMAD & 2)) ? st.k256 : (uint32_t)H[t]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[j] = mad_lo(tmp[j], h, t == 0 ? acc0 : acc[j]);
                /* pre-add of the next tap (or of tap 0 of the next group) */
                const int gn = t < 15 ? g : g + 1, tn = t < 15 ? t + 1 : 0;
                if (gn < 4) {
                    const uint32_t a = w[8 * gn + 32 + j - tn], b = w[8 * gn + 1 + j + tn];
                    tmp[j] = tn < SDRD_HB_FMA_ADD_TAPS ? mad_lo(a, st.one, b) : add3(a, b, st.zero);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
            y[8 * g + j] = asr32((SDRD_HB_POW2_IMAD & 4) ? mad_lo(e[8 * g + 1 + j], st.k8192, acc[j]) : (e[8 * g + 1 + j] << HB_SHIFT) + acc[j], HB_SHIFT);
        if (DB && a0 + 8 * g < 0) {
            /* DB: the reference's stages start from all-zero state, but a DB stage maps zero input to 1;
             * outputs that lie before the stream origin must read as 0. */
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (a0 + 8 * g + j < 0) y[8 * g + j] = 0;
        }
    }
}

template <int M, int DB, int PRO>
SDRD_KERNEL(32, SDRD_K1_WARPS_PER_SM) decimate_warp_kernel(Params p)
{
    static_assert(M >= 1 && M <= 6, "1..6 half-band stages");
    constexpr int FN = wfin_n(M);
    constexpr int FS = wfin_stride(M);
    constexpr int MACRO = wmacro(M);
    constexpr int DELAY = wdelay(M);
    constexpr int OPL = FN / 32; /* outputs a lane packs per event */
    constexpr int RAWW = wraw_words(PRO);
    SDRD_DYN_SMEM(smem);
    const int lane = (int)threadIdx.x;
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem);
    mbar_t* bars = reinterpret_cast<mbar_t*>(smem + (size_t)RAWW * 4);
    int* planes = reinterpret_cast<int*>(smem + (size_t)RAWW * 4 + 128);
    int* fin = planes + wplanes_words(M) + 8;
    constexpr int warm_macro = wwarm_chunks(M) / MACRO;
    constexpr uint32_t chunk_bytes = (uint32_t)RAWW * 4u;
    const Steer steer = {p.steer_zero, p.steer_one, p.steer_k32, p.steer_k256, p.steer_k8192};

    /* ---- this lane's FIR task: source stage tm (0 = unpacked input), component, task index ---- */
    int tm, tcomp, ti;
    if (lane < 16)      { tm = 0; tcomp = lane >> 3;       ti = lane & 7; }
    else if (lane < 24) { tm = 1; tcomp = (lane >> 2) & 1; ti = lane & 3; }
    else if (lane < 28) { tm = 2; tcomp = (lane >> 1) & 1; ti = lane & 1; }
    else if (lane < 30) { tm = 3; tcomp = lane & 1;        ti = 0; }
    else                { tm = 4; tcomp = lane & 1;        ti = 0; }
    const bool has_task = tm < M;
    const bool sub_rate = lane >= 30; /* only meaningful when M >= 5 */
    if (!has_task) tm = 0;
    const int* srcE = planes + wplane_off(M, tm, 0, tcomp) + BLK * ti;
    const int* srcO = planes + wplane_off(M, tm, 1, tcomp) + BLK * ti;
    int* dstE;
    int* dstO;
    if (tm + 1 < M) {
        dstE = planes + wplane_off(M, tm + 1, 0, tcomp) + wphys(32 + 16 * ti);
        dstO = planes + wplane_off(M, tm + 1, 1, tcomp) + wphys(32 + 16 * ti);
    } else {
        dstE = fin + tcomp * FS + 16 * ti;
        dstO = dstE + FN / 2;
    }
    /* M = 6: lanes 30/31 alternate between stage 5 (regions 4 -> 5) and stage 6 (regions 5 -> fin) */
    const int* srcE_b = planes + wplane_off(M, M == 6 ? 5 : 0, 0, tcomp);
    const int* srcO_b = planes + wplane_off(M, M == 6 ? 5 : 0, 1, tcomp);
    int* dstE_b = fin + tcomp * FS;
    int* dstO_b = dstE_b + FN / 2;
    /* producers of regions 4 and 5 fill them in two halves of 16 entries */
    const bool dst_halves_4 = M >= 5 && (lane == 28 || lane == 29);
    const bool dst_halves_5 = M == 6 && sub_rate;

    /* ---- tail shift: 16-byte unit t of the 24 * M (per region pair: O 8 units, E the upper 4) ---- */
    constexpr int TSLOTS = (24 * M + 31) / 32;
    const int* tl_src[TSLOTS];
    int tl_back[TSLOTS]; /* words from source to destination */
    int tl_m[TSLOTS];
#pragma unroll
    for (int k = 0; k < TSLOTS; k++) {
        const int t = lane + 32 * k;
        const bool on = t < 24 * M;
        const int m = on ? t / 24 : 0, r = t % 24;
        const int comp = r / 12, rr = r % 12;
        const int parity = rr < 8 ? 1 : 0;
        const int unit = parity ? rr : 4 + (rr - 8);
        tl_back[k] = (wregion_entries(m) - 32) / 32 * BLK;
        tl_src[k] = planes + wplane_off(M, m, parity, comp) + 4 * unit + tl_back[k];
        tl_m[k] = on ? m : -1;
    }

    /* ---- unpack (no prologue): lane q reads 16-byte unit 32 k + q of the raw chunk, k = 0..3: samples
     *      4 (32 k + q) .. + 3 -> entries 64 k + 2 q, + 1 of the four stage-0 regions.
     *      (/4 prologue: lane q turns raw unit 32 k + q, k = 0..15, into cascade input 32 k + q) ---- */
    int* const up = planes + (PRO ? wphys(32 + (lane >> 1)) : wphys(32 + 2 * lane));
    constexpr int UP_EI = wplane_off(M, 0, 0, 0), UP_EQ = wplane_off(M, 0, 0, 1), UP_OI = wplane_off(M, 0, 1, 0), UP_OQ = wplane_off(M, 0, 1, 1);

    /* ---- pack: lane l packs outputs OPL*l .. OPL*l+OPL-1 of an event; output r of a component sits at
     *      fin[r even ? r/2 : FN/2 + r/2] ---- */
    const int* const pk = fin + (OPL == 1 ? ((lane & 1) ? FN / 2 + (lane >> 1) : (lane >> 1)) : lane * (OPL / 2));

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_fence_init();
    }
    SDRD_SYNCWARP();

    /* The warp's share of the global event axis (all streams laid end to end, so that the resident warps of
     * the device get equally long shares whatever the number of streams).  A share that runs over the end of
     * a stream continues at the start of the next one as a new piece with its own filter warm-up. */
    long long g = (long long)blockIdx.x * p.ev_warp;
    long long g_end = g + p.ev_warp;
    if (g_end > p.ev_total) g_end = p.ev_total;
    uint32_t tma_n = 0; /* chunks requested before this piece (mbarrier phase) */
    while (g < g_end) {
        const int s = (int)(g / p.ev_stream);
        const long long e0 = g - (long long)s * p.ev_stream;
        long long e1 = e0 + (g_end - g);
        if (e1 > p.ev_stream) e1 = p.ev_stream;
        g += e1 - e0;
        const long long seg_first_out = e0 * FN;
        const int n_macro = warm_macro + (int)(e1 - e0);
        const int NC = n_macro * MACRO; /* chunks to unpack */
        const int u_last = (n_macro - 1) * MACRO + DELAY;
        const long long first_in = (seg_first_out << M) - (long long)wwarm_chunks(M) * WC0;
        const uint32_t* src = p.in + (long long)s * p.in_stride + first_in * (PRO ? 4 : 1);
        uint32_t* dst = p.out + (long long)s * p.out_stride + seg_first_out - (long long)warm_macro * FN;
        const long long seg_room = p.n_out - seg_first_out + (long long)warm_macro * FN; /* valid: index < seg_room */
        const long long abs0 = p.origin + first_in;
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[0], chunk_bytes);
            tma_load_1d(raw, src, chunk_bytes, &bars[0]);
        }

        for (int u = 0; u <= u_last; u++) {
            /* ================= phase A: every read of the stage regions, and the arithmetic ================= */
            bool task_on = has_task;
            const int* sE = srcE;
            const int* sO = srcO;
            int* dE = dstE;
            int* dO = dstO;
            int c_first = u - 1 - tm; /* first chunk this task's outputs come from (DB origin test) */
            int t_stage = tm + 1;
            if (M >= 5) {
                if (sub_rate) {
                    const bool run_a = (u & 1) == 0;           /* stage 5 */
                    const bool run_b = M == 6 && (u & 3) == 1; /* stage 6 */
                    task_on = run_a || run_b;
                    if (run_b) { sE = srcE_b; sO = srcO_b; dE = dstE_b; dO = dstO_b; c_first = u - 9; t_stage = 6; }
                    else { c_first = u - 6; }
                    if (dst_halves_5 && run_a && (((u >> 1) & 1) == 0)) { dE += 16; dO += 16; } /* pair (u-6)/2 odd */
                }
                if (dst_halves_4 && (u & 1)) { dE += 16; dO += 16; } /* chunk u-4 odd */
            }
            int4 tl_v[TSLOTS];
            bool tl_on[TSLOTS];
#pragma unroll
            for (int k = 0; k < TSLOTS; k++) {
                tl_on[k] = tl_m[k] >= 0 && (M <= 4 || tl_m[k] <= 3 || (tl_m[k] == 4 ? (u & 1) == 0 : (u & 3) == 1));
                tl_v[k] = make_int4(0, 0, 0, 0);
                if (tl_on[k]) tl_v[k] = *reinterpret_cast<const int4*>(tl_src[k]);
            }

            const int ev = u - DELAY; /* pack event e = ev / MACRO when ev is a non-negative multiple of MACRO */
            const bool pack_now = ev >= warm_macro * MACRO && (MACRO == 1 || (ev & (MACRO - 1)) == 0);
            int pki[OPL], pkq[OPL];
#pragma unroll
            for (int j = 0; j < OPL; j++) pki[j] = pkq[j] = 0;
            if (pack_now) {
                if (OPL == 1) {
                    pki[0] = pk[0];
                    pkq[0] = pk[FS];
                } else {
#pragma unroll
                    for (int j = 0; j < OPL; j++) { /* output OPL*l + j: even -> entry (OPL*l + j)/2, odd -> FN/2 + .. */
                        const int idx = (j & 1) ? FN / 2 + (j >> 1) : (j >> 1);
                        pki[j] = pk[idx];
                        pkq[j] = pk[FS + idx];
                    }
                }
            }

            /* the FIR task itself: loads and arithmetic, results stay in registers until the barrier */
            int y[32];
            if (task_on) {
                const long long a0 = DB ? ((abs0 + (long long)c_first * WC0) >> t_stage) + 32 * ti : 0;
                fir32_stream<DB>(sE, sO, steer, a0, y);
            }
            SDRD_SYNCWARP();

            /* ================= phase B: every shared-memory write of the step =========
             * (the raw chunk is written by the TMA only, so its loads may follow the barrier) */
            const bool unpack_now = u < NC;
            uint4 rw[4];
            int2 x[16];
            if (unpack_now) {
                mbar_wait(&bars[0], (tma_n + (uint32_t)u) & 1u);
                const uint4* r4 = reinterpret_cast<const uint4*>(raw);
                if (!PRO) {
#pragma unroll
                    for (int k = 0; k < 4; k++) rw[k] = r4[32 * k + lane];
                } else {
#pragma unroll
                    for (int k = 0; k < 16; k++) x[k] = rot4(r4[32 * k + lane], p.prologue);
                }
            }
#pragma unroll
            for (int k = 0; k < TSLOTS; k++)
                if (tl_on[k]) *reinterpret_cast<int4*>(const_cast<int*>(tl_src[k]) - tl_back[k]) = tl_v[k];

            if (pack_now) {
                /* (y << norm_shift) >> trunk_shift truncated to 16 bits (Decimators.cpp:408-409, SDRDaemon.h:59) */
                uint32_t o[OPL];
#pragma unroll
                for (int j = 0; j < OPL; j++) {
                    const uint32_t a = (uint32_t)asr32((uint32_t)pki[j] << p.norm_shift, p.trunk_shift);
                    const uint32_t b = (uint32_t)asr32((uint32_t)pkq[j] << p.norm_shift, p.trunk_shift);
                    o[j] = (a & 0xFFFFu) | (b << 16);
                }
                const long long n = (long long)(ev / MACRO) * FN + lane * OPL; /* relative to dst */
                if (n + OPL <= seg_room) {
                    if (OPL == 1) {
                        dst[n] = o[0];
                    } else if (OPL == 2) {
                        *reinterpret_cast<uint2*>(dst + n) = make_uint2(o[0], o[OPL - 1]);
                    } else {
#pragma unroll
                        for (int j = 0; j + 3 < OPL; j += 4)
                            *reinterpret_cast<uint4*>(dst + n + j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < OPL; j++)
                        if (n + j < seg_room) dst[n + j] = o[j];
                }
            }

            if (task_on) {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    reinterpret_cast<int4*>(dE)[g] = make_int4(y[8 * g], y[8 * g + 2], y[8 * g + 4], y[8 * g + 6]);
                    reinterpret_cast<int4*>(dO)[g] = make_int4(y[8 * g + 1], y[8 * g + 3], y[8 * g + 5], y[8 * g + 7]);
                }
            }

            if (unpack_now) {
                if (!PRO) {
#pragma unroll
                    for (int k = 0; k < 4; k++) { /* entries 64 k + 2 q, + 1: two 32-entry blocks further per k */
                        int* o = up + 2 * BLK * k;
                        *reinterpret_cast<int2*>(o + UP_EI) = make_int2(s16lo(rw[k].x), s16lo(rw[k].z));
                        *reinterpret_cast<int2*>(o + UP_EQ) = make_int2(s16hi(rw[k].x), s16hi(rw[k].z));
                        *reinterpret_cast<int2*>(o + UP_OI) = make_int2(s16lo(rw[k].y), s16lo(rw[k].w));
                        *reinterpret_cast<int2*>(o + UP_OQ) = make_int2(s16hi(rw[k].y), s16hi(rw[k].w));
                    }
                } else {
                    /* cascade input 32 k + q: parity q & 1, entry 16 k + (q >> 1): half a block further per k */
                    int* oi = up + ((lane & 1) ? UP_OI : UP_EI);
                    int* oq = up + ((lane & 1) ? UP_OQ : UP_EQ);
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        oi[(k >> 1) * BLK + (k & 1) * 16] = x[k].x;
                        oq[(k >> 1) * BLK + (k & 1) * 16] = x[k].y;
                    }
                }
            }

            SDRD_SYNCWARP();
            /* every lane has read chunk u out of the raw buffer: request chunk u + 1 into it; it has the whole
             * arithmetic phase of the next step to arrive */
            if (lane == 0 && u + 1 < NC) {
                mbar_arrive_expect_tx(&bars[0], chunk_bytes);
                tma_load_1d(raw, src + (size_t)(u + 1) * RAWW, chunk_bytes, &bars[0]);
            }
        }
        tma_n += (uint32_t)NC;
    } /* pieces */
}

/* ------------------------------------------------------------------------------------------
 * Explicit per-stage state: the reference keeps six persistent stage objects m_decimator2 ..
 * m_decimator64 (include/Decimators.h:57-62) and cascade position k always runs object k
 * (Decimators.cpp:189-201, 290-322, 356-359 ...).  While the configuration does not change, the state of
 * every active stage is a function of the raw input history and decimate_warp_kernel re-derives it
 * from there.  After Downsampler::configure (Downsampler.cpp:32-67) that is no longer true: a stage the
 * new cascade uses continues from whatever it saw last -- under the previous configuration, possibly long
 * ago, or never (zeros).  stateful_kernel runs the cascade from explicit stage states
 *     state[stream][stage 0..5][component I/Q][64]  =  the stage's last 64 inputs, oldest first
 * (what m_even/m_odd of IntHalfbandFilterEO1.h:68-69 resp. m_samples of IntHalfbandFilterDB.h:73 hold)
 * and writes them back.  The library uses it for the first SDRD_DEC_HEAD raw samples after a reconfiguration
 * (from then on every value the cascade can still reach was computed from samples of the new run, and the
 * warp kernel takes over) and, with out == nullptr, to turn the raw history into explicit states when a
 * configuration is left.  One CTA per stream, chunks of SCH cascade inputs through shared memory; a few
 * thousand samples per stream and reconfiguration, so simplicity counts here, not speed.
 * ------------------------------------------------------------------------------------------ */
constexpr int SNT = 256;    /* threads per CTA */
constexpr int SCH = 1024;   /* cascade-input samples per chunk */
constexpr int SST = 64;     /* state entries per stage and component */
constexpr int STATE_WORDS = 6 * 2 * SST; /* per stream */
SDRD_HD constexpr int sbuf_off(int k) { return k == 0 ? 0 : sbuf_off(k - 1) + 2 * (SST + (SCH >> (k - 1))); }
SDRD_HD constexpr size_t stateful_smem_bytes() { return (size_t)sbuf_off(6) * 4; }

struct StateParams {
    const uint32_t* in;    /* stream s, raw sample i: in[s * in_stride + i] */
    long long in_stride;
    uint32_t* out;         /* out[s * out_stride + n], or nullptr: only the states are wanted */
    long long out_stride;
    int* state;            /* [n_streams][STATE_WORDS] */
    long long n_casc;      /* cascade-input samples per stream (raw / 4 with the prologue); a multiple of 2^M */
    int M;                 /* half-band stages 1..6 */
    int round_add, norm_shift, trunk_shift, prologue;
};

SDRD_KERNEL(SNT, 1) stateful_kernel(StateParams p)
{
    constexpr int H[16] = SDRD_HB64_TAPS;
    SDRD_DYN_SMEM(smem);
    int* const buf = reinterpret_cast<int*>(smem);
    const int tid = (int)threadIdx.x;
    const int s = (int)blockIdx.x;
    const int M = p.M;
    const uint32_t* in = p.in + (long long)s * p.in_stride;
    int* state = p.state + (long long)s * STATE_WORDS;
    /* buffer of stage k's input, component c: [SST state | chunk >> k] */
    auto B = [&](int k, int c) -> int* { return buf + sbuf_off(k) + c * (SST + (SCH >> k)); };

    for (int i = tid; i < M * 2 * SST; i += SNT) {
        const int k = i / (2 * SST), c = (i / SST) & 1, j = i % SST;
        B(k, c)[j] = state[(k * 2 + c) * SST + j];
    }
    __syncthreads();
    for (long long c0 = 0; c0 < p.n_casc; c0 += SCH) {
        const int len = (int)(p.n_casc - c0 < SCH ? p.n_casc - c0 : SCH);
        for (int i = tid; i < len; i += SNT) {
            int xi, xq;
            if (p.prologue) {
                const int2 r = rot4(reinterpret_cast<const uint4*>(in)[c0 + i], p.prologue);
                xi = r.x; xq = r.y;
            } else {
                const uint32_t v = in[c0 + i];
                xi = s16lo(v); xq = s16hi(v);
            }
            B(0, 0)[SST + i] = xi;
            B(0, 1)[SST + i] = xq;
        }
        __syncthreads();
        for (int k = 0; k < M; k++) {
            const int nk = len >> (k + 1);
            for (int n = tid; n < nk; n += SNT) {
                int y[2];
                for (int c = 0; c < 2; c++) {
                    const int* x = B(k, c) + SST; /* x[j], j >= -SST */
                    uint32_t acc = 0;
                    for (int t = 0; t < 16; t++) acc += ((uint32_t)x[2 * n + 1 - 2 * t] + (uint32_t)x[2 * n - 61 + 2 * t]) * (uint32_t)H[t];
                    acc += ((uint32_t)x[2 * n - 30] + (uint32_t)p.round_add) << HB_SHIFT;
                    y[c] = asr32(acc, HB_SHIFT);
                }
                if (k + 1 < M) {
                    B(k + 1, 0)[SST + n] = y[0];
                    B(k + 1, 1)[SST + n] = y[1];
                } else if (p.out) {
                    const uint32_t a = (uint32_t)asr32((uint32_t)y[0] << p.norm_shift, p.trunk_shift);
                    const uint32_t b = (uint32_t)asr32((uint32_t)y[1] << p.norm_shift, p.trunk_shift);
                    p.out[(long long)s * p.out_stride + (c0 >> M) + n] = (a & 0xFFFFu) | (b << 16);
                }
            }
            __syncthreads();
        }
        /* the last SST inputs of every stage become its state: read, barrier, write (the ranges overlap when a
         * stage got fewer than SST inputs) */
        int keep[6 * 2 * SST / SNT];
        for (int r = 0; r < 6 * 2 * SST / SNT; r++) {
            const int i = tid + r * SNT;
            const int k = i / (2 * SST), c = (i / SST) & 1, j = i % SST;
            keep[r] = k < M ? B(k, c)[(len >> k) + j] : 0;
        }
        __syncthreads();
        for (int r = 0; r < 6 * 2 * SST / SNT; r++) {
            const int i = tid + r * SNT;
            const int k = i / (2 * SST), c = (i / SST) & 1, j = i % SST;
            if (k < M) B(k, c)[j] = keep[r];
        }
        __syncthreads();
    }
    for (int i = tid; i < M * 2 * SST; i += SNT) {
        const int k = i / (2 * SST), c = (i / SST) & 1, j = i % SST;
        state[(k * 2 + c) * SST + j] = B(k, c)[j];
    }
}

/* ------------------------------------------------------------------------------------------
 * Filter-less routines of the reference: decimate1 (Decimators.cpp:22-35, left-justify sources
 * with fewer than 16 bits), decimate2_inf/sup (:38-91) and decimate4_inf/sup (:127-170).
 * Element-wise, HBM-bound; one thread per output group.
 * ------------------------------------------------------------------------------------------ */
struct PlainParams {
    const uint32_t* in;
    long long in_stride;
    uint32_t* out;
    long long out_stride;
    long long n_units;  /* mode 0: samples; mode 1: quads (4 in -> 2 out); mode 2: quads (4 in -> 1 out) */
    int mode;           /* 0: rescale, 1: decimate2 inf/sup, 2: decimate4 inf/sup */
    int supra;          /* 0: infradyne, 1: supradyne */
    int norm_shift, trunk_shift;
    long long n_zero_tail; /* mode 1: out.resize(len/2) leaves one default sample when len % 4 >= 2 */
};

SDRD_DEVICE uint32_t pack_out(int x, int y, int norm_shift, int trunk_shift)
{
    uint32_t vi = (uint32_t)asr32((uint32_t)x << norm_shift, trunk_shift);
    uint32_t vq = (uint32_t)asr32((uint32_t)y << norm_shift, trunk_shift);
    return (vi & 0xFFFFu) | (vq << 16);
}

SDRD_KERNEL(256, 4) plain_kernel(PlainParams p)
{
    const int s = (int)blockIdx.y;
    const uint32_t* in = p.in + (long long)s * p.in_stride;
    uint32_t* out = p.out + (long long)s * p.out_stride;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < p.n_units; g += step) {
        if (p.mode == 0) {
            uint32_t v = in[g];
            int x = s16lo(v), y = s16hi(v);
            out[g] = (((uint32_t)x << p.norm_shift) & 0xFFFFu) | (((uint32_t)y << p.norm_shift) << 16);
        } else {
            uint4 v = reinterpret_cast<const uint4*>(in)[g];
            int r0 = s16lo(v.x), i0 = s16hi(v.x), r1 = s16lo(v.y), i1 = s16hi(v.y);
            int r2 = s16lo(v.z), i2 = s16hi(v.z), r3 = s16lo(v.w), i3 = s16hi(v.w);
            if (p.mode == 1) {
                int x0, y0, x1, y1;
                if (!p.supra) { x0 = r0 - i1; y0 = i0 + r1; x1 = i3 - r2; y1 = -i2 - r3; }
                else          { x0 = i0 - r1; y0 = -r0 - i1; x1 = r3 - i2; y1 = r2 + i3; }
                out[2 * g] = pack_out(x0, y0, p.norm_shift, p.trunk_shift);
                out[2 * g + 1] = pack_out(x1, y1, p.norm_shift, p.trunk_shift);
            } else {
                int2 r = rot4(v, p.supra ? 2 : 1);
                out[g] = pack_out(r.x, r.y, p.norm_shift, p.trunk_shift);
            }
        }
    }
    if (p.mode == 1 && blockIdx.x == 0 && (long long)threadIdx.x < p.n_zero_tail) out[2 * p.n_units + threadIdx.x] = 0u;
}

} /* namespace hb */
} /* namespace sdrd */
