/*
 * hb_decimate_warp.cuh -- K1, warp-private form: one WARP (a 32-thread CTA) walks one segment of one
 * stream through all M half-band stages.  Same arithmetic, Params and results as
 * hb::decimate_kernel (hb_decimate.cuh; reference include/IntHalfbandFilterEO1.h:34-42,100-147,
 * include/IntHalfbandFilterDB.h:80-107, sdmnbase/Decimators.cpp:94-1305), different decomposition:
 *
 *   - a step consumes WC0 = 512 cascade-input samples (2 KB, one TMA bulk copy, double buffered) and
 *     every lane owns one FIR task of 32 outputs x one component for the life of the warp:
 *         lanes  0-15  stage 1 (8 tasks per component)      lanes 24-27  stage 3
 *         lanes 16-23  stage 2                              lanes 28-29  stage 4
 *         lanes 30-31  stage 5 on every 2nd step (32 outputs need two chunks) and, on the steps in
 *                      between, stage 6 on every 4th
 *     so 30 of 32 lanes (M = 4) .. 31.5 of 32 (M = 6) do FIR work in every step and there is no
 *     CTA-wide barrier anywhere: the stages are software-pipelined across steps (stage j works on the
 *     chunk unpacked j steps earlier) and a step is  loads | __syncwarp | arithmetic + stores |
 *     __syncwarp.  Warps drift apart freely, which is what keeps the four schedulers of an SM fed;
 *   - a task of 32 outputs reads a window of 64 odd-phase + 33 even-phase inputs (3 words per output;
 *     a 16-output task needs 4): shared-memory bandwidth (128 B/clk/SM) is the second limit of this
 *     kernel right behind instruction issue, see DESIGN.md;
 *   - because every load of a step precedes every store, a stage buffer is a single linear region
 *     [32 history entries | the step's new entries] per (stage, component, parity) -- no double
 *     buffer; a few lanes shift the last 32 entries down after the consumers have read them;
 *   - regions are padded 32 -> 36 words and placed at chosen residues mod 32 words so that the eight
 *     lanes of every quarter-warp hit eight different 16-byte bank groups on every LDS.128 / STS.128.
 */
#pragma once
#include "hb_decimate.cuh"

namespace sdrd {
namespace hb {

constexpr int WC0 = 512; /* cascade-input samples per step */
constexpr int BLK = 36;  /* 32 entries + 4 words of padding */

/* logical entry -> word offset inside a region */
SDRD_HD constexpr int wphys(int k) { return k + 4 * (k >> 5); }
/* logical entries of the region holding stage m's output (m = 0: the unpacked input), one parity, one
 * component: 32 history + what one consumer step reads as new (m >= 4: two producer steps' worth) */
SDRD_HD constexpr int wregion_entries(int m) { return m <= 3 ? 32 + (256 >> m) : 64; }
SDRD_HD constexpr int wregion_words(int m) { return wregion_entries(m) / 32 * BLK; }
/* start of a region modulo 32 words (a multiple of 4 words = one 16-byte bank group), chosen so that
 * the mixed quarter-warps 16-23 (stage 2: 4 I + 4 Q tasks) and 24-31 (stages 3..6) are conflict free
 * both when they load their windows and when they store their results */
SDRD_HD constexpr int wresidue(int m, int comp)
{
    return 4 * (m == 0 ? 0 : m == 1 ? (comp ? 4 : 0) : m == 2 ? (comp ? 2 : 0) : m == 3 ? (comp ? 5 : 4) : (comp ? 7 : 6));
}
/* word offset of region (m, parity, comp) from the start of the plane area; m = 6 gives the total */
SDRD_HD constexpr int wplane_off(int m, int parity, int comp)
{
    int off = 0;
    for (int mm = 0; mm < 6; mm++)
        for (int pp = 0; pp < 2; pp++)
            for (int cc = 0; cc < 2; cc++) {
                const int r = wresidue(mm, cc);
                off += ((r - off) % 32 + 32) % 32;
                if (mm == m && pp == parity && cc == comp) return off;
                off += wregion_words(mm);
            }
    return off;
}
SDRD_HD constexpr int wfin_n(int M) { return (WC0 >> M) > 32 ? (WC0 >> M) : 32; } /* outputs per pack event */
SDRD_HD constexpr int wfin_stride(int M) { return wfin_n(M) + 4; }                  /* words between the I and Q results */
SDRD_HD constexpr int wmacro(int M) { return M <= 4 ? 1 : 1 << (M - 4); }           /* steps per pack event */
/* steps between unpacking the first chunk of a pack event and packing its outputs */
SDRD_HD constexpr int wdelay(int M) { return M <= 4 ? M + 1 : (M == 5 ? 7 : 10); }
SDRD_HD constexpr int wraw_words(int PRO) { return PRO ? 4 * WC0 : WC0; }
/* raw double buffer | 2 mbarriers | regions of stages 0..M-1 | 8 words | last stage's results */
SDRD_HD constexpr size_t wsmem_bytes(int M, int PRO)
{
    return (size_t)2 * wraw_words(PRO) * 4 + 128 + (size_t)(wplane_off(M, 0, 0) + 40) * 4 + (size_t)2 * wfin_stride(M) * 4;
}
/* warm-up chunks in front of a segment: >= 61 * (2^M - 1) samples, whole pack events */
SDRD_HD constexpr int wwarm_chunks(int M)
{
    return ((61 * ((1 << M) - 1) + WC0 - 1) / WC0 + wmacro(M) - 1) / wmacro(M) * wmacro(M);
}

/* 32 consecutive outputs n0 .. n0+31 of one half-band stage, one component.  srcO/srcE point at the
 * task's first 32-entry block of the consumed regions: w[j] = O[n0 - 32 + j], e[j] = E[n0 - 16 + j]. */
struct Fir32Regs {
    uint32_t w[64];
    uint32_t e[33];
};

SDRD_DEVICE void fir32_load(const int* SDRD_RESTRICT srcE, const int* SDRD_RESTRICT srcO, Fir32Regs& f)
{
#pragma unroll
    for (int b = 0; b < 2; b++) {
        const int4* po = reinterpret_cast<const int4*>(srcO + BLK * b);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int4 v = po[q];
            f.w[32 * b + 4 * q] = (uint32_t)v.x;
            f.w[32 * b + 4 * q + 1] = (uint32_t)v.y;
            f.w[32 * b + 4 * q + 2] = (uint32_t)v.z;
            f.w[32 * b + 4 * q + 3] = (uint32_t)v.w;
        }
    }
    /* entries 16 .. 31 of the first block, 0 .. 16 of the second */
#pragma unroll
    for (int q = 0; q < 8; q++) {
        int4 v = *reinterpret_cast<const int4*>(srcE + (q < 4 ? 16 + 4 * q : BLK + 4 * (q - 4)));
        f.e[4 * q] = (uint32_t)v.x;
        f.e[4 * q + 1] = (uint32_t)v.y;
        f.e[4 * q + 2] = (uint32_t)v.z;
        f.e[4 * q + 3] = (uint32_t)v.w;
    }
    f.e[32] = (uint32_t)srcE[BLK + 16];
}

/* outputs 16 h .. 16 h + 15 of the task (h = 0, 1) */
template <int HALF>
SDRD_DEVICE void fir32_half(const Fir32Regs& f, uint32_t acc0, const Steer st, int (&y)[16])
{
    constexpr int H[16] = SDRD_HB64_TAPS;
    uint32_t acc[16];
#pragma unroll
    for (int r = 0; r < 16; r++) acc[r] = acc0;
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            const uint32_t a = f.w[16 * HALF + 32 + r - t], b = f.w[16 * HALF + 1 + r + t];
            const uint32_t sum = t < SDRD_HB_FMA_ADD_TAPS ? mad_lo(a, st.one, b) : add3(a, b, st.zero);
            const uint32_t h = H[t] == 32 ? st.k32 : (H[t] == 256 ? st.k256 : (uint32_t)H[t]);
            acc[r] = mad_lo(sum, h, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 16; r++) y[r] = asr32(mad_lo(f.e[16 * HALF + 1 + r], st.k8192, acc[r]), HB_SHIFT);
}

template <int M, int DB, int PRO>
SDRD_KERNEL(32, 12) decimate_warp_kernel(Params p)
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
    const int seg = (int)blockIdx.x;
    const int s = (int)blockIdx.y;
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem);
    mbar_t* bars = reinterpret_cast<mbar_t*>(smem + (size_t)2 * RAWW * 4);
    int* planes = reinterpret_cast<int*>(smem + (size_t)2 * RAWW * 4 + 128);
    int* fin = planes + wplane_off(M, 0, 0) + 8; /* residue 8 words: the last stage's stores miss the others' banks */

    const long long seg_first_out = (long long)seg * p.seg_out;
    long long seg_n_out = p.n_out - seg_first_out;
    if (seg_n_out > p.seg_out) seg_n_out = p.seg_out;
    constexpr int warm_macro = wwarm_chunks(M) / MACRO;
    const int n_macro = warm_macro + (int)((seg_n_out + FN - 1) / FN);
    const int NC = n_macro * MACRO; /* chunks to unpack */
    const int u_last = (n_macro - 1) * MACRO + DELAY;
    const long long first_in = (seg_first_out << M) - (long long)wwarm_chunks(M) * WC0;
    const uint32_t* src = p.in + (long long)s * p.in_stride + first_in * (PRO ? 4 : 1);
    uint32_t* dst = p.out + (long long)s * p.out_stride + seg_first_out - (long long)warm_macro * FN;
    const long long seg_room = p.n_out - seg_first_out + (long long)warm_macro * FN; /* valid: index < seg_room */
    const long long abs0 = p.origin + first_in;
    constexpr uint32_t acc0 = (uint32_t)DB << HB_SHIFT;
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
    const int* srcE = planes + wplane_off(tm, 0, tcomp) + BLK * ti;
    const int* srcO = planes + wplane_off(tm, 1, tcomp) + BLK * ti;
    int* dstE;
    int* dstO;
    if (tm + 1 < M) {
        dstE = planes + wplane_off(tm + 1, 0, tcomp) + wphys(32 + 16 * ti);
        dstO = planes + wplane_off(tm + 1, 1, tcomp) + wphys(32 + 16 * ti);
    } else {
        dstE = fin + tcomp * FS + 16 * ti;
        dstO = dstE + FN / 2;
    }
    /* M = 6: lanes 30/31 alternate between stage 5 (regions 4 -> 5) and stage 6 (regions 5 -> fin) */
    const int* srcE_b = planes + wplane_off(M == 6 ? 5 : 0, 0, tcomp);
    const int* srcO_b = planes + wplane_off(M == 6 ? 5 : 0, 1, tcomp);
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
        tl_src[k] = planes + wplane_off(m, parity, comp) + 4 * unit + tl_back[k];
        tl_m[k] = on ? m : -1;
    }

    /* ---- unpack (no prologue): lane q reads 16-byte unit 32 k + q of the raw chunk, k = 0..3: samples
     *      4 (32 k + q) .. + 3 -> entries 64 k + 2 q, + 1 of the four stage-0 regions.
     *      (/4 prologue: lane q turns raw unit 32 k + q, k = 0..15, into cascade input 32 k + q) ---- */
    int* const up = planes + (PRO ? wphys(32 + (lane >> 1)) : wphys(32 + 2 * lane));
    constexpr int UP_EI = wplane_off(0, 0, 0), UP_EQ = wplane_off(0, 0, 1), UP_OI = wplane_off(0, 1, 0), UP_OQ = wplane_off(0, 1, 1);

    /* ---- pack: lane l packs outputs OPL*l .. OPL*l+OPL-1 of an event; output r of a component sits at
     *      fin[r even ? r/2 : FN/2 + r/2] ---- */
    const int* const pk = fin + (OPL == 1 ? ((lane & 1) ? FN / 2 + (lane >> 1) : (lane >> 1)) : lane * (OPL / 2));

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        for (int c = 0; c < 2 && c < NC; c++) {
            mbar_arrive_expect_tx(&bars[c], chunk_bytes);
            tma_load_1d(raw + (size_t)c * RAWW, src + (size_t)c * RAWW, chunk_bytes, &bars[c]);
        }
    }
    SDRD_SYNCWARP();

    for (int u = 0; u <= u_last; u++) {
        /* ================= phase A: every shared-memory read of the step ================= */
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
        Fir32Regs fr;
        if (task_on) fir32_load(sE, sO, fr);

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

        const bool unpack_now = u < NC;
        uint4 rw[4];
        int2 x[16];
        if (unpack_now) {
            mbar_wait(&bars[u & 1], (uint32_t)((u >> 1) & 1));
            const uint4* r4 = reinterpret_cast<const uint4*>(raw + (size_t)(u & 1) * RAWW);
            if (!PRO) {
#pragma unroll
                for (int k = 0; k < 4; k++) rw[k] = r4[32 * k + lane];
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = rot4(r4[32 * k + lane], p.prologue);
            }
        }
        SDRD_SYNCWARP();

        /* ================= phase B: arithmetic and every shared-memory write of the step ========= */
#pragma unroll
        for (int k = 0; k < TSLOTS; k++)
            if (tl_on[k]) *reinterpret_cast<int4*>(const_cast<int*>(tl_src[k]) - tl_back[k]) = tl_v[k];

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
            /* DB: the reference's stages start from all-zero state, but a DB stage maps zero input to 1;
             * outputs that lie before the stream origin must read as 0. */
            const long long a0 = DB ? ((abs0 + (long long)c_first * WC0) >> t_stage) + 32 * ti : 0;
            int y[16];
            fir32_half<0>(fr, acc0, steer, y);
            if (DB && a0 < 0) {
#pragma unroll
                for (int r = 0; r < 16; r++)
                    if (a0 + r < 0) y[r] = 0;
            }
            reinterpret_cast<int4*>(dE)[0] = make_int4(y[0], y[2], y[4], y[6]);
            reinterpret_cast<int4*>(dE)[1] = make_int4(y[8], y[10], y[12], y[14]);
            reinterpret_cast<int4*>(dO)[0] = make_int4(y[1], y[3], y[5], y[7]);
            reinterpret_cast<int4*>(dO)[1] = make_int4(y[9], y[11], y[13], y[15]);
            fir32_half<1>(fr, acc0, steer, y);
            if (DB && a0 + 16 < 0) {
#pragma unroll
                for (int r = 0; r < 16; r++)
                    if (a0 + 16 + r < 0) y[r] = 0;
            }
            reinterpret_cast<int4*>(dE)[2] = make_int4(y[0], y[2], y[4], y[6]);
            reinterpret_cast<int4*>(dE)[3] = make_int4(y[8], y[10], y[12], y[14]);
            reinterpret_cast<int4*>(dO)[2] = make_int4(y[1], y[3], y[5], y[7]);
            reinterpret_cast<int4*>(dO)[3] = make_int4(y[9], y[11], y[13], y[15]);
        }
        SDRD_SYNCWARP();
        if (lane == 0 && u + 2 < NC) {
            mbar_arrive_expect_tx(&bars[u & 1], chunk_bytes);
            tma_load_1d(raw + (size_t)(u & 1) * RAWW, src + (size_t)(u + 2) * RAWW, chunk_bytes, &bars[u & 1]);
        }
    }
}

} /* namespace hb */
} /* namespace sdrd */
