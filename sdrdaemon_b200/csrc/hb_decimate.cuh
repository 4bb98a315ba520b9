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
 * Design (B200, HBM/ALU streaming kernel, no tensor cores):
 *   - one CTA walks one SEGMENT of one stream's time axis, chunk by chunk (C0 raw samples), so the
 *     61*(2^M-1)-sample filter warm-up is paid once per segment, not once per tile;
 *   - raw chunks arrive through the TMA 1-D bulk copy (cp.async.bulk + mbarrier), double buffered;
 *   - all M stages run software-pipelined ACROSS chunks: in step u the "unpack" work handles chunk u
 *     and stage j+1 handles chunk u-1-j, so every step has ~C0/8 independent FIR tasks and a single
 *     __syncthreads();
 *   - a FIR task produces 8 consecutive outputs of one stage for both I and Q from a 40-entry window
 *     of odd-phase inputs held in registers (pre-add of the symmetric taps, then IMAD with immediate
 *     coefficients: 16 IADD + 16 IMAD per output and component);
 *   - stage buffers are (I,Q) int32 pairs split by sample parity (E = x[2k], O = x[2k+1]) and padded
 *     8 -> 10 entries so that 128-bit shared loads at an 8-entry thread stride are bank-conflict free;
 *     each double-buffer slot is preceded by a 32-entry tail copy of the previous chunk.
 *
 * This file is single-source: nvcc builds the product kernel, tests/emu builds the same code for
 * the host (see sdrd_platform.cuh).
 */
#pragma once
#include "sdrd_platform.cuh"

namespace sdrd {
namespace hb {

constexpr int TAIL = 32;    /* entries (per parity) of the previous chunk kept in front of a slot */
constexpr int MAX_STAGES = 6;

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
    int seg_out;           /* outputs per segment, multiple of C0 >> M */
    int warm_chunks;       /* chunks processed ahead of every segment: warm_chunks * C0 >= 61 * (2^M - 1) */
    int round_add;         /* 0: EO1, 1: DB */
    int norm_shift, trunk_shift;
    int prologue;          /* 0: centred; 1: infradyne /4; 2: supradyne /4 in front of the cascade */
    long long origin;      /* cascade-input samples consumed since reset (only consulted when round_add) */
    uint32_t steer_zero, steer_one, steer_k32, steer_k256, steer_k8192; /* 0, 1, 32, 256, 8192: see Steer */
};

/* Geometry for a chunk of C0 cascade-input samples handled by C0/8 threads. */
template <int C0>
struct Geo {
    static constexpr int NT = C0 / 8;
    static constexpr int LOG2_2N1 = (C0 == 512 ? 6 : C0 == 1024 ? 7 : 8); /* log2(C0 / 8) */
    static_assert(C0 == 512 || C0 == 1024 || C0 == 2048, "supported chunk sizes");
    SDRD_HD static constexpr int stage_base(int m) { return 5 * (TAIL * m + C0 - (C0 >> m)); }
    SDRD_HD static constexpr int region_phys(int m) { return ((TAIL + (C0 >> (m + 1))) >> 3) * 10; }
    /* [T | S] region of stage m's output buffer: parity eo (0 = even samples), double-buffer slot */
    SDRD_HD static int2* region(int2* sbuf, int m, int eo, int slot)
    {
        return sbuf + stage_base(m) + (eo * 2 + slot) * region_phys(m);
    }
    SDRD_HD static constexpr size_t raw_bytes(int prologue) { return (size_t)2 * (prologue ? 4 : 1) * C0 * 4; }
    SDRD_HD static constexpr size_t smem_bytes(int M, int prologue)
    {
        return raw_bytes(prologue) + 128 + (size_t)(5 * (TAIL * M + C0 - (C0 >> M))) * 8;
    }
};

/* chunk size used for an M-stage cascade: the last stage's input needs >= TAIL entries per chunk
 * (C0 >= 32 * 2^M); smaller chunks mean smaller CTAs and more of them per SM */
#ifndef SDRD_HB_C0_MIN
#define SDRD_HB_C0_MIN 1024 /* build-time floor, for experiments */
#endif
constexpr int chunk_for(int M)
{
    return (M == 6 || SDRD_HB_C0_MIN >= 2048) ? 2048 : ((M == 5 || SDRD_HB_C0_MIN >= 1024) ? 1024 : 512);
}

/* logical entry -> physical entry (8 -> 10 padding) */
SDRD_DEVICE int phys(int k) { return k + 2 * (k >> 3); }

/* Pipe steering.  The FIR body is issue-bound: per output and component 16 pre-adds + 16
 * multiply-accumulates.  Left alone, ptxas turns about half of the pre-adds into IMAD.IADD, which
 * piles them onto the FMA pipe next to the IMADs (measured: fmaheavy 66 % busy, ALU 39 %, long
 * same-pipe runs).  `Steer` carries run-time constants the compiler cannot fold:
 *   zero   added as the THIRD operand of a pre-add  -> IADD3 with three sources, ALU pipe only;
 *   one    multiplier of an add written as IMAD      -> FMA pipe;
 *   k32, k256, k8192  multipliers of the power-of-two taps / centre tap -> IMAD instead of LEA.
 * Which taps use which form is fixed below so that ALU and FMA work per output are equal. */
struct Steer {
    uint32_t zero, one, k32, k256, k8192;
};

#ifndef SDRD_HB_FMA_ADD_TAPS
#define SDRD_HB_FMA_ADD_TAPS 1 /* taps (outermost first) whose pre-add runs on the FMA pipe */
#endif

/* 8 consecutive outputs n0 .. n0+7 (n0 = 8 i, chunk-local) of one half-band stage.
 * srcE/srcO point at the [T | S] region of the consumed chunk: logical entry TAIL + k is
 * E[k] = x[2k] resp. O[k] = x[2k+1] of the chunk, entries 0..TAIL-1 the previous chunk's tail. */
SDRD_DEVICE void fir8(const int2* SDRD_RESTRICT srcE, const int2* SDRD_RESTRICT srcO, int i, uint32_t acc0,
                      const Steer st, int2 (&y)[8])
{
    constexpr int H[16] = SDRD_HB64_TAPS;
    /* y[n] needs O[n-31 .. n] and E[n-15].  Window w[j] = O[n0 - 32 + j], j = 0..39: logical entries
     * 8i .. 8i+39 = padded groups i .. i+4. */
    int2 w[40];
    {
        const int4* po = reinterpret_cast<const int4*>(srcO + 10 * i);
#pragma unroll
        for (int g = 0; g < 5; g++) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int4 v = po[g * 5 + q];
                w[g * 8 + 2 * q] = make_int2(v.x, v.y);
                w[g * 8 + 2 * q + 1] = make_int2(v.z, v.w);
            }
        }
    }
    uint32_t aI[8], aQ[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        aI[r] = acc0;
        aQ[r] = acc0;
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
            /* O[n - t] = w[32 + r - t], O[n - 31 + t] = w[1 + r + t] */
            const uint32_t aIx = (uint32_t)w[32 + r - t].x, bIx = (uint32_t)w[1 + r + t].x;
            const uint32_t aQx = (uint32_t)w[32 + r - t].y, bQx = (uint32_t)w[1 + r + t].y;
            uint32_t sI, sQ;
            if (t < SDRD_HB_FMA_ADD_TAPS) {
                sI = mad_lo(aIx, st.one, bIx);
                sQ = mad_lo(aQx, st.one, bQx);
            } else {
                sI = add3(aIx, bIx, st.zero);
                sQ = add3(aQx, bQx, st.zero);
            }
            const uint32_t h = H[t] == 32 ? st.k32 : (H[t] == 256 ? st.k256 : (uint32_t)H[t]);
            aI[r] = mad_lo(sI, h, aI[r]);
            aQ[r] = mad_lo(sQ, h, aQ[r]);
        }
    }
    /* centre taps E[n0 - 15 + r]: logical entries 8i + 17 + r -> group i+2 (entries 16..23) and the
     * first entry of group i+3 */
    {
        const int4* pe = reinterpret_cast<const int4*>(srcE + 10 * (i + 2));
        int2 e[10];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int4 v = pe[q];
            e[2 * q] = make_int2(v.x, v.y);
            e[2 * q + 1] = make_int2(v.z, v.w);
        }
        e[8] = srcE[10 * (i + 3)];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            aI[r] = mad_lo((uint32_t)e[1 + r].x, st.k8192, aI[r]);
            aQ[r] = mad_lo((uint32_t)e[1 + r].y, st.k8192, aQ[r]);
            y[r] = make_int2(asr32(aI[r], HB_SHIFT), asr32(aQ[r], HB_SHIFT));
        }
    }
}

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

template <int M, int C0>
SDRD_KERNEL((C0 / 8), (512 / (C0 / 8))) decimate_kernel(Params p)
{
    typedef Geo<C0> G;
    constexpr int NT = G::NT;
    static_assert((C0 >> M) >= TAIL, "chunk too small for this many stages");
    SDRD_DYN_SMEM(smem);
    const int tid = (int)threadIdx.x;
    const int seg = (int)blockIdx.x;
    const int s = (int)blockIdx.y;
    const int pro = p.prologue;
    const int raw_per_chunk = pro ? 4 * C0 : C0; /* raw samples feeding one chunk of cascade input */
    uint32_t* raw = reinterpret_cast<uint32_t*>(smem);
    mbar_t* bars = reinterpret_cast<mbar_t*>(smem + G::raw_bytes(pro));
    int2* sbuf = reinterpret_cast<int2*>(smem + G::raw_bytes(pro) + 128);

    constexpr int out_per_chunk = C0 >> M;
    const long long seg_first_out = (long long)seg * p.seg_out;
    long long seg_n_out = p.n_out - seg_first_out;
    if (seg_n_out > p.seg_out) seg_n_out = p.seg_out;
    const int data_chunks = (int)((seg_n_out + out_per_chunk - 1) / out_per_chunk);
    const int NC = p.warm_chunks + data_chunks;
    /* cascade-input index (relative to this call's first new sample) of chunk 0 */
    const long long first_in = (seg_first_out << M) - (long long)p.warm_chunks * C0;
    const uint32_t* src = p.in + (long long)s * p.in_stride + first_in * (pro ? 4 : 1);
    uint32_t* dst = p.out + (long long)s * p.out_stride;
    const long long abs0 = p.origin + first_in;
    const uint32_t acc0 = (uint32_t)p.round_add << HB_SHIFT;
    const uint32_t chunk_bytes = (uint32_t)raw_per_chunk * 4u;
    const Steer steer = {p.steer_zero, p.steer_one, p.steer_k32, p.steer_k256, p.steer_k8192};

    /* ---- this thread's fixed role: FIR task t = tid of every step (stage j+1, outputs 8i..8i+7 of
     *      the chunk); there are C0/8 * (1 - 2^-M) < NT tasks per step ---- */
    constexpr int two_n1 = C0 / 8;
    constexpr int n_tasks = two_n1 - (two_n1 >> M);
    const bool has_task = tid < n_tasks;
    const int tj = has_task ? __clz(two_n1 - 1 - tid) - (32 - G::LOG2_2N1) : 0;
    const int ti = tid - (two_n1 - (two_n1 >> tj));
    const bool t_final = tj + 1 == M;
    const int2* const t_srcE = G::region(sbuf, tj, 0, 0);
    const int2* const t_srcO = G::region(sbuf, tj, 1, 0);
    const int t_src_slot = G::region_phys(tj);           /* entries between slot 0 and slot 1 */
    int2* const t_dstE = G::region(sbuf, t_final ? tj : tj + 1, 0, 0) + phys(TAIL + 4 * ti);
    int2* const t_dstO = G::region(sbuf, t_final ? tj : tj + 1, 1, 0) + phys(TAIL + 4 * ti);
    const int t_dst_slot = G::region_phys(t_final ? tj : tj + 1);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < 2 && c < NC; c++) {
            mbar_arrive_expect_tx(&bars[c], chunk_bytes);
            tma_load_1d(raw + (size_t)c * raw_per_chunk, src + (size_t)c * raw_per_chunk, chunk_bytes, &bars[c]);
        }
    }

    for (int u = 0; u < NC + M; u++) {
        /* ---- stage 0: unpack raw chunk u into (I,Q) int32 pairs, split by parity ---- */
        if (u < NC) {
            mbar_wait(&bars[u & 1], (uint32_t)((u >> 1) & 1));
            const uint4* r4 = reinterpret_cast<const uint4*>(raw + (size_t)(u & 1) * raw_per_chunk);
            int2* E0 = G::region(sbuf, 0, 0, u & 1);
            int2* O0 = G::region(sbuf, 0, 1, u & 1);
            if (!pro) {
#pragma unroll
                for (int qq = 0; qq < 2; qq++) {
                    const int q = tid + qq * NT;
                    uint4 v = r4[q];
                    int ph = phys(TAIL + 2 * q);
                    *reinterpret_cast<int4*>(E0 + ph) = make_int4(s16lo(v.x), s16hi(v.x), s16lo(v.z), s16hi(v.z));
                    *reinterpret_cast<int4*>(O0 + ph) = make_int4(s16lo(v.y), s16hi(v.y), s16lo(v.w), s16hi(v.w));
                }
            } else {
                /* 16 raw samples -> 4 cascade inputs -> 2 E entries + 2 O entries */
                for (int q = tid; q < C0 / 4; q += NT) {
                    int2 a = rot4(r4[4 * q], pro), b = rot4(r4[4 * q + 1], pro);
                    int2 c = rot4(r4[4 * q + 2], pro), d = rot4(r4[4 * q + 3], pro);
                    int ph = phys(TAIL + 2 * q);
                    *reinterpret_cast<int4*>(E0 + ph) = make_int4(a.x, a.y, c.x, c.y);
                    *reinterpret_cast<int4*>(O0 + ph) = make_int4(b.x, b.y, d.x, d.y);
                }
            }
        }

        /* ---- half-band task: stage tj+1 consumes stage-tj chunk u-1-tj ---- */
        const int c = u - 1 - tj;
        if (has_task && c >= 0 && c < NC) {
            const int slot = c & 1;
            int2 y[8];
            fir8(t_srcE + slot * t_src_slot, t_srcO + slot * t_src_slot, ti, acc0, steer, y);
            if (p.round_add) {
                /* DB: the reference's stages start from all-zero state, but a DB stage maps zero
                 * input to 1; outputs that lie before the stream origin must read as 0. */
                const long long a = ((abs0 + (long long)c * C0) >> (tj + 1)) + 8 * ti;
                if (a < 0) {
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (a + r < 0) y[r] = make_int2(0, 0);
                }
            }
            if (!t_final) {
                int4* qe = reinterpret_cast<int4*>(t_dstE + slot * t_dst_slot);
                int4* qo = reinterpret_cast<int4*>(t_dstO + slot * t_dst_slot);
                qe[0] = make_int4(y[0].x, y[0].y, y[2].x, y[2].y);
                qe[1] = make_int4(y[4].x, y[4].y, y[6].x, y[6].y);
                qo[0] = make_int4(y[1].x, y[1].y, y[3].x, y[3].y);
                qo[1] = make_int4(y[5].x, y[5].y, y[7].x, y[7].y);
            } else if (c >= p.warm_chunks) {
                const long long n = seg_first_out + (long long)(c - p.warm_chunks) * out_per_chunk + 8 * ti;
                uint32_t o[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    uint32_t vi = (uint32_t)asr32((uint32_t)y[r].x << p.norm_shift, p.trunk_shift);
                    uint32_t vq = (uint32_t)asr32((uint32_t)y[r].y << p.norm_shift, p.trunk_shift);
                    o[r] = (vi & 0xFFFFu) | (vq << 16);
                }
                if (n + 8 <= p.n_out) {
                    uint4* q = reinterpret_cast<uint4*>(dst + n);
                    q[0] = make_uint4(o[0], o[1], o[2], o[3]);
                    q[1] = make_uint4(o[4], o[5], o[6], o[7]);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; r++)
                        if (n + r < p.n_out) dst[n + r] = o[r];
                }
            }
        }

        /* ---- tail copies: last TAIL entries of the chunk consumed in this step -> front of the
         *      other slot, where the next chunk's consumer expects its history ---- */
        for (int t = tid; t < 32 * M; t += NT) {
            const int j = t >> 5, eo = (t >> 4) & 1, unit = t & 15;
            const int cc = u - 1 - j;
            if (cc < 0 || cc >= NC) continue;
            const int n = C0 >> (j + 1);
            const int4* sp = reinterpret_cast<const int4*>(G::region(sbuf, j, eo, cc & 1) + phys(n)); /* entry TAIL+n-32 */
            int4* dp = reinterpret_cast<int4*>(G::region(sbuf, j, eo, (cc & 1) ^ 1));
            /* 4 padded groups of 8 entries: 5 int4 per group, the first 4 carry data */
            dp[(unit >> 2) * 5 + (unit & 3)] = sp[(unit >> 2) * 5 + (unit & 3)];
        }
        __syncthreads();
        if (tid == 0 && u + 2 < NC) {
            mbar_arrive_expect_tx(&bars[u & 1], chunk_bytes);
            tma_load_1d(raw + (size_t)(u & 1) * raw_per_chunk, src + (size_t)(u + 2) * raw_per_chunk, chunk_bytes,
                        &bars[u & 1]);
        }
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
