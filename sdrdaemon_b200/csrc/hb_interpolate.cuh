/*
 * hb_interpolate.cuh -- K4: the Tx side's half-band interpolation cascade for interleaved int16 I/Q.
 *
 * Replaces Upsampler::process (reference sdmnbase/Upsampler.cpp:57-84) ->
 * Interpolators::interpolate{2,4,8,16,32,64}_cen (sdmnbase/Interpolators.cpp:23-606) built from
 * IntHalfbandFilterEO1<N>::myInterpolate / doInterpolateFIR (include/IntHalfbandFilterEO1.h:44-65,
 * 149-168; IntHalfbandFilterDB.h:51-72,109-127 is the same arithmetic) with N = 64 for the first
 * stage, 32 for the second and 16 for the others (include/Interpolators.h:31-33).  Per stage, with
 * L = N/2 and c[] the N/4 unique taps (HBFilterTraits.cpp, hbShift = 14),
 *
 *     y[2k]   = x[k - L/2]
 *     y[2k+1] = ( sum_{i < L/2} c[i] * ( x[k - L + 1 + i] + x[k - i] ) ) >> 13        (wrapping int32)
 *
 * no scaling between stages, and the final samples are truncated to int16 (IQSample::setReal).
 * Reference quirk, reproduced: interpolate64_cen runs the five stages of interpolate32_cen and then
 * emits 64 samples per input sample of which the last 32 are zero (Interpolators.cpp:363-605).
 *
 * One kernel for the steady state, the warp-private, barrier-free interpolate_warp_kernel (second half of this file),
 * and i_stateful_kernel for the samples right after a reconfiguration.  (A CTA-wide tiled kernel -- a tile of
 * 4096 >> S input samples plus halo per CTA, stage after stage through shared memory with a barrier in between --
 * was the product path of round 1 and stayed for interp = 64 until the warp kernel passed it there as well: x64
 * 0.633 -> 0.575 ms.  It is gone; profiles/r1 holds its measurements.)
 *
 * Single-source: nvcc builds the product kernel, tests/emu the same code for the host.
 */
#pragma once
#include "sdrd_platform.cuh"

namespace sdrd {
namespace hbi {

constexpr int NT = 256;
constexpr int HIST = 64; /* input samples of history kept in front of a call's samples (>= 42) */

#define SDRD_HB64_ITAPS {-7, 11, -20, 32, -49, 71, -101, 140, -190, 256, -345, 469, -656, 978, -1698, 5201}
#define SDRD_HB32_ITAPS {-30, 63, -135, 261, -469, 830, -1605, 5176}
#define SDRD_HB16_ITAPS {-85, 380, -1246, 5041}

/* ring length L (= filter order / 2) of stage s = 1..5 */
SDRD_HD constexpr int ring_len(int s) { return s == 1 ? 32 : s == 2 ? 16 : 8; }
struct Params {
    const uint32_t* in;   /* stream s, sample k (k >= -HIST): in[s * in_stride + k] */
    long long in_stride;  /* words */
    uint32_t* out;        /* out[s * out_stride + n] */
    long long out_stride;
    long long n_in;       /* input samples per stream */
    int log2_interp;      /* 1..6; stages run S = min(log2_interp, 5) */
};

SDRD_DEVICE uint32_t pack16(int2 v) { return ((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16); }

/* ------------------------------------------------------------------------------------------------------------
 * Explicit per-stage state (the Tx twin of hb::stateful_kernel): the reference keeps persistent stage objects
 * m_interpolator2 .. m_interpolator64 (include/Interpolators.h:52-58), cascade position k always runs object k
 * (Interpolators.cpp:36, 60-63, 93-101 ...; object 6 is never used, see the interpolate64_cen quirk above).  After
 * Upsampler::configure (Upsampler.cpp:32-55) a stage the new cascade uses continues from what it saw last under
 * the previous configuration (or zeros).  i_stateful_kernel runs the cascade from
 *     state[stream][stage 0..4][component I/Q][32]  =  the stage's last 32 inputs, oldest first (m_samples)
 * and writes the states back; out == nullptr: states only.  One CTA per stream, chunks of ICH input samples.
 * ------------------------------------------------------------------------------------------------------------ */
constexpr int ICH = 64;
constexpr int IST = 32;
constexpr int ISTATE_WORDS = 5 * 2 * IST;
SDRD_HD constexpr int ibuf_off(int k) { return k == 0 ? 0 : ibuf_off(k - 1) + 2 * (IST + (ICH << (k - 1))); }
SDRD_HD constexpr size_t i_stateful_smem_bytes() { return (size_t)ibuf_off(5) * 4; }

struct IStateParams {
    const uint32_t* in;   /* in[s * in_stride + k] */
    long long in_stride;
    uint32_t* out;        /* out[s * out_stride + n] or nullptr */
    long long out_stride;
    int* state;           /* [n_streams][ISTATE_WORDS] */
    long long n_in;       /* input samples per stream */
    int log2_interp;      /* 1..6; stages run S = min(log2_interp, 5) */
};

SDRD_KERNEL(NT, 1) i_stateful_kernel(IStateParams p)
{
    constexpr int C64[16] = SDRD_HB64_ITAPS;
    constexpr int C32[8] = SDRD_HB32_ITAPS;
    constexpr int C16[4] = SDRD_HB16_ITAPS;
    SDRD_DYN_SMEM(smem);
    int* const buf = reinterpret_cast<int*>(smem);
    const int tid = (int)threadIdx.x;
    const int s = (int)blockIdx.x;
    const int wo = p.log2_interp;
    const int S = wo < 5 ? wo : 5;
    const uint32_t* in = p.in + (long long)s * p.in_stride;
    uint32_t* out = p.out ? p.out + (long long)s * p.out_stride : nullptr;
    int* state = p.state + (long long)s * ISTATE_WORDS;
    auto B = [&](int k, int c) -> int* { return buf + ibuf_off(k) + c * (IST + (ICH << k)); };

    for (int i = tid; i < S * 2 * IST; i += NT) {
        const int k = i / (2 * IST), c = (i / IST) & 1, j = i % IST;
        B(k, c)[j] = state[(k * 2 + c) * IST + j];
    }
    __syncthreads();
    for (long long k0 = 0; k0 < p.n_in; k0 += ICH) {
        const int len = (int)(p.n_in - k0 < ICH ? p.n_in - k0 : ICH);
        for (int i = tid; i < len; i += NT) {
            const uint32_t v = in[k0 + i];
            B(0, 0)[IST + i] = (int)(int16_t)(v & 0xFFFFu);
            B(0, 1)[IST + i] = ((int)v) >> 16;
        }
        __syncthreads();
        for (int k = 0; k < S; k++) {
            const int L = ring_len(k + 1);
            const int nk = len << k; /* inputs of this stage in the chunk */
            for (int j = tid; j < nk; j += NT) {
                int ev[2], od[2];
                for (int c = 0; c < 2; c++) {
                    const int* x = B(k, c) + IST; /* x[j], j >= -IST */
                    uint32_t acc = 0;
                    for (int t = 0; t < L / 2; t++) {
                        const int co = L == 32 ? C64[t & 15] : L == 16 ? C32[t & 7] : C16[t & 3];
                        acc += ((uint32_t)x[j - L + 1 + t] + (uint32_t)x[j - t]) * (uint32_t)co;
                    }
                    ev[c] = x[j - L / 2];
                    od[c] = asr32(acc, 13);
                }
                if (k + 1 < S) {
                    B(k + 1, 0)[IST + 2 * j] = ev[0];
                    B(k + 1, 1)[IST + 2 * j] = ev[1];
                    B(k + 1, 0)[IST + 2 * j + 1] = od[0];
                    B(k + 1, 1)[IST + 2 * j + 1] = od[1];
                } else if (out) {
                    /* cascade sample n is emitted at (n >> S << wo) + (n & (2^S - 1)) */
                    const long long n = (k0 << S) + 2 * (long long)j;
                    const long long pos = ((n >> S) << wo) + (n & ((1 << S) - 1));
                    out[pos] = pack16(make_int2(ev[0], ev[1]));
                    out[pos + 1] = pack16(make_int2(od[0], od[1]));
                }
            }
            __syncthreads();
        }
        if (out && wo > S) { /* interpolate64_cen: 32 zero samples after every 32 */
            const int zw = (1 << wo) - (1 << S);
            for (int i = tid; i < len * zw; i += NT) out[((k0 + i / zw) << wo) + (1 << S) + i % zw] = 0u;
        }
        int keep[(5 * 2 * IST + NT - 1) / NT];
        for (int r = 0; r < (5 * 2 * IST + NT - 1) / NT; r++) {
            const int i = tid + r * NT;
            const int k = i / (2 * IST), c = (i / IST) & 1, j = i % IST;
            keep[r] = k < S ? B(k, c)[(len << k) + j] : 0;
        }
        __syncthreads();
        for (int r = 0; r < (5 * 2 * IST + NT - 1) / NT; r++) {
            const int i = tid + r * NT;
            const int k = i / (2 * IST), c = (i / IST) & 1, j = i % IST;
            if (k < S) B(k, c)[j] = keep[r];
        }
        __syncthreads();
    }
    for (int i = tid; i < S * 2 * IST; i += NT) {
        const int k = i / (2 * IST), c = (i / IST) & 1, j = i % IST;
        state[(k * 2 + c) * IST + j] = B(k, c)[j];
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * Warp-private form (the product kernel).
 * One warp (a 32-thread CTA) walks a segment of one stream in steps of WC = 64 input samples and runs the
 * stages one after the other, every lane busy in every stage and no CTA barrier anywhere:
 *     stage 1: 64 steps = 32 lanes x 2     stage 3: 256 steps = 32 lanes x 8
 *     stage 2: 128 steps = 32 lanes x 4    stage 4 / 5: 2 / 4 passes of 32 lanes x 8
 * A lane computes its N steps from one register window of L + N entries ((L + N) / (2 N) entries read per output
 * sample instead of (L + 2) / 4).  Each stage buffer is [L entries of history | the step's new entries]: nothing is
 * recomputed at segment edges beyond one warm-up step and the history is a copy of the
 * last L entries at the end of a step.  Buffers are XOR-swizzled for their reader (16-byte unit u lives at
 * u ^ ((u >> 3) & SW), SW = 3 for windows 4 units apart, 1 for 2 units apart): conflict-free window loads.
 * The last stage's samples (16 per lane and pass) go through a swizzled staging area so that the global stores are
 * whole 512-byte rows per instruction: a global store costs one L1 wavefront per 128-byte line it touches, and 64
 * bytes per lane straight from registers touch 16 lines per instruction.
 * A segment starts with one warm-up step (the cascade looks back 42 input samples, HIST = 64 are there).
 * ------------------------------------------------------------------------------------------------------------ */
constexpr int WC = 64;
#ifndef SDRD_K4_N16
#define SDRD_K4_N16 1 /* stages 4 and 5 take 16 steps per lane (one 128-byte swizzle row per lane and 8 steps) */
#endif
SDRD_HD constexpr int w_nstep(int s) { return s == 1 ? 2 : s == 2 ? 4 : (s == 3 || !SDRD_K4_N16) ? 8 : 16; }
SDRD_HD constexpr int w_passes(int s) { return (WC << (s - 1)) / (32 * w_nstep(s)); }
SDRD_HD constexpr int w_sw_of_n(int n) { return n == 16 ? 7 : n == 8 ? 3 : n == 4 ? 1 : 0; }
SDRD_HD constexpr int w_sw(int b) { return w_sw_of_n(w_nstep(b + 1)); }  /* buffer b is read by stage b + 1 */
/* history entries in front of buffer b: L of the reader, so that a task's window starts at its first step's unit --
 * or 16 for a reader with 16 steps per lane, whose windows then start 4 units into a swizzle row of 8 */
SDRD_HD constexpr int w_hist(int b) { return w_nstep(b + 1) == 16 ? 16 : ring_len(b + 1); }
SDRD_HD constexpr int w_woff(int b) { return (w_hist(b) - ring_len(b + 1)) / 2; } /* unit of task 0's window */
SDRD_HD constexpr int w_len(int b) { return (w_hist(b) + (WC << b) + 15) & ~15; }
SDRD_HD constexpr int w_off(int b)
{
    int o = 0;
    for (int t = 0; t < b; t++) o += w_len(t);
    return o;
}
SDRD_HD constexpr int w_stage_words(int S) { return 64 * w_nstep(S); } /* one pass of the last stage: 32 lanes x 2 N samples */
/* the staging tile of the TMA tensor store has to sit on a 1024-byte boundary (128-byte swizzle): 1 KB of slack */
SDRD_HD constexpr size_t w_smem_bytes(int S) { return (size_t)w_off(S) * 8 + (size_t)w_stage_words(S) * 4 + 1024; }

template <int SW>
SDRD_DEVICE int swz(int u) { return SW ? (u ^ ((u >> 3) & SW)) : u; }
SDRD_DEVICE void ld_unit(const int2* SDRD_RESTRICT buf, int unit, int2* w)
{
    const int4 v = *reinterpret_cast<const int4*>(buf + 2 * unit);
    w[0] = make_int2(v.x, v.y);
    w[1] = make_int2(v.z, v.w);
}
SDRD_DEVICE void st_unit(int2* SDRD_RESTRICT buf, int unit, int2 a, int2 b)
{
    *reinterpret_cast<int4*>(buf + 2 * unit) = make_int4(a.x, a.y, b.x, b.y);
}

/* N consecutive steps from the window that starts at unit u0 of a buffer swizzled for N (u0 = N / 2 * task, for
 * N = 16: 4 + 8 * task); every finished step i is handed to sink(i, ev, od): ev = x[k + i - L/2], od = FIR at k + i */
#ifndef SDRD_K4_PACKED_X0
#define SDRD_K4_PACKED_X0 1 /* buffer 0 keeps the raw {int16 I, int16 Q} words (half the window bytes of stage 1), sign-extended in registers */
#endif
/* Pipe steering, as in K1 (hb_decimate.cuh, Steer): a step is T pre-adds + T multiply-accumulates per component; left
 * alone ptxas writes 63 % of the pre-adds as IMAD.IADD, next to the IMADs on the FMA pipe (67 % busy, ALU 42 %).
 * Tap t's pre-add is a three-source IADD3 (ALU only) unless t % SDRD_K4_FMA_ADD_MOD == 0, where it is an IMAD by a
 * run-time 1 (FMA pipe); 0 = every pre-add on the ALU pipe, -1 = left to ptxas.
 * Measured (x16, config 6, ms): -1: 0.2172, 2: 0.2162, 3: 0.2162, 0 / 4 / 5 / 6 / 8 / 16: 0.2069 / 0.2066 / 0.2058 /
 * 0.2068 / 0.2062 / 0.2058.  Moving the >> 13 to the FMA pipe as well (high word of acc * 2^19, IMAD.HI) loses what
 * the steering gains (0.217 - 0.227) and is not kept. */
#ifndef SDRD_K4_FMA_ADD_MOD
#define SDRD_K4_FMA_ADD_MOD 16
#endif
struct ISteer {
    uint32_t zero, one;
};
SDRD_DEVICE uint32_t pre_add(uint32_t a, uint32_t b, int t, ISteer st)
{
    if (SDRD_K4_FMA_ADD_MOD < 0) return a + b;
    if (SDRD_K4_FMA_ADD_MOD > 0 && t % (SDRD_K4_FMA_ADD_MOD > 0 ? SDRD_K4_FMA_ADD_MOD : 1) == 0) return mad_lo(a, st.one, b);
    return add3(a, b, st.zero);
}

/* RAW8 (the last stage, whose samples only live on as int16): the taps are taken times 8 and the sums handed over
 * unshifted -- bits 16..31 of the wrapping sum 8 * acc are bits 13..28 of acc, i.e. (int16)(acc >> 13), which the
 * packing byte-permute picks up directly: no shift instruction for half of all the cascade's outputs. */
#ifndef SDRD_K4_RAW8
#define SDRD_K4_RAW8 1
#endif
#ifndef SDRD_K4_PACKED_Q_PLAIN
#define SDRD_K4_PACKED_Q_PLAIN 1
#endif
template <int L, int N, bool PACKED = false, bool RAW8 = false, class Sink>
SDRD_DEVICE void fir_steps(const int2* SDRD_RESTRICT buf, int u0, ISteer steer, Sink&& sink)
{
    constexpr int SW = w_sw_of_n(N);
    constexpr int T = L / 2;
    constexpr int C64[16] = SDRD_HB64_ITAPS;
    constexpr int C32[8] = SDRD_HB32_ITAPS;
    constexpr int C16[4] = SDRD_HB16_ITAPS;
    constexpr int NU = (L + N) / 2; /* units in the window */
    int2 win[L + N];
    if (PACKED) {
        /* entry e of buffer 0 is the 32-bit word e: the window is NU 8-byte loads, 8 bytes apart from lane to lane */
        const uint2* rw = reinterpret_cast<const uint2*>(reinterpret_cast<const uint32_t*>(buf) + 2 * u0);
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const uint2 r = rw[j];
            win[2 * j] = make_int2((int)(int16_t)(r.x & 0xFFFFu), ((int)r.x) >> 16);
            win[2 * j + 1] = make_int2((int)(int16_t)(r.y & 0xFFFFu), ((int)r.y) >> 16);
        }
    } else if (SW == 7) {
        /* u0 = 4 + 8 r: units 4..7 of row r, then rows r + 1, ..; unit low of row r sits at 8 r + (low ^ (r & 7)) */
#pragma unroll
        for (int j = 0; j < NU; j++) {
            const int row = (u0 >> 3) + ((4 + j) >> 3);
            const int b0 = (row << 3) | (row & 7);
            ld_unit(buf, b0 ^ ((4 + j) & 7), &win[2 * j]);
        }
    } else if (SW == 3) {
        /* chunks of 4 units share one swizzle value: unit 4 hc + i sits at 4 hc + (i ^ g), g = (hc >> 1) & 3 */
#pragma unroll
        for (int c = 0; c < NU / 4; c++) {
            const int hc = (u0 >> 2) + c;
            const int b0 = 4 * hc + ((hc >> 1) & 3);
#pragma unroll
            for (int i = 0; i < 4; i++) ld_unit(buf, b0 ^ i, &win[8 * c + 2 * i]);
        }
    } else if (SW == 1) {
#pragma unroll
        for (int c = 0; c < NU / 2; c++) {
            const int pr = (u0 >> 1) + c;
            const int b0 = 2 * pr + ((pr >> 2) & 1);
            ld_unit(buf, b0, &win[4 * c]);
            ld_unit(buf, b0 ^ 1, &win[4 * c + 2]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NU; j++) ld_unit(buf, u0 + j, &win[2 * j]);
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t ia = 0, qa = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int c = (L == 32 ? C64[t & 15] : L == 16 ? C32[t & 7] : C16[t & 3]) * (RAW8 ? 8 : 1);
            const int2 a = win[i + 1 + t], b = win[i + L - t]; /* x[k+i-L+1+t], x[k+i-t] */
            ia += pre_add((uint32_t)a.x, (uint32_t)b.x, t, steer) * (uint32_t)c;
            /* packed input: the Q of a raw word is its high half, and ptxas folds that shift into the pre-add
             * (LEA.HI.SX32, an ALU instruction already) -- a third operand would only add an instruction */
            qa += ((PACKED && SDRD_K4_PACKED_Q_PLAIN) ? (uint32_t)a.y + (uint32_t)b.y : pre_add((uint32_t)a.y, (uint32_t)b.y, t, steer)) * (uint32_t)c;
        }
        sink(i, win[i + L - L / 2], RAW8 ? make_int2((int)ia, (int)qa) : make_int2(asr32(ia, 13), asr32(qa, 13)));
    }
}

/* {int16 I, int16 Q} in one byte-permute: the low halves, or the high halves of sums that were left unshifted (RAW8) */
SDRD_DEVICE uint32_t pack16p(int2 v) { return prmt((uint32_t)v.x, (uint32_t)v.y, 0x5410u); }
SDRD_DEVICE uint32_t pack16hi(int2 v) { return prmt((uint32_t)v.x, (uint32_t)v.y, 0x7632u); }

struct WarpParams {
    const uint32_t* in;   /* stream s, sample k (k >= -HIST): in[s * in_stride + k] */
    long long in_stride;  /* words */
    uint32_t* out;        /* out[s * out_stride + n] */
    long long out_stride;
    long long n_in;       /* input samples per stream */
    int log2_interp;      /* 1..6; stages run S = min(log2_interp, 5) */
    int steps_per_warp;   /* steps of WC input samples per warp; warp blockIdx.x takes steps [x * spw, (x + 1) * spw) */
    /* TMA tensor store of the last stage (16 steps per lane: x16, x32): the output buffer as rows of 32 words; a pass of
     * the last stage is one 32 x 32-word tile = 1024 consecutive samples.  use_tma = 0: plain stores. */
    int use_tma;
    uint32_t steer_zero, steer_one; /* 0 and 1 the compiler cannot fold: see ISteer */
    long long out_word0;  /* word index of out[0] inside the array the map describes */
    TileMap tmap;
};

/* one stage of the step: reads buffer ST - 1, writes buffer ST (ST < S) or the output (ST == S) */
template <int S, int ST>
SDRD_DEVICE void warp_stage(int2* SDRD_RESTRICT buf, uint32_t* SDRD_RESTRICT stage, int lane, uint32_t* SDRD_RESTRICT out_step, int n_left,
                            int wo, bool emit, const TileMap* tmap, long long tile_row0, ISteer steer)
{
    constexpr int L = ring_len(ST);
    constexpr int N = w_nstep(ST);
    constexpr bool PK = ST == 1 && SDRD_K4_PACKED_X0;
    const int2* src = buf + w_off(ST - 1);
#pragma unroll 1
    for (int pass = 0; pass < w_passes(ST); pass++) {
        const int task = 32 * pass + lane;
        const int u0 = w_woff(ST - 1) + (N / 2) * task;
        if (ST < S) {
            constexpr int SWD = w_sw(ST);
            int2* dst = buf + w_off(ST);
            const int ub = w_hist(ST) / 2 + N * task; /* first unit of the task's 2 N new entries: one unit per step */
            if (SWD == 7 && N == 8) { /* one whole swizzle row */
                const int row = ub >> 3;
                const int b0 = (row << 3) | (row & 7);
                fir_steps<L, N, PK>(src, u0, steer, [&](int i, int2 ev, int2 od) { st_unit(dst, b0 ^ i, ev, od); });
            } else if (N == 4 && SWD == 3) { /* one chunk of 4 units */
                const int q = ub >> 2;
                const int b0 = 4 * q + ((q >> 1) & 3);
                fir_steps<L, N, PK>(src, u0, steer, [&](int i, int2 ev, int2 od) { st_unit(dst, b0 ^ i, ev, od); });
            } else if (N == 2 && SWD == 1) { /* one pair of units */
                const int q = ub >> 1;
                const int b0 = 2 * q + ((q >> 2) & 1);
                fir_steps<L, N, PK>(src, u0, steer, [&](int i, int2 ev, int2 od) { st_unit(dst, b0 ^ i, ev, od); });
            } else {
                fir_steps<L, N, PK>(src, u0, steer, [&](int i, int2 ev, int2 od) { st_unit(dst, swz<SWD>(ub + i), ev, od); });
            }
        } else {
            /* last stage: pack to int16 pairs (IQSample::setReal/setImag), through the staging area, whole rows out */
            constexpr int SWO = w_sw_of_n(N);
            constexpr int UL = N / 2; /* 16-byte units per lane */
            uint32_t wd[2 * N];
            fir_steps<L, N, PK, SDRD_K4_RAW8 != 0>(src, u0, steer, [&](int i, int2 ev, int2 od) {
                wd[2 * i] = pack16p(ev);
                wd[2 * i + 1] = SDRD_K4_RAW8 ? pack16hi(od) : pack16p(od);
            });
            /* A whole pass of 16 steps per lane is one 32 x 32-word tile of consecutive output samples: written to the
             * staging area in the 128-byte swizzle (which is the conflict-free layout the read-back form uses anyway) and
             * sent with ONE TMA tensor store -- no read-back, no STG, nothing of it on the LSU data pipe. */
            if (N == 16 && tmap != nullptr && wo == S && (64 * N) * (pass + 1) <= n_left) {
                if (lane == 0) tma_store_wait_read(); /* the previous tile has been read out of the staging area */
                SDRD_SYNCWARP();
                uint4* sgt = reinterpret_cast<uint4*>(stage);
#pragma unroll
                for (int i = 0; i < UL; i++) sgt[swz<SWO>(UL * lane + i)] = make_uint4(wd[4 * i], wd[4 * i + 1], wd[4 * i + 2], wd[4 * i + 3]);
                fence_proxy_async_smem();
                SDRD_SYNCWARP();
                if (lane == 0 && emit) {
                    tma_store_tile32(tmap, stage, (int)(tile_row0 + 32 * pass));
                    tma_store_commit();
                }
                continue;
            }
            if (N == 16 && tmap != nullptr) { /* a ragged last pass after tensor stores: the staging area may still be read */
                if (lane == 0) tma_store_wait_read();
                SDRD_SYNCWARP();
            }
            uint4* sg = reinterpret_cast<uint4*>(stage);
#pragma unroll
            for (int i = 0; i < UL; i++) sg[swz<SWO>(UL * lane + i)] = make_uint4(wd[4 * i], wd[4 * i + 1], wd[4 * i + 2], wd[4 * i + 3]);
            SDRD_SYNCWARP();
            uint4 v[UL];
#pragma unroll
            for (int j = 0; j < UL; j++) v[j] = sg[swz<SWO>(32 * j + lane)];
            SDRD_SYNCWARP();
            if (emit) {
#pragma unroll
                for (int j = 0; j < UL; j++) {
                    /* sample n of the cascade is emitted at (n >> S << wo) + (n & (2^S - 1)); here relative to the
                     * step's first sample (out_step, n_left: 32-bit offsets) */
                    const int n = (64 * N) * pass + 4 * (32 * j + lane);
                    if (S >= 2) {
                        if (n + 4 <= n_left) *reinterpret_cast<uint4*>(out_step + ((n >> S) << wo) + (n & ((1 << S) - 1))) = v[j];
                    } else { /* S = 1: two input samples' pairs */
                        if (n + 2 <= n_left) *reinterpret_cast<uint2*>(out_step + ((n >> 1) << wo)) = make_uint2(v[j].x, v[j].y);
                        if (n + 4 <= n_left) *reinterpret_cast<uint2*>(out_step + (((n + 2) >> 1) << wo)) = make_uint2(v[j].z, v[j].w);
                    }
                }
            }
        }
    }
}

#ifndef SDRD_K4_WARPS_PER_SM
#define SDRD_K4_WARPS_PER_SM 16 /* x16, ms: before the pipe steering 8 / 12 / 16 warps = 0.234 / 0.232 / 0.233 (2 per scheduler reached the rate);
                                 * with the pipes balanced 8 / 12 / 16 = 0.214 / 0.201 / 0.197 (113 registers at 16, no spills).  The launch
                                 * asks for what also fits in shared memory, in whole warps per scheduler: 16 at x16, 8 at x32 / x64. */
#endif

template <int S>
SDRD_KERNEL(32, SDRD_K4_WARPS_PER_SM) interpolate_warp_kernel(SDRD_GRID_CONSTANT WarpParams p)
{
    static_assert(S >= 1 && S <= 5, "1..5 stages");
    SDRD_DYN_SMEM(smem);
    int2* const buf = reinterpret_cast<int2*>(smem);
    /* the staging tile on a 1024-byte boundary of the shared-memory address space (TMA's 128-byte swizzle is a function
     * of the address) */
    unsigned char* const stage_raw = smem + (size_t)w_off(S) * 8;
    uint32_t* const stage = reinterpret_cast<uint32_t*>(stage_raw + ((1024u - (unsigned)(smem_addr(stage_raw) & 1023u)) & 1023u));
    const TileMap* const tmap = (p.use_tma && w_nstep(S) == 16) ? &p.tmap : nullptr;
    const int lane = (int)threadIdx.x;
    const ISteer steer = {p.steer_zero, p.steer_one};
    const uint32_t* in = p.in + (long long)blockIdx.y * p.in_stride;
    uint32_t* out = p.out + (long long)blockIdx.y * p.out_stride;
    const int wo = p.log2_interp;
    const long long steps_total = (p.n_in + WC - 1) / WC;
    const long long first = (long long)blockIdx.x * p.steps_per_warp;
    long long last = first + p.steps_per_warp;
    if (last > steps_total) last = steps_total;
    if (first >= last) return;
    const long long n_valid = p.n_in << S;

    /* the two input samples of this lane for a step: k = 64 step + 2 lane, + 1 (0 past the end of the stream) */
    auto fetch = [&](long long step) -> uint2 {
        const long long k = step * WC + 2 * lane;
        uint2 v = make_uint2(0u, 0u);
        if (k + 1 < p.n_in) v = *reinterpret_cast<const uint2*>(in + k);
        else if (k < p.n_in) v.x = in[k];
        return v;
    };
    uint2 nxt = fetch(first - 1);
#pragma unroll 1
    for (long long step = first - 1; step < last; step++) {
        const uint2 cur = nxt;
        if (step + 1 < last) nxt = fetch(step + 1);
        const bool emit = step >= first; /* the first step only fills the histories */
        if (SDRD_K4_PACKED_X0)
            reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(buf + w_off(0)) + w_hist(0))[lane] = cur;
        else
            st_unit(buf + w_off(0), w_hist(0) / 2 + lane, make_int2((int)(int16_t)(cur.x & 0xFFFFu), ((int)cur.x) >> 16),
                    make_int2((int)(int16_t)(cur.y & 0xFFFFu), ((int)cur.y) >> 16));
        SDRD_SYNCWARP();
        /* where the step's first output sample goes, and how many stage-S samples exist from there on */
        uint32_t* const out_step = out + ((step * WC) << wo);
        const long long left = n_valid - ((step * WC) << S);
        const int n_left = left > (WC << S) ? (WC << S) : (int)left;
        /* row (32 words) of the step's first output sample in the array the tensor map describes */
        const long long tile_row0 = (p.out_word0 + (long long)blockIdx.y * p.out_stride + ((step * WC) << wo)) >> 5;
        warp_stage<S, 1>(buf, stage, lane, out_step, n_left, wo, emit, tmap, tile_row0, steer);
        if (S >= 2) { SDRD_SYNCWARP(); warp_stage<S, S >= 2 ? 2 : 1>(buf, stage, lane, out_step, n_left, wo, emit, tmap, tile_row0, steer); }
        if (S >= 3) { SDRD_SYNCWARP(); warp_stage<S, S >= 3 ? 3 : 1>(buf, stage, lane, out_step, n_left, wo, emit, tmap, tile_row0, steer); }
        if (S >= 4) { SDRD_SYNCWARP(); warp_stage<S, S >= 4 ? 4 : 1>(buf, stage, lane, out_step, n_left, wo, emit, tmap, tile_row0, steer); }
        if (S >= 5) { SDRD_SYNCWARP(); warp_stage<S, S >= 5 ? 5 : 1>(buf, stage, lane, out_step, n_left, wo, emit, tmap, tile_row0, steer); }
        if (wo > S && emit) { /* interpolate64_cen: 32 zero samples after every 32 (Interpolators.cpp:370,413-603) */
            const int zq = ((1 << wo) - (1 << S)) / 4; /* zero uint4 per input sample */
            for (int i = lane; i < WC * zq; i += 32) {
                const long long k = step * WC + i / zq;
                if (k < p.n_in) *reinterpret_cast<uint4*>(out + (k << wo) + (1 << S) + 4 * (i % zq)) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        /* histories: the last w_hist(b) entries of every buffer move to its front (16, 8, 4, 4..8 units) */
        SDRD_SYNCWARP();
        {
            int4 hv[5];
#define SDRD_K4_TL(B, OP)                                                                                      \
            if ((B) < S && lane < w_hist((B) < S ? (B) : 0) / 2) {                                             \
                constexpr int BB = (B) < S ? (B) : 0;                                                          \
                constexpr int SWB = w_sw(BB);                                                                  \
                int2* bb = buf + w_off(BB);                                                                    \
                const int su = (WC << BB) / 2 + lane;                                                          \
                OP                                                                                             \
            }
            if (SDRD_K4_PACKED_X0) { /* 32 history words = 8 units of raw words */
                if (lane < w_hist(0) / 4) hv[0] = reinterpret_cast<const int4*>(reinterpret_cast<const uint32_t*>(buf + w_off(0)) + WC)[lane];
            } else {
                SDRD_K4_TL(0, hv[0] = *reinterpret_cast<const int4*>(bb + 2 * swz<SWB>(su));)
            }
            SDRD_K4_TL(1, hv[1] = *reinterpret_cast<const int4*>(bb + 2 * swz<SWB>(su));)
            SDRD_K4_TL(2, hv[2] = *reinterpret_cast<const int4*>(bb + 2 * swz<SWB>(su));)
            SDRD_K4_TL(3, hv[3] = *reinterpret_cast<const int4*>(bb + 2 * swz<SWB>(su));)
            SDRD_K4_TL(4, hv[4] = *reinterpret_cast<const int4*>(bb + 2 * swz<SWB>(su));)
            SDRD_SYNCWARP();
            if (SDRD_K4_PACKED_X0) {
                if (lane < w_hist(0) / 4) reinterpret_cast<int4*>(buf + w_off(0))[lane] = hv[0];
            } else {
                SDRD_K4_TL(0, (void)su; *reinterpret_cast<int4*>(bb + 2 * swz<SWB>(lane)) = hv[0];)
            }
            SDRD_K4_TL(1, (void)su; *reinterpret_cast<int4*>(bb + 2 * swz<SWB>(lane)) = hv[1];)
            SDRD_K4_TL(2, (void)su; *reinterpret_cast<int4*>(bb + 2 * swz<SWB>(lane)) = hv[2];)
            SDRD_K4_TL(3, (void)su; *reinterpret_cast<int4*>(bb + 2 * swz<SWB>(lane)) = hv[3];)
            SDRD_K4_TL(4, (void)su; *reinterpret_cast<int4*>(bb + 2 * swz<SWB>(lane)) = hv[4];)
#undef SDRD_K4_TL
        }
        SDRD_SYNCWARP();
    }
    if (tmap != nullptr && lane == 0) tma_store_wait_read(); /* the staging tile stays allocated until the last store has read it */
}

} /* namespace hbi */
} /* namespace sdrd */
