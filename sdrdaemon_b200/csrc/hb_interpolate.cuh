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
 * Design: the cascade's receptive field is short (42 input samples of history for five stages), so a
 * CTA takes a tile of K = 4096 >> S input samples plus that halo, runs stage after stage through
 * shared memory as int32 {I, Q} pairs (two output pairs per thread from one register window,
 * conflict-free 16-byte shared loads),
 * and the last stage packs and writes its 4096 output samples straight to global memory with 16-byte
 * stores.  Arithmetic per OUTPUT sample is small (16 S / 2^S multiply-adds per component), the kernel
 * is bound by the 4 + 4 / 2^S bytes it moves per output sample and by shared-memory reads.
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
/* samples of stage s needed in front of a tile's first sample so that stages s+1..S can be computed,
 * rounded up to a multiple of 4 (two output pairs per thread, 16-byte aligned window loads) */
SDRD_HD constexpr int halo(int s, int S)
{
    int h = 0;
    for (int t = S; t > s; t--) h = (((h + 1) >> 1) + ring_len(t) - 1 + 3) & ~3;
    return h;
}
SDRD_HD constexpr int tile_in(int S) { return 4096 >> S; }
/* int2 entries of the buffer holding stage s (0 = unpacked input) for one tile */
SDRD_HD constexpr int buf_len(int s, int S) { return (tile_in(S) << s) + halo(s, S); }
SDRD_HD constexpr int buf_off(int s, int S)
{
    int o = 0;
    for (int t = 0; t < s; t++) o += (buf_len(t, S) + 1) & ~1;
    return o;
}
SDRD_HD constexpr size_t smem_bytes(int S) { return (size_t)buf_off(S, S) * 8; }

struct Params {
    const uint32_t* in;   /* stream s, sample k (k >= -HIST): in[s * in_stride + k] */
    long long in_stride;  /* words */
    uint32_t* out;        /* out[s * out_stride + n] */
    long long out_stride;
    long long n_in;       /* input samples per stream */
    int log2_interp;      /* 1..6; stages run S = min(log2_interp, 5) */
};

/* Two consecutive steps k, k + 1 (k even) of one stage from a register window win[j] = x[k - L + j],
 * j = 0 .. L + 1, fetched with (L + 2) / 2 conflict-free 16-byte shared loads (a thread's window starts
 * 16 bytes after its neighbour's):  ev[i] = x[k + i - L/2],  od[i] = FIR at step k + i. */
template <int L>
SDRD_DEVICE void fir_two(const int2* SDRD_RESTRICT x /* points at x[k - L], 16-byte aligned */, int2 (&ev)[2], int2 (&od)[2])
{
    constexpr int T = L / 2;
    constexpr int C64[16] = SDRD_HB64_ITAPS;
    constexpr int C32[8] = SDRD_HB32_ITAPS;
    constexpr int C16[4] = SDRD_HB16_ITAPS;
    int2 win[L + 2];
#pragma unroll
    for (int j = 0; j < (L + 2) / 2; j++) {
        const int4 v = reinterpret_cast<const int4*>(x)[j];
        win[2 * j] = make_int2(v.x, v.y);
        win[2 * j + 1] = make_int2(v.z, v.w);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        uint32_t ia = 0, qa = 0;
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int c = L == 32 ? C64[t & 15] : L == 16 ? C32[t & 7] : C16[t & 3];
            const int2 a = win[i + 1 + t], b = win[i + L - t]; /* x[k+i-L+1+t], x[k+i-t] */
            ia += ((uint32_t)a.x + (uint32_t)b.x) * (uint32_t)c;
            qa += ((uint32_t)a.y + (uint32_t)b.y) * (uint32_t)c;
        }
        od[i] = make_int2(asr32(ia, 13), asr32(qa, 13));
        ev[i] = win[i + L - L / 2];
    }
}

/* stage s: src holds x_{s-1}[src0 ...], dst receives x_s[dst0 ... dst0 + 4 * nquads) (dst0 a multiple of 4,
 * src0 even) */
template <int L>
SDRD_DEVICE void run_stage(const int2* SDRD_RESTRICT src, long long src0, int2* SDRD_RESTRICT dst, long long dst0, int nquads, int tid)
{
    for (int q = tid; q < nquads; q += NT) {
        const long long k = (dst0 >> 1) + 2 * q;
        int2 ev[2], od[2];
        fir_two<L>(src + (k - L - src0), ev, od);
        *reinterpret_cast<int4*>(dst + 4 * q) = make_int4(ev[0].x, ev[0].y, od[0].x, od[0].y);
        *reinterpret_cast<int4*>(dst + 4 * q + 2) = make_int4(ev[1].x, ev[1].y, od[1].x, od[1].y);
    }
}

SDRD_DEVICE uint32_t pack16(int2 v) { return ((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16); }

template <int S>
SDRD_KERNEL(NT, 2) interpolate_kernel(Params p)
{
    static_assert(S >= 1 && S <= 5, "1..5 stages");
    SDRD_DYN_SMEM(smem);
    int2* const buf = reinterpret_cast<int2*>(smem);
    const int tid = (int)threadIdx.x;
    const int s = (int)blockIdx.y;
    constexpr int K = tile_in(S);
    const long long k0 = (long long)blockIdx.x * K; /* first input sample of the tile */
    const uint32_t* in = p.in + (long long)s * p.in_stride;
    uint32_t* out = p.out + (long long)s * p.out_stride;
    const int wo = p.log2_interp; /* log2 of the samples emitted per input sample */

    /* stage 0: unpack x_0[k0 - halo .. k0 + K); samples past n_in read as 0 (their outputs are not stored) */
    {
        constexpr int h0 = halo(0, S);
        for (int i = tid; i < K + h0; i += NT) {
            const long long k = k0 - h0 + i;
            const uint32_t v = k < p.n_in ? in[k] : 0u;
            buf[buf_off(0, S) + i] = make_int2((int)(int16_t)(v & 0xFFFFu), ((int)v) >> 16);
        }
    }
    __syncthreads();

    /* stages 1 .. S-1 through shared memory */
#define SDRD_HBI_STAGE(t)                                                                                            \
    if (S > (t)) {                                                                                                   \
        constexpr int hs = halo((t), S), hp = halo((t) - 1, S);                                                      \
        run_stage<ring_len(t)>(buf + buf_off((t) - 1, S), (k0 << ((t) - 1)) - hp, buf + buf_off((t), S),           \
                               (k0 << (t)) - hs, (buf_len((t), S)) >> 2, tid);                                      \
        __syncthreads();                                                                                             \
    }
    SDRD_HBI_STAGE(1)
    SDRD_HBI_STAGE(2)
    SDRD_HBI_STAGE(3)
    SDRD_HBI_STAGE(4)
#undef SDRD_HBI_STAGE

    /* stage S: straight to global memory, truncated to int16 (IQSample::setReal/setImag) */
    {
        constexpr int L = ring_len(S);
        constexpr int hp = halo(S - 1, S);
        const int2* src = buf + buf_off(S - 1, S);
        const long long src0 = (k0 << (S - 1)) - hp;
        const long long n_valid = p.n_in << S; /* stage-S samples that exist */
        for (int q = tid; q < (K << S) / 4; q += NT) {
            const long long k = (k0 << (S - 1)) + 2 * q;
            const long long n = 2 * k; /* stage-S index of the first of the four samples */
            if (n >= n_valid) break;
            int2 ev[2], od[2];
            fir_two<L>(src + (k - L - src0), ev, od);
            /* sample n of the cascade is emitted at (n >> S << wo) + (n & (2^S - 1)); S >= 2: the four stay together */
            if (S >= 2) {
                const long long pos = ((n >> S) << wo) + (n & ((1 << S) - 1));
                *reinterpret_cast<uint4*>(out + pos) = make_uint4(pack16(ev[0]), pack16(od[0]), pack16(ev[1]), pack16(od[1]));
            } else {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const long long ni = n + 2 * i;
                    if (ni < n_valid) {
                        const long long pos = ((ni >> S) << wo) + (ni & ((1 << S) - 1));
                        *reinterpret_cast<uint2*>(out + pos) = make_uint2(pack16(ev[i]), pack16(od[i]));
                    }
                }
            }
        }
        if (wo > S) { /* interpolate64_cen: 32 zero samples after every 32 (Interpolators.cpp:370,413-603) */
            const int zw = (1 << wo) - (1 << S); /* zero words per input sample */
            for (int i = tid; i < K * (zw / 4); i += NT) {
                const long long k = k0 + i / (zw / 4);
                if (k >= p.n_in) break;
                *reinterpret_cast<uint4*>(out + (k << wo) + (1 << S) + 4 * (i % (zw / 4))) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    }
}

} /* namespace hbi */
} /* namespace sdrd */
