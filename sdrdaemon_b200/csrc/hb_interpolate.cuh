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
 * shared memory as int32 {I, Q} pairs (one thread per output pair, coalesced 8-byte shared loads),
 * and the last stage packs and writes its 4096 output samples straight to global memory with 8-byte
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
 * rounded up to whole output pairs */
SDRD_HD constexpr int halo(int s, int S)
{
    int h = 0;
    for (int t = S; t > s; t--) h = (((h + 1) >> 1) + ring_len(t) - 1 + 1) & ~1;
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

template <int L>
SDRD_DEVICE int2 fir_pair(const int2* SDRD_RESTRICT x /* points at x[k] */)
{
    constexpr int T = L / 2;
    constexpr int C64[16] = SDRD_HB64_ITAPS;
    constexpr int C32[8] = SDRD_HB32_ITAPS;
    constexpr int C16[4] = SDRD_HB16_ITAPS;
    uint32_t ia = 0, qa = 0;
#pragma unroll
    for (int i = 0; i < T; i++) {
        const int c = L == 32 ? C64[i & 15] : L == 16 ? C32[i & 7] : C16[i & 3];
        const int2 a = x[-(L - 1) + i], b = x[-i];
        ia += ((uint32_t)a.x + (uint32_t)b.x) * (uint32_t)c;
        qa += ((uint32_t)a.y + (uint32_t)b.y) * (uint32_t)c;
    }
    return make_int2(asr32(ia, 13), asr32(qa, 13));
}

/* stage s: src holds x_{s-1}[src0 ...], dst receives x_s[dst0 ... dst0 + 2 * npairs) (dst0 even) */
template <int L>
SDRD_DEVICE void run_stage(const int2* SDRD_RESTRICT src, long long src0, int2* SDRD_RESTRICT dst, long long dst0, int npairs, int tid)
{
    for (int q = tid; q < npairs; q += NT) {
        const long long k = (dst0 >> 1) + q;
        const int2* x = src + (k - src0);
        const int2 ev = x[-(L / 2)];
        const int2 od = fir_pair<L>(x);
        *reinterpret_cast<int4*>(dst + 2 * q) = make_int4(ev.x, ev.y, od.x, od.y);
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
                               (k0 << (t)) - hs, (buf_len((t), S)) >> 1, tid);                                      \
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
        for (int q = tid; q < (K << S) / 2; q += NT) {
            const long long k = (k0 << (S - 1)) + q;
            const long long n = 2 * k; /* stage-S index of the pair */
            if (n >= n_valid) break;
            const int2* x = src + (k - src0);
            const uint32_t ev = pack16(x[-(L / 2)]), od = pack16(fir_pair<L>(x));
            /* sample n of the cascade is emitted at (n >> S << wo) + (n & (2^S - 1)) */
            const long long pos = ((n >> S) << wo) + (n & ((1 << S) - 1));
            *reinterpret_cast<uint2*>(out + pos) = make_uint2(ev, od);
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
