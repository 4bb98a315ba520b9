/*
 * fec_kernels.cuh -- K2 (superframe framing + CM256 encode) and K3 (CM256 recover) for sdrdaemon's
 * 128-block superframes.
 *
 *   K2 replaces UDPSinkFEC::write (sdmnbase/UDPSinkFEC.cpp:79-191: 127 samples per block, block 0 =
 *      meta data, 128 blocks per frame) and CM256::cm256_encode as called from
 *      UDPSinkFEC::transmitUDP (:228-256).
 *   K3 replaces the store / cm256_decode / copy-back part of SDRdaemonFECBuffer::writeAndRead
 *      (sdmnbase/SDRdaemonFECBuffer.cpp:143-213).
 *
 * Both are the same computation: R output blocks, each a GF(2^8)-linear combination of the 128
 * blocks of the frame,  out[r] = XOR_j C[r][j] (x) blk[j]  over 508 payload bytes.  For the encoder C
 * is the fixed Cauchy matrix; for the decoder C = [A^-1 | A^-1 M] is built per frame from the
 * erasure pattern (A = Cauchy sub-matrix of the received recovery rows x erased columns, inverted in
 * closed form).
 *
 * One CTA owns one superframe, kept in shared memory as 128 datagram images of 128 words (header
 * word + 127 payload words).  A constant-times-block product works on 4 packed bytes at a time:
 * each byte is split 3+3+2 bits, the three partial products come from 8-entry byte tables held in
 * registers and are looked up with the byte-permute unit (PRMT), 3 PRMT + 2 LOP3 per word.
 * Warp w handles columns 16w..16w+15 for a block of 16 output rows (64 accumulator registers per
 * thread: 16 rows x 4 words) and the 8 partial sums meet in shared memory through XOR atomics.
 * No tensor cores: this is byte-wise table arithmetic, bound by the ALU/PRMT issue rate.
 */
#pragma once
#include "sdrd_platform.cuh"

namespace sdrd {
namespace fec {

constexpr int NT = 256;
constexpr int ROW_WORDS = 128;               /* one 512-byte datagram */
constexpr int IMG_WORDS = 128 * ROW_WORDS;   /* 128 datagrams */
constexpr int FRAME_SAMPLES = 127 * 127;
constexpr int RB = 16;                       /* output rows per pass */

struct Tables {
    const uint4* tabA;      /* [256] */
    const uint32_t* tabB;   /* [256] */
    const uint8_t* cauchy;  /* [128][128] encoder matrix */
    const uint8_t* gfexp;   /* [512] */
    const uint8_t* gflog;   /* [256] */
};

/* Nibble selectors for PRMT from the 3+3+2-bit digits of four packed bytes: selector nibble i = digit of
 * byte i.  The ALU pipe (PRMT, LOP3, SHF) is what bounds these kernels and the FMA pipe idles, so the shifts
 * that bring two digits into one byte are done by an integer multiply: the masked word times (2^a + 2^b)
 * puts digit i - 1 into the low and digit i into the high nibble of byte i (no two fields overlap, so the
 * sum has no carries); for the top digit the high word of the product (IMAD.HI) does the right shift.  One
 * PRMT then moves the two useful bytes into the low half.  2 ALU + 1 FMA instruction per digit (was 4-5 ALU). */
SDRD_DEVICE void selectors(uint32_t x, uint32_t& s0, uint32_t& s1, uint32_t& s2)
{
    s0 = prmt(mad_lo(x & 0x07070707u, (1u << 4) + (1u << 8), 0u), 0u, 0x0031u);   /* bits 0-2: bytes 1, 3 of the product */
    s1 = prmt(mad_lo(x & 0x38383838u, (1u << 1) + (1u << 5), 0u), 0u, 0x0031u);   /* bits 3-5 */
    s2 = prmt(mul_hi(x & 0xC0C0C0C0u, (1u << 26) + (1u << 22)), 0u, 0x0020u);     /* bits 6-7: bytes 0, 2 of the high word */
}
/* One pass: rows [row0, row0 + nrows) (nrows <= 16) of  out = C * img  accumulated into rec16
 * (16 x 128 words, zeroed by the caller).  coefT[j * cstride + r] = 32 * C[r][j]: the byte offset of
 * the multiplier's entry in `tab` ({8-entry tables for digits 0 and 1: 16 bytes, digit 2: 4 bytes,
 * padding}), fetched with a 16-bit shared load so that no ALU instruction is spent on it -- the ALU
 * pipe (PRMT, LOP3) is what bounds this kernel.  Two columns are folded into an accumulator with three
 * 3-input XORs. */
template <bool FULL> /* FULL: nrows == RB, the row loop carries no predicates */
SDRD_DEVICE void matvec_pass_t(const uint32_t* SDRD_RESTRICT img, const uint16_t* SDRD_RESTRICT coefT, int cstride,
                               int row0, int nrows, const unsigned char* SDRD_RESTRICT tab, uint32_t* rec16, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t acc[RB][4];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
        for (int w = 0; w < 4; w++) acc[r][w] = 0u;

    for (int jj = 0; jj < 16; jj += 2) {
        const int j = warp * 16 + jj;
        uint32_t s0[2][4], s1[2][4], s2[2][4];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int w = 0; w < 4; w++) selectors(img[(j + c) * ROW_WORDS + lane + 32 * w], s0[c][w], s1[c][w], s2[c][w]);
        const uint16_t* co0 = coefT + j * cstride + row0;
        const uint16_t* co1 = co0 + cstride;
#pragma unroll
        for (int r = 0; r < RB; r++) {
            if (FULL || r < nrows) {
                const unsigned char* e0 = tab + co0[r];
                const unsigned char* e1 = tab + co1[r];
                const uint4 a0 = *reinterpret_cast<const uint4*>(e0);
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(e0 + 16);
                const uint4 a1 = *reinterpret_cast<const uint4*>(e1);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(e1 + 16);
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    uint32_t v = acc[r][w];
                    v = v ^ prmt(a0.x, a0.y, s0[0][w]) ^ prmt(a0.z, a0.w, s1[0][w]);
                    v = v ^ prmt(b0, b0, s2[0][w]) ^ prmt(a1.x, a1.y, s0[1][w]);
                    v = v ^ prmt(a1.z, a1.w, s1[1][w]) ^ prmt(b1, b1, s2[1][w]);
                    acc[r][w] = v;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (FULL || r < nrows) {
#pragma unroll
            for (int w = 0; w < 4; w++) atomicXor(&rec16[r * ROW_WORDS + lane + 32 * w], acc[r][w]);
        }
    }
}

SDRD_DEVICE void matvec_pass(const uint32_t* SDRD_RESTRICT img, const uint16_t* SDRD_RESTRICT coefT, int cstride,
                             int row0, int nrows, const unsigned char* SDRD_RESTRICT tab, uint32_t* rec16, int tid)
{
    if (nrows == RB) matvec_pass_t<true>(img, coefT, cstride, row0, nrows, tab, rec16, tid);
    else matvec_pass_t<false>(img, coefT, cstride, row0, nrows, tab, rec16, tid);
}

/* shared-memory carve-up common to both kernels */
struct Smem {
    uint32_t* img;    /* [128][128] */
    uint32_t* rec16;  /* [16][128] */
    unsigned char* tab; /* [256] entries of 32 bytes: tabA[c] (16), tabB[c] (4), padding */
    uint16_t* coefT;    /* [128][cstride], 32 * coefficient */
    uint8_t* extra;
};
constexpr int TAB_ENTRY = 32;
constexpr size_t SMEM_FIXED = (size_t)IMG_WORDS * 4 + (size_t)RB * ROW_WORDS * 4 + 256 * TAB_ENTRY;
SDRD_DEVICE Smem carve(unsigned char* base, int cstride)
{
    Smem s;
    s.img = reinterpret_cast<uint32_t*>(base);
    s.rec16 = s.img + IMG_WORDS;
    s.tab = reinterpret_cast<unsigned char*>(s.rec16 + RB * ROW_WORDS);
    s.coefT = reinterpret_cast<uint16_t*>(s.tab + 256 * TAB_ENTRY);
    s.extra = reinterpret_cast<uint8_t*>(s.coefT + 128 * cstride);
    return s;
}
/* CRC-32/IEEE (boost::crc_32_type, UDPSinkFEC.cpp:106-109) over the 20 meta bytes held in five words */
SDRD_DEVICE uint32_t crc32_meta(const uint32_t* w5)
{
    uint32_t c = 0xFFFFFFFFu;
    for (int i = 0; i < 20; i++) {
        c ^= (w5[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
    }
    return c ^ 0xFFFFFFFFu;
}

SDRD_DEVICE void load_tables(const Smem& s, const Tables& t, int tid)
{
    for (int i = tid; i < 256; i += NT) {
        *reinterpret_cast<uint4*>(s.tab + i * TAB_ENTRY) = t.tabA[i];
        *reinterpret_cast<uint32_t*>(s.tab + i * TAB_ENTRY + 16) = t.tabB[i];
    }
}

/* ============================================================================ encode ==== */

struct EncParams {
    int mode;  /* 0: frame mode (samples -> datagram images), 1: raw mode (originals -> recovery blocks) */
    int F;     /* recovery blocks per frame, 0..128 */
    int cstride; /* F rounded up to a multiple of 16 (>= 16) */
    /* frame mode */
    const uint32_t* samples;   /* stream s: samples + s * sample_stride */
    long long sample_stride;
    const uint32_t* pending;   /* [S][FRAME_SAMPLES]: samples carried over from earlier calls */
    int n_pending;
    uint32_t meta_first[6];    /* block-0 payload (24 bytes) of a frame begun in an earlier call */
    uint32_t meta_next[6];     /* block-0 payload of frames begun in this call */
    /* per-frame time stamps: the reference reads the clock when the first sample of a frame is written
     * (UDPSinkFEC.cpp:89-95).  With stamp_rate != 0 a frame begun `o` samples into this call is stamped
     * meta_next's time + o / stamp_rate (whole microseconds) and its CRC-32 is computed here. */
    uint32_t stamp_rate;
    unsigned frame_index0;     /* frame counter of the first frame completed by this call */
    uint32_t* dgrams;          /* stream s, frame f: dgrams + s * dgram_stride + f * (128 + F) * 128 */
    long long dgram_stride;    /* words */
    /* raw mode */
    const uint8_t* originals;  /* frame f block j: originals + (f * 128 + j) * block_pitch */
    long long block_pitch;
    uint8_t* recovery;         /* frame f row r: recovery + (f * F + r) * 508 */
    int n_frames, n_streams;   /* work items = n_streams * n_frames, spread over a persistent grid */
    Tables tab;
};

/* Persistent encode kernel, two shapes (template parameter TWO):
 *   false  one 512-thread CTA per SM with two datagram images: the next superframe is gathered while this one is
 *          encoded.  A single frame gets the whole SM: what small batches (the reference's call granularity) need.
 *   true   two 256-thread CTAs per SM with one image each: the gather, the barriers and the stores of one run under the
 *          products of the other.  5 % faster on batches of more than one frame per SM (config 2: 64.5 -> 61.0 us);
 *          a lone frame takes twice as long. */
template <bool TWO>
struct EncShape {
    static constexpr int NT = TWO ? 256 : 512;
    static constexpr int CTAS = TWO ? 2 : 1;
    static constexpr int CPW = 128 / (NT / 32); /* columns per warp and pass */
};
inline size_t enc_smem_bytes(int cstride, bool two = false)
{
    return SMEM_FIXED + (two ? 0 : (size_t)IMG_WORDS * 4) + (size_t)128 * cstride * 4;
}
/* what the two-CTA shape needs twice per SM (227 KB, 1 KB per CTA reserved) */
inline bool enc_two_fits(int cstride) { return 2 * (enc_smem_bytes(cstride, true) + 1024) <= (size_t)227 * 1024; }

/* warp w takes columns CPW w .. CPW (w + 1) - 1 of a pass (16 warps: 8 each) */
template <bool FULL, int CPW>
SDRD_DEVICE void enc_matvec_pass_t(const uint32_t* SDRD_RESTRICT img, const uint32_t* SDRD_RESTRICT coef, int row0,
                                   int nrows, const unsigned char* SDRD_RESTRICT tab, uint32_t* rec16, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t acc[RB][4];
#pragma unroll
    for (int r = 0; r < RB; r++)
#pragma unroll
        for (int w = 0; w < 4; w++) acc[r][w] = 0u;
    for (int jj = 0; jj < CPW; jj += 2) {
        const int j = warp * CPW + jj;
        uint32_t s0[2][4], s1[2][4], s2[2][4];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int w = 0; w < 4; w++) selectors(img[(j + c) * ROW_WORDS + lane + 32 * w], s0[c][w], s1[c][w], s2[c][w]);
        /* coef[r][j] = 32 * M[r][j] as a full word: the two columns' entries arrive in one 8-byte load and go straight
         * into the table addresses (16-bit entries cost two ALU instructions per row to pull apart) */
        const uint2* co = reinterpret_cast<const uint2*>(coef + row0 * 128 + j);
#pragma unroll
        for (int r = 0; r < RB; r++) {
            if (FULL || r < nrows) {
                const uint2 cc = co[r * 64];
                const unsigned char* e0 = tab + cc.x;
                const unsigned char* e1 = tab + cc.y;
                const uint4 a0 = *reinterpret_cast<const uint4*>(e0);
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(e0 + 16);
                const uint4 a1 = *reinterpret_cast<const uint4*>(e1);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(e1 + 16);
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    uint32_t v = acc[r][w];
                    v = v ^ prmt(a0.x, a0.y, s0[0][w]) ^ prmt(a0.z, a0.w, s1[0][w]);
                    v = v ^ prmt(b0, b0, s2[0][w]) ^ prmt(a1.x, a1.y, s0[1][w]);
                    v = v ^ prmt(a1.z, a1.w, s1[1][w]) ^ prmt(b1, b1, s2[1][w]);
                    acc[r][w] = v;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
        if (FULL || r < nrows) {
#pragma unroll
            for (int w = 0; w < 4; w++) atomicXor(&rec16[r * ROW_WORDS + lane + 32 * w], acc[r][w]);
        }
    }
}

template <int CPW>
SDRD_DEVICE void enc_matvec_pass(const uint32_t* SDRD_RESTRICT img, const uint32_t* SDRD_RESTRICT coef, int row0,
                                 int nrows, const unsigned char* SDRD_RESTRICT tab, uint32_t* rec16, int tid)
{
    if (nrows == RB) enc_matvec_pass_t<true, CPW>(img, coef, row0, nrows, tab, rec16, tid);
    else enc_matvec_pass_t<false, CPW>(img, coef, row0, nrows, tab, rec16, tid);
}

/* Persistent: CTA c encodes work items c, c + gridDim.x, ... (item = stream * n_frames + frame).  While
 * item i is being encoded out of one image buffer, the samples of item i + gridDim.x arrive in the other
 * through cp.async (4-byte LDGSTS: the 127-sample blocks are neither 16-byte aligned nor a multiple of 16
 * bytes long, which rules the bulk copy out). */
#ifndef SDRD_K2_BULK_ORIGINALS
#define SDRD_K2_BULK_ORIGINALS 1
#endif
template <bool TWO>
SDRD_KERNEL(EncShape<TWO>::NT, EncShape<TWO>::CTAS) encode_kernel(EncParams p)
{
    constexpr int NTH = EncShape<TWO>::NT;
    SDRD_DYN_SMEM(smem_raw);
    const int tid = (int)threadIdx.x;
    Smem sm = carve(smem_raw, 2 * p.cstride); /* the coefficient area holds full words here */
    uint32_t* const img2[2] = {sm.img, TWO ? sm.img : reinterpret_cast<uint32_t*>(sm.extra)};
    uint32_t* const coef = reinterpret_cast<uint32_t*>(sm.coefT); /* [cstride rows][128 blocks] */
    const long long n_items = (long long)p.n_frames * p.n_streams;

    /* request the payload words of work item `it` into image `im`, write its header words */
    auto gather = [&](long long it, uint32_t* im) {
        const int s = (int)(it / p.n_frames), f = (int)(it - (long long)s * p.n_frames);
        const unsigned frame_index = (p.frame_index0 + (unsigned)f) & 0xFFFFu;
        if (p.mode == 0) {
            /* UDPSinkFEC::write: block 0 = meta data, blocks 1..127 = 127 samples each */
            const bool begun_earlier = f == 0 && p.n_pending > 0;
            const uint32_t* meta = begun_earlier ? p.meta_first : p.meta_next;
            for (int k = tid; k < ROW_WORDS; k += NTH) {
                uint32_t v = 0;
                if (k == 0) v = frame_index;
                else if (k <= 6) v = meta[k - 1];
                if (p.stamp_rate && !begun_earlier && k >= 4 && k <= 6) {
                    /* frame begun (long long)f * FRAME_SAMPLES - n_pending samples into this call */
                    uint32_t m5[5] = {meta[0], meta[1], meta[2], 0u, 0u};
                    const unsigned long long o = (unsigned long long)((long long)f * FRAME_SAMPLES - p.n_pending);
                    const unsigned long long us = (unsigned long long)meta[4] + o * 1000000ull / p.stamp_rate;
                    m5[3] = meta[3] + (uint32_t)(us / 1000000ull);
                    m5[4] = (uint32_t)(us % 1000000ull);
                    v = k == 4 ? m5[3] : k == 5 ? m5[4] : crc32_meta(m5);
                }
                im[k] = v;
            }
            for (int k = tid; k < 127; k += NTH) im[(k + 1) * ROW_WORDS] = frame_index | ((uint32_t)(k + 1) << 16);
            const uint32_t* src = p.samples + (long long)s * p.sample_stride;
            const uint32_t* pend = p.pending + (long long)s * FRAME_SAMPLES;
            const long long g0 = (long long)f * FRAME_SAMPLES - p.n_pending; /* index into this call's samples */
            /* one warp per block of 127 samples: four coalesced requests, no index division */
            for (int b = tid >> 5; b < 127; b += NTH / 32) {
                const long long gb = g0 + (long long)b * 127;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = (tid & 31) + 32 * q;
                    if (i < 127) {
                        const long long g = gb + i;
                        cp_async4(&im[(b + 1) * ROW_WORDS + 1 + i], g < 0 ? &pend[g + p.n_pending] : &src[g]);
                    }
                }
            }
        } else {
            for (int j = tid >> 5; j < 128; j += NTH / 32) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(p.originals + ((long long)f * 128 + j) * p.block_pitch);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = (tid & 31) + 32 * q;
                    if (i < 127) cp_async4(&im[j * ROW_WORDS + 1 + i], &row[i]);
                }
            }
            for (int k = tid; k < 128; k += NTH) im[k * ROW_WORDS] = 0u;
        }
        cp_async_commit();
    };

    /* the first item's samples are requested before the tables are loaded: their way from HBM runs under that */
    long long it = blockIdx.x;
    int cur = 0;
    if (it < n_items) gather(it, img2[0]);
    load_tables(sm, p.tab, tid);
    /* the Cauchy rows 128 .. 128 + F - 1 as table offsets */
    for (int k = tid; k < 128 * p.cstride; k += NTH) {
        const int r = k >> 7;
        coef[k] = (uint32_t)TAB_ENTRY * (r < p.F ? (uint32_t)p.tab.cauchy[k] : 0u);
    }
    for (; it < n_items; it += gridDim.x, cur ^= 1) {
        cp_async_wait_all();
#if SDRD_K2_BULK_ORIGINALS
        fence_proxy_async_smem();                 /* this thread's share of image `cur`, visible to the bulk copy below */
        if (tid == 0) tma_store_wait_read();      /* the bulk copy out of image `cur ^ 1` (previous item) has read it */
#endif
        __syncthreads(); /* image `cur` complete; everybody is done with image `cur ^ 1` */
        if (!TWO && it + gridDim.x < n_items) gather(it + gridDim.x, img2[cur ^ 1]);
        const uint32_t* img = img2[cur];
        const int s = (int)(it / p.n_frames), f = (int)(it - (long long)s * p.n_frames);
        const unsigned frame_index = (p.frame_index0 + (unsigned)f) & 0xFFFFu;

        uint32_t* out = nullptr;
        if (p.mode == 0) {
            out = p.dgrams + (long long)s * p.dgram_stride + (long long)f * (128 + p.F) * ROW_WORDS;
            /* the 128 original datagrams leave as they are: 64 KB, contiguous on both sides */
#if SDRD_K2_BULK_ORIGINALS
            /* ... as bulk copies by the async engine (UBLKCP shared -> global), under the products below */
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < 4; q++) tma_store_1d(out + q * (IMG_WORDS / 4), img + q * (IMG_WORDS / 4), IMG_WORDS);
                tma_store_commit();
            }
#else
            const uint4* src4 = reinterpret_cast<const uint4*>(img);
            uint4* dst4 = reinterpret_cast<uint4*>(out);
            for (int k = tid; k < IMG_WORDS / 4; k += NTH) dst4[k] = src4[k];
#endif
        }
        for (int row0 = 0; row0 < p.F; row0 += RB) {
            const int nrows = p.F - row0 < RB ? p.F - row0 : RB;
            for (int k = tid; k < RB * ROW_WORDS; k += NTH) sm.rec16[k] = 0u;
            __syncthreads();
            enc_matvec_pass<EncShape<TWO>::CPW>(img, coef, row0, nrows, sm.tab, sm.rec16, tid);
            __syncthreads();
            if (p.mode == 0) {
                /* recovery datagram = header {frameIndex, blockIndex = 128 + r, filler 0} + payload */
                for (int k = tid; k < nrows; k += NTH)
                    sm.rec16[k * ROW_WORDS] = frame_index | ((uint32_t)(128 + row0 + k) << 16);
                __syncthreads();
                const uint4* src4 = reinterpret_cast<const uint4*>(sm.rec16);
                uint4* dst4 = reinterpret_cast<uint4*>(out + (128 + row0) * ROW_WORDS);
                for (int k = tid; k < nrows * ROW_WORDS / 4; k += NTH) dst4[k] = src4[k];
            } else {
                for (int r = tid >> 5; r < nrows; r += NTH / 32) {
                    uint32_t* row = reinterpret_cast<uint32_t*>(p.recovery + ((long long)f * p.F + row0 + r) * 508);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int i = (tid & 31) + 32 * q;
                        if (i < 127) row[i] = sm.rec16[r * ROW_WORDS + 1 + i];
                    }
                }
            }
            __syncthreads();
        }
        if (TWO && it + gridDim.x < n_items) { /* one image: the next item is gathered once this one has left it */
#if SDRD_K2_BULK_ORIGINALS
            if (tid == 0) tma_store_wait_read();
#endif
            __syncthreads();
            gather(it + gridDim.x, img2[0]);
        }
    }
    cp_async_wait_all();
#if SDRD_K2_BULK_ORIGINALS
    if (tid == 0) tma_store_wait_read(); /* the images stay allocated until the last bulk copy has read them */
#endif
}

/* ============================================================================ decode ==== */

constexpr int ST_INCOMPLETE = 0, ST_COMPLETE = 1, ST_RECOVERED = 2, ST_FAILED = -1, ST_NEEDS_BIG = -100;

struct DecParams {
    const uint32_t* sb;        /* frame f datagram i: sb + (f * blocks_pitch + i) * 128 */
    long long blocks_pitch;    /* datagrams */
    const long long* frame_start; /* optional: frame f begins at datagram frame_start[f] instead of f * blocks_pitch */
    const int* n_blocks;       /* [n_frames] */
    uint32_t* payload;         /* frame f: payload + f * 127 * 127 words (blocks 1..127) */
    uint32_t* block0;          /* frame f: block0 + f * 127 words, may be null */
    int* status;               /* [n_frames] */
    int pass;                  /* 0: frames needing more than DCAP rows are flagged and left; 1: only flagged frames */
    /* Bare cm256_decode form (sdrd_cm256_decode_blocks): when `recovered` is set nothing is written to payload /
     * block0; instead the block solved into recovery descriptor k (arrival order) -- erased original k in ascending
     * order -- goes to recovered[(f * 128 + k) * 127 ..], for EVERY k: copying back only the last N descriptors is
     * SDRdaemonFECBuffer's doing (.cpp:208-213), not the library's. */
    uint32_t* recovered;
    int n_frames;              /* decode_stream_kernel: frames of the batch (its CTAs are persistent) */
    int general_single;        /* 1: a lone recovery block is solved like any other (cm256 takes its XOR shortcut only
                                  when params.RecoveryCount == 1, not when one recovery block happens to be present) */
    Tables tab;
};

template <int DCAP>
inline size_t dec_smem_bytes()
{
    /* coefT [128][DCAP] + scratch [DCAP][2*DCAP] (log of the inverse) + exp/log + small lists */
    return SMEM_FIXED + (size_t)128 * DCAP * 2 + (size_t)DCAP * 2 * DCAP + 512 + 256 + 2048;
}

template <int DCAP>
SDRD_KERNEL(NT, (DCAP <= 32 ? 2 : 1)) decode_kernel(DecParams p)
{
    SDRD_DYN_SMEM(smem_raw);
    const int tid = (int)threadIdx.x;
    const int f = (int)blockIdx.x;
    if (p.pass == 1 && p.status[f] != ST_NEEDS_BIG) return;

    Smem sm = carve(smem_raw, DCAP);
    uint8_t* aug = sm.extra;                          /* [DCAP][2*DCAP] */
    uint8_t* gfexp = aug + DCAP * 2 * DCAP;           /* [512] */
    uint8_t* gflog = gfexp + 512;                     /* [256] */
    int* lists = reinterpret_cast<int*>(gflog + 256); /* 2048 bytes */
    int* origRow = lists;                 /* [128] row of the image holding original b, or -1 */
    int* origCnt = lists + 128;           /* [128] times original b was received */
    unsigned* recMask = reinterpret_cast<unsigned*>(lists + 256);  /* [4] bit i: datagram i is a recovery block */
    unsigned* missMask = recMask + 4;                              /* [4] bit b: original b missing */
    int* flags = lists + 264;             /* [0] repeated original, [1] repeated recovery row (singular system) */
    unsigned* recSeen = reinterpret_cast<unsigned*>(lists + 266);  /* [4] bit r: recovery row 128 + r received */
    uint8_t* recRowOf = reinterpret_cast<uint8_t*>(lists + 272);   /* [128] k -> image row */
    uint8_t* recIdxOf = recRowOf + 128;                            /* [128] k -> block index (128..255) */
    uint8_t* erased = recIdxOf + 128;                              /* [128] c -> erased original index */

    load_tables(sm, p.tab, tid);
    for (int i = tid; i < 512; i += NT) gfexp[i] = p.tab.gfexp[i];
    for (int i = tid; i < 256; i += NT) gflog[i] = p.tab.gflog[i];

    int nb = p.n_blocks[f];
    if (nb > 128) nb = 128; /* blocks beyond the first 128 received are dropped (.cpp:143) */
    if (nb > p.blocks_pitch) nb = (int)p.blocks_pitch; /* never read past the frame's slots */
    if (nb < 0) nb = 0;
    {
        const long long first = p.frame_start ? p.frame_start[f] : (long long)f * p.blocks_pitch;
        const uint4* src4 = reinterpret_cast<const uint4*>(p.sb + first * ROW_WORDS);
        uint4* dst4 = reinterpret_cast<uint4*>(sm.img);
        for (int k = tid; k < nb * (ROW_WORDS / 4); k += NT) dst4[k] = src4[k];
        for (int k = nb * (ROW_WORDS / 4) + tid; k < IMG_WORDS / 4; k += NT) dst4[k] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 128) {
        origRow[tid] = -1;
        origCnt[tid] = 0;
    }
    if (tid < 8) recMask[tid] = 0u; /* recMask + missMask */
    if (tid < 6) flags[tid] = 0;    /* flags + recSeen */
    __syncthreads();

    /* classify the received datagrams by header.blockIndex (.cpp:143-166) */
    if (tid < nb) {
        const int idx = (int)((sm.img[tid * ROW_WORDS] >> 16) & 0xFFu);
        if (idx < 128) {
            atomicMax(&origRow[idx], tid); /* a repeated original overwrites the earlier copy */
            atomicAdd(&origCnt[idx], 1);
        } else {
            atomicOr(&recMask[tid >> 5], 1u << (tid & 31));
            /* the same recovery row twice (a duplicated datagram): two equal rows, the system is singular */
            if (atomicOr(&recSeen[(idx - 128) >> 5], 1u << (idx & 31)) & (1u << (idx & 31))) flags[1] = 1;
        }
    }
    __syncthreads();
    if (tid < 128) {
        if (origRow[tid] < 0) atomicOr(&missMask[tid >> 5], 1u << (tid & 31));
        if (origCnt[tid] > 1) flags[0] = 1;
    }
    __syncthreads();
    const int N = __popc(recMask[0]) + __popc(recMask[1]) + __popc(recMask[2]) + __popc(recMask[3]);
    const int n_missing = __popc(missMask[0]) + __popc(missMask[1]) + __popc(missMask[2]) + __popc(missMask[3]);
    /* ordered lists: recovery blocks in arrival order, erased originals ascending */
    if (tid < 128) {
        const int wq = tid >> 5;
        const unsigned below = (1u << (tid & 31)) - 1u;
        int rr = 0, rm = 0;
        for (int q = 0; q < wq; q++) {
            rr += __popc(recMask[q]);
            rm += __popc(missMask[q]);
        }
        if (recMask[wq] & (1u << (tid & 31))) {
            const int k = rr + __popc(recMask[wq] & below);
            recRowOf[k] = (uint8_t)tid;
            recIdxOf[k] = (uint8_t)((sm.img[tid * ROW_WORDS] >> 16) & 0xFFu);
        }
        if (missMask[wq] & (1u << (tid & 31))) erased[rm + __popc(missMask[wq] & below)] = (uint8_t)tid;
    }
    __syncthreads();

    int st;
    bool do_decode = false;
    if (nb < 128) st = ST_INCOMPLETE;
    else if (N == 0) st = ST_COMPLETE;
    else if (flags[0] || n_missing < N) st = ST_FAILED; /* repeated original: cm256_decode refuses */
    else if (flags[1] && N > 1) st = ST_FAILED;         /* repeated recovery row: no solution; the originals pass through */
    else if (N > DCAP) st = ST_NEEDS_BIG;
    else {
        st = ST_RECOVERED;
        do_decode = true;
    }

    if (do_decode) {
        /* coefficient matrix D [N][128 columns = image rows] */
        if (N == 1 && !p.general_single) {
            /* cm256's single-recovery shortcut: XOR of everything received, whatever the row */
            for (int k = tid; k < 128 * DCAP; k += NT) sm.coefT[k] = (uint16_t)((k % DCAP) == 0 ? TAB_ENTRY : 0);
        } else {
            /* A[k][c] = M[x_k][y_c] = a_c / (x_k ^ y_c) with x_k = index of recovery block k (128..255),
             * y_c = erased original c (0..127), a_y = y ^ 128: a Cauchy matrix with scaled columns.  Both halves
             * of the decode matrix D = [A^-1 | A^-1 M] have closed forms (no elimination, no pivots, no
             * N-term sums per entry).  With
             *   P_c = prod_j (y_c ^ x_j) / prod_{j != c} (y_c ^ y_j),   Q_k = prod_j (x_k ^ y_j) / prod_{j != k} (x_k ^ x_j),
             *   g(y) = prod_c (y ^ y_c) / prod_k (y ^ x_k)   (= 1 + sum_k Q_k / (y ^ x_k), partial fractions)
             * the weight of recovery block k in erased original c is   P_c Q_k / (a_c (x_k ^ y_c))
             * and, because 1/((x^a)(x^b)) = (1/(x^a) + 1/(x^b)) / (a^b) in characteristic 2, the weight of a
             * received original o is   sum_k A^-1[c][k] M[x_k][o] = P_c a_o g(y_o) / (a_c (y_c ^ y_o)).
             * Both read  D[c][i] = exp( lp[c] + lbase[i] - log(z_i ^ y_c) )  with z_i the block index of image
             * row i, lp[c] = log(P_c / a_c) and lbase[i] = log Q_k resp. log(a_o g(y_o)): one pass of 2N
             * logarithms per image row, then two table look-ups per entry.  Everything is kept as discrete
             * logarithms (none of these factors is ever zero). */
            uint8_t* lbase = aug;               /* [128] per image row */
            uint8_t* lp = erased + 128;         /* [N] log (P_c / a_c) */
            if (tid < 128) {
                const int i = tid;
                const int z = (int)((sm.img[i * ROW_WORDS] >> 16) & 0xFFu);
                int me = -1; /* recovery rows: their own slot is left out of the x-product */
                bool used = true;
                if (z >= 128) {
                    const int wq = i >> 5;
                    me = __popc(recMask[wq] & ((1u << (i & 31)) - 1u));
                    for (int q = 0; q < wq; q++) me += __popc(recMask[q]);
                } else {
                    used = origRow[z] == i;
                }
                int acc = 0;
                if (used) {
                    if (z < 128) acc = gflog[z ^ 128];
                    for (int j = 0; j < N; j++) {
                        acc += gflog[z ^ erased[j]];
                        if (j != me) acc += 255 - gflog[z ^ recIdxOf[j]];
                    }
                }
                lbase[i] = (uint8_t)(acc % 255);
            } else if (tid < 128 + N) {
                const int me = tid - 128;
                const int v = erased[me];
                int acc = 255 - gflog[v ^ 128];
                for (int j = 0; j < N; j++) {
                    acc += gflog[v ^ recIdxOf[j]];
                    if (j != me) acc += 255 - gflog[v ^ erased[j]];
                }
                lp[me] = (uint8_t)(acc % 255);
            }
            __syncthreads();
            /* thread (i, h) fills image row i's column of D for the rows c in every other block of 16 */
            {
                const int i = tid & 127, h = tid >> 7;
                const int z = (int)((sm.img[i * ROW_WORDS] >> 16) & 0xFFu);
                const bool used = z >= 128 || origRow[z] == i;
                const int lb = lbase[i];
                for (int c0 = 16 * h; c0 < N; c0 += 32) {
#pragma unroll
                    for (int cc = 0; cc < 16; cc++) {
                        const int c = c0 + cc;
                        if (c < N) {
                            int t = lp[c] + lb;
                            if (t >= 255) t -= 255;
                            const uint32_t v = used ? gfexp[t + 255 - gflog[z ^ erased[c]]] : 0u;
                            sm.coefT[i * DCAP + c] = (uint16_t)(TAB_ENTRY * v);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    uint32_t* pay = p.payload + (long long)f * 127 * 127;
    uint32_t* b0 = p.block0 ? p.block0 + (long long)f * 127 : nullptr;

    /* originals that arrived go out as they are; blocks that did not arrive read as zero (.cpp:109) */
    for (int b = tid >> 5; b < 128 && !p.recovered; b += NT / 32) { /* one warp per block: four coalesced stores, no index division */
        const int row = origRow[b];
        uint32_t* dstb = b == 0 ? b0 : pay + (b - 1) * 127;
        if (!dstb || !(b == 0 || row >= 0 || !do_decode)) continue;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int i = (tid & 31) + 32 * q;
            if (i < 127) dstb[i] = row >= 0 ? sm.img[row * ROW_WORDS + 1 + i] : 0u;
        }
    }

    if (do_decode) {
        for (int row0 = 0; row0 < N; row0 += RB) {
            const int nrows = N - row0 < RB ? N - row0 : RB;
            for (int k = tid; k < RB * ROW_WORDS; k += NT) sm.rec16[k] = 0u;
            __syncthreads();
            matvec_pass(sm.img, sm.coefT, DCAP, row0, nrows, sm.tab, sm.rec16, tid);
            __syncthreads();
            for (int r = tid >> 5; r < nrows && p.recovered; r += NT / 32) {
                uint32_t* dstb = p.recovered + ((long long)f * 128 + row0 + r) * 127;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = (tid & 31) + 32 * q;
                    if (i < 127) dstb[i] = sm.rec16[r * ROW_WORDS + 1 + i];
                }
            }
            for (int r = tid >> 5; r < nrows && !p.recovered; r += NT / 32) {
                const int b = erased[row0 + r];
                /* The reference copies back only the LAST N descriptors (.cpp:208-213), i.e. it
                 * assumes the recovery blocks arrived after the originals; a recovery block that
                 * arrived earlier keeps its result to itself and the erased block stays zero. */
                const bool copied = (int)recRowOf[row0 + r] >= 128 - N;
                uint32_t* dstb = b == 0 ? b0 : pay + (b - 1) * 127;
                if (!dstb) continue;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int i = (tid & 31) + 32 * q;
                    if (i < 127) dstb[i] = copied ? sm.rec16[r * ROW_WORDS + 1 + i] : 0u;
                }
            }
            __syncthreads();
        }
        /* erased originals beyond the N recovered ones stay zero */
        for (int c = N; c < n_missing && !p.recovered; c++) {
            const int b = erased[c];
            for (int i = tid; i < 127; i += NT) {
                if (b == 0) {
                    if (b0) b0[i] = 0u;
                } else {
                    pay[(b - 1) * 127 + i] = 0u;
                }
            }
        }
    }
    if (tid == 0) p.status[f] = st;
}


/* ============================================================================ decode, streaming form ====
 * K3 as of round 2.  The CTA-wide kernel above stages the whole frame (64 KB) in shared memory and walks it once
 * per block of 16 output rows; with the 20 erasures of BASELINE config 4 that is two passes, the second with 4 of
 * its 16 rows in use, a dozen CTA barriers per frame and two resident CTAs per SM.  Here every datagram word is read
 * ONCE, straight from global memory into registers:
 *   - a CTA is four warps = two halves of the datagram words x two groups of output rows; lanes hold 2 words (8-byte
 *     loads, 256 B per warp) of every one of the 128 datagrams, fetched two column pairs ahead into registers that
 *     are reloaded the moment their selectors are built;
 *   - all N <= 32 output rows accumulate in registers in one pass (2 words x N/2 rows per warp, N rounded up to a
 *     multiple of 4, the loop picked per frame from eight instantiations): a warp owns its words of its rows, so
 *     there is nothing to merge and the recovered blocks leave from the accumulators (the price: the digit selectors
 *     of a data word, 3 % of the work, are built by both row groups);
 *   - received originals go out from the same registers (no shared-memory image, no second read);
 *   - the coefficient matrix is held as full-word table offsets [row][datagram]: the entries of a column pair arrive
 *     in one 8-byte load and go straight into the address of the table loads (packed 16-bit entries cost two ALU
 *     instructions per row to pull apart, and the ALU pipe is what bounds this kernel);
 *   - 28 KB of shared memory per CTA (tables, coefficient matrix, lists), four CTAs per SM.
 * Classification and the closed-form decode matrix are those of the kernel above (which stays for frames with more
 * than 32 recovery blocks). */
#ifndef SDRD_K3_CTAS_PER_SM
#define SDRD_K3_CTAS_PER_SM 4 /* resident 128-thread CTAs per SM the register budget is set for (4: 128 registers) */
#endif
constexpr int DS_NT = 128;
constexpr int DS_CAP = 32;   /* rows the streaming form holds */
constexpr size_t ds_smem_bytes() { return 256 * TAB_ENTRY + 128 * DS_CAP * 4 + 512 + 256 + 3072; }

struct DsOut {
    uint32_t* pay;        /* payload of the frame (blocks 1..127) or null */
    uint32_t* b0;         /* block 0 or null */
    uint32_t* recovered;  /* bare cm256 form: recovered + (f * 128 + c) * 127, or null */
    const uint32_t* hdr;
    const int* origRow;
    const uint8_t* erased;
    const uint8_t* recRowOf;
    int N;
};

/* NR = rows per warp: warp (half, rg) holds rows rg * NR .. rg * NR + NR - 1 for the word half `half` */
template <int NR>
SDRD_DEVICE void ds_pass(const uint32_t* SDRD_RESTRICT frame, int half, int rg, int lane, const uint32_t* SDRD_RESTRICT coef_all,
                         const unsigned char* SDRD_RESTRICT tab, const DsOut& o)
{
    const int m = 32 * half + lane;           /* this lane's word pair: words 2m, 2m + 1 of every datagram */
    const uint2* src = reinterpret_cast<const uint2*>(frame) + m;
    const bool pass_through = o.recovered == nullptr && rg == 0; /* one row group sends the received originals on */
    const uint32_t* coef = coef_all + rg * NR * 128;
    uint32_t acc[NR][2];
#pragma unroll
    for (int r = 0; r < NR; r++) acc[r][0] = acc[r][1] = 0u;
    /* two column pairs in flight (A, B): a pair's words are dead once its selectors are built, so its registers are
     * reloaded with the pair two ahead right there -- no register moves, the loads have a whole row loop to arrive */
    uint2 xa0 = src[0], xa1 = src[ROW_WORDS / 2];
    uint2 xb0 = src[2 * (ROW_WORDS / 2)], xb1 = src[3 * (ROW_WORDS / 2)];
    auto column_pair = [&](int jj, uint2& c0, uint2& c1) {
        if (pass_through) { /* a received original leaves as it is: payload word k of the datagram is word k + 1 */
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int i = jj + c;
                const int z = (int)((o.hdr[i] >> 16) & 0xFFu);
                if (z < 128 && o.origRow[z] == i) {
                    uint32_t* d = z == 0 ? o.b0 : o.pay + (z - 1) * 127;
                    if (d) {
                        const uint2 v = c ? c1 : c0;
                        if (m > 0) d[2 * m - 1] = v.x;
                        d[2 * m] = v.y;
                    }
                }
            }
        }
        uint32_t s0[2][2], s1[2][2], s2[2][2];
        selectors(c0.x, s0[0][0], s1[0][0], s2[0][0]);
        selectors(c0.y, s0[0][1], s1[0][1], s2[0][1]);
        selectors(c1.x, s0[1][0], s1[1][0], s2[1][0]);
        selectors(c1.y, s0[1][1], s1[1][1], s2[1][1]);
        if (jj + 4 < 128) {
            c0 = src[(size_t)(jj + 4) * (ROW_WORDS / 2)];
            c1 = src[(size_t)(jj + 5) * (ROW_WORDS / 2)];
        }
        const uint2* co = reinterpret_cast<const uint2*>(coef + jj);
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const uint2 cc = co[r * (128 / 2)];
            const unsigned char* e0 = tab + cc.x;
            const unsigned char* e1 = tab + cc.y;
            const uint4 a0 = *reinterpret_cast<const uint4*>(e0);
            const uint32_t q0 = *reinterpret_cast<const uint32_t*>(e0 + 16);
            const uint4 a1 = *reinterpret_cast<const uint4*>(e1);
            const uint32_t q1 = *reinterpret_cast<const uint32_t*>(e1 + 16);
#pragma unroll
            for (int w = 0; w < 2; w++) {
                uint32_t v = acc[r][w];
                v = v ^ prmt(a0.x, a0.y, s0[0][w]) ^ prmt(a0.z, a0.w, s1[0][w]);
                v = v ^ prmt(q0, q0, s2[0][w]) ^ prmt(a1.x, a1.y, s0[1][w]);
                v = v ^ prmt(a1.z, a1.w, s1[1][w]) ^ prmt(q1, q1, s2[1][w]);
                acc[r][w] = v;
            }
        }
    };
#pragma unroll 1
    for (int jj = 0; jj < 128; jj += 4) {
        column_pair(jj, xa0, xa1);
        column_pair(jj + 2, xb0, xb1);
    }
    /* row c = the original solved into recovery descriptor c = erased original c: straight from the accumulators */
#pragma unroll
    for (int r = 0; r < NR; r++) {
        const int c = rg * NR + r;
        if (c < o.N) {
            uint32_t* d;
            bool copied = true;
            if (o.recovered) {
                d = o.recovered + c * 127;
            } else {
                const int b = o.erased[c];
                /* the reference copies back only the LAST N descriptors (.cpp:208-213): a recovery block that arrived
                 * before an original keeps its result to itself and the erased block stays zero */
                copied = (int)o.recRowOf[c] >= 128 - o.N;
                d = b == 0 ? o.b0 : o.pay + (b - 1) * 127;
            }
            if (d) {
                if (m > 0) d[2 * m - 1] = copied ? acc[r][0] : 0u;
                d[2 * m] = copied ? acc[r][1] : 0u;
            }
        }
    }
}

SDRD_KERNEL(DS_NT, SDRD_K3_CTAS_PER_SM) decode_stream_kernel(DecParams p)
{
    const int n_frames = p.n_frames;
    SDRD_DYN_SMEM(smem_raw);
    const int tid = (int)threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    unsigned char* tab = smem_raw;                                                   /* [256][32] */
    uint32_t* coef = reinterpret_cast<uint32_t*>(tab + 256 * TAB_ENTRY);             /* [DS_CAP][128] */
    uint8_t* gfexp = reinterpret_cast<uint8_t*>(coef + DS_CAP * 128);                /* [512] */
    uint8_t* gflog = gfexp + 512;                                                    /* [256] */
    int* lists = reinterpret_cast<int*>(gflog + 256);                                /* 3072 bytes */
    uint32_t* hdr = reinterpret_cast<uint32_t*>(lists);   /* [128] header word of datagram i */
    int* origRow = lists + 128;                           /* [128] datagram holding original b, or -1 */
    int* origCnt = lists + 256;                           /* [128] */
    unsigned* recMask = reinterpret_cast<unsigned*>(lists + 384);  /* [4] */
    unsigned* missMask = recMask + 4;                              /* [4] */
    int* flags = lists + 392;                                      /* [2] */
    unsigned* recSeen = reinterpret_cast<unsigned*>(lists + 394);  /* [4] */
    uint8_t* recRowOf = reinterpret_cast<uint8_t*>(lists + 400);   /* [128] */
    uint8_t* recIdxOf = recRowOf + 128;
    uint8_t* erased = recIdxOf + 128;
    uint8_t* lbase = erased + 128;
    uint8_t* lp = lbase + 128;                                     /* [DS_CAP] */

    for (int i = tid; i < 256; i += DS_NT) {
        *reinterpret_cast<uint4*>(tab + i * TAB_ENTRY) = p.tab.tabA[i];
        *reinterpret_cast<uint32_t*>(tab + i * TAB_ENTRY + 16) = p.tab.tabB[i];
    }
    for (int i = tid; i < 512; i += DS_NT) gfexp[i] = p.tab.gfexp[i];
    for (int i = tid; i < 256; i += DS_NT) gflog[i] = p.tab.gflog[i];

    for (int f = (int)blockIdx.x; f < n_frames; f += (int)gridDim.x) {
        __syncthreads(); /* the lists of the previous frame are no longer read */
        int nb = p.n_blocks[f];
        if (nb > 128) nb = 128; /* blocks beyond the first 128 received are dropped (.cpp:143) */
        if (nb > p.blocks_pitch) nb = (int)p.blocks_pitch;
        if (nb < 0) nb = 0;
        const long long first = p.frame_start ? p.frame_start[f] : (long long)f * p.blocks_pitch;
        const uint32_t* frame = p.sb + first * ROW_WORDS;
        for (int i = tid; i < 128; i += DS_NT) {
            hdr[i] = i < nb ? frame[(size_t)i * ROW_WORDS] : 0u;
            origRow[i] = -1;
            origCnt[i] = 0;
        }
        if (tid < 8) recMask[tid] = 0u;
        if (tid < 6) flags[tid] = 0;
        __syncthreads();

        /* classify the received datagrams by header.blockIndex (.cpp:143-166) */
        for (int i = tid; i < nb; i += DS_NT) {
            const int idx = (int)((hdr[i] >> 16) & 0xFFu);
            if (idx < 128) {
                atomicMax(&origRow[idx], i); /* a repeated original overwrites the earlier copy */
                atomicAdd(&origCnt[idx], 1);
            } else {
                atomicOr(&recMask[i >> 5], 1u << (i & 31));
                if (atomicOr(&recSeen[(idx - 128) >> 5], 1u << (idx & 31)) & (1u << (idx & 31))) flags[1] = 1;
            }
        }
        __syncthreads();
        for (int i = tid; i < 128; i += DS_NT) {
            if (origRow[i] < 0) atomicOr(&missMask[i >> 5], 1u << (i & 31));
            if (origCnt[i] > 1) flags[0] = 1;
        }
        __syncthreads();
        const int N = __popc(recMask[0]) + __popc(recMask[1]) + __popc(recMask[2]) + __popc(recMask[3]);
        const int n_missing = __popc(missMask[0]) + __popc(missMask[1]) + __popc(missMask[2]) + __popc(missMask[3]);
        for (int i = tid; i < 128; i += DS_NT) {
            const int wq = i >> 5;
            const unsigned below = (1u << (i & 31)) - 1u;
            int rr = 0, rm = 0;
            for (int q = 0; q < wq; q++) {
                rr += __popc(recMask[q]);
                rm += __popc(missMask[q]);
            }
            if (recMask[wq] & (1u << (i & 31))) {
                const int k = rr + __popc(recMask[wq] & below);
                recRowOf[k] = (uint8_t)i;
                recIdxOf[k] = (uint8_t)((hdr[i] >> 16) & 0xFFu);
            }
            if (missMask[wq] & (1u << (i & 31))) erased[rm + __popc(missMask[wq] & below)] = (uint8_t)i;
        }
        __syncthreads();

        int st;
        bool do_decode = false;
        if (nb < 128) st = ST_INCOMPLETE;
        else if (N == 0) st = ST_COMPLETE;
        else if (flags[0] || n_missing < N) st = ST_FAILED; /* repeated original: cm256_decode refuses */
        else if (flags[1] && N > 1) st = ST_FAILED;         /* repeated recovery row: no solution */
        else if (N > DS_CAP) st = ST_NEEDS_BIG;
        else {
            st = ST_RECOVERED;
            do_decode = true;
        }
        if (tid == 0) p.status[f] = st;
        if (st == ST_NEEDS_BIG) continue; /* left to decode_kernel<128>, which writes everything of this frame */
        const int NR = (N + 3) & ~3;

        if (do_decode) {
            /* the decode matrix D [NR][128 datagrams], see decode_kernel for the closed form */
            const bool shortcut = N == 1 && !p.general_single;
            if (!shortcut) {
                for (int i = tid; i < 128; i += DS_NT) {
                    const int z = (int)((hdr[i] >> 16) & 0xFFu);
                    const bool used = z >= 128 || origRow[z] == i;
                    int me = -1;
                    if (z >= 128) {
                        const int wq = i >> 5;
                        me = __popc(recMask[wq] & ((1u << (i & 31)) - 1u));
                        for (int q = 0; q < wq; q++) me += __popc(recMask[q]);
                    }
                    int a = 0;
                    if (used) {
                        if (z < 128) a = gflog[z ^ 128];
                        for (int j = 0; j < N; j++) {
                            a += gflog[z ^ erased[j]];
                            if (j != me) a += 255 - gflog[z ^ recIdxOf[j]];
                        }
                    }
                    lbase[i] = (uint8_t)(a % 255);
                }
                if (tid < N) {
                    const int v = erased[tid];
                    int b = 255 - gflog[v ^ 128];
                    for (int j = 0; j < N; j++) {
                        b += gflog[v ^ recIdxOf[j]];
                        if (j != tid) b += 255 - gflog[v ^ erased[j]];
                    }
                    lp[tid] = (uint8_t)(b % 255);
                }
                __syncthreads();
            }
            for (int i = tid; i < 128; i += DS_NT) {
                const int z = (int)((hdr[i] >> 16) & 0xFFu);
                const bool used = z >= 128 || origRow[z] == i;
                const int lb = shortcut ? 0 : lbase[i];
                for (int c = 0; c < NR; c++) {
                    uint32_t v = 0u;
                    if (shortcut) {
                        v = c == 0 ? 1u : 0u; /* cm256's single-recovery shortcut: XOR of everything received */
                    } else if (c < N && used) {
                        int t = lp[c] + lb;
                        if (t >= 255) t -= 255;
                        v = gfexp[t + 255 - gflog[z ^ erased[c]]];
                    }
                    coef[c * 128 + i] = (uint32_t)TAB_ENTRY * v;
                }
            }
            __syncthreads();
        }

        DsOut o;
        o.pay = p.payload ? p.payload + (long long)f * 127 * 127 : nullptr;
        o.b0 = p.block0 ? p.block0 + (long long)f * 127 : nullptr;
        o.recovered = p.recovered ? p.recovered + (long long)f * 128 * 127 : nullptr;
        o.hdr = hdr;
        o.origRow = origRow;
        o.erased = erased;
        o.recRowOf = recRowOf;
        o.N = N;
        const bool raw = p.recovered != nullptr;

        if (do_decode) {
            const int half = warp & 1, rg = warp >> 1;
            switch (NR) { /* two row groups of NR / 2 rows */
                case 4: ds_pass<2>(frame, half, rg, lane, coef, tab, o); break;
                case 8: ds_pass<4>(frame, half, rg, lane, coef, tab, o); break;
                case 12: ds_pass<6>(frame, half, rg, lane, coef, tab, o); break;
                case 16: ds_pass<8>(frame, half, rg, lane, coef, tab, o); break;
                case 20: ds_pass<10>(frame, half, rg, lane, coef, tab, o); break;
                case 24: ds_pass<12>(frame, half, rg, lane, coef, tab, o); break;
                case 28: ds_pass<14>(frame, half, rg, lane, coef, tab, o); break;
                default: ds_pass<16>(frame, half, rg, lane, coef, tab, o); break;
            }
            /* erased originals beyond the N recovered ones stay zero */
            if (!raw)
                for (int c = N + warp; c < n_missing; c += DS_NT / 32) {
                    const int b = erased[c];
                    uint32_t* dstb = b == 0 ? o.b0 : o.pay + (b - 1) * 127;
                    if (!dstb) continue;
                    for (int k = lane; k < 127; k += 32) dstb[k] = 0u;
                }
        } else if (!raw) {
            /* nothing to solve: originals that arrived go out as they are, blocks that did not read as zero (.cpp:109) */
            for (int b = warp; b < 128; b += DS_NT / 32) {
                const int row = origRow[b];
                uint32_t* dstb = b == 0 ? o.b0 : o.pay + (b - 1) * 127;
                if (!dstb) continue;
                const uint32_t* srcb = frame + (size_t)(row < 0 ? 0 : row) * ROW_WORDS + 1;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = lane + 32 * q;
                    if (k < 127) dstb[k] = row >= 0 ? srcb[k] : 0u;
                }
            }
        }
    }
}

} /* namespace fec */
} /* namespace sdrd */
