/*
 * sdrd_oracle.c -- CPU restatement of sdrdaemon's decimate + FEC hot path (see sdrd_oracle.h).
 * TEST INFRASTRUCTURE ONLY: the product path (sdrdaemon_b200/) never links or loads this.
 * Citations are path:line under /root/reference.
 */
#include "sdrd_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ======================================================================= decimator === */

/* HBFIRFilterTraits<64>::hbCoeffs, sdmnbase/HBFilterTraits.cpp:210-228: (int32_t)(c * (1<<14)),
 * truncated toward zero; outermost tap first.  hbShift = 14 (include/HBFilterTraits.h:64-72). */
static const int32_t HB64[16] = {-7,   11,  -20,  32,  -49, 71,   -101,  140,
                                 -190, 256, -345, 469, -656, 978, -1698, 5201};
#define HB_SHIFT 14

/* One half-band decimate-by-2 stage.  The reference keeps even/odd double buffers
 * (IntHalfbandFilterEO1.h:68-97) or one 63-deep double buffer (IntHalfbandFilterDB.h:73-76);
 * both compute, for the pair (x[2n], x[2n+1]) just stored,
 *   acc = sum_i H[i] * (x[2n+1-2i] + x[2n-61+2i]) + ((x[2n-30] (+1 for DB)) << 13)
 *   y[n] = acc >> 13                     (EO1.h:100-147 with EO1i.h:34-87; DB.h:80-107)
 * in wrapping int32 arithmetic.  Here a 64-deep ring of the most recent inputs holds the state. */
typedef struct {
    uint32_t ring[2][64];
    unsigned count;
} hb_stage;

static void hb_reset(hb_stage* s) { memset(s, 0, sizeof(*s)); }

static inline void hb_push(hb_stage* s, int32_t i, int32_t q)
{
    s->ring[0][s->count & 63] = (uint32_t)i;
    s->ring[1][s->count & 63] = (uint32_t)q;
    s->count++;
}

static inline int32_t asr(uint32_t v, unsigned sh)
{
    /* arithmetic shift right of the int32 bit pattern */
    return (int32_t)((v >> sh) | ((v & 0x80000000u) ? ~(0xFFFFFFFFu >> sh) : 0u));
}

static inline void hb_fir(const hb_stage* s, int variant, int32_t* yi, int32_t* yq)
{
    unsigned t = s->count - 1; /* index of x[2n+1] */
    for (int c = 0; c < 2; c++) {
        const uint32_t* r = s->ring[c];
        uint32_t acc = 0;
        for (int i = 0; i < 16; i++) {
            uint32_t a = r[(t - 2 * i) & 63];
            uint32_t b = r[(t - 62 + 2 * i) & 63];
            acc += (a + b) * (uint32_t)HB64[i];
        }
        uint32_t centre = r[(t - 31) & 63];
        if (variant == SDRO_HB_DB) centre += 1u;
        acc += centre << (HB_SHIFT - 1);
        int32_t y = asr(acc, HB_SHIFT - 1);
        if (c == 0) *yi = y; else *yq = y;
    }
}

struct sdro_dec {
    int log2_decim, fcpos, variant;
    hb_stage stage[6];   /* m_decimator2 .. m_decimator64, Decimators.h:57-62 */
};

sdro_dec* sdro_dec_create(int log2_decim, int fcpos, int variant)
{
    if (log2_decim < 0 || log2_decim > 6 || fcpos < 0 || fcpos > 2 || variant < 0 || variant > 1)
        return NULL; /* Downsampler.cpp:39-43,57-61 */
    sdro_dec* d = (sdro_dec*)calloc(1, sizeof(*d));
    d->log2_decim = log2_decim;
    d->fcpos = fcpos;
    d->variant = variant;
    return d;
}
void sdro_dec_destroy(sdro_dec* d) { free(d); }
void sdro_dec_reset(sdro_dec* d) { for (int i = 0; i < 6; i++) hb_reset(&d->stage[i]); }
/* Downsampler::configure (Downsampler.cpp:32-67): only m_decim / m_fcPos change.  The six stage objects
 * of m_decimators persist (Decimators.h:57-62): a stage the new cascade uses continues from whatever it
 * last saw -- its state of the previous configuration, or zeros if it never ran. */
int sdro_dec_configure(sdro_dec* d, int log2_decim, int fcpos)
{
    if (log2_decim < 0 || log2_decim > 6 || fcpos < 0 || fcpos > 2) return -1;
    d->log2_decim = log2_decim;
    d->fcpos = fcpos;
    return 0;
}

/* feed one sample into stage k of an n_stages cascade; returns 1 when the cascade emitted */
static int cascade_feed(sdro_dec* d, int k, int n_stages, int32_t i, int32_t q, int32_t* oi, int32_t* oq)
{
    if (k == n_stages) { *oi = i; *oq = q; return 1; }
    hb_stage* s = &d->stage[k];
    hb_push(s, i, q);
    if (s->count & 1) return 0;
    int32_t yi, yq;
    hb_fir(s, d->variant, &yi, &yq);
    return cascade_feed(d, k + 1, n_stages, yi, yq, oi, oq);
}

static inline int16_t scale_out(int32_t v, unsigned norm_shift, unsigned trunk_shift)
{
    /* `x << norm_shift >> trunk_shift` then the implicit int32 -> FixReal truncation of
     * IQSample::setReal (SDRDaemon.h:59). */
    uint32_t u = (uint32_t)v << norm_shift;
    return (int16_t)(uint16_t)(uint32_t)asr(u, trunk_shift);
}

size_t sdro_dec_process(sdro_dec* d, unsigned* sample_bits, const int16_t* in, size_t n_in, int16_t* out)
{
    const int M = d->log2_decim;
    unsigned ss = *sample_bits;

    if (M == 0) { /* Downsampler.cpp:76-80 + Decimators::decimate1 (Decimators.cpp:22-35) */
        if (ss < 16) {
            unsigned ns = 16 - ss;
            for (size_t k = 0; k < 2 * n_in; k++) out[k] = (int16_t)(uint16_t)((uint32_t)(int32_t)in[k] << ns);
        } else {
            memmove(out, in, n_in * 4);
        }
        return n_in;
    }

    /* common shift rule of every decimateN routine, e.g. Decimators.cpp:408-409,515 */
    const unsigned thresh = 16 - M;
    const unsigned trunk_shift = ss < thresh ? 0 : ss - thresh;
    const unsigned norm_shift = ss < thresh ? thresh - ss : 0;
    const size_t group = (size_t)1 << M;
    const size_t n_groups = n_in >= group ? n_in / group : 0;
    size_t n_out = 0;

    if (d->fcpos == SDRO_FC_CENTER) {
        for (size_t p = 0; p < n_groups * group; p++) {
            int32_t oi, oq;
            if (cascade_feed(d, 0, M, in[2 * p], in[2 * p + 1], &oi, &oq)) {
                out[2 * n_out] = scale_out(oi, norm_shift, trunk_shift);
                out[2 * n_out + 1] = scale_out(oq, norm_shift, trunk_shift);
                n_out++;
            }
        }
    } else if (M == 1) { /* decimate2_inf / decimate2_sup, Decimators.cpp:38-91 */
        /* the reference walks 4 input samples per iteration (two outputs) */
        size_t quads = n_in >= 4 ? n_in / 4 : 0;
        for (size_t g = 0; g < quads; g++) {
            const int16_t* s = in + 8 * g;
            int32_t r0 = s[0], i0 = s[1], r1 = s[2], i1 = s[3], r2 = s[4], i2 = s[5], r3 = s[6], i3 = s[7];
            int32_t x0, y0, x1, y1;
            if (d->fcpos == SDRO_FC_INFRA) {
                x0 = r0 - i1;  y0 = i0 + r1;
                x1 = i3 - r2;  y1 = -i2 - r3;
            } else {
                x0 = i0 - r1;  y0 = -r0 - i1;
                x1 = r3 - i2;  y1 = r2 + i3;
            }
            out[2 * n_out] = scale_out(x0, norm_shift, trunk_shift);
            out[2 * n_out + 1] = scale_out(y0, norm_shift, trunk_shift);
            out[2 * n_out + 2] = scale_out(x1, norm_shift, trunk_shift);
            out[2 * n_out + 3] = scale_out(y1, norm_shift, trunk_shift);
            n_out += 2;
        }
        /* out.resize(len/2): when len % 4 >= 2 the reference leaves one extra default
         * (zero) sample at the tail */
        if (n_in / 2 > n_out) { out[2 * n_out] = 0; out[2 * n_out + 1] = 0; n_out++; }
    } else { /* divide-by-4 rotate-and-sum prologue then M-2 half-band stages,
                Decimators.cpp:127-170 (M=2), :337-400 (M=3), :519-592, :808-899 ... */
        for (size_t g = 0; g < n_groups * (group / 4); g++) {
            const int16_t* s = in + 8 * g;
            int32_t r0 = s[0], i0 = s[1], r1 = s[2], i1 = s[3], r2 = s[4], i2 = s[5], r3 = s[6], i3 = s[7];
            int32_t x, y;
            if (d->fcpos == SDRO_FC_INFRA) {
                x = r0 - i1 + i3 - r2;
                y = i0 - i2 + r1 - r3;
            } else {
                x = i0 - r1 - i2 + r3;
                y = -r0 - i1 + r2 + i3;
            }
            int32_t oi, oq;
            if (cascade_feed(d, 0, M - 2, x, y, &oi, &oq)) {
                out[2 * n_out] = scale_out(oi, norm_shift, trunk_shift);
                out[2 * n_out + 1] = scale_out(oq, norm_shift, trunk_shift);
                n_out++;
            }
        }
    }
    *sample_bits = ss + (unsigned)M - trunk_shift;
    return n_out;
}

/* ======================================================================= GF(256) ===== */

/* Restated from the published cm256/gf256 algorithm (library absent from the reference tree):
 * polynomial index 3 of gf256's table, 0xa6 -> (0xa6 << 1) | 1 = 0x14D; exp[i] = 2^i. */
#define GF_POLY 0x14Du
static uint8_t GF_EXP[512 + 1];
static uint8_t GF_LOG[256];
static int gf_ready = 0;

static void gf_init(void)
{
    if (gf_ready) return;
    unsigned v = 1;
    for (int i = 0; i < 255; i++) {
        GF_EXP[i] = (uint8_t)v;
        GF_LOG[v] = (uint8_t)i;
        v <<= 1;
        if (v & 0x100) v ^= GF_POLY;
    }
    for (int i = 255; i < 513; i++) GF_EXP[i] = GF_EXP[i - 255];
    GF_LOG[0] = 0;
    gf_ready = 1;
}

uint8_t sdro_gf_exp(int i) { gf_init(); return GF_EXP[((i % 255) + 255) % 255]; }
uint8_t sdro_gf_log(uint8_t a) { gf_init(); return GF_LOG[a]; }
uint8_t sdro_gf_mul(uint8_t a, uint8_t b)
{
    gf_init();
    if (!a || !b) return 0;
    return GF_EXP[GF_LOG[a] + GF_LOG[b]];
}
uint8_t sdro_gf_div(uint8_t a, uint8_t b)
{
    gf_init();
    if (!a) return 0;          /* b == 0 never occurs on the paths below (x_i != y_j) */
    return GF_EXP[GF_LOG[a] + 255 - GF_LOG[b]];
}
uint8_t sdro_cm256_matrix_element(uint8_t x_i, uint8_t x_0, uint8_t y_j)
{
    return sdro_gf_div((uint8_t)(y_j ^ x_0), (uint8_t)(x_i ^ y_j));
}

/* cm256cc / gf256 multiply blocks with the SSSE3 byte shuffle, 16 bytes per instruction: two 16-entry tables per
 * multiplier (its products with the low and with the high nibble), z ^= lo[x & 15] ^ hi[x >> 4].  Same arithmetic as
 * the scalar table below -- used when the CPU has SSSE3 (as the library does), so that the CPU arm of bench.py is not
 * slower than the real thing; sdro_set_simd(0) forces the scalar path (tests compare the two). */
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define SDRO_HAVE_SSSE3_PATH 1
static int g_simd = -1; /* -1: not probed yet */
__attribute__((target("ssse3"))) static void gf_muladd_mem_ssse3(uint8_t* z, uint8_t c, const uint8_t* x, int n)
{
    uint8_t lo[16], hi[16];
    const unsigned lc = GF_LOG[c];
    lo[0] = hi[0] = 0;
    for (int v = 1; v < 16; v++) {
        lo[v] = GF_EXP[GF_LOG[v] + lc];
        hi[v] = GF_EXP[GF_LOG[v << 4] + lc];
    }
    const __m128i tlo = _mm_loadu_si128((const __m128i*)lo), thi = _mm_loadu_si128((const __m128i*)hi);
    const __m128i mask = _mm_set1_epi8(0x0F);
    int i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128i v = _mm_loadu_si128((const __m128i*)(x + i));
        const __m128i l = _mm_shuffle_epi8(tlo, _mm_and_si128(v, mask));
        const __m128i h = _mm_shuffle_epi8(thi, _mm_and_si128(_mm_srli_epi64(v, 4), mask));
        __m128i o = _mm_loadu_si128((const __m128i*)(z + i));
        o = _mm_xor_si128(o, _mm_xor_si128(l, h));
        _mm_storeu_si128((__m128i*)(z + i), o);
    }
    for (; i < n; i++) z[i] ^= (uint8_t)(lo[x[i] & 15] ^ hi[x[i] >> 4]);
}
#endif
void sdro_set_simd(int on)
{
#ifdef SDRO_HAVE_SSSE3_PATH
    g_simd = on && __builtin_cpu_supports("ssse3") ? 1 : 0;
#else
    (void)on;
#endif
}
int sdro_simd(void)
{
#ifdef SDRO_HAVE_SSSE3_PATH
    if (g_simd < 0) g_simd = __builtin_cpu_supports("ssse3") ? 1 : 0;
    return g_simd;
#else
    return 0;
#endif
}

static void gf_muladd_mem(uint8_t* z, uint8_t c, const uint8_t* x, int n)
{
    if (c == 0) return;
    if (c == 1) { for (int i = 0; i < n; i++) z[i] ^= x[i]; return; }
#ifdef SDRO_HAVE_SSSE3_PATH
    if (sdro_simd()) { gf_muladd_mem_ssse3(z, c, x, n); return; }
#endif
    uint8_t t[256];
    t[0] = 0;
    unsigned lc = GF_LOG[c];
    for (int v = 1; v < 256; v++) t[v] = GF_EXP[GF_LOG[v] + lc];
    for (int i = 0; i < n; i++) z[i] ^= t[x[i]];
}

static void gf_mul_mem(uint8_t* z, const uint8_t* x, uint8_t c, int n)
{
    memset(z, 0, (size_t)n);
    gf_muladd_mem(z, c, x, n);
}

/* cm256_encode: every recovery row r = 0..RecoveryCount-1 is
 *   rec[r] = XOR_j M[x_0 + r][j] * orig[j],   row 0 is plain XOR parity.
 * The original's position j in the array (not its Index) is the column, as upstream. */
int sdro_cm256_encode(sdro_cm256_params p, const sdro_cm256_block* originals, void* recovery)
{
    gf_init();
    if (p.OriginalCount <= 0 || p.RecoveryCount <= 0 || p.BlockBytes <= 0) return -1;
    if (p.OriginalCount + p.RecoveryCount > 256) return -2;
    if (!originals || !recovery) return -3;
    uint8_t* out = (uint8_t*)recovery;
    for (int r = 0; r < p.RecoveryCount; r++, out += p.BlockBytes) {
        if (p.OriginalCount == 1) { memcpy(out, originals[0].Block, (size_t)p.BlockBytes); continue; }
        const uint8_t x_0 = (uint8_t)p.OriginalCount;
        const uint8_t x_i = (uint8_t)(p.OriginalCount + r);
        if (r == 0) {
            memcpy(out, originals[0].Block, (size_t)p.BlockBytes);
            for (int j = 1; j < p.OriginalCount; j++) gf_muladd_mem(out, 1, (const uint8_t*)originals[j].Block, p.BlockBytes);
            continue;
        }
        gf_mul_mem(out, (const uint8_t*)originals[0].Block, sdro_cm256_matrix_element(x_i, x_0, 0), p.BlockBytes);
        for (int j = 1; j < p.OriginalCount; j++)
            gf_muladd_mem(out, sdro_cm256_matrix_element(x_i, x_0, (uint8_t)j), (const uint8_t*)originals[j].Block, p.BlockBytes);
    }
    return 0;
}

/* cm256_decode.  Blocks with Index < OriginalCount are originals, the others recovery rows
 * (kept in arrival order); erasures are the missing original indices in ascending order.
 * RecoveryCount == 1 takes upstream's parity shortcut: XOR of all received originals into the
 * lone recovery block *whatever its row is* (a latent mismatch with sdrdaemon passing "number
 * received" as RecoveryCount, SDRdaemonFECBuffer.cpp:176 -- replicated, not fixed).
 * Otherwise: subtract the known originals from every recovery row, then solve the N x N Cauchy
 * system.  Upstream factors it as LDU; the solution is unique, so Gaussian elimination over
 * GF(256) yields the same bytes. */
int sdro_cm256_decode(sdro_cm256_params p, sdro_cm256_block* blocks)
{
    gf_init();
    if (p.OriginalCount <= 0 || p.RecoveryCount <= 0 || p.BlockBytes <= 0) return -1;
    if (p.OriginalCount + p.RecoveryCount > 256) return -2;
    if (!blocks) return -3;
    if (p.OriginalCount == 1) { blocks[0].Index = 0; return 0; }

    sdro_cm256_block* orig[256];
    sdro_cm256_block* rec[256];
    uint8_t present[256];
    uint8_t erased[256];
    int n_orig = 0, n_rec = 0;
    memset(present, 0, sizeof(present));
    for (int i = 0; i < p.OriginalCount; i++) {
        int row = blocks[i].Index;
        if (row < p.OriginalCount) {
            if (present[row]) return -5; /* repeated original */
            present[row] = 1;
            orig[n_orig++] = &blocks[i];
        } else {
            rec[n_rec++] = &blocks[i];
        }
    }
    if (n_rec <= 0) return 0;
    for (int i = 0, k = 0; i < 256 && k < n_rec; i++)
        if (!present[i]) erased[k++] = (uint8_t)i;

    const int nb = p.BlockBytes;
    if (p.RecoveryCount == 1) {
        uint8_t* out = (uint8_t*)rec[0]->Block;
        for (int i = 0; i < n_orig; i++) gf_muladd_mem(out, 1, (const uint8_t*)orig[i]->Block, nb);
        rec[0]->Index = erased[0];
        return 0;
    }

    const int N = n_rec;
    const uint8_t x_0 = (uint8_t)p.OriginalCount;
    for (int o = 0; o < n_orig; o++)
        for (int r = 0; r < N; r++)
            gf_muladd_mem((uint8_t*)rec[r]->Block, sdro_cm256_matrix_element(rec[r]->Index, x_0, orig[o]->Index),
                          (const uint8_t*)orig[o]->Block, nb);

    /* A[r][c] = M[row of recovery r][erased c]; solve A * X = R in place (Gauss-Jordan). */
    uint8_t* A = (uint8_t*)malloc((size_t)N * N);
    for (int r = 0; r < N; r++)
        for (int c = 0; c < N; c++) A[r * N + c] = sdro_cm256_matrix_element(rec[r]->Index, x_0, erased[c]);
    uint8_t** R = (uint8_t**)malloc(sizeof(uint8_t*) * (size_t)N);
    for (int r = 0; r < N; r++) R[r] = (uint8_t*)rec[r]->Block;
    uint8_t* tmp = (uint8_t*)malloc((size_t)nb);
    int rc = 0;
    for (int c = 0; c < N && rc == 0; c++) {
        int piv = -1;
        for (int r = c; r < N; r++) if (A[r * N + c]) { piv = r; break; }
        if (piv < 0) { rc = -6; break; } /* cannot happen: Cauchy sub-matrices are non-singular */
        if (piv != c) {
            for (int k = 0; k < N; k++) { uint8_t t = A[c * N + k]; A[c * N + k] = A[piv * N + k]; A[piv * N + k] = t; }
            /* swap the *contents* so that block pointers keep their arrival order */
            memcpy(tmp, R[c], (size_t)nb); memcpy(R[c], R[piv], (size_t)nb); memcpy(R[piv], tmp, (size_t)nb);
        }
        uint8_t inv = sdro_gf_div(1, A[c * N + c]);
        if (inv != 1) {
            for (int k = 0; k < N; k++) A[c * N + k] = sdro_gf_mul(A[c * N + k], inv);
            gf_mul_mem(tmp, R[c], inv, nb); memcpy(R[c], tmp, (size_t)nb);
        }
        for (int r = 0; r < N; r++) {
            if (r == c) continue;
            uint8_t f = A[r * N + c];
            if (!f) continue;
            for (int k = 0; k < N; k++) A[r * N + k] ^= sdro_gf_mul(A[c * N + k], f);
            gf_muladd_mem(R[r], f, R[c], nb);
        }
    }
    free(tmp); free(R); free(A);
    if (rc) return rc;
    for (int r = 0; r < N; r++) rec[r]->Index = erased[r];
    return 0;
}

/* boost::crc_32_type: reflected 0xEDB88320, init and final xor 0xFFFFFFFF (UDPSinkFEC.cpp:106-109) */
static uint32_t crc_table[256];
static void crc_init(void)
{
    static volatile int ready = 0;
    if (ready) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        crc_table[i] = c;
    }
    ready = 1;
}
/* running form (zlib convention: start from 0, feed the previous result back in) */
uint32_t sdro_crc32_update(uint32_t crc, const void* data, size_t n)
{
    crc_init();
    uint32_t c = crc ^ 0xFFFFFFFFu;
    const uint8_t* p = (const uint8_t*)data;
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
uint32_t sdro_crc32(const void* data, size_t n) { return sdro_crc32_update(0u, data, n); }

/* ======================================================================= sink ======== */

struct sdro_sink {
    sdro_frame_cb cb;
    void* user;
    uint32_t center_freq_khz, sample_rate;
    uint8_t sample_bytes, sample_bits;
    int nb_fec;
    uint32_t tv_sec, tv_usec;
    uint8_t super_block[SDRO_UDPSIZE];             /* m_superBlock */
    uint8_t tx_blocks[256][SDRO_UDPSIZE];          /* one row of m_txBlocks */
    uint8_t fec[256 * SDRO_BLOCK_BYTES];
    int tx_block_index;                            /* m_txBlockIndex */
    int sample_index;                              /* m_sampleIndex */
    uint16_t frame_count;                          /* m_frameCount */
};

sdro_sink* sdro_sink_create(sdro_frame_cb cb, void* user)
{
    gf_init();
    sdro_sink* s = (sdro_sink*)calloc(1, sizeof(*s));
    s->cb = cb;
    s->user = user;
    s->sample_bytes = 2;   /* UDPSink defaults are set by the mains through the setters */
    s->sample_bits = 16;
    return s;
}
void sdro_sink_destroy(sdro_sink* s) { free(s); }
void sdro_sink_set_meta(sdro_sink* s, uint32_t f, uint32_t r, uint8_t sb, uint8_t bits)
{
    s->center_freq_khz = f; s->sample_rate = r; s->sample_bytes = sb; s->sample_bits = bits;
}
void sdro_sink_set_nb_fec(sdro_sink* s, int nb_fec) { s->nb_fec = nb_fec; }
void sdro_sink_set_time(sdro_sink* s, uint32_t sec, uint32_t usec) { s->tv_sec = sec; s->tv_usec = usec; }

static void put_u32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

static void sink_finish_frame(sdro_sink* s)
{
    /* encode half of UDPSinkFEC::transmitUDP, UDPSinkFEC.cpp:217-256 */
    const int F = s->nb_fec;
    int n_blocks = SDRO_NB_ORIGINAL;
    if (F > 0) {
        sdro_cm256_params p = {SDRO_NB_ORIGINAL, F, SDRO_BLOCK_BYTES};
        sdro_cm256_block desc[256];
        uint8_t* fec = s->fec;           /* transmitUDP's local fecBlocks[256], UDPSinkFEC.cpp:197 */
        for (int i = 0; i < SDRO_NB_ORIGINAL + F; i++) {
            if (i >= SDRO_NB_ORIGINAL) memset(s->tx_blocks[i], 0, SDRO_UDPSIZE); /* filler byte: the
                reference leaves it uninitialised (:233-243); zero here, masked in comparisons */
            s->tx_blocks[i][0] = (uint8_t)(s->frame_count & 0xFF);
            s->tx_blocks[i][1] = (uint8_t)(s->frame_count >> 8);
            s->tx_blocks[i][2] = (uint8_t)i;
            desc[i].Block = s->tx_blocks[i] + 4;
            desc[i].Index = (uint8_t)i;
        }
        if (sdro_cm256_encode(p, desc, fec) == 0) {
            for (int i = 0; i < F; i++) memcpy(s->tx_blocks[SDRO_NB_ORIGINAL + i] + 4, fec + (size_t)i * SDRO_BLOCK_BYTES, SDRO_BLOCK_BYTES);
            n_blocks += F;
        }
    }
    if (s->cb) s->cb(s->user, &s->tx_blocks[0][0], n_blocks, s->frame_count);
}

void sdro_sink_write(sdro_sink* s, const int16_t* iq, size_t n)
{
    size_t it = 0;
    while (it < n) {
        size_t remaining = n - it;
        if (s->tx_block_index == 0) { /* block 0 carries only meta data, UDPSinkFEC.cpp:89-136 */
            uint8_t meta[24];
            put_u32(meta + 0, s->center_freq_khz);
            put_u32(meta + 4, s->sample_rate);
            meta[8] = s->sample_bytes;
            meta[9] = s->sample_bits;
            meta[10] = SDRO_NB_ORIGINAL;
            meta[11] = (uint8_t)s->nb_fec;
            put_u32(meta + 12, s->tv_sec);
            put_u32(meta + 16, s->tv_usec);
            put_u32(meta + 20, sdro_crc32(meta, 20));
            memset(s->super_block, 0, SDRO_UDPSIZE);
            s->super_block[0] = (uint8_t)(s->frame_count & 0xFF);
            s->super_block[1] = (uint8_t)(s->frame_count >> 8);
            s->super_block[2] = 0;
            memcpy(s->super_block + 4, meta, 24);
            memcpy(s->tx_blocks[0], s->super_block, SDRO_UDPSIZE);
            s->tx_block_index = 1;
        }
        if ((size_t)s->sample_index + remaining < SDRO_SAMPLES_PER_BLK) { /* :138-144 */
            memcpy(s->super_block + 4 + 4 * s->sample_index, iq + 2 * it, remaining * 4);
            s->sample_index += (int)remaining;
            it = n;
        } else { /* :145-190 */
            size_t take = (size_t)(SDRO_SAMPLES_PER_BLK - s->sample_index);
            memcpy(s->super_block + 4 + 4 * s->sample_index, iq + 2 * it, take * 4);
            it += take;
            s->sample_index = 0;
            s->super_block[0] = (uint8_t)(s->frame_count & 0xFF);
            s->super_block[1] = (uint8_t)(s->frame_count >> 8);
            s->super_block[2] = (uint8_t)s->tx_block_index;
            memcpy(s->tx_blocks[s->tx_block_index], s->super_block, SDRO_UDPSIZE);
            if (s->tx_block_index == SDRO_NB_ORIGINAL - 1) {
                sink_finish_frame(s);
                s->tx_block_index = 0;
                s->frame_count++;
            } else {
                s->tx_block_index++;
            }
        }
    }
}

/* ======================================================================= source ====== */

typedef struct {
    int sum, idx, n;
    int hist[10];
} mavg10; /* MovingAverage<int,int,10>, include/MovingAverage.h:27-58 (stats only) */

struct sdro_fecbuf {
    uint8_t frame[SDRO_NB_ORIGINAL][SDRO_BLOCK_BYTES];          /* m_decoderSlot.m_frame */
    uint8_t recovery[SDRO_NB_ORIGINAL][SDRO_BLOCK_BYTES];       /* m_recoveryBlocks */
    sdro_cm256_block desc[SDRO_NB_ORIGINAL];                    /* m_cm256DescriptorBlocks */
    int block_count, recovery_count, decoded, meta_retrieved;
    int frame_head;
    int cur_nb_blocks, cur_nb_recovery, min_nb_blocks, max_nb_recovery;
    uint8_t current_meta[20], output_meta[20];
};

sdro_fecbuf* sdro_fecbuf_create(void)
{
    sdro_fecbuf* b = (sdro_fecbuf*)calloc(1, sizeof(*b));
    b->frame_head = -1;
    b->min_nb_blocks = 256;
    b->current_meta[11] = 0xFF; /* MetaDataFEC::init sets m_nbFECBlocks = -1 */
    b->output_meta[11] = 0xFF;
    return b;
}
void sdro_fecbuf_destroy(sdro_fecbuf* b) { free(b); }
int sdro_fecbuf_cur_nb_blocks(const sdro_fecbuf* b) { return b->cur_nb_blocks; }
int sdro_fecbuf_cur_nb_recovery(const sdro_fecbuf* b) { return b->cur_nb_recovery; }
int sdro_fecbuf_min_nb_blocks(sdro_fecbuf* b) { int v = b->min_nb_blocks; b->min_nb_blocks = 256; return v; }
int sdro_fecbuf_max_nb_recovery(sdro_fecbuf* b) { int v = b->max_nb_recovery; b->max_nb_recovery = 0; return v; }
void sdro_fecbuf_current_meta(const sdro_fecbuf* b, uint8_t meta20[20]) { memcpy(meta20, b->current_meta, 20); }

/* store + decode steps shared by the streaming buffer and the one-shot helper;
 * SDRdaemonFECBuffer.cpp:143-247 */
static int slot_store(sdro_fecbuf* b, const uint8_t* sb)
{
    int status = 0;
    if (b->block_count < SDRO_NB_ORIGINAL) {
        int bc = b->block_count, rcnt = b->recovery_count;
        int block_index = sb[2];
        b->desc[bc].Index = (uint8_t)block_index;
        if (block_index == 0) b->meta_retrieved = 1;
        if (block_index < SDRO_NB_ORIGINAL) {
            memcpy(b->frame[block_index], sb + 4, SDRO_BLOCK_BYTES);
            b->desc[bc].Block = b->frame[block_index];
        } else {
            memcpy(b->recovery[rcnt], sb + 4, SDRO_BLOCK_BYTES);
            b->desc[bc].Block = b->recovery[rcnt];
            b->recovery_count++;
        }
    }
    b->block_count++;
    if (b->block_count == SDRO_NB_ORIGINAL) {
        b->decoded = 1;
        status = 1;
        if (b->recovery_count > 0) {
            sdro_cm256_params p = {SDRO_NB_ORIGINAL, b->recovery_count, SDRO_BLOCK_BYTES};
            if (sdro_cm256_decode(p, b->desc)) {
                status = -1;
            } else {
                status = 2;
                for (int ir = 0; ir < b->recovery_count; ir++) { /* :208-213 */
                    int ri = SDRO_NB_ORIGINAL - b->recovery_count + ir;
                    int bi = b->desc[ri].Index;
                    if (b->desc[ri].Block != (void*)b->frame[bi])
                        memcpy(b->frame[bi], b->desc[ri].Block, SDRO_BLOCK_BYTES);
                }
            }
        }
        if (b->meta_retrieved && memcmp(b->frame[0], b->current_meta, 12) != 0)
            memcpy(b->current_meta, b->frame[0], 20);
    }
    return status;
}

static void slot_init(sdro_fecbuf* b)
{
    /* initDecodeSlot, SDRdaemonFECBuffer.cpp:95-110 */
    b->cur_nb_blocks = b->block_count;
    b->cur_nb_recovery = b->recovery_count;
    if (b->cur_nb_blocks < b->min_nb_blocks) b->min_nb_blocks = b->cur_nb_blocks;
    if (b->cur_nb_recovery > b->max_nb_recovery) b->max_nb_recovery = b->cur_nb_recovery;
    b->block_count = 0;
    b->recovery_count = 0;
    b->decoded = 0;
    b->meta_retrieved = 0;
    memset(b->frame, 0, sizeof(b->frame));
}

int sdro_fecbuf_write_and_read(sdro_fecbuf* b, const uint8_t* sb, uint8_t* data, size_t* data_len)
{
    int available = 0;
    *data_len = 0;
    int frame_index = sb[0] | (sb[1] << 8);
    if (b->frame_head != frame_index) { /* :133-139 + getSlotData :72-93 */
        *data_len = (size_t)(SDRO_NB_ORIGINAL - 1) * SDRO_BLOCK_BYTES;
        memcpy(data, b->frame[1], *data_len);
        if (b->meta_retrieved && memcmp(b->frame[0], b->output_meta, 12) != 0) memcpy(b->output_meta, b->frame[0], 20);
        available = 1;
        slot_init(b);
        b->frame_head = frame_index;
    }
    slot_store(b, sb);
    return available;
}

int sdro_decode_frame(const uint8_t* superblocks, int n_blocks, uint8_t* payload, uint8_t* block0)
{
    sdro_fecbuf* b = sdro_fecbuf_create();
    int status = 0;
    for (int i = 0; i < n_blocks; i++) {
        int st = slot_store(b, superblocks + (size_t)i * SDRO_UDPSIZE);
        if (st) status = st;
    }
    memcpy(payload, b->frame[1], (size_t)(SDRO_NB_ORIGINAL - 1) * SDRO_BLOCK_BYTES);
    if (block0) memcpy(block0, b->frame[0], SDRO_BLOCK_BYTES);
    sdro_fecbuf_destroy(b);
    return status;
}

/* ======================================================================= bench leg ==== */
typedef struct {
    const uint8_t* sb;
    size_t pitch;
    const int* n_blocks;
    int n_frames;
    uint8_t *payload, *block0;
    int* status;
    volatile int* next;
} dec_job;
static void* dec_worker(void* arg);

/* Whole Rx hot path (decimate -> pack -> encode) for n_streams streams on n_threads POSIX threads:
 * the "port" CPU baseline of bench.py when the reference build (oracle/_ref) is not available. */
#include <pthread.h>

typedef struct {
    int log2_decim, fcpos, variant, nb_fec, n_streams;
    const int16_t* in;
    size_t n_in, stride, block;
    volatile int* next;
    long long frames;
    uint32_t digest;
    uint32_t* stream_crc; /* optional: CRC-32 of every stream's datagram bytes, in send order */
} rx_job;

typedef struct { long long frames; uint32_t digest; uint32_t crc; int want_crc; } rx_acc;
static void rx_cb(void* user, const uint8_t* dg, int n_blocks, uint16_t fi)
{
    (void)fi;
    rx_acc* a = (rx_acc*)user;
    const uint32_t* w = (const uint32_t*)dg;
    uint32_t d = 0;
    for (int i = 0; i < n_blocks * 128; i++) d ^= w[i];
    a->frames++;
    a->digest ^= d;
    if (a->want_crc) a->crc = sdro_crc32_update(a->crc, dg, (size_t)n_blocks * SDRO_UDPSIZE);
}

static void* rx_worker(void* arg)
{
    rx_job* j = (rx_job*)arg;
    int16_t* out = (int16_t*)malloc(j->block * 4 + 16);
    for (;;) {
        int s = __sync_fetch_and_add(j->next, 1);
        if (s >= j->n_streams) break;
        sdro_dec* d = sdro_dec_create(j->log2_decim, j->fcpos, j->variant);
        rx_acc acc = {0, 0, 0, j->stream_crc != NULL};
        sdro_sink* k = sdro_sink_create(rx_cb, &acc);
        sdro_sink_set_meta(k, 435000, 625000, 2, 16);
        sdro_sink_set_nb_fec(k, j->nb_fec);
        sdro_sink_set_time(k, 1700000000u, 0);
        const int16_t* in = j->in + (size_t)s * j->stride * 2;
        size_t done = 0;
        while (done < j->n_in) {
            size_t n = j->n_in - done < j->block ? j->n_in - done : j->block;
            unsigned ss = 16;
            size_t no = sdro_dec_process(d, &ss, in + 2 * done, n, out);
            sdro_sink_write(k, out, no);
            done += n;
        }
        sdro_sink_destroy(k);
        sdro_dec_destroy(d);
        j->frames += acc.frames;
        j->digest ^= acc.digest;
        if (j->stream_crc) j->stream_crc[s] = acc.crc;
    }
    free(out);
    return NULL;
}

long long sdro_rx_streams(int log2_decim, int fcpos, int variant, int nb_fec, int n_streams, int n_threads,
                          const int16_t* iq_in, size_t n_in_per_stream, size_t in_stride, size_t block,
                          uint32_t* digest)
{
    return sdro_rx_streams_crc(log2_decim, fcpos, variant, nb_fec, n_streams, n_threads, iq_in, n_in_per_stream, in_stride,
                               block, digest, NULL);
}

long long sdro_rx_streams_crc(int log2_decim, int fcpos, int variant, int nb_fec, int n_streams, int n_threads,
                              const int16_t* iq_in, size_t n_in_per_stream, size_t in_stride, size_t block,
                              uint32_t* digest, uint32_t* stream_crc)
{
    gf_init();
    crc_init();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    rx_job jobs[256];
    volatile int next = 0;
    for (int t = 0; t < n_threads; t++) {
        rx_job j = {log2_decim, fcpos, variant, nb_fec, n_streams, iq_in, n_in_per_stream, in_stride, block, &next, 0, 0, stream_crc};
        jobs[t] = j;
        pthread_create(&th[t], NULL, rx_worker, &jobs[t]);
    }
    long long frames = 0;
    uint32_t dg = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        frames += jobs[t].frames;
        dg ^= jobs[t].digest;
    }
    if (digest) *digest = dg;
    return frames;
}
static void* dec_worker(void* arg)
{
    dec_job* j = (dec_job*)arg;
    for (;;) {
        int f = __sync_fetch_and_add(j->next, 1);
        if (f >= j->n_frames) break;
        j->status[f] = sdro_decode_frame(j->sb + (size_t)f * j->pitch * SDRO_UDPSIZE, j->n_blocks[f],
                                         j->payload + (size_t)f * 127 * SDRO_BLOCK_BYTES,
                                         j->block0 ? j->block0 + (size_t)f * SDRO_BLOCK_BYTES : NULL);
    }
    return NULL;
}

/* sdro_decode_frame over a batch of frames (frame f: n_blocks[f] datagrams at sb + f * pitch * 512) on n_threads
 * threads: the all-frames checker of BASELINE config 4 and its CPU baseline */
void sdro_decode_frames(const uint8_t* sb, size_t pitch, const int* n_blocks, int n_frames, int n_threads, uint8_t* payload,
                        uint8_t* block0, int* status)
{
    gf_init();
    crc_init();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    volatile int next = 0;
    dec_job j = {sb, pitch, n_blocks, n_frames, payload, block0, status, &next};
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, dec_worker, &j);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------- interpolator ---- */

/* HBFIRFilterTraits<32>/<16>::hbCoeffs as integers (sdmnbase/HBFilterTraits.cpp:62-72 and :25-31,
 * (int32_t)(c * (1 << 14)) truncated toward zero; values printed by a probe against the reference). */
static const int32_t HB32[8] = {-30, 63, -135, 261, -469, 830, -1605, 5176};
static const int32_t HB16[4] = {-85, 380, -1246, 5041};

/* One interpolating half-band stage: IntHalfbandFilterEO1<N>::myInterpolate + doInterpolateFIR
 * (include/IntHalfbandFilterEO1.h:44-65,149-168; IntHalfbandFilterDB.h:51-72,109-127 is the same
 * arithmetic).  Ring of N/2 samples, double-buffered so that a window never wraps. */
typedef struct {
    int order;                /* 64, 32 or 16 */
    int ptr;
    int32_t smp[64][2];
} int_stage;

static void int_stage_reset(int_stage* st, int order)
{
    memset(st, 0, sizeof *st);
    st->order = order;
}

/* in: (x1, y1); out: first sample (x1, y1) = the middle of the window, second (x2, y2) = the FIR */
static void int_stage_run(int_stage* st, int32_t* x1, int32_t* y1, int32_t* x2, int32_t* y2)
{
    const int half = st->order / 2, quarter = st->order / 4;
    const int32_t* co = st->order == 64 ? HB64 : st->order == 32 ? HB32 : HB16;
    st->smp[st->ptr][0] = *x1;
    st->smp[st->ptr][1] = *y1;
    st->smp[st->ptr + half][0] = *x1;
    st->smp[st->ptr + half][1] = *y1;
    st->ptr = st->ptr < half - 1 ? st->ptr + 1 : 0;
    *x1 = st->smp[st->ptr + quarter - 1][0];
    *y1 = st->smp[st->ptr + quarter - 1][1];
    uint32_t ia = 0, qa = 0; /* wrapping int32 arithmetic */
    int a = st->ptr, b = st->ptr + half - 1;
    for (int i = 0; i < quarter; i++, a++, b--) {
        ia += ((uint32_t)st->smp[a][0] + (uint32_t)st->smp[b][0]) * (uint32_t)co[i];
        qa += ((uint32_t)st->smp[a][1] + (uint32_t)st->smp[b][1]) * (uint32_t)co[i];
    }
    *x2 = (int32_t)ia >> 13; /* hbShift - 1 */
    *y2 = (int32_t)qa >> 13;
}

struct sdro_int {
    int log2_interp;
    int_stage st[6];
};

sdro_int* sdro_int_create(int log2_interp)
{
    if (log2_interp < 0 || log2_interp > 6) return NULL; /* Upsampler.cpp:38-42 */
    sdro_int* u = (sdro_int*)calloc(1, sizeof *u);
    if (!u) return NULL;
    u->log2_interp = log2_interp;
    sdro_int_reset(u);
    return u;
}
void sdro_int_destroy(sdro_int* u) { free(u); }
/* Upsampler::configure (Upsampler.cpp:32-55): only m_interp changes; the stage objects of
 * m_interpolators persist (Interpolators.h:52-58). */
int sdro_int_configure(sdro_int* u, int log2_interp)
{
    if (log2_interp < 0 || log2_interp > 6) return -1;
    u->log2_interp = log2_interp;
    return 0;
}
void sdro_int_reset(sdro_int* u)
{
    static const int orders[6] = {64, 32, 16, 16, 16, 16}; /* Interpolators.h:31-33 */
    for (int i = 0; i < 6; i++) int_stage_reset(&u->st[i], orders[i]);
}

/* stage s (0-based) expands buf[0..n) (interleaved int32 I/Q pairs at stride `step` pairs) in place the
 * way interpolateN_cen threads its intbuf through the stages (Interpolators.cpp:60-71 etc.): the sample
 * at slot k gets its second output at slot k + step/2. */
size_t sdro_int_process(sdro_int* u, const int16_t* iq_in, size_t n_in, int16_t* iq_out)
{
    const int M = u->log2_interp;
    if (M == 0) { /* samples_out = samples_in */
        memcpy(iq_out, iq_in, n_in * 4);
        return n_in;
    }
    /* Reference quirk, reproduced: interpolate64_cen (Interpolators.cpp:363-605) runs only the five
     * stages of interpolate32_cen into intbuf[0..63] and then emits intbuf[0..127], whose upper half
     * is zeroed once before the loop (:370) and never written: per input sample 32 interpolated
     * samples followed by 32 zero samples.  m_interpolator64 is never used. */
    const int S = M == 6 ? 5 : M;   /* stages actually run */
    const int W = 1 << S;           /* samples they produce per input sample */
    const int WO = 1 << M;          /* samples emitted per input sample */
    int32_t buf[64][2];
    memset(buf, 0, sizeof buf);
    for (size_t k = 0; k < n_in; k++) {
        buf[0][0] = iq_in[2 * k];
        buf[0][1] = iq_in[2 * k + 1];
        for (int s = 0; s < S; s++) {
            const int step = W >> s; /* distance between the samples stage s consumes */
            for (int j = 0; j < W; j += step)
                int_stage_run(&u->st[s], &buf[j][0], &buf[j][1], &buf[j + step / 2][0], &buf[j + step / 2][1]);
        }
        for (int j = 0; j < WO; j++) { /* IQSample::setReal/setImag take int16_t: truncation */
            iq_out[2 * (k * WO + j)] = j < W ? (int16_t)buf[j][0] : (int16_t)0;
            iq_out[2 * (k * WO + j) + 1] = j < W ? (int16_t)buf[j][1] : (int16_t)0;
        }
    }
    return n_in << M;
}


