/*
 * sdrd_oracle.h -- CPU restatement of sdrdaemon's decimate + FEC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker (or as the CPU arm that is timed beside the GPU path).
 *
 * Parity status
 *   - decimator half (sdro_dec_*): PINNED.  tests/test_oracle_vs_ref.py checks it byte-for-byte
 *     against the reference's own Decimators.cpp/HBFilterTraits.cpp compiled from
 *     /root/reference into oracle/_ref (both the EO1 and the DB build), and against the
 *     committed vectors in tests/golden/ generated from that build.
 *   - framing half (sdro_sink_*, sdro_fecbuf_*): PINNED against the reference's own
 *     UDPSinkFEC.cpp / SDRdaemonFECBuffer.cpp compiled in oracle/_ref *with the restated
 *     CM256 below standing in for cm256cc*.
 *   - GF(256)/CM256 arithmetic (sdro_gf_*, sdro_cm256_*): PARITY UNPINNED.  cm256cc
 *     (f4exb/cm256cc, a C++ class port of catid/cm256; no version pinned by the reference,
 *     located via find_package at reference CMakeLists.txt:97) is not in /root/reference and
 *     the reference holds no vectors for it.  The arithmetic is restated from the published
 *     algorithm: GF(2^8) with polynomial 0x14D, generator 2, Cauchy element
 *     M[x_i][y_j] = (y_j ^ x_0) / (x_i ^ y_j), x_0 = OriginalCount.  It is anchored on the
 *     reference's call sites (sdmnbase/UDPSinkFEC.cpp:228-246, SDRdaemonFECBuffer.cpp:143-213)
 *     and on algebraic properties (MDS round trips, row-128 = XOR parity).  The field
 *     arithmetic alone (products, quotients, exp/log, matrix elements given the polynomial
 *     and the formula) is checked against sympy's GF(2)[x] routines in tests/.
 */
#ifndef SDRD_ORACLE_H
#define SDRD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- decimator ---------- */

enum { SDRO_FC_INFRA = 0, SDRO_FC_SUPRA = 1, SDRO_FC_CENTER = 2 };  /* Downsampler.h fcPos_t */
enum { SDRO_HB_EO1 = 0, SDRO_HB_DB = 1 };                            /* Decimators.h:56-70  */

typedef struct sdro_dec sdro_dec;

/* One Downsampler+Decimators state (six persistent half-band stages, Decimators.h:57-62). */
sdro_dec* sdro_dec_create(int log2_decim, int fcpos, int variant);
void      sdro_dec_destroy(sdro_dec* d);
void      sdro_dec_reset(sdro_dec* d);
/* Downsampler::configure (Downsampler.cpp:32-67) between blocks: stage states persist. 0 = ok. */
int       sdro_dec_configure(sdro_dec* d, int log2_decim, int fcpos);
/* Downsampler::process (Downsampler.cpp:74-162).  iq_in: n_in interleaved {I,Q} int16 samples.
 * Writes n_in >> log2_decim samples to iq_out, returns that count.  *sample_bits is the
 * reference's in/out sampleSize. Samples beyond the last whole group of 2^M are dropped
 * exactly as the reference's loop bounds do (Decimators.cpp:412). */
size_t    sdro_dec_process(sdro_dec* d, unsigned* sample_bits, const int16_t* iq_in, size_t n_in,
                           int16_t* iq_out);

/* ---------------------------------------------------------------- interpolator ---- */

typedef struct sdro_int sdro_int;

/* One Upsampler+Interpolators state: up to six persistent half-band stages of orders 64, 32, 16, 16,
 * 16, 16 (Interpolators.h:31-33,52-58). */
sdro_int* sdro_int_create(int log2_interp);
void      sdro_int_destroy(sdro_int* u);
void      sdro_int_reset(sdro_int* u);
/* Upsampler::configure (Upsampler.cpp:32-55) between blocks: stage states persist. 0 = ok. */
int       sdro_int_configure(sdro_int* u, int log2_interp);
/* Upsampler::process (Upsampler.cpp:57-84) -> Interpolators::interpolate{2..64}_cen
 * (Interpolators.cpp:23-606).  Writes n_in << log2_interp samples, returns that count. */
size_t    sdro_int_process(sdro_int* u, const int16_t* iq_in, size_t n_in, int16_t* iq_out);

/* ---------------------------------------------------------------- GF(256) / CM256 ---- */

uint8_t sdro_gf_mul(uint8_t a, uint8_t b);
uint8_t sdro_gf_div(uint8_t a, uint8_t b);
uint8_t sdro_gf_exp(int i);              /* 2^i */
uint8_t sdro_gf_log(uint8_t a);          /* a != 0 */
uint8_t sdro_cm256_matrix_element(uint8_t x_i, uint8_t x_0, uint8_t y_j);
/* block multiply-add: SSSE3 byte shuffle (16 bytes per instruction, what cm256cc uses on x86) when the CPU has it,
 * else the scalar 256-entry table; sdro_set_simd(0) forces the scalar path, sdro_simd() tells which is active */
void sdro_set_simd(int on);
int sdro_simd(void);

typedef struct {
    void*   Block;
    uint8_t Index;
} sdro_cm256_block;                      /* CM256::cm256_block as used at UDPSinkFEC.cpp:241-242 */

typedef struct {
    int OriginalCount;
    int RecoveryCount;
    int BlockBytes;
} sdro_cm256_params;                     /* CM256::cm256_encoder_params, UDPSinkFEC.cpp:228-230 */

/* recovery: RecoveryCount * BlockBytes contiguous output bytes. 0 = success. */
int sdro_cm256_encode(sdro_cm256_params p, const sdro_cm256_block* originals, void* recovery);
/* In place; recovered originals replace the recovery blocks' data, Index rewritten. 0 = success. */
int sdro_cm256_decode(sdro_cm256_params p, sdro_cm256_block* blocks);

uint32_t sdro_crc32(const void* data, size_t n);   /* boost::crc_32_type */

/* ---------------------------------------------------------------- sink framing ------- */

#define SDRO_UDPSIZE          512
#define SDRO_NB_ORIGINAL      128
#define SDRO_SAMPLES_PER_BLK  127
#define SDRO_BLOCK_BYTES      508
#define SDRO_FRAME_SAMPLES    (127 * 127)

typedef struct sdro_sink sdro_sink;
/* Called with every completed frame: (128 + n_fec) * 512 bytes of datagram images in send order. */
typedef void (*sdro_frame_cb)(void* user, const uint8_t* datagrams, int n_blocks, uint16_t frame_index);

sdro_sink* sdro_sink_create(sdro_frame_cb cb, void* user);
void       sdro_sink_destroy(sdro_sink* s);
void       sdro_sink_set_meta(sdro_sink* s, uint32_t center_freq_khz, uint32_t sample_rate,
                              uint8_t sample_bytes, uint8_t sample_bits);
void       sdro_sink_set_nb_fec(sdro_sink* s, int nb_fec);
/* Stand-in for gettimeofday at UDPSinkFEC.cpp:95: stamp used for the next block 0 created. */
void       sdro_sink_set_time(sdro_sink* s, uint32_t tv_sec, uint32_t tv_usec);
/* UDPSinkFEC::write (UDPSinkFEC.cpp:79-191) + the encode half of transmitUDP (:193-256). */
void       sdro_sink_write(sdro_sink* s, const int16_t* iq, size_t n_samples);

/* ---------------------------------------------------------------- source FEC buffer -- */

typedef struct sdro_fecbuf sdro_fecbuf;
sdro_fecbuf* sdro_fecbuf_create(void);
void         sdro_fecbuf_destroy(sdro_fecbuf* b);
/* SDRdaemonFECBuffer::writeAndRead (SDRdaemonFECBuffer.cpp:112-250).  Returns 1 when a frame
 * (127*508 bytes) was copied to data, with *data_len set. */
int          sdro_fecbuf_write_and_read(sdro_fecbuf* b, const uint8_t* superblock, uint8_t* data,
                                        size_t* data_len);
int          sdro_fecbuf_cur_nb_blocks(const sdro_fecbuf* b);
int          sdro_fecbuf_cur_nb_recovery(const sdro_fecbuf* b);
int          sdro_fecbuf_min_nb_blocks(sdro_fecbuf* b);     /* resets like the reference getter */
int          sdro_fecbuf_max_nb_recovery(sdro_fecbuf* b);   /* resets like the reference getter */
void         sdro_fecbuf_current_meta(const sdro_fecbuf* b, uint8_t meta20[20]);

/* One-shot helper for batch parity tests: the decode a single frame's first 128 received
 * superblocks go through (store, cm256_decode, copy-back), returning blocks 1..127.
 * n_blocks < 128 leaves the frame undecoded (missing blocks are zero). status: 0 = not decoded,
 * 1 = complete without recovery, 2 = recovered, -1 = cm256 error. */
int sdro_decode_frame(const uint8_t* superblocks, int n_blocks, uint8_t* payload /*127*508*/,
                      uint8_t* block0 /*508, may be NULL*/);

/* bench.py "port" CPU baseline: decimate -> pack -> encode for n_streams streams on n_threads threads.
 * Returns superframes produced; *digest = XOR of all datagram words. */
long long sdro_rx_streams(int log2_decim, int fcpos, int variant, int nb_fec, int n_streams, int n_threads,
                          const int16_t* iq_in, size_t n_in_per_stream, size_t in_stride, size_t block,
                          uint32_t* digest);
/* the same with, optionally, stream_crc[s] = CRC-32 (zlib convention) of stream s's datagram bytes in send order:
 * the every-stream checker of the full-size parity tests (BASELINE configs 3 and 5) */
long long sdro_rx_streams_crc(int log2_decim, int fcpos, int variant, int nb_fec, int n_streams, int n_threads,
                              const int16_t* iq_in, size_t n_in_per_stream, size_t in_stride, size_t block,
                              uint32_t* digest, uint32_t* stream_crc);
uint32_t sdro_crc32_update(uint32_t crc, const void* data, size_t n);
/* sdro_decode_frame over n_frames frames on n_threads threads (frame f: n_blocks[f] datagrams at
 * superblocks + f * pitch * 512): the every-frame checker and CPU baseline of BASELINE config 4 */
void sdro_decode_frames(const uint8_t* superblocks, size_t pitch, const int* n_blocks, int n_frames, int n_threads,
                        uint8_t* payload, uint8_t* block0, int* status);

#ifdef __cplusplus
}
#endif
#endif
