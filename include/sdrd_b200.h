/*
 * sdrd_b200.h -- C ABI of libsdrd_b200.so, the B200 (sm_100a) implementation of sdrdaemon's
 * decimate + FEC hot path.  Plain pointers and sizes only; every entry point states the reference
 * interface it stands in for (paths relative to f4exb/sdrdaemon v3.1.2).
 *
 * Conventions
 *   - IQ samples are the reference's `IQSample` (include/SDRDaemon.h:52-68): packed little-endian
 *     {int16 re, int16 im}, 4 bytes; an `IQSampleVector` is passed as (pointer, sample count).
 *   - "streams" are independent instances of the same reference object (one Downsampler /
 *     UDPSinkFEC each) processed in one call; stream s starts `stride` samples after stream s-1.
 *   - All functions return 0 on success and a negative SDRD_E* code on failure;
 *     sdrd_last_error() returns the message of the calling thread's last failure (the reference
 *     objects keep `std::string m_error`, include/Downsampler.h:65-76, include/UDPSink.h:85-108).
 *   - Host entry points take HOST pointers and perform the host<->device copies themselves.
 *     `_dev` entry points work on the handle's device buffers (for callers that already keep the
 *     stream in HBM) and take the CUDA stream as `void*` (cudaStream_t, may be NULL).
 *   - There is no CPU fallback: without a usable sm_100 device every create call fails with
 *     SDRD_ENODEV.
 *   - Errors: a call refused with SDRD_EINVAL / SDRD_ERANGE has not changed the handle (arguments are
 *     checked before any state moves).  After SDRD_ECUDA / SDRD_ENOMEM the handle's stream state is
 *     undefined: call the handle's reset function (or destroy it) before using it again.
 *   - Devices and threads: a handle belongs to the device that was current when it was created; its entry
 *     points may be called from any thread whatever that thread's current device (they switch and switch
 *     back).  A handle is not re-entrant: one call at a time per handle, as for the reference's objects.
 */
#ifndef SDRD_B200_H
#define SDRD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDRD_OK 0
#define SDRD_EINVAL (-1)  /* invalid argument (Downsampler.cpp:39-43,57-61; cm256 count checks) */
#define SDRD_ENODEV (-2)  /* no CUDA device / wrong architecture */
#define SDRD_ECUDA (-3)   /* CUDA runtime error, see sdrd_last_error() */
#define SDRD_ENOMEM (-4)
#define SDRD_ERANGE (-5)  /* call exceeds the capacity given at create time */

/* Downsampler::fcPos_t, include/Downsampler.h:31-35 */
#define SDRD_FC_INFRA 0
#define SDRD_FC_SUPRA 1
#define SDRD_FC_CENTER 2
/* which IntHalfbandFilter the reference was built with, include/Decimators.h:56-70 */
#define SDRD_HB_EO1 0 /* USE_SSE4_1 builds (x86): IntHalfbandFilterEO1<64> */
#define SDRD_HB_DB 1  /* other builds: IntHalfbandFilterDB<64> (+1 rounding on the centre tap) */

/* Wire format constants, include/UDPSinkFEC.h:34-36,77-121 */
#define SDRD_UDPSIZE 512
#define SDRD_NB_ORIGINAL 128
#define SDRD_BLOCK_BYTES 508
#define SDRD_SAMPLES_PER_BLOCK 127
#define SDRD_FRAME_SAMPLES (127 * 127)
#define SDRD_MAX_FEC 128

const char* sdrd_last_error(void);
const char* sdrd_version(void);
/* Number of usable sm_100 devices (0 if none); never fails. */
int sdrd_device_count(void);
/* Select the device used by handles created afterwards on this thread (cudaSetDevice). */
int sdrd_set_device(int device);

/* ------------------------------------------------------------------------------------------
 * Decimator: Downsampler + Decimators + IntHalfbandFilter{EO1,DB}<64>
 *   replaces  Downsampler::Downsampler/configure/process (include/Downsampler.h:42-57,
 *             sdmnbase/Downsampler.cpp:32-162) and the routines it dispatches to
 *             (sdmnbase/Decimators.cpp:22-1305), called from sdrdaemonrx.cpp:640.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrd_dec sdrd_dec;

/* log2_decim 0..6, fcpos SDRD_FC_*, variant SDRD_HB_*; max_in = largest n_in per stream per call. */
int sdrd_dec_create(sdrd_dec** dec, int log2_decim, int fcpos, int variant, int n_streams, size_t max_in);
void sdrd_dec_destroy(sdrd_dec* dec);
/* Forget all filter state (a freshly constructed Decimators, include/Decimators.h:57-62). */
int sdrd_dec_reset(sdrd_dec* dec);
/* Downsampler::configure (Downsampler.cpp:32-67): change decim / fcpos between blocks.  As in the
 * reference, only m_decim / m_fcPos change: the six half-band stage objects persist
 * (include/Decimators.h:57-62), so a stage the new cascade uses continues from the state it was left in
 * under an earlier configuration (zeros if it never ran) -- the first outputs after a change are
 * bit-identical to the reference's.  Waits for the handle's pending work. */
int sdrd_dec_configure(sdrd_dec* dec, int log2_decim, int fcpos);
int sdrd_dec_log2_decim(const sdrd_dec* dec);

/* Downsampler::process (include/Downsampler.h:57) for n_streams streams at once.
 *   iq_in   n_in samples per stream, stream pitch in_stride samples (HOST)
 *   iq_out  receives n_in >> log2_decim samples per stream, pitch out_stride samples (HOST)
 *   n_out   samples written per stream
 *   sample_bits  the reference's in/out `sampleSize` (effective bits per component)
 * Input samples beyond the last whole group of 2^log2_decim are ignored exactly as the
 * reference's loop bounds do (Decimators.cpp:412). */
int sdrd_dec_process(sdrd_dec* dec, const int16_t* iq_in, size_t n_in, size_t in_stride, int16_t* iq_out,
                     size_t out_stride, size_t* n_out, unsigned* sample_bits);

/* Downsampler::rescale (Downsampler.cpp:69-72) = the static Decimators::decimate1 (Decimators.cpp:22-35): sources
 * with fewer than 16 bits are left-justified in place; no filter state is read or changed, whatever the handle's
 * current decimation.  iq_inout: n samples per stream (HOST), pitch `stride` samples. */
int sdrd_dec_rescale(sdrd_dec* dec, int16_t* iq_inout, size_t n, size_t stride, unsigned* sample_bits);

/* Device-resident form.  The caller writes the next n_in samples of every stream to
 * sdrd_dec_dev_input() (pitch *stride samples) and reads the result from sdrd_dec_dev_output(). */
void* sdrd_dec_dev_input(sdrd_dec* dec, size_t* stride);
void* sdrd_dec_dev_output(sdrd_dec* dec, size_t* stride);
int sdrd_dec_process_dev(sdrd_dec* dec, size_t n_in, size_t* n_out, unsigned* sample_bits, void* cuda_stream);
/* Number of kernel launches issued by this handle so far (for bench accounting). */
long long sdrd_dec_launches(const sdrd_dec* dec);

/* Feeding a handle from ANOTHER process (drop-in mode: one process holds every stream, the GPUs of the box
 * each run a range of them in a process of their own).  The owner exports its input buffer as a CUDA IPC handle
 * (64 bytes); the feeding process opens it and writes the samples with the copy engines over NVLink --
 * sdrd_ipc_copy_rows is a device-to-device cudaMemcpy2DAsync, it occupies no SM and therefore overlaps the kernels
 * of both processes completely (an NCCL send / recv pair is a kernel on either side and has to wait for SMs the
 * decimator's one-wave grid has filled).  *offset_bytes / *stride are where stream 0's next input sample goes inside
 * the exported allocation and the stream pitch in samples (= sdrd_dec_dev_input). */
#define SDRD_IPC_HANDLE_BYTES 64
int sdrd_dec_ipc_export(sdrd_dec* dec, void* handle_out, size_t* offset_bytes, size_t* stride);
int sdrd_ipc_open(const void* handle, void** dev_ptr);
int sdrd_ipc_close(void* dev_ptr);
/* n_rows rows of row_bytes each from src (pitch src_pitch bytes) to dst (pitch dst_pitch bytes), device to device,
 * either pointer may belong to a peer device; asynchronous on cuda_stream */
int sdrd_ipc_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t n_rows,
                       void* cuda_stream);

/* ------------------------------------------------------------------------------------------
 * Interpolator (Tx side): Upsampler + Interpolators + IntHalfbandFilter{EO1,DB}<64/32/16>
 *   replaces  Upsampler::Upsampler/configure/process (include/Upsampler.h:36-50,
 *             sdmnbase/Upsampler.cpp:22-84) and Interpolators::interpolate{2..64}_cen
 *             (sdmnbase/Interpolators.cpp:23-606), called from sdrdaemontx.cpp's main loop.
 *   The EO1 and DB builds of the reference compute the same samples here.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrd_int sdrd_int;

/* log2_interp 0..6; max_in = largest n_in per stream per call. */
int sdrd_int_create(sdrd_int** up, int log2_interp, int n_streams, size_t max_in);
void sdrd_int_destroy(sdrd_int* up);
/* Forget all filter state (a freshly constructed Interpolators, include/Interpolators.h:52-58). */
int sdrd_int_reset(sdrd_int* up);
/* Upsampler::configure (Upsampler.cpp:32-55): change interp between blocks.  The stage objects persist
 * (include/Interpolators.h:52-58) exactly as for sdrd_dec_configure. */
int sdrd_int_configure(sdrd_int* up, int log2_interp);
int sdrd_int_log2_interp(const sdrd_int* up);
/* Upsampler::process (include/Upsampler.h:50) for n_streams streams at once: iq_out receives
 * n_in << log2_interp samples per stream (HOST pointers, pitches in samples). */
int sdrd_int_process(sdrd_int* up, const int16_t* iq_in, size_t n_in, size_t in_stride, int16_t* iq_out,
                     size_t out_stride, size_t* n_out);
/* Device-resident form, as for the decimator. */
void* sdrd_int_dev_input(sdrd_int* up, size_t* stride);
void* sdrd_int_dev_output(sdrd_int* up, size_t* stride);
int sdrd_int_process_dev(sdrd_int* up, size_t n_in, size_t* n_out, void* cuda_stream);
long long sdrd_int_launches(const sdrd_int* up);

/* ------------------------------------------------------------------------------------------
 * CM256 (Cauchy MDS GF(256) erasure code), batched over superframes
 *   replaces  CM256::cm256_encode as called at sdmnbase/UDPSinkFEC.cpp:228-246 and
 *             CM256::cm256_decode as called at sdmnbase/SDRdaemonFECBuffer.cpp:170-213
 *             (OriginalCount = 128, BlockBytes = 508).
 * ------------------------------------------------------------------------------------------ */
/* originals: n_frames x 128 blocks of 508 bytes, consecutive blocks `block_pitch` bytes apart
 * (508 for bare payloads, 512 for datagram images pointed at their payload);
 * recovery: n_frames x recovery_count x 508 bytes, contiguous (the layout cm256_encode writes). */
int sdrd_cm256_encode(const uint8_t* originals, size_t block_pitch, int n_frames, int recovery_count,
                      uint8_t* recovery);

/* The library-level seam under the FEC classes: cm256cc's descriptor API exactly as the reference calls it
 *   CM256::cm256_encode(params, blocks, recoveryBlocks)   sdmnbase/UDPSinkFEC.cpp:228-246
 *   CM256::cm256_decode(params, blocks)                   sdmnbase/SDRdaemonFECBuffer.cpp:148-163,176,197
 * (and the C form cm256_encode / cm256_decode of gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp:40,191).
 * include/cm256.h wraps these two calls in the `CM256` class and the C functions, so that the reference's
 * UDPSinkFEC.cpp / SDRdaemonFECBuffer.cpp compile unmodified against this library.
 *   - HOST pointers; one superframe per call; returns 0 on success, non-zero otherwise, like cm256.
 *   - This library implements sdrdaemon's superframe only: OriginalCount must be 128, BlockBytes 1..508
 *     (UDPSinkFEC.cpp:228-230, SDRdaemonFECBuffer.cpp:32-34); other shapes are refused with SDRD_EINVAL.
 *   - encode: blocks[j].Block is original j (Index is ignored, as in cm256); recovery receives RecoveryCount
 *     contiguous blocks of BlockBytes.
 *   - decode: blocks[0 .. 127] are the received blocks in any order, Index < 128 = original, else recovery row.
 *     On success every recovery descriptor's buffer holds a recovered original and its Index is rewritten to
 *     that original's index (erased indices ascending, paired with the recovery descriptors in array order).
 *     A repeated original or a repeated recovery row is refused.  params.RecoveryCount == 1 takes cm256's XOR
 *     shortcut (the lone recovery block is assumed to be row 128). */
typedef struct sdrd_cm256_block {
    void* Block;
    unsigned char Index;
} sdrd_cm256_block;
typedef struct sdrd_cm256_params {
    int OriginalCount;
    int RecoveryCount;
    int BlockBytes;
} sdrd_cm256_params;
int sdrd_cm256_encode_blocks(sdrd_cm256_params params, const sdrd_cm256_block* originals, void* recovery);
int sdrd_cm256_decode_blocks(sdrd_cm256_params params, sdrd_cm256_block* blocks);

/* ------------------------------------------------------------------------------------------
 * Sender: UDPSinkFEC framing + encode
 *   replaces  UDPSinkFEC::write (sdmnbase/UDPSinkFEC.cpp:79-191) and the encode half of
 *             UDPSinkFEC::transmitUDP (:193-256); the datagram images it returns are what the
 *             reference hands to UDPSocket::SendDataGram (:259-282), in send order.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrd_sink sdrd_sink;

int sdrd_sink_create(sdrd_sink** sink, int n_streams, size_t max_samples);
void sdrd_sink_destroy(sdrd_sink* sink);
/* UDPSinkFEC::reset: frame counter, block and sample indices back to zero. */
int sdrd_sink_reset(sdrd_sink* sink);
/* UDPSink::setCenterFrequency (already divided to kHz) / setSampleRate / setSampleBytes / setSampleBits */
int sdrd_sink_set_meta(sdrd_sink* sink, uint32_t center_freq_khz, uint32_t sample_rate, uint8_t sample_bytes,
                       uint8_t sample_bits);
/* UDPSinkFEC::setNbBlocksFEC */
int sdrd_sink_set_nb_fec(sdrd_sink* sink, int nb_fec);
/* Time stamp written into block 0 of frames started from now on.  The reference reads the wall clock when the
 * first sample of a frame is written (gettimeofday, UDPSinkFEC.cpp:89-95).  `use_fixed` is a set of two flags:
 *   bit 0  the time of a call is (tv_sec, tv_usec) instead of the wall clock read at the call;
 *   bit 1  per-frame stamps: a frame begun `o` samples into a call is stamped  time of the call + o / sample_rate
 *          (whole microseconds) -- what a caller feeding the sink in real time, block by block, sees from the
 *          reference.  Without it every frame begun in one call carries the call's time.
 * 0 = wall clock, one stamp per call (default). */
int sdrd_sink_set_time(sdrd_sink* sink, int use_fixed, uint32_t tv_sec, uint32_t tv_usec);
/* Datagrams per completed frame with the current FEC setting: 128 + nb_fec. */
int sdrd_sink_blocks_per_frame(const sdrd_sink* sink);
/* Frames a write of n_samples per stream would complete given the samples already pending. */
size_t sdrd_sink_frames_for(const sdrd_sink* sink, size_t n_samples);

/* UDPSinkFEC::write for n_streams streams.  iq: n_samples per stream (HOST), pitch `stride` samples.
 * datagrams (HOST) receives, per stream, the completed frames one after the other, each
 * (128 + nb_fec) x 512 bytes in send order; stream pitch = frame_capacity frames.
 * *n_frames = frames completed per stream (identical for all streams). */
int sdrd_sink_write(sdrd_sink* sink, const int16_t* iq, size_t n_samples, size_t stride, uint8_t* datagrams,
                    size_t frame_capacity, size_t* n_frames);

/* Device-resident form: samples is a DEVICE pointer (stream pitch `stride` samples); the datagram
 * images stay on the device at sdrd_sink_dev_datagrams() (stream pitch *frame_pitch frames of
 * (128 + nb_fec) x 512 bytes). */
int sdrd_sink_write_dev(sdrd_sink* sink, const void* samples, size_t n_samples, size_t stride, size_t* n_frames,
                        void* cuda_stream);
void* sdrd_sink_dev_datagrams(sdrd_sink* sink, size_t* frame_pitch);
long long sdrd_sink_launches(const sdrd_sink* sink);

/* ------------------------------------------------------------------------------------------
 * Fused receiver-side pipeline  TestSource -> Downsampler -> UDPSinkFEC  (sdrdaemonrx.cpp:579-663)
 *   = sdrd_dec_process feeding sdrd_sink_write without leaving HBM.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrd_rx sdrd_rx;

int sdrd_rx_create(sdrd_rx** rx, int log2_decim, int fcpos, int variant, int n_streams, size_t max_in);
void sdrd_rx_destroy(sdrd_rx* rx);
int sdrd_rx_reset(sdrd_rx* rx);
sdrd_dec* sdrd_rx_dec(sdrd_rx* rx);   /* borrowed */
sdrd_sink* sdrd_rx_sink(sdrd_rx* rx); /* borrowed: use the sdrd_sink_set_* calls on it */
/* iq_in as sdrd_dec_process, datagrams / frame_capacity / n_frames as sdrd_sink_write.
 * sample_bits: in/out as for sdrd_dec_process (the source's get_sample_bits() in, the decimator's sampleSize
 * out; NULL = 16).  As the reference's main loop does after every block (sdrdaemonrx.cpp:618-643), the sink's
 * sample bits / bytes are set from it before the block is framed: the decimator's output size, or the source's
 * own when decim = 0; bytes = (bits - 1) / 8 + 1.  Centre frequency and sample rate stay what
 * sdrd_sink_set_meta(sdrd_rx_sink(rx), ...) last set.
 * A call that is refused (SDRD_EINVAL / SDRD_ERANGE) leaves the handle untouched. */
int sdrd_rx_process(sdrd_rx* rx, const int16_t* iq_in, size_t n_in, size_t in_stride, uint8_t* datagrams,
                    size_t frame_capacity, size_t* n_frames, unsigned* sample_bits);
/* Calls of at least min_call_bytes input bytes (all streams) go through in 8 slices so that the host -> device
 * copy of slice i + 1 overlaps the kernels and the copy-back of slice i (default 32 MiB; 0 restores it). */
int sdrd_rx_set_slice_bytes(sdrd_rx* rx, size_t min_call_bytes);
/* Queued form, for callers that feed one block after the other the way the reference's threads do
 * (sdrdaemonrx.cpp:579-663: the source thread pushes blocks, the main thread decimates and writes, the sink's own
 * thread encodes and sends).
 *   sdrd_rx_submit   copies the block (n_in samples per stream, a multiple of 2^decim; any host memory) into the
 *                    handle's page-locked accumulation buffer and returns without waiting for the device.  Whenever
 *                    the device is idle, everything accumulated so far goes out as ONE chain of copy -> decimate ->
 *                    frame + encode -> copy back: blocks are batched exactly as far as the device lags behind the
 *                    producer, so a slow producer sees single-block latency and a fast one full batches.  Up to two
 *                    chains are in flight: the copy of one runs under the kernels of the one before (a second
 *                    chain is sent only when it is at least half as long as the one in flight).  The call waits
 *                    only when a block no longer fits behind what has accumulated (max_in samples per stream)
 *                    while two chains are still in flight.  The queued path alternates the decimator between two
 *                    device input buffers: pointers obtained from sdrd_dec_dev_input / sdrd_dec_ipc_export of this
 *                    handle's decimator are not stable across sdrd_rx_submit.
 *                    sample_bits as for sdrd_rx_process; on return it holds the decimator's output sample size.
 *   sdrd_rx_collect  hands over the frames completed so far, oldest first, layout as sdrd_rx_process (stream pitch
 *                    frame_capacity frames); wait != 0 first waits for everything submitted before the call.
 *                    Frames of different sizes (setNbBlocksFEC changed in between) come in separate calls;
 *                    *blocks_per_frame tells which.  May be called from another thread than sdrd_rx_submit.
 * The datagrams are bit-identical to those of sdrd_rx_process over the same sample stream. */
int sdrd_rx_submit(sdrd_rx* rx, const int16_t* iq_in, size_t n_in, size_t in_stride, unsigned* sample_bits);
int sdrd_rx_collect(sdrd_rx* rx, uint8_t* datagrams, size_t frame_capacity, size_t* n_frames, int* blocks_per_frame,
                    int wait);
/* chains sent to the device by sdrd_rx_submit / sdrd_rx_collect so far (blocks submitted / chains = batching achieved) */
long long sdrd_rx_chains(const sdrd_rx* rx);
/* Latency against throughput: by default (0) an idle device gets whatever has accumulated at once.  Starting a chain
 * costs the submitting thread some tens of microseconds of driver calls whatever its size; a producer that is faster
 * than that can ask for chains of at least min_samples samples per stream (sdrd_rx_collect with wait != 0 still sends
 * what is there). */
int sdrd_rx_set_min_chain(sdrd_rx* rx, size_t min_samples);
/* sdrd_rx_submit is bounded by one host core copying the block into page-locked memory (15 - 20 GB/s against the
 * 55 GB/s the copy engine takes).  n_helpers (0..15, default 0) threads owned by the handle share that copy with the
 * caller: they spin for some tens of microseconds after a block (a streaming producer finds them awake), sleep
 * otherwise, and end with sdrd_rx_destroy.  The reference spreads the same work over its four threads
 * (sdrdaemonrx.cpp:579-663).  No effect on the results. */
int sdrd_rx_set_staging_threads(sdrd_rx* rx, int n_helpers);

/* Device-resident form: input at sdrd_dec_dev_input(sdrd_rx_dec(rx)), datagram images left at
 * sdrd_rx_dev_datagrams() (stream pitch *frame_pitch frames of (128 + nb_fec) x 512 bytes). */
void* sdrd_rx_dev_datagrams(sdrd_rx* rx, size_t* frame_pitch);
int sdrd_rx_process_dev(sdrd_rx* rx, size_t n_in, size_t* n_frames, unsigned* sample_bits, void* cuda_stream);
long long sdrd_rx_launches(const sdrd_rx* rx);

/* ------------------------------------------------------------------------------------------
 * Receiver: SDRdaemonFECBuffer decode
 *   replaces  the store/decode/copy-back part of SDRdaemonFECBuffer::writeAndRead
 *             (sdmnbase/SDRdaemonFECBuffer.cpp:143-213) batched over frames: for each frame, the
 *             first min(n_blocks, 128) received datagrams are stored, cm256_decode runs if recovery
 *             blocks are among them, and blocks 1..127 are returned.
 * ------------------------------------------------------------------------------------------ */
#define SDRD_FRAME_INCOMPLETE 0 /* fewer than 128 blocks: missing blocks read as zero (.cpp:109) */
#define SDRD_FRAME_COMPLETE 1   /* 128 originals, nothing to recover */
#define SDRD_FRAME_RECOVERED 2  /* cm256_decode ran and succeeded */
#define SDRD_FRAME_FAILED (-1)  /* cm256_decode refused the block set */

/* superblocks: n_frames x blocks_pitch received 512-byte datagrams in arrival order (HOST);
 * n_blocks[f] <= blocks_pitch datagrams of frame f are valid;
 * payload: n_frames x 127 x 508 bytes; block0: n_frames x 508 bytes (meta block) or NULL;
 * status: n_frames SDRD_FRAME_* codes. */
int sdrd_fec_decode(const uint8_t* superblocks, size_t blocks_pitch, const int* n_blocks, int n_frames,
                    uint8_t* payload, uint8_t* block0, int* status);
/* Device-resident form of the same (all pointers are device pointers). */
int sdrd_fec_decode_dev(const uint8_t* superblocks, size_t blocks_pitch, const int* n_blocks, int n_frames,
                        uint8_t* payload, uint8_t* block0, int* status, void* cuda_stream);
/* Device-resident encode (all pointers are device pointers). */
int sdrd_cm256_encode_dev(const uint8_t* originals, size_t block_pitch, int n_frames, int recovery_count,
                          uint8_t* recovery, void* cuda_stream);

/* ------------------------------------------------------------------------------------------
 * Receiver framing, batched: SDRdaemonFECBuffer::writeAndRead (sdmnbase/SDRdaemonFECBuffer.cpp:112-250)
 * applied to a whole burst of datagrams in arrival order -- what UDPSourceFEC::read does one
 * datagram at a time (sdmnbase/UDPSourceFEC.cpp:52-78).  Every change of header.frameIndex closes the
 * current slot (a frame is the run of datagrams between two changes; only its first 128 count, :143);
 * all slots closed by a call are decoded in ONE kernel launch; the open slot is carried to the next call.
 * As in the reference, the very first datagram closes the initial empty slot, which comes out as a frame
 * of zeros with 0 blocks.
 * ------------------------------------------------------------------------------------------ */
typedef struct sdrd_src sdrd_src;

/* max_datagrams = most datagrams a single feed call may carry */
int sdrd_src_create(sdrd_src** src, size_t max_datagrams);
void sdrd_src_destroy(sdrd_src* src);
int sdrd_src_reset(sdrd_src* src);
/* datagrams: n x 512 bytes (HOST).  Outputs for the frames closed by this call, in order:
 *   payload      n_frames x 127 x 508 bytes (blocks 1..127; missing blocks zero, :109)
 *   block0       n_frames x 508 bytes or NULL
 *   status       SDRD_FRAME_* per frame
 *   nb_blocks    datagrams received for the frame (m_curNbBlocks, uncapped), nb_recovery (m_curNbRecovery)
 * frame_capacity bounds n_frames (SDRD_ERANGE otherwise; at most n + 1 frames can close). */
int sdrd_src_feed(sdrd_src* src, const uint8_t* datagrams, size_t n, uint8_t* payload, uint8_t* block0,
                  size_t frame_capacity, size_t* n_frames, int* status, int* nb_blocks, int* nb_recovery);
/* the reference's getters: the min / max getters reset their value (include/SDRdaemonFECBuffer.h:95-110) */
int sdrd_src_cur_nb_blocks(const sdrd_src* src);
int sdrd_src_cur_nb_recovery(const sdrd_src* src);
int sdrd_src_min_nb_blocks(sdrd_src* src);
int sdrd_src_max_nb_recovery(sdrd_src* src);
long long sdrd_src_launches(const sdrd_src* src);

#ifdef __cplusplus
}
#endif
#endif /* SDRD_B200_H */
