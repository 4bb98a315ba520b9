/*
 * ref_harness.cpp -- C entry points around the UNMODIFIED reference sources, compiled where they
 * lie under /root/reference by oracle/Makefile into oracle/_ref/ (git-ignored, travels to the GPU
 * box).  TEST INFRASTRUCTURE ONLY: used to pin oracle/sdrd_oracle.c, to generate tests/golden/ and
 * as the "reference" CPU baseline in bench.py.  No reference source is copied into this repo.
 *
 * What runs here is the reference's own code: Downsampler::process -> Decimators ->
 * IntHalfbandFilterEO1<64> (with -DUSE_SSE4_1) or IntHalfbandFilterDB<64> (without);
 * SDRdaemonFECBuffer::writeAndRead; UDPSinkFEC::write + transmitUDP over a loop-back UDP socket;
 * TestSource::read_samples.  Only cm256cc is replaced, by oracle/ref_stubs/cm256.h.
 */
#include <arpa/inet.h>
#include <netinet/in.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include <atomic>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>

#include "Downsampler.h"
#include "Upsampler.h"
#include "SDRdaemonFECBuffer.h"
#include "UDPSinkFEC.h"
#include "DeviceSource.h"
/* TestSource::read_samples is a private static; open it for this translation unit only
 * (access control, not behaviour -- TestSource.cpp itself is compiled untouched). */
#define private public
#include "TestSource.h"
#undef private

extern "C" {

int ref_variant(void)
{
#if defined(USE_SSE4_1)
    return 0; /* IntHalfbandFilterEO1<64>, Decimators.h:56-62 */
#else
    return 1; /* IntHalfbandFilterDB<64>,  Decimators.h:63-69 */
#endif
}

/* ------------------------------------------------------------------ Downsampler ---- */

void* ref_ds_create(int log2_decim, int fcpos)
{
    return new Downsampler((unsigned)log2_decim, (Downsampler::fcPos_t)fcpos);
}
void ref_ds_destroy(void* h) { delete (Downsampler*)h; }

/* Downsampler::process, include/Downsampler.h:57.  Returns samples_out.size(). */
size_t ref_ds_process(void* h, unsigned* sample_bits, const int16_t* iq_in, size_t n_in, int16_t* iq_out)
{
    Downsampler* ds = (Downsampler*)h;
    IQSampleVector in(n_in), out;
    memcpy((void*)in.data(), iq_in, n_in * sizeof(IQSample));
    ds->process(*sample_bits, in, out);
    memcpy(iq_out, (const void*)out.data(), out.size() * sizeof(IQSample));
    return out.size();
}

/* Downsampler::configure (sdmnbase/Downsampler.cpp:32-67) with the key/value pairs the control port would
 * deliver ("decim=..,fcpos=.."); returns 1 when the reference accepted them. */
int ref_ds_configure(void* h, int log2_decim, int fcpos)
{
    parsekv::pairs_type m;
    m["decim"] = std::to_string(log2_decim);
    m["fcpos"] = std::to_string(fcpos);
    std::ostringstream quiet; /* the reference logs every key to std::cerr */
    std::streambuf* old = std::cerr.rdbuf(quiet.rdbuf());
    const bool ok = ((Downsampler*)h)->configure(m);
    std::cerr.rdbuf(old);
    return ok ? 1 : 0;
}

/* ------------------------------------------------------------------ Upsampler ---- */

void* ref_us_create(int log2_interp) { return new Upsampler((unsigned)log2_interp); }
void ref_us_destroy(void* h) { delete (Upsampler*)h; }

/* Upsampler::configure (sdmnbase/Upsampler.cpp:32-55) */
int ref_us_configure(void* h, int log2_interp)
{
    parsekv::pairs_type m;
    m["interp"] = std::to_string(log2_interp);
    std::ostringstream quiet;
    std::streambuf* old = std::cerr.rdbuf(quiet.rdbuf());
    const bool ok = ((Upsampler*)h)->configure(m);
    std::cerr.rdbuf(old);
    return ok ? 1 : 0;
}

/* Upsampler::process, include/Upsampler.h:50.  Returns samples_out.size(). */
size_t ref_us_process(void* h, const int16_t* iq_in, size_t n_in, int16_t* iq_out)
{
    Upsampler* us = (Upsampler*)h;
    IQSampleVector in(n_in), out;
    memcpy((void*)in.data(), iq_in, n_in * sizeof(IQSample));
    us->process(in, out);
    memcpy(iq_out, (const void*)out.data(), out.size() * sizeof(IQSample));
    return out.size();
}

/* Multi-threaded throughput leg for bench.py: n_streams independent Downsamplers, one
 * std::thread each (up to n_threads at a time), block-wise like sdrdaemonrx's main loop
 * (sdrdaemonrx.cpp:590-654, TestSource block = 65536 samples, TestSource.h:33). */
void ref_ds_process_streams(int log2_decim, int fcpos, int n_streams, int n_threads, const int16_t* iq_in,
                            size_t n_in_per_stream, int16_t* iq_out, size_t block)
{
    std::vector<std::thread> pool;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= n_streams) return;
            Downsampler ds((unsigned)log2_decim, (Downsampler::fcPos_t)fcpos);
            const int16_t* in = iq_in + (size_t)s * n_in_per_stream * 2;
            int16_t* out = iq_out + (size_t)s * (n_in_per_stream >> log2_decim) * 2;
            IQSampleVector vin, vout;
            size_t done = 0, produced = 0;
            while (done < n_in_per_stream) {
                size_t n = n_in_per_stream - done < block ? n_in_per_stream - done : block;
                vin.resize(n);
                memcpy((void*)vin.data(), in + 2 * done, n * sizeof(IQSample));
                unsigned ss = 16;
                ds.process(ss, vin, vout);
                memcpy(out + 2 * produced, (const void*)vout.data(), vout.size() * sizeof(IQSample));
                produced += vout.size();
                done += n;
            }
        }
    };
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

/* Whole Rx hot path on the CPU for bench.py's reference arm: per stream, the reference's
 * Downsampler::process block by block (sdrdaemonrx.cpp:590-654) feeding the packer + encoder
 * (oracle sink = UDPSinkFEC::write + cm256_encode restated, since cm256cc itself is not available;
 * driving the reference's own UDPSinkFEC would add its sockets and usleep pacing).  One std::thread per
 * stream, n_threads at a time.  Returns the number of superframes produced; *digest = XOR of all
 * datagram words (keeps the work observable). */
extern "C" {
#include "sdrd_oracle.h"
}
struct RxAcc { long long frames; uint32_t digest; };
static void rx_frame_cb(void* user, const uint8_t* dg, int n_blocks, uint16_t)
{
    RxAcc* a = (RxAcc*)user;
    a->frames++;
    const uint32_t* w = (const uint32_t*)dg;
    uint32_t d = 0;
    for (int i = 0; i < n_blocks * 128; i++) d ^= w[i];
    a->digest ^= d;
}
extern "C" long long ref_rx_streams(int log2_decim, int fcpos, int nb_fec, int n_streams, int n_threads,
                                    const int16_t* iq_in, size_t n_in_per_stream, size_t in_stride, size_t block,
                                    uint32_t* digest)
{
    std::vector<std::thread> pool;
    std::atomic<int> next(0);
    std::atomic<long long> frames(0);
    std::atomic<uint32_t> dig(0);
    auto work = [&]() {
        for (;;) {
            int s = next.fetch_add(1);
            if (s >= n_streams) return;
            Downsampler ds((unsigned)log2_decim, (Downsampler::fcPos_t)fcpos);
            RxAcc acc = {0, 0};
            sdro_sink* sink = sdro_sink_create(rx_frame_cb, &acc);
            sdro_sink_set_meta(sink, 435000, 625000, 2, 16);
            sdro_sink_set_nb_fec(sink, nb_fec);
            sdro_sink_set_time(sink, 1700000000u, 0);
            const int16_t* in = iq_in + (size_t)s * in_stride * 2;
            IQSampleVector vin, vout;
            size_t done = 0;
            while (done < n_in_per_stream) {
                size_t n = n_in_per_stream - done < block ? n_in_per_stream - done : block;
                vin.resize(n);
                memcpy((void*)vin.data(), in + 2 * done, n * sizeof(IQSample));
                unsigned ss = 16;
                ds.process(ss, vin, vout);
                sdro_sink_write(sink, (const int16_t*)vout.data(), vout.size());
                done += n;
            }
            sdro_sink_destroy(sink);
            frames.fetch_add(acc.frames);
            dig.fetch_xor(acc.digest);
        }
    };
    for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    if (digest) *digest = dig.load();
    return frames.load();
}

/* ------------------------------------------------------------------ FEC buffer ----- */

void* ref_fecbuf_create(void) { return new SDRdaemonFECBuffer(); }
void ref_fecbuf_destroy(void* h) { delete (SDRdaemonFECBuffer*)h; }
int ref_fecbuf_write_and_read(void* h, const uint8_t* superblock, uint8_t* data, size_t* data_len)
{
    uint8_t sb[SDRDAEMONFEC_UDPSIZE];
    memcpy(sb, superblock, sizeof(sb));
    std::size_t len = 0;
    bool r = ((SDRdaemonFECBuffer*)h)->writeAndRead(sb, data, len);
    *data_len = len;
    return r ? 1 : 0;
}
int ref_fecbuf_cur_nb_blocks(void* h) { return ((SDRdaemonFECBuffer*)h)->getCurNbBlocks(); }
int ref_fecbuf_cur_nb_recovery(void* h) { return ((SDRdaemonFECBuffer*)h)->getCurNbRecovery(); }
int ref_fecbuf_min_nb_blocks(void* h) { return ((SDRdaemonFECBuffer*)h)->getMinNbBlocks(); }
int ref_fecbuf_max_nb_recovery(void* h) { return ((SDRdaemonFECBuffer*)h)->getMaxNbRecovery(); }
void ref_fecbuf_current_meta(void* h, uint8_t meta20[20])
{
    memcpy(meta20, &((SDRdaemonFECBuffer*)h)->getCurrentMeta(), 20);
}

/* BASELINE config 4 on the CPU, all cores: every frame's received datagrams go through the reference's own
 * SDRdaemonFECBuffer::writeAndRead (store, cm256_decode, copy-back; sdmnbase/SDRdaemonFECBuffer.cpp:112-250), one
 * buffer object per thread; a datagram of another frame index flushes the frame out (:133-139).
 * payload: n_frames x 127 x 508 bytes.  The reference logs every decode to std::cerr: silenced for the call. */
struct DiscardBuf : std::streambuf { /* stateless: safe to share between the threads that log into it */
    int overflow(int c) override { return c; }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
void ref_fecbuf_decode_frames(const uint8_t* sb, size_t pitch, const int* n_blocks, int n_frames, int n_threads, uint8_t* payload)
{
    static DiscardBuf quiet;
    std::streambuf* old = std::cerr.rdbuf(&quiet);
    std::atomic<int> next(0);
    auto work = [&]() {
        SDRdaemonFECBuffer* buf = new SDRdaemonFECBuffer();
        std::vector<uint8_t> data(128 * 512);
        uint8_t dg[SDRDAEMONFEC_UDPSIZE];
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_frames) break;
            std::size_t len = 0;
            for (int i = 0; i < n_blocks[f]; i++) {
                memcpy(dg, sb + ((size_t)f * pitch + i) * 512, 512);
                dg[0] = (uint8_t)(f & 0xFF); /* distinct frame indices per frame of the batch */
                dg[1] = (uint8_t)((f >> 8) & 0x7F);
                buf->writeAndRead(dg, data.data(), len);
            }
            memset(dg, 0, 512);
            dg[0] = (uint8_t)((f + 1) & 0xFF);
            dg[1] = (uint8_t)((((f + 1) >> 8) & 0x7F) | 0x80); /* never equal to a frame's own index */
            if (buf->writeAndRead(dg, data.data(), len)) memcpy(payload + (size_t)f * 127 * 508, data.data(), 127 * 508);
        }
        delete buf;
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
    std::cerr.rdbuf(old);
}

/* ------------------------------------------------------------------ UDPSinkFEC ----- */

/* Drives the reference sender over 127.0.0.1:port and captures the datagrams it emits.
 * Writes `iq` (n_samples) in chunks of `chunk` samples through UDPSinkFEC::write.  The Tx thread
 * lags one superframe behind the writer (UDPSinkFEC.cpp:160,208), so callers pass one frame more
 * than they want back.  Returns the number of 512-byte datagrams stored in `out`. */
int ref_sink_run(int port, uint32_t freq_khz, uint32_t rate, int nb_fec, const int16_t* iq, size_t n_samples,
                 size_t chunk, uint8_t* out, int max_datagrams, int expect_datagrams)
{
    int fd = socket(AF_INET, SOCK_DGRAM, 0);
    if (fd < 0) return -1;
    int rcvbuf = 64 << 20;
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
    sockaddr_in a;
    memset(&a, 0, sizeof(a));
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    a.sin_port = htons((uint16_t)port);
    if (bind(fd, (sockaddr*)&a, sizeof(a)) < 0) { close(fd); return -2; }

    /* m_running is never initialised before the Tx thread starts (UDPSinkFEC.cpp:29-44): give the
     * object storage whose bytes read as `true` so the thread's first m_running.load() passes. */
    void* mem = operator new(sizeof(UDPSinkFEC));
    memset(mem, 0x01, sizeof(UDPSinkFEC));
    UDPSinkFEC* sink = new (mem) UDPSinkFEC("127.0.0.1", (unsigned)port);
    sink->setCenterFrequency(freq_khz * 1000ULL);
    sink->setSampleRate(rate);
    sink->setSampleBytes(2);
    sink->setSampleBits(16);
    sink->setNbBlocksFEC(nb_fec);
    sink->setTxDelay(0);

    int got = 0;
    auto drain = [&](int timeout_ms) {
        pollfd p = {fd, POLLIN, 0};
        while (got < max_datagrams && poll(&p, 1, timeout_ms) > 0) {
            uint8_t buf[2048];
            ssize_t n = recv(fd, buf, sizeof(buf), 0);
            if (n == 512) { memcpy(out + (size_t)got * 512, buf, 512); got++; }
            if (got >= expect_datagrams) break;
        }
    };
    size_t done = 0;
    IQSampleVector v;
    while (done < n_samples) {
        size_t n = n_samples - done < chunk ? n_samples - done : chunk;
        v.resize(n);
        memcpy((void*)v.data(), iq + 2 * done, n * sizeof(IQSample));
        sink->write(v);
        done += n;
        drain(0);
    }
    for (int tries = 0; tries < 200 && got < expect_datagrams; tries++) drain(10);
    sink->~UDPSinkFEC();
    operator delete(mem);
    close(fd);
    return got;
}

/* ------------------------------------------------------------------ TestSource ----- */

/* TestSource::read_samples (TestSource.cpp:395-422), including its real-time usleep. */
int ref_testsource_read(int16_t* data, int n_samples, float* phasor, int sample_rate, float delta_phase,
                        float amplitude)
{
    int got = 0;
    int r = TestSource::read_samples(data, 4 * n_samples, got, *phasor, sample_rate, delta_phase, amplitude);
    return r < 0 ? r : got / 4;
}

} /* extern "C" */
