/*
 * sdrd_host.hpp -- C++ host side of the B200 hot path: the reference's plugin / operator interfaces
 * for this path, re-implemented on top of the C ABI (include/sdrd_b200.h).  Same class names,
 * method names, argument meaning and error behaviour as f4exb/sdrdaemon v3.1.2 so that a main()
 * written against the reference (sdrdaemonrx.cpp / sdrdaemontx.cpp) compiles against this header
 * with `using namespace sdrd_b200;`.  Everything numerical happens in libsdrd_b200.so (CUDA);
 * this header only moves vectors, datagrams and configuration strings.
 *
 *   reference                                   here
 *   IQSample / IQSampleVector (SDRDaemon.h:52)  IQSample / IQSampleVector (same 4-byte layout)
 *   DataBuffer<T>             (DataBuffer.h)    DataBuffer<T>
 *   Downsampler               (Downsampler.h)   Downsampler      -> sdrd_dec_*
 *   UDPSink / UDPSinkFEC      (UDPSinkFEC.h)    UDPSink / UDPSinkFEC   -> sdrd_sink_* (+ sendto)
 *   UDPSource / UDPSourceFEC  (UDPSourceFEC.h)  UDPSource / UDPSourceFEC -> recvfrom + sdrd_fec_decode
 *   DeviceSource / TestSource (TestSource.h)    DeviceSource / TestSource (front plug, host float math
 *                                               exactly as TestSource.cpp:395-416)
 *   DeviceSink / FileSink     (FileSink.h)      DeviceSink / FileSink (back plug, .sdriq writer)
 *
 * Out of scope (kept as in the reference, not re-implemented): nanomsg control port, hardware
 * sources, the CLI.  configure() takes the same "key=value,key=value" strings (parsekv.h:40-43).
 */
#pragma once

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdrd_b200.h"

#if defined(SDRD_HOST_REFERENCE_TYPES)
/* Built inside the reference tree (host/compat/ in front of the reference's include/): the glue types are the
 * reference's own dependency-free headers -- IQSample / IQSampleVector / SampleVector (include/SDRDaemon.h:52-70)
 * and DataBuffer<T> (include/DataBuffer.h:29-126) -- so that sdrdaemonrx.cpp / sdrdaemontx.cpp see exactly the types
 * they were written against; only the compute classes below are this library's. */
#include "DataBuffer.h"
#include "SDRDaemon.h"
namespace sdrd_b200 {
using ::DataBuffer;
using ::IQSample;
using ::IQSampleVector;
static_assert(sizeof(IQSample) == 4, "IQSample is 4 bytes on the wire");
#else
namespace sdrd_b200 {

/* ---------------------------------------------------------------- sample types --------------- */

#pragma pack(push, 1)
struct IQSample {
    int16_t m_real, m_imag;
    IQSample() : m_real(0), m_imag(0) {}
    IQSample(int16_t re, int16_t im) : m_real(re), m_imag(im) {}
    int16_t real() const { return m_real; }
    int16_t imag() const { return m_imag; }
    void setReal(int16_t v) { m_real = v; }
    void setImag(int16_t v) { m_imag = v; }
};
#pragma pack(pop)
static_assert(sizeof(IQSample) == 4, "IQSample is 4 bytes on the wire");
typedef std::vector<IQSample> IQSampleVector;

/* thread queue of sample vectors, include/DataBuffer.h:39-126 */
template <class T>
class DataBuffer {
public:
    DataBuffer() : m_qlen(0), m_end(false) {}
    void push(std::vector<T>&& v)
    {
        if (v.empty()) return;
        std::unique_lock<std::mutex> lk(m_mutex);
        m_qlen += v.size();
        m_queue.push(std::move(v));
        lk.unlock();
        m_cond.notify_all();
    }
    void push_end()
    {
        std::unique_lock<std::mutex> lk(m_mutex);
        m_end = true;
        lk.unlock();
        m_cond.notify_all();
    }
    std::size_t queued_samples()
    {
        std::unique_lock<std::mutex> lk(m_mutex);
        return m_qlen;
    }
    /* empty vector = end marker reached */
    std::vector<T> pull()
    {
        std::vector<T> r;
        std::unique_lock<std::mutex> lk(m_mutex);
        while (m_queue.empty() && !m_end) m_cond.wait(lk);
        if (!m_queue.empty()) {
            m_qlen -= m_queue.front().size();
            std::swap(r, m_queue.front());
            m_queue.pop();
        }
        return r;
    }
    bool pull_end_reached()
    {
        std::unique_lock<std::mutex> lk(m_mutex);
        return m_qlen == 0 && m_end;
    }
    void wait_buffer_fill(std::size_t minfill)
    {
        std::unique_lock<std::mutex> lk(m_mutex);
        while (m_qlen < minfill && !m_end) m_cond.wait(lk);
    }

private:
    std::size_t m_qlen;
    bool m_end;
    std::queue<std::vector<T>> m_queue;
    std::mutex m_mutex;
    std::condition_variable m_cond;
};
#endif

/* "key=value,key=value" (separators , or &), the grammar of include/parsekv.h:40-43 */
namespace parsekv {
typedef std::map<std::string, std::string> pairs_type;
inline bool parse(const std::string& s, pairs_type& m)
{
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find_first_of(",&", pos);
        if (end == std::string::npos) end = s.size();
        std::string item = s.substr(pos, end - pos);
        if (!item.empty()) {
            size_t eq = item.find('=');
            std::string k = item.substr(0, eq), v = eq == std::string::npos ? "" : item.substr(eq + 1);
            if (k.empty() || !(isalpha((unsigned char)k[0]) || k[0] == '_')) return false;
            m[k] = v;
        }
        pos = end + 1;
    }
    return true;
}
} /* namespace parsekv */

/* ---------------------------------------------------------------- Downsampler ---------------- */

/* include/Downsampler.h:25-85 / sdmnbase/Downsampler.cpp.  One stream per object, like the reference;
 * a batch of streams is available through the C ABI directly. */
class Downsampler {
public:
    typedef enum { FC_POS_INFRA = 0, FC_POS_SUPRA, FC_POS_CENTER } fcPos_t;

    /* variant: which IntHalfbandFilter the peer reference build uses (SDRD_HB_EO1 on x86) */
    Downsampler(unsigned int decim = 0, fcPos_t fcPos = FC_POS_CENTER, int variant = SDRD_HB_EO1,
                std::size_t max_block = 1 << 20)
        : m_decim(decim), m_fcPos(fcPos), m_dec(nullptr)
    {
        if (sdrd_dec_create(&m_dec, (int)decim, (int)fcPos, variant, 1, max_block) != 0) m_error = sdrd_last_error();
    }
    ~Downsampler() { sdrd_dec_destroy(m_dec); }
    Downsampler(const Downsampler&) = delete;
    Downsampler& operator=(const Downsampler&) = delete;

    /* Downsampler.cpp:32-67: keys "decim" (0..6) and "fcpos" (0..2) */
    bool configure(parsekv::pairs_type& m)
    {
        unsigned decim = m_decim;
        int fcpos = (int)m_fcPos;
        if (m.find("decim") != m.end()) {
            int log2Decim = atoi(m["decim"].c_str());
            if (log2Decim < 0 || log2Decim > 6) {
                m_error = "Invalid log2 decimation factor";
                return false;
            }
            decim = (unsigned)log2Decim;
        }
        if (m.find("fcpos") != m.end()) {
            fcpos = atoi(m["fcpos"].c_str());
            if (fcpos < 0 || fcpos > 2) {
                m_error = "Invalid Fc position index";
                return false;
            }
        }
        if (!m_dec || sdrd_dec_configure(m_dec, (int)decim, fcpos) != 0) {
            m_error = sdrd_last_error();
            return false;
        }
        m_decim = decim;
        m_fcPos = (fcPos_t)fcpos;
        return true;
    }
    unsigned int getLog2Decimation() const { return m_decim; }

    void process(unsigned int& sampleSize, const IQSampleVector& samples_in, IQSampleVector& samples_out)
    {
        if (!m_dec) return;
        samples_out.resize(samples_in.size() ? samples_in.size() : 1);
        std::size_t n_out = 0;
        unsigned ss = sampleSize;
        if (sdrd_dec_process(m_dec, reinterpret_cast<const int16_t*>(samples_in.data()), samples_in.size(), samples_in.size(),
                             reinterpret_cast<int16_t*>(samples_out.data()), samples_out.size(), &n_out, &ss) != 0) {
            m_error = sdrd_last_error();
            samples_out.clear();
            return;
        }
        samples_out.resize(n_out);
        sampleSize = ss;
    }
    /* decimation 1 path (Downsampler.cpp:69-72): Decimators::decimate1 in place, no filter state involved */
    void rescale(unsigned int& sampleSize, IQSampleVector& samples_inout)
    {
        if (!m_dec) return;
        unsigned ss = sampleSize;
        if (sdrd_dec_rescale(m_dec, reinterpret_cast<int16_t*>(samples_inout.data()), samples_inout.size(), samples_inout.size(), &ss) != 0)
            m_error = sdrd_last_error();
        sampleSize = ss;
    }
    operator bool() const { return m_error.empty(); }
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }

private:
    unsigned int m_decim;
    fcPos_t m_fcPos;
    sdrd_dec* m_dec;
    std::string m_error;
};

/* ------------------------------------------------------------------------------------------
 * Upsampler (reference include/Upsampler.h:28-76, sdmnbase/Upsampler.cpp): the Tx side's
 * interpolation by 2^interp, behind sdrd_int_*.
 * ------------------------------------------------------------------------------------------ */
class Upsampler {
public:
    Upsampler(unsigned int interp = 0, std::size_t max_block = 1 << 16) : m_interp(interp), m_int(nullptr)
    {
        if (sdrd_int_create(&m_int, (int)interp, 1, max_block) != 0) m_error = sdrd_last_error();
    }
    ~Upsampler() { sdrd_int_destroy(m_int); }
    Upsampler(const Upsampler&) = delete;
    Upsampler& operator=(const Upsampler&) = delete;

    /* Upsampler.cpp:32-55: key "interp" (0..6) */
    bool configure(parsekv::pairs_type& m)
    {
        if (m.find("interp") != m.end()) {
            int log2Interp = atoi(m["interp"].c_str());
            if (log2Interp < 0 || log2Interp > 6) {
                m_error = "Invalid log2 interpolation factor";
                return false;
            }
            if (!m_int || sdrd_int_configure(m_int, log2Interp) != 0) {
                m_error = sdrd_last_error();
                return false;
            }
            m_interp = (unsigned)log2Interp;
        }
        return true;
    }
    unsigned int getLog2Interpolation() const { return m_interp; }

    void process(const IQSampleVector& samples_in, IQSampleVector& samples_out)
    {
        if (!m_int) return;
        samples_out.resize(samples_in.size() ? (samples_in.size() << m_interp) : 1);
        std::size_t n_out = 0;
        if (sdrd_int_process(m_int, reinterpret_cast<const int16_t*>(samples_in.data()), samples_in.size(), samples_in.size(),
                             reinterpret_cast<int16_t*>(samples_out.data()), samples_out.size(), &n_out) != 0) {
            m_error = sdrd_last_error();
            samples_out.clear();
            return;
        }
        samples_out.resize(n_out);
    }
    operator bool() const { return m_error.empty(); }
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }

private:
    unsigned int m_interp;
    sdrd_int* m_int;
    std::string m_error;
};

/* ---------------------------------------------------------------- UDP helpers ---------------- */

class UdpTx {
public:
    UdpTx() : m_fd(-1) { memset(&m_to, 0, sizeof(m_to)); }
    ~UdpTx() { if (m_fd >= 0) close(m_fd); }
    bool open(const std::string& address, unsigned port, std::string& err)
    {
        m_fd = socket(AF_INET, SOCK_DGRAM, 0);
        if (m_fd < 0) { err = "socket() failed"; return false; }
        m_to.sin_family = AF_INET;
        m_to.sin_port = htons((uint16_t)port);
        if (inet_pton(AF_INET, address.c_str(), &m_to.sin_addr) != 1) { err = "bad address " + address; return false; }
        return true;
    }
    void send(const void* p, size_t n) { if (m_fd >= 0) sendto(m_fd, p, n, 0, (const sockaddr*)&m_to, sizeof(m_to)); }
private:
    int m_fd;
    sockaddr_in m_to;
};

class UdpRx {
public:
    UdpRx() : m_fd(-1) {}
    ~UdpRx() { if (m_fd >= 0) close(m_fd); }
    bool open(const std::string& address, unsigned port, std::string& err)
    {
        m_fd = socket(AF_INET, SOCK_DGRAM, 0);
        if (m_fd < 0) { err = "socket() failed"; return false; }
        int rcvbuf = 64 << 20;
        setsockopt(m_fd, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
        timeval tv = {0, 200000};
        setsockopt(m_fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        sockaddr_in a;
        memset(&a, 0, sizeof(a));
        a.sin_family = AF_INET;
        a.sin_port = htons((uint16_t)port);
        if (inet_pton(AF_INET, address.c_str(), &a.sin_addr) != 1) { err = "bad address " + address; return false; }
        if (bind(m_fd, (sockaddr*)&a, sizeof(a)) < 0) { err = "bind() failed"; return false; }
        return true;
    }
    int recv(void* p, size_t n) { return m_fd < 0 ? -1 : (int)::recv(m_fd, p, n, 0); }
private:
    int m_fd;
};

/* ---------------------------------------------------------------- UDPSink -------------------- */

/* include/UDPSink.h:30-129 (the members the hot path uses) */
class UDPSink {
public:
    UDPSink(const std::string& address, unsigned int port, unsigned int udpSize)
        : m_address(address), m_port(port), m_udpSize(udpSize), m_centerFrequency(0), m_sampleRate(48000),
          m_sampleBytes(1), m_sampleBits(8), m_nbSamples(0) {}
    virtual ~UDPSink() {}
    virtual void write(const IQSampleVector& samples_in) = 0;
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }
    void setCenterFrequency(uint64_t centerFrequency) { m_centerFrequency = (uint32_t)(centerFrequency / 1000); }
    void setSampleRate(uint32_t sampleRate) { m_sampleRate = sampleRate; }
    void setSampleBytes(uint8_t sampleBytes) { m_sampleBytes = (sampleBytes & 0x0F) + (m_sampleBytes & 0xF0); }
    void setSampleBits(uint8_t sampleBits) { m_sampleBits = sampleBits; }
    virtual void setNbBlocksFEC(int) {}
    virtual void setTxDelay(int) {}
    operator bool() const { return m_error.empty(); }

protected:
    std::string m_address;
    unsigned int m_port;
    unsigned int m_udpSize;
    std::string m_error;
    uint32_t m_centerFrequency; /* kHz */
    uint32_t m_sampleRate;
    uint8_t m_sampleBytes, m_sampleBits;
    uint32_t m_nbSamples;
};

/* include/UDPSinkFEC.h + sdmnbase/UDPSinkFEC.cpp.  write() frames and encodes on the GPU; a Tx thread
 * sends the datagram images with usleep(txDelay) between them like transmitUDP (:259-282).  A frame
 * is handed to the Tx thread as soon as it completes (the reference's one-frame lag, SURVEY H4b(ii),
 * is a defect of its slot hand-over and is not reproduced).  An optional tap sees every datagram. */
class UDPSinkFEC : public UDPSink {
public:
    typedef void (*datagram_tap)(void* user, const uint8_t* datagram, int frame_blocks, int block);

    UDPSinkFEC(const std::string& address, unsigned int port, std::size_t max_block = 1 << 20)
        : UDPSink(address, port, SDRD_UDPSIZE), m_sink(nullptr), m_nbBlocksFEC(0) /* UDPSinkFEC.cpp:31 */, m_txDelay(0), m_running(true),
          m_tap(nullptr), m_tapUser(nullptr), m_puncture(-1)
    {
        if (sdrd_sink_create(&m_sink, 1, max_block + SDRD_FRAME_SAMPLES) != 0) m_error = sdrd_last_error(); /* a write plus what was collected before it */
        /* one stamp per frame (UDPSinkFEC.cpp:89-95): the wall clock of the call that brings the frame's samples to the
         * library, advanced by the sample clock within the call */
        if (m_sink) sdrd_sink_set_time(m_sink, 2, 0, 0);
        std::string err;
        if (!m_tx.open(address, port, err)) m_error = err;
        m_txThread = std::thread(&UDPSinkFEC::transmitUDP, this);
    }
    virtual ~UDPSinkFEC()
    {
        {
            std::unique_lock<std::mutex> lk(m_mutex);
            m_running = false;
        }
        m_cond.notify_all();
        if (m_txThread.joinable()) m_txThread.join();
        sdrd_sink_destroy(m_sink);
    }
    virtual void setNbBlocksFEC(int nbBlocksFEC) { m_nbBlocksFEC = nbBlocksFEC; }
    virtual void setTxDelay(int txDelay) { m_txDelay = txDelay; }
    void setTimestamp(uint32_t tv_sec, uint32_t tv_usec) { if (m_sink) sdrd_sink_set_time(m_sink, 1, tv_sec, tv_usec); } /* tests: a fixed stamp */
    void setTap(datagram_tap tap, void* user) { m_tap = tap; m_tapUser = user; }
    void setPuncture(int block) { m_puncture = block; } /* SDRDAEMON_PUNCTURE, UDPSinkFEC.cpp:261-265 */
    void setTxEnabled(bool on) { m_txEnabled = on; }    /* false: datagrams are built (and tapped) but not handed to sendto */
    void reset()
    {
        m_pending.clear();
        if (m_sink) sdrd_sink_reset(m_sink);
    }

    /* UDPSinkFEC::write.  Samples are collected on the host while they complete no frame (the reference's write()
     * likewise only memcpy's them into the slot being filled, UDPSinkFEC.cpp:138-155) and go to the GPU -- framing +
     * encode, sdrd_sink_write -- as soon as a frame completes; what was collected is pushed through first whenever a
     * setter has changed a value since, so the library sees settings and samples in exactly the order of the calls. */
    virtual void write(const IQSampleVector& samples_in)
    {
        if (!m_sink) return;
        const Settings now = {m_centerFrequency, m_sampleRate, m_sampleBytes, m_sampleBits, (int)m_nbBlocksFEC};
        if (!(now == m_applied)) {
            if (!push()) return; /* with the settings those samples were written under */
            sdrd_sink_set_meta(m_sink, now.freq, now.rate, now.bytes, now.bits);
            if (sdrd_sink_set_nb_fec(m_sink, now.nbFEC) != 0) { m_error = sdrd_last_error(); return; }
            m_applied = now;
        }
        m_pending.insert(m_pending.end(), samples_in.begin(), samples_in.end());
        if (sdrd_sink_frames_for(m_sink, m_pending.size()) > 0) push();
    }
    /* block until everything queued has been sent */
    void flush()
    {
        std::unique_lock<std::mutex> lk(m_mutex);
        while (!m_queue.empty() || m_sending) m_cond.wait(lk);
    }

private:
    struct Batch {
        int blocks_per_frame, tx_delay;
        std::vector<uint8_t> data;
    };
    struct Settings {
        uint32_t freq, rate;
        uint8_t bytes, bits;
        int nbFEC;
        bool operator==(const Settings& o) const { return freq == o.freq && rate == o.rate && bytes == o.bytes && bits == o.bits && nbFEC == o.nbFEC; }
    };
    /* hand the collected samples to the library; completed frames go to the Tx thread */
    bool push()
    {
        if (m_pending.empty()) return true;
        const int bpf = sdrd_sink_blocks_per_frame(m_sink);
        const std::size_t cap = sdrd_sink_frames_for(m_sink, m_pending.size());
        Batch b;
        b.blocks_per_frame = bpf;
        b.tx_delay = m_txDelay;
        b.data.resize((cap ? cap : 1) * (std::size_t)bpf * SDRD_UDPSIZE);
        std::size_t n_frames = 0;
        const int rc = sdrd_sink_write(m_sink, reinterpret_cast<const int16_t*>(m_pending.data()), m_pending.size(), m_pending.size(),
                                       b.data.data(), cap ? cap : 1, &n_frames);
        m_pending.clear();
        if (rc != 0) {
            m_error = sdrd_last_error();
            return false;
        }
        if (!n_frames) return true;
        b.data.resize(n_frames * (std::size_t)bpf * SDRD_UDPSIZE);
        {
            std::unique_lock<std::mutex> lk(m_mutex);
            m_queue.push_back(std::move(b));
        }
        m_cond.notify_all();
        return true;
    }
    void transmitUDP()
    {
        for (;;) {
            Batch b;
            {
                std::unique_lock<std::mutex> lk(m_mutex);
                while (m_queue.empty() && m_running) m_cond.wait(lk);
                if (m_queue.empty()) return;
                b = std::move(m_queue.front());
                m_queue.pop_front();
                m_sending = true;
            }
            const std::size_t n = b.data.size() / SDRD_UDPSIZE;
            for (std::size_t i = 0; i < n; i++) {
                const int blk = (int)(i % (std::size_t)b.blocks_per_frame);
                if (blk == m_puncture) continue;
                const uint8_t* dg = b.data.data() + i * SDRD_UDPSIZE;
                if (m_tap) m_tap(m_tapUser, dg, b.blocks_per_frame, blk);
                if (m_txEnabled) m_tx.send(dg, SDRD_UDPSIZE);
                if (b.tx_delay > 0) usleep(b.tx_delay);
            }
            {
                std::unique_lock<std::mutex> lk(m_mutex);
                m_sending = false;
            }
            m_cond.notify_all();
        }
    }
    sdrd_sink* m_sink;
    IQSampleVector m_pending;           /* samples written since the last frame completed */
    Settings m_applied = {0, 0, 0, 0, -1}; /* what the library was last told */
    std::atomic_int m_nbBlocksFEC, m_txDelay;
    bool m_running;
    bool m_sending = false;
    bool m_txEnabled = true;
    datagram_tap m_tap;
    void* m_tapUser;
    int m_puncture;
    UdpTx m_tx;
    std::deque<Batch> m_queue;
    std::mutex m_mutex;
    std::condition_variable m_cond;
    std::thread m_txThread;
};

/* ---------------------------------------------------------------- UDPSource ------------------ */

class UDPSource {
public:
    UDPSource(const std::string& address, unsigned int port, unsigned int udpSize)
        : m_address(address), m_port((unsigned short)port), m_udpSize(udpSize), m_sampleBytes(2), m_sampleBits(16) {}
    virtual ~UDPSource() {}
    virtual void read(IQSampleVector& samples_out) = 0;
    virtual void getStatusMessage(char* messageBuffer) = 0;
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }
    uint8_t getSampleBytes() const { return m_sampleBytes; }
    uint8_t getSampleBits() { return m_sampleBits; }
    operator bool() const { return m_error.empty(); }

protected:
    std::string m_address;
    unsigned short m_port;
    unsigned int m_udpSize;
    std::string m_error;
    uint8_t m_sampleBytes, m_sampleBits;
};

/* include/UDPSourceFEC.h + include/SDRdaemonFECBuffer.h.  Datagrams of the frame being received are
 * collected on the host in arrival order (the first 128, SDRdaemonFECBuffer.cpp:143); when a datagram
 * of another frame index arrives the collected frame is decoded on the GPU (sdrd_fec_decode) and
 * returned -- the same hand-over point as SDRdaemonFECBuffer::writeAndRead (:133-139).  Counters as
 * initDecodeSlot (:95-110). */
class UDPSourceFEC : public UDPSource {
public:
#pragma pack(push, 1)
    struct MetaDataFEC {
        uint32_t m_centerFrequency, m_sampleRate;
        uint8_t m_sampleBytes, m_sampleBits, m_nbOriginalBlocks, m_nbFECBlocks;
        uint32_t m_tv_sec, m_tv_usec, m_crc32;
    };
#pragma pack(pop)

    UDPSourceFEC(const std::string& address, unsigned int port)
        : UDPSource(address, port, SDRD_UDPSIZE), m_frameHead(-1), m_blockCount(0), m_recoveryCount(0), m_metaRetrieved(false),
          m_curNbBlocks(0), m_curNbRecovery(0), m_minNbBlocks(256), m_maxNbRecovery(0), m_lastStatus(0), m_stop(nullptr)
    {
        memset(&m_currentMeta, 0, sizeof(m_currentMeta));
        m_currentMeta.m_nbFECBlocks = 0xFF;
        m_frame.resize(128 * SDRD_UDPSIZE);
        std::string err;
        if (!m_rx.open(address, port, err)) m_error = err;
    }
    void setStopFlag(std::atomic_bool* stop) { m_stop = stop; }

    /* feed one received datagram; returns true when a frame (127*127 samples) was written to out */
    bool writeAndRead(const uint8_t* superBlock, IQSampleVector& out)
    {
        bool available = false;
        const int frameIndex = superBlock[0] | (superBlock[1] << 8);
        if (m_frameHead != frameIndex) {
            decodeSlot(out);
            available = true;
            m_curNbBlocks = m_blockCount;
            m_curNbRecovery = m_recoveryCount;
            if (m_curNbBlocks < m_minNbBlocks) m_minNbBlocks = m_curNbBlocks;
            if (m_curNbRecovery > m_maxNbRecovery) m_maxNbRecovery = m_curNbRecovery;
            m_blockCount = 0;
            m_recoveryCount = 0;
            m_metaRetrieved = false;
            m_frameHead = frameIndex;
        }
        if (m_blockCount < 128) {
            memcpy(&m_frame[(std::size_t)m_blockCount * SDRD_UDPSIZE], superBlock, SDRD_UDPSIZE);
            if (superBlock[2] == 0) m_metaRetrieved = true;
            if (superBlock[2] >= 128) m_recoveryCount++;
        }
        m_blockCount++;
        return available;
    }

    virtual void read(IQSampleVector& samples_out)
    {
        uint8_t sb[2048];
        for (;;) {
            if (m_stop && m_stop->load()) { samples_out.clear(); return; }
            int n = m_rx.recv(sb, sizeof(sb));
            if (n != SDRD_UDPSIZE) continue; /* UDPSourceFEC.cpp:64: other sizes are ignored */
            if (writeAndRead(sb, samples_out)) return;
        }
    }
    /* UDPSourceFEC.cpp:80-95 */
    virtual void getStatusMessage(char* messageBuffer)
    {
        int msgLen = (int)strlen(messageBuffer);
        int statusCode;
        int minNbBlocks = getMinNbBlocks();
        if (minNbBlocks < 128) statusCode = 1;
        else if (minNbBlocks < 128 + m_currentMeta.m_nbFECBlocks) statusCode = 0;
        else statusCode = 2;
        sprintf(&messageBuffer[msgLen], ":%d:%03d/%03d", statusCode, minNbBlocks, getMaxNbRecovery());
    }
    int getCurNbBlocks() const { return m_curNbBlocks; }
    int getCurNbRecovery() const { return m_curNbRecovery; }
    int getMinNbBlocks() { int v = m_minNbBlocks; m_minNbBlocks = 256; return v; }
    int getMaxNbRecovery() { int v = m_maxNbRecovery; m_maxNbRecovery = 0; return v; }
    const MetaDataFEC& getCurrentMeta() const { return m_currentMeta; }
    int getLastFrameStatus() const { return m_lastStatus; } /* SDRD_FRAME_* of the frame just returned */

private:
    void decodeSlot(IQSampleVector& out)
    {
        out.resize(SDRD_FRAME_SAMPLES);
        uint8_t block0[SDRD_BLOCK_BYTES];
        int nb = m_blockCount < 128 ? m_blockCount : 128;
        int status = 0;
        if (sdrd_fec_decode(m_frame.data(), 128, &nb, 1, reinterpret_cast<uint8_t*>(out.data()), block0, &status) != 0) {
            m_error = sdrd_last_error();
            memset((void*)out.data(), 0, out.size() * sizeof(IQSample));
            return;
        }
        m_lastStatus = status;
        /* meta data only of a complete frame whose block 0 ARRIVED (SDRdaemonFECBuffer.cpp:237-246; accepting a
         * recovered block 0 is commented out in the reference, :215-218) */
        if (nb == 128 && m_metaRetrieved) {
            if (memcmp(block0, &m_currentMeta, 12) != 0) memcpy(&m_currentMeta, block0, sizeof(m_currentMeta));
            m_sampleBytes = m_currentMeta.m_sampleBytes & 0x0F;
            m_sampleBits = m_currentMeta.m_sampleBits;
        }
    }
    int m_frameHead, m_blockCount, m_recoveryCount;
    bool m_metaRetrieved;
    int m_curNbBlocks, m_curNbRecovery, m_minNbBlocks, m_maxNbRecovery, m_lastStatus;
    MetaDataFEC m_currentMeta;
    std::vector<uint8_t> m_frame;
    UdpRx m_rx;
    std::atomic_bool* m_stop;
};

/* ---------------------------------------------------------------- device plugs --------------- */

/* include/DeviceSource.h:30-156 (without the nanomsg control socket) */
class DeviceSource {
public:
    DeviceSource() : m_confFreq(0), m_decim(0), m_nbFECBlocks(1), m_txDelay(0), m_fcPos(2), m_buf(0), m_stop_flag(0), m_downsampler(0) {}
    virtual ~DeviceSource() {}
    void associateDownsampler(Downsampler* downsampler) { m_downsampler = downsampler; }
    /* include/DeviceSource.h:65-78 binds the nanomsg control socket there; the control plane is out of scope
     * (DESIGN 7): the call is accepted and does nothing, dynamic reconfiguration goes through configure() */
    void setConfigurationPort(std::uint32_t) {}
    /* include/DeviceSource.h:96-99 */
    std::uint64_t get_received_frequency() const { return m_confFreq; }
    /* include/DeviceSource.h:112 */
    virtual void print_specific_parms() {}
    /* sdmnbase/DeviceSource.cpp:25-72 */
    bool configure(std::string& configureStr)
    {
        parsekv::pairs_type m;
        if (!parsekv::parse(configureStr, m)) { m_error = "Configuration parsing failed"; return false; }
        if (m_downsampler && !m_downsampler->configure(m)) { m_error = m_downsampler->error(); return false; }
        if (m.find("decim") != m.end()) m_decim = (unsigned)atoi(m["decim"].c_str());
        if (m.find("fecblk") != m.end()) {
            int nbFECBlocks = atoi(m["fecblk"].c_str());
            if (nbFECBlocks >= 1 || nbFECBlocks < 128) m_nbFECBlocks = (unsigned)nbFECBlocks; /* sic, DeviceSource.cpp:55 */
        }
        if (m.find("txdelay") != m.end()) {
            int txDelay = atoi(m["txdelay"].c_str());
            m_txDelay = txDelay < 0 ? 0u : (unsigned)txDelay; /* DeviceSource.cpp:61-66 */
        }
        return configure(m);
    }
    virtual std::uint32_t get_sample_bits() = 0;
    virtual std::uint32_t get_sample_rate() = 0;
    virtual std::uint32_t get_frequency() = 0;
    unsigned int get_nb_fec_blocks() const { return m_nbFECBlocks; }
    unsigned int get_tx_delay() const { return m_txDelay; }
    virtual bool start(DataBuffer<IQSample>* buf, std::atomic_bool* stop_flag) = 0;
    virtual bool stop() = 0;
    virtual operator bool() const = 0;
    std::string get_device_name() const { return m_devname; }
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }

protected:
    std::string m_devname, m_error;
    uint64_t m_confFreq;
    unsigned int m_decim, m_nbFECBlocks, m_txDelay;
    int m_fcPos;
    DataBuffer<IQSample>* m_buf;
    std::atomic_bool* m_stop_flag;
    Downsampler* m_downsampler;
    virtual bool configure(parsekv::pairs_type& m) = 0;
};

/* include/TestSource.h + sdmnbase/TestSource.cpp: synthetic carrier.  Keys: srate (8000..10000000),
 * freq, dfp / dfn (carrier offset above / below centre, Hz), power (dB attenuation), blklen;
 * `pace=0` (not in the reference) drops the real-time usleep for tests and benchmarks. */
class TestSource : public DeviceSource {
public:
    static const int default_block_length = 65536;
    TestSource(int dev_index = 0)
        : m_dev(dev_index), m_block_length(default_block_length), m_srate(5000000), m_freq(435000000), m_carrierOffset(100000.0),
          m_deltaPhase(0), m_amplitude(1.0f), m_phase(0), m_pace(true), m_thread(0)
    {
        m_devname = "Test source";
    }
    virtual ~TestSource() { stop(); }
    using DeviceSource::configure;
    /* sdmnbase/TestSource.cpp:271-275, 371-393 */
    virtual void print_specific_parms()
    {
        fprintf(stderr, "Delta phase:       %g radians\n", (double)m_deltaPhase);
        fprintf(stderr, "Amplitude:         %g\n", (double)m_amplitude);
    }
    static void get_device_names(std::vector<std::string>& devices)
    {
        devices.clear();
        devices.push_back("Test Test dummy device 0000001 0");
    }
    virtual std::uint32_t get_sample_bits() { return 16; }
    virtual std::uint32_t get_sample_rate() { return (uint32_t)m_srate; }
    virtual std::uint32_t get_frequency() { return (uint32_t)m_freq; }
    virtual operator bool() const { return m_error.empty(); }

    /* TestSource.cpp:395-416, the same float/double arithmetic */
    static int read_samples(int16_t* data, int iqBlockSize, int& getSize, float& phasor, int sampleRate, float deltaPhase,
                            float amplitude, bool pace = true)
    {
        const int m_sampleHalfWidth = 1 << 15;
        int nbSamples = iqBlockSize / 4;
        float dt = (float)nbSamples / (float)sampleRate;
        int dtMicroseconds = (int)(dt * 1e6);
        for (int i = 0; i < nbSamples * 2; i += 2) {
            data[i] = amplitude * cos(phasor) * m_sampleHalfWidth;
            data[i + 1] = amplitude * sin(phasor) * m_sampleHalfWidth;
            phasor += deltaPhase;
            if (phasor > 2.0 * M_PI) {
                phasor -= 2.0 * M_PI;
            } else if (phasor < 2.0 * M_PI) {
                phasor += 2.0 * M_PI;
            }
        }
        if (pace) usleep(dtMicroseconds);
        getSize = iqBlockSize;
        return 0;
    }
    bool get_samples(IQSampleVector* samples)
    {
        std::vector<int16_t> buf(2 * (size_t)m_block_length);
        int n_read = 0;
        if (read_samples(buf.data(), 4 * m_block_length, n_read, m_phase, m_srate, m_deltaPhase, m_amplitude, m_pace) < 0) {
            m_error = "TestSource::get_samples: read_samples failed";
            return false;
        }
        samples->resize(m_block_length);
        for (int i = 0; i < m_block_length; i++) (*samples)[i] = IQSample(buf[2 * i], buf[2 * i + 1]);
        return true;
    }
    virtual bool start(DataBuffer<IQSample>* buf, std::atomic_bool* stop_flag)
    {
        m_buf = buf;
        m_stop_flag = stop_flag;
        if (m_thread == 0) m_thread = new std::thread(run, this);
        return true;
    }
    virtual bool stop()
    {
        if (m_thread) {
            m_thread->join();
            delete m_thread;
            m_thread = 0;
        }
        return true;
    }

protected:
    virtual bool configure(parsekv::pairs_type& m)
    {
        if (m.find("srate") != m.end()) {
            int v = atoi(m["srate"].c_str());
            if (v < 8000 || v > 10000000) { m_error = "Invalid sample rate"; return false; }
            m_srate = v;
        }
        if (m.find("freq") != m.end()) m_freq = strtoull(m["freq"].c_str(), 0, 10);
        if (m.find("blklen") != m.end()) {
            int v = atoi(m["blklen"].c_str());
            if (v > 0) m_block_length = v;
        }
        if (m.find("dfp") != m.end()) m_carrierOffset = atof(m["dfp"].c_str());
        if (m.find("dfn") != m.end()) m_carrierOffset = -atof(m["dfn"].c_str());
        if (m.find("power") != m.end()) m_amplitude = (float)pow(10.0, -atof(m["power"].c_str()) / 20.0);
        if (m.find("pace") != m.end()) m_pace = atoi(m["pace"].c_str()) != 0;
        m_deltaPhase = (float)(2.0 * M_PI * m_carrierOffset / (double)m_srate);
        m_confFreq = m_freq;
        return true;
    }

private:
    static void run(TestSource* self)
    {
        IQSampleVector iqsamples;
        while (!self->m_stop_flag->load() && self->get_samples(&iqsamples)) self->m_buf->push(std::move(iqsamples));
        self->m_buf->push_end();
    }
    int m_dev, m_block_length, m_srate;
    uint64_t m_freq;
    double m_carrierOffset;
    float m_deltaPhase, m_amplitude, m_phase;
    bool m_pace;
    std::thread* m_thread;
};

/* include/DeviceSink.h (the members the path uses) */
class DeviceSink {
public:
    DeviceSink() : m_confFreq(0), m_buf(0), m_stop_flag(0), m_upsampler(0), m_udpSource(0) {}
    virtual ~DeviceSink() {}
    /* include/DeviceSink.h:55-83: the upsampler is configured through the sink's configuration string ("interp"),
     * the UDP source only feeds the status message of the (out of scope) control port */
    void associateUpsampler(Upsampler* upsampler) { m_upsampler = upsampler; }
    void associateUDPSource(UDPSource* udpSource) { m_udpSource = udpSource; }
    void setConfigurationPort(std::uint32_t) {} /* nanomsg control socket: out of scope, see DeviceSource */
    virtual bool configure(std::string& configureStr) = 0;
    virtual std::uint32_t get_device_sample_bits() { return 16; }
    virtual std::uint32_t get_sample_rate() = 0;
    virtual std::uint64_t get_frequency() = 0;
    std::uint64_t get_transmit_frequency() const { return m_confFreq; }
    virtual void print_specific_parms() {}
    std::string get_device_name() const { return m_devname; }
    virtual bool start(DataBuffer<IQSample>* buf, std::atomic_bool* stop_flag) = 0;
    virtual bool stop() = 0;
    virtual operator bool() const = 0;
    std::string error()
    {
        std::string ret(m_error);
        m_error.clear();
        return ret;
    }

protected:
    std::string m_devname, m_error;
    uint64_t m_confFreq;
    DataBuffer<IQSample>* m_buf;
    std::atomic_bool* m_stop_flag;
    Upsampler* m_upsampler;
    UDPSource* m_udpSource;
};

/* include/FileSink.h + sdmnbase/FileSink.cpp: .sdriq = {u32 rate, u64 freq, time_t stamp} then raw
 * int16 I/Q (FileSink.cpp:179-193,242-246).  Keys: file, srate, freq. */
class FileSink : public DeviceSink {
public:
    FileSink(int dev_index = 0) : m_srate(48000), m_freq(435000000), m_thread(0), m_fixedStamp(-1)
    {
        (void)dev_index;
        m_devname = "FileSink";
    }
    virtual ~FileSink() { stop(); }
    virtual bool configure(std::string& configureStr)
    {
        parsekv::pairs_type m;
        if (!parsekv::parse(configureStr, m)) { m_error = "Configuration parsing failed"; return false; }
        if (m.find("file") != m.end()) m_fileName = m["file"];
        if (m.find("srate") != m.end()) m_srate = (uint32_t)atoi(m["srate"].c_str());
        if (m.find("freq") != m.end()) m_freq = strtoull(m["freq"].c_str(), 0, 10);
        if (m.find("stamp") != m.end()) m_fixedStamp = atoll(m["stamp"].c_str());
        /* DeviceSink::configure (sdmnbase/DeviceSink.cpp:26-66): the associated upsampler takes "interp" */
        if (m_upsampler && !m_upsampler->configure(m)) { m_error = m_upsampler->error(); return false; }
        m_confFreq = m_freq;
        if (m_fileName.empty()) { m_error = "No file name"; return false; }
        return openFile();
    }
    /* sdmnbase/FileSink.cpp:58-61,73-76 */
    static void get_device_names(std::vector<std::string>& devices) { devices.push_back("file"); }
    virtual void print_specific_parms() { fprintf(stderr, "File name:         %s\n", m_fileName.c_str()); }
    virtual std::uint32_t get_sample_rate() { return m_srate; }
    virtual std::uint64_t get_frequency() { return m_freq; }
    virtual operator bool() const { return m_error.empty(); }
    virtual bool start(DataBuffer<IQSample>* buf, std::atomic_bool* stop_flag)
    {
        m_buf = buf;
        m_stop_flag = stop_flag;
        if (m_thread == 0) m_thread = new std::thread(run, this);
        return true;
    }
    virtual bool stop()
    {
        if (m_thread) {
            m_thread->join();
            delete m_thread;
            m_thread = 0;
        }
        if (m_ofstream.is_open()) m_ofstream.close();
        return true;
    }

private:
    bool openFile()
    {
        if (m_ofstream.is_open()) m_ofstream.close();
        m_ofstream.open(m_fileName.c_str(), std::ios::binary);
        if (!m_ofstream) { m_error = "Cannot open " + m_fileName; return false; }
        uint32_t rate = m_srate;
        uint64_t freq = m_freq;
        time_t stamp = m_fixedStamp >= 0 ? (time_t)m_fixedStamp : time(0);
        m_ofstream.write((const char*)&rate, sizeof(rate));
        m_ofstream.write((const char*)&freq, sizeof(freq));
        m_ofstream.write((const char*)&stamp, sizeof(stamp));
        return true;
    }
    /* FileSink::run (sdmnbase/FileSink.cpp:216-277) polls the queue instead of blocking in pull(), so that a stop
     * request is seen while the queue is empty; what is still queued at that moment is written out (the reference
     * drops it) */
    static void run(FileSink* self)
    {
        auto write_one = [self]() {
            IQSampleVector v = self->m_buf->pull();
            if (!v.empty()) self->m_ofstream.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(IQSample)));
        };
        while (!self->m_stop_flag->load() && !self->m_buf->pull_end_reached()) {
            if (self->m_buf->queued_samples() > 0) write_one();
            else usleep(500);
        }
        while (self->m_buf->queued_samples() > 0) write_one();
        self->m_ofstream.flush();
    }
    std::string m_fileName;
    uint32_t m_srate;
    uint64_t m_freq;
    std::ofstream m_ofstream;
    std::thread* m_thread;
    long long m_fixedStamp;
};

} /* namespace sdrd_b200 */
