/*
 * ref_seam_loopback.cpp -- the reference's OWN sender and receiver classes, compiled unmodified from
 * /root/reference, run end to end over loop-back UDP with cm256 supplied through a header named cm256.h:
 *
 *     UDPSinkFEC::write -> transmitUDP -> CM256::cm256_encode   (sdmnbase/UDPSinkFEC.cpp:79-288)
 *         --- 127.0.0.1 ---> recvfrom
 *     SDRdaemonFECBuffer::writeAndRead -> CM256::cm256_decode   (sdmnbase/SDRdaemonFECBuffer.cpp:112-250)
 *
 * oracle/Makefile builds it three times: against oracle/ref_stubs/cm256.h (the CPU restatement: the baseline),
 * against include/cm256.h + tests/emu/libsdrd_emu.so, and against include/cm256.h + libsdrd_b200.so -- the
 * last is the drop-in claim: reference code, GPU arithmetic, nothing else changed.  The sender is compiled with
 * the reference's own erasure injection -DSDRDAEMON_PUNCTURE=101 (UDPSinkFEC.cpp:27,261-265) and this program
 * drops `extra` more original blocks per frame, so that every frame needs 1 + extra recovery blocks.
 * Every recovered frame is compared with the samples that were sent.  TEST INFRASTRUCTURE ONLY.
 *
 *   ref_seam_loopback <port> <n_fec> <n_frames> <extra_drop> <seed> [capture-file]
 * prints one JSON line; exit status 0 when every checked frame matched.
 */
#include <arpa/inet.h>
#include <netinet/in.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <array>
#include <atomic>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <mutex>
#include <new>
#include <sstream>
#include <string>
#include <thread>

#include "SDRdaemonFECBuffer.h"
#include "UDPSinkFEC.h"
#include "ref_seam_common.h"

#ifndef SDRDAEMON_PUNCTURE
#define SDRDAEMON_PUNCTURE (-1)
#endif

int main(int argc, char** argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s port n_fec n_frames extra_drop seed [capture-file]\n", argv[0]);
        return 2;
    }
    const int port = atoi(argv[1]), n_fec = atoi(argv[2]), n_frames = atoi(argv[3]), extra = atoi(argv[4]);
    const uint64_t seed = strtoull(argv[5], 0, 10);
    FILE* cap = argc > 6 ? fopen(argv[6], "wb") : nullptr;

    int fd = socket(AF_INET, SOCK_DGRAM, 0);
    int rcvbuf = 64 << 20;
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &rcvbuf, sizeof(rcvbuf));
    sockaddr_in a;
    memset(&a, 0, sizeof(a));
    a.sin_family = AF_INET;
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    a.sin_port = htons((uint16_t)port);
    if (fd < 0 || bind(fd, (sockaddr*)&a, sizeof(a)) < 0) {
        fprintf(stderr, "cannot bind 127.0.0.1:%d\n", port);
        return 2;
    }
    /* the reference logs every decode to std::cerr: keep it, but out of the way */
    std::ostringstream ref_log;
    std::streambuf* old_cerr = std::cerr.rdbuf(ref_log.rdbuf());

    /* m_running is never initialised before the Tx thread starts (UDPSinkFEC.cpp:29-44): storage whose bytes
     * read as `true` lets the thread's first m_running.load() pass (as oracle/ref_harness.cpp does) */
    void* mem = operator new(sizeof(UDPSinkFEC));
    memset(mem, 0x01, sizeof(UDPSinkFEC));
    UDPSinkFEC* sink = new (mem) UDPSinkFEC("127.0.0.1", (unsigned)port);
    sink->setCenterFrequency(435000000ULL);
    sink->setSampleRate(625000);
    sink->setSampleBytes(2);
    sink->setSampleBits(16);
    sink->setNbBlocksFEC(n_fec);
    /* the sender paces itself a little (the reference's own knob, usleep per datagram) ... */
    sink->setTxDelay(5);
    SDRdaemonFECBuffer* fecbuf = new SDRdaemonFECBuffer();

    /* ... and the socket is emptied by a thread of its own into a queue: a datagram lost in the kernel's receive buffer
     * while the main thread sits in a decode call (the first GPU call of the process takes a while) would be an
     * erasure the test did not ask for */
    std::mutex q_mutex;
    std::deque<std::array<uint8_t, 512>> queue;
    std::atomic<bool> rx_stop(false);
    std::thread rx_thread([&] {
        pollfd p = {fd, POLLIN, 0};
        while (!rx_stop.load()) {
            if (poll(&p, 1, 5) <= 0) continue;
            uint8_t buf[2048];
            if (recv(fd, buf, sizeof(buf), 0) == 512) {
                std::array<uint8_t, 512> d;
                memcpy(d.data(), buf, 512);
                std::lock_guard<std::mutex> lk(q_mutex);
                queue.push_back(d);
            }
        }
    });

    std::vector<int16_t> frame(2 * seam::FRAME_SAMPLES), want(2 * seam::FRAME_SAMPLES);
    std::vector<uint8_t> data(128 * 512);
    int head = -1, checked = 0, ok = 0, datagrams = 0, fed = 0;
    uint32_t digest = 2166136261u;
    bool first = true;
    auto feed = [&](uint8_t* dg) {
        const int fi = dg[0] | (dg[1] << 8), bi = dg[2];
        datagrams++;
        if (cap) fwrite(dg, 1, 512, cap);
        if (seam::dropped(seed, fi, bi, extra, SDRDAEMON_PUNCTURE)) return;
        fed++;
        std::size_t len = 0;
        if (fecbuf->writeAndRead(dg, data.data(), len)) {
            if (!first && head >= 0 && head < n_frames) { /* the very first emission is the empty initial slot */
                seam::gen_frame(seed, head, want.data());
                checked++;
                if (len == want.size() * 2 && memcmp(data.data(), want.data(), len) == 0) ok++;
                digest = seam::fold(digest, data.data(), len);
            }
            first = false;
        }
        head = fi;
    };
    auto drain = [&](int timeout_ms) {
        for (int waited = 0;;) {
            std::array<uint8_t, 512> d;
            bool have = false;
            {
                std::lock_guard<std::mutex> lk(q_mutex);
                if (!queue.empty()) {
                    d = queue.front();
                    queue.pop_front();
                    have = true;
                }
            }
            if (have) {
                feed(d.data());
                waited = 0;
            } else if (waited >= timeout_ms) {
                return;
            } else {
                usleep(1000);
                waited++;
            }
        }
    };
    /* the Tx thread lags one superframe behind the writer (UDPSinkFEC.cpp:160,208) and the receiver emits a frame
     * when the next one starts: two frames more than are checked */
    IQSampleVector v;
    for (int f = 0; f < n_frames + 2; f++) {
        seam::gen_frame(seed, f, frame.data());
        for (int done = 0; done < seam::FRAME_SAMPLES;) {
            const int n = seam::FRAME_SAMPLES - done < 4096 ? seam::FRAME_SAMPLES - done : 4096;
            v.resize(n);
            memcpy((void*)v.data(), &frame[2 * done], (size_t)n * 4);
            sink->write(v);
            done += n;
            drain(0);
        }
    }
    /* The sender's first cm256_encode may be the first GPU call of the process: context creation and module load take
     * from a fraction of a second to several seconds on a freshly started box.  Wait for the first datagram for as long
     * as a minute; once they flow, give up only after five seconds of silence. */
    for (int idle = 0, waited = 0; checked < n_frames && waited < 9000; waited++) {
        const int before = datagrams;
        drain(10);
        idle = datagrams != before ? 0 : idle + 1;
        if (datagrams == 0 ? waited >= 6000 : idle >= 500) break;
    }
    rx_stop.store(true);
    rx_thread.join();
    sink->~UDPSinkFEC();
    operator delete(mem);
    delete fecbuf;
    close(fd);
    if (cap) fclose(cap);
    std::cerr.rdbuf(old_cerr);

    const std::string log = ref_log.str();
    size_t n_success = 0, n_error = 0;
    for (size_t p = 0; (p = log.find("CM256 decode success", p)) != std::string::npos; p++) n_success++;
    for (size_t p = 0; (p = log.find("CM256 decode error", p)) != std::string::npos; p++) n_error++;
    printf("{\"frames_checked\": %d, \"frames_ok\": %d, \"datagrams\": %d, \"fed\": %d, \"decode_success\": %zu, "
           "\"decode_error\": %zu, \"digest\": %u, \"n_fec\": %d, \"extra_drop\": %d, \"puncture\": %d}\n",
           checked, ok, datagrams, fed, n_success, n_error, digest, n_fec, extra, (int)SDRDAEMON_PUNCTURE);
    return (checked >= n_frames && ok == checked && n_error == 0) ? 0 : 1;
}
