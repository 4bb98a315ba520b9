/*
 * host_pipeline.cpp -- drives the C++ host layer (sdrdaemon_b200/host/sdrd_host.hpp) the way the
 * reference's two mains do (sdrdaemonrx.cpp:579-663 and sdrdaemontx.cpp:448-500), in one process:
 *
 *   TestSource -> DataBuffer -> Downsampler::process -> UDPSinkFEC::write ==UDP 127.0.0.1==>
 *   UDPSourceFEC::read -> DataBuffer -> FileSink (.sdriq)
 *
 *   host_pipeline testsource <n_samples> <srate> <dfp> <power_db> <out.raw>
 *   host_pipeline upsample <interp> <block> <in.raw> <out.raw>       Upsampler::process block by block, the
 *       Tx side's step after UDPSourceFEC::read (sdrdaemontx.cpp); interp is re-configured through
 *       Upsampler::configure("interp=..") after the first block like a control message would
 *   host_pipeline pipeline <port> <config> <n_blocks> <out.sdriq> <datagrams.bin> [puncture]
 *       config e.g. "srate=2400000,decim=4,fecblk=16,dfp=100000,power=6,blklen=65536"
 *
 * Linked against libsdrd_b200.so on a GPU box, or against tests/emu/libsdrd_emu.so for the CPU-only
 * check of the host logic.
 */
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../sdrdaemon_b200/host/sdrd_host.hpp"

using namespace sdrd_b200;

static std::mutex g_tap_mutex;
static void tap(void* user, const uint8_t* dg, int, int) { fwrite(dg, 1, SDRD_UDPSIZE, (FILE*)user); }

static int run_testsource(int argc, char** argv)
{
    if (argc < 7) return 2;
    int n = atoi(argv[2]), srate = atoi(argv[3]);
    double dfp = atof(argv[4]), power = atof(argv[5]);
    std::vector<int16_t> buf(2 * (size_t)n);
    float phase = 0;
    int got = 0;
    float amp = (float)pow(10.0, -power / 20.0);
    float dphi = (float)(2.0 * M_PI * dfp / (double)srate);
    TestSource::read_samples(buf.data(), 4 * n, got, phase, srate, dphi, amp, false);
    FILE* f = fopen(argv[6], "wb");
    fwrite(buf.data(), 2, buf.size(), f);
    fclose(f);
    return 0;
}

static int run_upsample(int argc, char** argv)
{
    if (argc < 6) return 2;
    const int interp = atoi(argv[2]);
    const size_t block = (size_t)atoi(argv[3]);
    FILE* fi = fopen(argv[4], "rb");
    FILE* fo = fopen(argv[5], "wb");
    if (!fi || !fo) return 1;
    Upsampler up(0, block);
    if (!up) { fprintf(stderr, "Upsampler: %s\n", up.error().c_str()); return 1; }
    parsekv::pairs_type bad;
    bad["interp"] = "9";
    if (up.configure(bad) || up.error() != "Invalid log2 interpolation factor") return 3; /* Upsampler.cpp:38-42 */
    parsekv::pairs_type kv;
    kv["interp"] = std::to_string(interp);
    if (!up.configure(kv) || (int)up.getLog2Interpolation() != interp) return 4;
    IQSampleVector in(block), out;
    for (;;) {
        size_t n = fread((void*)in.data(), sizeof(IQSample), block, fi);
        if (n == 0) break;
        in.resize(n);
        up.process(in, out);
        if (!up) { fprintf(stderr, "process: %s\n", up.error().c_str()); return 1; }
        fwrite((const void*)out.data(), sizeof(IQSample), out.size(), fo);
        in.resize(block);
    }
    fclose(fi);
    fclose(fo);
    return 0;
}

static int run_pipeline(int argc, char** argv)
{
    if (argc < 7) return 2;
    unsigned port = (unsigned)atoi(argv[2]);
    std::string config = argv[3];
    int n_blocks = atoi(argv[4]);
    std::string out_file = argv[5];
    FILE* dgf = fopen(argv[6], "wb");
    int puncture = argc > 7 ? atoi(argv[7]) : -1;

    std::atomic_bool stop_rx(false), stop_src(false), stop_file(false);

    /* ---- Tx side of the link (sdrdaemontx): UDPSourceFEC -> FileSink ---- */
    UDPSourceFEC udp_input("127.0.0.1", port);
    if (!udp_input) { fprintf(stderr, "UDPSourceFEC: %s\n", udp_input.error().c_str()); return 1; }
    udp_input.setStopFlag(&stop_rx);
    FileSink file_sink;
    DataBuffer<IQSample> sink_buffer;
    std::string fcfg = "file=" + out_file + ",srate=150000,freq=435000000,stamp=1700000000";
    if (!file_sink.configure(fcfg)) { fprintf(stderr, "FileSink: %s\n", file_sink.error().c_str()); return 1; }
    file_sink.start(&sink_buffer, &stop_file);
    int frames_in = 0, frames_recovered = 0;
    std::thread rx_thread([&]() {
        IQSampleVector frame;
        bool first = true;
        for (;;) {
            udp_input.read(frame);
            if (frame.empty()) break;
            if (first) { first = false; continue; } /* the empty slot emitted before the first frame */
            frames_in++;
            if (udp_input.getLastFrameStatus() == SDRD_FRAME_RECOVERED) frames_recovered++;
            sink_buffer.push(std::move(frame));
        }
        sink_buffer.push_end();
    });

    /* ---- Rx side (sdrdaemonrx): TestSource -> Downsampler -> UDPSinkFEC ---- */
    Downsampler dn;
    if (!dn) { fprintf(stderr, "Downsampler: %s\n", dn.error().c_str()); return 1; }
    UDPSinkFEC udp_output("127.0.0.1", port);
    if (!udp_output) { fprintf(stderr, "UDPSinkFEC: %s\n", udp_output.error().c_str()); return 1; }
    udp_output.setTap(tap, dgf);
    udp_output.setTimestamp(1700000000u, 0);
    udp_output.setPuncture(puncture);
    TestSource src;
    src.associateDownsampler(&dn);
    std::string cfg = config + ",pace=0";
    DeviceSource* srcsdr = &src;
    if (!srcsdr->configure(cfg)) { fprintf(stderr, "configure: %s\n", srcsdr->error().c_str()); return 1; }
    DataBuffer<IQSample> source_buffer;
    srcsdr->start(&source_buffer, &stop_src);

    long long samples_out = 0;
    for (int blk = 0; blk < n_blocks; blk++) {
        IQSampleVector iqsamples = source_buffer.pull();
        if (iqsamples.empty()) break;
        unsigned int sampleSize = srcsdr->get_sample_bits();
        udp_output.setCenterFrequency(srcsdr->get_frequency());
        udp_output.setSampleRate(srcsdr->get_sample_rate() >> dn.getLog2Decimation());
        udp_output.setNbBlocksFEC((int)srcsdr->get_nb_fec_blocks());
        udp_output.setTxDelay((int)srcsdr->get_tx_delay());
        IQSampleVector outsamples;
        dn.process(sampleSize, iqsamples, outsamples);
        if (!dn) { fprintf(stderr, "process: %s\n", dn.error().c_str()); return 1; }
        udp_output.setSampleBytes(2);
        udp_output.setSampleBits((uint8_t)sampleSize);
        if (blk == 0) continue; /* sdrdaemonrx.cpp:646-648: the first block is thrown away */
        samples_out += (long long)outsamples.size();
        udp_output.write(outsamples);
        if (!udp_output) { fprintf(stderr, "write: %s\n", udp_output.error().c_str()); return 1; }
    }
    stop_src.store(true);
    while (!source_buffer.pull().empty()) {}
    srcsdr->stop();
    /* the receiver hands a frame over when the next one starts: pad with zeros up to the start of
     * one more frame */
    {
        long long rem = samples_out % SDRD_FRAME_SAMPLES;
        IQSampleVector pad((size_t)(SDRD_FRAME_SAMPLES - rem) + SDRD_FRAME_SAMPLES);
        udp_output.write(pad);
    }
    udp_output.flush();
    usleep(300000);
    stop_rx.store(true);
    rx_thread.join();
    stop_file.store(true);
    file_sink.stop();
    fclose(dgf);
    char status[64] = "status";
    udp_input.getStatusMessage(status);
    printf("samples_out=%lld frames_received=%d frames_recovered=%d %s\n", samples_out, frames_in, frames_recovered, status);
    return 0;
}

/* host_pipeline blocks <n_blocks> <decim> <fecblk> <blklen> [warmup]
 * The reference's call granularity, timed: one stream, TestSource-sized blocks in pageable std::vectors through
 * Downsampler::process + UDPSinkFEC::write (sdrdaemonrx.cpp:636-660), datagrams built but not sent.  Prints one JSON
 * line; bench.py's `small_block` leg runs this. */
static int run_blocks(int argc, char** argv)
{
    if (argc < 6) return 2;
    const int n_blocks = atoi(argv[2]), decim = atoi(argv[3]), fecblk = atoi(argv[4]), blklen = atoi(argv[5]);
    const int warmup = argc > 6 ? atoi(argv[6]) : 8;
    Downsampler dn((unsigned)decim, Downsampler::FC_POS_CENTER, SDRD_HB_EO1, (size_t)blklen);
    if (!dn) { fprintf(stderr, "Downsampler: %s\n", dn.error().c_str()); return 1; }
    UDPSinkFEC udp_output("127.0.0.1", 19399, (size_t)blklen);
    if (!udp_output) { fprintf(stderr, "UDPSinkFEC: %s\n", udp_output.error().c_str()); return 1; }
    udp_output.setTxEnabled(false);
    udp_output.setNbBlocksFEC(fecblk);
    udp_output.setSampleRate(10000000u >> decim);
    udp_output.setCenterFrequency(435000000ull);
    udp_output.setSampleBytes(2);
    /* a few distinct source blocks (TestSource arithmetic), cycled */
    const int n_src = 8;
    std::vector<IQSampleVector> src(n_src);
    float phase = 0;
    std::vector<int16_t> buf(2 * (size_t)blklen);
    for (int i = 0; i < n_src; i++) {
        int got = 0;
        TestSource::read_samples(buf.data(), 4 * blklen, got, phase, 10000000, 0.0628f, 0.5f, false);
        src[i].resize(blklen);
        memcpy((void*)src[i].data(), buf.data(), (size_t)blklen * 4);
    }
    IQSampleVector outsamples;
    long long frames_bytes = 0;
    struct timeval t0, t1;
    for (int blk = -warmup; blk < n_blocks; blk++) {
        if (blk == 0) {
            udp_output.flush();
            gettimeofday(&t0, 0);
        }
        unsigned int sampleSize = 16;
        dn.process(sampleSize, src[(blk + warmup) % n_src], outsamples);
        udp_output.setSampleBits((uint8_t)sampleSize);
        udp_output.write(outsamples);
        if (!dn || !udp_output) { fprintf(stderr, "error: %s %s\n", dn.error().c_str(), udp_output.error().c_str()); return 1; }
    }
    udp_output.flush();
    gettimeofday(&t1, 0);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_usec - t0.tv_usec);
    (void)frames_bytes;
    printf("{\"blocks\": %d, \"blklen\": %d, \"decim\": %d, \"fecblk\": %d, \"seconds\": %.6f, \"msamples_per_s\": %.3f, "
           "\"us_per_block\": %.2f}\n",
           n_blocks, blklen, decim, fecblk, dt, (double)n_blocks * blklen / dt * 1e-6, dt / n_blocks * 1e6);
    return 0;
}

/* host_pipeline blocksq <n_blocks> <decim> <fecblk> <blklen> <min_chain_blocks> [warmup] [staging_threads]
 * The same blocks through the queued C entry points (sdrd_rx_submit / sdrd_rx_collect) from a native caller: what a
 * main loop gets that hands its blocks over instead of waiting for each (route C). */
static int run_blocksq(int argc, char** argv)
{
    if (argc < 7) return 2;
    const int n_blocks = atoi(argv[2]), decim = atoi(argv[3]), fecblk = atoi(argv[4]), blklen = atoi(argv[5]), min_chain = atoi(argv[6]);
    const int warmup = argc > 7 ? atoi(argv[7]) : 64;
    const int helpers = argc > 8 ? atoi(argv[8]) : 0;
    sdrd_rx* rx = nullptr;
    if (sdrd_rx_create(&rx, decim, SDRD_FC_CENTER, SDRD_HB_EO1, 1, (size_t)blklen * 32) != 0) { fprintf(stderr, "%s\n", sdrd_last_error()); return 1; }
    sdrd_sink_set_nb_fec(sdrd_rx_sink(rx), fecblk);
    sdrd_sink_set_meta(sdrd_rx_sink(rx), 435000, 10000000u >> decim, 2, 16);
    sdrd_rx_set_min_chain(rx, (size_t)min_chain * blklen);
    if (sdrd_rx_set_staging_threads(rx, helpers) != 0) { fprintf(stderr, "%s\n", sdrd_last_error()); return 1; }
    const int n_src = 8;
    std::vector<IQSampleVector> src(n_src);
    float phase = 0;
    std::vector<int16_t> buf(2 * (size_t)blklen);
    for (int i = 0; i < n_src; i++) {
        int got = 0;
        TestSource::read_samples(buf.data(), 4 * blklen, got, phase, 10000000, 0.0628f, 0.5f, false);
        src[i].resize(blklen);
        memcpy((void*)src[i].data(), buf.data(), (size_t)blklen * 4);
    }
    std::vector<uint8_t> out((size_t)64 * (128 + fecblk) * SDRD_UDPSIZE);
    size_t nfr = 0, frames = 0;
    int bpf = 0;
    long long chains0 = 0;
    struct timeval t0, t1;
    for (int blk = -warmup; blk < n_blocks; blk++) {
        if (blk == 0) {
            do { sdrd_rx_collect(rx, out.data(), 64, &nfr, &bpf, 1); } while (nfr);
            chains0 = sdrd_rx_chains(rx);
            gettimeofday(&t0, 0);
        }
        if (sdrd_rx_submit(rx, reinterpret_cast<const int16_t*>(src[(blk + warmup) % n_src].data()), (size_t)blklen, (size_t)blklen, nullptr) != 0 ||
            sdrd_rx_collect(rx, out.data(), 64, &nfr, &bpf, 0) != 0) { fprintf(stderr, "%s\n", sdrd_last_error()); return 1; }
        frames += nfr;
    }
    do {
        if (sdrd_rx_collect(rx, out.data(), 64, &nfr, &bpf, 1) != 0) { fprintf(stderr, "%s\n", sdrd_last_error()); return 1; }
        frames += nfr;
    } while (nfr);
    gettimeofday(&t1, 0);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_usec - t0.tv_usec);
    const long long chains = sdrd_rx_chains(rx) - chains0;
    printf("{\"blocks\": %d, \"blklen\": %d, \"decim\": %d, \"fecblk\": %d, \"seconds\": %.6f, \"msamples_per_s\": %.3f, "
           "\"us_per_block\": %.2f, \"blocks_per_chain\": %.2f, \"frames\": %zu, \"staging_threads\": %d}\n",
           n_blocks, blklen, decim, fecblk, dt, (double)n_blocks * blklen / dt * 1e-6, dt / n_blocks * 1e6,
           (double)n_blocks / (double)(chains > 0 ? chains : 1), frames, helpers);
    sdrd_rx_destroy(rx);
    return 0;
}

/* host_pipeline chain <decim> <fecblk> <blklen>: how long ONE chain of n blocks takes on the device (submit n blocks
 * with the start held back, then time collect(wait): launch + copy in + kernels + copy back), n = 1 .. 32 */
static int run_chain(int argc, char** argv)
{
    if (argc < 5) return 2;
    const int decim = atoi(argv[2]), fecblk = atoi(argv[3]), blklen = atoi(argv[4]);
    sdrd_rx* rx = nullptr;
    if (sdrd_rx_create(&rx, decim, SDRD_FC_CENTER, SDRD_HB_EO1, 1, (size_t)blklen * 32) != 0) { fprintf(stderr, "%s\n", sdrd_last_error()); return 1; }
    sdrd_sink_set_nb_fec(sdrd_rx_sink(rx), fecblk);
    sdrd_sink_set_meta(sdrd_rx_sink(rx), 435000, 10000000u >> decim, 2, 16);
    sdrd_rx_set_min_chain(rx, (size_t)32 * blklen);
    std::vector<int16_t> buf(2 * (size_t)blklen);
    float phase = 0;
    int got = 0;
    TestSource::read_samples(buf.data(), 4 * blklen, got, phase, 10000000, 0.0628f, 0.5f, false);
    std::vector<uint8_t> out((size_t)64 * (128 + fecblk) * SDRD_UDPSIZE);
    size_t nfr = 0;
    int bpf = 0;
    for (int n = 1; n <= 32; n *= 2) {
        double best = 1e9;
        for (int rep = 0; rep < 40; rep++) {
            for (int b = 0; b < n - (n == 32); b++) sdrd_rx_submit(rx, buf.data(), (size_t)blklen, (size_t)blklen, nullptr);
            struct timeval t0, t1;
            gettimeofday(&t0, 0);
            if (n == 32) sdrd_rx_submit(rx, buf.data(), (size_t)blklen, (size_t)blklen, nullptr); /* this one starts the chain */
            do { sdrd_rx_collect(rx, out.data(), 64, &nfr, &bpf, 1); } while (nfr);
            gettimeofday(&t1, 0);
            const double dt = (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_usec - t0.tv_usec);
            if (rep >= 5 && dt < best) best = dt;
        }
        printf("{\"chain_blocks\": %d, \"us_per_chain\": %.1f, \"us_per_block\": %.2f}\n", n, best, best / n);
    }
    sdrd_rx_destroy(rx);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 2 && std::string(argv[1]) == "chain") return run_chain(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "blocksq") return run_blocksq(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "blocks") return run_blocks(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "testsource") return run_testsource(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "upsample") return run_upsample(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "pipeline") return run_pipeline(argc, argv);
    fprintf(stderr, "usage: host_pipeline testsource|upsample|pipeline|blocks ...\n");
    return 2;
}
