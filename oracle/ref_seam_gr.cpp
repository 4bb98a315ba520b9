/*
 * ref_seam_gr.cpp -- gr-sdrdaemon's own receiver buffer (gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp, the GNU Radio
 * client of the same wire format, which uses the C API of cm256: cm256_init / cm256_decode,
 * gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp:40,191) compiled unmodified and fed with the datagram stream that
 * ref_seam_loopback captured from the reference sender.  Built against oracle/ref_stubs (CPU restatement) and
 * against include/cm256.h + the emulation / product library.  TEST INFRASTRUCTURE ONLY.
 *
 *   ref_seam_gr <capture-file> <n_frames> <extra_drop> <seed> <puncture>
 */
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "SDRdaemonFECBuffer.h" /* gr-sdrdaemon/lib */
#include "ref_seam_common.h"

int main(int argc, char** argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s capture-file n_frames extra_drop seed puncture\n", argv[0]);
        return 2;
    }
    FILE* cap = fopen(argv[1], "rb");
    if (!cap) {
        fprintf(stderr, "cannot open %s\n", argv[1]);
        return 2;
    }
    const int n_frames = atoi(argv[2]), extra = atoi(argv[3]), puncture = atoi(argv[5]);
    const uint64_t seed = strtoull(argv[4], 0, 10);
    std::ostringstream ref_log;
    std::streambuf* old_cerr = std::cerr.rdbuf(ref_log.rdbuf());
    SDRdaemonFECBuffer* fecbuf = new SDRdaemonFECBuffer();
    std::vector<int16_t> want(2 * seam::FRAME_SAMPLES);
    std::vector<uint8_t> data(128 * 512);
    uint8_t dg[512];
    int head = -1, checked = 0, ok = 0;
    uint32_t digest = 2166136261u;
    bool first = true;
    while (fread(dg, 1, 512, cap) == 512) {
        const int fi = dg[0] | (dg[1] << 8), bi = dg[2];
        if (seam::dropped(seed, fi, bi, extra, puncture)) continue;
        uint32_t len = 0;
        if (fecbuf->writeAndRead(dg, 512, data.data(), len)) {
            if (!first && head >= 0 && head < n_frames) {
                seam::gen_frame(seed, head, want.data());
                checked++;
                if (len == want.size() * 2 && memcmp(data.data(), want.data(), len) == 0) ok++;
                digest = seam::fold(digest, data.data(), len);
            }
            first = false;
        }
        head = fi;
    }
    fclose(cap);
    delete fecbuf;
    std::cerr.rdbuf(old_cerr);
    const std::string log = ref_log.str();
    size_t n_success = 0, n_error = 0;
    for (size_t p = 0; (p = log.find("CM256 decode success", p)) != std::string::npos; p++) n_success++;
    for (size_t p = 0; (p = log.find("CM256 decode error", p)) != std::string::npos; p++) n_error++;
    printf("{\"frames_checked\": %d, \"frames_ok\": %d, \"decode_success\": %zu, \"decode_error\": %zu, \"digest\": %u}\n", checked,
           ok, n_success, n_error, digest);
    return (checked >= n_frames && ok == checked && n_error == 0) ? 0 : 1;
}
