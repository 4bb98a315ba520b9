/*
 * ref_seam_common.h -- shared by the two "reference code over the GPU seam" programs
 * (ref_seam_loopback.cpp, ref_seam_gr.cpp).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef REF_SEAM_COMMON_H
#define REF_SEAM_COMMON_H
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace seam {
static const int FRAME_SAMPLES = 127 * 127;

inline uint64_t splitmix64(uint64_t& x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* frame f of the test signal: 16129 random int16 I/Q pairs */
inline void gen_frame(uint64_t seed, int f, int16_t* iq)
{
    uint64_t st = seed * 1000003ull + (uint64_t)f;
    for (int i = 0; i < FRAME_SAMPLES; i++) {
        const uint64_t r = splitmix64(st);
        iq[2 * i] = (int16_t)(r & 0xFFFF);
        iq[2 * i + 1] = (int16_t)((r >> 16) & 0xFFFF);
    }
}

/* does the harness drop original block `block` of frame `frame_index` on the way to the receiver?
 * `extra` blocks per frame, chosen by a hash of the frame index; never the block the sender punctures itself */
inline bool dropped(uint64_t seed, int frame_index, int block, int extra, int puncture)
{
    if (block >= 128 || extra <= 0) return false;
    uint64_t st = seed ^ (0xD1CEull << 32) ^ (uint64_t)frame_index;
    bool pick[128] = {false};
    int n = 0;
    while (n < extra) {
        const int b = (int)(splitmix64(st) % 128);
        if (b == puncture || pick[b]) continue;
        pick[b] = true;
        n++;
    }
    return pick[block];
}

inline uint32_t fold(uint32_t d, const void* p, size_t bytes)
{
    const uint8_t* b = (const uint8_t*)p;
    for (size_t i = 0; i < bytes; i++) d = (d ^ b[i]) * 16777619u;
    return d;
}
} /* namespace seam */
#endif
