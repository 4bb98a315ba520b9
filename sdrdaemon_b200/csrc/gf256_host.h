/*
 * gf256_host.h -- host-side GF(2^8) helpers of the product library: field tables, the CM256
 * Cauchy matrix and the per-constant lookup tables the device kernels multiply with.
 *
 * cm256cc (f4exb/cm256cc, C++ port of catid/cm256) is an external dependency of the reference
 * (CMakeLists.txt:97, cm256cc/CMakeLists.txt:12-20) and is not in its tree; the arithmetic is
 * restated from the published algorithm: GF(2^8) modulo x^8+x^6+x^3+x^2+1 (0x14D, gf256 polynomial
 * index 3), generator 2, and the Cauchy element  M[x_i][y_j] = (y_j ^ x_0) / (x_i ^ y_j)  with
 * x_0 = OriginalCount, x_i = OriginalCount + recovery row, y_j = original column.
 */
#pragma once
#include <stdint.h>
#include <string.h>

namespace sdrd {
namespace gf {

struct Field {
    uint8_t exp[512];
    uint8_t log[256];
    Field()
    {
        unsigned v = 1;
        for (int i = 0; i < 255; i++) {
            exp[i] = (uint8_t)v;
            log[v] = (uint8_t)i;
            v <<= 1;
            if (v & 0x100) v ^= 0x14Du;
        }
        for (int i = 255; i < 512; i++) exp[i] = exp[i - 255];
        log[0] = 0;
    }
    uint8_t mul(uint8_t a, uint8_t b) const { return (a && b) ? exp[log[a] + log[b]] : 0; }
    uint8_t div(uint8_t a, uint8_t b) const { return a ? exp[log[a] + 255 - log[b]] : 0; }
    uint8_t inv(uint8_t a) const { return exp[255 - log[a]]; }
    /* CM256 matrix element for 128 originals: recovery row index x (128..255), column y (0..127) */
    uint8_t cauchy(uint8_t x, uint8_t y) const { return div((uint8_t)(y ^ 128), (uint8_t)(x ^ y)); }
};

inline const Field& field()
{
    static const Field f;
    return f;
}

/* Device multiply tables.  A byte b is split into 3+3+2 bits; for a constant c
 *   c*b = T0[b & 7] ^ T1[(b >> 3) & 7] ^ T2[b >> 6]
 * and the three look-ups are done four bytes at a time with the byte-permute unit (PRMT):
 * tabA[c] = {T0[0..3], T0[4..7], T1[0..3], T1[4..7]} (little-endian packed), tabB[c] = T2[0..3]. */
struct MulTables {
    uint32_t tabA[256][4];
    uint32_t tabB[256];
    /* log/antilog tables for scalar work on the device (matrix inversion in the decoder) */
    uint8_t exp[512];
    uint8_t log[256];
};

inline void build_mul_tables(MulTables& t)
{
    const Field& f = field();
    for (int c = 0; c < 256; c++) {
        uint32_t w[5] = {0, 0, 0, 0, 0};
        for (int v = 0; v < 8; v++) {
            w[v >> 2] |= (uint32_t)f.mul((uint8_t)c, (uint8_t)v) << (8 * (v & 3));
            w[2 + (v >> 2)] |= (uint32_t)f.mul((uint8_t)c, (uint8_t)(v << 3)) << (8 * (v & 3));
        }
        for (int v = 0; v < 4; v++) w[4] |= (uint32_t)f.mul((uint8_t)c, (uint8_t)(v << 6)) << (8 * v);
        memcpy(t.tabA[c], w, 16);
        t.tabB[c] = w[4];
    }
    memcpy(t.exp, f.exp, 512);
    memcpy(t.log, f.log, 256);
}

/* Encoder matrix, row-major [128][128]: row r = recovery block 128 + r (row 0 is all ones). */
inline void build_cauchy_matrix(uint8_t* m)
{
    const Field& f = field();
    for (int r = 0; r < 128; r++)
        for (int j = 0; j < 128; j++) m[r * 128 + j] = f.cauchy((uint8_t)(128 + r), (uint8_t)j);
}

} /* namespace gf */
} /* namespace sdrd */
