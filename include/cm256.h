/*
 * cm256.h -- cm256cc's interface (f4exb/cm256cc `class CM256`, and the C functions of f4exb/cm256 that
 * gr-sdrdaemon uses) bound to libsdrd_b200.so, the B200 implementation of sdrdaemon's FEC hot path.
 *
 * Put this directory in front of the include path instead of cm256cc's and link with -lsdrd_b200: the
 * reference's sdmnbase/UDPSinkFEC.cpp, sdmnbase/SDRdaemonFECBuffer.cpp and
 * gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp then compile UNMODIFIED and run their encode / decode on the GPU
 * (oracle/Makefile `seam` builds exactly that; tests/test_reference_seam.py runs it).
 *
 * What the reference uses of the library (and therefore what is here):
 *   CM256::isInitialized()                                   UDPSinkFEC.cpp:38, SDRdaemonFECBuffer.cpp:42
 *   CM256::cm256_encoder_params {OriginalCount, RecoveryCount, BlockBytes}
 *   CM256::cm256_block {void* Block; unsigned char Index}    UDPSinkFEC.cpp:195-196,228-243
 *   CM256::cm256_encode(params, originals, recoveryBlocks)   UDPSinkFEC.cpp:246
 *   CM256::cm256_decode(params, blocks)                      SDRdaemonFECBuffer.cpp:197
 *   cm256_init(), cm256_encode(), cm256_decode(), cm256_encoder_params, cm256_block   (C form,
 *                                                            gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp:40,191)
 * Return values as in cm256: 0 = success.  Shapes other than sdrdaemon's 128-block superframe of up to 508
 * bytes are refused (non-zero), see sdrd_cm256_encode_blocks in sdrd_b200.h.  One superframe per call is the
 * library's own granularity; callers that can batch should use sdrd_sink_* / sdrd_fec_decode / sdrd_src_feed.
 */
#ifndef SDRD_B200_CM256_H
#define SDRD_B200_CM256_H

#include "sdrd_b200.h"

#define CM256_VERSION 2

/* C form (f4exb/cm256) */
typedef sdrd_cm256_params cm256_encoder_params;
typedef sdrd_cm256_block cm256_block;

#ifdef __cplusplus
extern "C" {
#endif
/* cm256_init() of the C library: 0 when the codec can be used -- here: when an sm_100 device is present */
static inline int cm256_init(void) { return sdrd_device_count() > 0 ? 0 : -1; }
/* -DSDRD_CM256_TRACE: a call that fails says why on stderr (cm256 itself only returns non-zero; the reference's
 * sender then stops transmitting without a reason, UDPSinkFEC.cpp:246-250) */
#ifdef SDRD_CM256_TRACE
#include <stdio.h>
#define SDRD_CM256_REPORT(what, rc)                                                       \
    do {                                                                                  \
        if (rc) fprintf(stderr, "cm256 (sdrd_b200) %s: %s\n", what, sdrd_last_error());   \
    } while (0)
#else
#define SDRD_CM256_REPORT(what, rc) ((void)0)
#endif
static inline int cm256_encode(cm256_encoder_params params, cm256_block* originals, void* recoveryBlocks)
{
    const int rc = sdrd_cm256_encode_blocks(params, originals, recoveryBlocks);
    SDRD_CM256_REPORT("encode", rc);
    return rc;
}
static inline int cm256_decode(cm256_encoder_params params, cm256_block* blocks)
{
    const int rc = sdrd_cm256_decode_blocks(params, blocks);
    SDRD_CM256_REPORT("decode", rc);
    return rc;
}
#ifdef __cplusplus
}

/* C++ form (f4exb/cm256cc) */
class CM256 {
public:
    typedef ::cm256_encoder_params cm256_encoder_params;
    typedef ::cm256_block cm256_block;

    CM256() : m_initialized(::cm256_init() == 0) {}
    bool isInitialized() const { return m_initialized; }
    int cm256_encode(cm256_encoder_params params, cm256_block* originals, void* recoveryBlocks)
    {
        return ::cm256_encode(params, originals, recoveryBlocks);
    }
    int cm256_decode(cm256_encoder_params params, cm256_block* blocks) { return ::cm256_decode(params, blocks); }

private:
    bool m_initialized;
};
#endif

#endif /* SDRD_B200_CM256_H */
