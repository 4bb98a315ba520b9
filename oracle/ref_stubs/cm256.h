/* Stand-in for cm256cc's cm256.h (the library is NOT in /root/reference and not installed).
 * Exposes exactly the class API the reference uses -- CM256::isInitialized, cm256_encode,
 * cm256_decode, cm256_encoder_params, cm256_block (UDPSinkFEC.cpp:38,195-196,228-246;
 * SDRdaemonFECBuffer.cpp:42,148-163,197) -- and forwards to the restated arithmetic in
 * oracle/sdrd_oracle.c.  PARITY UNPINNED for the GF(256) arithmetic itself. */
#ifndef SDRD_STUB_CM256_H
#define SDRD_STUB_CM256_H
#include "../sdrd_oracle.h"
/* the C form (f4exb/cm256) that gr-sdrdaemon/lib/SDRdaemonFECBuffer.cpp:40,191 uses */
typedef sdro_cm256_params cm256_encoder_params;
typedef sdro_cm256_block cm256_block;
static inline int cm256_init(void) { return 0; }
static inline int cm256_encode(cm256_encoder_params params, cm256_block* originals, void* recoveryBlocks)
{ return sdro_cm256_encode(params, originals, recoveryBlocks); }
static inline int cm256_decode(cm256_encoder_params params, cm256_block* blocks)
{ return sdro_cm256_decode(params, blocks); }
class CM256 {
public:
    typedef sdro_cm256_params cm256_encoder_params;
    typedef sdro_cm256_block cm256_block;
    CM256() {}
    bool isInitialized() const { return true; }
    int cm256_encode(cm256_encoder_params params, cm256_block* originals, void* recoveryBlocks)
    { return sdro_cm256_encode(params, originals, recoveryBlocks); }
    int cm256_decode(cm256_encoder_params params, cm256_block* blocks)
    { return sdro_cm256_decode(params, blocks); }
};
#endif
