/* UDPSinkFEC.h -- the reference's header name (f4exb/sdrdaemon include/UDPSinkFEC.h) resolved to the B200 host layer.  Put this
 * directory in FRONT of the reference's include/ (-Icompat -I<reference>/include) and link with -lsdrd_b200:
 * sdrdaemonrx.cpp / sdrdaemontx.cpp then compile unchanged, with the reference's own glue types (SDRDaemon.h,
 * DataBuffer.h, util.h) and this library's compute classes.  oracle/Makefile `mains` and
 * tests/test_reference_mains.py do exactly that. */
#ifndef SDRD_B200_COMPAT_HOST_H
#define SDRD_B200_COMPAT_HOST_H
#define SDRD_HOST_REFERENCE_TYPES 1
#include "../sdrd_host.hpp"
using namespace sdrd_b200;
#endif
