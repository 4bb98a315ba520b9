/* Stand-in for <nanomsg/nn.h>: the control socket is out of scope; these
 * no-ops only let the reference's DeviceSource.h / TestSource.cpp compile unmodified. */
#ifndef SDRD_STUB_NN_H
#define SDRD_STUB_NN_H
#include <cstddef>
#define AF_SP 1
#define NN_MSG ((size_t)-1)
#define NN_DONTWAIT 1
static inline int nn_socket(int, int) { return 0; }
static inline int nn_bind(int, const char*) { return 0; }
static inline int nn_recv(int, void*, size_t, int) { return -1; }
static inline int nn_send(int, const void*, size_t, int) { return 0; }
static inline int nn_freemsg(void*) { return 0; }
#endif
