#ifndef SDRD_STUB_NN_PAIR_H
#define SDRD_STUB_NN_PAIR_H
#define NN_PAIR 16
#endif
