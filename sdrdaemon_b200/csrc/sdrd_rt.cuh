/*
 * sdrd_rt.cuh -- the few runtime services the host side of the library needs (device memory,
 * copies, launches), CUDA in the product build, plain host memory under -DSDRD_EMU (tests/emu only,
 * see sdrd_platform.cuh).
 */
#pragma once
#include "sdrd_platform.cuh"
#if !defined(SDRD_EMU)
#include <cuda.h> /* CUtensorMap and the enums of cuTensorMapEncodeTiled: types only, no libcuda symbol is linked */
#endif

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

namespace sdrd {
namespace rt {

enum CopyKind { H2D, D2H, D2D };

#if defined(SDRD_EMU)

typedef void* stream_t;
inline bool device_ok(std::string&) { return true; }
inline int device_count() { return 1; }
inline int set_device(int) { return 0; }
inline int current_device() { return 0; }
inline int sm_count() { return 148; }
inline int alloc(void** p, size_t n)
{
    *p = malloc(n ? n : 1);
    if (!*p) return -1;
    memset(*p, 0xCD, n); /* cudaMalloc does not zero either */
    return 0;
}
inline void release(void* p) { free(p); }
inline int fill(void* p, int v, size_t n, stream_t) { memset(p, v, n); return 0; }
inline int copy(void* d, const void* s, size_t n, CopyKind, stream_t) { memmove(d, s, n); return 0; }
inline int copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, CopyKind, stream_t)
{
    for (size_t i = 0; i < h; i++) memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return 0;
}
inline int sync(stream_t) { return 0; }
inline int sync_device() { return 0; }
inline int stream_create(stream_t* s) { *s = nullptr; return 0; }
inline void stream_destroy(stream_t) {}
typedef void* event_t;
inline int event_create(event_t* e) { *e = nullptr; return 0; }
inline void event_destroy(event_t) {}
inline int event_record(event_t, stream_t) { return 0; }
inline int stream_wait(stream_t, event_t) { return 0; }
inline int event_done(event_t) { return 1; }
inline int event_sync(event_t) { return 0; }
inline int host_alloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : -1; }
inline void host_release(void* p) { free(p); }
inline bool host_is_pageable(const void*) { return true; }
/* inter-process handles: within the emulation a "handle" is the pointer itself */
inline int ipc_export(void* p, void* handle64) { memset(handle64, 0, 64); memcpy(handle64, &p, sizeof p); return 0; }
inline int ipc_open(const void* handle64, void** p) { memcpy(p, handle64, sizeof *p); return 0; }
inline int ipc_close(void*) { return 0; }
inline const char* last_error() { return "emu"; }
/* descriptor of a global array of 32-word rows for tma_store_tile32: in the emulation just the base pointer */
inline int make_tile_map(TileMap* map, void* base, unsigned long long)
{
    memset(map->opaque, 0, sizeof map->opaque);
    memcpy(map->opaque, &base, sizeof base);
    return 0;
}

#define SDRD_LAUNCH(kernel, gx, gy, nthreads, smem, stream, params)                                            \
    sdrd_emu::launch(sdrd_emu::Dim3{(unsigned)(gx), (unsigned)(gy), 1u}, (unsigned)(nthreads), (size_t)(smem), \
                     [&]() { kernel(params); })
#define SDRD_LAUNCH_OK() (true)

#else

typedef cudaStream_t stream_t;
inline const char* last_error() { return cudaGetErrorString(cudaGetLastError()); }
inline int device_count()
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ok++;
    }
    return ok;
}
inline bool device_ok(std::string& why)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        why = std::string("no CUDA device: ") + last_error();
        return false;
    }
    /* cudaGetDeviceProperties costs milliseconds: ask the two attributes needed, once per device */
    static int cached_major[64];
    static bool cached[64];
    int major = 0;
    if (dev >= 0 && dev < 64 && cached[dev]) {
        major = cached_major[dev];
    } else {
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            why = std::string("cudaDeviceGetAttribute: ") + last_error();
            return false;
        }
        if (dev >= 0 && dev < 64) {
            cached_major[dev] = major;
            cached[dev] = true;
        }
    }
    if (major != 10) {
        int minor = 0;
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
        char b[400];
        snprintf(b, sizeof b, "device %d is sm_%d%d; this library carries sm_100a code only", dev, major, minor);
        why = b;
        return false;
    }
    return true;
}
inline int set_device(int d) { return cudaSetDevice(d) == cudaSuccess ? 0 : -1; }
inline int current_device()
{
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
inline int sm_count()
{
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
inline int alloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : -1; }
inline void release(void* p) { if (p) cudaFree(p); }
inline int fill(void* p, int v, size_t n, stream_t s) { return cudaMemsetAsync(p, v, n, s) == cudaSuccess ? 0 : -1; }
inline cudaMemcpyKind kind_of(CopyKind k)
{
    return k == H2D ? cudaMemcpyHostToDevice : k == D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
}
inline int copy(void* d, const void* s, size_t n, CopyKind k, stream_t st)
{
    if (!n) return 0;
    return cudaMemcpyAsync(d, s, n, kind_of(k), st) == cudaSuccess ? 0 : -1;
}
inline int copy2d(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, CopyKind k, stream_t st)
{
    if (!w || !h) return 0;
    /* one row (one stream, the reference's case) or rows that abut: a plain copy costs the host less to enqueue */
    if (h == 1 || (dp == w && sp == w)) return cudaMemcpyAsync(d, s, w * h, kind_of(k), st) == cudaSuccess ? 0 : -1;
    return cudaMemcpy2DAsync(d, dp, s, sp, w, h, kind_of(k), st) == cudaSuccess ? 0 : -1;
}
inline int sync(stream_t s) { return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1; }
inline int sync_device() { return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1; }
inline int stream_create(stream_t* s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking) == cudaSuccess ? 0 : -1; }
inline void stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
typedef cudaEvent_t event_t;
inline int event_create(event_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess ? 0 : -1; }
inline void event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
inline int event_record(event_t e, stream_t s) { return cudaEventRecord(e, s) == cudaSuccess ? 0 : -1; }
inline int stream_wait(stream_t s, event_t e) { return cudaStreamWaitEvent(s, e, 0) == cudaSuccess ? 0 : -1; }
/* 1: everything recorded before the event has completed, 0: not yet, -1: error */
inline int event_done(event_t e)
{
    const cudaError_t r = cudaEventQuery(e);
    return r == cudaSuccess ? 1 : (r == cudaErrorNotReady ? 0 : -1);
}
inline int event_sync(event_t e) { return cudaEventSynchronize(e) == cudaSuccess ? 0 : -1; }
/* page-locked host memory (staging rings of the queued entry points) */
inline int host_alloc(void** p, size_t n) { return cudaHostAlloc(p, n ? n : 1, cudaHostAllocDefault) == cudaSuccess ? 0 : -1; }
inline void host_release(void* p) { if (p) cudaFreeHost(p); }
inline int ipc_export(void* p, void* handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    return cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p) == cudaSuccess ? 0 : -1;
}
inline int ipc_open(const void* handle64, void** p)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    return cudaIpcOpenMemHandle(p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? 0 : -1;
}
inline int ipc_close(void* p) { return cudaIpcCloseMemHandle(p) == cudaSuccess ? 0 : -1; }
/* true for ordinary (unregistered) host memory: copies from / to it are staged by the driver, synchronously */
inline bool host_is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

/* CUtensorMap over a global array of n_rows rows of 32 words (128 bytes), box 32 x 32, 128-byte swizzle: what
 * tma_store_tile32 stores through.  cuTensorMapEncodeTiled is a driver entry point: fetched through the runtime, so
 * the library does not link libcuda. */
inline int make_tile_map(TileMap* map, void* base, unsigned long long n_rows)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -1;
        encode = (encode_fn)fn;
    }
    static_assert(sizeof(CUtensorMap) == sizeof(TileMap), "a CUtensorMap is 128 bytes");
    const cuuint64_t dims[2] = {32, n_rows};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
    return encode(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
               ? 0
               : -1;
}

/* the opt-in for more than 48 KB of dynamic shared memory is per function and device: set it when a launch site
 * first needs that much on a device, not on every launch (a microsecond each on the small-block path) */
#define SDRD_LAUNCH(kernel, gx, gy, nthreads, smem, stream, params)                                       \
    do {                                                                                                  \
        static int smem_set_[64];                                                                         \
        int dev_ = 0;                                                                                     \
        cudaGetDevice(&dev_);                                                                             \
        if ((int)(smem) > smem_set_[dev_ & 63]) {                                                         \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));       \
            smem_set_[dev_ & 63] = (int)(smem);                                                           \
        }                                                                                                 \
        kernel<<<dim3((unsigned)(gx), (unsigned)(gy), 1), dim3((unsigned)(nthreads), 1, 1), (smem), (stream)>>>(params); \
    } while (0)
#define SDRD_LAUNCH_OK() (cudaPeekAtLastError() == cudaSuccess)

#endif

} /* namespace rt */
} /* namespace sdrd */
