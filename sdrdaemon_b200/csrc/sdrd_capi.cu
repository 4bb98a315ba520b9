/*
 * sdrd_capi.cu -- implementation of the C ABI declared in include/sdrd_b200.h: handle management,
 * launch geometry and the host<->device plumbing around the kernels in hb_decimate.cuh and
 * fec_kernels.cuh.  No arithmetic of the hot path happens on the host.
 *
 * (The same file is compiled with -DSDRD_EMU by tests/emu to check the host logic and the kernels'
 * indexing in a container without a GPU; that build is never part of libsdrd_b200.so.)
 */
#include "../../include/sdrd_b200.h"

#include "fec_kernels.cuh"
#include "gf256_host.h"
#include "hb_decimate.cuh"
#include "hb_interpolate.cuh"
#include "sdrd_rt.cuh"

#include <sys/time.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <new>
#include <string>
#include <vector>

using namespace sdrd;

#if defined(SDRD_EMU)
thread_local sdrd_emu::Cta* sdrd_emu::t_cta = nullptr;
thread_local sdrd_emu::Dim3 sdrd_emu::t_tid;
#endif

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
int fail_cuda(const char* what) { return fail(SDRD_ECUDA, std::string(what) + ": " + rt::last_error()); }

#define SDRD_TRY(expr, what)                  \
    do {                                      \
        if ((expr) != 0) return fail_cuda(what); \
    } while (0)

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

/* A handle's buffers and streams live on the device that was current when it was created: every entry point
 * that takes a handle makes that device current for its duration, whatever the calling thread had selected. */
struct DeviceGuard {
    int prev;
    bool switched = false;
    explicit DeviceGuard(int dev) : prev(rt::current_device())
    {
        if (dev >= 0 && dev != prev) switched = rt::set_device(dev) == 0;
    }
    ~DeviceGuard()
    {
        if (switched) rt::set_device(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define SDRD_ON_DEVICE_OF(h) DeviceGuard device_guard_((h)->device)

/* Small transfers between PAGEABLE host memory and the device (the reference's callers hand over std::vectors of
 * 65536 samples): the driver stages such copies itself and blocks while it does; going through the handle's own
 * page-locked buffer -- one memcpy, then an asynchronous copy -- costs the calling thread less.  Large transfers
 * and page-locked user buffers go directly. */
struct HostStage {
    static constexpr size_t MAX_BYTES = (size_t)4 << 20;
    uint8_t* in = nullptr;
    uint8_t* out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    /* a device -> host copy that still has to be unpacked into the user's buffer after the stream has been synchronised */
    uint8_t* pend_dst = nullptr;
    size_t pend_dpitch = 0, pend_w = 0, pend_h = 0;
    ~HostStage()
    {
        rt::host_release(in);
        rt::host_release(out);
    }
    static int grow(uint8_t** p, size_t* cap, size_t need)
    {
        if (need <= *cap) return 0;
        rt::host_release(*p);
        *p = nullptr;
        *cap = 0;
        const size_t n = std::max<size_t>(need, (size_t)1 << 20);
        if (rt::host_alloc((void**)p, n) != 0) return -1;
        *cap = n;
        return 0;
    }
    int to_device(void* dst, size_t dpitch, const void* src, size_t spitch, size_t w, size_t h, rt::stream_t st)
    {
        if (!w || !h) return 0;
        if (w * h <= MAX_BYTES && rt::host_is_pageable(src) && grow(&in, &in_cap, w * h) == 0) {
            for (size_t i = 0; i < h; i++) memcpy(in + i * w, (const uint8_t*)src + i * spitch, w);
            return rt::copy2d(dst, dpitch, in, w, w, h, rt::H2D, st);
        }
        return rt::copy2d(dst, dpitch, src, spitch, w, h, rt::H2D, st);
    }
    int to_host(void* dst, size_t dpitch, const void* src, size_t spitch, size_t w, size_t h, rt::stream_t st)
    {
        if (!w || !h) return 0;
        pend_dst = nullptr; /* one staged copy per call; a call that failed half-way must not leave its destination behind */
        if (w * h <= MAX_BYTES && rt::host_is_pageable(dst) && grow(&out, &out_cap, w * h) == 0) {
            pend_dst = (uint8_t*)dst;
            pend_dpitch = dpitch;
            pend_w = w;
            pend_h = h;
            return rt::copy2d(out, w, src, spitch, w, h, rt::D2H, st);
        }
        return rt::copy2d(dst, dpitch, src, spitch, w, h, rt::D2H, st);
    }
    /* after the stream has been synchronised */
    void finish()
    {
        if (!pend_dst) return;
        for (size_t i = 0; i < pend_h; i++) memcpy(pend_dst + i * pend_dpitch, out + i * pend_w, pend_w);
        pend_dst = nullptr;
    }
};

/* raw samples of input history kept per stream: two chunks of the /4-prologue cascade */
constexpr size_t HISTW = 16384;
constexpr size_t MAX_CHUNK_RAW = 4 * 2048; /* largest raw chunk the kernel reads: /4 prologue, C0 = 2048 */
/* Raw samples after which a cascade no longer reaches back across a reconfiguration: the deepest cascade (six
 * stages) looks back 61 * 63 = 3843 raw samples, and the last 64 inputs of its last stage depend on 3939. */
constexpr long long DEC_HEAD = 4096;
constexpr long long INT_HEAD = 128; /* interpolator: five stages look back 42 input samples */

/* CRC-32/IEEE (boost::crc_32_type, UDPSinkFEC.cpp:106-109) over the 20 meta bytes -- 20 bytes per
 * call, host side like the reference */
uint32_t crc32_ieee(const uint8_t* p, size_t n)
{
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
    }
    return c ^ 0xFFFFFFFFu;
}

/* ---- per-device GF(256) tables ---------------------------------------------------------- */
struct DeviceTables {
    fec::Tables t;
    void* blob = nullptr;
};
std::mutex g_tab_mutex;
DeviceTables g_tables[64];
bool g_tables_ready[64];

int current_device()
{
#if defined(SDRD_EMU)
    return 0;
#else
    int d = 0;
    cudaGetDevice(&d);
    return d;
#endif
}

int get_tables(fec::Tables* out)
{
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    const int dev = current_device();
    if (dev < 0 || dev >= 64) return fail(SDRD_EINVAL, "device index out of range");
    if (!g_tables_ready[dev]) {
        struct Blob {
            uint32_t tabA[256][4];
            uint32_t tabB[256];
            uint8_t cauchy[128 * 128];
            uint8_t ex[512];
            uint8_t lg[256];
        };
        std::vector<uint8_t> hostbuf(sizeof(Blob));
        Blob* hb = reinterpret_cast<Blob*>(hostbuf.data());
        gf::MulTables mt;
        gf::build_mul_tables(mt);
        memcpy(hb->tabA, mt.tabA, sizeof(mt.tabA));
        memcpy(hb->tabB, mt.tabB, sizeof(mt.tabB));
        memcpy(hb->ex, mt.exp, 512);
        memcpy(hb->lg, mt.log, 256);
        gf::build_cauchy_matrix(hb->cauchy);
        void* d = nullptr;
        if (rt::alloc(&d, sizeof(Blob)) != 0) return fail_cuda("alloc GF tables");
        if (rt::copy(d, hb, sizeof(Blob), rt::H2D, 0) != 0 || rt::sync(0) != 0) return fail_cuda("upload GF tables");
        Blob* db = reinterpret_cast<Blob*>(d);
        g_tables[dev].blob = d;
        g_tables[dev].t.tabA = reinterpret_cast<const uint4*>(&db->tabA[0][0]);
        g_tables[dev].t.tabB = &db->tabB[0];
        g_tables[dev].t.cauchy = &db->cauchy[0];
        g_tables[dev].t.gfexp = &db->ex[0];
        g_tables[dev].t.gflog = &db->lg[0];
        g_tables_ready[dev] = true;
    }
    *out = g_tables[dev].t;
    return 0;
}

/* the reference's shift rule shared by every decimateN routine (Decimators.cpp:408-409,515) */
void shift_rule(unsigned ss, int log2_decim, int* norm_shift, int* trunk_shift, unsigned* ss_out)
{
    const unsigned thresh = 16u - (unsigned)log2_decim;
    const unsigned trunk = ss < thresh ? 0u : ss - thresh;
    const unsigned norm = ss < thresh ? thresh - ss : 0u;
    *norm_shift = (int)norm;
    *trunk_shift = (int)trunk;
    *ss_out = ss + (unsigned)log2_decim - trunk;
}

/* persistent encode grid: equal shares over one 512-thread CTA per SM, or -- batches of more than one frame per SM,
 * when two fit -- over two 256-thread CTAs per SM (fec::EncShape) */
bool enc_two(long long items, int cstride) { return items > rt::sm_count() && fec::enc_two_fits(cstride); }
int enc_grid(long long items, bool two)
{
    const long long slots = (long long)rt::sm_count() * (two ? 2 : 1);
    if (items <= slots) return (int)items;
    const long long per = (items + slots - 1) / slots;
    return (int)((items + per - 1) / per);
}
void launch_encode(const fec::EncParams& p, long long items, rt::stream_t st)
{
    if (enc_two(items, p.cstride))
        SDRD_LAUNCH(fec::encode_kernel<true>, enc_grid(items, true), 1, fec::EncShape<true>::NT, fec::enc_smem_bytes(p.cstride, true), st, p);
    else
        SDRD_LAUNCH(fec::encode_kernel<false>, enc_grid(items, false), 1, fec::EncShape<false>::NT, fec::enc_smem_bytes(p.cstride, false), st, p);
}

/* persistent decode grid: SDRD_K3_CTAS_PER_SM two-warp CTAs per SM, each walking frames blockIdx.x, + gridDim.x, .. */
#ifndef SDRD_K3_GRID_WAVES
#define SDRD_K3_GRID_WAVES 1000 /* CTAs launched per resident slot: 1 = persistent (measured slower: the static shares leave a tail), large = one CTA per frame */
#endif
int dec_grid(long long n_frames)
{
    const long long slots = (long long)rt::sm_count() * SDRD_K3_CTAS_PER_SM;
    return (int)(n_frames < slots * SDRD_K3_GRID_WAVES ? n_frames : slots * SDRD_K3_GRID_WAVES);
}

template <int M, int PRO>
void launch_decimate_warp2(const hb::Params& p, int n_seg, rt::stream_t st)
{
    if (p.round_add)
        SDRD_LAUNCH((hb::decimate_warp_kernel<M, 1, PRO>), n_seg, 1, 32, hb::wsmem_bytes(M, PRO), st, p);
    else
        SDRD_LAUNCH((hb::decimate_warp_kernel<M, 0, PRO>), n_seg, 1, 32, hb::wsmem_bytes(M, PRO), st, p);
}
template <int M>
void launch_decimate_warp(const hb::Params& p, int n_seg, rt::stream_t st)
{
    if constexpr (M <= 4) { /* the /4 prologue leaves at most 4 half-band stages */
        if (p.prologue) return launch_decimate_warp2<M, 1>(p, n_seg, st);
    }
    launch_decimate_warp2<M, 0>(p, n_seg, st);
}

} /* namespace */

/* ========================================================================================== */

struct sdrd_dec {
    int log2_decim = 0, fcpos = SDRD_FC_CENTER, variant = SDRD_HB_EO1, S = 1;
    int device = -1;             /* the device the handle lives on */
    size_t max_in = 0;
    uint32_t* d_in = nullptr;    /* [S][in_pitch]: HISTW history words, then the new samples */
    uint32_t* d_in_alt = nullptr; /* a second buffer of the same shape, allocated by the queued Rx path: a chain's samples are
                                     copied into it while the chain before still reads d_in, then the two swap roles */
    uint32_t* d_hist = nullptr;  /* [S][HISTW] */
    bool hist_in_front = false;  /* the history already sits in d_in[0 .. HISTW) (a call of >= HISTW samples moves its
                                    tail there directly); otherwise it is in d_hist and is restored at the next call */
    uint32_t* d_out = nullptr;   /* [S][out_pitch] */
    size_t in_pitch = 0, out_pitch = 0;
    long long consumed = 0;      /* raw samples consumed since reset (saturating) */
    /* Per-stage state across configure (the reference's six persistent stage objects, Decimators.h:57-62).
     * Either `consistent`: every stage of the current cascade is in the state the raw history implies (the warp
     * kernel re-derives it from d_hist) and d_state holds the true state of the OTHER stages; or not: d_state
     * holds the true state of all six stages and hb::stateful_kernel runs the cascade from it until `run` raw
     * samples of the current configuration (>= DEC_HEAD) have made the two views agree again. */
    int* d_state = nullptr;      /* [S][hb::STATE_WORDS] */
    bool consistent = true;
    long long run = 0;           /* raw samples consumed under the current configuration (saturating) */
    long long launches = 0;
    int sms = 148;
    rt::stream_t stream = 0;
    HostStage stage;
};

extern "C" const char* sdrd_last_error(void) { return g_err.c_str(); }
extern "C" const char* sdrd_version(void) { return "sdrd_b200 0.1 (sm_100a)"; }
extern "C" int sdrd_device_count(void) { return rt::device_count(); }
extern "C" int sdrd_set_device(int device)
{
    if (rt::set_device(device) != 0) return fail_cuda("cudaSetDevice");
    return 0;
}

static int check_decim(int log2_decim, int fcpos)
{
    if (log2_decim < 0 || log2_decim > 6) return fail(SDRD_EINVAL, "Invalid log2 decimation factor"); /* Downsampler.cpp:39-43 */
    if (fcpos < 0 || fcpos > 2) return fail(SDRD_EINVAL, "Invalid Fc position index");               /* :57-61 */
    return 0;
}

extern "C" int sdrd_dec_create(sdrd_dec** out, int log2_decim, int fcpos, int variant, int n_streams, size_t max_in)
{
    if (!out) return fail(SDRD_EINVAL, "null handle pointer");
    *out = nullptr;
    if (int rc = check_decim(log2_decim, fcpos)) return rc;
    if (variant != SDRD_HB_EO1 && variant != SDRD_HB_DB) return fail(SDRD_EINVAL, "Invalid half-band variant");
    if (n_streams < 1 || max_in < 1) return fail(SDRD_EINVAL, "n_streams and max_in must be positive");
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    sdrd_dec* d = new (std::nothrow) sdrd_dec();
    if (!d) return fail(SDRD_ENOMEM, "out of host memory");
    d->log2_decim = log2_decim;
    d->fcpos = fcpos;
    d->variant = variant;
    d->S = n_streams;
    d->device = rt::current_device();
    d->max_in = max_in;
    d->sms = rt::sm_count();
    /* the kernel reads whole chunks (up to 4*C0 raw samples with the /4 prologue) */
    d->in_pitch = HISTW + round_up(max_in, MAX_CHUNK_RAW) + MAX_CHUNK_RAW;
    d->out_pitch = round_up(max_in, 8) + 8;
    if (rt::alloc((void**)&d->d_in, d->in_pitch * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&d->d_hist, HISTW * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&d->d_out, d->out_pitch * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&d->d_state, (size_t)hb::STATE_WORDS * 4 * (size_t)n_streams) != 0 || rt::stream_create(&d->stream) != 0) {
        int rc = fail_cuda("allocating decimator buffers");
        sdrd_dec_destroy(d);
        return rc;
    }
    /* bytes past the valid samples are read by the last chunk (results discarded): keep them defined */
    rt::fill(d->d_in, 0, d->in_pitch * 4 * (size_t)n_streams, d->stream);
    if (int rc = sdrd_dec_reset(d)) {
        sdrd_dec_destroy(d);
        return rc;
    }
    *out = d;
    return 0;
}

extern "C" void sdrd_dec_destroy(sdrd_dec* d)
{
    if (!d) return;
    SDRD_ON_DEVICE_OF(d);
    rt::sync(d->stream);
    rt::release(d->d_in);
    rt::release(d->d_in_alt);
    rt::release(d->d_hist);
    rt::release(d->d_out);
    rt::release(d->d_state);
    rt::stream_destroy(d->stream);
    delete d;
}

extern "C" int sdrd_dec_reset(sdrd_dec* d)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    SDRD_TRY(rt::sync_device(), "reset history"); /* whatever stream the caller's last process call used */
    SDRD_TRY(rt::fill(d->d_hist, 0, HISTW * 4 * (size_t)d->S, d->stream), "reset history");
    SDRD_TRY(rt::fill(d->d_state, 0, (size_t)hb::STATE_WORDS * 4 * (size_t)d->S, d->stream), "reset stage states");
    SDRD_TRY(rt::sync(d->stream), "reset history");
    d->hist_in_front = false;
    d->consumed = 0;
    d->consistent = true;
    d->run = 0;
    return 0;
}

/* half-band stages and /4 prologue of a configuration (0 stages: the filter-less routines) */
static void dec_shape(int log2_decim, int fcpos, int* M, int* pro)
{
    *pro = fcpos == SDRD_FC_CENTER ? 0 : (fcpos == SDRD_FC_INFRA ? 1 : 2);
    if (log2_decim == 0 || (*pro && log2_decim <= 2)) {
        *M = 0;
        *pro = 0;
    } else {
        *M = *pro ? log2_decim - 2 : log2_decim;
    }
}

extern "C" int sdrd_dec_configure(sdrd_dec* d, int log2_decim, int fcpos)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    if (int rc = check_decim(log2_decim, fcpos)) return rc;
    if (log2_decim == d->log2_decim && fcpos == d->fcpos) return 0;
    if (d->consumed > 0) {
        /* Leaving a configuration: the stages it ran keep their state in the reference (Decimators.h:57-62).  While
         * `consistent` that state exists only implicitly, as the raw history: make it explicit by running the old
         * cascade over the last min(run, DEC_HEAD) raw samples, starting from the states as they were when the run
         * began (run < DEC_HEAD only happens for the first run after a reset: zeros). */
        int M, pro;
        dec_shape(d->log2_decim, d->fcpos, &M, &pro);
        if (d->consistent && M > 0 && d->run > 0) {
            SDRD_TRY(rt::sync_device(), "configure"); /* the history of the last process call, whatever stream it used */
            const long long n_raw = std::min<long long>(d->run, DEC_HEAD);
            hb::StateParams p{};
            p.in = (d->hist_in_front ? d->d_in : d->d_hist) + (HISTW - (size_t)n_raw);
            p.in_stride = (long long)(d->hist_in_front ? d->in_pitch : HISTW);
            p.out = nullptr;
            p.state = d->d_state;
            p.n_casc = n_raw / (pro ? 4 : 1);
            p.M = M;
            p.round_add = d->variant == SDRD_HB_DB ? 1 : 0;
            p.prologue = pro;
            SDRD_LAUNCH(hb::stateful_kernel, d->S, 1, hb::SNT, hb::stateful_smem_bytes(), d->stream, p);
            d->launches++;
            if (!SDRD_LAUNCH_OK()) return fail_cuda("stage-state kernel launch");
            SDRD_TRY(rt::sync(d->stream), "configure");
        }
        d->consistent = false;
        d->run = 0;
    }
    d->log2_decim = log2_decim;
    d->fcpos = fcpos;
    return 0;
}
extern "C" int sdrd_dec_log2_decim(const sdrd_dec* d) { return d ? d->log2_decim : -1; }
extern "C" long long sdrd_dec_launches(const sdrd_dec* d) { return d ? d->launches : 0; }

extern "C" void* sdrd_dec_dev_input(sdrd_dec* d, size_t* stride)
{
    if (!d) return nullptr;
    if (stride) *stride = d->in_pitch;
    return d->d_in + HISTW;
}
extern "C" void* sdrd_dec_dev_output(sdrd_dec* d, size_t* stride)
{
    if (!d) return nullptr;
    if (stride) *stride = d->out_pitch;
    return d->d_out;
}

extern "C" int sdrd_dec_ipc_export(sdrd_dec* d, void* handle_out, size_t* offset_bytes, size_t* stride)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    if (!handle_out) return fail(SDRD_EINVAL, "null handle buffer");
    SDRD_TRY(rt::ipc_export(d->d_in, handle_out), "cudaIpcGetMemHandle");
    if (offset_bytes) *offset_bytes = HISTW * 4;
    if (stride) *stride = d->in_pitch;
    return 0;
}
extern "C" int sdrd_ipc_open(const void* handle, void** dev_ptr)
{
    if (!handle || !dev_ptr) return fail(SDRD_EINVAL, "null pointer");
    SDRD_TRY(rt::ipc_open(handle, dev_ptr), "cudaIpcOpenMemHandle");
    return 0;
}
extern "C" int sdrd_ipc_close(void* dev_ptr)
{
    if (!dev_ptr) return 0;
    SDRD_TRY(rt::ipc_close(dev_ptr), "cudaIpcCloseMemHandle");
    return 0;
}
extern "C" int sdrd_ipc_copy_rows(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t n_rows,
                                  void* cuda_stream)
{
    if (!dst || !src) return fail(SDRD_EINVAL, "null pointer");
    if (dst_pitch < row_bytes || src_pitch < row_bytes) return fail(SDRD_EINVAL, "pitch smaller than the row");
    SDRD_TRY(rt::copy2d(dst, dst_pitch, src, src_pitch, row_bytes, n_rows, rt::D2D, (rt::stream_t)cuda_stream), "device-to-device copy");
    return 0;
}

/* samples a process call of n_in samples per stream produces under the current configuration */
static size_t dec_out_count(const sdrd_dec* d, size_t n_in)
{
    const int L = d->log2_decim;
    if (L == 0) return n_in;
    if (d->fcpos != SDRD_FC_CENTER && L == 1) return n_in / 2; /* out.resize(len/2), Decimators.cpp:41 */
    return n_in >> L;
}

/* in_off / first / last: a call may be processed in consecutive slices of the input buffer (sdrd_rx_process
 * overlaps the host copies with the kernels that way): slice i > 0 finds its history in place, right in
 * front of it; only the first restores it and only the last saves it.  Every slice but the last must be a
 * multiple of 2^log2_decim samples. */
static int dec_run(sdrd_dec* d, size_t n_in, size_t* n_out_p, unsigned* sample_bits, rt::stream_t st, size_t in_off = 0,
                   bool first = true, bool last = true)
{
    if (in_off + n_in > d->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    const int L = d->log2_decim;
    unsigned ss = sample_bits ? *sample_bits : 16u;
    if (ss < 1 || ss > 16) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    size_t n_out = 0;
    size_t consumed_now = 0;
    const uint32_t* in0 = d->d_in + HISTW + in_off;
    uint32_t* out0 = d->d_out + (in_off >> L);

    /* history of the previous calls in front of the new samples */
    if (first && !d->hist_in_front)
        SDRD_TRY(rt::copy2d(d->d_in, d->in_pitch * 4, d->d_hist, HISTW * 4, HISTW * 4, (size_t)d->S, rt::D2D, st),
                 "restore history");

    if (L == 0) {
        /* Downsampler.cpp:76-80: copy, then decimate1's left-justification for < 16-bit sources */
        n_out = n_in;
        consumed_now = n_in;
        if (n_in) {
            hb::PlainParams p{};
            p.in = in0; p.in_stride = (long long)d->in_pitch;
            p.out = out0; p.out_stride = (long long)d->out_pitch;
            p.n_units = (long long)n_in; p.mode = 0;
            p.norm_shift = ss < 16 ? (int)(16 - ss) : 0;
            const int gx = (int)std::min<size_t>((n_in + 255) / 256, (size_t)d->sms * 8);
            SDRD_LAUNCH(hb::plain_kernel, gx, d->S, 256, 0, st, p);
            d->launches++;
        }
        /* sampleSize is left unchanged by decimate1 (Decimators.cpp:22-35) */
    } else if (d->fcpos != SDRD_FC_CENTER && L <= 2) {
        int norm, trunk;
        unsigned ss_out;
        shift_rule(ss, L, &norm, &trunk, &ss_out);
        hb::PlainParams p{};
        p.in = in0; p.in_stride = (long long)d->in_pitch;
        p.out = out0; p.out_stride = (long long)d->out_pitch;
        p.supra = d->fcpos == SDRD_FC_SUPRA;
        p.norm_shift = norm; p.trunk_shift = trunk;
        const size_t quads = n_in / 4;
        p.n_units = (long long)quads;
        if (L == 1) {
            p.mode = 1;
            n_out = n_in / 2; /* out.resize(len/2), Decimators.cpp:41 */
            p.n_zero_tail = (long long)(n_out - 2 * quads);
        } else {
            p.mode = 2;
            n_out = quads;
        }
        consumed_now = quads * 4;
        if (quads || p.n_zero_tail) {
            const int gx = (int)std::max<size_t>(1, std::min<size_t>((quads + 255) / 256, (size_t)d->sms * 8));
            SDRD_LAUNCH(hb::plain_kernel, gx, d->S, 256, 0, st, p);
            d->launches++;
        }
        ss = ss_out;
    } else {
        const int pro = d->fcpos == SDRD_FC_CENTER ? 0 : (d->fcpos == SDRD_FC_INFRA ? 1 : 2);
        const int M = pro ? L - 2 : L; /* half-band stages */
        n_out = n_in >> L;             /* whole groups only, Decimators.cpp:412 */
        consumed_now = n_out << L;
        int norm, trunk;
        unsigned ss_out;
        shift_rule(ss, L, &norm, &trunk, &ss_out);
        /* After a reconfiguration the first DEC_HEAD raw samples run from the explicit stage states; the head ends
         * on a multiple of 4 outputs so that the warp kernel's vector accesses behind it stay aligned. */
        size_t head = 0;
        if (!d->consistent && n_out) {
            const size_t G = (size_t)4 << L;
            const size_t want = round_up((size_t)(DEC_HEAD - d->run), G);
            head = std::min(consumed_now, want);
            hb::StateParams sp{};
            sp.in = in0; sp.in_stride = (long long)d->in_pitch;
            sp.out = out0; sp.out_stride = (long long)d->out_pitch;
            sp.state = d->d_state;
            sp.n_casc = (long long)(head / (pro ? 4 : 1));
            sp.M = M;
            sp.round_add = d->variant == SDRD_HB_DB ? 1 : 0;
            sp.norm_shift = norm; sp.trunk_shift = trunk;
            sp.prologue = pro;
            SDRD_LAUNCH(hb::stateful_kernel, d->S, 1, hb::SNT, hb::stateful_smem_bytes(), st, sp);
            d->launches++;
        }
        const size_t n_out_fast = n_out - (head >> L);
        if (n_out_fast) {
            hb::Params p{};
            p.in = in0 + head; p.in_stride = (long long)d->in_pitch;
            p.out = out0 + (head >> L); p.out_stride = (long long)d->out_pitch;
            p.n_out = (long long)n_out_fast;
            p.round_add = d->variant == SDRD_HB_DB ? 1 : 0;
            p.norm_shift = norm; p.trunk_shift = trunk;
            p.prologue = pro;
            p.origin = (d->consumed + (long long)head) / (pro ? 4 : 1);
            p.steer_zero = 0; p.steer_one = 1; p.steer_k32 = 32; p.steer_k256 = 256; p.steer_k8192 = 8192;
            /* One warp per share of the global event axis (the streams laid end to end), shares sized for ONE
             * wave of resident warps whatever the number of streams -- equally long, so there is no tail --
             * but never so short that the filter warm-up costs more than ~1/8 of a share (short streams: one
             * warp per stream). */
            const int FN = hb::wfin_n(M);
            const long long ev_stream = ((long long)n_out_fast + FN - 1) / FN;
            const long long ev_total = ev_stream * d->S;
            const long long warm_ev = hb::wwarm_chunks(M) / hb::wmacro(M);
            const size_t smem = hb::wsmem_bytes(M, pro ? 1 : 0);
            long long resident = (long long)((227 * 1024) / (smem + 1024));
            if (resident > SDRD_K1_WARPS_PER_SM) resident = SDRD_K1_WARPS_PER_SM;
            long long want = resident * d->sms;
            if (want < 1) want = 1;
            long long ev_warp = (ev_total + want - 1) / want;
            /* Small calls (the reference's 65536-sample blocks) do not fill one wave: there the length of a share is
             * what the caller waits for (a warp walks its share step by step, ~2.5 us each), so shares shrink until a
             * share is only as long as its own warm-up -- the redundant warm-up runs on SMs that would idle anyway. */
            const long long ev_min = std::min<long long>(std::max<long long>(warm_ev, 1), ev_stream);
            if (ev_warp < ev_min) ev_warp = ev_min;
            const long long n_seg = (ev_total + ev_warp - 1) / ev_warp;
            p.ev_stream = ev_stream;
            p.ev_total = ev_total;
            p.ev_warp = ev_warp;
            p.warm_chunks = hb::wwarm_chunks(M);
            switch (M) {
                case 1: launch_decimate_warp<1>(p, (int)n_seg, st); break;
                case 2: launch_decimate_warp<2>(p, (int)n_seg, st); break;
                case 3: launch_decimate_warp<3>(p, (int)n_seg, st); break;
                case 4: launch_decimate_warp<4>(p, (int)n_seg, st); break;
                case 5: launch_decimate_warp<5>(p, (int)n_seg, st); break;
                default: launch_decimate_warp<6>(p, (int)n_seg, st); break;
            }
            d->launches++;
        }
        ss = ss_out;
    }
    if (!SDRD_LAUNCH_OK()) return fail_cuda("kernel launch");

    /* the last HISTW consumed samples become the next call's history */
    if (last) {
        /* [.. | HISTW + in_off + consumed_now) ends the consumed stream: its last HISTW samples are the history.  When
         * they lie wholly behind the front region they move there in one copy; otherwise through d_hist (source and
         * destination would overlap). */
        const bool direct = in_off + consumed_now >= HISTW;
        SDRD_TRY(rt::copy2d(direct ? d->d_in : d->d_hist, (direct ? d->in_pitch : HISTW) * 4, d->d_in + in_off + consumed_now,
                            d->in_pitch * 4, HISTW * 4, (size_t)d->S, rt::D2D, st),
                 "save history");
        d->hist_in_front = direct;
    } else {
        d->hist_in_front = true; /* the next slice of this call finds its history in place */
    }
    d->consumed += (long long)consumed_now;
    if (d->consumed > (1LL << 50)) d->consumed = 1LL << 50;
    d->run += (long long)consumed_now;
    if (d->run > (1LL << 50)) d->run = 1LL << 50;
    if (d->run >= DEC_HEAD) d->consistent = true;
    if (n_out_p) *n_out_p = n_out;
    if (sample_bits) *sample_bits = ss;
    return 0;
}

extern "C" int sdrd_dec_process_dev(sdrd_dec* d, size_t n_in, size_t* n_out, unsigned* sample_bits, void* cuda_stream)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    return dec_run(d, n_in, n_out, sample_bits, (rt::stream_t)cuda_stream);
}

extern "C" int sdrd_dec_process(sdrd_dec* d, const int16_t* iq_in, size_t n_in, size_t in_stride, int16_t* iq_out,
                                size_t out_stride, size_t* n_out_p, unsigned* sample_bits)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    if ((!iq_in && n_in) || !iq_out) return fail(SDRD_EINVAL, "null sample pointer");
    if (n_in > d->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    if (d->S > 1 && in_stride < n_in) return fail(SDRD_EINVAL, "in_stride smaller than n_in");
    /* everything that can be refused is refused before the filter state moves */
    if (d->S > 1 && out_stride < dec_out_count(d, n_in)) return fail(SDRD_EINVAL, "out_stride smaller than the output length");
    if (sample_bits && (*sample_bits < 1 || *sample_bits > 16)) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    SDRD_TRY(d->stage.to_device(d->d_in + HISTW, d->in_pitch * 4, iq_in, in_stride * 4, n_in * 4, (size_t)d->S, d->stream),
             "copy samples to device");
    size_t n_out = 0;
    if (int rc = dec_run(d, n_in, &n_out, sample_bits, d->stream)) return rc;
    SDRD_TRY(d->stage.to_host(iq_out, out_stride * 4, d->d_out, d->out_pitch * 4, n_out * 4, (size_t)d->S, d->stream),
             "copy samples to host");
    SDRD_TRY(rt::sync(d->stream), "decimate");
    d->stage.finish();
    if (n_out_p) *n_out_p = n_out;
    return 0;
}

extern "C" int sdrd_dec_rescale(sdrd_dec* d, int16_t* iq, size_t n, size_t stride, unsigned* sample_bits)
{
    if (!d) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(d);
    if (!iq && n) return fail(SDRD_EINVAL, "null sample pointer");
    if (n > d->max_in) return fail(SDRD_ERANGE, "n exceeds the max_in given at create time");
    if (d->S > 1 && stride < n) return fail(SDRD_EINVAL, "stride smaller than n");
    const unsigned ss = sample_bits ? *sample_bits : 16u;
    if (ss < 1 || ss > 16) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    if (!n || ss == 16) return 0; /* decimate1 does nothing for 16-bit sources; sampleSize stays as it is */
    /* the staging input buffer is free between calls: the history lives in d_hist */
    rt::stream_t st = d->stream;
    SDRD_TRY(rt::copy2d(d->d_in + HISTW, d->in_pitch * 4, iq, stride * 4, n * 4, (size_t)d->S, rt::H2D, st), "copy samples to device");
    hb::PlainParams p{};
    p.in = d->d_in + HISTW; p.in_stride = (long long)d->in_pitch;
    p.out = d->d_out; p.out_stride = (long long)d->out_pitch;
    p.n_units = (long long)n; p.mode = 0;
    p.norm_shift = (int)(16 - ss);
    const int gx = (int)std::min<size_t>((n + 255) / 256, (size_t)d->sms * 8);
    SDRD_LAUNCH(hb::plain_kernel, gx, d->S, 256, 0, st, p);
    d->launches++;
    if (!SDRD_LAUNCH_OK()) return fail_cuda("kernel launch");
    SDRD_TRY(rt::copy2d(iq, stride * 4, d->d_out, d->out_pitch * 4, n * 4, (size_t)d->S, rt::D2H, st), "copy samples to host");
    SDRD_TRY(rt::sync(st), "rescale");
    return 0;
}

/* ========================================================================================== */
/* interpolator                                                                                */
/* ========================================================================================== */

struct sdrd_int {
    int log2_interp = 0, S = 1;
    int device = -1;             /* the device the handle lives on */
    size_t max_in = 0;
    uint32_t* d_in = nullptr;    /* [S][in_pitch]: hbi::HIST history words, then the new samples */
    uint32_t* d_hist = nullptr;  /* [S][hbi::HIST] */
    uint32_t* d_out = nullptr;   /* [S][out_pitch] */
    size_t in_pitch = 0, out_pitch = 0;
    TileMap out_map;             /* d_out as rows of 32 words: what the TMA tensor store of K4's last stage goes through */
    bool have_map = false;
    /* per-stage state across configure, as in sdrd_dec (Interpolators.h:52-58) */
    int* d_state = nullptr;      /* [S][hbi::ISTATE_WORDS] */
    bool consistent = true;
    long long consumed = 0, run = 0;
    long long launches = 0;
    rt::stream_t stream = 0;
    HostStage stage;
};

static int check_interp(int log2_interp)
{
    if (log2_interp < 0 || log2_interp > 6) return fail(SDRD_EINVAL, "Invalid log2 interpolation factor"); /* Upsampler.cpp:38-42 */
    return 0;
}

extern "C" int sdrd_int_create(sdrd_int** out, int log2_interp, int n_streams, size_t max_in)
{
    if (!out) return fail(SDRD_EINVAL, "null handle pointer");
    *out = nullptr;
    if (int rc = check_interp(log2_interp)) return rc;
    if (n_streams < 1 || max_in < 1) return fail(SDRD_EINVAL, "n_streams and max_in must be positive");
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    sdrd_int* u = new (std::nothrow) sdrd_int();
    if (!u) return fail(SDRD_ENOMEM, "out of host memory");
    u->log2_interp = log2_interp;
    u->S = n_streams;
    u->device = rt::current_device();
    u->max_in = max_in;
    u->in_pitch = hbi::HIST + round_up(max_in, 4) + 4;
    u->out_pitch = round_up(max_in, 4) << 6; /* room for any interp up to 64 (configure may raise it) */
    if (rt::alloc((void**)&u->d_in, u->in_pitch * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&u->d_hist, hbi::HIST * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&u->d_out, u->out_pitch * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&u->d_state, (size_t)hbi::ISTATE_WORDS * 4 * (size_t)n_streams) != 0 || rt::stream_create(&u->stream) != 0) {
        int rc = fail_cuda("allocating interpolator buffers");
        sdrd_int_destroy(u);
        return rc;
    }
    rt::fill(u->d_in, 0, u->in_pitch * 4 * (size_t)n_streams, u->stream);
    /* out_pitch is a multiple of 256 words, so every stream starts on a row of the map */
    u->have_map = rt::make_tile_map(&u->out_map, u->d_out, (unsigned long long)(u->out_pitch * (size_t)n_streams / 32)) == 0;
    if (int rc = sdrd_int_reset(u)) {
        sdrd_int_destroy(u);
        return rc;
    }
    *out = u;
    return 0;
}

extern "C" void sdrd_int_destroy(sdrd_int* u)
{
    if (!u) return;
    SDRD_ON_DEVICE_OF(u);
    rt::sync(u->stream);
    rt::release(u->d_in);
    rt::release(u->d_hist);
    rt::release(u->d_out);
    rt::release(u->d_state);
    rt::stream_destroy(u->stream);
    delete u;
}

extern "C" int sdrd_int_reset(sdrd_int* u)
{
    if (!u) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(u);
    SDRD_TRY(rt::sync_device(), "reset history");
    SDRD_TRY(rt::fill(u->d_hist, 0, hbi::HIST * 4 * (size_t)u->S, u->stream), "reset history");
    SDRD_TRY(rt::fill(u->d_state, 0, (size_t)hbi::ISTATE_WORDS * 4 * (size_t)u->S, u->stream), "reset stage states");
    SDRD_TRY(rt::sync(u->stream), "reset history");
    u->consistent = true;
    u->consumed = u->run = 0;
    return 0;
}

extern "C" int sdrd_int_configure(sdrd_int* u, int log2_interp)
{
    if (!u) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(u);
    if (int rc = check_interp(log2_interp)) return rc;
    if (log2_interp == u->log2_interp) return 0;
    if (u->consumed > 0) {
        /* leaving a configuration: make the state of the stages it ran explicit (see sdrd_dec_configure) */
        if (u->consistent && u->log2_interp > 0 && u->run > 0) {
            SDRD_TRY(rt::sync_device(), "configure");
            const long long n = std::min<long long>(u->run, (long long)hbi::HIST);
            hbi::IStateParams p{};
            p.in = u->d_hist + ((size_t)hbi::HIST - (size_t)n);
            p.in_stride = (long long)hbi::HIST;
            p.out = nullptr;
            p.state = u->d_state;
            p.n_in = n;
            p.log2_interp = u->log2_interp;
            SDRD_LAUNCH(hbi::i_stateful_kernel, u->S, 1, hbi::NT, hbi::i_stateful_smem_bytes(), u->stream, p);
            u->launches++;
            if (!SDRD_LAUNCH_OK()) return fail_cuda("stage-state kernel launch");
            SDRD_TRY(rt::sync(u->stream), "configure");
        }
        u->consistent = false;
        u->run = 0;
    }
    u->log2_interp = log2_interp;
    return 0;
}
extern "C" int sdrd_int_log2_interp(const sdrd_int* u) { return u ? u->log2_interp : -1; }
extern "C" long long sdrd_int_launches(const sdrd_int* u) { return u ? u->launches : 0; }
extern "C" void* sdrd_int_dev_input(sdrd_int* u, size_t* stride)
{
    if (!u) return nullptr;
    if (stride) *stride = u->in_pitch;
    return u->d_in + hbi::HIST;
}
extern "C" void* sdrd_int_dev_output(sdrd_int* u, size_t* stride)
{
    if (!u) return nullptr;
    if (stride) *stride = u->out_pitch;
    return u->d_out;
}

namespace {
template <int NS>
void launch_interpolate(const hbi::Params& p, int S, rt::stream_t st, const TileMap* map, long long out_word0)
{
    /* one wave of resident warps over all streams, each a contiguous range of 64-sample steps of one stream */
    hbi::WarpParams w{};
    w.in = p.in; w.in_stride = p.in_stride; w.out = p.out; w.out_stride = p.out_stride; w.n_in = p.n_in;
    w.log2_interp = p.log2_interp;
    /* rows of the map are 32 words: the call's first output word has to start one (it does unless a reconfiguration
     * head of odd length precedes it) */
    w.use_tma = map != nullptr && (out_word0 & 31) == 0 && (p.out_stride & 31) == 0;
    w.out_word0 = out_word0;
    w.steer_zero = 0; w.steer_one = 1;
    if (map) w.tmap = *map;
    const long long steps = (p.n_in + hbi::WC - 1) / hbi::WC;
    /* warps that are resident at once: the launch bound, or what fits in shared memory (1 KB per CTA is the system's) */
    long long resident = (long long)((227 * 1024) / (hbi::w_smem_bytes(NS) + 1024));
    if (resident > SDRD_K4_WARPS_PER_SM) resident = SDRD_K4_WARPS_PER_SM;
    if (resident > 4) resident &= ~3LL; /* the same number of warps on each of the four schedulers (x32: 8 against 9 warps = 0.391 against 0.414 ms) */
    long long warps = ((long long)rt::sm_count() * resident + S - 1) / S; /* per stream */
    if (warps > steps) warps = steps;
    if (warps < 1) warps = 1;
    w.steps_per_warp = (int)((steps + warps - 1) / warps);
    warps = (steps + w.steps_per_warp - 1) / w.steps_per_warp;
    SDRD_LAUNCH((hbi::interpolate_warp_kernel<NS>), (int)warps, S, 32, hbi::w_smem_bytes(NS), st, w);
}
} /* namespace */

static int int_run(sdrd_int* u, size_t n_in, size_t* n_out_p, rt::stream_t st)
{
    if (n_in > u->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    const int L = u->log2_interp;
    /* history of the previous calls in front of the new samples */
    SDRD_TRY(rt::copy2d(u->d_in, u->in_pitch * 4, u->d_hist, hbi::HIST * 4, hbi::HIST * 4, (size_t)u->S, rt::D2D, st),
             "restore history");
    if (n_in) {
        if (L == 0) { /* samples_out = samples_in, Upsampler.cpp:59-62 */
            SDRD_TRY(rt::copy2d(u->d_out, u->out_pitch * 4, u->d_in + hbi::HIST, u->in_pitch * 4, n_in * 4, (size_t)u->S,
                                rt::D2D, st),
                     "copy samples");
        } else {
            /* after a reconfiguration the first INT_HEAD input samples run from the explicit stage states */
            size_t head = 0;
            if (!u->consistent) {
                head = std::min(n_in, round_up((size_t)(INT_HEAD - u->run), 4));
                hbi::IStateParams sp{};
                sp.in = u->d_in + hbi::HIST;
                sp.in_stride = (long long)u->in_pitch;
                sp.out = u->d_out;
                sp.out_stride = (long long)u->out_pitch;
                sp.state = u->d_state;
                sp.n_in = (long long)head;
                sp.log2_interp = L;
                SDRD_LAUNCH(hbi::i_stateful_kernel, u->S, 1, hbi::NT, hbi::i_stateful_smem_bytes(), st, sp);
                u->launches++;
            }
            if (n_in > head) {
                hbi::Params p{};
                p.in = u->d_in + hbi::HIST + head;
                p.in_stride = (long long)u->in_pitch;
                p.out = u->d_out + (head << L);
                p.out_stride = (long long)u->out_pitch;
                p.n_in = (long long)(n_in - head);
                p.log2_interp = L;
                switch (L < 5 ? L : 5) {
                    case 1: launch_interpolate<1>(p, u->S, st, u->have_map ? &u->out_map : nullptr, (long long)(head << L)); break;
                    case 2: launch_interpolate<2>(p, u->S, st, u->have_map ? &u->out_map : nullptr, (long long)(head << L)); break;
                    case 3: launch_interpolate<3>(p, u->S, st, u->have_map ? &u->out_map : nullptr, (long long)(head << L)); break;
                    case 4: launch_interpolate<4>(p, u->S, st, u->have_map ? &u->out_map : nullptr, (long long)(head << L)); break;
                    default: launch_interpolate<5>(p, u->S, st, u->have_map ? &u->out_map : nullptr, (long long)(head << L)); break;
                }
                u->launches++;
            }
            if (!SDRD_LAUNCH_OK()) return fail_cuda("kernel launch");
        }
    }
    u->consumed += (long long)n_in;
    if (u->consumed > (1LL << 50)) u->consumed = 1LL << 50;
    u->run += (long long)n_in;
    if (u->run > (1LL << 50)) u->run = 1LL << 50;
    if (u->run >= INT_HEAD) u->consistent = true;
    /* the last HIST input samples become the next call's history */
    SDRD_TRY(rt::copy2d(u->d_hist, hbi::HIST * 4, u->d_in + n_in, u->in_pitch * 4, hbi::HIST * 4, (size_t)u->S, rt::D2D, st),
             "save history");
    if (n_out_p) *n_out_p = n_in << L;
    return 0;
}

extern "C" int sdrd_int_process_dev(sdrd_int* u, size_t n_in, size_t* n_out, void* cuda_stream)
{
    if (!u) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(u);
    return int_run(u, n_in, n_out, (rt::stream_t)cuda_stream);
}

extern "C" int sdrd_int_process(sdrd_int* u, const int16_t* iq_in, size_t n_in, size_t in_stride, int16_t* iq_out,
                                size_t out_stride, size_t* n_out_p)
{
    if (!u) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(u);
    if ((!iq_in && n_in) || !iq_out) return fail(SDRD_EINVAL, "null sample pointer");
    if (n_in > u->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    if (u->S > 1 && in_stride < n_in) return fail(SDRD_EINVAL, "in_stride smaller than n_in");
    if (u->S > 1 && out_stride < (n_in << u->log2_interp)) return fail(SDRD_EINVAL, "out_stride smaller than the output length");
    SDRD_TRY(u->stage.to_device(u->d_in + hbi::HIST, u->in_pitch * 4, iq_in, in_stride * 4, n_in * 4, (size_t)u->S, u->stream),
             "copy samples to device");
    size_t n_out = 0;
    if (int rc = int_run(u, n_in, &n_out, u->stream)) return rc;
    SDRD_TRY(u->stage.to_host(iq_out, out_stride * 4, u->d_out, u->out_pitch * 4, n_out * 4, (size_t)u->S, u->stream),
             "copy samples to host");
    SDRD_TRY(rt::sync(u->stream), "interpolate");
    u->stage.finish();
    if (n_out_p) *n_out_p = n_out;
    return 0;
}

/* ========================================================================================== */
/* sender                                                                                      */
/* ========================================================================================== */

struct sdrd_sink {
    int S = 1;
    int device = -1;             /* the device the handle lives on */
    size_t max_samples = 0;
    uint32_t* d_samples = nullptr;  /* host-API staging [S][samples_pitch] */
    size_t samples_pitch = 0;
    uint32_t* d_pending = nullptr;  /* [S][FRAME_SAMPLES] */
    int n_pending = 0;
    uint32_t* d_dgrams = nullptr;
    size_t dgram_words = 0;         /* allocated words */
    size_t frame_cap = 0;           /* frames per stream a call can complete */
    uint32_t center_freq_khz = 0, sample_rate = 0;
    uint8_t sample_bytes = 2, sample_bits = 16;
    int nb_fec = 0;
    bool fixed_time = false;
    bool frame_clock = false;       /* stamp every frame begun in a call with the call's time + its sample offset / rate */
    uint32_t tv_sec = 0, tv_usec = 0;
    uint32_t pending_meta[6] = {0, 0, 0, 0, 0, 0};
    unsigned frame_count = 0;       /* m_frameCount, uint16 wrap */
    long long launches = 0;
    fec::Tables tab;
    rt::stream_t stream = 0;
    /* layout of the last completed call */
    size_t last_frames = 0;
    size_t last_dgram_stride = 0;
    HostStage stage;
};

extern "C" int sdrd_sink_create(sdrd_sink** out, int n_streams, size_t max_samples)
{
    if (!out) return fail(SDRD_EINVAL, "null handle pointer");
    *out = nullptr;
    if (n_streams < 1 || max_samples < 1) return fail(SDRD_EINVAL, "n_streams and max_samples must be positive");
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    sdrd_sink* k = new (std::nothrow) sdrd_sink();
    if (!k) return fail(SDRD_ENOMEM, "out of host memory");
    k->S = n_streams;
    k->max_samples = max_samples;
    k->device = rt::current_device();
    k->samples_pitch = round_up(max_samples, 4);
    k->frame_cap = (max_samples + fec::FRAME_SAMPLES - 1) / fec::FRAME_SAMPLES + 1;
    if (get_tables(&k->tab) != 0) {
        delete k;
        return SDRD_ECUDA;
    }
    if (rt::alloc((void**)&k->d_samples, k->samples_pitch * 4 * (size_t)n_streams) != 0 ||
        rt::alloc((void**)&k->d_pending, (size_t)fec::FRAME_SAMPLES * 4 * (size_t)n_streams) != 0 ||
        rt::stream_create(&k->stream) != 0) {
        int rc = fail_cuda("allocating sink buffers");
        sdrd_sink_destroy(k);
        return rc;
    }
    *out = k;
    return 0;
}

extern "C" void sdrd_sink_destroy(sdrd_sink* k)
{
    if (!k) return;
    SDRD_ON_DEVICE_OF(k);
    rt::sync(k->stream);
    rt::release(k->d_samples);
    rt::release(k->d_pending);
    rt::release(k->d_dgrams);
    rt::stream_destroy(k->stream);
    delete k;
}

extern "C" int sdrd_sink_reset(sdrd_sink* k)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    k->n_pending = 0;
    k->frame_count = 0;
    return 0;
}
extern "C" int sdrd_sink_set_meta(sdrd_sink* k, uint32_t f_khz, uint32_t rate, uint8_t sbytes, uint8_t sbits)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    k->center_freq_khz = f_khz;
    k->sample_rate = rate;
    k->sample_bytes = sbytes;
    k->sample_bits = sbits;
    return 0;
}
extern "C" int sdrd_sink_set_nb_fec(sdrd_sink* k, int nb_fec)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    if (nb_fec < 0 || nb_fec > SDRD_MAX_FEC) return fail(SDRD_EINVAL, "nb_fec must be 0..128 (128 + nb_fec <= 256 blocks)");
    k->nb_fec = nb_fec;
    return 0;
}
extern "C" int sdrd_sink_set_time(sdrd_sink* k, int use_fixed, uint32_t tv_sec, uint32_t tv_usec)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    if (use_fixed < 0 || use_fixed > 3) return fail(SDRD_EINVAL, "time mode must be 0..3");
    k->fixed_time = (use_fixed & 1) != 0;
    k->frame_clock = (use_fixed & 2) != 0;
    k->tv_sec = tv_sec;
    k->tv_usec = tv_usec;
    return 0;
}
extern "C" int sdrd_sink_blocks_per_frame(const sdrd_sink* k) { return k ? 128 + k->nb_fec : 0; }
extern "C" size_t sdrd_sink_frames_for(const sdrd_sink* k, size_t n)
{
    return k ? ((size_t)k->n_pending + n) / fec::FRAME_SAMPLES : 0;
}

/* MetaDataFEC (include/UDPSinkFEC.h:77-99) as six little-endian words */
static void build_meta(const sdrd_sink* k, uint32_t out[6], uint32_t sec, uint32_t usec, unsigned long long offset_samples = 0)
{
    if (offset_samples && k->frame_clock && k->sample_rate) {
        const unsigned long long us = (unsigned long long)usec + offset_samples * 1000000ull / k->sample_rate;
        sec += (uint32_t)(us / 1000000ull);
        usec = (uint32_t)(us % 1000000ull);
    }
    uint8_t m[24];
    auto put = [&](int o, uint32_t v) { m[o] = (uint8_t)v; m[o + 1] = (uint8_t)(v >> 8); m[o + 2] = (uint8_t)(v >> 16); m[o + 3] = (uint8_t)(v >> 24); };
    put(0, k->center_freq_khz);
    put(4, k->sample_rate);
    m[8] = k->sample_bytes;
    m[9] = k->sample_bits;
    m[10] = SDRD_NB_ORIGINAL;
    m[11] = (uint8_t)k->nb_fec;
    put(12, sec);
    put(16, usec);
    put(20, crc32_ieee(m, 20));
    memcpy(out, m, 24);
}

/* samples: device pointer, stream pitch `stride` words */
static int sink_run(sdrd_sink* k, const uint32_t* samples, size_t stride, size_t n, size_t* n_frames_p, rt::stream_t st)
{
    const size_t total = (size_t)k->n_pending + n;
    const size_t n_frames = total / fec::FRAME_SAMPLES;
    const int new_pending = (int)(total % fec::FRAME_SAMPLES);
    if (n_frames > k->frame_cap) return fail(SDRD_ERANGE, "write completes more frames than the handle was sized for");
    const int F = k->nb_fec;
    /* the time of this call: the wall clock (gettimeofday, UDPSinkFEC.cpp:95) or the fixed value */
    uint32_t sec = k->tv_sec, usec = k->tv_usec;
    if (!k->fixed_time) {
        struct timeval tv;
        gettimeofday(&tv, 0);
        sec = (uint32_t)tv.tv_sec;
        usec = (uint32_t)tv.tv_usec;
    }
    uint32_t meta_next[6];
    build_meta(k, meta_next, sec, usec);
    const size_t frame_words = (size_t)(128 + F) * fec::ROW_WORDS;
    const size_t dgram_stride = k->frame_cap * frame_words;
    if (n_frames) {
        const size_t need = dgram_stride * (size_t)k->S;
        if (need > k->dgram_words) {
            rt::sync(st);
            rt::release(k->d_dgrams);
            k->d_dgrams = nullptr;
            k->dgram_words = 0;
            if (rt::alloc((void**)&k->d_dgrams, need * 4) != 0) return fail_cuda("allocating datagram buffer");
            k->dgram_words = need;
        }
        fec::EncParams p{};
        p.mode = 0;
        p.F = F;
        p.cstride = (int)std::max<size_t>(16, round_up((size_t)F, 16));
        p.samples = samples;
        p.sample_stride = (long long)stride;
        p.pending = k->d_pending;
        p.n_pending = k->n_pending;
        memcpy(p.meta_first, k->pending_meta, 24);
        memcpy(p.meta_next, meta_next, 24);
        p.stamp_rate = k->frame_clock ? k->sample_rate : 0u;
        p.frame_index0 = k->frame_count;
        p.dgrams = k->d_dgrams;
        p.dgram_stride = (long long)dgram_stride;
        p.tab = k->tab;
        p.n_frames = (int)n_frames;
        p.n_streams = k->S;
        launch_encode(p, (long long)n_frames * k->S, st);
        k->launches++;
        if (!SDRD_LAUNCH_OK()) return fail_cuda("encode kernel launch");
    }
    /* carry the samples of the unfinished frame (UDPSinkFEC keeps them in m_superBlock / m_txBlocks) */
    if (n_frames == 0) {
        SDRD_TRY(rt::copy2d(k->d_pending + k->n_pending, (size_t)fec::FRAME_SAMPLES * 4, samples, stride * 4, n * 4,
                            (size_t)k->S, rt::D2D, st),
                 "carry samples");
        if (k->n_pending == 0 && n > 0) memcpy(k->pending_meta, meta_next, 24);
    } else {
        const size_t from = n_frames * fec::FRAME_SAMPLES - (size_t)k->n_pending;
        SDRD_TRY(rt::copy2d(k->d_pending, (size_t)fec::FRAME_SAMPLES * 4, samples + from, stride * 4, (size_t)new_pending * 4,
                            (size_t)k->S, rt::D2D, st),
                 "carry samples");
        /* the unfinished frame began `from` samples into this call */
        build_meta(k, k->pending_meta, sec, usec, (unsigned long long)from);
    }
    k->n_pending = new_pending;
    k->frame_count = (k->frame_count + (unsigned)n_frames) & 0xFFFFu;
    k->last_frames = n_frames;
    k->last_dgram_stride = dgram_stride;
    if (n_frames_p) *n_frames_p = n_frames;
    return 0;
}

static int sink_fetch(sdrd_sink* k, uint8_t* datagrams, size_t frame_capacity, size_t n_frames, rt::stream_t st)
{
    if (!n_frames) return 0;
    if (!datagrams) return fail(SDRD_EINVAL, "null datagram buffer");
    if (frame_capacity < n_frames) return fail(SDRD_ERANGE, "frame_capacity smaller than the number of completed frames");
    const size_t frame_bytes = (size_t)(128 + k->nb_fec) * SDRD_UDPSIZE;
    SDRD_TRY(k->stage.to_host(datagrams, frame_capacity * frame_bytes, k->d_dgrams, k->last_dgram_stride * 4, n_frames * frame_bytes,
                              (size_t)k->S, st),
             "copy datagrams to host");
    return 0;
}

extern "C" int sdrd_sink_write(sdrd_sink* k, const int16_t* iq, size_t n, size_t stride, uint8_t* datagrams,
                               size_t frame_capacity, size_t* n_frames_p)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    if (!iq && n) return fail(SDRD_EINVAL, "null sample pointer");
    if (n > k->max_samples) return fail(SDRD_ERANGE, "n_samples exceeds the max_samples given at create time");
    /* refused before the frame counter moves */
    {
        const size_t will_close = sdrd_sink_frames_for(k, n);
        if (will_close > k->frame_cap) return fail(SDRD_ERANGE, "write completes more frames than the handle was sized for");
        if (will_close && !datagrams) return fail(SDRD_EINVAL, "null datagram buffer");
        if (will_close > frame_capacity) return fail(SDRD_ERANGE, "frame_capacity smaller than the number of completed frames");
    }
    SDRD_TRY(k->stage.to_device(k->d_samples, k->samples_pitch * 4, iq, stride * 4, n * 4, (size_t)k->S, k->stream),
             "copy samples to device");
    size_t n_frames = 0;
    if (int rc = sink_run(k, k->d_samples, k->samples_pitch, n, &n_frames, k->stream)) return rc;
    if (int rc = sink_fetch(k, datagrams, frame_capacity, n_frames, k->stream)) return rc;
    SDRD_TRY(rt::sync(k->stream), "sink write");
    k->stage.finish();
    if (n_frames_p) *n_frames_p = n_frames;
    return 0;
}

extern "C" int sdrd_sink_write_dev(sdrd_sink* k, const void* samples, size_t n, size_t stride, size_t* n_frames,
                                   void* cuda_stream)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    if (!samples && n) return fail(SDRD_EINVAL, "null sample pointer");
    return sink_run(k, reinterpret_cast<const uint32_t*>(samples), stride, n, n_frames, (rt::stream_t)cuda_stream);
}
extern "C" void* sdrd_sink_dev_datagrams(sdrd_sink* k, size_t* frame_pitch)
{
    if (!k) return nullptr;
    if (frame_pitch) *frame_pitch = k->frame_cap;
    return k->d_dgrams;
}
extern "C" long long sdrd_sink_launches(const sdrd_sink* k) { return k ? k->launches : 0; }

/* ========================================================================================== */
/* fused rx pipeline                                                                           */
/* ========================================================================================== */

#if defined(__x86_64__) && !defined(SDRD_EMU)
#include <emmintrin.h>
#endif
namespace {
/* Copy into the page-locked accumulation buffer.  The destination is megabytes that the CPU never reads back (the copy
 * engine does): ordinary stores would first fetch every line they overwrite; non-temporal stores do not, which roughly
 * doubles the rate of this copy -- and it is what bounds the queued path. */
void staging_copy(void* dst, const void* src, size_t n)
{
#if defined(__x86_64__) && !defined(SDRD_EMU)
    unsigned char* d = (unsigned char*)dst;
    const unsigned char* s = (const unsigned char*)src;
    const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (n < 256 || head > n) {
        memcpy(d, s, n);
        return;
    }
    memcpy(d, s, head);
    d += head; s += head; n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b = _mm_loadu_si128((const __m128i*)(s + i + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(s + i + 32)), e = _mm_loadu_si128((const __m128i*)(s + i + 48));
        _mm_stream_si128((__m128i*)(d + i), a);
        _mm_stream_si128((__m128i*)(d + i + 16), b);
        _mm_stream_si128((__m128i*)(d + i + 32), c);
        _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence(); /* visible to the copy engine before the transfer is enqueued */
    memcpy(d + i, s + i, n - i);
#else
    memcpy(dst, src, n);
#endif
}

inline void cpu_relax()
{
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}
} /* namespace */

/* Helper threads that share one staging copy with the thread that calls sdrd_rx_submit.  One core moves a block into
 * memory it does not cache at 15 - 20 GB/s, which is what bounds the queued path; the copy engine takes 55 GB/s.  The
 * helpers spin for a short while after a job (a producer that streams blocks finds them awake) and sleep otherwise. */
struct CopyCrew {
    explicit CopyCrew(int n_helpers)
    {
        try {
            for (int i = 0; i < n_helpers; i++) th.emplace_back([this, i] { run(i + 1); });
        } catch (...) {
            dismiss();
            throw;
        }
    }
    ~CopyCrew() { dismiss(); }
    CopyCrew(const CopyCrew&) = delete;
    CopyCrew& operator=(const CopyCrew&) = delete;
    int helpers() const { return (int)th.size(); }

    /* the caller's share is part 0; returns when every part is in place */
    void copy(void* dst_, const void* src_, size_t n_)
    {
        const int parts = (int)th.size() + 1;
        if (n_ < (size_t)parts * 16384) { /* not worth a hand-over */
            staging_copy(dst_, src_, n_);
            return;
        }
        dst = (unsigned char*)dst_;
        src = (const unsigned char*)src_;
        n = n_;
        pending.store((int)th.size(), std::memory_order_relaxed);
        job.fetch_add(1); /* seq_cst: publishes dst / src / n, ordered against `sleepers` below */
        if (sleepers.load() > 0) {
            std::lock_guard<std::mutex> lk(m);
            cv.notify_all();
        }
        part(0, parts);
        while (pending.load(std::memory_order_acquire) != 0) cpu_relax();
    }

private:
    void dismiss()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            quit.store(true);
        }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
    void part(int i, int parts)
    {
        const size_t chunk = ((n / (size_t)parts) + 63) & ~(size_t)63;
        const size_t b = std::min(n, chunk * (size_t)i), e = i + 1 == parts ? n : std::min(n, chunk * (size_t)(i + 1));
        if (e > b) staging_copy(dst + b, src + b, e - b);
    }
    void run(int me)
    {
        unsigned long long seen = 0;
        for (;;) {
            /* wait for the next job: spin first, then sleep */
            bool got = false;
            for (int spin = 0; spin < 20000; spin++) {
                if (job.load(std::memory_order_acquire) != seen || quit.load(std::memory_order_relaxed)) {
                    got = true;
                    break;
                }
                cpu_relax();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(m);
                sleepers.fetch_add(1);
                cv.wait(lk, [&] { return job.load() != seen || quit.load(); });
                sleepers.fetch_sub(1);
            }
            if (quit.load()) return;
            seen = job.load(std::memory_order_acquire);
            part(me, (int)th.size() + 1);
            pending.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv;
    std::atomic<unsigned long long> job{0};
    std::atomic<int> pending{0}, sleepers{0};
    std::atomic<bool> quit{false};
    unsigned char* dst = nullptr;
    const unsigned char* src = nullptr;
    size_t n = 0;
};

struct sdrd_rx {
    int device = -1;             /* the device the handle lives on */
    sdrd_dec* dec = nullptr;
    sdrd_sink* sink = nullptr;
    rt::stream_t copy_stream = 0;   /* host -> device copies of sdrd_rx_process, ahead of the kernels */
    rt::event_t copied[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t slice_threshold = (size_t)32 << 20; /* calls of at least this many input bytes go through in 8 slices */

    /* ---- queued form (sdrd_rx_submit / sdrd_rx_collect) ----
     * submit copies the caller's block into a page-locked accumulation buffer and returns; whenever the device is
     * idle everything accumulated so far goes out as ONE chain of copy -> decimate -> frame + encode -> copy back
     * (no synchronisation in submit), so small blocks are batched exactly as far as the device lags behind the
     * producer.  Completed frames wait in `ready` until collect hands them over. */
    struct Batch {
        int blocks_per_frame = 0;
        size_t n_frames = 0, taken = 0; /* per stream; frames already handed over */
        std::vector<uint8_t> data;      /* [S][n_frames][blocks_per_frame * 512] */
    };
    struct Flight {                          /* a chain on the device */
        int slot = 0;                        /* which q_out / q_done it owns */
        size_t samples = 0;                  /* per stream */
        size_t frames = 0, out_frames = 0;   /* frames completed per stream; frame pitch of q_out[slot] */
        int bpf = 0;
    };
    std::mutex q_mutex;
    uint32_t* q_in[3] = {nullptr, nullptr, nullptr}; /* page-locked [S][q_cap]: two chains in flight + the one being filled */
    uint8_t* q_out[2] = {nullptr, nullptr};  /* page-locked [S][out_frames][blocks per frame * 512] per slot, grown on demand */
    size_t q_out_bytes[2] = {0, 0};
    size_t q_cap = 0;
    size_t q_fill = 0;                      /* samples per stream accumulated in q_in[q_cur] */
    size_t q_min_chain = 0;                 /* submit starts a chain only once this many samples per stream have accumulated */
    int q_cur = 0;
    unsigned q_ss = 16;                     /* sample bits of the accumulated samples */
    std::deque<Flight> q_fly;               /* at most two, oldest first */
    rt::event_t q_done[2] = {0, 0};
    rt::event_t q_copied = 0;
    rt::stream_t q_stream = 0;
    std::deque<Batch> ready;
    size_t ready_frames = 0;                /* per stream */
    long long q_launches = 0;               /* chains sent so far (a measure of the batching achieved) */
    CopyCrew* crew = nullptr;               /* helper threads that share submit's staging copy (sdrd_rx_set_staging_threads) */
};

extern "C" int sdrd_rx_create(sdrd_rx** out, int log2_decim, int fcpos, int variant, int n_streams, size_t max_in)
{
    if (!out) return fail(SDRD_EINVAL, "null handle pointer");
    *out = nullptr;
    sdrd_rx* r = new (std::nothrow) sdrd_rx();
    if (!r) return fail(SDRD_ENOMEM, "out of host memory");
    r->device = rt::current_device();
    int rc = sdrd_dec_create(&r->dec, log2_decim, fcpos, variant, n_streams, max_in);
    if (!rc) rc = sdrd_sink_create(&r->sink, n_streams, max_in);
    if (!rc && rt::stream_create(&r->copy_stream) != 0) rc = fail_cuda("creating the copy stream");
    for (int i = 0; i < 8 && !rc; i++)
        if (rt::event_create(&r->copied[i]) != 0) rc = fail_cuda("creating events");
    if (rc) {
        sdrd_rx_destroy(r);
        return rc;
    }
    *out = r;
    return 0;
}
extern "C" void sdrd_rx_destroy(sdrd_rx* r)
{
    if (!r) return;
    SDRD_ON_DEVICE_OF(r);
    if (r->copy_stream) rt::sync(r->copy_stream);
    if (r->q_stream) rt::sync(r->q_stream);
    delete r->crew;
    for (int i = 0; i < 3; i++) rt::host_release(r->q_in[i]);
    for (int i = 0; i < 2; i++) {
        rt::host_release(r->q_out[i]);
        rt::event_destroy(r->q_done[i]);
    }
    rt::event_destroy(r->q_copied);
    rt::stream_destroy(r->q_stream);
    sdrd_dec_destroy(r->dec);
    sdrd_sink_destroy(r->sink);
    for (int i = 0; i < 8; i++) rt::event_destroy(r->copied[i]);
    rt::stream_destroy(r->copy_stream);
    delete r;
}
extern "C" int sdrd_rx_reset(sdrd_rx* r)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    {
        std::lock_guard<std::mutex> lk(r->q_mutex);
        for (const sdrd_rx::Flight& f : r->q_fly) rt::event_sync(r->q_done[f.slot]);
        r->q_fly.clear();
        r->q_fill = 0;
        r->ready.clear();
        r->ready_frames = 0;
    }
    if (int rc = sdrd_dec_reset(r->dec)) return rc;
    return sdrd_sink_reset(r->sink);
}
extern "C" int sdrd_rx_set_slice_bytes(sdrd_rx* r, size_t min_call_bytes)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    r->slice_threshold = min_call_bytes ? min_call_bytes : ((size_t)32 << 20);
    return 0;
}
extern "C" sdrd_dec* sdrd_rx_dec(sdrd_rx* r) { return r ? r->dec : nullptr; }
extern "C" sdrd_sink* sdrd_rx_sink(sdrd_rx* r) { return r ? r->sink : nullptr; }
extern "C" long long sdrd_rx_launches(const sdrd_rx* r) { return r ? r->dec->launches + r->sink->launches : 0; }
extern "C" void* sdrd_rx_dev_datagrams(sdrd_rx* r, size_t* frame_pitch)
{
    if (!r) return nullptr;
    if (frame_pitch) *frame_pitch = r->sink->frame_cap;
    return r->sink->d_dgrams;
}

static int q_drain(sdrd_rx* r);

/* sdrdaemonrx.cpp:618-643: the sink's sample size follows the decimator -- with decim = 0 it stays the source's
 * (rescale left-justifies the samples but the meta data keep get_sample_bits()), otherwise it is what process
 * returned; sample bytes = (bits - 1) / 8 + 1 */
static void rx_set_sample_size(sdrd_rx* r, unsigned ss_in, unsigned ss_out)
{
    const unsigned bits = r->dec->log2_decim == 0 ? ss_in : ss_out;
    r->sink->sample_bits = (uint8_t)bits;
    r->sink->sample_bytes = (uint8_t)((bits - 1) / 8 + 1);
}

extern "C" int sdrd_rx_process_dev(sdrd_rx* r, size_t n_in, size_t* n_frames, unsigned* sample_bits, void* cuda_stream)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    rt::stream_t st = (rt::stream_t)cuda_stream;
    const unsigned ss_in = sample_bits ? *sample_bits : 16u;
    if (ss_in < 1 || ss_in > 16) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    if (n_in > r->dec->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    if (int rc = q_drain(r)) return rc;
    size_t n_out = 0;
    unsigned ss = ss_in;
    if (int rc = dec_run(r->dec, n_in, &n_out, &ss, st)) return rc;
    rx_set_sample_size(r, ss_in, ss);
    if (sample_bits) *sample_bits = ss;
    return sink_run(r->sink, r->dec->d_out, r->dec->out_pitch, n_out, n_frames, st);
}

extern "C" int sdrd_rx_process(sdrd_rx* r, const int16_t* iq_in, size_t n_in, size_t in_stride, uint8_t* datagrams,
                               size_t frame_capacity, size_t* n_frames_p, unsigned* sample_bits)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    if (!iq_in && n_in) return fail(SDRD_EINVAL, "null sample pointer");
    sdrd_dec* d = r->dec;
    if (n_in > d->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    if (d->S > 1 && in_stride < n_in) return fail(SDRD_EINVAL, "in_stride smaller than n_in");
    const unsigned ss_in = sample_bits ? *sample_bits : 16u;
    if (ss_in < 1 || ss_in > 16) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    /* everything that can be refused is refused before the decimator and sink state move */
    {
        const size_t will_close = sdrd_sink_frames_for(r->sink, dec_out_count(d, n_in));
        if (will_close > frame_capacity) return fail(SDRD_ERANGE, "frame_capacity smaller than the number of completed frames");
        if (will_close > r->sink->frame_cap) return fail(SDRD_ERANGE, "write completes more frames than the handle was sized for");
        if (will_close && !datagrams) return fail(SDRD_EINVAL, "null datagram buffer");
    }
    if (int rc = q_drain(r)) return rc;
    rt::stream_t st = d->stream;
    /* Large calls go through in slices: the copy of slice i + 1 (copy stream) runs while slice i is being
     * decimated, framed, encoded and its datagrams copied back (compute stream), so the call costs little
     * more than its host -> device copy. */
    const int L = d->log2_decim;
    int n_slices = 1;
    /* SDRD_RX_SLICE_BYTES: tests lower the threshold to exercise the sliced path on small inputs */
    if ((size_t)d->S * n_in * 4 >= r->slice_threshold) n_slices = 8;
    size_t slice = n_in / (size_t)n_slices;
    slice -= slice % (((size_t)1 << L) * 4); /* whole decimation groups, 16-byte aligned */
    if (slice == 0) n_slices = 1;
    size_t n_frames = 0;
    unsigned ss_out = ss_in;
    const size_t frame_bytes = (size_t)(128 + r->sink->nb_fec) * SDRD_UDPSIZE;
    for (int i = 0; i < n_slices; i++) {
        const size_t off = (size_t)i * slice;
        const size_t len = i == n_slices - 1 ? n_in - off : slice;
        if (n_slices > 1) {
            SDRD_TRY(rt::copy2d(d->d_in + HISTW + off, d->in_pitch * 4, iq_in + 2 * off, in_stride * 4, len * 4, (size_t)d->S,
                                rt::H2D, r->copy_stream),
                     "copy samples to device");
            SDRD_TRY(rt::event_record(r->copied[i], r->copy_stream), "record copy event");
        } else {
            SDRD_TRY(d->stage.to_device(d->d_in + HISTW, d->in_pitch * 4, iq_in, in_stride * 4, n_in * 4, (size_t)d->S, st),
                     "copy samples to device");
        }
    }
    for (int i = 0; i < n_slices; i++) {
        const size_t off = (size_t)i * slice;
        const size_t len = i == n_slices - 1 ? n_in - off : slice;
        if (n_slices > 1) SDRD_TRY(rt::stream_wait(st, r->copied[i]), "wait for copy");
        size_t n_out = 0, nf = 0;
        unsigned ss = ss_in;
        if (int rc = dec_run(d, len, &n_out, &ss, st, off, i == 0, i == n_slices - 1)) return rc;
        rx_set_sample_size(r, ss_in, ss);
        ss_out = ss;
        if (int rc = sink_run(r->sink, d->d_out + (off >> L), d->out_pitch, n_out, &nf, st)) return rc;
        if (nf) {
            if (n_frames + nf > frame_capacity) return fail(SDRD_ERANGE, "frame_capacity smaller than the number of completed frames");
            if (n_slices == 1)
                SDRD_TRY(d->stage.to_host(datagrams + n_frames * frame_bytes, frame_capacity * frame_bytes, r->sink->d_dgrams,
                                          r->sink->last_dgram_stride * 4, nf * frame_bytes, (size_t)d->S, st),
                         "copy datagrams to host");
            else
                SDRD_TRY(rt::copy2d(datagrams + n_frames * frame_bytes, frame_capacity * frame_bytes, r->sink->d_dgrams,
                                    r->sink->last_dgram_stride * 4, nf * frame_bytes, (size_t)d->S, rt::D2H, st),
                         "copy datagrams to host");
        }
        n_frames += nf;
    }
    SDRD_TRY(rt::sync(st), "rx process");
    d->stage.finish();
    if (n_frames_p) *n_frames_p = n_frames;
    if (sample_bits) *sample_bits = ss_out;
    return 0;
}

/* ---- queued form ---- */

namespace {
/* under q_mutex: allocate the staging buffers on first use */
int q_prepare(sdrd_rx* r)
{
    if (r->q_in[0]) return 0;
    sdrd_dec* d = r->dec;
    r->q_cap = d->max_in;
    const size_t in_bytes = (size_t)d->S * r->q_cap * 4;
    bool ok = rt::alloc((void**)&d->d_in_alt, d->in_pitch * 4 * (size_t)d->S) == 0;
    for (int i = 0; i < 3 && ok; i++) ok = rt::host_alloc((void**)&r->q_in[i], in_bytes) == 0;
    for (int i = 0; i < 2 && ok; i++) ok = rt::event_create(&r->q_done[i]) == 0;
    ok = ok && rt::event_create(&r->q_copied) == 0 && rt::stream_create(&r->q_stream) == 0;
    /* (bytes past the valid samples are read by a launch's last chunk, results discarded: keep them defined) */
    ok = ok && rt::fill(d->d_in_alt, 0, d->in_pitch * 4 * (size_t)d->S, r->q_stream) == 0 && rt::sync(r->q_stream) == 0;
    if (!ok) {
        for (int i = 0; i < 3; i++) {
            rt::host_release(r->q_in[i]);
            r->q_in[i] = nullptr;
        }
        rt::release(d->d_in_alt);
        d->d_in_alt = nullptr;
        return fail_cuda("allocating the buffers of the queued path");
    }
    return 0;
}

/* under q_mutex: the oldest chain in flight has completed -> its frames join `ready` */
void q_harvest(sdrd_rx* r)
{
    if (r->q_fly.empty()) return;
    const sdrd_rx::Flight f = r->q_fly.front();
    r->q_fly.pop_front();
    if (!f.frames) return;
    sdrd_rx::Batch b;
    b.blocks_per_frame = f.bpf;
    b.n_frames = f.frames;
    const size_t per_stream = b.n_frames * (size_t)b.blocks_per_frame * SDRD_UDPSIZE;
    b.data.resize((size_t)r->dec->S * per_stream);
    for (int s = 0; s < r->dec->S; s++)
        memcpy(&b.data[(size_t)s * per_stream], r->q_out[f.slot] + (size_t)s * f.out_frames * (size_t)b.blocks_per_frame * SDRD_UDPSIZE,
               per_stream);
    r->ready_frames += b.n_frames;
    r->ready.push_back(std::move(b));
}
/* under q_mutex: harvest every chain that has completed (asks the driver once per chain in flight) */
void q_poll(sdrd_rx* r)
{
    while (!r->q_fly.empty() && rt::event_done(r->q_done[r->q_fly.front().slot]) != 0) q_harvest(r);
}
/* under q_mutex: wait for the oldest chain in flight */
int q_wait_oldest(sdrd_rx* r)
{
    if (r->q_fly.empty()) return 0;
    SDRD_TRY(rt::event_sync(r->q_done[r->q_fly.front().slot]), "waiting for the chain in flight");
    q_harvest(r);
    return 0;
}

/* under q_mutex, at most one chain in flight: send everything accumulated as one chain, no synchronisation.
 *
 * Two chains overlap on the device: the samples of chain k go into the decimator's ALTERNATE input buffer on the copy
 * stream while chain k - 1 (copy done, kernels running on q_stream) still reads the current one; then, in q_stream order,
 * the history chain k - 1 left at the front of its buffer is carried over and the two buffers swap roles.  Chain k - 2,
 * the last user of the alternate buffer, has completed before chain k is sent (at most two in flight).  Everything
 * behind the copy -- decimator state, the sink's pending samples, the datagram images -- is single and ordered by
 * q_stream.  (Cutting ONE chain into pieces instead was measured and gains nothing: behind its copy a piece costs
 * ~60 us of dependent launches whatever its size.) */
int q_launch(sdrd_rx* r)
{
    sdrd_dec* d = r->dec;
    const size_t n = r->q_fill;
    if (!n) return 0;
    if (r->q_fly.size() >= 2) return fail(SDRD_ECUDA, "internal: two chains already in flight");
    rt::stream_t st = r->q_stream;
    sdrd_rx::Flight f;
    f.slot = r->q_fly.empty() ? 0 : 1 - r->q_fly.back().slot;
    f.samples = n;
    /* room for the frames this chain will complete.  (will_close counts from the sink's state, which already includes
     * the chain in flight: sink_run advances it at launch time.) */
    {
        const size_t will_close = sdrd_sink_frames_for(r->sink, dec_out_count(d, n));
        const size_t need = (size_t)d->S * will_close * (size_t)(128 + r->sink->nb_fec) * SDRD_UDPSIZE;
        if (need > r->q_out_bytes[f.slot]) {
            rt::host_release(r->q_out[f.slot]);
            r->q_out[f.slot] = nullptr;
            r->q_out_bytes[f.slot] = 0;
            if (rt::host_alloc((void**)&r->q_out[f.slot], need) != 0) return fail_cuda("allocating the page-locked output buffer");
            r->q_out_bytes[f.slot] = need;
        }
        f.out_frames = will_close;
    }
    SDRD_TRY(rt::copy2d(d->d_in_alt + HISTW, d->in_pitch * 4, r->q_in[r->q_cur], r->q_cap * 4, n * 4, (size_t)d->S, rt::H2D, r->copy_stream),
             "copy samples to device");
    SDRD_TRY(rt::event_record(r->q_copied, r->copy_stream), "record copy event");
    SDRD_TRY(rt::stream_wait(st, r->q_copied), "wait for copy");
    if (d->hist_in_front)
        SDRD_TRY(rt::copy2d(d->d_in_alt, d->in_pitch * 4, d->d_in, d->in_pitch * 4, HISTW * 4, (size_t)d->S, rt::D2D, st), "carry the history over");
    std::swap(d->d_in, d->d_in_alt);
    size_t n_out = 0, nf = 0;
    unsigned ss = r->q_ss;
    if (int rc = dec_run(d, n, &n_out, &ss, st)) return rc;
    rx_set_sample_size(r, r->q_ss, ss);
    if (int rc = sink_run(r->sink, d->d_out, d->out_pitch, n_out, &nf, st)) return rc;
    f.bpf = 128 + r->sink->nb_fec;
    if (nf) {
        const size_t frame_bytes = (size_t)f.bpf * SDRD_UDPSIZE;
        SDRD_TRY(rt::copy2d(r->q_out[f.slot], f.out_frames * frame_bytes, r->sink->d_dgrams, r->sink->last_dgram_stride * 4, nf * frame_bytes,
                            (size_t)d->S, rt::D2H, st),
                 "copy datagrams to host");
    }
    SDRD_TRY(rt::event_record(r->q_done[f.slot], st), "record completion");
    f.frames = nf;
    r->q_fly.push_back(f);
    r->q_cur = (r->q_cur + 1) % 3;
    r->q_fill = 0;
    r->q_launches++;
    return 0;
}

/* under q_mutex: may what has accumulated go now?  An idle device takes it at once (latency).  Sending a chain costs
 * the submitting thread ~25 us of driver calls whatever its length, so next to a chain in flight a second one goes
 * only when that cost is small against its copy -- at least 1 MB of samples -- and when it is at least as long as the
 * one in flight: a device that lags gets longer chains, not more of them (measured with the greedy rule "whenever
 * fewer than two are in flight": 1.3 blocks per chain and 27 us per 65536-sample block from a single thread). */
bool q_may_launch(const sdrd_rx* r, size_t fill)
{
    if (!fill || fill < r->q_min_chain || r->q_fly.size() >= 2) return false;
    if (r->q_fly.empty()) return true;
    return fill >= r->q_fly.back().samples && (size_t)r->dec->S * fill * 4 >= ((size_t)1 << 20);
}
/* under q_mutex: everything submitted so far through the device and into `ready` */
int q_flush(sdrd_rx* r)
{
    while (!r->q_fly.empty() || r->q_fill) {
        if (r->q_fill && r->q_fly.size() < 2) {
            if (int rc = q_launch(r)) return rc;
        } else if (int rc = q_wait_oldest(r)) {
            return rc;
        }
    }
    return 0;
}
} /* namespace */

/* the synchronous entry points on a handle that also has queued work: that work comes first */
static int q_drain(sdrd_rx* r)
{
    std::unique_lock<std::mutex> lk(r->q_mutex);
    if (!r->q_in[0]) return 0;
    return q_flush(r);
}

extern "C" int sdrd_rx_submit(sdrd_rx* r, const int16_t* iq_in, size_t n_in, size_t in_stride, unsigned* sample_bits)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    sdrd_dec* d = r->dec;
    if (!iq_in && n_in) return fail(SDRD_EINVAL, "null sample pointer");
    if (n_in > d->max_in) return fail(SDRD_ERANGE, "n_in exceeds the max_in given at create time");
    if (d->S > 1 && in_stride < n_in) return fail(SDRD_EINVAL, "in_stride smaller than n_in");
    const unsigned ss_in = sample_bits ? *sample_bits : 16u;
    if (ss_in < 1 || ss_in > 16) return fail(SDRD_EINVAL, "sample_bits must be 1..16");
    /* blocks of one chain are decimated as one piece of stream: whole groups of 2^decim samples per block, as the
     * reference's own loop bound assumes of every vector (Decimators.cpp:412) */
    if (n_in & (((size_t)1 << d->log2_decim) - 1)) return fail(SDRD_EINVAL, "queued blocks must be a multiple of 2^decim samples");
    std::unique_lock<std::mutex> lk(r->q_mutex);
    if (int rc = q_prepare(r)) return rc;
    if (r->ready_frames > 65536) return fail(SDRD_ERANGE, "too many completed frames waiting: call sdrd_rx_collect");
    /* a block that does not fit behind what has accumulated, or that has another sample size, starts a new chain */
    if (r->q_fill && (r->q_fill + n_in > r->q_cap || ss_in != r->q_ss)) {
        if (r->q_fly.size() >= 2)
            if (int rc = q_wait_oldest(r)) return rc;
        if (int rc = q_launch(r)) return rc;
    }
    r->q_ss = ss_in;
    for (int s = 0; s < d->S; s++) {
        void* to = r->q_in[r->q_cur] + (size_t)s * r->q_cap + r->q_fill;
        const void* from = iq_in + 2 * (size_t)s * in_stride;
        if (r->crew) r->crew->copy(to, from, n_in * 4);
        else staging_copy(to, from, n_in * 4);
    }
    r->q_fill += n_in;
    /* asking the driver whether a chain has completed costs microseconds: only when the answer matters, i.e. when
     * what has accumulated is long enough to go but the chains in flight, as last seen, stand in its way */
    if (r->q_fill >= r->q_min_chain && !r->q_fly.empty() && !q_may_launch(r, r->q_fill)) q_poll(r);
    if (q_may_launch(r, r->q_fill))
        if (int rc = q_launch(r)) return rc;
    if (sample_bits) { /* what the decimator makes of ss_in (Decimators.cpp:408-409,515): known without running it */
        int norm, trunk;
        unsigned ss_out = ss_in;
        if (d->log2_decim > 0) shift_rule(ss_in, d->log2_decim, &norm, &trunk, &ss_out);
        *sample_bits = ss_out;
    }
    return 0;
}

extern "C" int sdrd_rx_collect(sdrd_rx* r, uint8_t* datagrams, size_t frame_capacity, size_t* n_frames_p, int* blocks_per_frame,
                               int wait)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(r);
    if (n_frames_p) *n_frames_p = 0;
    std::unique_lock<std::mutex> lk(r->q_mutex);
    if (wait) {
        /* everything submitted so far: the chains in flight, then what has accumulated behind them */
        if (int rc = q_flush(r)) return rc;
    } else if (r->ready.empty() && !r->q_fly.empty()) {
        /* (frames already waiting are handed over without asking the driver about the chains in flight) */
        q_poll(r);
        if (q_may_launch(r, r->q_fill))
            if (int rc = q_launch(r)) return rc;
    }
    if (r->ready.empty()) return 0;
    if (!datagrams && frame_capacity) return fail(SDRD_EINVAL, "null datagram buffer");
    /* hand over completed frames, oldest first, while they share one frame size */
    const int bpf = r->ready.front().blocks_per_frame;
    const size_t frame_bytes = (size_t)bpf * SDRD_UDPSIZE;
    size_t got = 0;
    while (!r->ready.empty() && got < frame_capacity && r->ready.front().blocks_per_frame == bpf) {
        sdrd_rx::Batch& b = r->ready.front();
        const size_t take = std::min(frame_capacity - got, b.n_frames - b.taken);
        for (int s = 0; s < r->dec->S; s++)
            memcpy(datagrams + ((size_t)s * frame_capacity + got) * frame_bytes,
                   &b.data[((size_t)s * b.n_frames + b.taken) * frame_bytes], take * frame_bytes);
        b.taken += take;
        got += take;
        r->ready_frames -= take;
        if (b.taken == b.n_frames) r->ready.pop_front();
    }
    if (n_frames_p) *n_frames_p = got;
    if (blocks_per_frame) *blocks_per_frame = bpf;
    return 0;
}
extern "C" long long sdrd_rx_chains(const sdrd_rx* r) { return r ? r->q_launches : 0; }
extern "C" int sdrd_rx_set_min_chain(sdrd_rx* r, size_t min_samples)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    if (min_samples > r->dec->max_in) return fail(SDRD_ERANGE, "min_samples exceeds the max_in given at create time");
    std::lock_guard<std::mutex> lk(r->q_mutex);
    r->q_min_chain = min_samples;
    return 0;
}
extern "C" int sdrd_rx_set_staging_threads(sdrd_rx* r, int n_helpers)
{
    if (!r) return fail(SDRD_EINVAL, "null handle");
    if (n_helpers < 0 || n_helpers > 15) return fail(SDRD_ERANGE, "0..15 helper threads");
    std::lock_guard<std::mutex> lk(r->q_mutex);
    if ((r->crew ? r->crew->helpers() : 0) == n_helpers) return 0;
    delete r->crew;
    r->crew = nullptr;
    if (n_helpers) {
        try {
            r->crew = new CopyCrew(n_helpers);
        } catch (...) {
            return fail(SDRD_ENOMEM, "could not start the staging threads");
        }
    }
    return 0;
}

/* ========================================================================================== */
/* stateless CM256 encode / frame decode                                                       */
/* ========================================================================================== */

namespace {
/* grow-only device scratch for the host-pointer entry points */
struct Scratch {
    void* p = nullptr;
    size_t n = 0;
    int device = -1;
    ~Scratch() { drop(); } /* thread exit */
    void drop()
    {
        if (p) {
            DeviceGuard g(device);
            rt::release(p);
        }
        p = nullptr;
        n = 0;
    }
    int ensure(size_t need)
    {
        if (device != rt::current_device()) { /* the thread moved to another device: start over there */
            drop();
            device = rt::current_device();
        }
        if (need <= n) return 0;
        rt::release(p);
        p = nullptr;
        n = 0;
        if (rt::alloc(&p, need) != 0) return -1;
        n = need;
        return 0;
    }
};
thread_local Scratch g_scr_a, g_scr_b, g_scr_c;
/* copy-in / kernel / copy-out streams of the host-pointer decode */
struct DecPipe {
    static constexpr int NS = 16;
    rt::stream_t s_in = 0, s_k = 0, s_out = 0;
    rt::event_t ev_in[NS] = {}, ev_k[NS] = {};
    bool ready = false;
    int device = -1;
    ~DecPipe() { drop(); } /* thread exit */
    void drop()
    {
        if (device < 0) return;
        DeviceGuard g(device);
        rt::stream_destroy(s_in);
        rt::stream_destroy(s_k);
        rt::stream_destroy(s_out);
        for (int i = 0; i < NS; i++) {
            rt::event_destroy(ev_in[i]);
            rt::event_destroy(ev_k[i]);
            ev_in[i] = ev_k[i] = 0;
        }
        s_in = s_k = s_out = 0;
        ready = false;
        device = -1;
    }
    int init()
    {
        if (ready && device == rt::current_device()) return 0;
        drop(); /* first use, or the thread moved to another device: streams and events belong to a device */
        device = rt::current_device();
        if (rt::stream_create(&s_in) || rt::stream_create(&s_k) || rt::stream_create(&s_out)) return -1;
        for (int i = 0; i < NS; i++)
            if (rt::event_create(&ev_in[i]) || rt::event_create(&ev_k[i])) return -1;
        ready = true;
        return 0;
    }
};
thread_local DecPipe g_dec_pipe;
} /* namespace */

extern "C" int sdrd_cm256_encode_dev(const uint8_t* originals, size_t block_pitch, int n_frames, int recovery_count,
                                     uint8_t* recovery, void* cuda_stream)
{
    /* cm256_encode's parameter checks */
    if (recovery_count <= 0 || recovery_count > SDRD_MAX_FEC) return fail(SDRD_EINVAL, "recovery_count must be 1..128");
    if (n_frames < 0) return fail(SDRD_EINVAL, "negative frame count");
    if (!originals || !recovery) return fail(SDRD_EINVAL, "null block pointer");
    if (block_pitch < SDRD_BLOCK_BYTES || (block_pitch & 3)) return fail(SDRD_EINVAL, "block_pitch must be >= 508 and a multiple of 4");
    if (n_frames == 0) return 0;
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    fec::EncParams p{};
    if (int rc = get_tables(&p.tab)) return rc;
    p.mode = 1;
    p.F = recovery_count;
    p.cstride = (int)round_up((size_t)recovery_count, 16);
    p.originals = originals;
    p.block_pitch = (long long)block_pitch;
    p.recovery = recovery;
    rt::stream_t st = (rt::stream_t)cuda_stream;
    p.n_frames = n_frames;
    p.n_streams = 1;
    launch_encode(p, (long long)n_frames, st);
    if (!SDRD_LAUNCH_OK()) return fail_cuda("encode kernel launch");
    return 0;
}

extern "C" int sdrd_cm256_encode(const uint8_t* originals, size_t block_pitch, int n_frames, int recovery_count,
                                 uint8_t* recovery)
{
    if (recovery_count <= 0 || recovery_count > SDRD_MAX_FEC) return fail(SDRD_EINVAL, "recovery_count must be 1..128");
    if (n_frames < 0) return fail(SDRD_EINVAL, "negative frame count");
    if (!originals || !recovery) return fail(SDRD_EINVAL, "null block pointer");
    if (block_pitch < SDRD_BLOCK_BYTES || (block_pitch & 3)) return fail(SDRD_EINVAL, "block_pitch must be >= 508 and a multiple of 4");
    if (n_frames == 0) return 0;
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    const size_t in_bytes = (size_t)n_frames * 128 * block_pitch;
    const size_t out_bytes = (size_t)n_frames * (size_t)recovery_count * SDRD_BLOCK_BYTES;
    if (g_scr_a.ensure(in_bytes) || g_scr_b.ensure(out_bytes)) return fail_cuda("allocating scratch");
    SDRD_TRY(rt::copy(g_scr_a.p, originals, in_bytes, rt::H2D, 0), "copy originals to device");
    if (int rc = sdrd_cm256_encode_dev((const uint8_t*)g_scr_a.p, block_pitch, n_frames, recovery_count, (uint8_t*)g_scr_b.p, 0))
        return rc;
    SDRD_TRY(rt::copy(recovery, g_scr_b.p, out_bytes, rt::D2H, 0), "copy recovery blocks to host");
    SDRD_TRY(rt::sync(0), "cm256 encode");
    return 0;
}

extern "C" int sdrd_fec_decode_dev(const uint8_t* superblocks, size_t blocks_pitch, const int* n_blocks, int n_frames,
                                   uint8_t* payload, uint8_t* block0, int* status, void* cuda_stream)
{
    if (n_frames < 0) return fail(SDRD_EINVAL, "negative frame count");
    if (!superblocks || !n_blocks || !payload || !status) return fail(SDRD_EINVAL, "null pointer");
    if (blocks_pitch < 1) return fail(SDRD_EINVAL, "blocks_pitch must be positive");
    if (n_frames == 0) return 0;
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    fec::DecParams p{};
    if (int rc = get_tables(&p.tab)) return rc;
    p.sb = reinterpret_cast<const uint32_t*>(superblocks);
    p.blocks_pitch = (long long)blocks_pitch;
    p.n_blocks = n_blocks;
    p.payload = reinterpret_cast<uint32_t*>(payload);
    p.block0 = reinterpret_cast<uint32_t*>(block0);
    p.status = status;
    rt::stream_t st = (rt::stream_t)cuda_stream;
    (void)st; /* (the emulation's launch macro runs the kernel in place and ignores the stream) */
    p.pass = 0;
    p.n_frames = n_frames;
    SDRD_LAUNCH(fec::decode_stream_kernel, dec_grid(n_frames), 1, fec::DS_NT, fec::ds_smem_bytes(), st, p);
    /* frames with more than 32 recovery blocks (flagged by the first pass) take the large-matrix build */
    p.pass = 1;
    SDRD_LAUNCH(fec::decode_kernel<128>, n_frames, 1, fec::NT, fec::dec_smem_bytes<128>(), st, p);
    if (!SDRD_LAUNCH_OK()) return fail_cuda("decode kernel launch");
    return 0;
}

extern "C" int sdrd_fec_decode(const uint8_t* superblocks, size_t blocks_pitch, const int* n_blocks, int n_frames,
                               uint8_t* payload, uint8_t* block0, int* status)
{
    if (n_frames < 0) return fail(SDRD_EINVAL, "negative frame count");
    if (!superblocks || !n_blocks || !payload || !status) return fail(SDRD_EINVAL, "null pointer");
    if (blocks_pitch < 1) return fail(SDRD_EINVAL, "blocks_pitch must be positive");
    if (n_frames == 0) return 0;
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    const size_t in_bytes = (size_t)n_frames * blocks_pitch * SDRD_UDPSIZE;
    const size_t pay_bytes = (size_t)n_frames * 127 * SDRD_BLOCK_BYTES;
    const size_t b0_bytes = (size_t)n_frames * SDRD_BLOCK_BYTES;
    const size_t int_bytes = (size_t)n_frames * sizeof(int);
    /* scratch layout: [superblocks][n_blocks][status] | [payload][block0] */
    const size_t a_need = round_up(in_bytes, 16) + 2 * round_up(int_bytes, 16);
    const size_t b_need = round_up(pay_bytes, 16) + round_up(b0_bytes, 16);
    if (g_scr_a.ensure(a_need) || g_scr_b.ensure(b_need)) return fail_cuda("allocating scratch");
    uint8_t* d_sb = (uint8_t*)g_scr_a.p;
    int* d_nb = (int*)(d_sb + round_up(in_bytes, 16));
    int* d_st = (int*)((uint8_t*)d_nb + round_up(int_bytes, 16));
    uint8_t* d_pay = (uint8_t*)g_scr_b.p;
    uint8_t* d_b0 = d_pay + round_up(pay_bytes, 16);
    SDRD_TRY(rt::copy(d_nb, n_blocks, int_bytes, rt::H2D, 0), "copy block counts to device");
    SDRD_TRY(rt::sync(0), "fec decode");
    /* Frames are independent: large calls go through in slices on three streams, so that the host -> device
     * copy of slice i + 1, the kernels of slice i and the device -> host copy of slice i - 1 overlap (PCIe is
     * full duplex; the call then costs little more than the larger of its two copies). */
    DecPipe& dp = g_dec_pipe;
    int n_slices = n_frames >= DecPipe::NS * 64 ? DecPipe::NS : (n_frames >= 128 ? n_frames / 64 : 1);
    if (dp.init()) return fail_cuda("creating streams");
    const int per = (n_frames + n_slices - 1) / n_slices;
    for (int i = 0; i < n_slices; i++) {
        const size_t f0 = (size_t)i * per, nf = std::min<size_t>(per, (size_t)n_frames - f0);
        SDRD_TRY(rt::copy(d_sb + f0 * blocks_pitch * SDRD_UDPSIZE, superblocks + f0 * blocks_pitch * SDRD_UDPSIZE,
                          nf * blocks_pitch * SDRD_UDPSIZE, rt::H2D, dp.s_in),
                 "copy datagrams to device");
        SDRD_TRY(rt::event_record(dp.ev_in[i], dp.s_in), "record copy event");
    }
    for (int i = 0; i < n_slices; i++) {
        const size_t f0 = (size_t)i * per, nf = std::min<size_t>(per, (size_t)n_frames - f0);
        SDRD_TRY(rt::stream_wait(dp.s_k, dp.ev_in[i]), "wait for copy");
        if (int rc = sdrd_fec_decode_dev(d_sb + f0 * blocks_pitch * SDRD_UDPSIZE, blocks_pitch, d_nb + f0, (int)nf,
                                         d_pay + f0 * 127 * SDRD_BLOCK_BYTES, d_b0 + f0 * SDRD_BLOCK_BYTES, d_st + f0, (void*)dp.s_k))
            return rc;
        SDRD_TRY(rt::event_record(dp.ev_k[i], dp.s_k), "record kernel event");
        SDRD_TRY(rt::stream_wait(dp.s_out, dp.ev_k[i]), "wait for kernels");
        SDRD_TRY(rt::copy(payload + f0 * 127 * SDRD_BLOCK_BYTES, d_pay + f0 * 127 * SDRD_BLOCK_BYTES, nf * 127 * SDRD_BLOCK_BYTES,
                          rt::D2H, dp.s_out),
                 "copy payload to host");
        if (block0)
            SDRD_TRY(rt::copy(block0 + f0 * SDRD_BLOCK_BYTES, d_b0 + f0 * SDRD_BLOCK_BYTES, nf * SDRD_BLOCK_BYTES, rt::D2H, dp.s_out),
                     "copy meta blocks to host");
    }
    SDRD_TRY(rt::stream_wait(dp.s_out, dp.ev_k[n_slices - 1]), "wait for kernels");
    SDRD_TRY(rt::copy(status, d_st, int_bytes, rt::D2H, dp.s_out), "copy status to host");
    SDRD_TRY(rt::sync(dp.s_out), "fec decode");
    return 0;
}

/* ---- cm256cc's descriptor API, one superframe per call (the seam include/cm256.h binds) ---- */

static int check_cm256_shape(const sdrd_cm256_params& p)
{
    /* cm256's own parameter checks first, then the shape this library implements */
    if (p.OriginalCount <= 0 || p.RecoveryCount <= 0 || p.BlockBytes <= 0) return fail(SDRD_EINVAL, "cm256: counts and block size must be positive");
    if (p.OriginalCount + p.RecoveryCount > 256) return fail(SDRD_EINVAL, "cm256: OriginalCount + RecoveryCount exceeds 256");
    if (p.OriginalCount != SDRD_NB_ORIGINAL || p.BlockBytes > SDRD_BLOCK_BYTES)
        return fail(SDRD_EINVAL, "cm256: this library implements sdrdaemon's superframe only (OriginalCount 128, BlockBytes <= 508)");
    return 0;
}

extern "C" int sdrd_cm256_encode_blocks(sdrd_cm256_params p, const sdrd_cm256_block* originals, void* recovery)
{
    if (int rc = check_cm256_shape(p)) return rc;
    if (!originals || !recovery) return fail(SDRD_EINVAL, "cm256: null pointer");
    for (int j = 0; j < 128; j++)
        if (!originals[j].Block) return fail(SDRD_EINVAL, "cm256: null block pointer");
    /* the blocks of a UDPSinkFEC slot lie 512 bytes apart (UDPSinkFEC.cpp:233-243): no gather needed then */
    const uint8_t* b0 = (const uint8_t*)originals[0].Block;
    const ptrdiff_t pitch = (const uint8_t*)originals[1].Block - b0;
    bool uniform = p.BlockBytes == SDRD_BLOCK_BYTES && pitch >= SDRD_BLOCK_BYTES && (pitch & 3) == 0;
    for (int j = 2; j < 128 && uniform; j++) uniform = (const uint8_t*)originals[j].Block - b0 == pitch * j;
    if (uniform) return sdrd_cm256_encode(b0, (size_t)pitch, 1, p.RecoveryCount, (uint8_t*)recovery);
    /* shorter blocks are zero-extended: the code is byte-wise linear, the padding encodes to zeros */
    std::vector<uint8_t> tmp((size_t)128 * SDRD_BLOCK_BYTES, 0), out((size_t)p.RecoveryCount * SDRD_BLOCK_BYTES);
    for (int j = 0; j < 128; j++) memcpy(&tmp[(size_t)j * SDRD_BLOCK_BYTES], originals[j].Block, (size_t)p.BlockBytes);
    if (int rc = sdrd_cm256_encode(tmp.data(), SDRD_BLOCK_BYTES, 1, p.RecoveryCount, out.data())) return rc;
    for (int r = 0; r < p.RecoveryCount; r++)
        memcpy((uint8_t*)recovery + (size_t)r * p.BlockBytes, &out[(size_t)r * SDRD_BLOCK_BYTES], (size_t)p.BlockBytes);
    return 0;
}

extern "C" int sdrd_cm256_decode_blocks(sdrd_cm256_params p, sdrd_cm256_block* blocks)
{
    if (int rc = check_cm256_shape(p)) return rc;
    if (!blocks) return fail(SDRD_EINVAL, "cm256: null pointer");
    /* classification as cm256's decoder initialisation does it */
    bool present[128] = {false};
    int rec_pos[128], n_rec = 0;
    bool rec_seen[128] = {false};
    for (int i = 0; i < 128; i++) {
        if (!blocks[i].Block) return fail(SDRD_EINVAL, "cm256: null block pointer");
        const int idx = blocks[i].Index;
        if (idx < 128) {
            if (present[idx]) return fail(SDRD_EINVAL, "cm256: repeated original block");
            present[idx] = true;
        } else {
            /* no bound against RecoveryCount: sdrdaemon passes the NUMBER of recovery blocks received there
             * (SDRdaemonFECBuffer.cpp:176), their rows are anywhere in 128 .. 255 */
            if (rec_seen[idx - 128]) return fail(SDRD_EINVAL, "cm256: repeated recovery block (singular system)");
            rec_seen[idx - 128] = true;
            rec_pos[n_rec++] = i;
        }
    }
    if (n_rec == 0) return 0; /* nothing erased */
    if (p.RecoveryCount == 1 && n_rec > 1)
        return fail(SDRD_EINVAL, "cm256: RecoveryCount is 1 but several recovery blocks were passed");
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    /* the frame as K3 takes it: 128 datagram images in array order, header word = {0, Index, 0} */
    std::vector<uint8_t> img((size_t)128 * SDRD_UDPSIZE, 0);
    for (int i = 0; i < 128; i++) {
        img[(size_t)i * SDRD_UDPSIZE + 2] = blocks[i].Index;
        memcpy(&img[(size_t)i * SDRD_UDPSIZE + 4], blocks[i].Block, (size_t)p.BlockBytes);
    }
    const size_t rec_bytes = (size_t)128 * SDRD_BLOCK_BYTES;
    if (g_scr_a.ensure(img.size() + 64) || g_scr_b.ensure(rec_bytes)) return fail_cuda("allocating scratch");
    uint8_t* d_img = (uint8_t*)g_scr_a.p;
    int* d_nb = (int*)(d_img + img.size());
    int* d_st = d_nb + 1;
    const int nb = 128;
    SDRD_TRY(rt::copy(d_img, img.data(), img.size(), rt::H2D, 0), "copy blocks to device");
    SDRD_TRY(rt::copy(d_nb, &nb, sizeof nb, rt::H2D, 0), "copy block count");
    fec::DecParams kp{};
    if (int rc = get_tables(&kp.tab)) return rc;
    kp.sb = reinterpret_cast<const uint32_t*>(d_img);
    kp.blocks_pitch = 128;
    kp.n_blocks = d_nb;
    kp.payload = nullptr;
    kp.block0 = nullptr;
    kp.status = d_st;
    kp.recovered = reinterpret_cast<uint32_t*>(g_scr_b.p);
    kp.general_single = p.RecoveryCount != 1;
    kp.pass = 0;
    kp.n_frames = 1;
    SDRD_LAUNCH(fec::decode_stream_kernel, 1, 1, fec::DS_NT, fec::ds_smem_bytes(), 0, kp);
    if (n_rec > 32) {
        kp.pass = 1;
        SDRD_LAUNCH(fec::decode_kernel<128>, 1, 1, fec::NT, fec::dec_smem_bytes<128>(), 0, kp);
    }
    if (!SDRD_LAUNCH_OK()) return fail_cuda("decode kernel launch");
    std::vector<uint8_t> out((size_t)n_rec * SDRD_BLOCK_BYTES);
    int st = 0;
    SDRD_TRY(rt::copy(out.data(), g_scr_b.p, out.size(), rt::D2H, 0), "copy recovered blocks to host");
    SDRD_TRY(rt::copy(&st, d_st, sizeof st, rt::D2H, 0), "copy status");
    SDRD_TRY(rt::sync(0), "cm256 decode");
    if (st != fec::ST_RECOVERED) return fail(SDRD_EINVAL, "cm256: the block set cannot be decoded");
    /* in place: recovery descriptor k (array order) receives erased original k (ascending), Index rewritten */
    int e = 0;
    for (int k = 0; k < n_rec; k++) {
        while (present[e]) e++;
        memcpy(blocks[rec_pos[k]].Block, &out[(size_t)k * SDRD_BLOCK_BYTES], (size_t)p.BlockBytes);
        blocks[rec_pos[k]].Index = (unsigned char)e++;
    }
    return 0;
}

/* ========================================================================================== */
/* batched receiver framing                                                                    */
/* ========================================================================================== */

struct sdrd_src {
    size_t max_dg = 0;
    int device = -1;             /* the device the handle lives on */
    /* the open slot: its first <= 128 datagrams (host copy), counters as SDRdaemonFECBuffer keeps them */
    std::vector<uint8_t> carry;      /* 128 x 512 */
    int frame_head = -1, block_count = 0, recovery_count = 0;
    int cur_nb_blocks = 0, cur_nb_recovery = 0, min_nb_blocks = 256, max_nb_recovery = 0;
    /* device side */
    uint8_t* d_dg = nullptr;         /* (128 + max_dg) datagrams: carried ones, then the call's */
    long long* d_start = nullptr;    /* [max_dg + 1] */
    int* d_nb = nullptr;             /* [max_dg + 1] */
    int* d_status = nullptr;
    uint8_t* d_payload = nullptr;    /* grow-only */
    uint8_t* d_block0 = nullptr;
    size_t frames_cap = 0;
    long long launches = 0;
    rt::stream_t stream = 0;
};

extern "C" int sdrd_src_create(sdrd_src** out, size_t max_datagrams)
{
    if (!out) return fail(SDRD_EINVAL, "null handle pointer");
    *out = nullptr;
    if (max_datagrams < 1) return fail(SDRD_EINVAL, "max_datagrams must be positive");
    std::string why;
    if (!rt::device_ok(why)) return fail(SDRD_ENODEV, why);
    sdrd_src* k = new (std::nothrow) sdrd_src();
    if (!k) return fail(SDRD_ENOMEM, "out of host memory");
    k->max_dg = max_datagrams;
    k->device = rt::current_device();
    k->carry.assign((size_t)128 * SDRD_UDPSIZE, 0);
    if (rt::alloc((void**)&k->d_dg, (128 + max_datagrams) * (size_t)SDRD_UDPSIZE) != 0 ||
        rt::alloc((void**)&k->d_start, (max_datagrams + 1) * sizeof(long long)) != 0 ||
        rt::alloc((void**)&k->d_nb, (max_datagrams + 1) * sizeof(int)) != 0 ||
        rt::alloc((void**)&k->d_status, (max_datagrams + 1) * sizeof(int)) != 0 || rt::stream_create(&k->stream) != 0) {
        int rc = fail_cuda("allocating receiver buffers");
        sdrd_src_destroy(k);
        return rc;
    }
    *out = k;
    return 0;
}

extern "C" void sdrd_src_destroy(sdrd_src* k)
{
    if (!k) return;
    SDRD_ON_DEVICE_OF(k);
    rt::sync(k->stream);
    rt::release(k->d_dg);
    rt::release(k->d_start);
    rt::release(k->d_nb);
    rt::release(k->d_status);
    rt::release(k->d_payload);
    rt::release(k->d_block0);
    rt::stream_destroy(k->stream);
    delete k;
}

extern "C" int sdrd_src_reset(sdrd_src* k)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    k->frame_head = -1;
    k->block_count = k->recovery_count = 0;
    k->cur_nb_blocks = k->cur_nb_recovery = 0;
    k->min_nb_blocks = 256;
    k->max_nb_recovery = 0;
    return 0;
}
extern "C" int sdrd_src_cur_nb_blocks(const sdrd_src* k) { return k ? k->cur_nb_blocks : -1; }
extern "C" int sdrd_src_cur_nb_recovery(const sdrd_src* k) { return k ? k->cur_nb_recovery : -1; }
extern "C" int sdrd_src_min_nb_blocks(sdrd_src* k)
{
    if (!k) return -1;
    int v = k->min_nb_blocks;
    k->min_nb_blocks = 256;
    return v;
}
extern "C" int sdrd_src_max_nb_recovery(sdrd_src* k)
{
    if (!k) return -1;
    int v = k->max_nb_recovery;
    k->max_nb_recovery = 0;
    return v;
}
extern "C" long long sdrd_src_launches(const sdrd_src* k) { return k ? k->launches : 0; }

extern "C" int sdrd_src_feed(sdrd_src* k, const uint8_t* dg, size_t n, uint8_t* payload, uint8_t* block0, size_t frame_capacity,
                             size_t* n_frames_p, int* status, int* nb_blocks, int* nb_recovery)
{
    if (!k) return fail(SDRD_EINVAL, "null handle");
    SDRD_ON_DEVICE_OF(k);
    if (n_frames_p) *n_frames_p = 0;
    if (n == 0) return 0;
    if (!dg) return fail(SDRD_EINVAL, "null datagram pointer");
    if (n > k->max_dg) return fail(SDRD_ERANGE, "more datagrams than the max_datagrams given at create time");
    /* ---- host: SDRdaemonFECBuffer::writeAndRead's bookkeeping, datagram by datagram (.cpp:118-168).
     * Device array = [stored datagrams of the open slot | this call's datagrams]; a frame's first <= 128
     * datagrams are contiguous in it. */
    const int carried = k->block_count < 128 ? k->block_count : 128; /* datagrams of the open slot held in `carry` */
    std::vector<long long> start;
    std::vector<int> nb, fr_blocks, fr_recovery;
    long long slot_start = 0; /* position of the open slot's first datagram in the device array */
    /* the counters move on a copy: they are committed, together with the carried datagrams, only once the call
     * can no longer fail (a refused call leaves the handle exactly as it was) */
    struct Book {
        int frame_head, block_count, recovery_count, cur_nb_blocks, cur_nb_recovery, min_nb_blocks, max_nb_recovery;
    } bk = {k->frame_head, k->block_count, k->recovery_count, k->cur_nb_blocks, k->cur_nb_recovery, k->min_nb_blocks, k->max_nb_recovery};
    for (size_t i = 0; i < n; i++) {
        const uint8_t* sb = dg + i * SDRD_UDPSIZE;
        const int frame_index = sb[0] | (sb[1] << 8);
        if (bk.frame_head != frame_index) {
            start.push_back(slot_start);
            nb.push_back(bk.block_count < 128 ? bk.block_count : 128);
            fr_blocks.push_back(bk.block_count);
            fr_recovery.push_back(bk.recovery_count);
            bk.cur_nb_blocks = bk.block_count;
            bk.cur_nb_recovery = bk.recovery_count;
            if (bk.cur_nb_blocks < bk.min_nb_blocks) bk.min_nb_blocks = bk.cur_nb_blocks;
            if (bk.cur_nb_recovery > bk.max_nb_recovery) bk.max_nb_recovery = bk.cur_nb_recovery;
            bk.block_count = 0;
            bk.recovery_count = 0;
            bk.frame_head = frame_index;
            slot_start = (long long)carried + (long long)i;
        }
        if (bk.block_count < 128 && sb[2] >= 128) bk.recovery_count++;
        bk.block_count++;
    }
    const size_t nf = start.size();
    if (nf > frame_capacity) return fail(SDRD_ERANGE, "frame_capacity smaller than the number of frames closed by this call");
    if (nf && (!payload || !status)) return fail(SDRD_EINVAL, "null output pointer");
    rt::stream_t st = k->stream;
    if (nf) {
        if (nf > k->frames_cap) {
            rt::sync(st);
            rt::release(k->d_payload);
            rt::release(k->d_block0);
            k->d_payload = k->d_block0 = nullptr;
            k->frames_cap = 0;
            if (rt::alloc((void**)&k->d_payload, nf * (size_t)127 * SDRD_BLOCK_BYTES) != 0 ||
                rt::alloc((void**)&k->d_block0, nf * (size_t)SDRD_BLOCK_BYTES) != 0)
                return fail_cuda("allocating receiver output buffers");
            k->frames_cap = nf;
        }
        if (carried) SDRD_TRY(rt::copy(k->d_dg, k->carry.data(), (size_t)carried * SDRD_UDPSIZE, rt::H2D, st), "copy carried datagrams");
        SDRD_TRY(rt::copy(k->d_dg + (size_t)carried * SDRD_UDPSIZE, dg, n * SDRD_UDPSIZE, rt::H2D, st), "copy datagrams to device");
        SDRD_TRY(rt::copy(k->d_start, start.data(), nf * sizeof(long long), rt::H2D, st), "copy frame table");
        SDRD_TRY(rt::copy(k->d_nb, nb.data(), nf * sizeof(int), rt::H2D, st), "copy frame table");
        fec::DecParams p{};
        if (int rc = get_tables(&p.tab)) return rc;
        p.sb = reinterpret_cast<const uint32_t*>(k->d_dg);
        p.blocks_pitch = 128;
        p.frame_start = k->d_start;
        p.n_blocks = k->d_nb;
        p.payload = reinterpret_cast<uint32_t*>(k->d_payload);
        p.block0 = reinterpret_cast<uint32_t*>(k->d_block0);
        p.status = k->d_status;
        p.pass = 0;
        p.n_frames = (int)nf;
        SDRD_LAUNCH(fec::decode_stream_kernel, dec_grid((long long)nf), 1, fec::DS_NT, fec::ds_smem_bytes(), st, p);
        p.pass = 1;
        SDRD_LAUNCH(fec::decode_kernel<128>, (int)nf, 1, fec::NT, fec::dec_smem_bytes<128>(), st, p);
        k->launches += 2;
        if (!SDRD_LAUNCH_OK()) return fail_cuda("decode kernel launch");
        SDRD_TRY(rt::copy(payload, k->d_payload, nf * (size_t)127 * SDRD_BLOCK_BYTES, rt::D2H, st), "copy payload to host");
        if (block0) SDRD_TRY(rt::copy(block0, k->d_block0, nf * (size_t)SDRD_BLOCK_BYTES, rt::D2H, st), "copy block 0 to host");
        SDRD_TRY(rt::copy(status, k->d_status, nf * sizeof(int), rt::D2H, st), "copy status to host");
        SDRD_TRY(rt::sync(st), "receiver decode");
        for (size_t f = 0; f < nf; f++) {
            if (nb_blocks) nb_blocks[f] = fr_blocks[f];
            if (nb_recovery) nb_recovery[f] = fr_recovery[f];
        }
    }
    /* ---- commit: counters, and the open slot's first <= 128 datagrams for the next call ---- */
    k->frame_head = bk.frame_head;
    k->block_count = bk.block_count;
    k->recovery_count = bk.recovery_count;
    k->cur_nb_blocks = bk.cur_nb_blocks;
    k->cur_nb_recovery = bk.cur_nb_recovery;
    k->min_nb_blocks = bk.min_nb_blocks;
    k->max_nb_recovery = bk.max_nb_recovery;
    {
        const int have = k->block_count < 128 ? k->block_count : 128;
        if (nf == 0) {
            /* the slot was open before this call: append what arrived */
            for (int j = carried; j < have; j++)
                memcpy(&k->carry[(size_t)j * SDRD_UDPSIZE], dg + (size_t)(j - carried) * SDRD_UDPSIZE, SDRD_UDPSIZE);
        } else {
            const size_t first = (size_t)(slot_start - carried); /* index in dg of the open slot's first datagram */
            memcpy(k->carry.data(), dg + first * SDRD_UDPSIZE, (size_t)have * SDRD_UDPSIZE);
        }
    }
    if (n_frames_p) *n_frames_p = nf;
    return 0;
}
