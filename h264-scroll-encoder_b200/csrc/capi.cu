/*
 * capi.cu - the extern "C" surface declared in include/composer_gpu.h.
 *
 * Thin shim: device bring-up, raw memory/stream/event plumbing so that C (and
 * ctypes) callers never see a CUDA header, and the two launches.  There is no
 * CPU fallback: without a usable sm_100 device every call fails with
 * H264GPU_ERR_NO_DEVICE and a message in h264gpu_last_error().
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <chrono>
#include "composer_gpu.h"
#include "compose_kernels.cuh"

#if defined(COMPOSE_TIMING) && COMPOSE_TIMING
namespace h264scroll { cudaError_t read_stage_cycles(unsigned long long *h8); }
#endif
#if defined(COMPOSE_TRACE) && COMPOSE_TRACE
namespace h264scroll { cudaError_t read_trace(unsigned long long *h); }
#endif

namespace {

constexpr int PIPE_STREAMS = 3;            /* upload, kernels, download */
constexpr int PIPE_LAG = 2;                /* chunks enqueued before the first packed chunk is fetched */
constexpr int PIPE_MAX_CHUNKS = 64;

/* scratch cached by h264gpu_compose_host */
struct Scratch {
    void *d = nullptr;
    size_t cap = 0;
};

/* streams, events and buffers of the pipelined host-batch call */
struct Pipe {
    cudaStream_t st[PIPE_STREAMS] = {};
    cudaEvent_t ready[PIPE_MAX_CHUNKS] = {};      /* chunk c: packed, its frame offsets are on the host */
    cudaEvent_t uploaded[PIPE_MAX_CHUNKS] = {};   /* chunk c: inputs are on the device */
    bool made = false;
    char *dev = nullptr;  size_t dev_cap = 0;
    char *pin = nullptr;  size_t pin_cap = 0;
};

}  // namespace

/*
 * One context per (device, user): the device index, what the launches need to know about it, and the scratch the
 * host-buffer calls cache (device and pinned memory, streams, events).  Nothing is shared between contexts: two host
 * threads with a context each - on two GPUs or on one - never wait for each other.  The host-buffer calls of ONE context
 * serialise on its mutex (they share its scratch).
 */
struct H264GpuCtx {
    int device = -1;
    int sm_count = 0;
    std::mutex mu;
    Scratch scratch;
    Pipe pipe;
};

namespace {

std::mutex g_mu;                           /* guards g_default */
H264GpuCtx *g_default = nullptr;           /* the context h264gpu_init makes */
thread_local H264GpuCtx *t_current = nullptr;   /* h264gpu_ctx_make_current: this thread's context (else the default one) */
std::atomic<uint64_t> g_launches{0};
thread_local char g_err[256] = "";

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void pipe_release(Pipe &p) {
    for (int i = 0; i < PIPE_STREAMS; i++) if (p.st[i]) cudaStreamDestroy(p.st[i]);
    for (int i = 0; i < PIPE_MAX_CHUNKS; i++) if (p.ready[i]) cudaEventDestroy(p.ready[i]);
    for (int i = 0; i < PIPE_MAX_CHUNKS; i++) if (p.uploaded[i]) cudaEventDestroy(p.uploaded[i]);
    if (p.dev) cudaFree(p.dev);
    if (p.pin) cudaFreeHost(p.pin);
    p = Pipe();
}

/* the calling thread's context, its device made current for the thread (the current device is per host thread) */
H264GpuCtx *current() {
    H264GpuCtx *c = t_current ? t_current : g_default;
    if (!c) { fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init / h264gpu_ctx_create was not called (or failed)"); return nullptr; }
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) { fail(H264GPU_ERR_CUDA, "cudaSetDevice", e); return nullptr; }
    return c;
}

H264GpuCtx *ctx_new(int device, int *rc) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); *rc = fail(H264GPU_ERR_NO_DEVICE, "no CUDA device visible", e); return nullptr; }
    if (device < 0 || device >= n) { *rc = fail(H264GPU_ERR_ARG, "device index out of range"); return nullptr; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) { *rc = fail(H264GPU_ERR_CUDA, "cudaGetDeviceProperties", e); return nullptr; }
    if (p.major != 10) {
        snprintf(g_err, sizeof g_err, "device %d is sm_%d%d; this library only carries sm_100a code", device, p.major, p.minor);
        *rc = H264GPU_ERR_NO_DEVICE;
        return nullptr;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { *rc = fail(H264GPU_ERR_CUDA, "cudaSetDevice", e); return nullptr; }
    H264GpuCtx *c = new (std::nothrow) H264GpuCtx();
    if (!c) { *rc = fail(H264GPU_ERR_NOMEM, "context"); return nullptr; }
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    *rc = H264GPU_OK;
    return c;
}

void ctx_free(H264GpuCtx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->scratch.d) cudaFree(c->scratch.d);
    pipe_release(c->pipe);
    delete c;
}

}  // namespace

namespace {

/* what both host-buffer calls check about a batch's frame and dyn descriptors: ranges, alignment, missing arrays */
int validate_frames(const H264GpuFrameDesc *desc, int n, size_t n_hints, size_t out_bytes,
                    const H264GpuDynDesc *dyn, int n_dyn, const uint32_t *tab, size_t n_tab,
                    const uint8_t *payload, size_t payload_bytes) {
    for (int f = 0; f < n; f++) {
        const H264GpuFrameDesc &d = desc[f];
        if (d.mb_w == 0 || d.mb_h == 0 || d.mb_w > H264GPU_MAX_MB_W) return fail(H264GPU_ERR_ARG, "frame width out of range");
        if ((d.hint_off & 1) || (d.out_off & 15) || (d.out_cap & 15)) return fail(H264GPU_ERR_ARG, "misaligned frame descriptor");
        if (d.hint_off + (size_t)d.mb_w * d.mb_h > n_hints) return fail(H264GPU_ERR_ARG, "hints out of range");
        if (d.out_off + d.out_cap > out_bytes) return fail(H264GPU_ERR_ARG, "output slot out of range");
        if (d.dyn_id != H264GPU_NO_DYN) {
            if (!dyn || (int)d.dyn_id >= n_dyn) return fail(H264GPU_ERR_ARG, "dyn id out of range");
            const H264GpuDynDesc &y = dyn[d.dyn_id];
            if (!tab || (y.payload_bytes && !payload)) return fail(H264GPU_ERR_ARG, "dyn table / payload missing");
            if (y.table_off + y.n_mbs + 1 > n_tab || y.payload_off + y.payload_bytes > payload_bytes || (y.payload_off & 3))
                return fail(H264GPU_ERR_ARG, "dyn table / payload out of range");
        }
    }
    return H264GPU_OK;
}

}  // namespace

extern "C" {

int h264gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int h264gpu_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_default && g_default->device == device) return H264GPU_OK;
    int rc = H264GPU_OK;
    H264GpuCtx *c = ctx_new(device, &rc);
    if (!c) return rc;
    ctx_free(g_default);                              /* (a different device: the old default context and its scratch go) */
    g_default = c;
    return H264GPU_OK;
}

void h264gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    ctx_free(g_default);
    g_default = nullptr;
}

H264GpuCtx *h264gpu_ctx_create(int device) {
    int rc = H264GPU_OK;
    return ctx_new(device, &rc);
}

void h264gpu_ctx_destroy(H264GpuCtx *ctx) {
    if (t_current == ctx) t_current = nullptr;
    ctx_free(ctx);
}

int h264gpu_ctx_make_current(H264GpuCtx *ctx) {
    t_current = ctx;
    return current() ? H264GPU_OK : H264GPU_ERR_NO_DEVICE;
}

int h264gpu_ctx_device(const H264GpuCtx *ctx) { return ctx ? ctx->device : -1; }

const char *h264gpu_last_error(void) { return g_err; }
int h264gpu_sm_count(void) { H264GpuCtx *c = t_current ? t_current : g_default; return c ? c->sm_count : 0; }
uint64_t h264gpu_launch_count(void) { return g_launches.load(); }

void *h264gpu_malloc(size_t bytes) {
    if (!current()) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(H264GPU_ERR_NOMEM, "cudaMalloc", e); return nullptr; }
    return p;
}
void h264gpu_free(void *d) { if (d) cudaFree(d); }

void *h264gpu_host_alloc(size_t bytes) {
    if (!current()) return nullptr;
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(H264GPU_ERR_NOMEM, "cudaMallocHost", e); return nullptr; }
    return p;
}
void h264gpu_host_free(void *h) { if (h) cudaFreeHost(h); }

int h264gpu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? H264GPU_OK : fail(H264GPU_ERR_CUDA, "memcpy h2d", e);
}
int h264gpu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return e == cudaSuccess ? H264GPU_OK : fail(H264GPU_ERR_CUDA, "memcpy d2h", e);
}
int h264gpu_memset(void *dst, int value, size_t bytes, void *stream) {
    cudaError_t e = cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? H264GPU_OK : fail(H264GPU_ERR_CUDA, "memset", e);
}

void *h264gpu_stream_create(void) {
    if (!current()) return nullptr;
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return (void *)s;
}
void h264gpu_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
int h264gpu_stream_sync(void *s) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)s);
    return e == cudaSuccess ? H264GPU_OK : fail(H264GPU_ERR_CUDA, "stream sync", e);
}
void *h264gpu_event_create(void) {
    if (!current()) return nullptr;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) return nullptr;
    return (void *)ev;
}
void h264gpu_event_destroy(void *ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int h264gpu_event_record(void *ev, void *stream) {
    cudaError_t e = cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream);
    return e == cudaSuccess ? H264GPU_OK : fail(H264GPU_ERR_CUDA, "event record", e);
}
float h264gpu_event_elapsed_ms(void *a, void *b) {
    float ms = -1.f;
    if (cudaEventSynchronize((cudaEvent_t)b) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.f;
    return ms;
}

int h264gpu_compose_batch(const H264GpuFrameDesc *d_desc, int n_frames,
                          const H264GpuMbHint *d_hints, const H264GpuDynDesc *d_dyn,
                          const uint32_t *d_dyn_tab, const uint8_t *d_payload,
                          uint8_t *d_out, uint32_t *d_sizes, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_frames < 0 || !d_desc || !d_hints || !d_out || !d_sizes) return fail(H264GPU_ERR_ARG, "null argument");
    if (((uintptr_t)d_out | (uintptr_t)d_hints | (uintptr_t)d_desc) & 15u) return fail(H264GPU_ERR_ARG, "d_desc, d_hints and d_out must be 16-byte aligned");
    cudaError_t e = h264scroll::launch_compose(d_desc, n_frames, d_hints, 0, d_dyn, d_dyn_tab, d_payload,
                                               d_out, d_sizes, ctx->sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "compose launch", e);
    if (n_frames > 0) g_launches.fetch_add(2);        /* reference-mode kernel + SPEC-mode kernel */
    return H264GPU_OK;
}

int h264gpu_compose_batch_ex(const H264GpuFrameDesc *d_desc, int n_frames,
                             const H264GpuMbHint *d_hints, size_t n_hints, const H264GpuDynDesc *d_dyn,
                             const uint32_t *d_dyn_tab, const uint8_t *d_payload,
                             uint8_t *d_out, uint32_t *d_sizes, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_frames < 0 || !d_desc || !d_hints || !d_out || !d_sizes) return fail(H264GPU_ERR_ARG, "null argument");
    if (((uintptr_t)d_out | (uintptr_t)d_hints | (uintptr_t)d_desc) & 15u) return fail(H264GPU_ERR_ARG, "d_desc, d_hints and d_out must be 16-byte aligned");
    cudaError_t e = h264scroll::launch_compose(d_desc, n_frames, d_hints, n_hints, d_dyn, d_dyn_tab, d_payload,
                                               d_out, d_sizes, ctx->sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "compose launch", e);
    if (n_frames > 0) g_launches.fetch_add(2);
    return H264GPU_OK;
}

int h264gpu_raster_regions(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs_per_frame,
                           const H264GpuRegion *d_regions, const H264GpuMbHint *d_dyn_info,
                           H264GpuMbHint *d_hints, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_frames < 0 || !d_rdesc || !d_regions || !d_hints) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_raster(d_rdesc, n_frames, max_mbs_per_frame, d_regions, d_dyn_info,
                                              d_hints, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "raster launch", e);
    if (n_frames > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_canonicalize_hints(H264GpuMbHint *d_hints, size_t n_hints, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (!d_hints && n_hints) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_canonicalize(d_hints, n_hints, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "canonicalize launch", e);
    if (n_hints > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

size_t h264gpu_parse_annexb_scratch_words(size_t n_bytes, uint32_t max_nals) { return h264scroll::parse_annexb_scratch_words(n_bytes, max_nals); }

int h264gpu_parse_annexb(const uint8_t *d_data, size_t n_bytes, H264GpuNalUnit *d_units, uint32_t max_nals,
                         uint32_t *d_scratch, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (!d_data || !d_units || !d_scratch) return fail(H264GPU_ERR_ARG, "null argument");
    if (n_bytes > 0xFFFFFFF0ull) return fail(H264GPU_ERR_ARG, "stream too large for 32-bit offsets");
    cudaError_t e = h264scroll::launch_parse_annexb(d_data, n_bytes, d_units, max_nals, d_scratch, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "annexb scan launch", e);
    if (n_bytes && max_nals) g_launches.fetch_add(2);
    return H264GPU_OK;
}

int h264gpu_unescape_chunk_bytes(void) { return h264scroll::unescape_chunk_bytes(); }

int h264gpu_unescape_nals(const H264GpuUnescapeDesc *d_desc, int n_nals, int max_chunks,
                          const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_scratch, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_nals < 0 || max_chunks < 0 || !d_desc || !d_src || !d_out || !d_sizes || !d_scratch) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_unescape(d_desc, n_nals, max_chunks, d_src, d_out, d_sizes, d_scratch, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "unescape launch", e);
    if (n_nals > 0 && max_chunks > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_rewrite_chunk_bytes(void) { return h264scroll::rewrite_chunk_bytes(); }

int h264gpu_rewrite_slices(const H264GpuRewriteDesc *d_desc, int n_slices, int max_chunks,
                           const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_counts, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_slices < 0 || max_chunks < 0 || !d_desc || !d_src || !d_out || !d_sizes || !d_counts) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_rewrite(d_desc, n_slices, max_chunks, d_src, d_out, d_sizes, d_counts, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "rewrite launch", e);
    if (n_slices > 0 && max_chunks > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_pack_frames(const H264GpuFrameDesc *d_desc, const uint32_t *d_sizes, int n_frames,
                        const uint8_t *d_slots, uint8_t *d_packed, uint64_t *d_offsets, void *stream) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_frames < 0 || !d_desc || !d_sizes || !d_slots || !d_packed || !d_offsets) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_pack(d_desc, d_sizes, n_frames, d_slots, d_packed,
                                            (unsigned long long *)d_offsets, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "pack launch", e);
    if (n_frames > 0) g_launches.fetch_add(2);
    return H264GPU_OK;
}

int h264gpu_compose_host(const H264GpuFrameDesc *h_desc, int n_frames,
                         const H264GpuMbHint *h_hints, size_t n_hints,
                         const H264GpuDynDesc *h_dyn, int n_dyn,
                         const uint32_t *h_dyn_tab, size_t n_tab,
                         const uint8_t *h_payload, size_t payload_bytes,
                         uint8_t *h_out, size_t out_bytes, uint32_t *h_sizes) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (n_frames <= 0) return H264GPU_OK;
    if (!h_desc || !h_hints || !h_out || !h_sizes) return fail(H264GPU_ERR_ARG, "null argument");
    {
        const int rc = validate_frames(h_desc, n_frames, n_hints, out_bytes, h_dyn, n_dyn, h_dyn_tab, n_tab, h_payload, payload_bytes);
        if (rc) return rc;
        /* a hint's dyn_idx must lie inside its frame's table: checked where the hints are (they are host memory here) */
        for (int f = 0; f < n_frames; f++) {
            const H264GpuFrameDesc &d = h_desc[f];
            const uint32_t n_dyn_mbs = d.dyn_id != H264GPU_NO_DYN ? h_dyn[d.dyn_id].n_mbs : 0;
            const H264GpuMbHint *h = h_hints + d.hint_off;
            for (size_t i = 0, m = (size_t)d.mb_w * d.mb_h; i < m; i++)
                if (h[i].kind != H264GPU_MB_INTER && h[i].dyn_idx >= n_dyn_mbs) return fail(H264GPU_ERR_ARG, "a hint's dyn_idx lies outside the frame's table");
        }
    }
    std::lock_guard<std::mutex> lk(ctx->mu);
    Scratch &g_scratch = ctx->scratch;
    /* device layout: desc | hints | dyn | tab | payload(+8) | sizes | out */
    size_t o_desc = 0;
    size_t o_hint = align_up(o_desc + sizeof(H264GpuFrameDesc) * n_frames, 256);
    size_t o_dyn = align_up(o_hint + sizeof(H264GpuMbHint) * n_hints, 256);
    size_t o_tab = align_up(o_dyn + sizeof(H264GpuDynDesc) * (size_t)(n_dyn > 0 ? n_dyn : 0), 256);
    size_t o_pay = align_up(o_tab + sizeof(uint32_t) * n_tab, 256);
    size_t o_sz = align_up(o_pay + payload_bytes + 8, 256);
    size_t o_out = align_up(o_sz + sizeof(uint32_t) * n_frames, 256);
    size_t total = o_out + align_up(out_bytes, 256);
    if (total > g_scratch.cap) {
        if (g_scratch.d) cudaFree(g_scratch.d);
        g_scratch.d = nullptr; g_scratch.cap = 0;
        cudaError_t e = cudaMalloc(&g_scratch.d, total);
        if (e != cudaSuccess) return fail(H264GPU_ERR_NOMEM, "scratch cudaMalloc", e);
        g_scratch.cap = total;
    }
    char *base = (char *)g_scratch.d;
    cudaStream_t st = 0;
    cudaError_t e;
    e = cudaMemcpyAsync(base + o_desc, h_desc, sizeof(H264GpuFrameDesc) * n_frames, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_hint, h_hints, sizeof(H264GpuMbHint) * n_hints, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_dyn > 0) e = cudaMemcpyAsync(base + o_dyn, h_dyn, sizeof(H264GpuDynDesc) * n_dyn, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_tab > 0) e = cudaMemcpyAsync(base + o_tab, h_dyn_tab, sizeof(uint32_t) * n_tab, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && payload_bytes > 0) {
        e = cudaMemsetAsync(base + o_pay + (payload_bytes & ~(size_t)3), 0, 8, st);   /* zero the padded tail */
        if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_pay, h_payload, payload_bytes, cudaMemcpyHostToDevice, st);
    }
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "upload", e);
    e = h264scroll::launch_canonicalize((H264GpuMbHint *)(base + o_hint), n_hints, st);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "canonicalize launch", e);
    g_launches.fetch_add(1);
    e = h264scroll::launch_compose((const H264GpuFrameDesc *)(base + o_desc), n_frames,
                                   (const H264GpuMbHint *)(base + o_hint), n_hints,
                                   (const H264GpuDynDesc *)(base + o_dyn),
                                   (const uint32_t *)(base + o_tab), (const uint8_t *)(base + o_pay),
                                   (uint8_t *)(base + o_out), (uint32_t *)(base + o_sz), ctx->sm_count, st);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "compose launch", e);
    g_launches.fetch_add(2);
    e = cudaMemcpyAsync(h_sizes, base + o_sz, sizeof(uint32_t) * n_frames, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, base + o_out, out_bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "download", e);
    for (int f = 0; f < n_frames; f++)
        if (h_sizes[f] > h_desc[f].out_cap) return fail(H264GPU_ERR_OVERFLOW, "a frame exceeded its output slot");
    return H264GPU_OK;
}

}  /* extern "C" */

/* ---- host batch -> packed stream, pipelined over frame chunks -------------------------------- */

namespace {

cudaError_t pipe_setup(Pipe &g_pipe, size_t dev_bytes, size_t pin_bytes) {
    cudaError_t e = cudaSuccess;
    if (!g_pipe.made) {
        for (int i = 0; i < PIPE_STREAMS && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&g_pipe.st[i], cudaStreamNonBlocking);
        for (int i = 0; i < PIPE_MAX_CHUNKS && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&g_pipe.ready[i], cudaEventDisableTiming);
        for (int i = 0; i < PIPE_MAX_CHUNKS && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&g_pipe.uploaded[i], cudaEventDisableTiming);
        if (e != cudaSuccess) { pipe_release(g_pipe); return e; }      /* (nothing half-made is kept) */
        g_pipe.made = true;
    }
    if (dev_bytes > g_pipe.dev_cap) {
        if (g_pipe.dev) cudaFree(g_pipe.dev);
        g_pipe.dev = nullptr; g_pipe.dev_cap = 0;
        e = cudaMalloc((void **)&g_pipe.dev, dev_bytes);
        if (e != cudaSuccess) return e;
        g_pipe.dev_cap = dev_bytes;
    }
    if (pin_bytes > g_pipe.pin_cap) {
        if (g_pipe.pin) cudaFreeHost(g_pipe.pin);
        g_pipe.pin = nullptr; g_pipe.pin_cap = 0;
        e = cudaMallocHost((void **)&g_pipe.pin, pin_bytes);
        if (e != cudaSuccess) return e;
        g_pipe.pin_cap = pin_bytes;
    }
    return cudaSuccess;
}

}  // namespace

extern "C" {

int h264gpu_compose_regions_host(const H264GpuHostBatch *b, uint8_t *h_stream, size_t stream_cap,
                                 uint64_t *h_offsets, uint32_t *h_sizes, int n_chunks) {
    H264GpuCtx *ctx = current();
    if (!ctx) return H264GPU_ERR_NO_DEVICE;
    if (!b || !h_stream || !h_offsets) return fail(H264GPU_ERR_ARG, "null argument");
    const int n = b->n_frames;
    h_offsets[0] = 0;
    if (n <= 0) return H264GPU_OK;
    if (!b->desc || !b->rdesc || !b->regions) return fail(H264GPU_ERR_ARG, "null argument");
    if (n_chunks <= 0) n_chunks = 8;
    if (n_chunks > n) n_chunks = n;
    if (n_chunks > PIPE_MAX_CHUNKS) n_chunks = PIPE_MAX_CHUNKS;

    /* ---- validation: ranges and alignment ---- */
    {
        const int rc = validate_frames(b->desc, n, b->n_hints, b->slots_bytes, b->dyn, b->n_dyn, b->dyn_tab, b->n_tab, b->payload, b->payload_bytes);
        if (rc) return rc;
    }
    for (int f = 0; f < n; f++) {
        const H264GpuFrameDesc &d = b->desc[f];
        const H264GpuRasterDesc &r = b->rdesc[f];
        if (r.hint_off != d.hint_off || r.mb_w != d.mb_w || r.mb_h != d.mb_h) return fail(H264GPU_ERR_ARG, "raster and frame descriptor disagree");
        if (r.region_count > H264GPU_MAX_REGIONS || (uint64_t)r.region_first + r.region_count > b->n_regions) return fail(H264GPU_ERR_ARG, "regions out of range");
        for (uint32_t i = 0; i < r.region_count; i++) {
            const H264GpuRegion &g = b->regions[r.region_first + i];
            if (g.kind != H264GPU_REGION_DYN || g.mb_x1 <= g.mb_x0 || g.mb_y1 <= g.mb_y0) continue;
            const size_t z = (size_t)r.dyn_info_off + g.dyn_base + (size_t)(g.mb_x1 - g.mb_x0) * (size_t)(g.mb_y1 - g.mb_y0);
            if (!b->dyn_info || z > b->n_dyn_info) return fail(H264GPU_ERR_ARG, "a spliced region's dyn_info lies outside the array");
        }
    }

    std::lock_guard<std::mutex> lk(ctx->mu);
    Pipe &g_pipe = ctx->pipe;
    int bounds[PIPE_MAX_CHUNKS + 1];
    int chunk_frames = 0;
    for (int c = 0; c <= n_chunks; c++) bounds[c] = (int)((long long)n * c / n_chunks);
    for (int c = 0; c < n_chunks; c++) chunk_frames = bounds[c + 1] - bounds[c] > chunk_frames ? bounds[c + 1] - bounds[c] : chunk_frames;
    /* packed bytes of a chunk never exceed its slots */
    size_t chunk_cap = 0;
    for (int c = 0; c < n_chunks; c++) {
        size_t s = 0;
        for (int f = bounds[c]; f < bounds[c + 1]; f++) s += b->desc[f].out_cap;
        chunk_cap = s > chunk_cap ? s : chunk_cap;
    }
    chunk_cap = align_up(chunk_cap, 256);

    /* device layout (whole batch resident, arrays at the offsets the descriptors use) */
    const size_t o_desc = 0;
    const size_t o_rd = align_up(o_desc + sizeof(H264GpuFrameDesc) * n, 256);
    const size_t o_reg = align_up(o_rd + sizeof(H264GpuRasterDesc) * n, 256);
    const size_t o_info = align_up(o_reg + sizeof(H264GpuRegion) * b->n_regions, 256);
    const size_t o_dyn = align_up(o_info + sizeof(H264GpuMbHint) * b->n_dyn_info, 256);
    const size_t o_tab = align_up(o_dyn + sizeof(H264GpuDynDesc) * (size_t)(b->n_dyn > 0 ? b->n_dyn : 0), 256);
    const size_t o_pay = align_up(o_tab + sizeof(uint32_t) * b->n_tab, 256);
    const size_t o_hint = align_up(o_pay + b->payload_bytes + 32, 256);
    const size_t o_sz = align_up(o_hint + sizeof(H264GpuMbHint) * b->n_hints, 256);
    const size_t o_offs = align_up(o_sz + sizeof(uint32_t) * n, 256);
    const size_t o_slots = align_up(o_offs + 8 * (size_t)(chunk_frames + 1) * n_chunks, 256);
    const size_t o_packed = align_up(o_slots + b->slots_bytes, 256);
    const size_t dev_total = o_packed + chunk_cap * n_chunks;
    const size_t p_offs = 0, p_sz = align_up(8 * (size_t)(chunk_frames + 1) * n_chunks, 256);
    cudaError_t e = pipe_setup(g_pipe, dev_total, p_sz + sizeof(uint32_t) * n);
    if (e != cudaSuccess) return fail(H264GPU_ERR_NOMEM, "pipeline scratch", e);
    char *D = g_pipe.dev;
    uint64_t *pin_offs = (uint64_t *)(g_pipe.pin + p_offs);
    uint32_t *pin_sz = (uint32_t *)(g_pipe.pin + p_sz);

    /* item range [lo, hi) of each input array that the frames of a chunk refer to (frame-ordered layouts make
       these ranges tile the arrays; any other layout still works, it just uploads some items more than once) */
    struct Range { size_t lo = ~(size_t)0, hi = 0; void add(size_t a, size_t z) { if (z > a) { lo = a < lo ? a : lo; hi = z > hi ? z : hi; } } };
    struct ChunkRanges { Range reg, info, dyn, tab, pay; };
    auto ranges_of = [&](int f0, int f1) {
        ChunkRanges R;
        for (int f = f0; f < f1; f++) {
            const H264GpuRasterDesc &r = b->rdesc[f];
            R.reg.add(r.region_first, (size_t)r.region_first + r.region_count);
            for (uint32_t i = 0; i < r.region_count; i++) {
                const H264GpuRegion &g = b->regions[r.region_first + i];
                if (g.kind != H264GPU_REGION_DYN || g.mb_x1 <= g.mb_x0 || g.mb_y1 <= g.mb_y0) continue;
                size_t z = (size_t)r.dyn_info_off + g.dyn_base + (size_t)(g.mb_x1 - g.mb_x0) * (size_t)(g.mb_y1 - g.mb_y0);
                R.info.add(r.dyn_info_off, z);
            }
            const uint32_t id = b->desc[f].dyn_id;
            if (id != H264GPU_NO_DYN) {
                const H264GpuDynDesc &y = b->dyn[id];
                R.dyn.add(id, (size_t)id + 1);
                R.tab.add(y.table_off, y.table_off + y.n_mbs + 1);
                R.pay.add(y.payload_off, y.payload_off + y.payload_bytes);
            }
        }
        return R;
    };
    auto up = [&](size_t dev_off, const void *src, size_t item, size_t i0, size_t i1, cudaStream_t st) -> cudaError_t {
        if (i1 <= i0 || !src) return cudaSuccess;
        return cudaMemcpyAsync(D + dev_off + item * i0, (const char *)src + item * i0, item * (i1 - i0), cudaMemcpyHostToDevice, st);
    };

    /* H264GPU_PIPE_DEBUG=1: device-side timeline of every chunk on stderr (experiments) */
    static const bool dbg = getenv("H264GPU_PIPE_DEBUG") != nullptr;
    static cudaEvent_t dbg_ev[PIPE_MAX_CHUNKS][4], dbg_t0;
    static bool dbg_made = false;
    if (dbg && !dbg_made) {
        cudaEventCreate(&dbg_t0);
        for (int i = 0; i < PIPE_MAX_CHUNKS; i++) for (int k = 0; k < 4; k++) cudaEventCreate(&dbg_ev[i][k]);
        dbg_made = true;
    }
    size_t total = 0;
    int rc = H264GPU_OK;
    e = cudaMemsetAsync(D + o_pay + (b->payload_bytes & ~(size_t)3), 0, 32, g_pipe.st[0]);   /* the splice reads whole words (+ read-ahead) */
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_pipe.st[0]);
    if (dbg) cudaEventRecord(dbg_t0, g_pipe.st[0]);
    const auto h0 = std::chrono::steady_clock::now();
    auto hms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count(); };
    double h_enq[PIPE_MAX_CHUNKS] = {}, h_enq_done[PIPE_MAX_CHUNKS] = {}, h_ready[PIPE_MAX_CHUNKS] = {}, h_d2h[PIPE_MAX_CHUNKS] = {};
    /* three role streams - upload, kernels, download - so that each copy engine works through one copy at a time in
       chunk order (copies of several chunks in flight on one engine share its bandwidth and finish late).  The pack
       kernel writes the chunk's frame offsets straight into pinned host memory (12 KB: not worth a copy that would
       queue behind the previous chunk's stream on the download engine). */
    cudaStream_t s_in = g_pipe.st[0], s_k = g_pipe.st[1], s_out = g_pipe.st[2];
    if (e == cudaSuccess) e = up(o_desc, b->desc, sizeof(H264GpuFrameDesc), 0, n, s_in);
    if (e == cudaSuccess) e = up(o_rd, b->rdesc, sizeof(H264GpuRasterDesc), 0, n, s_in);
    if (e == cudaSuccess) e = up(o_reg, b->regions, sizeof(H264GpuRegion), 0, b->n_regions, s_in);
    if (e == cudaSuccess) e = up(o_dyn, b->dyn, sizeof(H264GpuDynDesc), 0, (size_t)(b->n_dyn > 0 ? b->n_dyn : 0), s_in);
    for (int c = 0; c < n_chunks + PIPE_LAG && e == cudaSuccess && rc == H264GPU_OK; c++) {
        if (c < n_chunks) {
            const int f0 = bounds[c], f1 = bounds[c + 1];
            const ChunkRanges R = ranges_of(f0, f1);
            if (dbg) { h_enq[c] = hms(); cudaEventRecord(dbg_ev[c][0], s_in); }
            /* (every copy costs the upload engine tens of microseconds while the download engine is busy: the small
               arrays went up whole in front of the loop, a chunk is three copies) */
            e = up(o_info, b->dyn_info, sizeof(H264GpuMbHint), R.info.lo, R.info.hi, s_in);
            if (e == cudaSuccess) e = up(o_tab, b->dyn_tab, sizeof(uint32_t), R.tab.lo, R.tab.hi, s_in);
            if (e == cudaSuccess) e = up(o_pay, b->payload, 1, R.pay.lo, R.pay.hi, s_in);
            if (e == cudaSuccess) e = cudaEventRecord(g_pipe.uploaded[c], s_in);
            if (dbg) cudaEventRecord(dbg_ev[c][1], s_in);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, g_pipe.uploaded[c], 0);
            if (e == cudaSuccess) e = h264scroll::launch_raster((const H264GpuRasterDesc *)(D + o_rd) + f0, f1 - f0, b->max_mbs_per_frame,
                                                                (const H264GpuRegion *)(D + o_reg), (const H264GpuMbHint *)(D + o_info),
                                                                (H264GpuMbHint *)(D + o_hint), s_k);
            if (e == cudaSuccess) e = h264scroll::launch_compose((const H264GpuFrameDesc *)(D + o_desc) + f0, f1 - f0, (const H264GpuMbHint *)(D + o_hint), b->n_hints,
                                                                 (const H264GpuDynDesc *)(D + o_dyn), (const uint32_t *)(D + o_tab),
                                                                 (const uint8_t *)(D + o_pay), (uint8_t *)(D + o_slots), (uint32_t *)(D + o_sz) + f0,
                                                                 ctx->sm_count, s_k);
            if (e == cudaSuccess) e = h264scroll::launch_pack((const H264GpuFrameDesc *)(D + o_desc) + f0, (const uint32_t *)(D + o_sz) + f0, f1 - f0,
                                                              (const uint8_t *)(D + o_slots), (uint8_t *)(D + o_packed) + chunk_cap * c,
                                                              (unsigned long long *)(pin_offs + (size_t)(chunk_frames + 1) * c), s_k);
            g_launches.fetch_add(5);
            if (e == cudaSuccess) e = cudaEventRecord(g_pipe.ready[c], s_k);
            if (dbg) { cudaEventRecord(dbg_ev[c][2], s_k); h_enq_done[c] = hms(); }
        }
        const int d = c - PIPE_LAG;
        if (d >= 0 && e == cudaSuccess) {
            const int f0 = bounds[d], f1 = bounds[d + 1];
            e = cudaEventSynchronize(g_pipe.ready[d]);
            if (e != cudaSuccess) break;
            if (dbg) h_ready[d] = hms();
            const volatile uint64_t *lo = pin_offs + (size_t)(chunk_frames + 1) * d;
            const size_t tot = (size_t)lo[f1 - f0];
            for (int f = f0; f < f1; f++) h_offsets[f] = total + lo[f - f0];
            if (total + tot > stream_cap) rc = fail(H264GPU_ERR_OVERFLOW, "the packed stream exceeds stream_cap");
            if (rc == H264GPU_OK && tot)
                e = cudaMemcpyAsync(h_stream + total, D + o_packed + chunk_cap * d, tot, cudaMemcpyDeviceToHost, s_out);
            if (dbg) { cudaEventRecord(dbg_ev[d][3], s_out); h_d2h[d] = hms(); }
            total += tot;
            h_offsets[f1] = total;
        }
    }
    /* the frames' own sizes (a frame that overflowed its slot reports more than the slot holds) */
    if (e == cudaSuccess) e = cudaMemcpyAsync(pin_sz, D + o_sz, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s_k);
    cudaError_t e2 = cudaSuccess;
    for (int i = 0; i < PIPE_STREAMS; i++) { cudaError_t t = cudaStreamSynchronize(g_pipe.st[i]); if (t != cudaSuccess) e2 = t; }
    if (e == cudaSuccess && e2 == cudaSuccess) {
        for (int f = 0; f < n; f++) {
            if (pin_sz[f] > b->desc[f].out_cap && rc == H264GPU_OK) rc = fail(H264GPU_ERR_OVERFLOW, "a frame exceeded its output slot");
            if (h_sizes) h_sizes[f] = pin_sz[f];
        }
    }
    if (dbg && e == cudaSuccess && e2 == cudaSuccess) {
        for (int c = 0; c < n_chunks; c++) {
            float t[4];
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], dbg_t0, dbg_ev[c][k]);
            fprintf(stderr, "chunk %2d: h2d %.3f..%.3f  kernels ..%.3f  d2h ..%.3f ms | host: enqueue %.3f..%.3f ready %.3f d2h enqueued %.3f\n",
                    c, t[0], t[1], t[2], t[3], h_enq[c], h_enq_done[c], h_ready[c], h_d2h[c]);
        }
    }
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "pipelined compose", e);
    if (e2 != cudaSuccess) return fail(H264GPU_ERR_CUDA, "pipelined compose (sync)", e2);
    return rc;
}

#if defined(COMPOSE_TRACE) && COMPOSE_TRACE
/* experiments only: the per-warp event trace of one CTA (32 warps x 1024 events) */
int h264gpu_debug_trace(unsigned long long *h) {
    return h264scroll::read_trace(h) == cudaSuccess ? H264GPU_OK : H264GPU_ERR_CUDA;
}
#endif
#if defined(COMPOSE_TIMING) && COMPOSE_TIMING
/* experiments only: per-stage clock64() sums of thread 0 of every CTA since the last call */
int h264gpu_debug_stage_cycles(unsigned long long *h8) {
    return h264scroll::read_stage_cycles(h8) == cudaSuccess ? H264GPU_OK : H264GPU_ERR_CUDA;
}
#endif

}  /* extern "C" */
