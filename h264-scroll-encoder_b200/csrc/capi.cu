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
#include "composer_gpu.h"
#include "compose_kernels.cuh"

#if defined(COMPOSE_TIMING) && COMPOSE_TIMING
namespace h264scroll { cudaError_t read_stage_cycles(unsigned long long *h8); }
#endif

namespace {

std::mutex g_mu;
int g_device = -1;
int g_sm_count = 0;
std::atomic<uint64_t> g_launches{0};
thread_local char g_err[256] = "";

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess) snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(g_err, sizeof g_err, "%s", what);
    return code;
}

/* scratch cached by h264gpu_compose_host */
struct Scratch {
    void *d = nullptr;
    size_t cap = 0;
    void *pin = nullptr;
    size_t pin_cap = 0;
} g_scratch;

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" {

int h264gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int h264gpu_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device == device) return H264GPU_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(H264GPU_ERR_NO_DEVICE, "no CUDA device visible", e); }
    if (device < 0 || device >= n) return fail(H264GPU_ERR_ARG, "device index out of range");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "cudaGetDeviceProperties", e);
    if (p.major != 10) {
        snprintf(g_err, sizeof g_err, "device %d is sm_%d%d; this library only carries sm_100a code", device, p.major, p.minor);
        return H264GPU_ERR_NO_DEVICE;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "cudaSetDevice", e);
    g_device = device;
    g_sm_count = p.multiProcessorCount;
    return H264GPU_OK;
}

void h264gpu_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device < 0) return;
    if (g_scratch.d) cudaFree(g_scratch.d);
    if (g_scratch.pin) cudaFreeHost(g_scratch.pin);
    g_scratch = Scratch();
    g_device = -1;
}

const char *h264gpu_last_error(void) { return g_err; }
int h264gpu_sm_count(void) { return g_sm_count; }
uint64_t h264gpu_launch_count(void) { return g_launches.load(); }

void *h264gpu_malloc(size_t bytes) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(H264GPU_ERR_NOMEM, "cudaMalloc", e); return nullptr; }
    return p;
}
void h264gpu_free(void *d) { if (d) cudaFree(d); }

void *h264gpu_host_alloc(size_t bytes) {
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
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (n_frames < 0 || !d_desc || !d_hints || !d_out || !d_sizes) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_compose(d_desc, n_frames, d_hints, d_dyn, d_dyn_tab, d_payload,
                                               d_out, d_sizes, g_sm_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "compose launch", e);
    if (n_frames > 0) g_launches.fetch_add(2);        /* reference-mode kernel + SPEC-mode kernel */
    return H264GPU_OK;
}

int h264gpu_raster_regions(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs_per_frame,
                           const H264GpuRegion *d_regions, const H264GpuMbHint *d_dyn_info,
                           H264GpuMbHint *d_hints, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (n_frames < 0 || !d_rdesc || !d_regions || !d_hints) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_raster(d_rdesc, n_frames, max_mbs_per_frame, d_regions, d_dyn_info,
                                              d_hints, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "raster launch", e);
    if (n_frames > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_canonicalize_hints(H264GpuMbHint *d_hints, size_t n_hints, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (!d_hints && n_hints) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_canonicalize(d_hints, n_hints, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "canonicalize launch", e);
    if (n_hints > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

size_t h264gpu_parse_annexb_scratch_words(size_t n_bytes, uint32_t max_nals) { return h264scroll::parse_annexb_scratch_words(n_bytes, max_nals); }

int h264gpu_parse_annexb(const uint8_t *d_data, size_t n_bytes, H264GpuNalUnit *d_units, uint32_t max_nals,
                         uint32_t *d_scratch, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (!d_data || !d_units || !d_scratch) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_parse_annexb(d_data, n_bytes, d_units, max_nals, d_scratch, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "annexb scan launch", e);
    if (n_bytes && max_nals) g_launches.fetch_add(2);
    return H264GPU_OK;
}

int h264gpu_unescape_chunk_bytes(void) { return h264scroll::unescape_chunk_bytes(); }

int h264gpu_unescape_nals(const H264GpuUnescapeDesc *d_desc, int n_nals, int max_chunks,
                          const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_scratch, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (n_nals < 0 || max_chunks < 0 || !d_desc || !d_src || !d_out || !d_sizes || !d_scratch) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_unescape(d_desc, n_nals, max_chunks, d_src, d_out, d_sizes, d_scratch, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "unescape launch", e);
    if (n_nals > 0 && max_chunks > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_rewrite_chunk_bytes(void) { return h264scroll::rewrite_chunk_bytes(); }

int h264gpu_rewrite_slices(const H264GpuRewriteDesc *d_desc, int n_slices, int max_chunks,
                           const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_counts, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (n_slices < 0 || max_chunks < 0 || !d_desc || !d_src || !d_out || !d_sizes || !d_counts) return fail(H264GPU_ERR_ARG, "null argument");
    cudaError_t e = h264scroll::launch_rewrite(d_desc, n_slices, max_chunks, d_src, d_out, d_sizes, d_counts, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(H264GPU_ERR_CUDA, "rewrite launch", e);
    if (n_slices > 0 && max_chunks > 0) g_launches.fetch_add(1);
    return H264GPU_OK;
}

int h264gpu_pack_frames(const H264GpuFrameDesc *d_desc, const uint32_t *d_sizes, int n_frames,
                        const uint8_t *d_slots, uint8_t *d_packed, uint64_t *d_offsets, void *stream) {
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
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
    if (g_device < 0) return fail(H264GPU_ERR_NO_DEVICE, "h264gpu_init was not called (or failed)");
    if (n_frames <= 0) return H264GPU_OK;
    if (!h_desc || !h_hints || !h_out || !h_sizes) return fail(H264GPU_ERR_ARG, "null argument");
    for (int f = 0; f < n_frames; f++) {
        const H264GpuFrameDesc &d = h_desc[f];
        if (d.mb_w == 0 || d.mb_h == 0 || d.mb_w > H264GPU_MAX_MB_W) return fail(H264GPU_ERR_ARG, "frame width out of range");
        if ((d.hint_off & 1) || (d.out_off & 15) || (d.out_cap & 15)) return fail(H264GPU_ERR_ARG, "misaligned frame descriptor");
        if (d.hint_off + (size_t)d.mb_w * d.mb_h > n_hints) return fail(H264GPU_ERR_ARG, "hints out of range");
        if (d.out_off + d.out_cap > out_bytes) return fail(H264GPU_ERR_ARG, "output slot out of range");
        if (d.dyn_id != H264GPU_NO_DYN && (int)d.dyn_id >= n_dyn) return fail(H264GPU_ERR_ARG, "dyn id out of range");
    }
    std::lock_guard<std::mutex> lk(g_mu);
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
                                   (const H264GpuMbHint *)(base + o_hint),
                                   (const H264GpuDynDesc *)(base + o_dyn),
                                   (const uint32_t *)(base + o_tab), (const uint8_t *)(base + o_pay),
                                   (uint8_t *)(base + o_out), (uint32_t *)(base + o_sz), g_sm_count, st);
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

#if defined(COMPOSE_TIMING) && COMPOSE_TIMING
/* experiments only: per-stage clock64() sums of thread 0 of every CTA since the last call */
int h264gpu_debug_stage_cycles(unsigned long long *h8) {
    return h264scroll::read_stage_cycles(h8) == cudaSuccess ? H264GPU_OK : H264GPU_ERR_CUDA;
}
#endif

}  /* extern "C" */
