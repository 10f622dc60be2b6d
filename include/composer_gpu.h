/*
 * composer_gpu.h - C ABI of the B200 (sm_100a) P-slice composition path.
 *
 * This is the drop-in boundary for the reference's per-frame hot path:
 *
 *   reference                                             replaced by
 *   ---------------------------------------------------   -------------------------------
 *   h264_write_scroll_p_frame   src/h264_writer.c:541     h264gpu_compose_batch (dense hints)
 *   h264_write_waypoint_p_frame src/h264_writer.c:678     h264gpu_compose_batch (nal_header 0x41)
 *   get_mv_prediction/median3   src/h264_writer.c:362-432 kernel stage "predict"
 *   write_p16x16_mb             src/h264_writer.c:434     kernel stage "code"
 *   bitwriter_write_{ue,se,bit} src/bitwriter.c:13-101    closed-form (code,len) + scan + pack
 *   bitwriter_write_trailing_bits/_get_size  :103-131     kernel stage "finish"
 *   rbsp_to_ebsp/nal_write_unit src/nal.c:24-84           kernel stage "epb" (detect+scan+compact)
 *   per-MB region -> (ref,mv)   src/h264_writer.c:598-620 h264gpu_raster_regions (device rasteriser)
 *
 * The slice header, frame_num and the waypoint / long-term-reference
 * bookkeeping stay on the host (include/h264_writer.h); the host hands the
 * kernel the already-written slice-header bits.
 *
 * Plain C: pointers, sizes and PODs only.  No CPU fallback exists: every
 * entry point returns H264GPU_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef COMPOSER_GPU_H
#define COMPOSER_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (negative = failure) -------------------------------- */
#define H264GPU_OK               0
#define H264GPU_ERR_NO_DEVICE   -1   /* no CUDA device / wrong architecture  */
#define H264GPU_ERR_CUDA        -2   /* a CUDA call failed (see last_error)  */
#define H264GPU_ERR_ARG         -3   /* invalid argument                     */
#define H264GPU_ERR_OVERFLOW    -4   /* a frame did not fit its output slot  */
#define H264GPU_ERR_NOMEM       -5

/* ---- per-macroblock hint (8 bytes, raster order, dense) -------------- */
#define H264GPU_MB_INTER      0   /* P_L0_16x16 / P_Skip candidate with this MV      */
#define H264GPU_MB_DYN_INTER  1   /* spliced pre-encoded MB, neighbours see (ref,mv) */
#define H264GPU_MB_DYN_INTRA  2   /* spliced pre-encoded intra MB (ref -1, mv 0)     */

typedef struct {
    int16_t  mv_x_qpel;     /* quarter-pel, as coded (src/h264_writer.c:619-620) */
    int16_t  mv_y_qpel;
    uint8_t  ref_idx;       /* index into RefPicList0 (< 0xF0)                  */
    uint8_t  kind;          /* H264GPU_MB_*                                     */
    uint16_t dyn_idx;       /* kind != INTER: index into the frame's dyn table  */
} H264GpuMbHint;

/*
 * Canonical form (what h264gpu_compose_batch reads): a H264GPU_MB_DYN_INTRA hint carries mv 0 and
 * ref_idx 0xFF - "available, never matches", the (ref -1, mv 0) the reference publishes for an
 * intra neighbour (src/h264_writer.c:636-640).  h264gpu_raster_regions and h264gpu_compose_host
 * produce it; dense maps built elsewhere go through h264gpu_canonicalize_hints once.
 */

/* ---- frame flags ------------------------------------------------------ */
#define H264GPU_SKIP_OFF    0x00  /* reference behaviour: mb_skip_run always 0 (src/h264_writer.c:629-631) */
#define H264GPU_SKIP_SPEC   0x01  /* emit P_Skip where H.264 8.4.1.1 infers the hinted MV */
#define H264GPU_PRED_REF    0x00  /* reference predictor incl. its median3 / single-neighbour quirks */
#define H264GPU_PRED_SPEC   0x02  /* H.264 8.4.1.3 predictor                    */

#define H264GPU_NO_DYN      0xFFFFFFFFu
#define H264GPU_MAX_HDR_BITS 255

/* ---- frame descriptor (64 bytes) ------------------------------------- */
typedef struct {
    uint64_t hdr_bits[4];   /* slice-header bits, MSB first: stream bit k is
                               bit (63 - k%64) of hdr_bits[k/64]            */
    uint64_t hint_off;      /* index of the frame's first hint (even)       */
    uint64_t out_off;       /* byte offset of the output slot (16-aligned)  */
    uint32_t out_cap;       /* slot capacity in bytes (multiple of 16)      */
    uint32_t dyn_id;        /* index into the dyn descriptor array or NO_DYN */
    uint16_t mb_w, mb_h;
    uint8_t  hdr_nbits;     /* <= H264GPU_MAX_HDR_BITS                      */
    uint8_t  num_refs;      /* active refs: selects the ref_idx te() coding */
    uint8_t  flags;         /* H264GPU_SKIP_* | H264GPU_PRED_*              */
    uint8_t  nal_header;    /* (nal_ref_idc<<5)|nal_unit_type, src/nal.c:74 */
} H264GpuFrameDesc;

/* ---- dynamic-region (splice) descriptor (32 bytes) -------------------- */
typedef struct {
    uint64_t payload_off;   /* byte offset into the payload buffer (4-aligned)   */
    uint64_t table_off;     /* index into the bit-offset table: n_mbs+1 entries, */
                            /* MB j owns payload bits [t[j], t[j+1])             */
    uint32_t n_mbs;
    uint32_t payload_bytes; /* bytes readable from payload_off (padded to 4)     */
    uint64_t reserved;
} H264GpuDynDesc;

/* ---- region list (device rasteriser input), later regions occlude ---- */
typedef struct {
    int16_t  mb_x0, mb_y0, mb_x1, mb_y1;   /* half-open MB rectangle          */
    int16_t  mv_x_qpel, mv_y_qpel;
    uint8_t  ref_idx, kind; /* H264GPU_MB_INTER or H264GPU_REGION_DYN          */
    uint16_t dyn_base;      /* REGION_DYN: dyn-info index of the rect's first MB */
} H264GpuRegion;            /* 16 bytes */

#define H264GPU_REGION_DYN 3  /* region.kind: copy per-MB hints from the dyn-info table
                                 (entry = dyn_info_off + dyn_base + row-major index) */

typedef struct {
    uint32_t region_first;  /* index of the frame's first region              */
    uint32_t region_count;  /* <= H264GPU_MAX_REGIONS                          */
    uint64_t hint_off;      /* where to write the dense map (MB index)        */
    uint16_t mb_w, mb_h;
    uint32_t dyn_info_off;  /* frame's base index into the dyn-info table     */
} H264GpuRasterDesc;        /* 24 bytes */

#define H264GPU_MAX_REGIONS 512
#define H264GPU_MAX_MB_W    1023   /* widest frame the composer stages (16368 px) */

/* ---- device lifecycle --------------------------------------------------
 *
 * A context = one device + the scratch the host-buffer calls cache on it (device and pinned memory, streams, events).
 * h264gpu_init(device) makes the DEFAULT context, which every call uses unless the calling host thread has picked
 * another one with h264gpu_ctx_make_current.  The reference is single-threaded with no globals (distinct Composer
 * objects are independent); here distinct contexts are: one host thread + one context per GPU run side by side without
 * sharing a lock, a stream or a byte of scratch.  Calls on ONE context from several threads are safe, the
 * host-buffer calls among them serialise.  Every call (re)selects its context's device for the calling thread.
 */
typedef struct H264GpuCtx H264GpuCtx;
int         h264gpu_init(int device);          /* makes / replaces the default context; idempotent per device */
void        h264gpu_shutdown(void);            /* destroys the default context */
H264GpuCtx *h264gpu_ctx_create(int device);    /* NULL on failure (h264gpu_last_error) */
void        h264gpu_ctx_destroy(H264GpuCtx *ctx);
int         h264gpu_ctx_make_current(H264GpuCtx *ctx);   /* for the calling host thread; NULL = back to the default context */
int         h264gpu_ctx_device(const H264GpuCtx *ctx);
int         h264gpu_device_count(void);
const char *h264gpu_last_error(void);
int         h264gpu_sm_count(void);

/* ---- raw memory / stream plumbing (so callers need no CUDA headers) --- */
void *h264gpu_malloc(size_t bytes);            /* device memory            */
void  h264gpu_free(void *dptr);
void *h264gpu_host_alloc(size_t bytes);        /* pinned host memory       */
void  h264gpu_host_free(void *hptr);
int   h264gpu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int   h264gpu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int   h264gpu_memset(void *dst, int value, size_t bytes, void *stream);
void *h264gpu_stream_create(void);
void  h264gpu_stream_destroy(void *stream);
int   h264gpu_stream_sync(void *stream);       /* NULL = default stream    */
void *h264gpu_event_create(void);
void  h264gpu_event_destroy(void *ev);
int   h264gpu_event_record(void *ev, void *stream);
float h264gpu_event_elapsed_ms(void *ev_start, void *ev_stop); /* syncs on stop */

/* ---- the hot path ------------------------------------------------------ */

/*
 * Compose n_frames P-slices.  All pointers are DEVICE pointers; d_desc, d_hints and d_out 16-byte aligned (the kernel
 * moves descriptors, hints and output groups as 128-bit words; anything else is refused with H264GPU_ERR_ARG).
 *   d_desc      n_frames frame descriptors
 *   d_hints     dense hint maps, indexed by desc.hint_off
 *   d_dyn       dyn descriptors (may be NULL when no frame splices)
 *   d_dyn_tab   bit-offset tables, d_payload the pre-encoded bits
 *   d_out       output arena; frame f writes Annex-B bytes at out_off
 *   d_sizes     per frame: bytes of "00 00 00 01 | header | EBSP" that the
 *               frame needs; when it exceeds out_cap only out_cap bytes were
 *               stored and the caller must treat the frame as overflowed
 *               (the reference asserts instead, src/nal.c:34,59)
 * Asynchronous on `stream`; returns a launch status only.
 */
int h264gpu_compose_batch(const H264GpuFrameDesc *d_desc, int n_frames,
                          const H264GpuMbHint *d_hints,
                          const H264GpuDynDesc *d_dyn,
                          const uint32_t *d_dyn_tab,
                          const uint8_t *d_payload,
                          uint8_t *d_out, uint32_t *d_sizes,
                          void *stream);

/*
 * The same for a caller that can say how many hints lie behind d_hints (n_hints, e.g. the size it allocated): the dense
 * maps are then also described to the TMA unit as a 2-D tensor (rows of 128 bytes, 128-byte swizzle), and frames that start
 * on a row - hint_off a multiple of 16, mb_w a multiple of 8, d_hints 128-byte aligned - get their hint tiles by tensor
 * loads instead of per-thread copies.  Results are identical; h264gpu_compose_batch is this call with n_hints = 0.
 */
int h264gpu_compose_batch_ex(const H264GpuFrameDesc *d_desc, int n_frames,
                             const H264GpuMbHint *d_hints, size_t n_hints,
                             const H264GpuDynDesc *d_dyn,
                             const uint32_t *d_dyn_tab,
                             const uint8_t *d_payload,
                             uint8_t *d_out, uint32_t *d_sizes,
                             void *stream);

/*
 * Expand region lists into dense hint maps on the device
 * (replaces the per-MB region test of src/h264_writer.c:598-620).  The descriptors live in device memory, so nothing
 * about them can be refused here: region_count > H264GPU_MAX_REGIONS is invalid, the kernel then paints the first
 * H264GPU_MAX_REGIONS regions only (h264gpu_compose_regions_host, which sees the descriptors, rejects it).
 */
int h264gpu_raster_regions(const H264GpuRasterDesc *d_rdesc, int n_frames, int max_mbs_per_frame,
                           const H264GpuRegion *d_regions,
                           const H264GpuMbHint *d_dyn_info,   /* may be NULL without DYN regions */
                           H264GpuMbHint *d_hints, void *stream);

/* Rewrites spliced-intra hints of a dense map into canonical form, in place (asynchronous on `stream`). */
int h264gpu_canonicalize_hints(H264GpuMbHint *d_hints, size_t n_hints, void *stream);

/*
 * Exact output packing: gathers the frames' slots into one contiguous Annex-B
 * stream in frame order (the append order of src/nal.c:55-83, src/composer.c:255-264).
 * d_offsets receives n_frames+1 byte offsets, d_offsets[n_frames] = stream size.
 * A frame that overflowed its slot contributes out_cap bytes.
 */
int h264gpu_pack_frames(const H264GpuFrameDesc *d_desc, const uint32_t *d_sizes, int n_frames,
                        const uint8_t *d_slots, uint8_t *d_packed, uint64_t *d_offsets,
                        void *stream);

/*
 * Host-buffer convenience: copies descriptors/hints/dyn data up, composes,
 * copies the slots and sizes back, synchronises.  Frames are returned in
 * their slots (h_out + out_off).  Returns H264GPU_ERR_OVERFLOW if any frame
 * exceeded its slot.  Scratch device memory is cached between calls.
 */
int h264gpu_compose_host(const H264GpuFrameDesc *h_desc, int n_frames,
                         const H264GpuMbHint *h_hints, size_t n_hints,
                         const H264GpuDynDesc *h_dyn, int n_dyn,
                         const uint32_t *h_dyn_tab, size_t n_tab,
                         const uint8_t *h_payload, size_t payload_bytes,
                         uint8_t *h_out, size_t out_bytes, uint32_t *h_sizes);

/*
 * Host batch -> packed Annex-B stream, pipelined (the whole path behind composer_write_scroll_frames /
 * a hinted batch, src/composer.c:255-264): the batch is cut into frame chunks; while chunk c is uploaded,
 * chunk c-1 is rasterised, composed and packed and the stream of chunk c-2 is on its way back (PCIe is full
 * duplex).  Every input array is indexed by the batch's own descriptors (absolute offsets); a chunk uploads
 * the range of each array its frames refer to, so arrays laid out in frame order cross the bus exactly once
 * (any other layout still works).  desc[f].out_off / out_cap describe
 * device-side slots only (any non-overlapping 16-aligned layout); the result arrives packed:
 *   h_stream   the frames' NAL units back to back in frame order, h_offsets[f] = first byte of frame f,
 *              h_offsets[n_frames] = stream size (n_frames + 1 entries)
 *   h_sizes    may be NULL; bytes per frame
 * Host buffers should come from h264gpu_host_alloc (pinned): pageable memory works but does not overlap.
 * Device and pinned scratch, streams and events are cached between calls.  n_chunks <= 0 picks a default.
 * Returns H264GPU_ERR_OVERFLOW if a frame exceeded its slot or the stream exceeds stream_cap.
 */
typedef struct {
    const H264GpuFrameDesc  *desc;     int    n_frames;
    const H264GpuRasterDesc *rdesc;    /* n_frames */
    const H264GpuRegion     *regions;  size_t n_regions;
    const H264GpuMbHint     *dyn_info; size_t n_dyn_info;   /* may be NULL / 0 */
    const H264GpuDynDesc    *dyn;      int    n_dyn;        /* may be NULL / 0 */
    const uint32_t          *dyn_tab;  size_t n_tab;
    const uint8_t           *payload;  size_t payload_bytes;
    size_t n_hints;                    /* macroblocks the dense maps span (max hint_off + mb_w * mb_h) */
    size_t slots_bytes;                /* bytes the output slots span (max out_off + out_cap)          */
    int    max_mbs_per_frame;
} H264GpuHostBatch;

int h264gpu_compose_regions_host(const H264GpuHostBatch *batch, uint8_t *h_stream, size_t stream_cap,
                                 uint64_t *h_offsets, uint32_t *h_sizes, int n_chunks);

/* ---- slice rewrite (the reference's I-frame path) -------------------------
 *
 * h264_rewrite_idr_frame / h264_rewrite_as_non_idr_i_frame (src/h264_writer.c:242-350) for whole batches: every
 * slice becomes   00 00 00 01 | nal_header | EBSP(hdr_bits | source RBSP bits [src_start_bit, +n_bits) | zero pad),
 * which is copy_bits (src/h264_writer.c:228-240) + nal_write_unit (src/nal.c:52-84).  The host parses the source
 * slice header and writes the new one (as for P slices); the device moves the megabytes.
 */
typedef struct {
    uint64_t hdr_bits[4];   /* new slice header, MSB first (as in H264GpuFrameDesc)            */
    uint64_t src_off;       /* byte offset of the source RBSP in d_src (multiple of 4)        */
    uint64_t src_bytes;     /* its size; 16 more bytes must be readable behind it (d_src 16-aligned) */
    uint64_t src_start_bit; /* first source bit to copy (mb data start of the parsed header)  */
    uint64_t n_bits;        /* bits to copy (src_bytes * 8 - src_start_bit in the reference)  */
    uint64_t out_off;       /* byte offset of the output slot (16-aligned)                    */
    uint32_t out_cap;       /* slot capacity in bytes (multiple of 16)                        */
    uint16_t hdr_nbits;     /* <= H264GPU_MAX_HDR_BITS                                        */
    uint8_t  nal_header;    /* (nal_ref_idc << 5) | nal_unit_type                             */
    uint8_t  reserved;
} H264GpuRewriteDesc;       /* 80 bytes */

/* bytes of the pre-emulation-prevention string one CTA handles; a slice of B string bytes needs ceil(B / this) chunks */
int h264gpu_rewrite_chunk_bytes(void);

/*
 * All pointers are DEVICE pointers.  max_chunks >= chunks of the largest slice; d_scratch = 1 + n_slices * max_chunks
 * uint32 (chunk tickets and the look-back words that carry the insertion counts from chunk to chunk; zeroed by the
 * call).  d_sizes[i] = bytes the NAL unit needs (> out_cap: overflowed, out_cap bytes stored).  Asynchronous on `stream`.
 */
int h264gpu_rewrite_slices(const H264GpuRewriteDesc *d_desc, int n_slices, int max_chunks,
                           const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_scratch, void *stream);

/* ---- EBSP -> RBSP (the input side of the splice) ---------------------------
 *
 * ebsp_to_rbsp (src/nal_parser.c:67-88) for whole batches of NAL payloads: drops every emulation prevention byte
 * (a 0x03 behind two zero bytes and in front of a byte <= 0x03).
 */
typedef struct {
    uint64_t src_off;       /* byte offset of the EBSP payload in d_src (multiple of 16)      */
    uint64_t src_bytes;     /* its size; 16 more bytes must be readable behind it             */
    uint64_t out_off;       /* byte offset of the RBSP in d_out (multiple of 16)              */
    uint32_t out_cap;       /* capacity in bytes (src_bytes always suffices)                  */
    uint32_t reserved;
} H264GpuUnescapeDesc;      /* 32 bytes */

int h264gpu_unescape_chunk_bytes(void);   /* payload bytes one CTA handles */

/* Device pointers; max_chunks >= chunks of the largest payload; d_scratch = 1 + n_nals * max_chunks uint32 (zeroed by
 * the call); d_sizes[i] = RBSP bytes.  Asynchronous on `stream`. */
int h264gpu_unescape_nals(const H264GpuUnescapeDesc *d_desc, int n_nals, int max_chunks,
                          const uint8_t *d_src, uint8_t *d_out, uint32_t *d_sizes, uint32_t *d_scratch, void *stream);

/* ---- Annex-B scan -----------------------------------------------------------
 *
 * nal_parser_next (src/nal_parser.c:28-65) over a whole stream: one record per NAL unit, in stream order.
 */
typedef struct {
    uint32_t offset;        /* of the payload behind the NAL header byte (NALUnit.data - stream start) */
    uint32_t size;          /* payload bytes, trailing zero bytes removed (NALUnit.size)              */
    uint8_t  nal_ref_idc, nal_unit_type, reserved[2];
} H264GpuNalUnit;           /* 12 bytes */

/* words of scratch h264gpu_parse_annexb needs for a stream of n_bytes and up to max_nals units */
size_t h264gpu_parse_annexb_scratch_words(size_t n_bytes, uint32_t max_nals);

/* d_data must be readable 20 bytes past n_bytes and 16-byte aligned; n_bytes < 4 GiB.  The unit count lands in
 * d_scratch[max_nals]; it may exceed max_nals: then the first max_nals records were written, each of them complete (the
 * last one ends at the start code behind it).  Unlike nal_parser_next, which stops at an empty unit (two start codes
 * back to back, or one at the very end), the scan reports such a unit as a record of size 0 and type 0 and goes on.
 * Asynchronous on `stream`. */
int h264gpu_parse_annexb(const uint8_t *d_data, size_t n_bytes, H264GpuNalUnit *d_units, uint32_t max_nals,
                         uint32_t *d_scratch, void *stream);

/* Count of kernels this library has launched since init (for bench accounting). */
uint64_t h264gpu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* COMPOSER_GPU_H */
