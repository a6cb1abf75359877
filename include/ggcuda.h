/*
 * ggcuda.h -- C ABI of libggcuda.so, the B200 (sm_100a) scene rasteriser that plugs in
 * behind gogpu/gg's GPUAccelerator boundary.
 *
 * Every entry point is what the Go side's cgo file (accel_cuda.go, build tag `ggcuda`)
 * binds; INTEGRATION.md shows that binding. Reference interfaces replaced (gogpu/gg):
 *   gg.GPUAccelerator                      accelerator.go:104-140
 *   gg.GPURenderTarget                     accelerator.go:61-92
 *   internal/gpu.VelloAccelerator          vello_accelerator.go:197-386 (accumulate + Flush)
 *   internal/gpu.VelloComputeDispatcher    vello_compute.go:1112-1225 (9 WGSL passes + readback)
 *   scene.GPUSceneRenderer / scene.Encoding scene/gpu_renderer.go:73-213, scene/encoding.go:407-444
 *
 * Conventions: plain pointers and sizes, no ownership transfer (inputs are copied before the
 * call returns, as cgo requires), int status (0 = ok, negative = GGCUDA_ERR_*), the message
 * for the last failure via ggcuda_last_error(). Calls on one context must be serialised by
 * the caller (the reference accelerator holds one mutex, vello_accelerator.go:37,202); any
 * OS thread may make them (the library binds its device on every call).
 */
#ifndef GGCUDA_H
#define GGCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GGCUDA_API __attribute__((visibility("default")))
#else
#define GGCUDA_API
#endif

typedef struct ggcuda_ctx ggcuda_ctx;

enum {
    GGCUDA_OK = 0,
    GGCUDA_ERR_CUDA = -1,        /* CUDA runtime failure; maps to a Go error (gg falls back to CPU) */
    GGCUDA_ERR_INVALID = -2,     /* bad argument */
    GGCUDA_ERR_UNSUPPORTED = -3, /* content this path cannot render (== gg.ErrFallbackToCPU, accelerator.go:16) */
    GGCUDA_ERR_NOMEM = -4
};

/* gg.PathVerb (path.go): MoveTo 0, LineTo 1, QuadTo 2, CubicTo 3, Close 4 */
enum { GGCUDA_VERB_MOVE = 0, GGCUDA_VERB_LINE = 1, GGCUDA_VERB_QUAD = 2, GGCUDA_VERB_CUBIC = 3, GGCUDA_VERB_CLOSE = 4 };
/* gg.FillRule: NonZero 0, EvenOdd 1 */
enum { GGCUDA_FILL_NONZERO = 0, GGCUDA_FILL_EVENODD = 1 };
/* gg.LineCap / gg.LineJoin (paint.go): Butt 0, Round 1, Square 2 / Miter 0, Round 1, Bevel 2 */
enum { GGCUDA_CAP_BUTT = 0, GGCUDA_CAP_ROUND = 1, GGCUDA_CAP_SQUARE = 2 };
enum { GGCUDA_JOIN_MITER = 0, GGCUDA_JOIN_ROUND = 1, GGCUDA_JOIN_BEVEL = 2 };

/* flags for ggcuda_flush / ggcuda_render_device */
enum {
    GGCUDA_COMPOSITE_OVER = 1,  /* rasterise the scene on transparent and source-over it onto the pixels already in dst with
                                   VelloAccelerator.compositeOver's byte arithmetic, (d * (255 - sA) + 127) / 255
                                   (vello_accelerator.go:388-442), instead of starting from the background colour */
    GGCUDA_KEEP_SCENE = 2,      /* do not clear the accumulated scene after rendering */
    GGCUDA_TARGET_F32 = 4,      /* ggcuda_render_device only: dst holds premultiplied float RGBA, 16 bytes per pixel
                                   (stride in bytes, 16-byte aligned), instead of RGBA8 */
    GGCUDA_NO_WAIT = 8          /* ggcuda_render_device[_multi] only: return as soon as the pass is on the stream when the same
                                   scene has already been rendered once by this context with the buffers it has now (its
                                   element counts are known to fit: nothing to check). The host can then queue frames ahead
                                   of the device; ggcuda_sync waits for them. Ignored (the call waits) in every other case. */
};

/* flags for ggcuda_create */
enum {
    GGCUDA_CREATE_HOST_STROKES = 1  /* diagnostic: expand strokes with the host polyline stroker (the path the
                                       reference takes, internal/stroke/expander.go) instead of on the device */
};

/* ---- lifetime: GPUAccelerator.Init / Close (accelerator.go:108-112) ---- */
GGCUDA_API int ggcuda_create(int device, uint32_t flags, ggcuda_ctx** out);
GGCUDA_API void ggcuda_destroy(ggcuda_ctx* ctx);
GGCUDA_API const char* ggcuda_last_error(ggcuda_ctx* ctx);        /* ctx may be NULL for create failures */
/* Use an externally owned cudaStream_t (e.g. the caller's current stream); NULL restores the context's own. */
GGCUDA_API int ggcuda_set_stream(ggcuda_ctx* ctx, void* cuda_stream);

/* ---- frame set-up ---- */
/* Start accumulating a scene for a width x height target (GPURenderTarget.Width/Height). */
GGCUDA_API int ggcuda_begin(ggcuda_ctx* ctx, uint32_t width, uint32_t height);
/* Resident scenes (scene.Encoding.Hash, scene/encoding.go:752-802; dirty tiles, scene/renderer.go:395-433): like
 * ggcuda_begin, but the scene accumulated afterwards is remembered under `key` once it has been rendered. If the context
 * still holds the segments and command lists of a scene with the same key, size and band, *resident is set to 1, NOTHING
 * may be added, and the next render runs fine rasterisation only (no ingest, no upload, no flatten / binning / coarse). */
GGCUDA_API int ggcuda_begin_keyed(ggcuda_ctx* ctx, uint32_t width, uint32_t height, uint64_t key, int* resident);
/* scene.Encoding.Hash (scene/encoding.go:752-802: FNV-1a over the elements of the tag, path-data, draw-data and transform
 * streams) for hosts that do not have it; with brushes_rgba != NULL the brush colours are folded in behind it (the
 * reference's hash ignores them; a key for re-using pixels must not). n_transforms counts floats. */
GGCUDA_API uint64_t ggcuda_encoding_hash(const uint8_t* tags, size_t n_tags, const float* path_data, size_t n_path_data,
                                         const uint32_t* draw_data, size_t n_draw_data, const float* transforms, size_t n_transforms,
                                         const double* brushes_rgba, size_t n_brushes);
/* Register an image of the frame being built (scene.Image, scene/scene.go:769-779: premultiplied RGBA8, row-major, the bytes
 * gg.Pixmap.ToImage() produces). Images are numbered from 0 in the order they are added after ggcuda_begin; TagImage entries
 * of an encoding added afterwards refer to them by that number (scene/encoding.go:656-661) and are drawn as the CPU tile
 * renderer draws them (scene/renderer.go:1093-1243: inverse affine, bilinear in premultiplied space with clamp-to-edge,
 * the nearest texel when the sample falls on a texel centre, source-over). The pixels are copied. */
GGCUDA_API int ggcuda_add_image(ggcuda_ctx* ctx, uint32_t width, uint32_t height, const uint8_t* premul_rgba, uint32_t* index_out);
/* Restrict the next render (and its read-back) to the 16x16 tiles touching the pixel rectangle [x0, x1) x [y0, y1);
 * everything else in dst is left as it is. Applies to one render. */
GGCUDA_API int ggcuda_set_dirty_rect(ggcuda_ctx* ctx, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1);
/* Background (premultiplied RGBA8) used where GGCUDA_COMPOSITE_OVER is not set. Default transparent. */
GGCUDA_API int ggcuda_set_background(ggcuda_ctx* ctx, const uint8_t rgba_premul[4]);
/* Multi-GPU banding: this context renders only tile rows [y0, y1) (16-px rows). Default: whole canvas. Set it BEFORE
 * adding the scene: ggcuda_add_encoding drops fills and strokes that cannot reach the band. */
GGCUDA_API int ggcuda_set_band(ggcuda_ctx* ctx, uint32_t tile_row0, uint32_t tile_row1);

/* ---- per-draw accumulation: GPUAccelerator.FillPath / StrokePath (accelerator.go:118-128),
 *      VelloAccelerator.FillPath/StrokePath (vello_accelerator.go:197-268).
 *      Paths are in device space (the CTM was applied by gg, context.go:1815,1870);
 *      colour is straight-alpha RGBA8 as extractColorU8 produces (path_convert.go:116-128). ---- */
GGCUDA_API int ggcuda_fill_path(ggcuda_ctx* ctx, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                     const uint8_t rgba_straight[4], int fill_rule);
/* Gradient brushes (gg.LinearGradientBrush, gradient_linear.go:52-66; gg.RadialGradientBrush without focus,
 * gradient_radial.go computeTSimple; colour stops interpolated in linear light, gradient.go:62-131; sampled at pixel centres,
 * software.go:1086-1090). geom: linear x0, y0, x1, y1; radial cx, cy, r0, r1 (device space, two unused). stops: 5 doubles each
 * {offset, r, g, b, a}, straight alpha, any order. extend: gg.ExtendMode. Sweep (gg.SweepGradientBrush, gradient_sweep.go:79-150):
 * cx, cy, start angle, end angle (radians). Radial with a focus off the centre (gradient_radial.go:131-196): cx, cy, r0, r1, fx, fy. */
enum { GGCUDA_GRADIENT_LINEAR = 0, GGCUDA_GRADIENT_RADIAL = 1, GGCUDA_GRADIENT_SWEEP = 2, GGCUDA_GRADIENT_RADIAL_FOCAL = 3 };
enum { GGCUDA_EXTEND_PAD = 0, GGCUDA_EXTEND_REPEAT = 1, GGCUDA_EXTEND_REFLECT = 2 };
GGCUDA_API int ggcuda_fill_path_gradient(ggcuda_ctx* ctx, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                                         int kind, const double geom[6], const double* stops, uint32_t n_stops, int extend, int fill_rule);
GGCUDA_API int ggcuda_stroke_path(ggcuda_ctx* ctx, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                       const uint8_t rgba_straight[4], double width, int cap, int join, double miter_limit);
/* Clip / layer brackets for callers that own a clip stack (scene.Encoding's TagBeginClip /
 * TagPushLayer reach the library through ggcuda_add_encoding; these are the same operations
 * exposed per call). blend_mode is scene.BlendMode (scene/encoding.go:17-48). */
GGCUDA_API int ggcuda_push_clip(ggcuda_ctx* ctx, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords);
GGCUDA_API int ggcuda_push_layer(ggcuda_ctx* ctx, uint32_t blend_mode, float alpha);
GGCUDA_API int ggcuda_pop(ggcuda_ctx* ctx);

/* ---- whole-encoding accumulation: the streams of scene.Encoding (scene/encoding.go:407-444),
 *      consumed with the tag semantics of scene.Renderer.executeEncodingOnTile
 *      (scene/renderer.go:619-813). brushes: 4 doubles (straight RGBA) per brush.
 *      Returns GGCUDA_ERR_UNSUPPORTED for TagImage / TagText (caller falls back). ---- */
GGCUDA_API int ggcuda_add_encoding(ggcuda_ctx* ctx, const uint8_t* tags, size_t n_tags, const float* path_data, size_t n_path_data,
                        const uint32_t* draw_data, size_t n_draw_data, const float* transforms, size_t n_transforms,
                        const double* brushes_rgba, size_t n_brushes);

/* ---- render: GPUAccelerator.Flush (accelerator.go:139, vello_accelerator.go:335-386) ---- */
/* Upload the scene, run the pipeline, read the band back into dst (premultiplied RGBA8,
 * GPURenderTarget.Data/Stride). dst addresses row 0 of the CANVAS; only the band's rows are written. */
GGCUDA_API int ggcuda_flush(ggcuda_ctx* ctx, uint8_t* dst, size_t stride_bytes, uint32_t flags);
/* By default the frame travels through the library's own page-locked staging buffer and nothing of dst is retained.
 * A caller that keeps one pixmap alive (gg keeps one per Context) may page-lock it in place: flushes whose destination
 * lies inside a registered range are DMA'd straight into it, slice by slice while fine rasterisation is still running.
 * The memory must stay allocated AND at the same address until ggcuda_unregister_target / ggcuda_destroy -- from Go that
 * means a runtime.Pinner (or C-allocated pixels), see INTEGRATION.md. */
GGCUDA_API int ggcuda_register_target(ggcuda_ctx* ctx, uint8_t* data, size_t bytes);
GGCUDA_API int ggcuda_unregister_target(ggcuda_ctx* ctx, uint8_t* data);
/* Split form used by benchmarks and multi-GPU callers: upload once, render into device memory.
 * dst_device addresses the first row of this context's BAND. */
GGCUDA_API int ggcuda_upload(ggcuda_ctx* ctx);
GGCUDA_API int ggcuda_render_device(ggcuda_ctx* ctx, void* dst_device, size_t stride_bytes, uint32_t flags);
/* Multi-GPU form (bands assembled over NVLink): besides dst_device the band is stored, by the fine
 * rasterisation kernel itself, at n_mirrors further addresses -- the first row of the same band inside the frames of the
 * other devices (peer-mapped pointers, same stride), or, with multicast != 0, ONE NVSwitch multicast address (n_mirrors
 * == 1) that reaches every device of the group. The caller synchronises the group afterwards (a barrier, not an
 * all-gather). There is no counterpart in the reference (single device). */
GGCUDA_API int ggcuda_render_device_multi(ggcuda_ctx* ctx, void* dst_device, void* const* mirrors, uint32_t n_mirrors, int multicast,
                                          size_t stride_bytes, uint32_t flags);
/* Deferred assembly: copy the band the last render produced (band_device, `bytes` long, 16-byte aligned) into every address of
 * `mirrors` (the band's place in each device's frame -- this device's own frame included --, or ONE NVSwitch multicast address)
 * on `stream`, a stream of the caller's: the copy runs beside the context's next renders, so the receivers' NVLink ingress
 * ((N - 1) bands per frame) is hidden behind the next frame's flatten / binning / coarse. Render into two bands alternately;
 * the context waits before rendering into a band whose broadcast is still reading it. Synchronise `stream` (and the other
 * devices) before reading the assembled frame. */
GGCUDA_API int ggcuda_broadcast_band(ggcuda_ctx* ctx, const void* band_device, void* const* mirrors, uint32_t n_mirrors, int multicast,
                                     size_t bytes, void* stream);

/* Band assembly inside the library ("assembled with one NCCL all-gather"), for hosts that have no
 * collective library of their own (the Go binding). NCCL is resolved at run time (the copy already loaded in the process,
 * else libnccl.so.2); libggcuda.so does not link it. One rank calls ggcuda_comm_unique_id and hands the 128 bytes to the
 * others (any channel); every rank then calls ggcuda_comm_init. ggcuda_all_gather_bands gathers IN PLACE on the context's
 * stream: every rank's band (band_bytes, equal on all ranks) already lies at frame + rank * band_bytes, where
 * ggcuda_render_device put it. ggcuda_sync waits for the context's stream. */
GGCUDA_API int ggcuda_comm_unique_id(uint8_t id[128]);
GGCUDA_API int ggcuda_comm_init(ggcuda_ctx* ctx, int n_ranks, int rank, const uint8_t id[128]);
GGCUDA_API int ggcuda_comm_destroy(ggcuda_ctx* ctx);
GGCUDA_API int ggcuda_all_gather_bands(ggcuda_ctx* ctx, void* frame_device, size_t band_bytes);
GGCUDA_API int ggcuda_sync(ggcuda_ctx* ctx);

/* ---- introspection ---- */
typedef struct {
    uint32_t n_draws, n_paths, n_clips, n_tag_bytes;
    uint32_t n_lines, n_path_tiles, n_seg_counts, n_segments, n_hits, n_ptcl_words, n_spill;
    uint32_t passes;              /* pipeline executions of the last render (> 1 when a buffer had to grow) */
    uint32_t kernel_launches;     /* kernels launched by the last render */
    uint64_t scene_bytes;         /* host->device bytes of the last upload */
    uint64_t device_bytes;        /* device memory currently held */
    float ms_front, ms_binning, ms_coarse, ms_fine;   /* CUDA-event stage times of the last render (timing enabled) */
} ggcuda_stats;
GGCUDA_API int ggcuda_get_stats(ggcuda_ctx* ctx, ggcuda_stats* out);
GGCUDA_API int ggcuda_set_timing(ggcuda_ctx* ctx, int enabled);

/* Copy an intermediate buffer of the last render to host memory (parity tests). Returns bytes
 * written, or the required size when dst == NULL / cap too small (negative on error). */
enum {
    GGCUDA_BUF_SCENE = 0, GGCUDA_BUF_TAG_MONOIDS = 1, GGCUDA_BUF_DRAW_MONOIDS = 2, GGCUDA_BUF_INFO = 3,
    GGCUDA_BUF_CLIP_INPS = 4, GGCUDA_BUF_LINES = 5, GGCUDA_BUF_PATHS = 6, GGCUDA_BUF_TILES = 7,
    GGCUDA_BUF_SEG_START = 8, GGCUDA_BUF_SEGMENTS = 9, GGCUDA_BUF_PTCL_OFF = 10, GGCUDA_BUF_PTCL = 11,
    GGCUDA_BUF_HIT_CNT = 12, GGCUDA_BUF_LAYOUT = 13, GGCUDA_BUF_RESTART = 14
};
GGCUDA_API long long ggcuda_debug_read(ggcuda_ctx* ctx, int which, void* dst, size_t cap_bytes);

/* Host-only scene packing. ggcuda_create(device = -1) gives a context that accumulates scenes but
 * owns no device (every render call fails with GGCUDA_ERR_UNSUPPORTED: there is no CPU fallback).
 * ggcuda_pack_host writes the packed scene exactly as ggcuda_upload would send it (layout13: the 13
 * words of the packed layout) and returns its size in words (also when dst is NULL / too small). */
GGCUDA_API long long ggcuda_pack_host(ggcuda_ctx* ctx, uint32_t* dst, size_t cap_words, uint32_t layout13[13]);

#ifdef __cplusplus
}
#endif
#endif
