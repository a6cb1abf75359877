// gg_b200/csrc/api.cu -- C ABI of libggcuda.so (declared in include/ggcuda.h): context,
// device buffers, scene upload, pipeline execution with grow-and-retry sizing, read-back.
//
// Replaces, behind gg.GPUAccelerator: VelloAccelerator.flushLocked/dispatchComputeScene
// (internal/gpu/vello_accelerator.go:335-386, 512-630) and VelloComputeDispatcher
// (internal/gpu/vello_compute.go:734-793 buffer sizing, :1112-1225 dispatch + WaitIdle).
// Unlike the reference (18 buffers sized by worst-case formulas, zero-filled by upload,
// every intermediate read back for diagnostics), buffers persist across frames, are sized
// by what the previous pass measured (GGBump), and nothing but the bump block and the
// finished band leaves the device.
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved at run time (ggcuda_comm_init), libggcuda.so does not link it

#include <algorithm>
#include <chrono>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "../../include/ggcuda.h"
#include "host_scene.h"
#include "pipeline.cuh"
#include "flatten.cuh"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct Ctx {
    int device = 0;
    bool host_only = false;   // device < 0: scene accumulation and packing only, no CUDA call is ever made
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_part[GG_FINE_PARTS] = {}, ev_copied = nullptr;
    std::string err;
    HostScene scene;
    HostScene::Layout layout{};
    uint32_t width = 0, height = 0, band_y0 = 0, band_y1 = 0;
    bool band_set = false;
    uint8_t bg[4] = {0, 0, 0, 0};
    bool uploaded = false;
    // host staging
    uint32_t* h_scene = nullptr; size_t h_scene_words = 0;
    GGBump* h_bump = nullptr;
    uint8_t* h_frame = nullptr; size_t h_frame_bytes = 0;
    // device
    DevBuf scene_d, tag_monoids, draw_monoids, info, clip_inps, draw_recs, line_count, line_off, curve_list, esegs, lines, path_bbox, paths, path_row_off,
        tiles, seg_start, imp_mask, imp_seen, seg_counts, segments, tile_hits, hit_off, hit_cnt, hit_cursor, hits, ptcl_off, ptcl_len, ptcl, restart_pt, spill_off, spill, bump,
        scan_partials, frame_d;
    uint32_t imp_words = 0;
    size_t uploaded_words = 0;   // packed scene + tail + gradient table, as last uploaded
    uint32_t lines_cap = 0, tiles_cap = 0, rows_cap = 0, seg_counts_cap = 0, segments_cap = 0, hits_cap = 0, ptcl_cap = 0, spill_cap = 0, esegs_cap = 0;
    // Read-back targets the caller asked to page-lock (ggcuda_register_target): a flush whose destination lies inside
    // one is DMA'd straight into it, slice by slice while fine is still running; any other goes through the staging buffer.
    struct Pinned { uint8_t* p; size_t bytes; };
    std::vector<Pinned> pinned;
    // Resident scene: the key the caller gave the scene whose pipeline results (segments, PTCL) are in the
    // device buffers; a ggcuda_begin_keyed with the same key, size and band skips ingest, upload and every stage before fine.
    uint64_t resident_key = 0; bool resident_valid = false, reuse = false, keyed = false; uint64_t pending_key = 0;
    uint32_t res_w = 0, res_h = 0, res_y0 = 0, res_y1 = 0;
    bool dirty_set = false; uint32_t dirty[4] = {0, 0, 0, 0};   // pixels: x0, y0, x1, y1 -- fine and read-back cover only these tiles
    int sm_count = 148;
    // NCCL communicator for band assembly inside the library (ggcuda_comm_init)
    ncclComm_t comm = nullptr; int comm_ranks = 0, comm_rank = 0;
    GGFineMirrors mirrors{};   // set for one render by ggcuda_render_device_multi
    ggcuda_stats stats{};
    bool timing = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // One pass into a device target, captured as a CUDA graph and replayed while nothing it was built from changes (kernel
    // arguments are baked into the nodes: configuration, buffer addresses, target, fine ranges). 25 dependent launches cost
    // ~5 us each through the stream; a 512 x 512 frame spent half its time there.
    struct PassKey { GGConfig cfg; GGBuffers b; uint8_t* dst; size_t stride; GGFineRange rg; GGFineMirrors mir; uint32_t reuse; };
    struct PassGraph { PassKey key; cudaGraphExec_t exec = nullptr; uint32_t launches = 0; } graph[2];   // two: a double-buffered target alternates
    uint32_t graph_next = 0;
    // Deferred band broadcast (ggcuda_broadcast_band): the band a broadcast is still reading must not be rendered into
    struct Bcast { const void* band = nullptr; cudaEvent_t done = nullptr; } bcast[2];
    cudaEvent_t ev_rendered = nullptr;
    bool graphs = true;       // GGCUDA_NO_GRAPH=1 turns them off; so does a failed capture
    bool steady = false;      // the uploaded scene has been rendered to completion with the current buffers (GGCUDA_NO_WAIT)
    GGConfig cfg{};
    GGBump last_bump{};
};

thread_local std::string g_create_err;

int fail(Ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(c, GGCUDA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int ensure(Ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.bytes) return 0;
    size_t want = bytes + bytes / 4 + 256;
    if (b.p) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, GGCUDA_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    b.bytes = want;
    return 0;
}
size_t total_device_bytes(Ctx* c) {
    DevBuf* all[] = {&c->scene_d, &c->tag_monoids, &c->draw_monoids, &c->info, &c->clip_inps, &c->draw_recs, &c->line_count, &c->line_off, &c->curve_list, &c->esegs,
                     &c->lines, &c->path_bbox, &c->paths, &c->path_row_off, &c->tiles, &c->seg_start, &c->imp_mask, &c->imp_seen, &c->seg_counts, &c->segments,
                     &c->tile_hits, &c->hit_off, &c->hit_cnt, &c->hit_cursor, &c->hits, &c->ptcl_off, &c->ptcl_len, &c->ptcl, &c->restart_pt, &c->spill_off, &c->spill,
                     &c->bump, &c->scan_partials, &c->frame_d};
    size_t t = 0;
    for (DevBuf* b : all) t += b->bytes;
    return t;
}
void free_all(Ctx* c) {
    DevBuf* all[] = {&c->scene_d, &c->tag_monoids, &c->draw_monoids, &c->info, &c->clip_inps, &c->draw_recs, &c->line_count, &c->line_off, &c->curve_list, &c->esegs,
                     &c->lines, &c->path_bbox, &c->paths, &c->path_row_off, &c->tiles, &c->seg_start, &c->imp_mask, &c->imp_seen, &c->seg_counts, &c->segments,
                     &c->tile_hits, &c->hit_off, &c->hit_cnt, &c->hit_cursor, &c->hits, &c->ptcl_off, &c->ptcl_len, &c->ptcl, &c->restart_pt, &c->spill_off, &c->spill,
                     &c->bump, &c->scan_partials, &c->frame_d};
    for (DevBuf* b : all) { if (b->p) cudaFree(b->p); b->p = nullptr; b->bytes = 0; }
}

uint32_t band_tiles(const Ctx* c) {
    uint32_t wt = (c->width + GG_TILE_W - 1) / GG_TILE_W;
    return wt * (c->band_y1 - c->band_y0);
}

// Upload the accumulated scene (one pinned staging buffer, one H2D copy) and size the
// scene-proportional buffers.
int upload(Ctx* c) {
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context: no device to render on (there is no CPU fallback)");
    CK(cudaSetDevice(c->device));
    if (c->width == 0 || c->height == 0) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_begin was not called");
    c->scene.close_open_clips();
    size_t words = c->scene.packed_words();
    if (words > c->h_scene_words) {
        if (c->h_scene) { CK(cudaStreamSynchronize(c->stream)); cudaFreeHost(c->h_scene); c->h_scene = nullptr; }
        size_t want = words + words / 4 + 1024;
        CK(cudaHostAlloc((void**)&c->h_scene, want * 4, cudaHostAllocDefault));
        c->h_scene_words = want;
    }
    uint32_t ht = (c->height + GG_TILE_H - 1) / GG_TILE_H;
    if (!c->band_set) { c->band_y0 = 0; c->band_y1 = ht; }
    if (c->band_y1 > ht) c->band_y1 = ht;
    if (c->band_y0 > c->band_y1) c->band_y0 = c->band_y1;
    c->scene.pack(c->h_scene, &c->layout, band_tiles(c));
    const HostScene::Layout& L = c->layout;
    int r;
    if ((r = ensure(c, c->scene_d, words * 4))) return r;
    CK(cudaMemcpyAsync(c->scene_d.p, c->h_scene, words * 4, cudaMemcpyHostToDevice, c->stream));
    c->stats.scene_bytes = words * 4;
    c->uploaded_words = words;
    size_t nd = std::max<size_t>(L.n_draws, 1), np = std::max<size_t>(L.n_paths, 1), bt = std::max<size_t>(band_tiles(c), 1);
    if ((r = ensure(c, c->tag_monoids, sizeof(GGPathMonoid) * L.n_tag_words))) return r;
    if ((r = ensure(c, c->draw_monoids, sizeof(GGDrawMonoid) * nd))) return r;
    if ((r = ensure(c, c->info, 4 * nd))) return r;
    if ((r = ensure(c, c->clip_inps, sizeof(GGClipInp) * std::max<size_t>(L.n_clips, 1)))) return r;
    if ((r = ensure(c, c->draw_recs, sizeof(GGDrawRec) * nd))) return r;
    if ((r = ensure(c, c->line_count, 4 * std::max<size_t>(L.n_tag_bytes, 1)))) return r;
    if ((r = ensure(c, c->line_off, 4 * std::max<size_t>(L.n_tag_bytes, 1)))) return r;
    if ((r = ensure(c, c->curve_list, 4 * std::max<size_t>(L.n_tag_bytes, 1)))) return r;
    if ((r = ensure(c, c->path_bbox, 16 * np))) return r;
    if ((r = ensure(c, c->paths, sizeof(GGPath) * np))) return r;
    if ((r = ensure(c, c->path_row_off, 4 * np))) return r;
    if ((r = ensure(c, c->tile_hits, 8 * bt))) return r;
    if ((r = ensure(c, c->hit_off, 4 * bt))) return r;
    if ((r = ensure(c, c->hit_cnt, 4 * bt))) return r;
    if ((r = ensure(c, c->hit_cursor, 4 * bt))) return r;
    if ((r = ensure(c, c->ptcl_off, 4 * bt))) return r;
    if ((r = ensure(c, c->ptcl_len, 4 * bt))) return r;
    if ((r = ensure(c, c->restart_pt, 8 * bt))) return r;
    if ((r = ensure(c, c->spill_off, 8 * bt))) return r;   // + the list of heavy tiles (GG_FINE_HEAVY_WORDS)
    if ((r = ensure(c, c->bump, sizeof(GGBump)))) return r;
    if ((r = ensure(c, c->scan_partials, 32 * GG_SCAN_BLOCKS))) return r;
    // first guesses for the data-dependent buffers; the retry loop corrects them
    c->lines_cap = std::max<uint32_t>(c->lines_cap, 16u * c->scene.n_seg_tags + 1024u);
    c->esegs_cap = std::max<uint32_t>(c->esegs_cap, 4u * c->scene.n_seg_tags + 1024u);
    c->tiles_cap = std::max<uint32_t>(c->tiles_cap, 64u * L.n_paths + 4096u);
    c->rows_cap = 0xffffffffu;
    c->seg_counts_cap = std::max<uint32_t>(c->seg_counts_cap, 2u * c->lines_cap);
    c->segments_cap = std::max<uint32_t>(c->segments_cap, 2u * c->lines_cap);
    c->hits_cap = std::max<uint32_t>(c->hits_cap, c->tiles_cap);
    c->ptcl_cap = std::max<uint32_t>(c->ptcl_cap, 6u * c->hits_cap + 8u * (uint32_t)bt);
    c->uploaded = true;
    c->steady = false;
    return 0;
}

int size_dynamic(Ctx* c) {
    int r;
    if ((r = ensure(c, c->lines, sizeof(GGLine) * (size_t)c->lines_cap))) return r;
    if ((r = ensure(c, c->esegs, sizeof(GGESeg) * (size_t)c->esegs_cap))) return r;
    if ((r = ensure(c, c->tiles, sizeof(GGTile) * (size_t)c->tiles_cap))) return r;
    if ((r = ensure(c, c->seg_start, 4 * (size_t)c->tiles_cap))) return r;
    if ((r = ensure(c, c->imp_mask, (size_t)c->tiles_cap))) return r;
    {   // (implicit layer, tile) bitmap; beyond 256 MiB the hit lists simply keep their duplicates
        size_t words = ((size_t)band_tiles(c) + 31) / 32, bytes = 4 * words * c->scene.n_implicit;
        c->imp_words = (c->scene.n_implicit && bytes <= ((size_t)256 << 20)) ? (uint32_t)words : 0u;
        if (c->imp_words && (r = ensure(c, c->imp_seen, bytes))) return r;
    }
    if ((r = ensure(c, c->seg_counts, sizeof(GGSegCount) * (size_t)c->seg_counts_cap))) return r;
    // + 4: fine fetches segment slices in 16-byte aligned chunks of 4 segments and may read up to 3 past the last one
    if ((r = ensure(c, c->segments, sizeof(GGSegment) * ((size_t)c->segments_cap + 4)))) return r;
    if ((r = ensure(c, c->hits, sizeof(GGHit) * (size_t)c->hits_cap))) return r;
    if ((r = ensure(c, c->ptcl, 4 * (size_t)c->ptcl_cap))) return r;
    if ((r = ensure(c, c->spill, sizeof(float4) * 256 * (size_t)std::max<uint32_t>(c->spill_cap, 1)))) return r;
    return 0;
}

void fill_config(Ctx* c, uint32_t flags) {
    GGConfig& g = c->cfg;
    const HostScene::Layout& L = c->layout;
    g.width = c->width; g.height = c->height;
    g.width_in_tiles = (c->width + GG_TILE_W - 1) / GG_TILE_W;
    g.height_in_tiles = (c->height + GG_TILE_H - 1) / GG_TILE_H;
    g.band_y0 = c->band_y0; g.band_y1 = c->band_y1;
    g.n_tag_bytes = L.n_tag_bytes; g.n_tag_words = L.n_tag_words;
    g.n_draws = L.n_draws; g.n_paths = L.n_paths; g.n_clips = L.n_clips;
    g.path_tag_base = L.path_tag_base; g.path_data_base = L.path_data_base; g.draw_tag_base = L.draw_tag_base;
    g.draw_data_base = L.draw_data_base; g.transform_base = L.transform_base; g.style_base = L.style_base;
    g.clip_parent_base = L.clip_aux_base; g.n_scene_words = L.n_scene_words;
    g.lines_cap = c->lines_cap; g.tiles_cap = c->tiles_cap; g.rows_cap = c->rows_cap; g.seg_counts_cap = c->seg_counts_cap;
    g.segments_cap = c->segments_cap; g.hits_cap = c->hits_cap; g.ptcl_cap = c->ptcl_cap; g.spill_cap = c->spill_cap; g.esegs_cap = c->esegs_cap; g.imp_words = c->imp_words; g.n_implicit = c->scene.n_implicit;
    for (int i = 0; i < 4; i++) g.bg[i] = (float)c->bg[i] / 255.0f;
    g.flags = ((flags & GGCUDA_COMPOSITE_OVER) ? GG_FLAG_BG_FROM_DST : 0u) | ((flags & GGCUDA_TARGET_F32) ? GG_FLAG_TARGET_F32 : 0u);
    g.sm_count = (uint32_t)c->sm_count;
    g.grad_base = L.n_scene_words + 8; g.n_grads = c->scene.n_gradients;
}

GGBuffers buffers(Ctx* c) {
    GGBuffers b;
    b.scene = (uint32_t*)c->scene_d.p; b.tag_monoids = (GGPathMonoid*)c->tag_monoids.p; b.draw_monoids = (GGDrawMonoid*)c->draw_monoids.p;
    b.info = (uint32_t*)c->info.p; b.clip_inps = (GGClipInp*)c->clip_inps.p; b.draw_recs = (GGDrawRec*)c->draw_recs.p;
    b.line_count = (uint32_t*)c->line_count.p; b.line_off = (uint32_t*)c->line_off.p; b.curve_list = (uint32_t*)c->curve_list.p; b.esegs = (GGESeg*)c->esegs.p; b.lines = (GGLine*)c->lines.p;
    b.path_bbox_ord = (uint32_t*)c->path_bbox.p; b.paths = (GGPath*)c->paths.p; b.path_row_off = (uint32_t*)c->path_row_off.p;
    b.tiles = (GGTile*)c->tiles.p; b.seg_start = (uint32_t*)c->seg_start.p; b.imp_mask = (uint8_t*)c->imp_mask.p; b.imp_seen = (uint32_t*)c->imp_seen.p; b.seg_counts = (GGSegCount*)c->seg_counts.p;
    b.segments = (GGSegment*)c->segments.p; b.tile_hits = (unsigned long long*)c->tile_hits.p; b.hit_off = (uint32_t*)c->hit_off.p;
    b.hit_cnt = (uint32_t*)c->hit_cnt.p; b.hit_cursor = (uint32_t*)c->hit_cursor.p; b.hits = (GGHit*)c->hits.p;
    b.ptcl_off = (uint32_t*)c->ptcl_off.p; b.ptcl_len = (uint32_t*)c->ptcl_len.p; b.restart_pt = (uint32_t*)c->restart_pt.p; b.ptcl = (uint32_t*)c->ptcl.p; b.spill_off = (uint32_t*)c->spill_off.p;
    b.spill = (float4*)c->spill.p; b.bump = (GGBump*)c->bump.p; b.scan_partials = c->scan_partials.p;
    return b;
}

// Tile rows (relative to the band) and tile-pair columns fine has to cover: the whole band, or the tiles a dirty rectangle touches.
GGFineRange fine_range(const Ctx* c) {
    GGFineRange rg;
    const uint32_t wt = (c->width + GG_TILE_W - 1) / GG_TILE_W;
    rg.row0 = 0; rg.row1 = c->band_y1 - c->band_y0; rg.px0 = 0; rg.px1 = (wt + 1) / 2;
    if (c->dirty_set) {
        uint32_t tx0 = c->dirty[0] / GG_TILE_W, ty0 = c->dirty[1] / GG_TILE_H;
        uint32_t tx1 = (std::min(c->dirty[2], c->width) + GG_TILE_W - 1) / GG_TILE_W, ty1 = (std::min(c->dirty[3], c->height) + GG_TILE_H - 1) / GG_TILE_H;
        ty0 = std::max(ty0, c->band_y0); ty1 = std::min(ty1, c->band_y1);
        if (tx1 <= tx0 || ty1 <= ty0) { rg.row1 = rg.row0; return rg; }
        rg.row0 = ty0 - c->band_y0; rg.row1 = ty1 - c->band_y0; rg.px0 = tx0 / 2; rg.px1 = (tx1 + 1) / 2;
    }
    return rg;
}

// Run the pipeline into dst_device (band-relative). Re-runs with larger buffers while a stage overflowed.
// host_dst != nullptr (page-locked, row pitch host_stride): fine runs in GG_FINE_PARTS row slices and every slice is copied to
// the host on a second stream while the next one is rasterised.
// A resident scene (ggcuda_begin_keyed found its key) runs fine only: segments and command lists are still in place.
int render(Ctx* c, uint8_t* dst_device, size_t stride, uint32_t flags, uint8_t* host_dst = nullptr, size_t host_stride = 0) {
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context: no device to render on (there is no CPU fallback)");
    CK(cudaSetDevice(c->device));
    const bool reuse = c->reuse && c->resident_valid;
    if (!reuse && !c->uploaded) { int r = upload(c); if (r) return r; }
    c->stats.passes = 0; c->stats.kernel_launches = 0;
    for (auto& e : c->bcast) if (e.band == dst_device && e.done) { CK(cudaStreamWaitEvent(c->stream, e.done, 0)); e.band = nullptr; }
    const size_t px_bytes = (flags & GGCUDA_TARGET_F32) ? 16 : 4;
    for (int attempt = 0; attempt < 12; attempt++) {
        if (!reuse) { int r = size_dynamic(c); if (r) return r; }
        fill_config(c, flags);
        GGBuffers b = buffers(c);
        const GGFineRange full = fine_range(c);
        const uint32_t rows_t = full.row1 - full.row0;
        const uint32_t parts = (host_dst && rows_t >= 4 * GG_FINE_PARTS) ? GG_FINE_PARTS : 1u;
        uint32_t launches = 0;
        // everything one pass puts on the stream
        auto enqueue = [&]() -> int {
            launches = 0;
            if (c->timing) CK(cudaEventRecord(c->ev[0], c->stream));
            if (!reuse) launches += gg_launch_front(c->cfg, b, c->stream);
            if (c->timing) CK(cudaEventRecord(c->ev[1], c->stream));
            if (!reuse) launches += gg_launch_binning(c->cfg, b, c->stream);
            if (c->timing) CK(cudaEventRecord(c->ev[2], c->stream));
            if (!reuse) launches += gg_launch_coarse(c->cfg, b, c->stream);
            else CK(cudaMemsetAsync(&b.bump->fine_cursor[0], 0, sizeof(uint32_t) * GG_FINE_PARTS, c->stream));
            if (c->timing) CK(cudaEventRecord(c->ev[3], c->stream));
            if (!reuse) CK(cudaMemcpyAsync(c->h_bump, c->bump.p, sizeof(GGBump), cudaMemcpyDeviceToHost, c->stream));
            // fine is launched optimistically; if a stage overflowed its inputs are in-bounds garbage and the pass is redone
            const uint32_t col0 = std::min(full.px0 * 2 * GG_TILE_W, c->width), col1 = std::min(full.px1 * 2 * GG_TILE_W, c->width);
            for (uint32_t k = 0; k < parts && rows_t; k++) {
                GGFineRange rg = full;
                rg.row0 = full.row0 + (uint32_t)((uint64_t)rows_t * k / parts); rg.row1 = full.row0 + (uint32_t)((uint64_t)rows_t * (k + 1) / parts);
                gg_launch_fine(c->cfg, b, dst_device, stride, c->stream, rg, k, c->mirrors);
                if (host_dst) {
                    uint32_t y0 = rg.row0 * GG_TILE_H, y1 = std::min(rg.row1 * GG_TILE_H, std::min(c->band_y1 * GG_TILE_H, c->height) - c->band_y0 * GG_TILE_H);
                    if (y1 > y0 && col1 > col0) {
                        CK(cudaEventRecord(c->ev_part[k], c->stream));
                        CK(cudaStreamWaitEvent(c->copy_stream, c->ev_part[k], 0));
                        CK(cudaMemcpy2DAsync(host_dst + (size_t)y0 * host_stride + (size_t)col0 * px_bytes, host_stride,
                                             dst_device + (size_t)y0 * stride + (size_t)col0 * px_bytes, stride, (size_t)(col1 - col0) * px_bytes, y1 - y0,
                                             cudaMemcpyDeviceToHost, c->copy_stream));
                    }
                }
            }
            if (host_dst) {   // the main stream owns the frame buffer again only after the copies
                CK(cudaEventRecord(c->ev_copied, c->copy_stream));
                CK(cudaStreamWaitEvent(c->stream, c->ev_copied, 0));
            }
            if (c->timing) CK(cudaEventRecord(c->ev[4], c->stream));
            return 0;
        };
        bool replayed = false;
        if (c->graphs && !c->timing && !host_dst) {
            Ctx::PassKey key;
            memset(&key, 0, sizeof key);   // padding bytes take part in the comparison
            key.cfg = c->cfg; key.b = b; key.dst = dst_device; key.stride = stride; key.rg = full; key.mir = c->mirrors; key.reuse = reuse ? 1u : 0u;
            Ctx::PassGraph* pg = nullptr;
            for (auto& e : c->graph) if (e.exec && memcmp(&key, &e.key, sizeof key) == 0) pg = &e;
            if (!pg) {
                pg = &c->graph[c->graph_next]; c->graph_next ^= 1u;
                if (pg->exec) { cudaGraphExecDestroy(pg->exec); pg->exec = nullptr; }
                cudaGraph_t g = nullptr;
                bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
                if (ok) {
                    const int r = enqueue();
                    ok = cudaStreamEndCapture(c->stream, &g) == cudaSuccess && r == 0 && g;
                }
                if (ok) ok = cudaGraphInstantiate(&pg->exec, g, 0) == cudaSuccess;
                if (g) cudaGraphDestroy(g);
                if (!ok) { cudaGetLastError(); pg->exec = nullptr; c->graphs = false; }   // this context goes on without graphs
                else { memcpy(&pg->key, &key, sizeof key); pg->launches = launches; }
            }
            if (pg->exec) { CK(cudaGraphLaunch(pg->exec, c->stream)); replayed = true; launches = pg->launches; }
        }
        if (!replayed) { int r = enqueue(); if (r) return r; }
        c->stats.passes++;
        c->stats.kernel_launches += launches + (rows_t ? parts : 0u);
        // the same scene went through these buffers before: every count is known to fit, the caller may run ahead
        if ((flags & GGCUDA_NO_WAIT) && c->steady && !c->timing && !host_dst) return 0;
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
        if (reuse) {
            if (c->timing) { c->stats.ms_front = c->stats.ms_binning = c->stats.ms_coarse = 0; cudaEventElapsedTime(&c->stats.ms_fine, c->ev[3], c->ev[4]); }
            return 0;
        }
        GGBump bm = *c->h_bump;
        c->last_bump = bm;
        bool grow = false;
        auto need = [&](uint32_t& cap, uint64_t required) { if (required > cap) { cap = (uint32_t)std::min<uint64_t>(required + required / 4 + 1024, 0xfffffff0u); grow = true; } };
        need(c->lines_cap, bm.lines);
        need(c->esegs_cap, bm.esegs);
        need(c->tiles_cap, bm.path_tiles);
        need(c->seg_counts_cap, bm.seg_counts);
        need(c->segments_cap, bm.segments);
        need(c->hits_cap, bm.hits);
        need(c->ptcl_cap, bm.ptcl_words);
        need(c->spill_cap, bm.spill);
        if (!grow && bm.failed == 0) {
            ggcuda_stats& s = c->stats;
            s.n_draws = c->layout.n_draws; s.n_paths = c->layout.n_paths; s.n_clips = c->layout.n_clips; s.n_tag_bytes = c->layout.n_tag_bytes;
            s.n_lines = bm.lines; s.n_path_tiles = bm.path_tiles; s.n_seg_counts = bm.seg_counts; s.n_segments = bm.segments;
            s.n_hits = bm.hits; s.n_ptcl_words = bm.ptcl_words; s.n_spill = bm.spill;
            s.device_bytes = total_device_bytes(c);
            if (c->timing) {
                cudaEventElapsedTime(&s.ms_front, c->ev[0], c->ev[1]);
                cudaEventElapsedTime(&s.ms_binning, c->ev[1], c->ev[2]);
                cudaEventElapsedTime(&s.ms_coarse, c->ev[2], c->ev[3]);
                cudaEventElapsedTime(&s.ms_fine, c->ev[3], c->ev[4]);
            }
            // the device now holds this scene's segments and command lists: remember under which key
            c->steady = true;
            c->resident_valid = c->keyed; c->resident_key = c->pending_key;
            c->res_w = c->width; c->res_h = c->height; c->res_y0 = c->band_y0; c->res_y1 = c->band_y1;
            return 0;
        }
        if (!grow) return fail(c, GGCUDA_ERR_CUDA, "pipeline reported overflow without a growable buffer (failed mask " + std::to_string(bm.failed) + ")");
    }
    return fail(c, GGCUDA_ERR_NOMEM, "pipeline buffers did not converge");
}

// After a render: drop the accumulated scene unless the caller keeps it; a dirty rectangle applies to one render only.
void after_render(Ctx* c, uint32_t flags) {
    c->dirty_set = false;
    if (!(flags & GGCUDA_KEEP_SCENE) && !c->reuse) { c->scene.clear(c->width, c->height); c->uploaded = false; }
}

// ---- NCCL, resolved at run time: the copy already in the process (a host that links NCCL itself, torch's) or libnccl.so.2
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi* nccl_api(std::string* why) {
    static NcclApi api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { err = std::string("NCCL not found: ") + dlerror(); }
        else {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
            if (!api.ok) err = "libnccl.so.2 lacks a required symbol";
        }
    }
    if (!api.ok && why) *why = err;
    return api.ok ? &api : nullptr;
}

const Ctx::Pinned* find_pinned(const Ctx* c, const uint8_t* p, size_t bytes) {
    for (const auto& r : c->pinned) if (p >= r.p && p + bytes <= r.p + r.bytes) return &r;
    return nullptr;
}

}  // namespace

// No C++ exception may cross the C ABI (cgo would abort the process): every entry point runs inside GG_TRY / GG_CATCH.
#define GG_TRY try {
#define GG_CATCH(c)                                                                                                   \
    } catch (const std::bad_alloc&) { return fail((c), GGCUDA_ERR_NOMEM, "out of host memory"); }                     \
    catch (const std::exception& e_) { return fail((c), GGCUDA_ERR_INVALID, std::string("internal error: ") + e_.what()); } \
    catch (...) { return fail((c), GGCUDA_ERR_INVALID, "internal error"); }

extern "C" {

int ggcuda_create(int device, uint32_t flags, ggcuda_ctx** out) {
    Ctx* c = nullptr;
    if (!out) return fail(nullptr, GGCUDA_ERR_INVALID, "out is NULL");
    *out = nullptr;
    try {
    if (device < 0) {   // host-only context (scene packing for tests and the CPU baseline)
        c = new (std::nothrow) Ctx();
        if (!c) return fail(nullptr, GGCUDA_ERR_NOMEM, "out of host memory");
        c->host_only = true; c->device = -1;
        c->scene.host_strokes = (flags & GGCUDA_CREATE_HOST_STROKES) != 0;
        *out = reinterpret_cast<ggcuda_ctx*>(c);
        return 0;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(nullptr, GGCUDA_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e)); }
    if (device < 0 || device >= n) return fail(nullptr, GGCUDA_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, GGCUDA_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, GGCUDA_ERR_UNSUPPORTED, "libggcuda is built for sm_100a (B200) only; found compute " + std::to_string(prop.major) + "." + std::to_string(prop.minor));
    c = new (std::nothrow) Ctx();
    if (!c) return fail(nullptr, GGCUDA_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    c->scene.host_strokes = (flags & GGCUDA_CREATE_HOST_STROKES) != 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&c->h_bump, sizeof(GGBump), cudaHostAllocDefault) != cudaSuccess) {
        g_create_err = std::string("context set-up failed: ") + cudaGetErrorString(cudaGetLastError());
        delete c;
        return GGCUDA_ERR_CUDA;
    }
    c->stream = c->own_stream;
    if (const char* ng = getenv("GGCUDA_NO_GRAPH")) c->graphs = !(ng[0] && ng[0] != '0');
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto& ev : c->ev) cudaEventCreate(&ev);
    for (auto& ev : c->ev_part) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_copied, cudaEventDisableTiming);
    *out = reinterpret_cast<ggcuda_ctx*>(c);
    return 0;
    } catch (...) { delete c; return fail(nullptr, GGCUDA_ERR_NOMEM, "context creation failed"); }
}

void ggcuda_destroy(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return;
    if (c->host_only) { delete c; return; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm) { NcclApi* a = nccl_api(nullptr); if (a) a->CommDestroy(c->comm); c->comm = nullptr; }
    for (auto& r : c->pinned) cudaHostUnregister(r.p);
    c->pinned.clear();
    free_all(c);
    if (c->h_scene) cudaFreeHost(c->h_scene);
    for (auto& e : c->graph) if (e.exec) cudaGraphExecDestroy(e.exec);
    for (auto& e : c->bcast) if (e.done) cudaEventDestroy(e.done);
    if (c->ev_rendered) cudaEventDestroy(c->ev_rendered);
    if (c->h_bump) cudaFreeHost(c->h_bump);
    if (c->h_frame) cudaFreeHost(c->h_frame);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->ev_part) if (ev) cudaEventDestroy(ev);
    if (c->ev_copied) cudaEventDestroy(c->ev_copied);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* ggcuda_last_error(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    return c ? c->err.c_str() : g_create_err.c_str();
}

long long ggcuda_pack_host(ggcuda_ctx* h, uint32_t* dst, size_t cap_words, uint32_t layout13[13]) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->width == 0) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_begin was not called");
    c->scene.close_open_clips();
    size_t words = c->scene.packed_words();
    if (!dst || cap_words < words) return (long long)words;
    uint32_t ht = (c->height + GG_TILE_H - 1) / GG_TILE_H;
    uint32_t y0 = c->band_set ? c->band_y0 : 0, y1 = c->band_set ? std::min(c->band_y1, ht) : ht;
    HostScene::Layout L;
    c->scene.pack(dst, &L, ((c->width + GG_TILE_W - 1) / GG_TILE_W) * (y1 > y0 ? y1 - y0 : 0));
    if (layout13) memcpy(layout13, &L, sizeof(uint32_t) * 13);
    return (long long)words;
    GG_CATCH(c)
}

int ggcuda_set_stream(ggcuda_ctx* h, void* s) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
    return 0;
}

static int begin_common(Ctx* c, uint32_t width, uint32_t height) {
    if (width == 0 || height == 0 || width > 65536 || height > 65536) return fail(c, GGCUDA_ERR_INVALID, "bad target size");
    c->width = width; c->height = height;
    c->scene.clear(width, height);
    c->uploaded = false;
    c->reuse = false; c->keyed = false; c->dirty_set = false;
    return 0;
}

int ggcuda_begin(ggcuda_ctx* h, uint32_t width, uint32_t height) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    return begin_common(c, width, height);
    GG_CATCH(c)
}

int ggcuda_begin_keyed(ggcuda_ctx* h, uint32_t width, uint32_t height, uint64_t key, int* resident) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !resident) return c ? fail(c, GGCUDA_ERR_INVALID, "resident is NULL") : GGCUDA_ERR_INVALID;
    GG_TRY
    *resident = 0;
    uint32_t ht = (height + GG_TILE_H - 1) / GG_TILE_H;
    uint32_t y0 = c->band_set ? c->band_y0 : 0, y1 = c->band_set ? std::min(c->band_y1, ht) : ht;
    if (c->resident_valid && c->resident_key == key && c->res_w == width && c->res_h == height && c->res_y0 == y0 && c->res_y1 == y1 && !c->host_only) {
        // same scene as the one whose segments and command lists are on the device: nothing to ingest, upload or bin
        c->width = width; c->height = height;
        c->reuse = true; c->keyed = true; c->pending_key = key; c->dirty_set = false;
        *resident = 1;
        return 0;
    }
    int r = begin_common(c, width, height);
    if (r) return r;
    c->keyed = true; c->pending_key = key; c->resident_valid = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_set_dirty_rect(ggcuda_ctx* h, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    if (x1 < x0 || y1 < y0) return fail(c, GGCUDA_ERR_INVALID, "dirty rectangle reversed");
    c->dirty[0] = x0; c->dirty[1] = y0; c->dirty[2] = x1; c->dirty[3] = y1; c->dirty_set = true;
    return 0;
}

int ggcuda_set_background(ggcuda_ctx* h, const uint8_t rgba[4]) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !rgba) return GGCUDA_ERR_INVALID;
    memcpy(c->bg, rgba, 4);
    return 0;
}

int ggcuda_set_band(ggcuda_ctx* h, uint32_t y0, uint32_t y1) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    if (y1 < y0) return fail(c, GGCUDA_ERR_INVALID, "band rows reversed");
    if (c->scene.n_culled > 0 && !(c->band_set && c->band_y0 == y0 && c->band_y1 == y1))
        return fail(c, GGCUDA_ERR_INVALID, "ggcuda_set_band: paths outside the previous band were dropped when the scene was added; set the band before adding the scene");
    if (!(c->band_set && c->band_y0 == y0 && c->band_y1 == y1)) { c->uploaded = false; c->resident_valid = false; c->reuse = false; }
    c->band_y0 = y0; c->band_y1 = y1; c->band_set = true;
    c->scene.set_cull_band((float)(y0 * GG_TILE_H), (float)(y1 * GG_TILE_H));
    return 0;
}

static const float ID6[6] = {1, 0, 0, 0, 1, 0};
#define GG_NO_REUSE(c) if ((c)->reuse) return fail((c), GGCUDA_ERR_INVALID, "the scene is resident (ggcuda_begin_keyed): nothing may be added")

int ggcuda_fill_path(ggcuda_ctx* h, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                     const uint8_t rgba[4], int fill_rule) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || (!verbs && n_verbs) || (!coords && n_coords) || !rgba) return c ? fail(c, GGCUDA_ERR_INVALID, "null argument") : GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    if (n_verbs == 0) return 0;   // vello_accelerator.go:216-219: empty paths are dropped
    c->scene.begin_path(ID6, fill_rule == GGCUDA_FILL_EVENODD);
    c->scene.add_verbs(verbs, n_verbs, coords, n_coords);
    c->scene.end_path();
    c->scene.draw_color(gg_pack_color_straight(rgba));
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_add_image(ggcuda_ctx* h, uint32_t width, uint32_t height, const uint8_t* premul_rgba, uint32_t* index_out) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !premul_rgba) return c ? fail(c, GGCUDA_ERR_INVALID, "null argument") : GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    const int ix = c->scene.add_image(width, height, premul_rgba);
    if (ix < 0) return fail(c, GGCUDA_ERR_INVALID, "bad image size (1..16384 per side)");
    if (index_out) *index_out = (uint32_t)ix;
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_fill_path_gradient(ggcuda_ctx* h, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                              int kind, const double geom[6], const double* stops, uint32_t n_stops, int extend, int fill_rule) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || (!verbs && n_verbs) || (!coords && n_coords) || !geom || (!stops && n_stops)) return c ? fail(c, GGCUDA_ERR_INVALID, "null argument") : GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    if (kind < GGCUDA_GRADIENT_LINEAR || kind > GGCUDA_GRADIENT_RADIAL_FOCAL) return fail(c, GGCUDA_ERR_UNSUPPORTED, "unknown gradient kind");
    if (extend < 0 || extend > 2 || n_stops > 64) return fail(c, GGCUDA_ERR_INVALID, "bad gradient");
    if (n_verbs == 0) return 0;
    c->scene.begin_path(ID6, fill_rule == GGCUDA_FILL_EVENODD);
    c->scene.add_verbs(verbs, n_verbs, coords, n_coords);
    c->scene.end_path();
    c->scene.draw_gradient(kind, geom, stops, n_stops, extend);
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_stroke_path(ggcuda_ctx* h, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords,
                       const uint8_t rgba[4], double width, int cap, int join, double miter_limit) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || (!verbs && n_verbs) || (!coords && n_coords) || !rgba) return c ? fail(c, GGCUDA_ERR_INVALID, "null argument") : GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    if (n_verbs == 0) return 0;
    std::vector<float> cf(n_coords);
    for (uint32_t i = 0; i < n_coords; i++) cf[i] = (float)coords[i];
    StrokeStyleHost st = {width, miter_limit, cap, join};
    if (c->scene.host_strokes) {
        std::vector<uint8_t> v(verbs, verbs + n_verbs);
        StrokeSink sink;
        gg_stroke_to_fill(v, cf, st, &sink);
        c->scene.begin_path(ID6, false);
        c->scene.append_stroke(sink);
        c->scene.end_path();
    } else {
        c->scene.stroke_path(ID6, verbs, n_verbs, cf.data(), cf.size(), st);
    }
    c->scene.draw_color(gg_pack_color_straight(rgba));
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_push_clip(ggcuda_ctx* h, const uint8_t* verbs, uint32_t n_verbs, const double* coords, uint32_t n_coords) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    if ((!verbs && n_verbs) || (!coords && n_coords)) return fail(c, GGCUDA_ERR_INVALID, "null argument");
    GG_TRY
    GG_NO_REUSE(c);
    c->scene.begin_path(ID6, false);
    if (n_verbs) c->scene.add_verbs(verbs, n_verbs, coords, n_coords);
    c->scene.end_path();
    std::vector<float> cf(coords, coords + (n_verbs ? n_coords : 0));
    c->scene.set_next_clip_bounds(ID6, verbs, n_verbs, cf.data(), cf.size());
    c->scene.begin_clip(0x8003u, 1.0f, 0);
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_push_layer(ggcuda_ctx* h, uint32_t blend_mode, float alpha) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    c->scene.begin_layer(gg_blend_word(blend_mode), alpha);
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_pop(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    if (c->scene.clip_stack.empty()) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_pop: nothing to pop");
    c->scene.end_clip(c->scene.clip_kind.back());
    c->uploaded = false;
    return 0;
    GG_CATCH(c)
}

int ggcuda_add_encoding(ggcuda_ctx* h, const uint8_t* tags, size_t n_tags, const float* path_data, size_t n_path_data,
                        const uint32_t* draw_data, size_t n_draw_data, const float* transforms, size_t n_transforms,
                        const double* brushes, size_t n_brushes) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    GG_NO_REUSE(c);
    if (c->width == 0) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_begin was not called");
    std::string msg;
    int r = c->scene.add_encoding(tags, n_tags, path_data, n_path_data, draw_data, n_draw_data, transforms, n_transforms / 6, brushes, n_brushes, &msg);
    c->uploaded = false;
    if (r) return fail(c, r, msg);
    return 0;
    GG_CATCH(c)
}

int ggcuda_upload(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->reuse) return 0;
    return upload(c);
    GG_CATCH(c)
}

int ggcuda_render_device(ggcuda_ctx* h, void* dst_device, size_t stride, uint32_t flags) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !dst_device) return c ? fail(c, GGCUDA_ERR_INVALID, "dst_device is NULL") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (stride < (size_t)c->width * ((flags & GGCUDA_TARGET_F32) ? 16 : 4)) return fail(c, GGCUDA_ERR_INVALID, "stride smaller than a row");
    if ((flags & GGCUDA_TARGET_F32) && (flags & GGCUDA_COMPOSITE_OVER)) return fail(c, GGCUDA_ERR_INVALID, "composite-over is defined on RGBA8 targets only");
    if ((flags & GGCUDA_TARGET_F32) && ((reinterpret_cast<uintptr_t>(dst_device) | stride) & 15u)) return fail(c, GGCUDA_ERR_INVALID, "float targets must be 16-byte aligned");
    int r = render(c, (uint8_t*)dst_device, stride, flags);   // one fine launch, no host copy
    if (r == 0) after_render(c, flags);
    return r;
    GG_CATCH(c)
}

int ggcuda_render_device_multi(ggcuda_ctx* h, void* dst_device, void* const* mirrors, uint32_t n_mirrors, int multicast, size_t stride, uint32_t flags) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !dst_device || (n_mirrors && !mirrors)) return c ? fail(c, GGCUDA_ERR_INVALID, "null destination") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (n_mirrors > GG_MAX_MIRRORS || (multicast && n_mirrors != 1)) return fail(c, GGCUDA_ERR_INVALID, "bad mirror list");
    if (flags & GGCUDA_TARGET_F32) return fail(c, GGCUDA_ERR_INVALID, "mirrored bands are RGBA8");
    if (stride < (size_t)c->width * 4) return fail(c, GGCUDA_ERR_INVALID, "stride smaller than a row");
    c->mirrors = GGFineMirrors{};
    for (uint32_t i = 0; i < n_mirrors; i++) c->mirrors.p[i] = (uint8_t*)mirrors[i];
    c->mirrors.n = n_mirrors; c->mirrors.multicast = multicast ? 1u : 0u;
    int r = render(c, (uint8_t*)dst_device, stride, flags);
    c->mirrors = GGFineMirrors{};
    if (r == 0) after_render(c, flags);
    return r;
    GG_CATCH(c)
}

// Deferred multi-GPU assembly: the band of the LAST render, copied into every device's frame by a small kernel on `stream`
// (a stream of the caller's, not the context's) while the context's own stream already rasterises the next frame. The
// ingress of an N-way exchange -- (N - 1) bands per device and frame, 232 MB at N = 8 for 4K RGBA8, a quarter of a
// millisecond of NVLink time -- then hides behind flatten, binning and coarse instead of stretching fine (which it does
// when fine stores into the peers itself, ggcuda_render_device_multi). Render into two bands alternately: a band is not
// rendered into again before its broadcast has read it (the context waits if it has to).
int ggcuda_broadcast_band(ggcuda_ctx* h, const void* band_device, void* const* mirrors, uint32_t n_mirrors, int multicast, size_t bytes, void* stream) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !band_device || !mirrors || !n_mirrors || !stream) return c ? fail(c, GGCUDA_ERR_INVALID, "null argument") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context");
    if (n_mirrors > GG_MAX_MIRRORS || (multicast && n_mirrors != 1)) return fail(c, GGCUDA_ERR_INVALID, "bad mirror list");
    if ((reinterpret_cast<uintptr_t>(band_device) & 15u) || (bytes & 15u)) return fail(c, GGCUDA_ERR_INVALID, "band must be 16-byte aligned and sized");
    GGFineMirrors mir{};
    for (uint32_t i = 0; i < n_mirrors; i++) {
        if (reinterpret_cast<uintptr_t>(mirrors[i]) & 15u) return fail(c, GGCUDA_ERR_INVALID, "mirror must be 16-byte aligned");
        mir.p[i] = (uint8_t*)mirrors[i];
    }
    mir.n = n_mirrors; mir.multicast = multicast ? 1u : 0u;
    CK(cudaSetDevice(c->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!c->ev_rendered) CK(cudaEventCreateWithFlags(&c->ev_rendered, cudaEventDisableTiming));
    CK(cudaEventRecord(c->ev_rendered, c->stream));
    CK(cudaStreamWaitEvent(s, c->ev_rendered, 0));
    gg_launch_band_bcast(band_device, mir, bytes, c->sm_count, s);
    CK(cudaGetLastError());
    Ctx::Bcast* slot = nullptr;
    for (auto& e : c->bcast) if (e.band == band_device) slot = &e;            // this band again
    if (!slot) for (auto& e : c->bcast) if (!slot && !e.band) slot = &e;      // a free entry
    if (!slot) { slot = &c->bcast[0]; CK(cudaEventSynchronize(slot->done)); }  // a third band in flight: let the oldest finish
    if (!slot->done) CK(cudaEventCreateWithFlags(&slot->done, cudaEventDisableTiming));
    CK(cudaEventRecord(slot->done, s));
    slot->band = band_device;
    return 0;
    GG_CATCH(c)
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int ggcuda_register_target(ggcuda_ctx* h, uint8_t* data, size_t bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !data || !bytes) return c ? fail(c, GGCUDA_ERR_INVALID, "null target") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context");
    if (find_pinned(c, data, bytes)) return 0;
    CK(cudaSetDevice(c->device));
    cudaError_t e = cudaHostRegister(data, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, GGCUDA_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    c->pinned.push_back(Ctx::Pinned{data, bytes});
    return 0;
    GG_CATCH(c)
}

int ggcuda_unregister_target(ggcuda_ctx* h, uint8_t* data) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !data) return GGCUDA_ERR_INVALID;
    GG_TRY
    for (size_t i = 0; i < c->pinned.size(); i++) {
        if (c->pinned[i].p != data) continue;
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        cudaHostUnregister(data);
        c->pinned.erase(c->pinned.begin() + (long)i);
        return 0;
    }
    return fail(c, GGCUDA_ERR_INVALID, "target was not registered");
    GG_CATCH(c)
}

int ggcuda_flush(ggcuda_ctx* h, uint8_t* dst, size_t stride, uint32_t flags) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    const bool trace = getenv("GGCUDA_TRACE") != nullptr;
    double t0 = trace ? now_ms() : 0, t1 = 0, t2 = 0;
    if (!c || !dst) return c ? fail(c, GGCUDA_ERR_INVALID, "dst is NULL") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (flags & GGCUDA_TARGET_F32) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_flush reads back RGBA8 (GPURenderTarget.Data); float targets are device targets");
    if (stride < (size_t)c->width * 4) return fail(c, GGCUDA_ERR_INVALID, "stride smaller than a row");
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context: no device to render on (there is no CPU fallback)");
    CK(cudaSetDevice(c->device));
    if (!(c->reuse && c->resident_valid) && !c->uploaded) { int r = upload(c); if (r) return r; }
    if (trace) t1 = now_ms();
    uint32_t row0 = c->band_y0 * GG_TILE_H, row1 = std::min(c->band_y1 * GG_TILE_H, c->height);
    if (row1 <= row0) { after_render(c, flags); return 0; }
    size_t rows = row1 - row0, tight = (size_t)c->width * 4, bytes = rows * tight;
    int r;
    if ((r = ensure(c, c->frame_d, bytes))) return r;
    uint8_t* band_dst = dst + (size_t)row0 * stride;
    const size_t band_bytes = (rows - 1) * stride + tight;
    const bool direct = find_pinned(c, band_dst, band_bytes) != nullptr;
    // rows / columns this flush touches (the whole band unless a dirty rectangle was set)
    const GGFineRange rg = fine_range(c);
    const size_t ry0 = (size_t)rg.row0 * GG_TILE_H, ry1 = std::min<size_t>((size_t)rg.row1 * GG_TILE_H, rows);
    const size_t cx0 = std::min<size_t>((size_t)rg.px0 * 2 * GG_TILE_W, c->width), cx1 = std::min<size_t>((size_t)rg.px1 * 2 * GG_TILE_W, c->width);
    const bool any = ry1 > ry0 && cx1 > cx0;
    if (!direct && bytes > c->h_frame_bytes) {
        if (c->h_frame) { CK(cudaStreamSynchronize(c->stream)); cudaFreeHost(c->h_frame); c->h_frame = nullptr; c->h_frame_bytes = 0; }
        CK(cudaHostAlloc((void**)&c->h_frame, bytes, cudaHostAllocDefault));
        c->h_frame_bytes = bytes;
    }
    if ((flags & GGCUDA_COMPOSITE_OVER) && any) {   // the target's pixels go to the device: fine composites the scene over them
        if (direct) {
            CK(cudaMemcpy2DAsync((uint8_t*)c->frame_d.p + ry0 * tight + cx0 * 4, tight, band_dst + ry0 * stride + cx0 * 4, stride, (cx1 - cx0) * 4, ry1 - ry0,
                                 cudaMemcpyHostToDevice, c->stream));
        } else {
            for (size_t y = ry0; y < ry1; y++) memcpy(c->h_frame + y * tight + cx0 * 4, band_dst + y * stride + cx0 * 4, (cx1 - cx0) * 4);
            CK(cudaMemcpy2DAsync((uint8_t*)c->frame_d.p + ry0 * tight + cx0 * 4, tight, c->h_frame + ry0 * tight + cx0 * 4, tight, (cx1 - cx0) * 4, ry1 - ry0,
                                 cudaMemcpyHostToDevice, c->stream));
        }
    }
    // a page-locked target is filled slice by slice while fine is still running (render() queues the copies)
    r = render(c, (uint8_t*)c->frame_d.p, tight, flags, direct ? band_dst : nullptr, stride);
    if (r) return r;
    if (trace) t2 = now_ms();
    if (!direct && any) {
        CK(cudaMemcpy2DAsync(c->h_frame + ry0 * tight + cx0 * 4, tight, (uint8_t*)c->frame_d.p + ry0 * tight + cx0 * 4, tight, (cx1 - cx0) * 4, ry1 - ry0,
                             cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (size_t y = ry0; y < ry1; y++) memcpy(band_dst + y * stride + cx0 * 4, c->h_frame + y * tight + cx0 * 4, (cx1 - cx0) * 4);
    }
    if (trace) fprintf(stderr, "[ggcuda] flush: pack+upload %.2f ms, pipeline %s %.2f ms (%u passes), %s %.2f ms\n", t1 - t0,
                       direct ? "+ overlapped read-back" : "", t2 - t1, c->stats.passes, direct ? "tail" : "staged read-back", now_ms() - t2);
    after_render(c, flags);
    return 0;
    GG_CATCH(c)
}

// ---- band assembly inside the library: one NCCL all-gather, for hosts without torch (the Go binding)
int ggcuda_comm_unique_id(uint8_t id[128]) {
    if (!id) return GGCUDA_ERR_INVALID;
    try {
        std::string why;
        NcclApi* a = nccl_api(&why);
        if (!a) return fail(nullptr, GGCUDA_ERR_UNSUPPORTED, why);
        ncclUniqueId u;
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclResult_t r = a->GetUniqueId(&u);
        if (r != ncclSuccess) return fail(nullptr, GGCUDA_ERR_CUDA, std::string("ncclGetUniqueId: ") + a->GetErrorString(r));
        memcpy(id, &u, 128);
        return 0;
    } catch (...) { return GGCUDA_ERR_NOMEM; }
}

int ggcuda_comm_init(ggcuda_ctx* h, int n_ranks, int rank, const uint8_t id[128]) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !id) return c ? fail(c, GGCUDA_ERR_INVALID, "id is NULL") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->host_only) return fail(c, GGCUDA_ERR_UNSUPPORTED, "host-only context");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(c, GGCUDA_ERR_INVALID, "bad rank");
    std::string why;
    NcclApi* a = nccl_api(&why);
    if (!a) return fail(c, GGCUDA_ERR_UNSUPPORTED, why);
    CK(cudaSetDevice(c->device));
    if (c->comm) { a->CommDestroy(c->comm); c->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = a->CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != ncclSuccess) { c->comm = nullptr; return fail(c, GGCUDA_ERR_CUDA, std::string("ncclCommInitRank: ") + a->GetErrorString(r)); }
    c->comm_ranks = n_ranks; c->comm_rank = rank;
    return 0;
    GG_CATCH(c)
}

int ggcuda_comm_destroy(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->comm) { NcclApi* a = nccl_api(nullptr); if (a) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); a->CommDestroy(c->comm); } c->comm = nullptr; }
    return 0;
    GG_CATCH(c)
}

int ggcuda_all_gather_bands(ggcuda_ctx* h, void* frame_device, size_t band_bytes) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !frame_device) return c ? fail(c, GGCUDA_ERR_INVALID, "frame is NULL") : GGCUDA_ERR_INVALID;
    GG_TRY
    if (!c->comm) return fail(c, GGCUDA_ERR_INVALID, "ggcuda_comm_init was not called");
    NcclApi* a = nccl_api(nullptr);
    CK(cudaSetDevice(c->device));
    // in place: rank r's band already sits at frame + r * band_bytes (fine wrote it there)
    ncclResult_t r = a->AllGather((const uint8_t*)frame_device + (size_t)c->comm_rank * band_bytes, frame_device, band_bytes, ncclUint8, c->comm, c->stream);
    if (r != ncclSuccess) return fail(c, GGCUDA_ERR_CUDA, std::string("ncclAllGather: ") + a->GetErrorString(r));
    return 0;
    GG_CATCH(c)
}

int ggcuda_sync(ggcuda_ctx* h) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    if (c->host_only) return 0;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
    GG_CATCH(c)
}

// scene.Encoding.Hash (scene/encoding.go:752-802): 64-bit FNV-1a, one step per ELEMENT of the tag, path-data, draw-data and
// transform streams (text data: none on this path). The reference leaves the brushes out; a key that decides whether pixels
// may be re-used must not, so the brush colours' bit patterns are folded in behind the reference's value when given.
uint64_t ggcuda_encoding_hash(const uint8_t* tags, size_t n_tags, const float* path_data, size_t n_path_data, const uint32_t* draw_data, size_t n_draw_data,
                              const float* transforms, size_t n_transforms, const double* brushes_rgba, size_t n_brushes) {
    const uint64_t prime = 1099511628211ull;
    uint64_t hsh = 14695981039346656037ull;
    for (size_t i = 0; i < n_tags; i++) { hsh ^= tags[i]; hsh *= prime; }
    for (size_t i = 0; i < n_path_data; i++) { uint32_t b; memcpy(&b, path_data + i, 4); hsh ^= b; hsh *= prime; }
    for (size_t i = 0; i < n_draw_data; i++) { hsh ^= draw_data[i]; hsh *= prime; }
    for (size_t i = 0; i < n_transforms; i++) { uint32_t b; memcpy(&b, transforms + i, 4); hsh ^= b; hsh *= prime; }
    if (brushes_rgba) for (size_t i = 0; i < 4 * n_brushes; i++) { uint64_t b; memcpy(&b, brushes_rgba + i, 8); hsh ^= b; hsh *= prime; }
    return hsh;
}

int ggcuda_get_stats(ggcuda_ctx* h, ggcuda_stats* out) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c || !out) return GGCUDA_ERR_INVALID;
    *out = c->stats;
    return 0;
}

int ggcuda_set_timing(ggcuda_ctx* h, int enabled) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    c->timing = enabled != 0;
    return 0;
}

long long ggcuda_debug_read(ggcuda_ctx* h, int which, void* dst, size_t cap) {
    Ctx* c = reinterpret_cast<Ctx*>(h);
    if (!c) return GGCUDA_ERR_INVALID;
    GG_TRY
    const GGBump& bm = c->last_bump;
    const HostScene::Layout& L = c->layout;
    uint32_t bt = band_tiles(c);
    const void* src = nullptr; size_t bytes = 0;
    switch (which) {
    case GGCUDA_BUF_SCENE: src = c->scene_d.p; bytes = 4 * std::max<size_t>(c->uploaded_words, L.n_scene_words); break;   // + tail + gradient table
    case GGCUDA_BUF_TAG_MONOIDS: src = c->tag_monoids.p; bytes = sizeof(GGPathMonoid) * (size_t)L.n_tag_words; break;
    case GGCUDA_BUF_DRAW_MONOIDS: src = c->draw_monoids.p; bytes = sizeof(GGDrawMonoid) * (size_t)L.n_draws; break;
    case GGCUDA_BUF_INFO: src = c->info.p; bytes = 4 * (size_t)L.n_draws; break;
    case GGCUDA_BUF_CLIP_INPS: src = c->clip_inps.p; bytes = sizeof(GGClipInp) * (size_t)L.n_clips; break;
    case GGCUDA_BUF_LINES: src = c->lines.p; bytes = sizeof(GGLine) * (size_t)bm.lines; break;
    case GGCUDA_BUF_PATHS: src = c->paths.p; bytes = sizeof(GGPath) * (size_t)L.n_paths; break;
    case GGCUDA_BUF_TILES: src = c->tiles.p; bytes = sizeof(GGTile) * (size_t)bm.path_tiles; break;
    case GGCUDA_BUF_SEG_START: src = c->seg_start.p; bytes = 4 * (size_t)bm.path_tiles; break;
    case GGCUDA_BUF_SEGMENTS: src = c->segments.p; bytes = sizeof(GGSegment) * (size_t)bm.segments; break;
    case GGCUDA_BUF_PTCL_OFF: src = c->ptcl_off.p; bytes = 4 * (size_t)bt; break;
    case GGCUDA_BUF_PTCL: src = c->ptcl.p; bytes = 4 * (size_t)bm.ptcl_words; break;
    case GGCUDA_BUF_HIT_CNT: src = c->hit_cnt.p; bytes = 4 * (size_t)bt; break;
    case GGCUDA_BUF_RESTART: src = c->restart_pt.p; bytes = 8 * (size_t)bt; break;
    case GGCUDA_BUF_LAYOUT: {
        if (!dst || cap < sizeof(L)) return (long long)sizeof(L);
        memcpy(dst, &L, sizeof(L));
        return (long long)sizeof(L);
    }
    default: return fail(c, GGCUDA_ERR_INVALID, "unknown debug buffer");
    }
    if (!dst || cap < bytes) return (long long)bytes;
    if (bytes == 0) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(c, GGCUDA_ERR_CUDA, std::string("debug read: ") + cudaGetErrorString(cudaGetLastError()));
    if (which == GGCUDA_BUF_SEG_START) {   // the device array holds range ENDS after path_tiling; report starts (coarse.go:276-285)
        std::vector<GGTile> t(bm.path_tiles);
        if (cudaMemcpy(t.data(), c->tiles.p, sizeof(GGTile) * t.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(c, GGCUDA_ERR_CUDA, "debug read: tiles");
        uint32_t* s = (uint32_t*)dst;
        for (size_t i = 0; i < t.size(); i++) s[i] -= t[i].seg_count;
    }
    return (long long)bytes;
    GG_CATCH(c)
}

}  // extern "C"
