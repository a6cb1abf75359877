// gg_b200/csrc/fine.cu -- fine rasterisation: per-tile PTCL replay with analytic area coverage.
//
// Behavioural spec: gg internal/gpu/tilecompute/fine.go:40-289 (fineRasterizeTile, fillPath),
// the CPU twin of tilecompute/shaders/fine.wgsl. One warp owns one 16x16 tile; a lane owns
// 8 consecutive pixels of one row (2 lanes per row), so the per-(segment,row) terms are
// computed twice per row instead of 16 times as in a thread-per-pixel mapping, and every
// lane finishes with two 16-byte RGBA8 stores. Segments of a fill are fetched 32 at a time
// (one per lane, coalesced) and broadcast with shuffles; no shared memory, no block barrier.
#include "pipeline.cuh"

#define FINE_WARPS 4
#define PX 8

struct Seg { float p0x, p0y, p1x, p1y, y_edge; };

__device__ __forceinline__ float signum32(float x) {   // util.go:118-130
    return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : ((__float_as_uint(x) >> 31) ? -1.0f : 1.0f));
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// fine.go:219-289 fillPath for one row (yi) and PX pixels starting at column xb.
__device__ __forceinline__ void fill_row(float* area, const Seg& s, float yi, float xb) {
    float dx = s.p1x - s.p0x, dy_seg = s.p1y - s.p0y;
    float y = s.p0y - yi;
    float y0 = clamp01(y);
    float y1 = clamp01(y + dy_seg);
    float dy = y0 - y1;
    float y_edge = signum32(dx) * clamp01(yi - s.y_edge + 1.0f);
    if (dy != 0.0f) {
        float vec_y_recip = 1.0f / dy_seg;
        float t0 = (y0 - y) * vec_y_recip;
        float t1 = (y1 - y) * vec_y_recip;
        float x0 = s.p0x + t0 * dx;
        float x1 = s.p0x + t1 * dx;
        float xmin0 = fminf(x0, x1);
        float xmax0 = fmaxf(x0, x1);
#pragma unroll
        for (int i = 0; i < PX; i++) {
            float i_f = xb + (float)i;   // absolute column in the tile, as fine.go:259
            float xmin = fminf(xmin0 - i_f, 1.0f) - 1.0e-6f;
            float xmax = xmax0 - i_f;
            float b = fminf(xmax, 1.0f);
            float c = fmaxf(b, 0.0f);
            float d = fmaxf(xmin, 0.0f);
            float a = (b + 0.5f * (d * d - c * c) - xmin) / (xmax - xmin);
            area[i] += y_edge + a * dy;
        }
    } else if (y_edge != 0.0f) {
#pragma unroll
        for (int i = 0; i < PX; i++) area[i] += y_edge;
    }
}

__device__ __forceinline__ uint32_t pack_rgba8(float4 c) {   // fine.wgsl:305-323
    uint32_t r = (uint32_t)(clamp01(c.x) * 255.0f + 0.5f);
    uint32_t g = (uint32_t)(clamp01(c.y) * 255.0f + 0.5f);
    uint32_t b = (uint32_t)(clamp01(c.z) * 255.0f + 0.5f);
    uint32_t a = (uint32_t)(clamp01(c.w) * 255.0f + 0.5f);
    return r | (g << 8) | (b << 16) | (a << 24);
}
__device__ __forceinline__ float4 unpack_rgba8(uint32_t c) {
    return make_float4((float)(c & 0xffu) / 255.0f, (float)((c >> 8) & 0xffu) / 255.0f,
                       (float)((c >> 16) & 0xffu) / 255.0f, (float)((c >> 24) & 0xffu) / 255.0f);
}

__global__ void __launch_bounds__(FINE_WARPS * 32) fine_kernel(GGConfig cfg, const uint32_t* __restrict__ ptcl_off, const uint32_t* __restrict__ ptcl,
                                                               const GGSegment* __restrict__ segments, const uint32_t* __restrict__ spill_off,
                                                               float4* spill, const GGBump* __restrict__ bump, uint8_t* dst, size_t stride) {
    // a stage overflowed its buffer: PTCL / segments are incomplete, the host re-runs the pass with larger buffers
    if (bump->failed || bump->hits > cfg.hits_cap || bump->ptcl_words > cfg.ptcl_cap || bump->segments > cfg.segments_cap) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_global = blockIdx.x * FINE_WARPS + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * FINE_WARPS;
    const uint32_t n_tiles = cfg.width_in_tiles * (cfg.band_y1 - cfg.band_y0);
    const uint32_t row = lane >> 1;
    const uint32_t xb = (lane & 1u) * PX;
    const float yi = (float)row, xbf = (float)xb;
    float4 stack[GG_BLEND_STACK_SPLIT][PX];   // fine.go:58-62: first 4 clip levels local, deeper levels spill

    for (uint32_t T = warp_global; T < n_tiles; T += n_warps) {
        const uint32_t tx = T % cfg.width_in_tiles, ty = T / cfg.width_in_tiles + cfg.band_y0;
        const uint32_t px = tx * GG_TILE_W + xb, py = ty * GG_TILE_H + row;
        const bool row_in = py < cfg.height;
        uint8_t* out = dst + (size_t)(py - cfg.band_y0 * GG_TILE_H) * stride + (size_t)px * 4;
        float4 rgba[PX];
        float area[PX];
        if (cfg.flags & GG_FLAG_BG_FROM_DST) {
#pragma unroll
            for (int i = 0; i < PX; i++) {
                uint32_t c = 0;
                if (row_in && px + i < cfg.width) c = reinterpret_cast<const uint32_t*>(out)[i];
                rgba[i] = unpack_rgba8(c);
            }
        } else {
#pragma unroll
            for (int i = 0; i < PX; i++) rgba[i] = make_float4(cfg.bg[0], cfg.bg[1], cfg.bg[2], cfg.bg[3]);
        }
#pragma unroll
        for (int i = 0; i < PX; i++) area[i] = 0.0f;
        uint32_t clip_depth = 0;
        const uint32_t* cmd = ptcl + ptcl_off[T] + 1;   // word 0 = blend offset (ptcl.go:98)
        const uint32_t sp_off = spill_off[T];
        for (;;) {
            uint32_t tag = *cmd++;
            if (tag == GG_CMD_END) break;
            if (tag == GG_CMD_FILL) {
                uint32_t packed = cmd[0], seg_ix = cmd[1];
                float backdrop = (float)(int32_t)cmd[2];
                cmd += 3;
                uint32_t n = packed >> 1;
#pragma unroll
                for (int i = 0; i < PX; i++) area[i] = backdrop;
                for (uint32_t base = 0; base < n; base += 32) {
                    Seg mine = {0, 0, 0, 0, 1e9f};
                    if (base + lane < n) {
                        const GGSegment* sp = segments + seg_ix + base + lane;
                        mine.p0x = sp->p0x; mine.p0y = sp->p0y; mine.p1x = sp->p1x; mine.p1y = sp->p1y; mine.y_edge = sp->y_edge;
                    }
                    uint32_t cnt = min(32u, n - base);
                    for (uint32_t j = 0; j < cnt; j++) {
                        Seg s;
                        s.p0x = __shfl_sync(0xffffffffu, mine.p0x, j);
                        s.p0y = __shfl_sync(0xffffffffu, mine.p0y, j);
                        s.p1x = __shfl_sync(0xffffffffu, mine.p1x, j);
                        s.p1y = __shfl_sync(0xffffffffu, mine.p1y, j);
                        s.y_edge = __shfl_sync(0xffffffffu, mine.y_edge, j);
                        fill_row(area, s, yi, xbf);
                    }
                }
                if (packed & 1u) {
#pragma unroll
                    for (int i = 0; i < PX; i++) area[i] = fabsf(area[i] - 2.0f * roundf(0.5f * area[i]));   // fine.go:281
                } else {
#pragma unroll
                    for (int i = 0; i < PX; i++) area[i] = fminf(fabsf(area[i]), 1.0f);                      // fine.go:286
                }
            } else if (tag == GG_CMD_SOLID) {
#pragma unroll
                for (int i = 0; i < PX; i++) area[i] = 1.0f;
            } else if (tag == GG_CMD_COLOR) {
                float4 c = unpack_rgba8(*cmd++);
#pragma unroll
                for (int i = 0; i < PX; i++) {   // fine.go:104-123
                    float cov = area[i];
                    float fr = c.x * cov, fg = c.y * cov, fb = c.z * cov, fa = c.w * cov;
                    float inv = 1.0f - fa;
                    rgba[i].x = fmaf(rgba[i].x, inv, fr); rgba[i].y = fmaf(rgba[i].y, inv, fg);
                    rgba[i].z = fmaf(rgba[i].z, inv, fb); rgba[i].w = fmaf(rgba[i].w, inv, fa);
                }
            } else if (tag == GG_CMD_BEGIN_CLIP) {   // fine.go:125-138
                if (clip_depth < GG_BLEND_STACK_SPLIT) {
#pragma unroll
                    for (int i = 0; i < PX; i++) stack[clip_depth][i] = rgba[i];
                } else if (sp_off != 0xffffffffu) {
                    float4* sp = spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane * PX);
#pragma unroll
                    for (int i = 0; i < PX; i++) sp[i] = rgba[i];
                }
                clip_depth++;
#pragma unroll
                for (int i = 0; i < PX; i++) rgba[i] = make_float4(0, 0, 0, 0);
            } else if (tag == GG_CMD_END_CLIP) {     // fine.go:140-180
                float alpha = __uint_as_float(cmd[1]);
                cmd += 2;
                if (clip_depth == 0) continue;
                clip_depth--;
                const float4* saved;
                if (clip_depth < GG_BLEND_STACK_SPLIT) saved = stack[clip_depth];
                else saved = spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane * PX);
#pragma unroll
                for (int i = 0; i < PX; i++) {
                    float scale = area[i] * alpha;
                    float fr = rgba[i].x * scale, fg = rgba[i].y * scale, fb = rgba[i].z * scale, fa = rgba[i].w * scale;
                    float inv = 1.0f - fa;
                    float4 sv = saved[i];
                    rgba[i].x = fmaf(sv.x, inv, fr); rgba[i].y = fmaf(sv.y, inv, fg);
                    rgba[i].z = fmaf(sv.z, inv, fb); rgba[i].w = fmaf(sv.w, inv, fa);
                }
            } else {
                break;   // unknown command: stop (fine.go:182-185)
            }
        }
        if (row_in) {
            uint32_t o[PX];
#pragma unroll
            for (int i = 0; i < PX; i++) o[i] = pack_rgba8(rgba[i]);
            if (px + PX <= cfg.width && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0)) {
                reinterpret_cast<uint4*>(out)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4*>(out)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int i = 0; i < PX; i++) if (px + i < cfg.width) reinterpret_cast<uint32_t*>(out)[i] = o[i];
            }
        }
    }
}

void gg_launch_fine(const GGConfig& cfg, const GGBuffers& b, uint8_t* dst, size_t stride, cudaStream_t s) {
    uint32_t n_tiles = cfg.width_in_tiles * (cfg.band_y1 - cfg.band_y0);
    uint32_t blocks = (n_tiles + FINE_WARPS - 1) / FINE_WARPS;
    uint32_t max_blocks = GG_SM_COUNT * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) return;
    fine_kernel<<<blocks, FINE_WARPS * 32, 0, s>>>(cfg, b.ptcl_off, b.ptcl, b.segments, b.spill_off, b.spill, b.bump, dst, stride);
}
