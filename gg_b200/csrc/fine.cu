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

// ---------------------------------------------------------------- layer blending at CmdEndClip
// Blend word = (mix << 8) | compose in Vello/peniko numbering (scene.BlendMode 0-15 -> mix with SrcOver,
// 16-28 -> compose with Normal mix; 0x8003 = clip). The reference's fine stage ignores the word
// (tilecompute/fine.go:164, fine.wgsl:281) and its CPU scene renderer ignores layer blend
// (scene/renderer.go:716-721), so the semantics are those of gg's pixel blend library evaluated in float32 on
// premultiplied colour:  Co = (1-Da) S + (1-Sa) D + Sa Da B(Cs, Cd)   (internal/blend/advanced.go:52-100),
// B() from advanced.go:102-256 and hsl.go:15-121 (luma 0.30/0.59/0.11), Porter-Duff from porter_duff.go:117-216.
__device__ __forceinline__ float lum3(float r, float g, float b) { return 0.30f * r + 0.59f * g + 0.11f * b; }
__device__ __forceinline__ float min3f(float a, float b, float c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
__device__ __forceinline__ float max3f(float a, float b, float c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
__device__ __forceinline__ void clip_color(float& r, float& g, float& b) {   // hsl.go:24-43
    float l = lum3(r, g, b), n = min3f(r, g, b), x = max3f(r, g, b);
    if (n < 0) { r = l + (r - l) * l / (l - n); g = l + (g - l) * l / (l - n); b = l + (b - l) * l / (l - n); }
    if (x > 1) { r = l + (r - l) * (1 - l) / (x - l); g = l + (g - l) * (1 - l) / (x - l); b = l + (b - l) * (1 - l) / (x - l); }
}
__device__ __forceinline__ void set_lum(float& r, float& g, float& b, float l) { float d = l - lum3(r, g, b); r += d; g += d; b += d; clip_color(r, g, b); }
__device__ __forceinline__ void set_sat(float& r, float& g, float& b, float s) {   // hsl.go:53-88 (same tie-breaking order)
    float* mn; float* md; float* mx;
    if (r <= g && g <= b) { mn = &r; md = &g; mx = &b; }
    else if (r <= b && b <= g) { mn = &r; md = &b; mx = &g; }
    else if (b <= r && r <= g) { mn = &b; md = &r; mx = &g; }
    else if (g <= r && r <= b) { mn = &g; md = &r; mx = &b; }
    else if (g <= b && b <= r) { mn = &g; md = &b; mx = &r; }
    else { mn = &b; md = &g; mx = &r; }
    float lo = *mn, mi = *md, hi = *mx;
    if (hi > lo) { *md = ((mi - lo) * s) / (hi - lo); *mx = s; *mn = 0; }
}
__device__ __forceinline__ float sep_mix(uint32_t mix, float s, float d) {
    switch (mix) {
    case 1: return s * d;
    case 2: return 1.0f - (1.0f - s) * (1.0f - d);
    case 3: return d <= 0.5f ? 2.0f * d * s : 1.0f - 2.0f * (1.0f - d) * (1.0f - s);
    case 4: return s < d ? s : d;
    case 5: return s > d ? s : d;
    case 6: { if (s >= 1.0f) return 1.0f; float r = d / (1.0f - s); return r > 1.0f ? 1.0f : r; }
    case 7: { if (s <= 0.0f) return 0.0f; float r = (1.0f - d) / s; return r > 1.0f ? 0.0f : 1.0f - r; }
    case 8: return s <= 0.5f ? 2.0f * s * d : 1.0f - 2.0f * (1.0f - s) * (1.0f - d);
    case 9: {
        if (s <= 0.5f) return d - (1.0f - 2.0f * s) * d * (1.0f - d);
        float dx = d <= 0.25f ? ((16.0f * d - 12.0f) * d + 4.0f) * d : sqrtf(d);
        return d + (2.0f * s - 1.0f) * (dx - d);
    }
    case 10: return fabsf(s - d);
    case 11: return s + d - 2.0f * s * d;
    default: return s;
    }
}
// bg (blend) fg for a mix mode 1..15 (always composed SrcOver)
__device__ __noinline__ float4 blend_mix_px(uint32_t mix, float4 bg, float4 fg) {
    float sa = fg.w, da = bg.w;
    float isa = 1.0f / sa, ida = 1.0f / da;   // callers guarantee sa > 0 and da > 0
    float csr = fg.x * isa, csg = fg.y * isa, csb = fg.z * isa, cdr = bg.x * ida, cdg = bg.y * ida, cdb = bg.z * ida, br, bgc, bb;
    if (mix >= 12) {
        if (mix == 12) { br = csr; bgc = csg; bb = csb; set_sat(br, bgc, bb, max3f(cdr, cdg, cdb) - min3f(cdr, cdg, cdb)); set_lum(br, bgc, bb, lum3(cdr, cdg, cdb)); }
        else if (mix == 13) { br = cdr; bgc = cdg; bb = cdb; set_sat(br, bgc, bb, max3f(csr, csg, csb) - min3f(csr, csg, csb)); set_lum(br, bgc, bb, lum3(cdr, cdg, cdb)); }
        else if (mix == 14) { br = csr; bgc = csg; bb = csb; set_lum(br, bgc, bb, lum3(cdr, cdg, cdb)); }
        else { br = cdr; bgc = cdg; bb = cdb; set_lum(br, bgc, bb, lum3(csr, csg, csb)); }
    } else {
        br = sep_mix(mix, csr, cdr); bgc = sep_mix(mix, csg, cdg); bb = sep_mix(mix, csb, cdb);
    }
    float sada = sa * da;
    float4 o;
    o.x = (1.0f - da) * fg.x + (1.0f - sa) * bg.x + sada * br;
    o.y = (1.0f - da) * fg.y + (1.0f - sa) * bg.y + sada * bgc;
    o.z = (1.0f - da) * fg.z + (1.0f - sa) * bg.z + sada * bb;
    o.w = sa + da * (1.0f - sa);
    return o;
}
// Porter-Duff compose (Normal mix): Fa * S + Fb * D with Fa = a0 + a1 * Da and Fb = b0 + b1 * Sa
// (porter_duff.go:117-216); one table row per compose mode keeps the per-pixel code branch free.
__constant__ float4 COMPOSE_COEF[14] = {
    {0, 0, 0, 0},    // Clear
    {1, 0, 0, 0},    // Copy
    {0, 0, 1, 0},    // Dest
    {1, 0, 1, -1},   // SrcOver
    {1, -1, 1, 0},   // DestOver
    {0, 1, 0, 0},    // SrcIn
    {0, 0, 0, 1},    // DestIn
    {1, -1, 0, 0},   // SrcOut
    {0, 0, 1, -1},   // DestOut
    {0, 1, 1, -1},   // SrcAtop
    {1, -1, 0, 1},   // DestAtop
    {1, -1, 1, -1},  // Xor
    {1, 0, 1, 0},    // Plus (clamped)
    {1, 0, 1, -1},   // PlusLighter: not produced by gg, treated as SrcOver
};
// ---------------------------------------------------------------- PTCL streaming through shared memory (TMA)
// Each warp streams its tile's command list through a private two-slot ring in shared memory: 1-D bulk
// async copies (cp.async.bulk, SASS UBLKCP) complete on a per-slot mbarrier, the next chunk is in flight
// while the current one is decoded, and every command word is then a broadcast LDS instead of a dependent,
// L1-missing global load (the top stall of the first version: 14 % of samples on the tag compare).
#define PTCL_CHUNK 256   // words per ring slot (1 KiB)
#define FINE_SMEM_PER_WARP 8272

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct PtclStream {
    const uint32_t* src;   // tile's PTCL in global memory (16-byte aligned)
    uint32_t* ring;        // [2][PTCL_CHUNK] in shared memory
    uint64_t* bars;        // [2]
    uint32_t len;          // words in this tile's list
    uint32_t loaded_end;   // first word index not yet available
    uint32_t parity;       // bit s = phase parity to wait for on slot s
    uint32_t issued;       // chunks of the current tile handed to the copy engine
    uint32_t lane;

    __device__ __forceinline__ void issue(uint32_t chunk) {
        uint32_t w0 = chunk * PTCL_CHUNK;
        if (w0 >= len) return;
        uint32_t words = min((uint32_t)PTCL_CHUNK, len - w0);
        words = (words + 3u) & ~3u;   // bulk copies move multiples of 16 bytes; lists are padded to 4 words
        issued = chunk + 1;
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this slot precede the async write
            bulk_load(ring + (chunk & 1u) * PTCL_CHUNK, src + w0, words * 4u, bars + (chunk & 1u));
        }
    }
    __device__ __forceinline__ void begin(const uint32_t* s, uint32_t n, uint32_t first_chunk = 0) {
        // drain a chunk that was prefetched for the previous tile but never needed
        while (loaded_end < issued * PTCL_CHUNK) {
            uint32_t slot = (loaded_end / PTCL_CHUNK) & 1u;
            mbar_wait(bars + slot, (parity >> slot) & 1u);
            parity ^= 1u << slot;
            loaded_end += PTCL_CHUNK;
        }
        src = s; len = n; loaded_end = first_chunk * PTCL_CHUNK; issued = first_chunk;
        __syncwarp();
        issue(first_chunk);
        issue(first_chunk + 1);
    }
    // Called once per command with the index of its first word (the longest command is 4 words). The ring holds
    // the chunk the command starts in and the next one; the slot of the chunk BEHIND the command start is refilled
    // with the chunk after next. A command may straddle a chunk boundary, so nothing is refilled on the strength
    // of its later words (refilling when word cmd+3 entered a new chunk overwrote words cmd..cmd+2: long lists,
    // > 1024 hits in a tile, came out wrong).
    __device__ __forceinline__ void ensure(uint32_t cmd) {
        const uint32_t cur = cmd / PTCL_CHUNK;
        while (issued < cur + 2 && issued * PTCL_CHUNK < len) { __syncwarp(); issue(issued); }
        while (cmd + 3 >= loaded_end && loaded_end < issued * PTCL_CHUNK) {
            uint32_t slot = (loaded_end / PTCL_CHUNK) & 1u;
            mbar_wait(bars + slot, (parity >> slot) & 1u);
            parity ^= 1u << slot;
            loaded_end += PTCL_CHUNK;
        }
    }
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return ring[i & (2 * PTCL_CHUNK - 1)]; }
};

// Area of one CmdFill (fine.go:219-276 fillPath) for the 8 pixels of this lane.
//
// The reference walks every (segment, row, pixel) triple. Here the 32 lanes first take one segment each and
// count the rows it crosses, then the warp walks the flattened list of (segment, row) pairs 32 at a time, one
// pair per lane: a lane evaluates the reference's per-pixel trapezoid formula only for the columns its
// segment passes through in that row and records the constant winding step (a == 1, exactly dy) of every
// column to the right as one entry of a per-row suffix table. Both tables live in shared memory and are
// updated with shared-memory float atomics; a prefix sum over the suffix table at the end gives the same
// sums the reference accumulates pixel by pixel (up to float reassociation and the <= 1e-7 the reference
// adds to pixels left of a segment).
__device__ __forceinline__ void fill_area(float* area, float* acc, float* suf, const GGSegment* __restrict__ segs, uint32_t n,
                                          float backdrop, uint32_t lane) {
#pragma unroll
    for (int i = 0; i < PX; i++) acc[lane * PX + i] = 0.0f;
    for (uint32_t i = lane; i < 16 * 17; i += 32) suf[i] = 0.0f;
    __syncwarp();
    for (uint32_t base = 0; base < n; base += 32) {
        float p0x = 0, p0y = 0, dxs = 0, dys = 0;
        int r0 = 0, k = 0;
        if (base + lane < n) {
            const GGSegment* sp = segs + base + lane;
            p0x = sp->p0x; p0y = sp->p0y;
            float p1x = sp->p1x, p1y = sp->p1y, y_edge = sp->y_edge;
            dxs = p1x - p0x; dys = p1y - p0y;
            float ymin = fminf(p0y, p1y), ymax = fmaxf(p0y, p1y);
            r0 = max(0, min(15, (int)floorf(ymin)));
            int r1 = max(r0, min(16, (int)ceilf(ymax)));
            k = (dys != 0.0f) ? r1 - r0 : 0;
            if (y_edge < 16.0f) {   // segment touches the tile's left edge: winding step for the rows below (fine.go:244)
                float sgn = signum32(dxs);
                for (int yi = max(0, (int)floorf(y_edge)); yi < 16; yi++) {
                    float term = sgn * clamp01((float)yi - y_edge + 1.0f);
                    if (term != 0.0f) atomicAdd(&suf[yi * 17], term);
                }
            }
        }
        // exclusive prefix of rows-per-segment across the warp
        int incl = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += o; }
        const int excl = incl - k;
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int q0 = 0; q0 < total; q0 += 32) {
            const int q = q0 + (int)lane;
            // owner = last lane whose exclusive prefix is <= q
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                int cand = lo + step;
                int v = __shfl_sync(0xffffffffu, excl, cand & 31);
                if (v <= q) lo = cand;
            }
            const float sx = __shfl_sync(0xffffffffu, p0x, lo), sy = __shfl_sync(0xffffffffu, p0y, lo);
            const float sdx = __shfl_sync(0xffffffffu, dxs, lo), sdy = __shfl_sync(0xffffffffu, dys, lo);
            const int sr0 = __shfl_sync(0xffffffffu, r0, lo), sex = __shfl_sync(0xffffffffu, excl, lo);
            if (q < total) {
                const int rowq = sr0 + (q - sex);
                const float yi = (float)rowq;
                float y = sy - yi;
                float y0 = clamp01(y);
                float y1 = clamp01(y + sdy);
                float dy = y0 - y1;
                if (dy != 0.0f) {
                    float vec_y_recip = __frcp_rn(sdy);
                    float t0 = (y0 - y) * vec_y_recip;
                    float t1 = (y1 - y) * vec_y_recip;
                    float x0 = sx + t0 * sdx;
                    float x1 = sx + t1 * sdx;
                    float xmin0 = fminf(x0, x1);
                    float xmax0 = fmaxf(x0, x1);
                    int c0 = max(0, (int)floorf(xmin0));
                    int c1 = min(16, (int)ceilf(xmax0));
                    for (int col = c0; col < c1; col++) {
                        float i_f = (float)col;
                        float xmin = fminf(xmin0 - i_f, 1.0f) - 1.0e-6f;
                        float xmax = xmax0 - i_f;
                        float b = fminf(xmax, 1.0f);
                        float c = fmaxf(b, 0.0f);
                        float d = fmaxf(xmin, 0.0f);
                        // numerator evaluated exactly as the reference does (it cancels for near-vertical segments);
                        // the quotient may be 2 ulp off: far below what survives the 8-bit quantisation
                        float a = __fdividef(b + 0.5f * (d * d - c * c) - xmin, xmax - xmin);
                        atomicAdd(&acc[rowq * 16 + col], a * dy);
                    }
                    if (c1 < 16) atomicAdd(&suf[rowq * 17 + max(c1, 0)], dy);
                }
            }
        }
    }
    __syncwarp();
    const uint32_t row = lane >> 1, xb = (lane & 1u) * PX;
    float run = 0.0f;
    float pref[PX];
#pragma unroll
    for (int i = 0; i < PX; i++) { run += suf[row * 17 + xb + i]; pref[i] = run; }
    float left = __shfl_xor_sync(0xffffffffu, run, 1);
    if (!(lane & 1u)) left = 0.0f;
#pragma unroll
    for (int i = 0; i < PX; i++) area[i] = backdrop + acc[lane * PX + i] + (pref[i] + left);
    __syncwarp();
}

__global__ void __launch_bounds__(FINE_WARPS * 32, 6) fine_kernel(GGConfig cfg, const uint32_t* __restrict__ ptcl_off, const uint32_t* __restrict__ ptcl_len,
                                                               const uint32_t* __restrict__ ptcl, const uint32_t* __restrict__ restart_pt,
                                                               const GGSegment* __restrict__ segments, const uint32_t* __restrict__ spill_off,
                                                               float4* spill, GGBump* bump, uint8_t* dst, size_t stride, uint32_t tile0, uint32_t tile1, uint32_t part, GGFineMirrors mir) {
    // a stage overflowed its buffer: PTCL / segments are incomplete, the host re-runs the pass with larger buffers
    if (bump->failed || bump->hits > cfg.hits_cap || bump->ptcl_words > cfg.ptcl_cap || bump->segments > cfg.segments_cap) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t row = lane >> 1;
    const uint32_t xb = (lane & 1u) * PX;
    // per-warp slice of dynamic shared memory (FINE_SMEM_PER_WARP bytes, opt-in above 48 KiB):
    //   [0, 4096)      blend stack level 0: float4 [1][PX][32]
    //   [4096, 6144)   PTCL ring: u32 [2][PTCL_CHUNK]
    //   [6144, 7168)   area accumulators: float [256]
    //   [7168, 8256)   row suffix table: float [16][17]
    //   [8256, 8272)   two mbarriers
    extern __shared__ __align__(128) unsigned char fine_smem[];
    unsigned char* wsm = fine_smem + (size_t)(threadIdx.x >> 5) * FINE_SMEM_PER_WARP;
    float4 (*sstk)[PX][32] = reinterpret_cast<float4 (*)[PX][32]>(wsm);
    float* acc = reinterpret_cast<float*>(wsm + 6144);
    float* suf = reinterpret_cast<float*>(wsm + 7168);
    PtclStream ps;
    ps.ring = reinterpret_cast<uint32_t*>(wsm + 4096); ps.bars = reinterpret_cast<uint64_t*>(wsm + 8256); ps.parity = 0; ps.issued = 0; ps.loaded_end = 0; ps.lane = threadIdx.x & 31;
    if ((threadIdx.x & 31) == 0) {
        mbar_init(ps.bars + 0, 1); mbar_init(ps.bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // Blend stack (fine.go:58-62 keeps 4 levels "in registers" and spills deeper ones; where a level lives is
    // invisible in the output). Levels 0-1 live in shared memory ([level][pixel][lane]: conflict-free 128-bit
    // accesses), deeper levels in the global spill buffer coarse sized for this tile. A local-memory stack
    // pushed ~10 GB of write-through traffic to L2 per 4K frame of the benchmark scene (73 composites per tile).

    // Tiles are handed out from a shared cursor: their cost varies by orders of magnitude (restart points make most of
    // them trivial), a static stride left warps idle behind the heavy ones.
    for (;;) {
        uint32_t T = 0;
        if (lane == 0) T = tile0 + atomicAdd(&bump->fine_cursor[part], 1u);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= tile1) break;
        const uint32_t tx = T % cfg.width_in_tiles, ty = T / cfg.width_in_tiles + cfg.band_y0;
        const uint32_t px = tx * GG_TILE_W + xb, py = ty * GG_TILE_H + row;
        const bool row_in = py < cfg.height;
        uint8_t* out = dst + (size_t)(py - cfg.band_y0 * GG_TILE_H) * stride + (size_t)px * 4;
        float4 rgba[PX];
        float area[PX];
        // coarse found the last command of this tile that overwrites every pixel whatever came before (an opaque
        // solid colour or a backdrop-wiping layer at clip depth 0): start right after it, from that colour
        const uint32_t restart = restart_pt[2 * T];
        if (restart) {
            const float4 c0 = unpack_rgba8(restart_pt[2 * T + 1]);
#pragma unroll
            for (int i = 0; i < PX; i++) rgba[i] = c0;
        } else if (cfg.flags & GG_FLAG_BG_FROM_DST) {
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
        uint32_t cmd = restart ? restart : 1u;   // word 0 = blend offset (ptcl.go:98)
        ps.begin(ptcl + ptcl_off[T], ptcl_len[T], cmd / PTCL_CHUNK);
        const uint32_t sp_off = spill_off[T];
        for (;;) {
            ps.ensure(cmd);
            const uint32_t tag = ps.word(cmd);
            if (tag == GG_CMD_FILL) {
                const uint32_t packed = ps.word(cmd + 1), seg_ix = ps.word(cmd + 2);
                const float backdrop = (float)(int32_t)ps.word(cmd + 3);
                cmd += 4;
                fill_area(area, acc, suf, segments + seg_ix, packed >> 1, backdrop, lane);
                const bool even_odd = packed & 1u;
#pragma unroll
                for (int i = 0; i < PX; i++) {
                    float a = area[i];
                    area[i] = even_odd ? fabsf(a - 2.0f * roundf(0.5f * a))    // fine.go:281
                                       : fminf(fabsf(a), 1.0f);                // fine.go:286
                }
            } else if (tag == GG_CMD_SOLID) {
                cmd += 1;
#pragma unroll
                for (int i = 0; i < PX; i++) area[i] = 1.0f;
            } else if (tag == GG_CMD_COLOR) {
                const float4 c = unpack_rgba8(ps.word(cmd + 1));
                cmd += 2;
#pragma unroll
                for (int i = 0; i < PX; i++) {   // fine.go:104-123
                    float cov = area[i];
                    float fr = c.x * cov, fg = c.y * cov, fb = c.z * cov, fa = c.w * cov;
                    float inv = 1.0f - fa;
                    rgba[i].x = fmaf(rgba[i].x, inv, fr); rgba[i].y = fmaf(rgba[i].y, inv, fg);
                    rgba[i].z = fmaf(rgba[i].z, inv, fb); rgba[i].w = fmaf(rgba[i].w, inv, fa);
                }
            } else if (tag == GG_CMD_BEGIN_CLIP) {   // fine.go:125-138
                cmd += 1;
                float4* slot; uint32_t step;
                if (clip_depth < GG_BLEND_STACK_SPLIT) { slot = &sstk[clip_depth][0][lane]; step = 32; }
                else { slot = spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane * PX); step = 1; }
                if (clip_depth < GG_BLEND_STACK_SPLIT || sp_off != 0xffffffffu) {
#pragma unroll
                    for (int i = 0; i < PX; i++) { slot[i * step] = rgba[i]; }
                }
                clip_depth++;
#pragma unroll
                for (int i = 0; i < PX; i++) rgba[i] = make_float4(0, 0, 0, 0);
            } else if (tag == GG_CMD_END_CLIP) {     // fine.go:140-180
                const uint32_t blend = ps.word(cmd + 1) & 0x3fffffffu;   // bits 30-31 are coarse's layer flags
                const float alpha = __uint_as_float(ps.word(cmd + 2));
                cmd += 3;
                if (clip_depth == 0) continue;
                clip_depth--;
                const float4* slot; uint32_t step;
                if (clip_depth < GG_BLEND_STACK_SPLIT) { slot = &sstk[clip_depth][0][lane]; step = 32; }
                else { slot = spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane * PX); step = 1; }
                const uint32_t mix = (blend >> 8) & 0xffu, compose = blend & 0xffu;
                // Scaling the source by the coverage (fine.go:152-160) equals out = D + cov (blend(S, D) - D) for every mode
                // whose backdrop factor is 1 under a transparent source; the six compose modes that wipe their backdrop
                // are blended at full strength and interpolated instead (a pixel the layer's clip does not cover stays).
                const bool wipe = mix == 0u && (compose == 0u || compose == 1u || compose == 5u || compose == 6u || compose == 7u || compose == 10u);
#pragma unroll
                for (int i = 0; i < PX; i++) {
                    float scale = wipe ? alpha : area[i] * alpha;   // fg = rgba * area * alpha, in place
                    rgba[i].x *= scale; rgba[i].y *= scale; rgba[i].z *= scale; rgba[i].w *= scale;
                }
                if (mix != 0u && mix < 16u) {
                    // trivial pixels first (transparent source -> backdrop, transparent backdrop -> source); only
                    // pixels where both are visible take the (un-premultiply, mix, re-compose) path
#pragma unroll
                    for (int i = 0; i < PX; i++) {
                        const float4 sv = slot[i * step];
                        if (rgba[i].w <= 0.0f) rgba[i] = sv;
                        else if (sv.w > 0.0f) rgba[i] = blend_mix_px(mix, sv, rgba[i]);
                    }
                } else {
                    // Porter-Duff: Fa * S + Fb * D. Normal / clip SrcOver (fine.go:168-179) is the row {1, 0, 1, -1}.
                    const float4 k = COMPOSE_COEF[min(compose, 13u)];
                    const bool plus = compose == 12u;
#pragma unroll
                    for (int i = 0; i < PX; i++) {
                        const float4 sv = slot[i * step];
                        float fa = k.x + k.y * sv.w, fb = k.z + k.w * rgba[i].w;
                        float4 o;
                        o.x = fmaf(fb, sv.x, fa * rgba[i].x); o.y = fmaf(fb, sv.y, fa * rgba[i].y);
                        o.z = fmaf(fb, sv.z, fa * rgba[i].z); o.w = fmaf(fb, sv.w, fa * rgba[i].w);
                        if (plus) { o.x = fminf(o.x, 1.0f); o.y = fminf(o.y, 1.0f); o.z = fminf(o.z, 1.0f); o.w = fminf(o.w, 1.0f); }
                        if (wipe) { const float cv = area[i]; o.x = sv.x + cv * (o.x - sv.x); o.y = sv.y + cv * (o.y - sv.y); o.z = sv.z + cv * (o.z - sv.z); o.w = sv.w + cv * (o.w - sv.w); }
                        rgba[i] = o;
                    }
                }
            } else {
                break;   // CmdEnd, or an unknown command: stop (fine.go:182-185)
            }
        }
        if (row_in) {
            uint32_t o[PX];
#pragma unroll
            for (int i = 0; i < PX; i++) o[i] = pack_rgba8(rgba[i]);
            const size_t off = (size_t)(py - cfg.band_y0 * GG_TILE_H) * stride + (size_t)px * 4;
            if (px + PX <= cfg.width && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0)) {
                reinterpret_cast<uint4*>(out)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4*>(out)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int i = 0; i < PX; i++) if (px + i < cfg.width) reinterpret_cast<uint32_t*>(out)[i] = o[i];
            }
            // Multi-GPU: the band also goes straight into the other devices' frames (peer memory over NVLink / NVSwitch) while
            // the rest of the band is still being rasterised -- the all-gather of bands, fused into the kernel that makes them.
            if (mir.multicast) {   // one multimem store per 16 bytes: the switch replicates it to every device of the group
                uint8_t* m = mir.p[0] + off;
                if (px + PX <= cfg.width && ((reinterpret_cast<uintptr_t>(m) & 15u) == 0)) {
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(m), "f"(__uint_as_float(o[0])), "f"(__uint_as_float(o[1])),
                                 "f"(__uint_as_float(o[2])), "f"(__uint_as_float(o[3])) : "memory");
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(m + 16), "f"(__uint_as_float(o[4])), "f"(__uint_as_float(o[5])),
                                 "f"(__uint_as_float(o[6])), "f"(__uint_as_float(o[7])) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < PX; i++) if (px + i < cfg.width) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(m + 4 * i), "f"(__uint_as_float(o[i])) : "memory");
                }
            } else {
                for (uint32_t k = 0; k < mir.n; k++) {
                    uint8_t* m = mir.p[k] + off;
                    if (px + PX <= cfg.width && ((reinterpret_cast<uintptr_t>(m) & 15u) == 0)) {
                        reinterpret_cast<uint4*>(m)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        reinterpret_cast<uint4*>(m)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int i = 0; i < PX; i++) if (px + i < cfg.width) reinterpret_cast<uint32_t*>(m)[i] = o[i];
                    }
                }
            }
        }
    }
}

void gg_launch_fine(const GGConfig& cfg, const GGBuffers& b, uint8_t* dst, size_t stride, cudaStream_t s, uint32_t row0, uint32_t row1, uint32_t part,
                    const GGFineMirrors& mir) {
    // tile rows [row0, row1) relative to the band; `part` selects the work cursor (each launch of a frame needs its own)
    uint32_t n_tiles = cfg.width_in_tiles * (row1 - row0);
    uint32_t blocks = (n_tiles + FINE_WARPS - 1) / FINE_WARPS;
    uint32_t max_blocks = GG_SM_COUNT * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) return;
    const int smem = FINE_WARPS * FINE_SMEM_PER_WARP;
    cudaFuncSetAttribute(fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device; cheap
    fine_kernel<<<blocks, FINE_WARPS * 32, smem, s>>>(cfg, b.ptcl_off, b.ptcl_len, b.ptcl, b.restart_pt, b.segments, b.spill_off, b.spill, b.bump, dst, stride,
                                                      cfg.width_in_tiles * row0, cfg.width_in_tiles * row1, part, mir);
}
