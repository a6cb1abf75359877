// gg_b200/csrc/common.cuh -- device data layout, strict-f32 helpers and the chunked scan
// primitive shared by every stage of the ggcuda pipeline (sm_100a only).
//
// Layouts follow gg's tilecompute structs (internal/gpu/tilecompute/types.go:13-47) so that
// intermediate buffers can be compared byte-for-byte with the CPU twin.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>

#define GG_TILE_W 16
#define GG_TILE_H 16
#define GG_SCAN_BLOCKS 592         // CTAs of a chunked scan (4 per SM on a 148-SM B200; any device works, the count is only a partition)
#define GG_SCAN_THREADS 256
#define GG_SCAN_ITEMS 4
#define GG_FINE_PARTS 4            // ggcuda_flush runs fine in up to this many row slices so that read-back overlaps it

struct GGLine { uint32_t path_ix; float p0x, p0y, p1x, p1y; };                 // LineSoup, 20 B
struct GGPath { uint32_t bbox[4]; uint32_t tiles; };                           // Path, 20 B
struct GGTile { int32_t backdrop; uint32_t seg_count; };                       // Tile, 8 B (count stays a count; seg start lives in seg_start[])
struct GGSegCount { uint32_t line_ix; uint32_t counts; };                      // SegmentCount, 8 B
struct GGSegment { float p0x, p0y, p1x, p1y, y_edge; };                        // PathSegment, 20 B
struct GGPathMonoid { uint32_t trans_ix, path_seg_ix, path_seg_offset, style_ix, path_ix; };  // pathtag.go:16
struct GGDrawMonoid { uint32_t path_ix, clip_ix, scene_offset, info_offset; };                // draw_leaf.go:17
struct GGClipInp { uint32_t ix; int32_t path_ix; };                            // types.go:99-106

// Per-draw record produced by draw_leaf for coarse (our "info" superset).
// tag: draw tag. parent: draw index of the innermost enclosing BeginClip (-1 = none); for an
// EndClip it is the index of its own BeginClip. a/b: Color -> rgba8 premul (gradient: its index) / bit 0 even-odd, bit 1 gradient;
// BeginClip -> index of the matching EndClip / unused; EndClip -> blend word / alpha bits.
struct GGDrawRec { uint32_t tag; int32_t parent; uint32_t a; uint32_t b; };
// One (draw, tile) hit of coarse's per-tile lists, written by the backdrop pass that already holds the path tile in
// registers: the draw index (the sort key) and everything coarse needs from the path tile, so that the command writer
// reads one 16-byte record instead of walking draw monoid -> path -> tile -> segment start (ncu r1c: that chain was
// coarse's long-scoreboard stall). Hits of layers without geometry: seg_count 0, backdrop 1 (full coverage).
struct __align__(16) GGHit { uint32_t draw; uint32_t seg_count; uint32_t seg_start; int32_t backdrop; };

// Draw / path tags (scene_encode.go:68-86); 0x0C is our MoveTo: it feeds two floats into the
// path-data stream and no segment under the unchanged PathMonoid bit tricks.
#define GG_DRAWTAG_NOP 0u
#define GG_DRAWTAG_COLOR 0x44u
#define GG_DRAWTAG_BEGIN_CLIP 0x9u
#define GG_DRAWTAG_END_CLIP 0x21u
#define GG_DRAWTAG_GRADIENT 0x444u   // DrawTagColor's monoid increments + bit 10: the scene word is a gradient index (ours)
#define GG_PTAG_LINETO 0x09u
#define GG_PTAG_QUADTO 0x0Au
#define GG_PTAG_CUBICTO 0x0Bu
#define GG_PTAG_MOVETO 0x0Cu
#define GG_PTAG_PATH 0x10u
#define GG_PTAG_TRANSFORM 0x20u
#define GG_PTAG_STYLE 0x40u

// PTCL (ptcl.go:17-24)
#define GG_CMD_END 0u
#define GG_CMD_FILL 1u
#define GG_CMD_SOLID 3u
#define GG_CMD_COLOR 5u
#define GG_CMD_GRAD 6u               // {tag, gradient index}: CmdColor with a per-pixel colour (Vello's CmdLinGrad slot; ours)
#define GG_CMD_BEGIN_CLIP 10u
#define GG_CMD_END_CLIP 11u
#define GG_BLEND_STACK_SPLIT 1   // clip levels kept on chip by fine (the reference's BlendStackSplit is 4, ptcl.go:31; not observable)
// Blend word of BeginClip / CmdEndClip: (mix << 8) | compose in Vello/peniko numbering (0x8003 = plain clip).
// Bit 31 is ours: the layer may be dropped from a tile's PTCL when it would enclose no command there
// (set by the host for PushLayer blend modes that leave the backdrop unchanged where the layer is empty).
#define GG_BLEND_ELIDE_EMPTY 0x80000000u
// Bit 30, also ours: an "implicit" layer -- no clip geometry, coverage 1 everywhere (a PushLayer without a clip
// shape whose blend mode cannot change the backdrop where the layer is empty). It owns no tiles; its Begin/End
// reach a tile's command list only through the draws it encloses. Always set together with bit 31.
#define GG_BLEND_IMPLICIT 0x40000000u

// Bump allocators / required sizes, written by the device, read back once per frame.
struct GGBump {
    uint32_t lines;        //  0: total LineSoup produced by flatten
    uint32_t seg_counts;   //  4: SegmentCount entries bump-allocated by path_count (== segments)
    uint32_t path_tiles;   //  8: sum of per-path bbox tiles   } written as one u64 by the
    uint32_t path_rows;    // 12: sum of per-path bbox rows    } packed path-setup scan
    uint32_t hits;         // 16: (draw, tile) hits feeding coarse } one u64 from the
    uint32_t ptcl_words;   // 20: PTCL words laid out              } packed tile-hit scan
    uint32_t segments;     // 24: scan total of tile segment counts
    uint32_t spill;        // 28: blend-spill tile-levels (clip depth > 4)
    uint32_t failed;       // 32: bitmask of stages whose capacity was exceeded
    uint32_t curves;       // 36: curve tags compacted by flatten_classify
    uint32_t esegs;        // 40: Euler-segment records written by flatten_subdivide
    uint32_t sub_cursor;   // 44: next work-list entry flatten_subdivide hands to a lane
    uint32_t emit_cursor;  // 48: next Euler-segment record flatten_eseg_emit hands to a lane
    uint32_t heavy;        // 52: tiles whose command list (after the restart point) is long: fine starts those first, one warp each
    uint32_t coarse_cursor; // 56: next tile coarse hands to a warp (tile costs differ by orders of magnitude)
    uint32_t pad[1];       // 60
    uint32_t fine_cursor[GG_FINE_PARTS];   // 64: next tile (relative to the part's first) fine hands to a warp, one per launch of a frame
    uint32_t pad2[8 - GG_FINE_PARTS];
};
// A tile with more than this many PTCL words left to execute is `heavy`: coarse lists it (behind spill_off[]), fine hands the
// listed tiles out first and singly instead of in pairs (a frame ends when its slowest warp does).
#define GG_FINE_HEAVY_WORDS 96u
#define GG_FAIL_LINES 1u
#define GG_FAIL_TILES 2u
#define GG_FAIL_SEGCOUNTS 4u
#define GG_FAIL_SEGMENTS 8u
#define GG_FAIL_HITS 16u
#define GG_FAIL_PTCL 32u
#define GG_FAIL_SPILL 64u
#define GG_FAIL_ESEGS 128u

// Per-frame configuration (kernel argument, by value).
struct GGConfig {
    uint32_t width, height;                 // canvas in pixels
    uint32_t width_in_tiles, height_in_tiles;
    uint32_t band_y0, band_y1;              // tile rows owned by this device: [y0, y1)
    uint32_t n_tag_bytes, n_tag_words;      // path tag stream (words padded to 256 like PackScene)
    uint32_t n_draws, n_paths, n_clips;
    uint32_t path_tag_base, path_data_base, draw_tag_base, draw_data_base, transform_base, style_base; // word offsets
    uint32_t clip_parent_base;              // word offset of the host-resolved clip-parent array (n_draws words)
    uint32_t n_scene_words;
    uint32_t lines_cap, tiles_cap, rows_cap, seg_counts_cap, segments_cap, hits_cap, ptcl_cap, spill_cap, esegs_cap;
    uint32_t n_implicit, imp_words;         // implicit layers; words per implicit layer in the (layer, tile) bitmap; 0 = no de-duplication
    float bg[4];                            // premultiplied background
    uint32_t flags;
    uint32_t sm_count;                      // multiprocessors of the device (queried at context creation): grids are multiples of it
    uint32_t grad_base, n_grads;            // gradient table (word offset in the scene buffer): 16-word records | stops | ramps
};
#define GG_FLAG_BG_FROM_DST 1u              // composite-over: the scene is rasterised on transparent and source-overed onto the
                                            // destination's pixels with the reference's byte formula (vello_accelerator.go:388-442)
#define GG_FLAG_TARGET_F32 2u               // destination holds premultiplied float4 pixels (16 bytes) instead of RGBA8

// ---------------------------------------------------------------- strict float32 helpers
// The integer stages must reproduce the Go reference's float32 arithmetic: no FMA
// contraction (this TU is also compiled with -fmad=false), IEEE division/sqrt, and
// float->int conversions that behave like Go on amd64 for the values that occur.
__device__ __forceinline__ float f_floor(float x) { return floorf(x); }
__device__ __forceinline__ float f_ceil(float x) { return ceilf(x); }
__device__ __forceinline__ float f_round(float x) { return roundf(x); }          // half away from zero == math.Round
__device__ __forceinline__ float f_min(float a, float b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; }  // util.go:62
__device__ __forceinline__ float f_max(float a, float b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; }  // util.go:76
__device__ __forceinline__ float f_clamp(float x, float lo, float hi) { if (x < lo) return lo; if (x > hi) return hi; return x; }
__device__ __forceinline__ uint32_t f2u(float f) { return (uint32_t)(long long)f; }   // Go uint32(f32) on amd64
__device__ __forceinline__ int32_t f2i(float f) { return (int32_t)f; }
__device__ __forceinline__ uint32_t span_u(float a, float b) {                        // util.go:41-55
    float mx = a, mn = a;
    if (b > mx) mx = b;
    if (b < mn) mn = b;
    float r = f_ceil(mx) - f_floor(mn);
    if (r < 1.0f) r = 1.0f;
    return f2u(r);
}
// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ uint32_t f_ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float f_unord(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ---------------------------------------------------------------- chunked two-launch scan
// n elements are cut into GG_SCAN_BLOCKS contiguous chunks (one per CTA):
//   A) every CTA reduces its chunk              -> partials[b]
//   B) every CTA sums the partials before its own (592 values: two or three per thread and a block reduction -- cheaper
//      than the separate one-CTA launch that scanned them in round 1), re-reads its chunk and writes the exclusive
//      prefix of each element; CTA 0 also writes the grand total.
// Traffic: 2 reads + 1 write per element; deterministic, no atomics, no look-back spinning.
// `n` lives in device memory (it is usually the output of a previous stage).
template <typename T> struct ScanTraits;
template <> struct ScanTraits<uint32_t> {
    __device__ static uint32_t identity() { return 0; }
    __device__ static uint32_t combine(uint32_t a, uint32_t b) { return a + b; }
};
template <> struct ScanTraits<unsigned long long> {
    __device__ static unsigned long long identity() { return 0; }
    __device__ static unsigned long long combine(unsigned long long a, unsigned long long b) { return a + b; }
};
template <> struct ScanTraits<GGPathMonoid> {
    __device__ static GGPathMonoid identity() { return GGPathMonoid{0, 0, 0, 0, 0}; }
    __device__ static GGPathMonoid combine(const GGPathMonoid& a, const GGPathMonoid& b) {   // pathtag.go:66-74
        return GGPathMonoid{a.trans_ix + b.trans_ix, a.path_seg_ix + b.path_seg_ix, a.path_seg_offset + b.path_seg_offset,
                            a.style_ix + b.style_ix, a.path_ix + b.path_ix};
    }
};
template <> struct ScanTraits<GGDrawMonoid> {
    __device__ static GGDrawMonoid identity() { return GGDrawMonoid{0, 0, 0, 0}; }
    __device__ static GGDrawMonoid combine(const GGDrawMonoid& a, const GGDrawMonoid& b) {   // draw_leaf.go:44-51
        return GGDrawMonoid{a.path_ix + b.path_ix, a.clip_ix + b.clip_ix, a.scene_offset + b.scene_offset, a.info_offset + b.info_offset};
    }
};

template <typename T> __device__ __forceinline__ T shfl_up_t(T v, int delta) {
    static_assert(sizeof(T) % 4 == 0, "word sized");
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) d[i] = __shfl_up_sync(0xffffffffu, s[i], delta);
    return r;
}
template <typename T> __device__ __forceinline__ T shfl_idx_t(T v, int lane) {
    T r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(&v);
    uint32_t* d = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], lane);
    return r;
}

// Inclusive scan across the CTA (GG_SCAN_THREADS threads). Returns inclusive value; *block_total gets the sum.
template <typename T> __device__ __forceinline__ T block_inclusive_scan(T v, T* warp_sums /*[8]*/, T* block_total) {
    typedef ScanTraits<T> Tr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = shfl_up_t(v, d);
        if (lane >= d) v = Tr::combine(o, v);
    }
    if (lane == 31) warp_sums[warp] = v;
    __syncthreads();
    T pre = Tr::identity();
    T tot = Tr::identity();
#pragma unroll
    for (int w = 0; w < GG_SCAN_THREADS / 32; w++) {
        T s = warp_sums[w];
        if (w < warp) pre = Tr::combine(pre, s);
        tot = Tr::combine(tot, s);
    }
    __syncthreads();
    *block_total = tot;
    return Tr::combine(pre, v);
}

__device__ __forceinline__ void scan_chunk_range(uint32_t n, uint32_t* begin, uint32_t* end) {
    const uint32_t tile = GG_SCAN_THREADS * GG_SCAN_ITEMS;
    uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + tile - 1) / tile * tile;
    uint64_t b = (uint64_t)blockIdx.x * chunk, e = b + chunk;
    *begin = b < n ? (uint32_t)b : n;
    *end = e < n ? (uint32_t)e : n;
}

// Load: T operator()(uint32_t i) const.   Store: void operator()(uint32_t i, const T& exclusive, const T& value) const.
template <typename T, typename Load>
__global__ void __launch_bounds__(GG_SCAN_THREADS) scan_reduce_kernel(const uint32_t* n_ptr, uint32_t n_cap, Load load, T* partials) {
    typedef ScanTraits<T> Tr;
    __shared__ T warp_sums[GG_SCAN_THREADS / 32];
    uint32_t n = min(*n_ptr, n_cap), b, e;
    scan_chunk_range(n, &b, &e);
    T acc = Tr::identity();
    // keep element order inside each thread's partial irrelevant: all monoids used are commutative sums
    for (uint32_t i = b + threadIdx.x; i < e; i += GG_SCAN_THREADS) acc = Tr::combine(acc, load(i));
    T tot;
    block_inclusive_scan(acc, warp_sums, &tot);
    if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

template <typename T, typename Load, typename Store>
__global__ void __launch_bounds__(GG_SCAN_THREADS) scan_apply_kernel(const uint32_t* n_ptr, uint32_t n_cap, Load load, Store store, const T* partials, T* total_out) {
    typedef ScanTraits<T> Tr;
    __shared__ T warp_sums[GG_SCAN_THREADS / 32];
    uint32_t n = min(*n_ptr, n_cap), b, e;
    scan_chunk_range(n, &b, &e);
    // exclusive prefix of this CTA's chunk = sum of the partials before it (all monoids here are commutative sums);
    // CTA 0 sums them all for the grand total
    T carry;
    {
        const uint32_t upto = (blockIdx.x == 0 && total_out) ? gridDim.x : blockIdx.x;
        T acc = Tr::identity();
        for (uint32_t i = threadIdx.x; i < upto; i += GG_SCAN_THREADS) acc = Tr::combine(acc, partials[i]);
        T tot;
        block_inclusive_scan(acc, warp_sums, &tot);
        if (blockIdx.x == 0) { if (total_out && threadIdx.x == 0) *total_out = tot; carry = Tr::identity(); }
        else carry = tot;
    }
    const uint32_t tile = GG_SCAN_THREADS * GG_SCAN_ITEMS;
    for (uint32_t t0 = b; t0 < e; t0 += tile) {
        // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile
        T v[GG_SCAN_ITEMS];
        T sum = Tr::identity();
#pragma unroll
        for (int k = 0; k < GG_SCAN_ITEMS; k++) {
            uint32_t i = t0 + threadIdx.x * GG_SCAN_ITEMS + k;
            v[k] = i < e ? load(i) : Tr::identity();
            sum = Tr::combine(sum, v[k]);
        }
        T tot;
        T inc = block_inclusive_scan(sum, warp_sums, &tot);
        // exclusive prefix of this thread's first item = carry + (inc "minus" sum): recompute by shuffling
        T prev = shfl_up_t(inc, 1);
        __shared__ T warp_last[GG_SCAN_THREADS / 32];
        if ((threadIdx.x & 31) == 31) warp_last[threadIdx.x >> 5] = inc;
        __syncthreads();
        T excl;
        if (threadIdx.x == 0) excl = Tr::identity();
        else if ((threadIdx.x & 31) == 0) excl = warp_last[(threadIdx.x >> 5) - 1];
        else excl = prev;
        __syncthreads();
        T run = Tr::combine(carry, excl);
#pragma unroll
        for (int k = 0; k < GG_SCAN_ITEMS; k++) {
            uint32_t i = t0 + threadIdx.x * GG_SCAN_ITEMS + k;
            if (i < e) store(i, run, v[k]);
            run = Tr::combine(run, v[k]);
        }
        carry = Tr::combine(carry, tot);
    }
}

// Host helper: launches the two phases on `stream`. `partials` must hold GG_SCAN_BLOCKS elements of T.
template <typename T, typename Load, typename Store>
static inline void gg_scan(cudaStream_t stream, const uint32_t* n_ptr, uint32_t n_cap, Load load, Store store, T* partials, T* total_out) {
    scan_reduce_kernel<T, Load><<<GG_SCAN_BLOCKS, GG_SCAN_THREADS, 0, stream>>>(n_ptr, n_cap, load, partials);
    scan_apply_kernel<T, Load, Store><<<GG_SCAN_BLOCKS, GG_SCAN_THREADS, 0, stream>>>(n_ptr, n_cap, load, store, partials, total_out);
}
