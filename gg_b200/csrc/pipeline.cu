// gg_b200/csrc/pipeline.cu -- every stage of the ggcuda pipeline before fine rasterisation.
// Compiled with -fmad=false: the float32 results of these stages feed integer decisions
// (line counts, tile walks, backdrops) that must match the CPU twin bit-for-bit.
//
// Stage map (reference = gogpu/gg internal/gpu/tilecompute, CPU twin of its WGSL kernels):
//   pathtag scan      pathtag.go:26-121            -> tag_monoids
//   draw scan + leaf  draw_leaf.go:29-151          -> draw_monoids, info, clip_inps, draw_recs
//   clip leaf fix-up  clip_leaf.go:27-56           (matching resolved on the host, applied here)
//   flatten           flatten.go / euler.go / path_convert.go:29-112 -> lines (+ per-path bbox)
//   path setup        coarse.go:169-223            -> paths (tile bbox, tile offset)
//   path_count        path_count.go:11-205         -> tiles (backdrop deltas, counts), seg_counts
//   backdrop          coarse.go:291-299            -> tiles (prefix-summed backdrop) + tile hit histogram
//   seg alloc         coarse.go:276-285            -> seg_start (global exclusive scan of counts)
//   path_tiling       path_tiling.go:11-199        -> segments
//   coarse            coarse.go:322-625, ptcl.go   -> per-tile PTCL
#include "pipeline.cuh"
#include "stroke.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define GG_GRID(blocks_per_sm) (cfg.sm_count * (blocks_per_sm))

// ------------------------------------------------------------------ monoids
__device__ __forceinline__ GGPathMonoid path_monoid_new(uint32_t tag_word) {   // pathtag.go:26-63
    GGPathMonoid m;
    uint32_t point_count = tag_word & 0x03030303u;
    m.path_seg_ix = __popc((point_count * 7) & 0x04040404u);
    m.trans_ix = __popc(tag_word & 0x20202020u);
    uint32_t n_points = point_count + ((tag_word >> 2) & 0x01010101u);
    uint32_t a = n_points + (n_points & (((tag_word >> 3) & 0x01010101u) * 15));
    a += a >> 8;
    a += a >> 16;
    m.path_seg_offset = a & 0xff;
    m.path_ix = __popc(tag_word & 0x10101010u);
    m.style_ix = __popc(tag_word & 0x40404040u);
    return m;
}
__device__ __forceinline__ GGDrawMonoid draw_monoid_new(uint32_t tag) {   // draw_leaf.go:29-41
    GGDrawMonoid m;
    m.path_ix = tag != GG_DRAWTAG_NOP ? 1u : 0u;
    m.clip_ix = tag & 1u;
    m.scene_offset = (tag >> 2) & 0x7u;
    m.info_offset = (tag >> 6) & 0xfu;
    return m;
}

struct LoadTagMonoid {
    const uint32_t* tags;
    __device__ GGPathMonoid operator()(uint32_t i) const { return path_monoid_new(tags[i]); }
};
struct StoreTagMonoid {
    GGPathMonoid* out;
    __device__ void operator()(uint32_t i, const GGPathMonoid& ex, const GGPathMonoid&) const { out[i] = ex; }
};
struct LoadDrawMonoid {
    const uint32_t* tags;
    __device__ GGDrawMonoid operator()(uint32_t i) const { return draw_monoid_new(tags[i]); }
};
struct StoreDrawMonoid {
    GGDrawMonoid* out;
    __device__ void operator()(uint32_t i, const GGDrawMonoid& ex, const GGDrawMonoid&) const { out[i] = ex; }
};

// draw_leaf.go:112-150 (info + clip inputs) fused with the clip_leaf.go:27-56 fix-up. The
// Begin/End matching itself is resolved while the host packs the scene (clip aux words:
// [2d] = enclosing BeginClip or -1, [2d+1] = matching End (for Begin) / Begin (for End)).
__global__ void draw_leaf_kernel(GGConfig cfg, const uint32_t* __restrict__ scene, GGDrawMonoid* dm,
                                 uint32_t* info, GGClipInp* clip_inps, GGDrawRec* recs) {
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < cfg.n_draws; d += gridDim.x * blockDim.x) {
        uint32_t tag = scene[cfg.draw_tag_base + d];
        GGDrawMonoid m = dm[d];
        int32_t parent = (int32_t)scene[cfg.clip_parent_base + 2 * d];
        uint32_t link = scene[cfg.clip_parent_base + 2 * d + 1];
        GGDrawRec r; r.tag = tag; r.parent = parent; r.a = 0; r.b = 0;
        if (tag == GG_DRAWTAG_COLOR || tag == GG_DRAWTAG_GRADIENT) {
            uint32_t rgba = scene[cfg.draw_data_base + m.scene_offset];   // gradient: its index in the gradient table
            info[m.info_offset] = rgba;
            r.a = rgba;
            // fill rule: style word of this path (coarse.go:683-705 indexes styles by path_ix;
            // our encoder emits one style per path marker so the lookup is exact)
            r.b = ((scene[cfg.style_base + GG_STYLE_WORDS * m.path_ix] & 0x02u) ? 1u : 0u) | (tag == GG_DRAWTAG_GRADIENT ? 2u : 0u);
            r.tag = GG_DRAWTAG_COLOR;   // a gradient fill is a colour draw everywhere downstream; coarse reads bit 1 of b
        } else if (tag == GG_DRAWTAG_BEGIN_CLIP) {
            if (m.clip_ix < cfg.n_clips) { clip_inps[m.clip_ix].ix = d; clip_inps[m.clip_ix].path_ix = (int32_t)m.path_ix; }
            r.a = link;
            r.b = scene[cfg.draw_data_base + m.scene_offset];   // blend word (GG_BLEND_ELIDE_EMPTY lives here)
        } else if (tag == GG_DRAWTAG_END_CLIP) {
            if (m.clip_ix < cfg.n_clips) { clip_inps[m.clip_ix].ix = d; clip_inps[m.clip_ix].path_ix = ~(int32_t)d; }
            if (link < cfg.n_draws) {
                GGDrawMonoid bm = dm[link];          // BeginClip monoids are never rewritten
                m.path_ix = bm.path_ix;              // clip_leaf.go:48-51
                m.scene_offset = bm.scene_offset;
                dm[d] = m;
                r.parent = (int32_t)link;
                r.a = scene[cfg.draw_data_base + bm.scene_offset];
                r.b = scene[cfg.draw_data_base + bm.scene_offset + 1];
            }
        }
        recs[d] = r;
    }
}

// ------------------------------------------------------------------ flatten
struct CurveIn { V2 p0, p1, p2, p3; uint32_t path_ix; uint32_t kind; uint32_t tag; uint32_t style_ix; };   // kind: 0 none, 1 line, 3 cubic

__device__ __forceinline__ V2 xform(const float* t, float x, float y) {   // scene/encoding.go:348-350
    return mk(t[0] * x + t[1] * y + t[2], t[3] * x + t[4] * y + t[5]);
}

__device__ __forceinline__ uint32_t tag_byte(const GGConfig& cfg, const uint32_t* __restrict__ scene, uint32_t i) {
    return (scene[cfg.path_tag_base + (i >> 2)] >> ((i & 3u) * 8u)) & 0xffu;
}
// Exclusive PathMonoid at tag byte i: the scanned per-word monoid + the bytes of the word before i.
__device__ __forceinline__ GGPathMonoid tag_monoid_at(const GGConfig& cfg, const uint32_t* __restrict__ scene,
                                                      const GGPathMonoid* __restrict__ tag_monoids, uint32_t i) {
    uint32_t w = scene[cfg.path_tag_base + (i >> 2)];
    uint32_t sh = (i & 3u) * 8u;
    return ScanTraits<GGPathMonoid>::combine(tag_monoids[i >> 2], path_monoid_new(w & ((1u << sh) - 1u)));
}
// Style of the path a tag byte belongs to (style_ix = number of style tags up to it). False for fills.
__device__ __forceinline__ bool load_stroke_style(const GGConfig& cfg, const uint32_t* __restrict__ scene, uint32_t style_ix, StrokeStyle* st) {
    if (style_ix == 0) return false;
    const uint32_t* s = scene + cfg.style_base + GG_STYLE_WORDS * (style_ix - 1);
    uint32_t f = s[0];
    if (!(f & GG_STYLE_STROKE)) return false;
    st->hw = 0.5f * __uint_as_float(s[1]);
    st->miter_limit = __uint_as_float(s[2]);
    st->join = (f >> 2) & 3u;
    st->cap = (f >> 4) & 3u;
    return true;
}

__device__ inline bool load_curve(const GGConfig& cfg, const uint32_t* __restrict__ scene,
                                  const GGPathMonoid* __restrict__ tag_monoids, uint32_t i, CurveIn* c) {
    uint32_t tag = tag_byte(cfg, scene, i);
    uint32_t seg = tag & 3u;
    if (seg == 0) return false;
    GGPathMonoid m = tag_monoid_at(cfg, scene, tag_monoids, i);
    const float* data = reinterpret_cast<const float*>(scene + cfg.path_data_base) + m.path_seg_offset;
    float t[6] = {1, 0, 0, 0, 1, 0};
    if (m.trans_ix > 0) {
        const float* tp = reinterpret_cast<const float*>(scene + cfg.transform_base) + 6 * (m.trans_ix - 1);
#pragma unroll
        for (int k = 0; k < 6; k++) t[k] = tp[k];
    }
    c->path_ix = m.path_ix; c->tag = tag; c->style_ix = m.style_ix;
    V2 a = xform(t, data[-2], data[-1]);
    if (seg == 1) {
        V2 b = xform(t, data[0], data[1]);
        c->p0 = a; c->p3 = b; c->kind = 1;
    } else if (seg == 2) {   // quad -> cubic, path_convert.go:60-72
        V2 ctrl = xform(t, data[0], data[1]);
        V2 end = xform(t, data[2], data[3]);
        const float k = (float)(2.0 / 3.0);
        c->p0 = a;
        c->p1 = mk(a.x + k * (ctrl.x - a.x), a.y + k * (ctrl.y - a.y));
        c->p2 = mk(end.x + k * (ctrl.x - end.x), end.y + k * (ctrl.y - end.y));
        c->p3 = end; c->kind = 3;
    } else {
        c->p0 = a;
        c->p1 = xform(t, data[0], data[1]);
        c->p2 = xform(t, data[2], data[3]);
        c->p3 = xform(t, data[4], data[5]);
        c->kind = 3;
    }
    return true;
}

// Flatten runs as: classify (every tag byte: fill lines are counted on the spot; curve tags and every segment of a
// stroked path are compacted into a dense work list), subdivide (one thread per work item runs the adaptive
// Euler-spiral subdivision -- serial but cheap -- and leaves one GGESeg record per accepted Euler segment plus the
// item's line count), scan of the per-tag line counts, line_emit (fill LineTo tags, trivial) and eseg_emit (one
// thread per RECORD evaluates its points: the float64 transcendentals run ~200 k wide instead of ~45 k curves wide).
// History (ncu): one thread per tag byte running everything kept 5 of 32 lanes busy; one thread per curve running
// count and emit serially was latency bound at ~2 warps per scheduler once strokes were expanded here too.
__device__ __forceinline__ void fold_bbox(uint32_t* path_bbox_ord, uint32_t path_ix, const float* bb) {
    uint32_t* pb = path_bbox_ord + 4 * (size_t)path_ix;
    atomicMin(pb + 0, f_ord(bb[0]));
    atomicMin(pb + 1, f_ord(bb[1]));
    atomicMax(pb + 2, f_ord(bb[2]));
    atomicMax(pb + 3, f_ord(bb[3]));
}

__global__ void __launch_bounds__(256) flatten_classify_kernel(GGConfig cfg, const uint32_t* __restrict__ scene,
                                                               const GGPathMonoid* __restrict__ tag_monoids, uint32_t* line_count,
                                                               uint32_t* curve_list, GGBump* bump) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_round = (cfg.n_tag_bytes + 31u) & ~31u;
    // Multi-GPU bands: a segment whose control points all lie above or below this device's band of tile rows
    // cannot touch its tiles (backdrop only propagates along x inside a tile row), so it is dropped here and
    // the flatten work scales with the band, not with the scene.
    const bool banded = cfg.band_y0 > 0 || cfg.band_y1 < cfg.height_in_tiles;
    const float band_lo = (float)(cfg.band_y0 * GG_TILE_H), band_hi = (float)(cfg.band_y1 * GG_TILE_H);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool is_curve = false;   // goes to the dense work list: curves, and every segment of a stroked path
        if (i < cfg.n_tag_bytes) {
            uint32_t seg = tag_byte(cfg, scene, i) & 3u;
            uint32_t n = 0;
            if (seg != 0) {
                StrokeStyle st;
                const bool stroke = load_stroke_style(cfg, scene, tag_monoid_at(cfg, scene, tag_monoids, i).style_ix, &st);
                is_curve = seg >= 2 || stroke;
                if (seg == 1 || banded) {
                    CurveIn c;
                    load_curve(cfg, scene, tag_monoids, i, &c);
                    if (seg == 1 && !stroke) n = veq(c.p0, c.p3) ? 0u : 1u;   // path_convert.go:55
                    if (banded) {
                        float ymin = fminf(c.p0.y, c.p3.y), ymax = fmaxf(c.p0.y, c.p3.y);
                        if (seg >= 2) { ymin = fminf(ymin, fminf(c.p1.y, c.p2.y)); ymax = fmaxf(ymax, fmaxf(c.p1.y, c.p2.y)); }
                        if (stroke) {   // the outline reaches a half width (miter tips, square caps: a bit more) beyond the centre line
                            float margin = st.hw * fmaxf(st.join == 0u ? st.miter_limit : 1.0f, 1.5f);
                            ymin -= margin; ymax += margin;
                        }
                        if (ymax < band_lo || ymin > band_hi) { n = 0; is_curve = false; }
                    }
                }
            }
            if (!is_curve) line_count[i] = n;
        }
        uint32_t m = __ballot_sync(0xffffffffu, is_curve);
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&bump->curves, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (is_curve) curve_list[base + __popc(m & ((1u << lane) - 1u))] = i;
        }
    }
}

struct ESegSink {
    GGESeg* esegs; uint32_t cap; GGBump* bump;
    uint32_t tag_ix, path_ix, flags, lines_per_piece, line_rel, prev;
    __device__ void operator()(GGESeg& r) {
        uint32_t ix = atomicAdd(&bump->esegs, 1u);
        r.tag_ix = tag_ix; r.path_ix = path_ix; r.line_rel = line_rel; r.prev = prev; r.flags |= flags;
        if (ix < cap) esegs[ix] = r; else atomicOr(&bump->failed, GG_FAIL_ESEGS);
        prev = ix;
        line_rel += lines_per_piece * (uint32_t)r.n;
    }
};

// Lines of the join to the next segment / of the end cap, written (EMIT) or counted after a stroked segment's sides.
template <bool EMIT>
__device__ inline void stroke_tail(LineOut& o, const GGConfig& cfg, const uint32_t* __restrict__ scene, const GGPathMonoid* __restrict__ tag_monoids,
                                   uint32_t i, const CurveIn& c, V2 ne, const StrokeStyle& st) {
    CurveIn nx;
    if (i + 1 < cfg.n_tag_bytes && load_curve(cfg, scene, tag_monoids, i + 1, &nx))
        stroke_join<EMIT>(o, c.p3, ne, stroke_normal(seg_start_tangent(nx.p0, nx.p1, nx.p2, nx.p3, nx.kind), st.hw), st);
    else
        stroke_cap<EMIT>(o, c.p3, ne, st);
}

__global__ void __launch_bounds__(128, 6) flatten_subdivide_kernel(GGConfig cfg, const uint32_t* __restrict__ scene,
                                                                const GGPathMonoid* __restrict__ tag_monoids,
                                                                const uint32_t* __restrict__ curve_list, uint32_t* line_count,
                                                                GGESeg* esegs, GGBump* bump) {
    const uint32_t n = bump->curves;
    // per-lane curve in flight
    SubdivState ss;
    ESegSink sink; sink.esegs = esegs; sink.cap = cfg.esegs_cap; sink.bump = bump;
    CurveIn c;
    StrokeStyle st;
    uint32_t i = 0;
    bool active = false, stroke = false, finish = false, exhausted = false;
    for (;;) {
        if (!active) {
            if (finish) {   // the curve this lane just finished: a stroked one still owes the lines of its join / cap
                uint32_t lines = sink.line_rel;
                if (stroke) {
                    LineOut o; o.out = nullptr; o.n = 0; o.cap = 0; o.path_ix = c.path_ix;
                    V2 ne = stroke_normal(seg_end_tangent(c.p0, c.p1, c.p2, c.p3, c.kind), st.hw);
                    stroke_tail<false>(o, cfg, scene, tag_monoids, i, c, ne, st);
                    lines += o.n;
                }
                line_count[i] = lines;
                finish = false;
            }
            uint32_t k = exhausted ? n : atomicAdd(&bump->sub_cursor, 1u);
            if (k >= n) {
                exhausted = true;
            } else {
                i = curve_list[k];
                load_curve(cfg, scene, tag_monoids, i, &c);
                sink.tag_ix = i; sink.path_ix = c.path_ix; sink.flags = 0; sink.lines_per_piece = 1; sink.line_rel = 0; sink.prev = 0xffffffffu;
                stroke = load_stroke_style(cfg, scene, c.style_ix, &st);
                if (!stroke) {
                    active = subdiv_begin(ss, c.p0, c.p1, c.p2, c.p3, 0.0f);
                    if (!active) line_count[i] = 0;
                } else {
                    sink.flags = GG_ESEG_STROKE; sink.lines_per_piece = 2;
                    GGESeg r;
                    r.p0 = c.p0; r.p1 = c.p3; r.th0 = r.k0 = r.k1 = 0; r.ch = 1; r.chord_len = 0; r.a = r.b = r.integral = r.int0 = 0;
                    if (c.tag & GG_PTAG_MARKER) {
                        // copy of the subpath's first segment: it exists so that the last segment can read the tangent of a
                        // closing join; after a marker MoveTo (open subpath) it draws the start cap
                        LineOut o; o.out = nullptr; o.n = 0; o.cap = 0; o.path_ix = c.path_ix;
                        if (i > 0 && tag_byte(cfg, scene, i - 1) == GG_PTAG_MARKER_MOVE) {
                            V2 ns = stroke_normal(seg_start_tangent(c.p0, c.p1, c.p2, c.p3, c.kind), st.hw);
                            stroke_cap<false>(o, c.p0, mk(-ns.x, -ns.y), st);
                            r.n = 0; r.flags = GG_ESEG_CAP;
                            sink(r);
                        }
                        line_count[i] = o.n;
                    } else if (c.kind == 1) {
                        r.n = 1; r.flags = GG_ESEG_LINE | GG_ESEG_LAST; sink(r);
                        finish = true;    // join / cap lines are counted on the next round
                    } else {
                        active = subdiv_begin(ss, c.p0, c.p1, c.p2, c.p3, st.hw);
                        finish = !active;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted && !active && !finish)) break;
        if (active) {
            GGESeg r;
            int res = subdiv_step(ss, r);
            if (res == 1) sink(r);
            else if (res == 2) { active = false; finish = true; }
        }
    }
}

__global__ void __launch_bounds__(256) flatten_line_emit_kernel(GGConfig cfg, const uint32_t* __restrict__ scene,
                                                                const GGPathMonoid* __restrict__ tag_monoids,
                                                                const uint32_t* __restrict__ line_count, const uint32_t* __restrict__ line_off,
                                                                GGLine* lines, uint32_t* path_bbox_ord, GGBump* bump) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cfg.n_tag_bytes; i += gridDim.x * blockDim.x) {
        uint32_t w = scene[cfg.path_tag_base + (i >> 2)];
        if (((w >> ((i & 3u) * 8u)) & 3u) != 1u || line_count[i] == 0) continue;
        uint32_t off = line_off[i];
        if (off + 1 > cfg.lines_cap) { atomicOr(&bump->failed, GG_FAIL_LINES); continue; }
        CurveIn c;
        load_curve(cfg, scene, tag_monoids, i, &c);
        StrokeStyle st;
        if (load_stroke_style(cfg, scene, c.style_ix, &st)) continue;   // stroked segments are emitted from the work list
        GGLine l; l.path_ix = c.path_ix; l.p0x = c.p0.x; l.p0y = c.p0.y; l.p1x = c.p3.x; l.p1y = c.p3.y;
        lines[off] = l;
        float bb[4] = {fminf(c.p0.x, c.p3.x), fminf(c.p0.y, c.p3.y), fmaxf(c.p0.x, c.p3.x), fmaxf(c.p0.y, c.p3.y)};
        fold_bbox(path_bbox_ord, c.path_ix, bb);
    }
}

// One lane holds one record at a time and takes the next one from a shared cursor when its own is finished; a step
// evaluates ONE point of an Euler segment (the float64 transcendentals), whichever record and whichever kind of path
// it belongs to, so the lanes of a warp share that instruction stream. The first step of a record that continues a
// curve re-evaluates the last point of the previous record (bit-identical: same inputs, same code).
__global__ void __launch_bounds__(128) flatten_eseg_emit_kernel(GGConfig cfg, const uint32_t* __restrict__ scene,
                                                                const GGPathMonoid* __restrict__ tag_monoids,
                                                                const GGESeg* __restrict__ esegs, const uint32_t* __restrict__ line_off,
                                                                GGLine* lines, uint32_t* path_bbox_ord, GGBump* bump) {
    if (bump->failed) return;
    const uint32_t n_rec = min(bump->esegs, cfg.esegs_cap);
    GGESeg r;
    CurveIn c;
    StrokeStyle st;
    V2 ns = mk(0, 0), ne = mk(0, 0);
    StrokeVertex v0 = stroke_vertex(mk(0, 0), mk(0, 0));
    float bb[4] = {3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f};
    uint32_t base = 0;
    int j = 0, n = 0;
    bool active = false, finish = false, exhausted = false, stroke = false;
    r.flags = 0; r.n = 0; r.prev = 0xffffffffu; r.path_ix = 0; r.tag_ix = 0;
    st.hw = 0; st.miter_limit = 0; st.join = 0; st.cap = 0;
    for (;;) {
        if (!active) {
            if (finish) {
                if (stroke && (r.flags & GG_ESEG_LAST)) {   // the join to the next segment / the end cap
                    LineOut o; o.n = 0; o.path_ix = r.path_ix;
                    o.bb[0] = o.bb[1] = 3.0e38f; o.bb[2] = o.bb[3] = -3.0e38f;
                    uint32_t tb = base + 2u * (uint32_t)n;
                    o.out = lines + tb; o.cap = tb < cfg.lines_cap ? cfg.lines_cap - tb : 0;
                    stroke_tail<true>(o, cfg, scene, tag_monoids, r.tag_ix, c, ne, st);
                    if (o.n > o.cap) atomicOr(&bump->failed, GG_FAIL_LINES);
                    bb[0] = fminf(bb[0], o.bb[0]); bb[1] = fminf(bb[1], o.bb[1]); bb[2] = fmaxf(bb[2], o.bb[2]); bb[3] = fmaxf(bb[3], o.bb[3]);
                }
                if (bb[0] <= bb[2]) fold_bbox(path_bbox_ord, r.path_ix, bb);
                finish = false;
            }
            uint32_t k = exhausted ? n_rec : atomicAdd(&bump->emit_cursor, 1u);
            if (k >= n_rec) {
                exhausted = true;
            } else {
                r = esegs[k];
                n = (int)r.n;
                base = line_off[r.tag_ix] + r.line_rel;
                bb[0] = bb[1] = 3.0e38f; bb[2] = bb[3] = -3.0e38f;
                stroke = (r.flags & (GG_ESEG_STROKE | GG_ESEG_CAP)) != 0;
                j = r.prev != 0xffffffffu ? -1 : 0;
                if (!stroke) {
                    if ((uint64_t)base + (uint32_t)n > cfg.lines_cap) { atomicOr(&bump->failed, GG_FAIL_LINES); n = 0; }
                    v0.p = r.p0;   // first Euler segment: the curve's start point (last_p == p0)
                    active = n > 0;
                } else {
                    load_curve(cfg, scene, tag_monoids, r.tag_ix, &c);
                    load_stroke_style(cfg, scene, c.style_ix, &st);
                    ns = stroke_normal(seg_start_tangent(c.p0, c.p1, c.p2, c.p3, c.kind), st.hw);
                    ne = stroke_normal(seg_end_tangent(c.p0, c.p1, c.p2, c.p3, c.kind), st.hw);
                    if (r.flags & GG_ESEG_CAP) {
                        LineOut o; o.n = 0; o.path_ix = r.path_ix;
                        o.bb[0] = o.bb[1] = 3.0e38f; o.bb[2] = o.bb[3] = -3.0e38f;
                        o.out = lines + base; o.cap = base < cfg.lines_cap ? cfg.lines_cap - base : 0;
                        stroke_cap<true>(o, c.p0, mk(-ns.x, -ns.y), st);
                        if (o.n > o.cap) atomicOr(&bump->failed, GG_FAIL_LINES);
                        if (o.n) fold_bbox(path_bbox_ord, r.path_ix, o.bb);
                    } else {
                        if ((uint64_t)base + 2u * (uint32_t)n > cfg.lines_cap) { atomicOr(&bump->failed, GG_FAIL_LINES); n = 0; r.flags &= ~GG_ESEG_LAST; }
                        v0 = stroke_vertex(r.p0, ns);
                        active = n > 0;
                        finish = !active;
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, exhausted && !active && !finish)) break;
        if (active) {
            // ---- one point: of this record, or (first step of a continuing record) the last one of the previous record
            const bool from_prev = j < 0;
            GGESeg q = r;
            if (from_prev) q = esegs[r.prev];
            const int jj = from_prev ? (int)q.n - 1 : j;
            StrokeVertex v1;
            if ((q.flags & (GG_ESEG_LAST | GG_ESEG_LINE)) && jj == (int)q.n - 1) {
                v1 = stroke_vertex(q.p1, ne);   // exact end point (flatten.go:141-142) with the segment's end offset
            } else {
                EulerParams ep; ep.th0 = q.th0; ep.k0 = q.k0; ep.k1 = q.k1; ep.ch = q.ch;
                const float s = eseg_param(q, jj);
                v1.p = euler_seg_eval(q.p0, q.p1, ep, s);
                v1.n = ns;
                if (stroke) {
                    V2 chord = vsub(q.p1, q.p0);
                    if (!(vlen_sq(chord) < 1e-12f)) {   // euler.go:44 otherwise: no usable chord direction, the curve's start offset
                        float th = (ep.k0 + 0.5f * ep.k1 * (s - 1.0f)) * s - ep.th0;   // euler.go:121-123
                        float sx = sin32(th), sy = cos32(th);                           // euler.go:133-137: offset direction
                        float nscale = st.hw / q.chord_len;
                        v1.n = mk((chord.x * sx - chord.y * sy) * nscale, (chord.x * sy + chord.y * sx) * nscale);
                    }
                    v1.l = vadd(v1.p, v1.n); v1.r = vsub(v1.p, v1.n);
                }
            }
            if (!from_prev) {
                if (!stroke) write_line(lines + base + j, r.path_ix, v0.p, v1.p, bb);
                else stroke_piece_emit(lines, base + 2u * (uint32_t)j, cfg.lines_cap, &bump->lines, &bump->failed, r.path_ix, v0, v1, st.hw, bb);
            }
            v0 = v1;
            j = from_prev ? 0 : j + 1;
            if (j == n) { active = false; finish = true; }
        }
    }
}

__global__ void init_frame_kernel(GGConfig cfg, uint32_t* path_bbox_ord, GGBump* bump) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cfg.n_paths; i += gridDim.x * blockDim.x) {
        uint32_t* pb = path_bbox_ord + 4 * (size_t)i;
        pb[0] = 0xffffffffu; pb[1] = 0xffffffffu; pb[2] = 0u; pb[3] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x < sizeof(GGBump) / 4) reinterpret_cast<uint32_t*>(bump)[threadIdx.x] = 0;
}

struct LoadU32 {
    const uint32_t* p;
    __device__ uint32_t operator()(uint32_t i) const { return p[i]; }
};
struct StoreU32Ex {
    uint32_t* p;
    __device__ void operator()(uint32_t i, const uint32_t& ex, const uint32_t&) const { p[i] = ex; }
};

// ------------------------------------------------------------------ path setup (coarse.go:169-223)
// Tile bbox of a path from the order-mapped float bbox, clamped to the canvas and to this
// device's band of tile rows. Off-canvas / untouched paths come out empty (0 tiles); the
// reference's uint32 wrap-around for those cases is deliberately not reproduced.
__device__ inline void path_tile_bbox(const GGConfig& cfg, const uint32_t* pb, uint32_t bbox[4]) {
    bbox[0] = bbox[1] = bbox[2] = bbox[3] = 0;
    if (pb[0] == 0xffffffffu && pb[2] == 0u) return;   // no line touched this path
    float min_x = f_unord(pb[0]), min_y = f_unord(pb[1]), max_x = f_unord(pb[2]), max_y = f_unord(pb[3]);
    if (!(min_x <= max_x) || !(min_y <= max_y)) return;
    if (min_x < 0) min_x = 0;
    if (min_y < 0) min_y = 0;
    if (max_x > (float)cfg.width) max_x = (float)cfg.width;
    if (max_y > (float)cfg.height) max_y = (float)cfg.height;
    if (max_x < min_x || max_y < min_y) return;          // entirely off-canvas
    int64_t x0 = (int64_t)floor((double)(min_x / (float)GG_TILE_W));
    int64_t y0 = (int64_t)floor((double)(min_y / (float)GG_TILE_H));
    int64_t x1 = (int64_t)ceil((double)(max_x / (float)GG_TILE_W));
    int64_t y1 = (int64_t)ceil((double)(max_y / (float)GG_TILE_H));
    if (x1 > (int64_t)cfg.width_in_tiles) x1 = cfg.width_in_tiles;
    if (y1 > (int64_t)cfg.height_in_tiles) y1 = cfg.height_in_tiles;
    // band clamp (multi-GPU): rows outside [band_y0, band_y1) belong to another device
    if (y0 < (int64_t)cfg.band_y0) y0 = cfg.band_y0;
    if (y1 > (int64_t)cfg.band_y1) y1 = cfg.band_y1;
    if (x1 <= x0 || y1 <= y0) {
        // keep the reference's bbox for degenerate-but-on-canvas paths (e.g. a vertical
        // hairline on a tile boundary) when it is representable; tile count is 0 either way
        if (x1 >= x0 && y1 >= y0) { bbox[0] = (uint32_t)x0; bbox[1] = (uint32_t)y0; bbox[2] = (uint32_t)x1; bbox[3] = (uint32_t)y1; }
        return;
    }
    bbox[0] = (uint32_t)x0; bbox[1] = (uint32_t)y0; bbox[2] = (uint32_t)x1; bbox[3] = (uint32_t)y1;
}
struct LoadPathTiles {   // packed: tiles in the low word, rows in the high word
    GGConfig cfg; const uint32_t* path_bbox_ord;
    __device__ unsigned long long operator()(uint32_t p) const {
        uint32_t bb[4];
        path_tile_bbox(cfg, path_bbox_ord + 4 * (size_t)p, bb);
        unsigned long long w = bb[2] - bb[0], h = bb[3] - bb[1];
        unsigned long long t = w * h;
        return t == 0 ? 0ull : (t | (h << 32));
    }
};
struct StorePath {
    GGConfig cfg; const uint32_t* path_bbox_ord; GGPath* paths; uint32_t* path_row_off;
    __device__ void operator()(uint32_t p, const unsigned long long& ex, const unsigned long long&) const {
        GGPath o;
        path_tile_bbox(cfg, path_bbox_ord + 4 * (size_t)p, o.bbox);
        o.tiles = (uint32_t)ex;
        paths[p] = o;
        path_row_off[p] = (uint32_t)(ex >> 32);
    }
};

__global__ void zero_tiles_kernel(GGConfig cfg, GGTile* tiles, GGBump* bump) {
    uint32_t n = bump->path_tiles;
    if (n > cfg.tiles_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&bump->failed, GG_FAIL_TILES); n = cfg.tiles_cap; }
    uint2* t = reinterpret_cast<uint2*>(tiles);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) t[i] = make_uint2(0, 0);
}

// ------------------------------------------------------------------ DDA set-up shared by path_count / path_tiling
struct DDA {
    V2 xy0, xy1, s0, s1;
    bool is_down, is_positive_slope;
    uint32_t count_x, count;
    float dx, dy, a, b, sign, x0, y0;
};
__device__ __forceinline__ void dda_setup(const GGLine& l, DDA& d) {   // path_count.go:19-39
    const float TILE_SCALE = 1.0f / 16.0f;
    V2 p0 = mk(l.p0x, l.p0y), p1 = mk(l.p1x, l.p1y);
    d.is_down = p1.y >= p0.y;
    if (d.is_down) { d.xy0 = p0; d.xy1 = p1; } else { d.xy0 = p1; d.xy1 = p0; }
    d.s0 = vmul(d.xy0, TILE_SCALE);
    d.s1 = vmul(d.xy1, TILE_SCALE);
    d.count_x = span_u(d.s0.x, d.s1.x) - 1;
    d.count = d.count_x + span_u(d.s0.y, d.s1.y);
    d.dx = fabsf(d.s1.x - d.s0.x);
    d.dy = d.s1.y - d.s0.y;
}
__device__ __forceinline__ void dda_finish(DDA& d) {   // path_count.go:40-69
    const float ONE_MINUS_ULP = 0.99999994f, ROBUST_EPSILON = 2e-7f;
    float idxdy = 1.0f / (d.dx + d.dy);
    float a = d.dx * idxdy;
    d.is_positive_slope = d.s1.x >= d.s0.x;
    d.sign = d.is_positive_slope ? 1.0f : -1.0f;
    float xt0 = f_floor(d.s0.x * d.sign);
    float c = d.s0.x * d.sign - xt0;
    d.y0 = f_floor(d.s0.y);
    float ytop = (d.s0.y == d.s1.y) ? f_ceil(d.s0.y) : d.y0 + 1.0f;
    d.b = f_min((d.dy * c + d.dx * (ytop - d.s0.y)) * idxdy, ONE_MINUS_ULP);
    float robust_err = f_floor(a * (float)(d.count - 1) + d.b) - (float)d.count_x;
    if (robust_err != 0.0f) a -= copysignf(ROBUST_EPSILON, robust_err);
    d.a = a;
    d.x0 = d.is_positive_slope ? xt0 * d.sign : xt0 * d.sign - 1.0f;
}

// path_count.go:11-205. One thread per line; tile counters and backdrops are global atomics
// (the slot a segment gets inside its tile is the value returned by the count atomic).
__global__ void __launch_bounds__(256) path_count_kernel(GGConfig cfg, const GGLine* __restrict__ lines, const GGPath* __restrict__ paths,
                                                         GGTile* tiles, GGSegCount* seg_counts, GGBump* bump) {
    if (bump->failed) return;   // an upstream buffer overflowed: this pass is redone with larger buffers
    uint32_t n_lines = min(bump->lines, cfg.lines_cap);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n_round = (n_lines + 31u) & ~31u;   // whole warps run every iteration: the slot allocation below is warp-wide
    for (uint32_t line_ix = blockIdx.x * blockDim.x + threadIdx.x; line_ix < n_round; line_ix += gridDim.x * blockDim.x) {
        uint32_t imin = 0, imax = 0;
        int32_t ymin = 0, ymax = 0;
        bool live = false;
        GGLine line; line.path_ix = 0; line.p0x = line.p0y = line.p1x = line.p1y = 0.0f;
        DDA d;
        GGPath path; path.bbox[0] = path.bbox[1] = path.bbox[2] = path.bbox[3] = 0; path.tiles = 0;
        if (line_ix < n_lines) {
            line = lines[line_ix];
            dda_setup(line, d);
            live = !(d.dx + d.dy == 0.0f) && !(d.dy == 0.0f && f_floor(d.s0.y) == d.s0.y);
        }
        if (live) {
        dda_finish(d);
        const float a = d.a, b = d.b, sign = d.sign, x0 = d.x0, y0 = d.y0;
        const V2 s0 = d.s0, s1 = d.s1;
        path = paths[line.path_ix];
        int32_t bx0 = (int32_t)path.bbox[0], by0 = (int32_t)path.bbox[1], bx1 = (int32_t)path.bbox[2], by1 = (int32_t)path.bbox[3];
        float xmin = f_min(s0.x, s1.x);
        int32_t stride = bx1 - bx0;
        if (s0.y >= (float)by1 || s1.y < (float)by0 || xmin >= (float)bx1 || stride == 0) live = false;
        if (by1 <= by0) live = false;   // empty bbox (band-clamped away)
        if (live) {
        if (s0.y < (float)by0) {
            float iminf = f_round(((float)by0 - y0 + b - a) / (1.0f - a)) - 1.0f;
            if (y0 + iminf - f_floor(a * iminf + b) < (float)by0) iminf += 1.0f;
            imin = f2u(iminf);
        }
        imax = d.count;
        if (s1.y > (float)by1) {
            float imaxf = f_round(((float)by1 - y0 + b - a) / (1.0f - a)) - 1.0f;
            if (y0 + imaxf - f_floor(a * imaxf + b) < (float)by1) imaxf += 1.0f;
            imax = f2u(imaxf);
        }
        if (f_max(s0.x, s1.x) < (float)bx0) {
            ymin = f2i(f_ceil(s0.y));
            ymax = f2i(f_ceil(s1.y));
            imax = imin;
        } else {
            float fudge = d.is_positive_slope ? 0.0f : 1.0f;
            if (xmin < (float)bx0) {
                float f = f_round((sign * ((float)bx0 - x0) - b + fudge) / a);
                if ((x0 + sign * f_floor(a * f + b) < (float)bx0) == d.is_positive_slope) f += 1.0f;
                int32_t ynext = f2i(y0 + f - f_floor(a * f + b) + 1.0f);
                if (d.is_positive_slope) {
                    if (f2u(f) > imin) {
                        float y_off = (y0 != s0.y) ? 1.0f : 0.0f;
                        ymin = f2i(y0 + y_off);
                        ymax = ynext;
                        imin = f2u(f);
                    }
                } else if (f2u(f) < imax) {
                    ymin = ynext;
                    ymax = f2i(f_ceil(s1.y));
                    imax = f2u(f);
                }
            }
            if (f_max(s0.x, s1.x) > (float)bx1) {
                float f = f_round((sign * ((float)bx1 - x0) - b + fudge) / a);
                if ((x0 + sign * f_floor(a * f + b) < (float)bx1) == d.is_positive_slope) f += 1.0f;
                if (d.is_positive_slope) imax = min(imax, f2u(f));
                else imin = max(imin, f2u(f));
            }
        }
        imax = max(imin, imax);
        ymin = max(ymin, by0);
        ymax = min(ymax, by1);
        }   // live (bbox)
        }   // live (non-degenerate)
        if (!live) { imin = imax = 0; ymin = ymax = 0; }
        // SegmentCount slots: one returning atomic per WARP (the lines' counts are prefix-summed with shuffles); one per
        // line serialised 1.6 M returning atomics on a single address (long scoreboard 89 cycles per issue, ncu r1c)
        const uint32_t n_seg = imax - imin;
        uint32_t incl = n_seg;
#pragma unroll
        for (int dl = 1; dl < 32; dl <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, dl); if ((int)lane >= dl) incl += o; }
        const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t warp_base = 0;
        if (lane == 31 && warp_total) warp_base = atomicAdd(&bump->seg_counts, warp_total);
        warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
        if (!live) continue;
        const float a = d.a, b = d.b, sign = d.sign, x0 = d.x0, y0 = d.y0;
        const V2 s0 = d.s0;
        const int32_t bx0 = (int32_t)path.bbox[0], by0 = (int32_t)path.bbox[1], bx1 = (int32_t)path.bbox[2];
        const int32_t stride = bx1 - bx0;
        const int32_t delta = d.is_down ? -1 : 1;
        for (int32_t y = ymin; y < ymax; y++) {
            int32_t base = (int32_t)path.tiles + (y - by0) * stride;
            atomicAdd(&tiles[base].backdrop, delta);
        }
        float last_z = f_floor(a * (float)(imin - 1) + b);
        const uint32_t seg_base = warp_base + incl - n_seg;
        bool store = false;
        if (n_seg) {
            store = (uint64_t)seg_base + n_seg <= cfg.seg_counts_cap;
            if (!store) atomicOr(&bump->failed, GG_FAIL_SEGCOUNTS);
        }
        for (uint32_t i = imin; i < imax; i++) {
            float zf = a * (float)i + b;
            float z = f_floor(zf);
            int32_t y = f2i(y0 + (float)i - z);
            int32_t x = f2i(x0 + sign * z);
            int32_t base = (int32_t)path.tiles + (y - by0) * stride - bx0;
            bool top_edge = (i == 0) ? (y0 == s0.y) : (last_z == z);
            if (top_edge && x + 1 < bx1) {
                int32_t x_bump = max(x + 1, bx0);
                atomicAdd(&tiles[base + x_bump].backdrop, delta);
            }
            // The count needs no return value here (a fire-and-forget RED, not a scoreboard-stalling ATOM): the slot a
            // segment takes inside its tile is claimed later by path_tiling, one thread per segment, where the
            // returning atomic's latency hides behind millions of independent threads (it stalled this kernel for
            // 97 cycles per issue when the reference's seg_within_slice was taken here, path_count.go:192-194).
            atomicAdd(&tiles[base + x].seg_count, 1u);
            const uint32_t seg_within_slice = 0;
            if (store) {
                GGSegCount sc; sc.line_ix = line_ix; sc.counts = (seg_within_slice << 16) | i;
                seg_counts[seg_base + i - imin] = sc;
            }
            last_z = z;
        }
    }
}

// ------------------------------------------------------------------ backdrop prefix sum + tile-hit pass
// Rows of all per-path tile rectangles are enumerated flat; a group of 8 lanes owns one row
// and walks it 8 tiles at a time with a shuffle prefix sum (coarse.go:291-299, backdrop.wgsl).
// PASS 0: write the summed backdrop back and add this path's contribution to the per-tile
//         hit/word histogram of coarse. PASS 1: scatter draw indices into the per-tile hit lists.
__device__ __forceinline__ uint32_t find_row_path(const uint32_t* __restrict__ path_row_off, uint32_t n_paths, uint32_t row) {
    uint32_t lo = 0, hi = n_paths;   // largest p with path_row_off[p] <= row
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (path_row_off[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

template <int PASS>
__global__ void __launch_bounds__(256) tile_rows_kernel(GGConfig cfg, const uint32_t* __restrict__ scene, const GGPath* __restrict__ paths,
                                                        const uint32_t* __restrict__ path_row_off, GGTile* tiles,
                                                        const GGDrawRec* __restrict__ recs,
                                                        unsigned long long* tile_hits, const uint32_t* __restrict__ hit_off,
                                                        uint32_t* hit_cursor, GGHit* hits, const uint32_t* __restrict__ seg_start,
                                                        uint8_t* imp_mask, uint32_t* imp_seen, GGBump* bump) {
    cg::thread_block_tile<8> g = cg::tiled_partition<8>(cg::this_thread_block());
    const uint32_t n_rows = min(bump->path_rows, cfg.rows_cap);
    if (bump->failed || bump->path_tiles > cfg.tiles_cap) return;
    if (PASS == 1 && bump->hits > cfg.hits_cap) return;
    const uint32_t groups = gridDim.x * blockDim.x / 8;
    for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) / 8; row < n_rows; row += groups) {
        uint32_t p = find_row_path(path_row_off, cfg.n_paths, row);
        GGPath path = paths[p];
        uint32_t bw = path.bbox[2] - path.bbox[0];
        uint32_t y = row - path_row_off[p];
        uint32_t base = path.tiles + y * bw;
        uint32_t gy = path.bbox[1] + y - cfg.band_y0;
        uint32_t tag = scene[cfg.draw_tag_base + p];   // one path marker per draw object: path p <-> draw p
        if (tag == GG_DRAWTAG_GRADIENT) tag = GG_DRAWTAG_COLOR;
        // Even-odd fills: a tile without segments is inside only where the winding is odd. (DEVIATION: coarse.go:425
        // paints every tile with backdrop != 0 solid whatever the rule -- the hole of two nested same-direction
        // contours, backdrop 2, came out filled. The oracle follows; ot_evenodd_solid_quirk restores the reference.)
        const bool even_odd = tag == GG_DRAWTAG_COLOR && (recs[p].b & 1u) != 0u;
        // Implicit layers (GG_BLEND_IMPLICIT: no geometry, full coverage) have no tiles of their own: a hit of something
        // they enclose brings their Begin/End pair into that tile's list -- once per (layer, tile): PASS 0 claims the
        // pair with one bit per (layer, tile) and remembers per path tile which ancestors it brought (imp_mask), PASS 1
        // scatters exactly those. (Without the bitmap every enclosed hit brought its ancestors and coarse dropped the
        // duplicates after sorting: 5.6 M hits instead of 2.x M on the benchmark scene.) The chain is the same for the
        // whole row: walk it once; ancestors beyond the first four are always brought.
        uint32_t n_imp = 0, i0 = 0, i1 = 0, i2 = 0, i3 = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;   // (scalars: a dynamically indexed array would live in local memory)
        uint32_t o0 = 0, o1 = 0, o2 = 0, o3 = 0;   // ordinals of those layers (style word 1 of their clip path)
        for (int32_t a = recs[p].parent; a >= 0;) {
            const GGDrawRec ra = recs[a];
            if (ra.b & GG_BLEND_IMPLICIT) {
                const uint32_t ord = n_imp < 4 ? scene[cfg.style_base + GG_STYLE_WORDS * (uint32_t)a + 1] : 0u;
                if (n_imp == 0) { i0 = (uint32_t)a; e0 = ra.a; o0 = ord; } else if (n_imp == 1) { i1 = (uint32_t)a; e1 = ra.a; o1 = ord; }
                else if (n_imp == 2) { i2 = (uint32_t)a; e2 = ra.a; o2 = ord; } else if (n_imp == 3) { i3 = (uint32_t)a; e3 = ra.a; o3 = ord; }
                n_imp++;
            }
            a = ra.parent;
        }
        int32_t carry = 0;
        for (uint32_t x0 = 0; x0 < bw; x0 += 8) {
            uint32_t x = x0 + g.thread_rank();
            GGTile t; t.backdrop = 0; t.seg_count = 0;
            if (x < bw) t = tiles[base + x];
            int32_t v = t.backdrop;
            if (PASS == 0) {
#pragma unroll
                for (int dlt = 1; dlt < 8; dlt <<= 1) {
                    int32_t o = g.shfl_up(v, dlt);
                    if ((int)g.thread_rank() >= dlt) v += o;
                }
                v += carry;
                carry = g.shfl(v, 7);
            }
            if (x < bw) {
                if (PASS == 0 && v != t.backdrop) tiles[base + x].backdrop = v;
                if (t.seg_count != 0 || (even_odd ? (v & 1) != 0 : v != 0)) {
                    uint32_t T = gy * cfg.width_in_tiles + path.bbox[0] + x;
                    if (PASS == 0) {
                        unsigned long long w;
                        if (tag == GG_DRAWTAG_COLOR) w = 1ull | ((t.seg_count ? 6ull : 3ull) << 32);
                        else if (tag == GG_DRAWTAG_BEGIN_CLIP) w = 2ull | ((1ull + (t.seg_count ? 7ull : 4ull)) << 32);
                        else w = 0;
                        if (w) {
                            uint32_t mask = 0xfu;   // which of the (up to four) cached ancestors this path tile brings
                            if (cfg.imp_words) {
                                const uint32_t Tb = T - 0u, bit = 1u << (Tb & 31u), wi = Tb >> 5;
                                mask = 0;
                                if (n_imp > 0 && !(atomicOr(&imp_seen[(size_t)o0 * cfg.imp_words + wi], bit) & bit)) mask |= 1u;
                                if (n_imp > 1 && !(atomicOr(&imp_seen[(size_t)o1 * cfg.imp_words + wi], bit) & bit)) mask |= 2u;
                                if (n_imp > 2 && !(atomicOr(&imp_seen[(size_t)o2 * cfg.imp_words + wi], bit) & bit)) mask |= 4u;
                                if (n_imp > 3 && !(atomicOr(&imp_seen[(size_t)o3 * cfg.imp_words + wi], bit) & bit)) mask |= 8u;
                            }
                            mask &= (1u << min(n_imp, 4u)) - 1u;
                            imp_mask[base + x] = (uint8_t)mask;
                            const uint32_t n_add = (uint32_t)__popc(mask) + (n_imp > 4 ? n_imp - 4 : 0u);
                            atomicAdd(&tile_hits[T], w + n_add * (2ull | (5ull << 32)));
                        }
                    } else if (tag == GG_DRAWTAG_COLOR || tag == GG_DRAWTAG_BEGIN_CLIP) {
                        const uint32_t mask = imp_mask[base + x];
                        const uint32_t n_add = (uint32_t)__popc(mask) + (n_imp > 4 ? n_imp - 4 : 0u);
                        uint32_t own = tag == GG_DRAWTAG_COLOR ? 1u : 2u;
                        uint32_t slot = hit_off[T] + atomicAdd(&hit_cursor[T], own + 2u * n_add);
                        if (slot + own + 2u * n_add <= cfg.hits_cap) {
                            // path_tiling advanced seg_start[] to the END of the tile's range
                            GGHit h; h.draw = p; h.seg_count = t.seg_count; h.seg_start = seg_start[base + x] - t.seg_count; h.backdrop = t.backdrop;
                            hits[slot++] = h;
                            if (own == 2u) { h.draw = recs[p].a; hits[slot++] = h; }   // the clip's EndClip fills with the same path tile
                            GGHit li; li.seg_count = 0; li.seg_start = 0; li.backdrop = 1;   // layers without geometry: full coverage
                            if (mask & 1u) { li.draw = i0; hits[slot++] = li; li.draw = e0; hits[slot++] = li; }
                            if (mask & 2u) { li.draw = i1; hits[slot++] = li; li.draw = e1; hits[slot++] = li; }
                            if (mask & 4u) { li.draw = i2; hits[slot++] = li; li.draw = e2; hits[slot++] = li; }
                            if (mask & 8u) { li.draw = i3; hits[slot++] = li; li.draw = e3; hits[slot++] = li; }
                            if (n_imp > 4) {   // deeper than the cached part of the chain: walk it
                                uint32_t k = 0;
                                for (int32_t a = recs[p].parent; a >= 0; a = recs[a].parent)
                                    if (recs[a].b & GG_BLEND_IMPLICIT) { if (k >= 4) { li.draw = (uint32_t)a; hits[slot++] = li; li.draw = recs[a].a; hits[slot++] = li; } k++; }
                            }
                        }
                    }
                }
            }
        }
    }
}

struct LoadTileCount {
    const GGTile* tiles;
    __device__ uint32_t operator()(uint32_t i) const { return tiles[i].seg_count; }
};
struct LoadTileHits {   // + 2 words per tile: blend-offset word 0 and the CmdEnd terminator (ptcl.go:72-78, coarse.go:148-151)
    const unsigned long long* h;
    __device__ unsigned long long operator()(uint32_t i) const {
        unsigned long long v = h[i] + (2ull << 32);
        unsigned long long words = ((v >> 32) + 3ull) & ~3ull;   // pad every list to 16 bytes (bulk-copy granularity in fine)
        return (v & 0xffffffffull) | (words << 32);
    }
};
struct StoreTileHits {
    uint32_t* hit_off; uint32_t* hit_cnt; uint32_t* ptcl_off; uint32_t* hit_cursor; uint32_t* spill_off;
    __device__ void operator()(uint32_t i, const unsigned long long& ex, const unsigned long long& v) const {
        hit_off[i] = (uint32_t)ex; hit_cnt[i] = (uint32_t)v; ptcl_off[i] = (uint32_t)(ex >> 32);
        hit_cursor[i] = 0; spill_off[i] = 0xffffffffu;
    }
};

// ------------------------------------------------------------------ path_tiling.go:11-199
__global__ void __launch_bounds__(256) path_tiling_kernel(GGConfig cfg, const GGSegCount* __restrict__ seg_counts, const GGLine* __restrict__ lines,
                                                          const GGPath* __restrict__ paths, const GGTile* __restrict__ tiles,
                                                          uint32_t* seg_start, GGSegment* segments, GGBump* bump) {
    uint32_t n = min(bump->seg_counts, cfg.seg_counts_cap);
    if (bump->failed) return;
    if (bump->segments > cfg.segments_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&bump->failed, GG_FAIL_SEGMENTS); return; }
    for (uint32_t seg_ix = blockIdx.x * blockDim.x + threadIdx.x; seg_ix < n; seg_ix += gridDim.x * blockDim.x) {
        GGSegCount sc = seg_counts[seg_ix];
        GGLine line = lines[sc.line_ix];
        uint32_t seg_within_line = sc.counts & 0xffffu;
        DDA d; dda_setup(line, d); dda_finish(d);
        const float a = d.a, b = d.b, sign = d.sign, x0 = d.x0, y0 = d.y0;
        V2 xy0 = d.xy0, xy1 = d.xy1;
        float z = f_floor(a * (float)seg_within_line + b);
        int32_t x = f2i(x0) + f2i(sign * z);                 // split truncation, path_tiling.go:74
        int32_t y = f2i(y0 + (float)seg_within_line - z);
        GGPath path = paths[line.path_ix];
        int32_t bx0 = (int32_t)path.bbox[0], by0 = (int32_t)path.bbox[1], bx1 = (int32_t)path.bbox[2];
        int32_t stride = bx1 - bx0;
        int32_t tile_ix = (int32_t)path.tiles + (y - by0) * stride + x - bx0;
        if (tiles[tile_ix].seg_count == 0) continue;
        uint32_t out_ix = atomicAdd(&seg_start[tile_ix], 1u);   // claims the next slot: seg_start[] ends up as the END of each tile's range
        const float TW = (float)GG_TILE_W, TH = (float)GG_TILE_H;
        V2 tile_xy = mk((float)x * TW, (float)y * TH);
        V2 tile_xy1 = mk(tile_xy.x + TW, tile_xy.y + TH);
        if (seg_within_line > 0) {
            float z_prev = f_floor(a * (float)(seg_within_line - 1) + b);
            if (z == z_prev) {
                float xt = xy0.x + (xy1.x - xy0.x) * (tile_xy.y - xy0.y) / (xy1.y - xy0.y);
                xt = f_clamp(xt, tile_xy.x + 1e-3f, tile_xy1.x);
                xy0 = mk(xt, tile_xy.y);
            } else {
                float x_clip = d.is_positive_slope ? tile_xy.x : tile_xy1.x;
                float yt = xy0.y + (xy1.y - xy0.y) * (x_clip - xy0.x) / (xy1.x - xy0.x);
                yt = f_clamp(yt, tile_xy.y + 1e-3f, tile_xy1.y);
                xy0 = mk(x_clip, yt);
            }
        }
        if (seg_within_line < d.count - 1) {
            float z_next = f_floor(a * (float)(seg_within_line + 1) + b);
            if (z == z_next) {
                float xt = xy0.x + (xy1.x - xy0.x) * (tile_xy1.y - xy0.y) / (xy1.y - xy0.y);
                xt = f_clamp(xt, tile_xy.x + 1e-3f, tile_xy1.x);
                xy1 = mk(xt, tile_xy1.y);
            } else {
                float x_clip = d.is_positive_slope ? tile_xy1.x : tile_xy.x;
                float yt = xy0.y + (xy1.y - xy0.y) * (x_clip - xy0.x) / (xy1.x - xy0.x);
                yt = f_clamp(yt, tile_xy.y + 1e-3f, tile_xy1.y);
                xy1 = mk(x_clip, yt);
            }
        }
        float y_edge = 1e9f;
        V2 p0o = vsub(xy0, tile_xy), p1o = vsub(xy1, tile_xy);
        const float epsilon = 1e-6f;
        if (p0o.x == 0.0f) {
            if (p1o.x == 0.0f) {
                p0o.x = epsilon;
                if (p0o.y == 0.0f) { p1o.x = epsilon; p1o.y = TH; }
                else { p1o.x = 2.0f * epsilon; p1o.y = p0o.y; }
            } else if (p0o.y == 0.0f) {
                p0o.x = epsilon;
            } else {
                y_edge = p0o.y;
            }
        } else if (p1o.x == 0.0f) {
            if (p1o.y == 0.0f) p1o.x = epsilon; else y_edge = p1o.y;
        }
        if (p0o.x == f_floor(p0o.x) && p0o.x != 0.0f) p0o.x -= epsilon;
        if (p1o.x == f_floor(p1o.x) && p1o.x != 0.0f) p1o.x -= epsilon;
        if (!d.is_down) { V2 t = p0o; p0o = p1o; p1o = t; }
        if (out_ix < cfg.segments_cap) {
            GGSegment s; s.p0x = p0o.x; s.p0y = p0o.y; s.p1x = p1o.x; s.p1y = p1o.y; s.y_edge = y_edge;
            segments[out_ix] = s;
        }
    }
}

// ------------------------------------------------------------------ coarse (coarse.go:322-625)
// One warp per tile. The tile's hit list (draw indices, scattered by atomics) is sorted in
// shared memory so commands come out in scene order; the clip state machine of the
// reference (clipDepth / clipZeroDepth) collapses to a single "innermost emitted clip"
// register because a draw is live in a tile iff every enclosing clip emitted there.
#define COARSE_WARPS 4
#define COARSE_CAP 1024   // hits per tile handled in shared memory; longer lists are sorted in place by lane 0 and replayed sequentially

__device__ inline void heap_sort_global(GGHit* a, uint32_t n) {   // by draw index; equal keys are identical records
    for (uint32_t start = n / 2; start-- > 0;) {
        uint32_t root = start;
        for (;;) { uint32_t c = 2 * root + 1; if (c >= n) break; if (c + 1 < n && a[c].draw < a[c + 1].draw) c++; if (a[root].draw >= a[c].draw) break; GGHit t = a[root]; a[root] = a[c]; a[c] = t; root = c; }
    }
    for (uint32_t end = n; end-- > 1;) {
        GGHit t = a[0]; a[0] = a[end]; a[end] = t;
        uint32_t root = 0;
        for (;;) { uint32_t c = 2 * root + 1; if (c >= end) break; if (c + 1 < end && a[c].draw < a[c + 1].draw) c++; if (a[root].draw >= a[c].draw) break; GGHit t2 = a[root]; a[root] = a[c]; a[c] = t2; root = c; }
    }
}

// Sequential form of the clip state machine (coarse.go:440-625 with lazy layers), one warp, hits in `sorted`:
// used for tiles with more than COARSE_CAP hits. The shared-memory path below computes the same PTCL in parallel.
__device__ __noinline__ void coarse_tile_sequential(const GGConfig& cfg, uint32_t T, uint32_t n, const GGHit* sorted, uint32_t pos,
                                                    const GGDrawRec* __restrict__ recs, const uint32_t* __restrict__ ptcl_off, uint32_t* ptcl_len,
                                                    uint32_t* ptcl, uint32_t* spill_off, uint32_t* restart_pt, GGBump* bump) {
    const uint32_t lane = threadIdx.x & 31;
        // Layers whose blend word carries GG_BLEND_ELIDE_EMPTY are opened lazily: their BeginClip is only
        // written (just before the first command they enclose in this tile) once something is drawn inside;
        // a layer still pending at its EndClip leaves no trace. `depth` counts logically open clips,
        // `mat` how many of them (always the outermost ones) have been written to the PTCL.
        int32_t top = -1;
        uint32_t depth = 0, mat = 0, max_depth = 0;
        // Restart point for fine: the PTCL position after the last command at clip depth 0 that fixes every pixel
        // of the tile regardless of what came before -- an opaque CmdSolid+CmdColor, or a full-coverage layer whose
        // compose mode wipes the backdrop (Clear always; Copy/SrcIn/DestIn/SrcOut/DestAtop when nothing was drawn
        // inside). The PTCL itself is complete (it must match the reference word for word); fine merely starts there.
        uint32_t pos_u = pos - ptcl_off[T], restart = 0, restart_rgba = 0, d0_begin_pos = 0xffffffffu;
        for (uint32_t base = 0; base < n; base += 32) {
            uint32_t i = base + lane;
            // parallel gather of everything the state machine and the emitters need
            GGDrawRec r; r.tag = 0; r.parent = -1; r.a = 0; r.b = 0;
            uint32_t d = 0; GGTile t; t.backdrop = 0; t.seg_count = 0; uint32_t sstart = 0; int32_t begin_parent = -1;
            if (i < n && !(i > 0 && sorted[i].draw == sorted[i - 1].draw)) {   // implicit-layer hits arrive once per enclosed hit: keep the first
                const GGHit h = sorted[i];
                d = h.draw;
                r = recs[d];
                t.backdrop = h.backdrop; t.seg_count = h.seg_count; sstart = h.seg_start;
                if (r.tag == GG_DRAWTAG_END_CLIP) begin_parent = recs[r.parent].parent;
            }
            // per-hit flags for the restart bookkeeping: bit 0 = tile has segments for this path,
            // bits 1-2: 1 = opaque colour, 2 = End of a wiping layer (if empty), 3 = End of a Clear layer
            uint32_t hflags = t.seg_count ? 1u : 0u;
            if (r.tag == GG_DRAWTAG_COLOR) { if ((r.a >> 24) == 255u && !(r.b & 2u)) hflags |= 1u << 1; }
            else if (r.tag == GG_DRAWTAG_END_CLIP) {
                uint32_t mixm = (r.a >> 8) & 0xffu, comp = r.a & 0xffu;
                if (mixm == 0u && comp == 0u) hflags |= 3u << 1;
                else if (mixm == 0u && (comp == 1u || comp == 5u || comp == 6u || comp == 7u || comp == 10u)) hflags |= 2u << 1;
            }
            // sequential, warp-uniform replay of the clip state over the (up to) 32 gathered hits
            bool emit = false;
            uint32_t pre = 0;   // pending BeginClip words this hit has to write before its own command
            uint32_t cnt = min(32u, n - base);
            for (uint32_t j = 0; j < cnt; j++) {
                uint32_t jt = __shfl_sync(0xffffffffu, r.tag, j);
                int32_t jp = __shfl_sync(0xffffffffu, r.parent, j);
                uint32_t jd = __shfl_sync(0xffffffffu, d, j);
                int32_t jbp = __shfl_sync(0xffffffffu, begin_parent, j);
                uint32_t jb = __shfl_sync(0xffffffffu, r.b, j);
                uint32_t jf = __shfl_sync(0xffffffffu, hflags, j);
                uint32_t ja = __shfl_sync(0xffffffffu, r.a, j);
                bool e = false;
                uint32_t pj = 0;
                if (jt == GG_DRAWTAG_COLOR) {
                    if (jp == top) {
                        e = true; pj = depth - mat; mat = depth;
                        uint32_t words = pj + ((jf & 1u) ? 6u : 3u);
                        if (depth == 0 && jf == (1u << 1)) { restart = pos_u + words; restart_rgba = ja; }   // opaque, no segments
                        pos_u += words;
                    }
                } else if (jt == GG_DRAWTAG_BEGIN_CLIP) {
                    if (jp == top) {
                        top = (int32_t)jd;
                        if (jb & GG_BLEND_ELIDE_EMPTY) { depth++; }                       // pending
                        else {                                                              // written now (+ pending parents)
                            e = true; pj = depth - mat;
                            pos_u += pj + 1u;
                            if (depth == 0) d0_begin_pos = pos_u;
                            depth++; mat = depth;
                        }
                    }
                } else if (jt == GG_DRAWTAG_END_CLIP) {
                    if (top == jp) {   // jp == index of the matching BeginClip
                        top = jbp;
                        if (depth > mat) { depth--; }                                     // never materialised: nothing to close
                        else {
                            e = true; depth--; mat--;
                            uint32_t words = (jf & 1u) ? 7u : 4u, kind = jf >> 1;
                            if (depth == 0 && !(jf & 1u) && (kind == 3u || (kind == 2u && pos_u == d0_begin_pos))) { restart = pos_u + words; restart_rgba = 0; }
                            pos_u += words;
                        }
                    }
                }
                max_depth = max(max_depth, mat);
                if (lane == j) { emit = e; pre = pj; }
            }
            // words per hit, warp exclusive scan, parallel emission
            uint32_t nw = 0;
            if (emit) {
                if (r.tag == GG_DRAWTAG_COLOR) nw = t.seg_count ? 6u : 3u;
                else if (r.tag == GG_DRAWTAG_BEGIN_CLIP) nw = 1u;
                else nw = t.seg_count ? 7u : 4u;
                nw += pre;
            }
            uint32_t inc = nw;
#pragma unroll
            for (int dl = 1; dl < 32; dl <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, dl); if ((int)lane >= dl) inc += o; }
            uint32_t o = pos + inc - nw;
            if (emit) {
                for (uint32_t k = 0; k < pre; k++) ptcl[o++] = GG_CMD_BEGIN_CLIP;
                if (r.tag == GG_DRAWTAG_COLOR) {
                    if (t.seg_count) { ptcl[o++] = GG_CMD_FILL; ptcl[o++] = (t.seg_count << 1) | (r.b & 1u); ptcl[o++] = sstart; ptcl[o++] = (uint32_t)t.backdrop; }
                    else ptcl[o++] = GG_CMD_SOLID;
                    ptcl[o++] = (r.b & 2u) ? GG_CMD_GRAD : GG_CMD_COLOR; ptcl[o++] = r.a;
                } else if (r.tag == GG_DRAWTAG_BEGIN_CLIP) {
                    ptcl[o++] = GG_CMD_BEGIN_CLIP;
                } else {
                    if (t.seg_count) { ptcl[o++] = GG_CMD_FILL; ptcl[o++] = (t.seg_count << 1); ptcl[o++] = sstart; ptcl[o++] = (uint32_t)t.backdrop; }
                    else ptcl[o++] = GG_CMD_SOLID;
                    ptcl[o++] = GG_CMD_END_CLIP; ptcl[o++] = r.a; ptcl[o++] = r.b;
                }
            }
            pos += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            ptcl[pos] = GG_CMD_END;
            ptcl_len[T] = pos + 1 - ptcl_off[T];
            restart_pt[2 * T] = restart; restart_pt[2 * T + 1] = restart_rgba;
            if (pos + 1 - ptcl_off[T] - max(restart, 1u) > GG_FINE_HEAVY_WORDS)
                spill_off[(cfg.band_y1 - cfg.band_y0) * cfg.width_in_tiles + atomicAdd(&bump->heavy, 1u)] = T;
            if (max_depth > GG_BLEND_STACK_SPLIT) {
                uint32_t lv = max_depth - GG_BLEND_STACK_SPLIT;
                uint32_t so = atomicAdd(&bump->spill, lv);
                if (so + lv > cfg.spill_cap) atomicOr(&bump->failed, GG_FAIL_SPILL); else spill_off[T] = so;
            }
        }
}

// Per-hit state byte of the parallel path
#define CH_TAG 3u        // 0 colour, 1 BeginClip, 2 EndClip
#define CH_ELIDE 4u      // BeginClip that may be dropped where nothing is drawn inside (GG_BLEND_ELIDE_EMPTY)
#define CH_VAL 8u        // colour: drawn; BeginClip: entered; EndClip: its BeginClip was entered
#define CH_NE 16u        // BeginClip: something inside it is written to this tile's PTCL
#define CH_NONE 0xffffu  // lpos: no enclosing clip
#define CH_MISSING 0xfffeu   // lpos: the enclosing clip has no hit in this tile (zero coverage here): the hit is culled

// One warp per tile. The tile's hits (draw indices scattered by tile_rows<1>) are sorted in shared memory so that
// they are in scene order, then everything the reference's sequential clip state machine decides
// (coarse.go:440-625: clipDepth / clipZeroDepth / blendDepth) is computed in parallel from one fact: a hit takes part
// iff its enclosing BeginClip has a hit in this tile and takes part itself. `lpos` links every hit to the position
// of that BeginClip (an EndClip: to its own BeginClip), liveness is a few rounds of pointer chasing over those links,
// "this lazily opened layer encloses something" an upward marking, PTCL offsets a warp scan. The first version
// replayed the state machine hit by hit with seven shuffles per hit: 224 of the kernel's 442 M warp instructions.
__global__ void __launch_bounds__(COARSE_WARPS * 32) coarse_kernel(GGConfig cfg, const GGDrawRec* __restrict__ recs,
                                                                   const uint32_t* __restrict__ hit_off, const uint32_t* __restrict__ hit_cnt,
                                                                   GGHit* hits, const uint32_t* __restrict__ ptcl_off, uint32_t* ptcl_len, uint32_t* ptcl,
                                                                   uint32_t* spill_off, uint32_t* restart_pt, GGBump* bump) {
    // per warp: sorted keys (4 KB) | radix ping-pong buffer (4 KB), reused after the sort for lpos (2 KB) + state (1 KB) | histogram (1 KB)
    __shared__ uint32_t sort_buf[COARSE_WARPS][COARSE_CAP];
    __shared__ uint32_t tmp_buf[COARSE_WARPS][COARSE_CAP];
    __shared__ uint32_t hist_buf[COARSE_WARPS][256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = cfg.width_in_tiles * (cfg.band_y1 - cfg.band_y0);
    if (bump->failed) return;
    const bool overflow = bump->hits > cfg.hits_cap || bump->ptcl_words > cfg.ptcl_cap;
    if (overflow) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&bump->failed, (bump->hits > cfg.hits_cap ? GG_FAIL_HITS : 0u) | (bump->ptcl_words > cfg.ptcl_cap ? GG_FAIL_PTCL : 0u)); return; }
    uint32_t* sb = sort_buf[warp];
    uint32_t* tmp = tmp_buf[warp];
    uint32_t* hist = hist_buf[warp];
    const uint32_t key_bits = 32u - (uint32_t)__clz((int)max(cfg.n_draws, 2u) - 1);   // draw indices are < n_draws
    uint16_t* lpos = reinterpret_cast<uint16_t*>(tmp);
    uint8_t* st = reinterpret_cast<uint8_t*>(tmp) + 2 * COARSE_CAP;
    for (;;) {
        uint32_t T = 0;
        if (lane == 0) T = atomicAdd(&bump->coarse_cursor, 1u);
        T = __shfl_sync(0xffffffffu, T, 0);
        if (T >= n_tiles) break;
        const uint32_t n = hit_cnt[T];
        GGHit* list = hits + hit_off[T];
        uint32_t pos = ptcl_off[T];
        if (lane == 0) ptcl[pos] = 0;   // word 0: blend offset, always 0 as in the reference (spill offsets live in spill_off[])
        pos += 1;
        if (n == 0) { if (lane == 0) { ptcl[pos] = GG_CMD_END; ptcl_len[T] = 2; restart_pt[2 * T] = 0; restart_pt[2 * T + 1] = 0; } continue; }
        if (n > COARSE_CAP || key_bits > 22u) {   // (keys pack the draw index above 10 position bits)
            if (lane == 0) heap_sort_global(list, n);
            __syncwarp();
            coarse_tile_sequential(cfg, T, n, list, pos, recs, ptcl_off, ptcl_len, ptcl, spill_off, restart_pt, bump);
            __syncwarp();
            continue;
        }
        // ---- sort: keys are (draw index << 10) | position in the unsorted list, so that a hit's record can be fetched by
        //      position afterwards; stable LSD radix sort on the draw-index bits, 8 bits per pass (2 passes up to 65 536 draws). Per pass a
        //      256-bin histogram (shared-memory atomics), an exclusive scan (8 bins per lane) and a stable scatter that
        //      ranks equal digits inside each group of 32 keys with match_any. The bitonic network this replaces cost
        //      ~36-45 passes over the padded list: most of the kernel's instructions (ncu, r1b).
        uint32_t* src = sb;
        uint32_t* dstb = tmp;
        for (uint32_t i = lane; i < n; i += 32) src[i] = (list[i].draw << 10) | i;
        for (uint32_t shift = 10; shift < 10 + key_bits; shift += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) hist[lane * 8 + k] = 0;
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) atomicAdd(&hist[(src[i] >> shift) & 255u], 1u);
            __syncwarp();
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) { loc[k] = sum; sum += hist[lane * 8 + k]; }
            uint32_t incl = sum;
#pragma unroll
            for (int dl = 1; dl < 32; dl <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, dl); if ((int)lane >= dl) incl += o; }
            const uint32_t excl = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; k++) hist[lane * 8 + k] = excl + loc[k];
            __syncwarp();
            for (uint32_t base = 0; base < n; base += 32) {
                const uint32_t i = base + lane;
                const bool valid = i < n;
                const uint32_t key = valid ? src[i] : 0u;
                const uint32_t digit = valid ? ((key >> shift) & 255u) : (256u + lane);
                const uint32_t peers = __match_any_sync(0xffffffffu, digit);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                uint32_t at = 0;
                if (valid) at = hist[digit] + rank;
                __syncwarp();
                if (valid) {
                    dstb[at] = key;
                    if (rank + 1 == (uint32_t)__popc(peers)) hist[digit] = at + 1;   // the last of the group moves the bin on
                }
                __syncwarp();
            }
            uint32_t* sw = src; src = dstb; dstb = sw;
        }
        if (src != sb) { for (uint32_t i = lane; i < n; i += 32) sb[i] = src[i]; }
        __syncwarp();
        // ---- A: classify every hit, link it to the hit of its enclosing BeginClip
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t d = sb[i] >> 10;
            const bool dup = i > 0 && (sb[i - 1] >> 10) == d;   // implicit-layer hits arrive once per enclosed hit: keep the first
            const GGDrawRec r = recs[d];
            uint32_t s8 = r.tag == GG_DRAWTAG_COLOR ? 0u : (r.tag == GG_DRAWTAG_BEGIN_CLIP ? 1u : 2u);
            if (s8 == 1u && (r.b & GG_BLEND_ELIDE_EMPTY)) s8 |= CH_ELIDE;
            uint32_t lp = CH_NONE;
            if (r.parent >= 0) {
                const uint32_t key = (uint32_t)r.parent;
                uint32_t lo = 0, hi = n;   // lower bound
                while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((sb[mid] >> 10) < key) lo = mid + 1; else hi = mid; }
                lp = (lo < n && (sb[lo] >> 10) == key) ? lo : CH_MISSING;
            }
            if (!dup && lp != CH_MISSING) s8 |= CH_VAL;
            st[i] = (uint8_t)s8;
            lpos[i] = (uint16_t)lp;
        }
        __syncwarp();
        // ---- B: a hit takes part iff its link does (chains are as long as clips nest: a few rounds)
        for (;;) {
            bool changed = false;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t s8 = st[i], lp = lpos[i];
                if ((s8 & CH_VAL) && lp < CH_MISSING && !(st[lp] & CH_VAL)) { st[i] = (uint8_t)(s8 & ~CH_VAL); changed = true; }
            }
            __syncwarp();
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        // ---- C: layers opened lazily are written iff they enclose something that is: drawn colours and clips that
        //         are always written mark every BeginClip above them
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t s8 = st[i];
            if (!(s8 & CH_VAL) || (s8 & CH_TAG) == 2u || (s8 & (CH_TAG | CH_ELIDE)) == (1u | CH_ELIDE)) continue;
            for (uint32_t p = lpos[i]; p < CH_MISSING; p = lpos[p]) {
                const uint32_t sp = st[p];
                if (sp & CH_NE) break;   // whoever set it keeps climbing
                st[p] = (uint8_t)(sp | CH_NE);
            }
        }
        __syncwarp();
        // ---- D: commands, 32 hits at a time in scene order
        uint32_t pos_u = pos - ptcl_off[T];          // offset of the next command inside this tile's list
        uint32_t restart = 0, restart_rgba = 0, max_depth = 0;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            uint32_t nw = 0, s8 = 0, d = 0, lp = CH_NONE, sstart = 0;
            GGDrawRec r; r.tag = 0; r.parent = -1; r.a = 0; r.b = 0;
            GGTile t; t.backdrop = 0; t.seg_count = 0;
            bool emit = false;
            if (i < n) {
                s8 = st[i]; lp = lpos[i];
                const uint32_t tg = s8 & CH_TAG;
                if (s8 & CH_VAL) {
                    if (tg == 0u) emit = true;
                    else if (tg == 1u) emit = !(s8 & CH_ELIDE) || (s8 & CH_NE);
                    else if (lp < CH_MISSING) { const uint32_t sbg = st[lp]; emit = !(sbg & CH_ELIDE) || (sbg & CH_NE); }
                }
            }
            if (emit) {
                d = sb[i] >> 10;
                r = recs[d];
                const uint32_t tg = s8 & CH_TAG;
                if (tg == 1u) {
                    nw = 1u;
                    uint32_t depth = 1;   // written clips above this one + itself
                    for (uint32_t p = lp; p < CH_MISSING; p = lpos[p]) depth++;
                    max_depth = max(max_depth, depth);
                } else {
                    const GGHit h = list[sb[i] & 1023u];   // the record the backdrop pass wrote for this (draw, tile)
                    t.backdrop = h.backdrop; t.seg_count = h.seg_count; sstart = h.seg_start;
                    nw = tg == 0u ? (t.seg_count ? 6u : 3u) : (t.seg_count ? 7u : 4u);
                }
            }
            uint32_t inc = nw;
#pragma unroll
            for (int dl = 1; dl < 32; dl <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, inc, dl); if ((int)lane >= dl) inc += o; }
            uint32_t o = pos + inc - nw;
            if (emit) {
                const uint32_t tg = s8 & CH_TAG;
                const uint32_t after = pos_u + inc;   // list offset right after this command
                if (tg == 0u) {
                    if (t.seg_count) { ptcl[o++] = GG_CMD_FILL; ptcl[o++] = (t.seg_count << 1) | (r.b & 1u); ptcl[o++] = sstart; ptcl[o++] = (uint32_t)t.backdrop; }
                    else ptcl[o++] = GG_CMD_SOLID;
                    ptcl[o++] = (r.b & 2u) ? GG_CMD_GRAD : GG_CMD_COLOR; ptcl[o++] = r.a;
                    // restart point for fine: an opaque colour over the whole tile outside every clip
                    if (lp == CH_NONE && !t.seg_count && (r.a >> 24) == 255u && !(r.b & 2u) && after > restart) { restart = after; restart_rgba = r.a; }
                } else if (tg == 1u) {
                    ptcl[o++] = GG_CMD_BEGIN_CLIP;
                } else {
                    if (t.seg_count) { ptcl[o++] = GG_CMD_FILL; ptcl[o++] = (t.seg_count << 1); ptcl[o++] = sstart; ptcl[o++] = (uint32_t)t.backdrop; }
                    else ptcl[o++] = GG_CMD_SOLID;
                    ptcl[o++] = GG_CMD_END_CLIP; ptcl[o++] = r.a; ptcl[o++] = r.b;
                    // ... or the end of an outermost full-coverage layer whose compose mode wipes the backdrop: Clear always,
                    // Copy / SrcIn / DestIn / SrcOut / DestAtop when nothing was drawn inside
                    if (!t.seg_count && lpos[lp] == CH_NONE) {
                        const uint32_t mixm = (r.a >> 8) & 0xffu, comp = r.a & 0xffu;
                        const bool wipes_empty = comp == 1u || comp == 5u || comp == 6u || comp == 7u || comp == 10u;
                        if (mixm == 0u && (comp == 0u || (wipes_empty && !(st[lp] & CH_NE))) && after > restart) { restart = after; restart_rgba = 0; }
                    }
                }
            }
            const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
            pos += tot; pos_u += tot;
        }
        // the last qualifying command wins
#pragma unroll
        for (int dl = 16; dl > 0; dl >>= 1) {
            uint32_t orr = __shfl_xor_sync(0xffffffffu, restart, dl), oc = __shfl_xor_sync(0xffffffffu, restart_rgba, dl);
            uint32_t od = __shfl_xor_sync(0xffffffffu, max_depth, dl);
            if (orr > restart) { restart = orr; restart_rgba = oc; }
            max_depth = max(max_depth, od);
        }
        if (lane == 0) {
            ptcl[pos] = GG_CMD_END;
            ptcl_len[T] = pos + 1 - ptcl_off[T];
            restart_pt[2 * T] = restart; restart_pt[2 * T + 1] = restart_rgba;
            if (pos + 1 - ptcl_off[T] - max(restart, 1u) > GG_FINE_HEAVY_WORDS)
                spill_off[(cfg.band_y1 - cfg.band_y0) * cfg.width_in_tiles + atomicAdd(&bump->heavy, 1u)] = T;
            if (max_depth > GG_BLEND_STACK_SPLIT) {
                uint32_t lv = max_depth - GG_BLEND_STACK_SPLIT;
                uint32_t so = atomicAdd(&bump->spill, lv);
                if (so + lv > cfg.spill_cap) atomicOr(&bump->failed, GG_FAIL_SPILL); else spill_off[T] = so;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ launchers
uint32_t gg_launch_front(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s) {
    init_frame_kernel<<<GG_GRID(2), 256, 0, s>>>(cfg, b.path_bbox_ord, b.bump);
    // a3: pathtag scan. n is host-known here; the scan primitive wants it in device memory, so a
    // constant slot at the tail of the scene buffer carries it (word n_scene_words).
    const uint32_t* n_tag_words = b.scene + cfg.n_scene_words + 0;
    const uint32_t* n_draws = b.scene + cfg.n_scene_words + 1;
    const uint32_t* n_tag_bytes = b.scene + cfg.n_scene_words + 2;
    const uint32_t* n_paths = b.scene + cfg.n_scene_words + 3;
    gg_scan<GGPathMonoid>(s, n_tag_words, cfg.n_tag_words, LoadTagMonoid{b.scene + cfg.path_tag_base}, StoreTagMonoid{b.tag_monoids},
                          (GGPathMonoid*)b.scan_partials, (GGPathMonoid*)nullptr);
    // a4 + a5: draw scan, draw leaf, clip leaf fix-up
    gg_scan<GGDrawMonoid>(s, n_draws, cfg.n_draws, LoadDrawMonoid{b.scene + cfg.draw_tag_base}, StoreDrawMonoid{b.draw_monoids},
                          (GGDrawMonoid*)b.scan_partials, (GGDrawMonoid*)nullptr);
    draw_leaf_kernel<<<GG_GRID(2), 256, 0, s>>>(cfg, b.scene, b.draw_monoids, b.info, b.clip_inps, b.draw_recs);
    // a6: flatten (count, scan, emit)
    flatten_classify_kernel<<<GG_GRID(8), 256, 0, s>>>(cfg, b.scene, b.tag_monoids, b.line_count, b.curve_list, b.bump);
    flatten_subdivide_kernel<<<GG_GRID(6), 128, 0, s>>>(cfg, b.scene, b.tag_monoids, b.curve_list, b.line_count, b.esegs, b.bump);
    gg_scan<uint32_t>(s, n_tag_bytes, cfg.n_tag_bytes, LoadU32{b.line_count}, StoreU32Ex{b.line_off}, (uint32_t*)b.scan_partials, &b.bump->lines);
    flatten_line_emit_kernel<<<GG_GRID(8), 256, 0, s>>>(cfg, b.scene, b.tag_monoids, b.line_count, b.line_off, b.lines, b.path_bbox_ord, b.bump);
    flatten_eseg_emit_kernel<<<GG_GRID(8), 128, 0, s>>>(cfg, b.scene, b.tag_monoids, b.esegs, b.line_off, b.lines, b.path_bbox_ord, b.bump);
    // a7: per-path tile bbox + tile / row offsets (one packed scan)
    gg_scan<unsigned long long>(s, n_paths, cfg.n_paths, LoadPathTiles{cfg, b.path_bbox_ord}, StorePath{cfg, b.path_bbox_ord, b.paths, b.path_row_off},
                                (unsigned long long*)b.scan_partials, reinterpret_cast<unsigned long long*>(&b.bump->path_tiles));
    return 6 + 4 * 2;
}

uint32_t gg_launch_binning(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s) {
    zero_tiles_kernel<<<GG_GRID(4), 256, 0, s>>>(cfg, b.tiles, b.bump);
    uint32_t band_tiles = cfg.width_in_tiles * (cfg.band_y1 - cfg.band_y0);
    cudaMemsetAsync(b.tile_hits, 0, sizeof(unsigned long long) * band_tiles, s);
    path_count_kernel<<<GG_GRID(8), 256, 0, s>>>(cfg, b.lines, b.paths, b.tiles, b.seg_counts, b.bump);
    if (cfg.imp_words) cudaMemsetAsync(b.imp_seen, 0, sizeof(uint32_t) * (size_t)cfg.imp_words * cfg.n_implicit, s);
    tile_rows_kernel<0><<<GG_GRID(8), 256, 0, s>>>(cfg, b.scene, b.paths, b.path_row_off, b.tiles, b.draw_recs, b.tile_hits, nullptr, nullptr, nullptr, nullptr, b.imp_mask, b.imp_seen, b.bump);
    gg_scan<uint32_t>(s, &b.bump->path_tiles, cfg.tiles_cap, LoadTileCount{b.tiles}, StoreU32Ex{b.seg_start}, (uint32_t*)b.scan_partials, &b.bump->segments);
    path_tiling_kernel<<<GG_GRID(8), 256, 0, s>>>(cfg, b.seg_counts, b.lines, b.paths, b.tiles, b.seg_start, b.segments, b.bump);
    return 4 + 2;
}

uint32_t gg_launch_coarse(const GGConfig& cfg, const GGBuffers& b, cudaStream_t s) {
    const uint32_t* n_band_tiles = b.scene + cfg.n_scene_words + 4;
    uint32_t band_tiles = cfg.width_in_tiles * (cfg.band_y1 - cfg.band_y0);
    gg_scan<unsigned long long>(s, n_band_tiles, band_tiles, LoadTileHits{b.tile_hits},
                                StoreTileHits{b.hit_off, b.hit_cnt, b.ptcl_off, b.hit_cursor, b.spill_off},
                                (unsigned long long*)b.scan_partials, reinterpret_cast<unsigned long long*>(&b.bump->hits));
    tile_rows_kernel<1><<<GG_GRID(8), 256, 0, s>>>(cfg, b.scene, b.paths, b.path_row_off, b.tiles, b.draw_recs, b.tile_hits, b.hit_off, b.hit_cursor, b.hits, b.seg_start, b.imp_mask, b.imp_seen, b.bump);
    coarse_kernel<<<GG_GRID(8), COARSE_WARPS * 32, 0, s>>>(cfg, b.draw_recs,
                                                           b.hit_off, b.hit_cnt, b.hits, b.ptcl_off, b.ptcl_len, b.ptcl, b.spill_off, b.restart_pt, b.bump);
    return 2 + 2;
}
