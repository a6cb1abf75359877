/*
 * oracle/twin.h -- TEST INFRASTRUCTURE ONLY (never linked into libggcuda.so).
 *
 * CPU restatement ("twin oracle") of gg's Vello-style tile pipeline,
 * reference: internal/gpu/tilecompute (all .go files) in gogpu/gg.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef GG_ORACLE_TWIN_H
#define GG_ORACLE_TWIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* tilecompute/types.go:13-47 */
typedef struct { uint32_t path_ix; float p0[2]; float p1[2]; } ot_line_soup;     /* 20 B */
typedef struct { uint32_t bbox[4]; uint32_t tiles; } ot_path;                    /* 20 B */
typedef struct { int32_t backdrop; uint32_t seg_count_or_ix; } ot_tile;          /*  8 B */
typedef struct { uint32_t line_ix; uint32_t counts; } ot_segment_count;          /*  8 B */
typedef struct { float p0[2]; float p1[2]; float y_edge; } ot_path_segment;      /* 20 B */

/* tilecompute/pathtag.go:16-22, draw_leaf.go:17-22 */
typedef struct { uint32_t trans_ix, path_seg_ix, path_seg_offset, style_ix, path_ix; } ot_path_monoid;
typedef struct { uint32_t path_ix, clip_ix, scene_offset, info_offset; } ot_draw_monoid;

/* tilecompute/types.go:70-95 : SceneElement flattened to arrays */
enum { OT_ELEM_DRAW = 0, OT_ELEM_BEGIN_CLIP = 1, OT_ELEM_END_CLIP = 2 };
typedef struct {
    uint32_t type;        /* OT_ELEM_* */
    uint32_t line_start;  /* first line in the lines array */
    uint32_t line_count;
    uint8_t  color[4];    /* straight RGBA (draw) */
    uint32_t even_odd;    /* fill rule (draw) */
    uint32_t blend;       /* begin clip */
    float    alpha;       /* begin clip */
    uint32_t packed_rgba; /* draw: premultiplied packed colour, used instead of `color` when OT_ELEM_PACKED is set in type */
} ot_element;
#define OT_ELEM_PACKED 0x100u
#define OT_ELEM_GRADIENT 0x200u   /* draw: packed_rgba is an index into the gradient table (ggcuda's DrawTag 0x444) */
#define OT_ELEM_TYPE(t) ((t) & 0xffu)

/* scene_encode.go:52-62 */
typedef struct {
    uint32_t n_draw_objects, n_paths, n_clips;
    uint32_t path_tag_base, path_data_base, draw_tag_base, draw_data_base, transform_base, style_base;
} ot_layout;

/* Output of the full coarse pass (coarse.go:17-43). All arrays malloc'd; free with ot_coarse_free. */
typedef struct {
    int width_in_tiles, height_in_tiles;
    uint32_t n_paths;        ot_path *paths;
    uint32_t n_tiles;        ot_tile *tiles;            /* per-path tiles, concatenated */
    uint32_t n_segments;     ot_path_segment *segments;
    uint32_t *path_seg_base; uint32_t *path_total_segs;
    uint32_t *ptcl_offsets;  /* [wt*ht + 1] word offsets into ptcl_words */
    uint32_t *ptcl_words;
    /* packed scene + scan results, kept for parity checks */
    uint32_t n_scene_words;  uint32_t *scene; ot_layout layout;
    uint32_t n_tag_words;    ot_path_monoid *tag_monoids;
    ot_draw_monoid *draw_monoids;
    uint32_t n_info;         uint32_t *info;
    /* ggcuda's gradient table (16-word records | stops | ramps), copied from behind the packed scene; NULL without gradients */
    uint32_t n_gtab_words;   uint32_t *gtab;
} ot_coarse;

/* ---- flatten (flatten.go, euler.go, path_convert.go) ---- */
/* cubics: n x 8 floats (p0,p1,p2,p3). Returns number of lines written (<= cap) or needed count if > cap. */
uint32_t ot_flatten_fill(const float *cubics, uint32_t n, ot_line_soup *out, uint32_t cap);
/* path_convert.go:29-112: verbs (0 MoveTo 1 LineTo 2 QuadTo 3 CubicTo 4 Close) + f64 coords.
 * auto_close != 0 additionally closes open subpaths (ggcuda ingest rule). */
uint32_t ot_flatten_path(const uint8_t *verbs, uint32_t n_verbs, const double *coords,
                         int auto_close, ot_line_soup *out, uint32_t cap);

/* stroke expansion of one centre-line segment (ggcuda's definition, see twin.c; parity unpinned against gg) */
uint32_t ot_stroke_segment(const float *seg, int kind, int role, const float *next, int next_kind,
                           float width, float miter_limit, int join, int cap, ot_line_soup *out, uint32_t out_cap);

/* ---- monoids (pathtag.go:26-63, draw_leaf.go:29-41) ---- */
void ot_path_monoid_new(uint32_t tag_word, ot_path_monoid *out);
void ot_draw_monoid_new(uint32_t tag, ot_draw_monoid *out);

/* ---- single-path pipeline (rasterizer.go:27-170) ---- */
/* alpha out: w*h floats */
void ot_rasterize(const ot_line_soup *lines, uint32_t n_lines, int even_odd, int w, int h, float *alpha);
/* rasterizer.go:176-224 + compositor.go: per-path composite; out: w*h*4 straight RGBA8 */
void ot_rasterize_scene(const uint8_t bg[4], const ot_element *elems, uint32_t n_elems,
                        const ot_line_soup *lines, int w, int h, uint8_t *out);

/* ---- full PTCL pipeline (rasterizer.go:321-422, coarse.go, fine.go) ---- */
extern int ot_style_per_path;   /* 0 = reference behaviour (default), 1 = one style word per path marker */
ot_coarse *ot_coarse_run(const ot_element *elems, uint32_t n_elems,
                         const ot_line_soup *lines, int w, int h);
void ot_coarse_free(ot_coarse *c);
/* fine for one tile -> 256 x 4 premultiplied float32 (fine.go:40-187) */
/* the same with the tile's pixel origin and the gradient table (CmdGrad evaluates gg's ColorAt at pixel centres) */
void ot_fine_tile_at(const uint32_t *ptcl, uint32_t n_words, const ot_path_segment *segs, uint32_t n_segs,
                     const float bg[4], float *rgba_out, int origin_x, int origin_y, const uint32_t *gtab);
void ot_fine_tile(const uint32_t *ptcl, uint32_t n_words, const ot_path_segment *segs, uint32_t n_segs,
                  const float bg[4], float *rgba_out);
/* whole frame: out_straight (w*h*4, premulToStraightU8 as rasterizer.go:405-414) and/or
 * out_premul (w*h*4, packed like fine.wgsl:305-323 u8(clamp(c)*255+0.5)); either may be NULL. */
void ot_fine_frame(const ot_coarse *c, const uint8_t bg[4], int w, int h,
                   uint8_t *out_straight, uint8_t *out_premul);

/* path stages on an explicit line list + path table (path_count.go, path_tiling.go) */
uint32_t ot_path_count(const ot_line_soup *lines, uint32_t n_lines, const ot_path *paths,
                       ot_tile *tiles, ot_segment_count *seg_counts);
void ot_path_tiling(const ot_segment_count *seg_counts, uint32_t n_seg_counts,
                    const ot_line_soup *lines, const ot_path *paths, const ot_tile *tiles,
                    ot_path_segment *segments);
void ot_line_bbox(const ot_line_soup *lines, uint32_t n, int w, int h, uint32_t bbox[4]);

/* ---- blending (oracle/blend.c) ---- */
/* bit-exact restatement of gg's byte blend functions, selected by scene.BlendMode; premultiplied RGBA8 */
void ot_blend_bytes(uint32_t scene_mode, const uint8_t s[4], const uint8_t d[4], uint8_t out[4]);
/* float32 layer composite applied at CmdEndClip: bg (blend) fg, premultiplied; blend = (mix << 8) | compose */
void ot_blend_f32(uint32_t blend, const float bg[4], const float fg[4], float out[4]);

/* ---- whole pipeline from ggcuda's packed scene (the a2 format: Vello path tags incl. 0x0C MoveTo,
 *      quads/cubics, real transforms). CPU restatement of the product's device stages a6..a13:
 *      flatten every curve (transform in f32 as scene/encoding.go:348-350, quad elevation as
 *      path_convert.go:60-72, FlattenFill), then ot_coarse_run + fine on `threads` host threads.
 *      layout: the 13 words of gg_b200/csrc/host_scene.h HostScene::Layout. out_premul: w*h*4. ---- */
typedef struct { double t_flatten, t_coarse, t_fine; uint32_t n_lines, n_segments, n_ptcl_words; } ot_timing;
int ot_render_packed(const uint32_t *scene, const uint32_t *layout13, int w, int h, const uint8_t bg_premul[4],
                     int threads, uint8_t *out_premul, ot_timing *timing);
ot_coarse *ot_coarse_from_packed(const uint32_t *scene, const uint32_t *layout13, int w, int h);
/* flatten stage alone: lines (path_ix set) for every path of the packed scene; returns count (cap as ot_flatten_fill) */
uint32_t ot_flatten_packed(const uint32_t *scene, const uint32_t *layout13, ot_line_soup *out, uint32_t cap);
/* multi-threaded ot_fine_frame */
void ot_fine_frame_mt(const ot_coarse *c, const float bg_premul[4], int w, int h, int threads, uint8_t *out_premul);

#ifdef __cplusplus
}
#endif
uint32_t ot_scan_stages(const uint32_t *scene, uint32_t n_scene_words, uint32_t path_tag_base, uint32_t n_tag_words,
                        uint32_t draw_tag_base, uint32_t draw_data_base, uint32_t n_draw, uint32_t n_clips,
                        ot_path_monoid *tag_monoids, ot_draw_monoid *dm, uint32_t *info, int32_t *clip_inps_out);
/* flatten.go:19 (0.25 px); a test may tighten it to measure what the tolerance costs against gg's CPU path */
extern float ot_flatten_tol;
/* 1 = reproduce coarse.go:425 (even-odd tile with an even non-zero backdrop and no segments is painted solid) */
extern int ot_evenodd_solid_quirk;
/* 1 = fine truncates the running colour to 8 bits after every CmdColor, as gg's CPU pixmap does (pixmap.go:218-228) */
extern int ot_truncate_per_draw;

#endif
