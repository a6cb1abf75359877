// gg_b200/csrc/host_scene.cpp -- see host_scene.h.
#include "host_scene.h"

#include <math.h>
#include <string.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../include/ggcuda.h"

enum { PT_LINETO = 0x09, PT_QUADTO = 0x0A, PT_CUBICTO = 0x0B, PT_MOVETO = 0x0C, PT_PATH = 0x10, PT_TRANSFORM = 0x20, PT_STYLE = 0x40,
       PT_MARKER = 0x80, PT_MARKER_MOVE = 0x8C };   // stroke markers, see stroke.cuh
enum { STYLE_STROKE = 0x01, STYLE_EVEN_ODD = 0x02 };
enum { DT_COLOR = 0x44, DT_BEGIN_CLIP = 0x9, DT_END_CLIP = 0x21,
       DT_GRADIENT = 0x444 };   // DrawTagColor's monoid increments (one scene word, one info word) + bit 10: the word is a gradient index

// scene.Tag values, scene/tag.go:25-110
enum {
    ST_TRANSFORM = 0x01, ST_SET_AA = 0x02, ST_BEGIN_PATH = 0x10, ST_MOVE_TO = 0x11, ST_LINE_TO = 0x12, ST_QUAD_TO = 0x13,
    ST_CUBIC_TO = 0x14, ST_CLOSE_PATH = 0x16, ST_END_PATH = 0x17, ST_FILL = 0x20, ST_STROKE = 0x21, ST_FILL_ROUND_RECT = 0x22,
    ST_PUSH_LAYER = 0x30, ST_POP_LAYER = 0x31, ST_BEGIN_CLIP = 0x40, ST_END_CLIP = 0x41, ST_BRUSH = 0x50, ST_IMAGE = 0x51, ST_TEXT = 0x60
};

static const float IDENTITY[6] = {1, 0, 0, 0, 1, 0};   // scene.Affine: x' = A x + B y + C, y' = D x + E y + F

uint8_t gg_clamp_u8(double v) {   // path_convert.go:131-140
    double x = v * 255.0 + 0.5;
    if (x < 0) return 0;
    if (x > 255) return 255;
    return (uint8_t)x;
}
uint32_t gg_pack_color_straight(const uint8_t c[4]) {   // scene_encode.go:162-168 (float32, +0.5)
    float a = (float)c[3] / 255.0f;
    uint32_t r = (uint32_t)(int64_t)((float)c[0] * a + 0.5f);
    uint32_t g = (uint32_t)(int64_t)((float)c[1] * a + 0.5f);
    uint32_t b = (uint32_t)(int64_t)((float)c[2] * a + 0.5f);
    return r | (g << 8) | (b << 16) | ((uint32_t)c[3] << 24);
}
uint32_t gg_blend_word(uint32_t m) {   // scene/encoding.go:17-48 -> (mix << 8) | compose
    if (m < 16) return (m << 8) | 3u;  // Normal..Luminosity mix with SrcOver compose
    if (m <= 28) return m - 16;        // Clear..Plus compose with Normal mix
    return 3u;
}

void HostScene::clear(uint32_t w, uint32_t h) {
    width = w; height = h;
    tags.clear(); path_data.clear(); draw_tags.clear(); draw_data.clear(); styles.clear(); transforms.clear();
    clip_aux.clear(); clip_stack.clear(); clip_kind.clear(); clip_bb.clear();
    next_clip_bb[0] = next_clip_bb[1] = -3.0e38f; next_clip_bb[2] = next_clip_bb[3] = 3.0e38f;
    n_paths = n_clips = n_seg_tags = n_implicit = n_culled = 0;
    grad_recs.clear(); grad_stops.clear(); images.clear(); image_words.clear(); n_gradients = 0;
    have_transform = false; in_path = false; has_move = false;
}

// True if every control point (the curve lies inside their hull), moved by `reach`, misses the band's pixel rows.
// A pixel of slack covers the flattening tolerance and the float32 rounding of the device-side transform.
bool HostScene::outside_band(const float t[6], const float* c, size_t n_coords, float reach) const {
    if (!cull || n_coords < 2) return false;
    if (t[3] == 0.0f) {
        // No shear into y (every transform of an axis-aligned scene): y' = fl(fl(t4 * y) + t5) is monotone in y, so the extremes
        // of the raw ordinates map to the extremes of the transformed ones, bit for bit what the general loop below computes --
        // two compares per point instead of two multiplies, two adds and two compares. A non-finite x makes the general
        // form's 0 * x a NaN (never culled): kept.
        float lo = 3.0e38f, hi = -3.0e38f, bad = 0.0f;
        for (size_t k = 0; k + 1 < n_coords; k += 2) {
            const float x = c[k], y = c[k + 1];
            bad += x - x;                       // 0 for finite x, NaN otherwise
            if (!(y == y)) return false;
            lo = y < lo ? y : lo; hi = y > hi ? y : hi;
        }
        if (!(bad == 0.0f)) return false;
        float a = t[4] * lo + t[5], b = t[4] * hi + t[5];
        if (!(a == a) || !(b == b)) return false;
        if (a > b) { const float sw = a; a = b; b = sw; }
        return b + reach + 1.0f < cull_lo || a - reach - 1.0f > cull_hi;
    }
    float lo = 3.0e38f, hi = -3.0e38f;
    for (size_t k = 0; k + 1 < n_coords; k += 2) {
        float y = t[3] * c[k] + t[4] * c[k + 1] + t[5];
        if (!(y == y)) return false;
        lo = y < lo ? y : lo; hi = y > hi ? y : hi;
    }
    return hi + reach + 1.0f < cull_lo || lo - reach - 1.0f > cull_hi;
}

void HostScene::begin_path(const float t[6], bool even_odd) {
    if (!have_transform || memcmp(t, last_transform, sizeof(float) * 6) != 0) {
        tags.push_back(PT_TRANSFORM);
        transforms.insert(transforms.end(), t, t + 6);
        memcpy(last_transform, t, sizeof(float) * 6);
        have_transform = true;
    }
    tags.push_back(PT_STYLE);
    styles.push_back(even_odd ? 0x02u : 0u);   // scene_encode.go:110-114 (+ two stroke words, unused for fills)
    styles.push_back(0); styles.push_back(0);
    in_path = true; has_move = false;
    memcpy(path_t, t, sizeof path_t);
}
void HostScene::move_to(float x, float y) {
    // An open subpath is closed implicitly, as every CPU filler in gg does (the Vello
    // path of the reference leaves it open, path_convert.go:44-49).
    if (has_move && (cur[0] != start[0] || cur[1] != start[1])) line_to(start[0], start[1]);
    tags.push_back(PT_MOVETO);
    path_data.push_back(x); path_data.push_back(y);
    cur[0] = start[0] = x; cur[1] = start[1] = y; has_move = true;
}
void HostScene::line_to(float x, float y) {
    if (!has_move) return;   // path_convert.go:52-54
    tags.push_back(PT_LINETO);
    path_data.push_back(x); path_data.push_back(y);
    cur[0] = x; cur[1] = y; n_seg_tags++;
}
void HostScene::quad_to(float cx, float cy, float x, float y) {
    if (!has_move) return;
    tags.push_back(PT_QUADTO);
    path_data.push_back(cx); path_data.push_back(cy); path_data.push_back(x); path_data.push_back(y);
    cur[0] = x; cur[1] = y; n_seg_tags++;
}
void HostScene::cubic_to(float c1x, float c1y, float c2x, float c2y, float x, float y) {
    if (!has_move) return;
    tags.push_back(PT_CUBICTO);
    path_data.push_back(c1x); path_data.push_back(c1y); path_data.push_back(c2x); path_data.push_back(c2y);
    path_data.push_back(x); path_data.push_back(y);
    cur[0] = x; cur[1] = y; n_seg_tags++;
}
void HostScene::close() {   // path_convert.go:86-92
    if (has_move && (cur[0] != start[0] || cur[1] != start[1])) line_to(start[0], start[1]);
    cur[0] = start[0]; cur[1] = start[1];
}
void HostScene::end_path() {
    if (has_move && (cur[0] != start[0] || cur[1] != start[1])) line_to(start[0], start[1]);
    tags.push_back(PT_PATH);
    n_paths++;
    in_path = false; has_move = false;
}
void HostScene::add_verbs(const uint8_t* verbs, uint32_t n_verbs, const double* c, uint32_t n_coords) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n_verbs; i++) {
        switch (verbs[i]) {
        case GGCUDA_VERB_MOVE: if (k + 2 > n_coords) return; move_to((float)c[k], (float)c[k + 1]); k += 2; break;
        case GGCUDA_VERB_LINE: if (k + 2 > n_coords) return; line_to((float)c[k], (float)c[k + 1]); k += 2; break;
        case GGCUDA_VERB_QUAD: if (k + 4 > n_coords) return; quad_to((float)c[k], (float)c[k + 1], (float)c[k + 2], (float)c[k + 3]); k += 4; break;
        case GGCUDA_VERB_CUBIC: if (k + 6 > n_coords) return;
            cubic_to((float)c[k], (float)c[k + 1], (float)c[k + 2], (float)c[k + 3], (float)c[k + 4], (float)c[k + 5]); k += 6; break;
        case GGCUDA_VERB_CLOSE: close(); break;
        default: break;
        }
    }
}

// Fill path straight from verb bytes + float coordinates: the hot loop of scene ingest. Tags and coordinates are
// written through raw pointers into space reserved for the worst case (every MoveTo / Close / the path end may add a
// closing LineTo); the rules are those of move_to / line_to / close / end_path above.
// Geometry of a fill path: tags and coordinates written through raw pointers into space the caller reserved for the worst
// case (2 * n_verbs + 3 tags, n_coords + 2 * (n_verbs + 2) floats). A pure function of its inputs: the ingest threads run it
// side by side. Returns the number of segment tags.
static uint32_t build_fill_geometry(const uint8_t* verbs, size_t n_verbs, const float* c, size_t n_coords, const uint8_t* verb_map,
                                    uint8_t*& tp, float*& dp) {
    bool have = false;
    float sx = 0, sy = 0, cx = 0, cy = 0;
    uint32_t nseg = 0;
    size_t k = 0;
#define GG_CLOSE_SUBPATH() do { if (have && (cx != sx || cy != sy)) { *tp++ = PT_LINETO; *dp++ = sx; *dp++ = sy; nseg++; } } while (0)
    for (size_t i = 0; i < n_verbs; i++) {
        switch (verb_map ? verb_map[verbs[i]] : verbs[i]) {
        case GGCUDA_VERB_MOVE:
            if (k + 2 > n_coords) { i = n_verbs; break; }
            GG_CLOSE_SUBPATH();
            *tp++ = PT_MOVETO; sx = cx = *dp++ = c[k]; sy = cy = *dp++ = c[k + 1]; k += 2; have = true;
            break;
        case GGCUDA_VERB_LINE:
            if (k + 2 > n_coords) { i = n_verbs; break; }
            if (have) { *tp++ = PT_LINETO; cx = *dp++ = c[k]; cy = *dp++ = c[k + 1]; nseg++; }
            k += 2; break;
        case GGCUDA_VERB_QUAD:
            if (k + 4 > n_coords) { i = n_verbs; break; }
            if (have) { *tp++ = PT_QUADTO; memcpy(dp, c + k, 16); dp += 4; cx = c[k + 2]; cy = c[k + 3]; nseg++; }
            k += 4; break;
        case GGCUDA_VERB_CUBIC:
            if (k + 6 > n_coords) { i = n_verbs; break; }
            if (have) { *tp++ = PT_CUBICTO; memcpy(dp, c + k, 24); dp += 6; cx = c[k + 4]; cy = c[k + 5]; nseg++; }
            k += 6; break;
        case GGCUDA_VERB_CLOSE:   // path_convert.go:86-92
            GG_CLOSE_SUBPATH();
            cx = sx; cy = sy;
            break;
        default: break;
        }
    }
    GG_CLOSE_SUBPATH();
#undef GG_CLOSE_SUBPATH
    return nseg;
}
static inline size_t fill_tag_bound(size_t n_verbs) { return 2 * n_verbs + 3; }
static inline size_t fill_data_bound(size_t n_verbs, size_t n_coords) { return n_coords + 2 * (n_verbs + 2); }

void HostScene::fill_verbs(const float t[6], bool even_odd, const uint8_t* verbs, size_t n_verbs, const float* c, size_t n_coords,
                           const uint8_t* verb_map) {
    begin_path(t, even_odd);
    const size_t t0 = tags.size(), d0 = path_data.size();
    tags.resize(t0 + fill_tag_bound(n_verbs));
    path_data.resize(d0 + fill_data_bound(n_verbs, n_coords));
    uint8_t* tp = tags.data() + t0;
    float* dp = path_data.data() + d0;
    uint32_t nseg = build_fill_geometry(verbs, n_verbs, c, n_coords, verb_map, tp, dp);
    *tp++ = PT_PATH;
    tags.resize((size_t)(tp - tags.data()));
    path_data.resize((size_t)(dp - path_data.data()));
    n_seg_tags += nseg;
    n_paths++;
    in_path = false; has_move = false;
}

// Centre line of a stroked path for the device-side expander. Per subpath: MoveTo, the segments (those whose points
// all coincide are dropped, so every segment has a tangent), then a marker: a copy of the first segment flagged
// PT_MARKER -- the last segment reads the tangent of its join from it when the subpath is closed; for an open
// subpath a PT_MARKER_MOVE back to the first point precedes it and the marker segment draws the start cap.
// Geometry of a stroke's centre line (see above); same contract as build_fill_geometry, bounds 4 * n_verbs + 8 tags and
// n_coords + 12 * (n_verbs + 2) floats.
static uint32_t build_stroke_geometry(const uint8_t* verbs, size_t n_verbs, const float* c, size_t n_coords, const StrokeStyleHost& st,
                                      const uint8_t* verb_map, uint8_t*& tp, float*& dp) {
    bool have = false, implicit = false;   // implicit: current point left behind by a Close, not set by a MoveTo
    float sx = 0, sy = 0, cx = 0, cy = 0;
    uint8_t* sub_t = nullptr; float* sub_d = nullptr;   // where the open subpath's MoveTo was written
    uint32_t nseg = 0, sub_seg = 0;
    auto open_sub = [&]() { if (!sub_t) { sub_t = tp; sub_d = dp; *tp++ = PT_MOVETO; *dp++ = sx; *dp++ = sy; sub_seg = 0; } };
    auto flush = [&](bool closed) {
        if (!have) return;
        have = false;
        if (!sub_t) sub_seg = 0;
        if (sub_seg == 0 && implicit) { if (sub_t) { tp = sub_t; dp = sub_d; sub_t = nullptr; } return; }
        if (closed && sub_seg && (cx != sx || cy != sy)) { *tp++ = PT_LINETO; *dp++ = sx; *dp++ = sy; sub_seg++; }
        if (sub_seg == 0) {   // a dot: round and square caps draw it (software.go strokes a zero-length subpath the same way)
            if (st.cap == GGCUDA_CAP_BUTT) { if (sub_t) { tp = sub_t; dp = sub_d; sub_t = nullptr; } return; }
            open_sub();
            *tp++ = PT_LINETO; *dp++ = sx + 1.0f / 1024.0f; *dp++ = sy; sub_seg++;
            closed = false;
        }
        if (!closed) { *tp++ = PT_MARKER_MOVE; *dp++ = sx; *dp++ = sy; }
        uint8_t first = sub_t[1];
        uint32_t nf = 2u * (first & 3u);
        *tp++ = (uint8_t)(first | PT_MARKER);
        memcpy(dp, sub_d + 2, sizeof(float) * nf); dp += nf;
        nseg += sub_seg + 1;
        sub_t = nullptr; sub_seg = 0;
    };
    size_t k = 0;
    for (size_t i = 0; i < n_verbs; i++) {
        switch (verb_map ? verb_map[verbs[i]] : verbs[i]) {
        case GGCUDA_VERB_MOVE:
            if (k + 2 > n_coords) { i = n_verbs; break; }
            flush(false);
            sx = cx = c[k]; sy = cy = c[k + 1]; k += 2; have = true; implicit = false; break;
        case GGCUDA_VERB_LINE:
            if (k + 2 > n_coords) { i = n_verbs; break; }
            if (have && (c[k] != cx || c[k + 1] != cy)) { open_sub(); *tp++ = PT_LINETO; cx = *dp++ = c[k]; cy = *dp++ = c[k + 1]; sub_seg++; }
            k += 2; break;
        case GGCUDA_VERB_QUAD:
            if (k + 4 > n_coords) { i = n_verbs; break; }
            if (have && (c[k] != cx || c[k + 1] != cy || c[k + 2] != cx || c[k + 3] != cy)) {
                open_sub(); *tp++ = PT_QUADTO; memcpy(dp, c + k, 16); dp += 4; cx = c[k + 2]; cy = c[k + 3]; sub_seg++;
            }
            k += 4; break;
        case GGCUDA_VERB_CUBIC:
            if (k + 6 > n_coords) { i = n_verbs; break; }
            if (have && (c[k] != cx || c[k + 1] != cy || c[k + 2] != cx || c[k + 3] != cy || c[k + 4] != cx || c[k + 5] != cy)) {
                open_sub(); *tp++ = PT_CUBICTO; memcpy(dp, c + k, 24); dp += 6; cx = c[k + 4]; cy = c[k + 5]; sub_seg++;
            }
            k += 6; break;
        case GGCUDA_VERB_CLOSE:
            if (have) { float x0 = sx, y0 = sy; flush(true); sx = cx = x0; sy = cy = y0; have = true; implicit = true; }   // path.go: Close moves back to the start
            break;
        default: break;
        }
    }
    flush(false);
    return nseg;
}
static inline size_t stroke_tag_bound(size_t n_verbs) { return 4 * n_verbs + 8; }
static inline size_t stroke_data_bound(size_t n_verbs, size_t n_coords) { return n_coords + 12 * (n_verbs + 2); }
static inline uint32_t stroke_style_flags(const StrokeStyleHost& st) { return STYLE_STROKE | ((uint32_t)(st.join & 3) << 2) | ((uint32_t)(st.cap & 3) << 4); }

void HostScene::stroke_path(const float t[6], const uint8_t* verbs, size_t n_verbs, const float* c, size_t n_coords, const StrokeStyleHost& st,
                            const uint8_t* verb_map) {
    begin_path(t, false);
    float w = (float)st.width, ml = (float)st.miter_limit;
    if (!(w > 0.0f)) { end_path(); return; }
    uint32_t wb, mb; memcpy(&wb, &w, 4); memcpy(&mb, &ml, 4);
    styles[styles.size() - 3] = stroke_style_flags(st); styles[styles.size() - 2] = wb; styles[styles.size() - 1] = mb;
    const size_t t0 = tags.size(), d0 = path_data.size();
    tags.resize(t0 + stroke_tag_bound(n_verbs));
    path_data.resize(d0 + stroke_data_bound(n_verbs, n_coords));
    uint8_t* tp = tags.data() + t0;
    float* dp = path_data.data() + d0;
    uint32_t nseg = build_stroke_geometry(verbs, n_verbs, c, n_coords, st, verb_map, tp, dp);
    tags.resize((size_t)(tp - tags.data()));
    path_data.resize((size_t)(dp - path_data.data()));
    n_seg_tags += nseg;
    has_move = false;
    end_path();
}

void HostScene::append_stroke(const StrokeSink& k) {
    // the outline loops are complete subpaths in device space (the path was begun with the identity transform)
    tags.insert(tags.end(), k.tags.begin(), k.tags.end());
    path_data.insert(path_data.end(), k.data.begin(), k.data.end());
    n_seg_tags += k.n_seg;
    has_move = false;
}

void HostScene::set_next_clip_bounds(const float t[6], const uint8_t* verbs, size_t n_verbs, const float* c, size_t n_coords, const uint8_t* verb_map) {
    float bb[4] = {3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f};
    size_t k = 0;
    for (size_t i = 0; i < n_verbs; i++) {
        uint8_t v = verb_map ? verb_map[verbs[i]] : verbs[i];
        size_t np = v == GGCUDA_VERB_MOVE || v == GGCUDA_VERB_LINE ? 1 : v == GGCUDA_VERB_QUAD ? 2 : v == GGCUDA_VERB_CUBIC ? 3 : 0;
        if (k + 2 * np > n_coords) break;
        for (size_t j = 0; j < np; j++, k += 2) {   // control points bound the curve
            float x = t[0] * c[k] + t[1] * c[k + 1] + t[2], y = t[3] * c[k] + t[4] * c[k + 1] + t[5];
            bb[0] = std::min(bb[0], x); bb[1] = std::min(bb[1], y); bb[2] = std::max(bb[2], x); bb[3] = std::max(bb[3], y);
        }
    }
    if (!(bb[0] <= bb[2])) { bb[0] = bb[1] = bb[2] = bb[3] = 0; }   // empty clip path: nothing shows
    memcpy(next_clip_bb, bb, sizeof bb);
}

void HostScene::draw_color(uint32_t rgba_premul) {
    draw_tags.push_back(DT_COLOR);
    draw_data.push_back(rgba_premul);
    clip_aux.push_back(clip_stack.empty() ? -1 : clip_stack.back());
    clip_aux.push_back(0);
}
// gg's colour stops interpolate in linear light: sRGB -> linear, lerp, linear -> sRGB, float32 (gradient.go:62-101,
// internal/color/convert.go:8-23); alpha is linear. The device does the lerp and the way back per pixel (fine.cu grad_color);
// the linear-light value of every stop is computed here, once.
static float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : (float)pow((double)((s + 0.055f) / 1.055f), 2.4); }
void HostScene::draw_gradient(int kind, const double geom[6], const double* stops_in, uint32_t n_stops, int extend) {
    // sortStops (gradient.go:29-40): by offset (insertion sort: stable; sort.Slice makes no promise for equal offsets)
    std::vector<double> st(stops_in, stops_in + 5 * (size_t)n_stops);
    for (uint32_t i = 1; i < n_stops; i++) {
        double key[5]; memcpy(key, &st[5 * i], sizeof key);
        int j = (int)i - 1;
        while (j >= 0 && st[5 * j] > key[0]) { memcpy(&st[5 * (j + 1)], &st[5 * j], sizeof key); j--; }
        memcpy(&st[5 * (j + 1)], key, sizeof key);
    }
    uint32_t rec[16] = {0};
    rec[0] = kind <= 1 ? (uint32_t)kind : (uint32_t)kind + 2u;   // table kinds: 0 linear, 1 radial, 4 sweep, 5 focal radial (2 / 3: SDF round rect / image)
    rec[1] = (uint32_t)extend; rec[2] = n_stops;
    rec[3] = (uint32_t)grad_stops.size(); rec[4] = 0;
    float g[6];
    for (int k = 0; k < 6; k++) g[k] = (float)geom[k];
    memcpy(rec + 5, g, sizeof g);
    grad_recs.insert(grad_recs.end(), rec, rec + 16);
    for (uint32_t i = 0; i < n_stops; i++) {   // 8 floats per stop: offset, straight sRGB r g b, a, linear-light r g b
        for (int k = 0; k < 5; k++) grad_stops.push_back((float)st[5 * i + k]);
        for (int k = 0; k < 3; k++) grad_stops.push_back(srgb_to_linear((float)st[5 * i + 1 + k]));
    }
    draw_tags.push_back(DT_GRADIENT);
    draw_data.push_back(n_gradients++);
    clip_aux.push_back(clip_stack.empty() ? -1 : clip_stack.back());
    clip_aux.push_back(0);
}
void HostScene::draw_sdf_round_rect(float cx, float cy, float half_w, float half_h, float radius, uint32_t rgba_premul) {
    uint32_t rec[16] = {0};
    rec[0] = 2u;
    float g[6] = {cx, cy, half_w, half_h, radius, 0.0f};
    memcpy(rec + 5, g, sizeof g);
    rec[10] = rgba_premul;
    grad_recs.insert(grad_recs.end(), rec, rec + 16);
    draw_tags.push_back(DT_GRADIENT);
    draw_data.push_back(n_gradients++);
    clip_aux.push_back(clip_stack.empty() ? -1 : clip_stack.back());
    clip_aux.push_back(0);
}
int HostScene::add_image(uint32_t w, uint32_t h, const uint8_t* premul_rgba) {
    if (w == 0 || h == 0 || w > 16384 || h > 16384 || !premul_rgba) return -1;
    HostImage im = {w, h, image_words.size()};
    image_words.resize(im.off + (size_t)w * h);
    memcpy(image_words.data() + im.off, premul_rgba, (size_t)w * h * 4);
    images.push_back(im);
    return (int)images.size() - 1;
}
bool HostScene::draw_image(uint32_t index, const float t[6]) {
    if (index >= images.size()) return false;
    const HostImage& im = images[index];
    // inverse affine and the canvas-space box of the image, in float32 and in the reference's order of operations
    // (renderer.go:1104-1137); a degenerate transform draws nothing
    const float A = t[0], B = t[1], C = t[2], D = t[3], E = t[4], F = t[5];
    const float det = A * E - B * D;
    if (det == 0.0f || !(det == det)) return true;
    const float inv_det = 1.0f / det;
    const float inv[6] = {E * inv_det, -B * inv_det, (B * F - E * C) * inv_det, -D * inv_det, A * inv_det, (D * C - A * F) * inv_det};
    const float cw = (float)im.w, ch = (float)im.h;
    const float cx[4] = {0, cw, cw, 0}, cy[4] = {0, 0, ch, ch};
    float x0 = 0, y0 = 0, x1 = 0, y1 = 0;   // the reference seeds the box with the UNtransformed corner (0, 0)
    for (int k = 0; k < 4; k++) {
        const float px = A * cx[k] + B * cy[k] + C, py = D * cx[k] + E * cy[k] + F;
        x0 = std::min(x0, px); y0 = std::min(y0, py); x1 = std::max(x1, px); y1 = std::max(y1, py);
    }
    auto trunc_i = [](float v) { return v >= 2147483520.0f ? 2147483647 : (v <= -2147483520.0f ? -2147483647 : (int32_t)v); };
    const int32_t bx0 = trunc_i(x0), by0 = trunc_i(y0), bx1 = trunc_i(x1), by1 = trunc_i(y1);
    // pixels [bx0, bx1] x [by0, by1], clamped to the canvas: the rectangle that bins the tiles
    const float rx0 = (float)std::max(bx0, 0), ry0 = (float)std::max(by0, 0);
    const float rx1 = (float)std::min<int64_t>((int64_t)bx1 + 1, width), ry1 = (float)std::min<int64_t>((int64_t)by1 + 1, height);
    begin_path(IDENTITY, false);
    if (rx1 > rx0 && ry1 > ry0) {
        const uint8_t v[5] = {GGCUDA_VERB_MOVE, GGCUDA_VERB_LINE, GGCUDA_VERB_LINE, GGCUDA_VERB_LINE, GGCUDA_VERB_CLOSE};
        const double c[8] = {rx0, ry0, rx1, ry0, rx1, ry1, rx0, ry1};
        add_verbs(v, 5, c, 8);
    }
    end_path();
    uint32_t rec[16] = {0};
    rec[0] = 3u; rec[1] = im.w; rec[2] = im.h; rec[3] = (uint32_t)im.off;
    memcpy(rec + 5, inv, sizeof inv);
    rec[11] = (uint32_t)bx0; rec[12] = (uint32_t)by0; rec[13] = (uint32_t)bx1; rec[14] = (uint32_t)by1;
    grad_recs.insert(grad_recs.end(), rec, rec + 16);
    draw_tags.push_back(DT_GRADIENT);
    draw_data.push_back(n_gradients++);
    clip_aux.push_back(clip_stack.empty() ? -1 : clip_stack.back());
    clip_aux.push_back(0);
    return true;
}
void HostScene::begin_clip(uint32_t blend_word, float alpha, uint8_t kind) {
    int32_t d = (int32_t)draw_tags.size();
    draw_tags.push_back(DT_BEGIN_CLIP);
    draw_data.push_back(blend_word);
    uint32_t ab; memcpy(&ab, &alpha, 4);
    draw_data.push_back(ab);
    clip_aux.push_back(clip_stack.empty() ? -1 : clip_stack.back());
    clip_aux.push_back(-1);   // link patched by end_clip
    clip_stack.push_back(d); clip_kind.push_back(kind);
    // device-space bounds of everything this clip can show: its own bounds (set_next_clip_bounds, else unbounded)
    // inside its parent's
    float bb[4] = {next_clip_bb[0], next_clip_bb[1], next_clip_bb[2], next_clip_bb[3]};
    if (!clip_bb.empty()) {
        const float* p = &clip_bb[clip_bb.size() - 4];
        bb[0] = std::max(bb[0], p[0]); bb[1] = std::max(bb[1], p[1]); bb[2] = std::min(bb[2], p[2]); bb[3] = std::min(bb[3], p[3]);
    }
    clip_bb.insert(clip_bb.end(), bb, bb + 4);
    next_clip_bb[0] = next_clip_bb[1] = -3.0e38f; next_clip_bb[2] = next_clip_bb[3] = 3.0e38f;
    n_clips++;
}
// Compose modes whose result differs from the backdrop where the layer is transparent:
// Clear, Copy, SrcIn, DestIn, SrcOut, DestAtop (Porter-Duff with Fb != 1 at Sa == 0).
static bool blend_erases_backdrop(uint32_t bw) {
    uint32_t mix = (bw >> 8) & 0xffu, compose = bw & 0xffu;
    return mix == 0 && (compose == 0 || compose == 1 || compose == 5 || compose == 6 || compose == 7 || compose == 10);
}
void HostScene::begin_layer(uint32_t blend_word, float alpha) {
    // A layer is a BeginClip/EndClip pair carrying the blend mode and alpha. If the mode cannot change the backdrop
    // where the layer is empty, the layer is "implicit": no clip geometry (an empty path), coverage 1 everywhere,
    // present in a tile only where something is drawn inside it (GG_BLEND_IMPLICIT | GG_BLEND_ELIDE_EMPTY).
    // The six wiping compose modes act on the whole canvas, so they get an explicit full-canvas rectangle.
    begin_path(IDENTITY, false);
    if (blend_erases_backdrop(blend_word)) {
        // ... the canvas as far as an enclosing clip lets anything show: outside the bounds of that clip's path its
        // coverage is zero and the layer is invisible, so the rectangle stops there (rounded outwards to whole pixels)
        float x0 = 0, y0 = 0, x1 = (float)width, y1 = (float)height;
        if (!clip_bb.empty()) {
            const float* p = &clip_bb[clip_bb.size() - 4];
            x0 = std::max(x0, floorf(p[0])); y0 = std::max(y0, floorf(p[1])); x1 = std::min(x1, ceilf(p[2])); y1 = std::min(y1, ceilf(p[3]));
            if (!(x1 > x0) || !(y1 > y0)) { x1 = x0; y1 = y0; }
        }
        move_to(x0, y0); line_to(x1, y0); line_to(x1, y1); line_to(x0, y1); close();
    } else {
        blend_word |= 0xC0000000u;
        // ordinal among the implicit layers, in the (otherwise unused) width word of this clip path's style: the device keeps
        // one bit per (implicit layer, tile) so that a layer enters a tile's hit list once, not once per enclosed draw
        styles[styles.size() - 2] = n_implicit++;
    }
    end_path();
    begin_clip(blend_word, alpha, 1);
}
bool HostScene::end_clip(uint8_t kind) {
    if (clip_stack.empty() || clip_kind.back() != kind) return false;
    int32_t b = clip_stack.back();
    clip_stack.pop_back(); clip_kind.pop_back(); clip_bb.resize(clip_bb.size() - 4);
    int32_t d = (int32_t)draw_tags.size();
    // EndClip: dummy path marker so that path index == draw index (scene_encode.go:258-268);
    // we also give it a style word so that styles[path_ix] is valid for every path.
    tags.push_back(PT_STYLE); styles.push_back(0); styles.push_back(0); styles.push_back(0);
    tags.push_back(PT_PATH); n_paths++;
    draw_tags.push_back(DT_END_CLIP);
    clip_aux.push_back(b);    // "parent" of an EndClip = its BeginClip
    clip_aux.push_back(b);
    clip_aux[2 * (size_t)b + 1] = d;
    n_clips++;
    return true;
}
void HostScene::close_open_clips() {
    while (!clip_stack.empty()) end_clip(clip_kind.back());
}

size_t HostScene::packed_words() const {
    size_t tw = (tags.size() + 3) / 4;
    size_t padded = (tw + 255) / 256 * 256;
    if (padded == 0) padded = 256;
    return padded + path_data.size() + draw_tags.size() + draw_data.size() + transforms.size() + styles.size() + clip_aux.size() + 8 + gradient_words();
}
void HostScene::pack(uint32_t* out, Layout* L, uint32_t band_tiles) const {   // scene_encode.go:280-356
    uint32_t tw = (uint32_t)((tags.size() + 3) / 4);
    uint32_t padded = (tw + 255) / 256 * 256;
    if (padded == 0) padded = 256;
    uint32_t off = 0;
    L->n_tag_bytes = (uint32_t)tags.size(); L->n_tag_words = padded;
    L->n_draws = (uint32_t)draw_tags.size(); L->n_paths = n_paths; L->n_clips = n_clips;
    L->path_tag_base = off; off += padded;
    L->path_data_base = off; off += (uint32_t)path_data.size();
    L->draw_tag_base = off; off += (uint32_t)draw_tags.size();
    L->draw_data_base = off; off += (uint32_t)draw_data.size();
    L->transform_base = off; off += (uint32_t)transforms.size();
    L->style_base = off; off += (uint32_t)styles.size();
    L->clip_aux_base = off; off += (uint32_t)clip_aux.size();
    L->n_scene_words = off;
    memset(out, 0, sizeof(uint32_t) * padded);
    memcpy(out, tags.data(), tags.size());   // little endian: byte i -> bits 8*(i%4) of word i/4 (packPathTags)
    if (!path_data.empty()) memcpy(out + L->path_data_base, path_data.data(), 4 * path_data.size());
    if (!draw_tags.empty()) memcpy(out + L->draw_tag_base, draw_tags.data(), 4 * draw_tags.size());
    if (!draw_data.empty()) memcpy(out + L->draw_data_base, draw_data.data(), 4 * draw_data.size());
    if (!transforms.empty()) memcpy(out + L->transform_base, transforms.data(), 4 * transforms.size());
    if (!styles.empty()) memcpy(out + L->style_base, styles.data(), 4 * styles.size());
    if (!clip_aux.empty()) memcpy(out + L->clip_aux_base, clip_aux.data(), 4 * clip_aux.size());
    uint32_t* tail = out + off;   // device-resident element counts for the scan primitive
    tail[0] = padded; tail[1] = L->n_draws; tail[2] = L->n_tag_bytes; tail[3] = n_paths; tail[4] = band_tiles;
    // gradient table behind the tail: records | stops | image pixels; tail[5] = its word offset, tail[6] = number of records, tail[7] = its words
    tail[5] = off + 8; tail[6] = n_gradients; tail[7] = (uint32_t)gradient_words();
    uint32_t* gt = out + off + 8;
    if (!grad_recs.empty()) {
        const uint32_t stops_base = (uint32_t)grad_recs.size();
        memcpy(gt, grad_recs.data(), 4 * grad_recs.size());
        const uint32_t images_base = stops_base + (uint32_t)grad_stops.size();
        for (uint32_t g = 0; g < n_gradients; g++) gt[16 * g + 3] += gt[16 * g] == 3u ? images_base : stops_base;   // offsets relative to the table
        if (!grad_stops.empty()) memcpy(gt + stops_base, grad_stops.data(), 4 * grad_stops.size());
        if (!image_words.empty()) memcpy(gt + images_base, image_words.data(), 4 * image_words.size());
    }
}

// ------------------------------------------------------------------ scene.Encoding ingest
static void emit_round_rect(HostScene* s, float x0, float y0, float x1, float y1, float rx, float ry) {
    // scene/path.go rounded rectangle with kappa arcs (the outline that bins the tiles of a TagFillRoundRect; its coverage
    // comes from the signed distance field, draw_sdf_round_rect)
    const float k = 0.5522847498f;
    float w = x1 - x0, h = y1 - y0;
    if (rx > w * 0.5f) rx = w * 0.5f;
    if (ry > h * 0.5f) ry = h * 0.5f;
    if (rx <= 0 || ry <= 0) {
        s->move_to(x0, y0); s->line_to(x1, y0); s->line_to(x1, y1); s->line_to(x0, y1); s->close();
        return;
    }
    float kx = rx * k, ky = ry * k;
    s->move_to(x0 + rx, y0);
    s->line_to(x1 - rx, y0);
    s->cubic_to(x1 - rx + kx, y0, x1, y0 + ry - ky, x1, y0 + ry);
    s->line_to(x1, y1 - ry);
    s->cubic_to(x1, y1 - ry + ky, x1 - rx + kx, y1, x1 - rx, y1);
    s->line_to(x0 + rx, y1);
    s->cubic_to(x0 + rx - kx, y1, x0, y1 - ry + ky, x0, y1 - ry);
    s->line_to(x0, y0 + ry);
    s->cubic_to(x0, y0 + ry - ky, x0 + rx - kx, y0, x0 + rx, y0);
    s->close();
}

static uint32_t brush_color(const double* brushes, size_t n_brushes, uint32_t ix) {
    uint8_t c[4] = {0, 0, 0, 0};   // missing brush: zero Brush{} -> transparent
    if (ix < n_brushes) {
        const double* b = brushes + 4 * (size_t)ix;
        c[0] = gg_clamp_u8(b[0]); c[1] = gg_clamp_u8(b[1]); c[2] = gg_clamp_u8(b[2]); c[3] = gg_clamp_u8(b[3]);
    }
    return gg_pack_color_straight(c);
}

// ---- two-pass ingest: path geometry on a small persistent thread pool
// A scene.Encoding is consumed in three steps: (A) one light walk over the tags finds every Fill / Stroke with the slices of
// the tag and coordinate streams its path occupies and the transform in force; (B) the pool builds the packed geometry of
// those paths side by side into per-chunk arenas (a pure function per path: band culling, verb translation, auto-close,
// stroke markers); (C) the ordinary sequential walk does the order-dependent bookkeeping -- transforms, styles, draw
// objects, layers, clips -- and copies each finished slice in place. The packed scene is byte for byte what the
// one-thread walk produces. Off unless GGCUDA_INGEST_THREADS asks for it (see IngestPool()).
namespace {
class IngestPool {
public:
    static IngestPool& get() { static IngestPool p; return p; }
    unsigned workers() const { return (unsigned)th_.size() + 1; }
    // fn(chunk) for chunk in [0, n): the caller works too; returns when all chunks are done
    void run(unsigned n, const std::function<void(unsigned)>& fn) {
        if (n == 0) return;
        if (th_.empty() || n == 1) { for (unsigned c = 0; c < n; c++) fn(c); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; n_ = n; next_.store(0); done_ = 0; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return done_ == n_; });
        fn_ = nullptr;
    }
private:
    IngestPool() {
        // Opt-in: measured on the B200 box's 16 host cores the 96 k-tag benchmark encoding ingests in 0.80 ms on one thread
        // and 1.0-1.6 ms on 2-8 (the per-path work, ~1 us, is smaller than the hand-off), so the default is the one-thread
        // walk; GGCUDA_INGEST_THREADS=n turns the pool on for encodings whose paths are long.
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned want = 1;
        if (const char* e = getenv("GGCUDA_INGEST_THREADS")) want = std::min(hw, (unsigned)std::max(1, atoi(e)));
        for (unsigned i = 1; i < want; i++) th_.emplace_back([this] { loop(); });
    }
    ~IngestPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void work() {
        for (;;) {
            unsigned c = next_.fetch_add(1);
            if (c >= n_) break;
            (*fn_)(c);
            std::lock_guard<std::mutex> lk(m_);
            if (++done_ == n_) cv_done_.notify_all();
        }
    }
    void loop() {
        unsigned seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, cv_done_;
    const std::function<void(unsigned)>* fn_ = nullptr;
    unsigned n_ = 0, done_ = 0, gen_ = 0;
    std::atomic<unsigned> next_{0};
    bool stop_ = false;
};
struct IngestJob {
    const uint8_t* tg; const float* pd; uint32_t nt, npd;
    float t[6]; StrokeStyleHost st; float reach; bool stroke;
    uint32_t chunk, t_off, n_t, d_off, n_d, nseg; bool culled;   // results
};
struct IngestArena { std::vector<uint8_t> tags; std::vector<float> data; };
}  // namespace

int HostScene::add_encoding(const uint8_t* tg, size_t n_tags, const float* pd, size_t n_pd, const uint32_t* dd, size_t n_dd,
                            const float* tr, size_t n_tr, const double* brushes, size_t n_brushes, std::string* msg) {
    // Strokes are expanded on the host (the reference's stroke expander is host code too).
    // Pass 1 collects every TagStroke as a job, the jobs run on all host cores, pass 2 is the ordinary
    // sequential walk that appends each finished outline in scene order.
    struct StrokeJob { std::vector<uint8_t> verbs; std::vector<float> dev; StrokeStyleHost st; StrokeSink out; };
    std::vector<StrokeJob> jobs;
    auto t_begin = std::chrono::steady_clock::now();
    if (host_strokes) {
        size_t pi = 0, di = 0, ti = 0;
        float t[6]; memcpy(t, IDENTITY, sizeof t);
        std::vector<uint8_t> pv; std::vector<float> pc; bool active = false;
        for (size_t i = 0; i < n_tags; i++) {
            switch (tg[i]) {
            case ST_TRANSFORM: if (ti + 1 <= n_tr) memcpy(t, tr + 6 * ti, sizeof t); ti++; break;
            case ST_SET_AA: di += 1; break;
            case ST_BEGIN_PATH: pv.clear(); pc.clear(); active = true; break;
            case ST_MOVE_TO: case ST_LINE_TO: if (active && pi + 2 <= n_pd) { pv.push_back(tg[i] == ST_MOVE_TO ? GGCUDA_VERB_MOVE : GGCUDA_VERB_LINE); pc.insert(pc.end(), pd + pi, pd + pi + 2); } pi += 2; break;
            case ST_QUAD_TO: if (active && pi + 4 <= n_pd) { pv.push_back(GGCUDA_VERB_QUAD); pc.insert(pc.end(), pd + pi, pd + pi + 4); } pi += 4; break;
            case ST_CUBIC_TO: if (active && pi + 6 <= n_pd) { pv.push_back(GGCUDA_VERB_CUBIC); pc.insert(pc.end(), pd + pi, pd + pi + 6); } pi += 6; break;
            case ST_CLOSE_PATH: if (active) pv.push_back(GGCUDA_VERB_CLOSE); break;
            case ST_FILL: di += 2; active = false; break;
            case ST_STROKE:
                if (di + 5 <= n_dd && active && !pv.empty()) {
                    jobs.emplace_back();
                    StrokeJob& j = jobs.back();
                    float w, ml; memcpy(&w, dd + di + 1, 4); memcpy(&ml, dd + di + 2, 4);
                    j.st = {(double)w, (double)ml, (int)dd[di + 3], (int)dd[di + 4]};
                    j.verbs = pv;
                    // CPU scene tiles transform the points, then stroke with the raw width
                    // (scene/renderer.go:655-684, 707-713): do the same, in device space.
                    j.dev.resize(pc.size());
                    for (size_t k = 0; k + 1 < pc.size(); k += 2) {
                        j.dev[k] = t[0] * pc[k] + t[1] * pc[k + 1] + t[2];
                        j.dev[k + 1] = t[3] * pc[k] + t[4] * pc[k + 1] + t[5];
                    }
                }
                di += 5; active = false; break;
            case ST_FILL_ROUND_RECT: di += 2; pi += 6; break;
            case ST_PUSH_LAYER: di += 2; break;
            case ST_BEGIN_CLIP: active = false; break;
            case ST_BRUSH: pi += 4; break;
            case ST_IMAGE: di += 1; ti++; break;
            default: break;
            }
        }
        auto t_collect = std::chrono::steady_clock::now();
        unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        nt = (unsigned)std::min<size_t>(nt, jobs.size() / 32 + 1);
        auto work = [&](unsigned tix) { for (size_t j = tix; j < jobs.size(); j += nt) gg_stroke_to_fill(jobs[j].verbs, jobs[j].dev, jobs[j].st, &jobs[j].out); };
        if (nt <= 1) work(0);
        else {
            std::vector<std::thread> th;
            for (unsigned k = 1; k < nt; k++) th.emplace_back(work, k);
            work(0);
            for (auto& x : th) x.join();
        }
        if (getenv("GGCUDA_TRACE")) {
            auto t_done = std::chrono::steady_clock::now();
            fprintf(stderr, "[ggcuda] ingest: collect %.2f ms, %zu stroke jobs on %u threads %.2f ms\n",
                    std::chrono::duration<double, std::milli>(t_collect - t_begin).count(), jobs.size(), nt,
                    std::chrono::duration<double, std::milli>(t_done - t_collect).count());
        }
    }
    size_t next_job = 0;
    static const struct VerbMap { uint8_t m[256]; VerbMap() { memset(m, 0xff, sizeof m); m[ST_MOVE_TO] = GGCUDA_VERB_MOVE; m[ST_LINE_TO] = GGCUDA_VERB_LINE;
                                  m[ST_QUAD_TO] = GGCUDA_VERB_QUAD; m[ST_CUBIC_TO] = GGCUDA_VERB_CUBIC; m[ST_CLOSE_PATH] = GGCUDA_VERB_CLOSE; } } verb_map;
    // ---- (A) find the paths of Fill / Stroke tags, (B) build their geometry on the pool
    std::vector<IngestJob> mt_jobs;
    std::vector<IngestArena> arenas;
    size_t next_mt = 0;
    const bool mt = !host_strokes && n_tags >= 16384 && IngestPool::get().workers() > 1;
    if (mt) {
        size_t pi = 0, di = 0, ti = 0, pt0 = 0, pt1 = 0, pp0 = 0;
        float t[6]; memcpy(t, IDENTITY, sizeof t);
        bool active = false, ok = true;
        mt_jobs.reserve(n_tags / 8 + 16);
        for (size_t i = 0; i < n_tags && ok; i++) {
            switch (tg[i]) {
            case ST_TRANSFORM: if (ti + 1 > n_tr) ok = false; else { memcpy(t, tr + 6 * ti, sizeof t); ti++; } break;
            case ST_SET_AA: di += 1; break;
            case ST_BEGIN_PATH: pt0 = pt1 = i + 1; pp0 = pi; active = true; break;
            case ST_MOVE_TO: case ST_LINE_TO: if (pi + 2 > n_pd) ok = false; else { pi += 2; if (active) pt1 = i + 1; } break;
            case ST_QUAD_TO: if (pi + 4 > n_pd) ok = false; else { pi += 4; if (active) pt1 = i + 1; } break;
            case ST_CUBIC_TO: if (pi + 6 > n_pd) ok = false; else { pi += 6; if (active) pt1 = i + 1; } break;
            case ST_CLOSE_PATH: if (active) pt1 = i + 1; break;
            case ST_END_PATH: break;
            case ST_FILL: case ST_STROKE: {
                const bool stroke = tg[i] == ST_STROKE;
                if (di + (stroke ? 5 : 2) > n_dd) { ok = false; break; }
                if (active && pt1 > pt0) {
                    IngestJob j; memset(&j, 0, sizeof j);
                    j.tg = tg + pt0; j.nt = (uint32_t)(pt1 - pt0); j.pd = pd + pp0; j.npd = (uint32_t)(pi - pp0);
                    memcpy(j.t, t, sizeof t); j.stroke = stroke;
                    if (stroke) {
                        float w, ml; memcpy(&w, dd + di + 1, 4); memcpy(&ml, dd + di + 2, 4);
                        j.st = {(double)w, (double)ml, (int)dd[di + 3], (int)dd[di + 4]};
                        j.reach = 0.5f * (w > 0 ? w : 0) * ((dd[di + 4] == 0 && ml > 1.5f) ? ml : 1.5f);
                    }
                    mt_jobs.push_back(j);
                }
                di += stroke ? 5 : 2; active = false;
            } break;
            case ST_FILL_ROUND_RECT: if (di + 2 > n_dd || pi + 6 > n_pd) ok = false; else { di += 2; pi += 6; } break;
            case ST_PUSH_LAYER: if (di + 2 > n_dd) ok = false; else di += 2; break;
            case ST_POP_LAYER: case ST_END_CLIP: break;
            case ST_BEGIN_CLIP: active = false; break;
            case ST_BRUSH: pi += 4; break;
            case ST_IMAGE: if (di + 1 > n_dd) ok = false; else { di += 1; ti++; } break;
            default: ok = false; break;   // TagText / unknown: the sequential walk reports it
            }
        }
        const unsigned n_chunks = (unsigned)std::min<size_t>(4 * IngestPool::get().workers(), mt_jobs.size() / 64 + 1);
        arenas.resize(n_chunks);
        const size_t nj = mt_jobs.size();
        IngestPool::get().run(n_chunks, [&](unsigned c) {
            const size_t j0 = nj * c / n_chunks, j1 = nj * (c + 1) / n_chunks;
            size_t tb = 0, db = 0;
            for (size_t k = j0; k < j1; k++) {
                const IngestJob& j = mt_jobs[k];
                tb += j.stroke ? stroke_tag_bound(j.nt) : fill_tag_bound(j.nt);
                db += j.stroke ? stroke_data_bound(j.nt, j.npd) : fill_data_bound(j.nt, j.npd);
            }
            IngestArena& a = arenas[c];
            a.tags.resize(tb); a.data.resize(db);
            uint8_t* tp = a.tags.data(); float* dp = a.data.data();
            for (size_t k = j0; k < j1; k++) {
                IngestJob& j = mt_jobs[k];
                j.chunk = c;
                j.culled = outside_band(j.t, j.pd, j.npd, j.reach);
                if (j.culled || (j.stroke && !((float)j.st.width > 0.0f))) { j.n_t = j.n_d = j.nseg = 0; continue; }
                uint8_t* t0p = tp; float* d0p = dp;
                j.nseg = j.stroke ? build_stroke_geometry(j.tg, j.nt, j.pd, j.npd, j.st, verb_map.m, tp, dp)
                                  : build_fill_geometry(j.tg, j.nt, j.pd, j.npd, verb_map.m, tp, dp);
                j.t_off = (uint32_t)(t0p - a.tags.data()); j.n_t = (uint32_t)(tp - t0p);
                j.d_off = (uint32_t)(d0p - a.data.data()); j.n_d = (uint32_t)(dp - d0p);
            }
        });
    }
    // (C) the sequential walk; with `mt` the geometry of Fill / Stroke paths comes finished from mt_jobs, in order
    auto append_job = [&](const IngestJob& j) {
        const IngestArena& a = arenas[j.chunk];
        tags.insert(tags.end(), a.tags.begin() + j.t_off, a.tags.begin() + j.t_off + j.n_t);
        path_data.insert(path_data.end(), a.data.begin() + j.d_off, a.data.begin() + j.d_off + j.n_d);
        n_seg_tags += j.nseg;
    };
    size_t pi = 0, di = 0, ti = 0;
    float cur_t[6]; memcpy(cur_t, IDENTITY, sizeof cur_t);
    // current path: a view into the input streams (tags [pt0, pt1), coordinates from pp0), consumed by the Fill / Stroke /
    // BeginClip tag that follows it
    size_t pt0 = 0, pt1 = 0, pp0 = 0;
    bool path_active = false;
    tags.reserve(tags.size() + 2 * n_tags + 64);
    path_data.reserve(path_data.size() + 2 * n_pd + 64);
    bool pend_layer = false; uint32_t pend_blend = 0; float pend_alpha = 1.0f;
    auto flush_layer = [&]() { if (pend_layer) { pend_layer = false; begin_layer(pend_blend, pend_alpha); } };
    for (size_t i = 0; i < n_tags; i++) {
        switch (tg[i]) {
        case ST_TRANSFORM:
            if (ti + 1 > n_tr) { *msg = "encoding: transform stream underrun"; return GGCUDA_ERR_INVALID; }
            memcpy(cur_t, tr + 6 * ti, sizeof cur_t); ti++;
            break;
        case ST_SET_AA: di += 1; break;   // anti-aliasing is always on in this path
        case ST_BEGIN_PATH: {
            pt0 = pt1 = i + 1; pp0 = pi; path_active = true;
            // the path's own element tags in one tight loop (most of an encoding's tags): coordinate counts from a table
            static const struct NCoord { uint8_t n[256]; NCoord() { memset(n, 0xff, sizeof n); n[ST_MOVE_TO] = 2; n[ST_LINE_TO] = 2; n[ST_QUAD_TO] = 4;
                                                                     n[ST_CUBIC_TO] = 6; n[ST_CLOSE_PATH] = 0; } } ncoord;
            size_t j = i + 1;
            for (; j < n_tags; j++) { const uint8_t nc = ncoord.n[tg[j]]; if (nc == 0xff) break; pi += nc; }
            if (pi > n_pd) { *msg = "encoding: path stream underrun"; return GGCUDA_ERR_INVALID; }
            pt1 = j; i = j - 1;
        } break;
        case ST_MOVE_TO: case ST_LINE_TO:
            if (pi + 2 > n_pd) { *msg = "encoding: path stream underrun"; return GGCUDA_ERR_INVALID; }
            pi += 2; if (path_active) pt1 = i + 1; break;
        case ST_QUAD_TO:
            if (pi + 4 > n_pd) { *msg = "encoding: path stream underrun"; return GGCUDA_ERR_INVALID; }
            pi += 4; if (path_active) pt1 = i + 1; break;
        case ST_CUBIC_TO:
            if (pi + 6 > n_pd) { *msg = "encoding: path stream underrun"; return GGCUDA_ERR_INVALID; }
            pi += 6; if (path_active) pt1 = i + 1; break;
        case ST_CLOSE_PATH: if (path_active) pt1 = i + 1; break;
        case ST_END_PATH: break;
        case ST_FILL: {
            if (di + 2 > n_dd) { *msg = "encoding: draw stream underrun"; return GGCUDA_ERR_INVALID; }
            flush_layer();
            uint32_t bix = dd[di], style = dd[di + 1]; di += 2;
            if (path_active && pt1 > pt0) {
                if (mt && next_mt < mt_jobs.size()) {
                    const IngestJob& j = mt_jobs[next_mt++];
                    if (j.culled) n_culled++;
                    else {   // fill_verbs with the geometry already built
                        begin_path(cur_t, style == 1);
                        append_job(j);
                        tags.push_back(PT_PATH); n_paths++; in_path = false; has_move = false;
                        draw_color(brush_color(brushes, n_brushes, bix));
                    }
                } else if (outside_band(cur_t, pd + pp0, pi - pp0, 0.0f)) n_culled++;
                else { fill_verbs(cur_t, style == 1, tg + pt0, pt1 - pt0, pd + pp0, pi - pp0, verb_map.m); draw_color(brush_color(brushes, n_brushes, bix)); }
            }
            path_active = false;
        } break;
        case ST_STROKE: {
            if (di + 5 > n_dd) { *msg = "encoding: draw stream underrun"; return GGCUDA_ERR_INVALID; }
            flush_layer();
            uint32_t bix = dd[di];
            float sw = 0, sml = 0;
            memcpy(&sw, dd + di + 1, 4); memcpy(&sml, dd + di + 2, 4);
            // reach of the outline beyond the centre line: half the width, miter tips up to miter_limit times that
            const float reach = 0.5f * (sw > 0 ? sw : 0) * ((dd[di + 4] == 0 && sml > 1.5f) ? sml : 1.5f);
            if (mt && path_active && pt1 > pt0 && next_mt < mt_jobs.size()) {
                const IngestJob& j = mt_jobs[next_mt++];
                if (j.culled) n_culled++;
                else {   // stroke_path with the geometry already built
                    begin_path(cur_t, false);
                    if ((float)j.st.width > 0.0f) {
                        float w = (float)j.st.width, ml = (float)j.st.miter_limit;
                        uint32_t wb, mb; memcpy(&wb, &w, 4); memcpy(&mb, &ml, 4);
                        styles[styles.size() - 3] = stroke_style_flags(j.st); styles[styles.size() - 2] = wb; styles[styles.size() - 1] = mb;
                        append_job(j);
                        has_move = false;
                    }
                    end_path();
                    draw_color(brush_color(brushes, n_brushes, bix));
                }
            } else if (path_active && pt1 > pt0 && outside_band(cur_t, pd + pp0, pi - pp0, reach)) {
                n_culled++;
                if (host_strokes && next_job < jobs.size()) next_job++;
            } else if (path_active && pt1 > pt0) {
                if (host_strokes) {
                    if (next_job < jobs.size()) { begin_path(IDENTITY, false); append_stroke(jobs[next_job++].out); end_path(); }
                    else { begin_path(IDENTITY, false); end_path(); }
                } else {
                    float w, ml; memcpy(&w, dd + di + 1, 4); memcpy(&ml, dd + di + 2, 4);
                    StrokeStyleHost st = {(double)w, (double)ml, (int)dd[di + 3], (int)dd[di + 4]};
                    stroke_path(cur_t, tg + pt0, pt1 - pt0, pd + pp0, pi - pp0, st, verb_map.m);
                }
                draw_color(brush_color(brushes, n_brushes, bix));
            }
            di += 5;
            path_active = false;
        } break;
        case ST_FILL_ROUND_RECT: {
            if (di + 2 > n_dd || pi + 6 > n_pd) { *msg = "encoding: round-rect underrun"; return GGCUDA_ERR_INVALID; }
            flush_layer();
            uint32_t bix = dd[di], style = dd[di + 1]; di += 2;
            const float* r = pd + pi; pi += 6;
            if (cur_t[1] == 0.0f && cur_t[3] == 0.0f) {
                // axis-aligned transform: the CPU renderer's SDF (scene/renderer.go:986-1043 transforms the two corners, takes
                // min(rx, ry) UNSCALED as the radius). The outline that bins the tiles is inflated by one pixel so that every
                // pixel the 0.7 px smoothstep reaches lies in a tile with a command.
                float x0 = cur_t[0] * r[0] + cur_t[2], y0 = cur_t[4] * r[1] + cur_t[5], x1 = cur_t[0] * r[2] + cur_t[2], y1 = cur_t[4] * r[3] + cur_t[5];
                if (x0 > x1) std::swap(x0, x1);
                if (y0 > y1) std::swap(y0, y1);
                const float hw = (x1 - x0) / 2, hh = (y1 - y0) / 2;
                const float radius = std::min(std::min(r[4], r[5]), std::min(hw, hh));
                begin_path(IDENTITY, false);
                emit_round_rect(this, x0 - 1.0f, y0 - 1.0f, x1 + 1.0f, y1 + 1.0f, std::max(radius, 0.0f) + 1.0f, std::max(radius, 0.0f) + 1.0f);
                end_path();
                draw_sdf_round_rect((x0 + x1) / 2, (y0 + y1) / 2, hw, hh, radius, brush_color(brushes, n_brushes, bix));
            } else {   // rotated / skewed: the reference's SDF ignores that; here the exact area of the transformed outline
                begin_path(cur_t, style == 1);
                emit_round_rect(this, r[0], r[1], r[2], r[3], r[4], r[5]);
                end_path();
                draw_color(brush_color(brushes, n_brushes, bix));
            }
        } break;
        case ST_PUSH_LAYER: {
            if (di + 2 > n_dd) { *msg = "encoding: draw stream underrun"; return GGCUDA_ERR_INVALID; }
            flush_layer();
            uint32_t blend = dd[di]; memcpy(&pend_alpha, dd + di + 1, 4); di += 2;
            pend_blend = gg_blend_word(blend); pend_layer = true;   // scene.PushLayer may follow up with its clip shape
        } break;
        case ST_POP_LAYER:
            flush_layer();
            while (!clip_stack.empty() && clip_kind.back() == 0) end_clip(0);   // unbalanced clips inside the layer
            if (!clip_stack.empty()) end_clip(clip_kind.back());
            break;
        case ST_BEGIN_CLIP:
            if (path_active && pt1 > pt0) {
                fill_verbs(cur_t, false, tg + pt0, pt1 - pt0, pd + pp0, pi - pp0, verb_map.m);
                set_next_clip_bounds(cur_t, tg + pt0, pt1 - pt0, pd + pp0, pi - pp0, verb_map.m);
            } else { begin_path(IDENTITY, false); end_path(); next_clip_bb[0] = next_clip_bb[1] = next_clip_bb[2] = next_clip_bb[3] = 0; }   // empty clip path clips everything (renderer.go:747-753)
            if (pend_layer) {
                // PushLayer(blend, alpha, clip) (scene/scene.go:307-341: PushLayer, the clip path, BeginClip): ONE clip whose
                // path is the clip shape and whose blend word / alpha are the layer's -- the layer composites where its clip
                // lets it (also for the modes that wipe their backdrop), and fine keeps one stack level instead of two.
                // Where nothing is drawn inside it may be dropped unless the mode wipes.
                begin_clip(blend_erases_backdrop(pend_blend) ? pend_blend : (pend_blend | 0x80000000u), pend_alpha, 2);
                pend_layer = false;
            } else {
                begin_clip(0x8003u, 1.0f, 0);
            }
            path_active = false;
            break;
        case ST_END_CLIP:
            if (!clip_kind.empty() && clip_kind.back() == 2) break;   // the clip of a merged layer: closed by its PopLayer
            end_clip(0);
            break;
        case ST_BRUSH: pi += 4; break;
        case ST_IMAGE: {
            // the image's affine is the next entry of the transform stream (EncodeImage appends one without a TagTransform,
            // decoder.go:318-335); it does not become the current transform
            if (di + 1 > n_dd) { *msg = "encoding: draw stream underrun"; return GGCUDA_ERR_INVALID; }
            flush_layer();
            const uint32_t index = dd[di]; di += 1;
            const float* it = ti + 1 <= n_tr ? tr + 6 * ti : IDENTITY;
            ti++;
            if (index < images.size()) draw_image(index, it);   // an unknown index draws nothing (renderer.go:773)
            path_active = false;
        } break;
        case ST_TEXT: *msg = "encoding: TagText must be resolved to outlines before the CUDA path"; return GGCUDA_ERR_UNSUPPORTED;
        default: *msg = "encoding: unknown tag"; return GGCUDA_ERR_INVALID;
        }
    }
    flush_layer();
    close_open_clips();   // renderer.go:789-797
    if (getenv("GGCUDA_TRACE"))
        fprintf(stderr, "[ggcuda] ingest: total %.2f ms (%zu tags in, %zu packed tags out)\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), n_tags, tags.size());
    return 0;
}

// ------------------------------------------------------------------ stroke outline (host)
// Polyline stroker: curves are subdivided to `tol`, each subpath becomes one closed outline
// (left side forward, cap, right side backward, cap) or two loops for closed subpaths; joins
// are added on both sides (the inner ones overlap harmlessly under NonZero).
namespace {
struct P2 { double x, y; bool smooth; P2() : x(0), y(0), smooth(false) {} P2(double x_, double y_, bool s_ = false) : x(x_), y(y_), smooth(s_) {} };   // smooth: interior vertex of a flattened curve (no join style)
const double STROKE_TOL = 0.25;   // same tolerance as the fill flattener (flatten.go:19)

// adaptive de Casteljau subdivision: stop when the control points are within tol of the chord
void flatten_cubic_rec(std::vector<P2>& o, P2 a, P2 b, P2 c, P2 d, int depth) {
    double ux = 3 * b.x - 2 * a.x - d.x, uy = 3 * b.y - 2 * a.y - d.y;
    double vx = 3 * c.x - 2 * d.x - a.x, vy = 3 * c.y - 2 * d.y - a.y;
    double m = fmax(ux * ux, vx * vx) + fmax(uy * uy, vy * vy);
    if (m <= 16 * STROKE_TOL * STROKE_TOL || depth >= 16) { P2 e = d; e.smooth = true; o.push_back(e); return; }
    P2 ab = {(a.x + b.x) / 2, (a.y + b.y) / 2}, bc = {(b.x + c.x) / 2, (b.y + c.y) / 2}, cd = {(c.x + d.x) / 2, (c.y + d.y) / 2};
    P2 abc = {(ab.x + bc.x) / 2, (ab.y + bc.y) / 2}, bcd = {(bc.x + cd.x) / 2, (bc.y + cd.y) / 2};
    P2 mid = {(abc.x + bcd.x) / 2, (abc.y + bcd.y) / 2};
    flatten_cubic_rec(o, a, ab, abc, mid, depth + 1);
    flatten_cubic_rec(o, mid, bcd, cd, d, depth + 1);
}
void flatten_cubic(std::vector<P2>& o, P2 a, P2 b, P2 c, P2 d) {
    flatten_cubic_rec(o, a, b, c, d, 0);
    o.back().smooth = false;   // the segment's end point is a real path vertex
}
void flatten_quad(std::vector<P2>& o, P2 a, P2 b, P2 c) {
    P2 c1 = {a.x + 2.0 / 3.0 * (b.x - a.x), a.y + 2.0 / 3.0 * (b.y - a.y)}, c2 = {c.x + 2.0 / 3.0 * (b.x - c.x), c.y + 2.0 / 3.0 * (b.y - c.y)};
    flatten_cubic(o, a, c1, c2, c);
}
void arc_points(std::vector<P2>& o, P2 c, double r, double a0, double a1) {   // a0 -> a1 (signed sweep), excluding start, including end
    double sweep = a1 - a0;
    double step = 2 * acos(fmax(0.0, 1 - STROKE_TOL / fmax(r, STROKE_TOL)));
    if (step < 0.05) step = 0.05;
    int n = (int)ceil(fabs(sweep) / step);
    if (n < 1) n = 1;
    for (int i = 1; i <= n; i++) { double a = a0 + sweep * i / n; o.push_back({c.x + r * cos(a), c.y + r * sin(a)}); }
}
// join at vertex v between directions d0 -> d1 on the side with normal sign `side` (+1 = left)
void add_join(std::vector<P2>& o, P2 v, P2 d0, P2 d1, double hw, int side, int join, double miter_limit) {
    P2 n0 = {-d0.y * side, d0.x * side}, n1 = {-d1.y * side, d1.x * side};
    if (v.smooth) {   // flattened-curve interior vertex: one point on the bisector (turning angle is small)
        double mx = n0.x + n1.x, my = n0.y + n1.y, ml2 = mx * mx + my * my;
        if (ml2 > 1.0) { double k = 2.0 * hw / ml2; o.push_back({v.x + mx * k, v.y + my * k}); return; }
    }
    P2 a = {v.x + n0.x * hw, v.y + n0.y * hw}, b = {v.x + n1.x * hw, v.y + n1.y * hw};
    double cross = d0.x * d1.y - d0.y * d1.x, dot = d0.x * d1.x + d0.y * d1.y;
    bool outer = cross * side < 0;   // turning away from this side
    o.push_back(a);
    if (outer && fabs(cross) > 1e-12) {
        if (join == GGCUDA_JOIN_ROUND) {
            double a0 = atan2(n0.y, n0.x), a1 = atan2(n1.y, n1.x);
            double sw = a1 - a0;
            while (sw > M_PI) sw -= 2 * M_PI;
            while (sw < -M_PI) sw += 2 * M_PI;
            arc_points(o, v, hw, a0, a0 + sw);
            return;
        } else if (join == GGCUDA_JOIN_MITER) {
            double cos_half = sqrt(fmax(0.0, (1 + dot) * 0.5));
            if (cos_half > 1e-9 && 1.0 / cos_half <= miter_limit) {
                P2 m = {n0.x + n1.x, n0.y + n1.y};
                double ml = sqrt(m.x * m.x + m.y * m.y);
                if (ml > 1e-12) { double k = hw / cos_half / ml; o.push_back({v.x + m.x * k, v.y + m.y * k}); }
            }
        }
    }
    o.push_back(b);
}
void add_cap(std::vector<P2>& o, P2 v, P2 d, double hw, int cap) {   // from left side to right side around the end, d = outgoing direction
    P2 n = {-d.y, d.x};
    P2 l = {v.x + n.x * hw, v.y + n.y * hw}, r = {v.x - n.x * hw, v.y - n.y * hw};
    o.push_back(l);
    if (cap == GGCUDA_CAP_ROUND) {
        double a0 = atan2(n.y, n.x);
        arc_points(o, v, hw, a0, a0 - M_PI);
        return;
    } else if (cap == GGCUDA_CAP_SQUARE) {
        o.push_back({l.x + d.x * hw, l.y + d.y * hw});
        o.push_back({r.x + d.x * hw, r.y + d.y * hw});
    }
    o.push_back(r);
}
void emit_loop(StrokeSink* s, const std::vector<P2>& pts) {
    if (pts.size() < 3) return;
    size_t n = pts.size();
    s->tags.push_back(PT_MOVETO);
    s->tags.insert(s->tags.end(), n - 1, (uint8_t)PT_LINETO);
    size_t base = s->data.size();
    s->data.resize(base + 2 * n);
    float* d = s->data.data() + base;
    for (size_t i = 0; i < n; i++) {
        float x = (float)pts[i].x, y = (float)pts[i].y;
        d[2 * i] = x; d[2 * i + 1] = y;
        if (x < s->bb[0]) s->bb[0] = x;
        if (y < s->bb[1]) s->bb[1] = y;
        if (x > s->bb[2]) s->bb[2] = x;
        if (y > s->bb[3]) s->bb[3] = y;
    }
    s->n_seg += (uint32_t)(n - 1);
    if (d[0] != d[2 * n - 2] || d[1] != d[2 * n - 1]) {   // close the loop (HostScene::close)
        s->tags.push_back(PT_LINETO); s->data.push_back(d[0]); s->data.push_back(s->data[base + 1]); s->n_seg++;
    }
}
void stroke_subpath(std::vector<P2>& pts, bool closed, const StrokeStyleHost& st, StrokeSink* out) {
    // drop consecutive duplicates
    static thread_local std::vector<P2> tls_p, tls_dir, tls_o, tls_left, tls_right;
    std::vector<P2>&p = tls_p, &dir = tls_dir, &o = tls_o, &left = tls_left, &right = tls_right;   // one TLS lookup each
    if (p.capacity() < 4096) { p.reserve(4096); dir.reserve(4096); o.reserve(16384); left.reserve(8192); right.reserve(8192); }
    p.clear();
    for (const P2& q : pts) if (p.empty() || fabs(q.x - p.back().x) > 1e-9 || fabs(q.y - p.back().y) > 1e-9) p.push_back(q);
    if (closed && p.size() > 1 && fabs(p.front().x - p.back().x) < 1e-9 && fabs(p.front().y - p.back().y) < 1e-9) p.pop_back();
    double hw = st.width * 0.5;
    if (hw <= 0) return;
    size_t n = p.size();
    if (n == 0) return;
    if (n == 1) {   // degenerate: dot for round / square caps
        if (st.cap == GGCUDA_CAP_BUTT) return;
        o.clear();
        if (st.cap == GGCUDA_CAP_ROUND) arc_points(o, p[0], hw, 0, 2 * M_PI);
        else { o.push_back({p[0].x - hw, p[0].y - hw}); o.push_back({p[0].x + hw, p[0].y - hw}); o.push_back({p[0].x + hw, p[0].y + hw}); o.push_back({p[0].x - hw, p[0].y + hw}); }
        emit_loop(out, o);
        return;
    }
    size_t ns = closed ? n : n - 1;
    dir.resize(ns);
    for (size_t i = 0; i < ns; i++) {
        P2 a = p[i], b = p[(i + 1) % n];
        double dx = b.x - a.x, dy = b.y - a.y, l = sqrt(dx * dx + dy * dy);
        dir[i] = {dx / l, dy / l};
    }
    if (closed && n >= 3) {
        left.clear(); right.clear();
        for (size_t i = 0; i < n; i++) {
            P2 d0 = dir[(i + n - 1) % n], d1 = dir[i];
            add_join(left, p[i], d0, d1, hw, +1, st.join, st.miter_limit);
            add_join(right, p[i], d0, d1, hw, -1, st.join, st.miter_limit);
        }
        emit_loop(out, left);
        std::reverse(right.begin(), right.end());
        emit_loop(out, right);
        return;
    }
    o.clear(); right.clear();
    // left side forward
    o.push_back({p[0].x - dir[0].y * hw, p[0].y + dir[0].x * hw});
    for (size_t i = 1; i + 1 < n; i++) add_join(o, p[i], dir[i - 1], dir[i], hw, +1, st.join, st.miter_limit);
    add_cap(o, p[n - 1], dir[ns - 1], hw, st.cap);
    // right side: built forward, appended backward
    for (size_t i = 1; i + 1 < n; i++) add_join(right, p[i], dir[i - 1], dir[i], hw, -1, st.join, st.miter_limit);
    o.insert(o.end(), right.rbegin(), right.rend());
    P2 back = {-dir[0].x, -dir[0].y};
    add_cap(o, p[0], back, hw, st.cap);
    emit_loop(out, o);
}
}  // namespace

void gg_stroke_to_fill(const std::vector<uint8_t>& verbs, const std::vector<float>& c, const StrokeStyleHost& st, StrokeSink* out) {
    static thread_local std::vector<P2> tls_pts;
    std::vector<P2>& pts = tls_pts;
    if (pts.capacity() < 4096) pts.reserve(4096);
    pts.clear();
    bool closed = false, have = false;
    P2 cur = {0, 0}, start = {0, 0};
    size_t k = 0;
    auto flush = [&]() {
        if (have && !pts.empty()) stroke_subpath(pts, closed, st, out);
        pts.clear(); closed = false;
    };
    auto seg_start = [&]() { if (pts.empty()) pts.push_back(cur); };
    for (uint8_t v : verbs) {
        switch (v) {
        case GGCUDA_VERB_MOVE: flush(); cur = start = {c[k], c[k + 1]}; k += 2; have = true; break;
        case GGCUDA_VERB_LINE: if (have) { seg_start(); cur = {c[k], c[k + 1]}; pts.push_back(cur); } k += 2; break;
        case GGCUDA_VERB_QUAD: if (have) { seg_start(); P2 b = {c[k], c[k + 1]}, e = {c[k + 2], c[k + 3]}; flatten_quad(pts, cur, b, e); cur = e; } k += 4; break;
        case GGCUDA_VERB_CUBIC: if (have) { seg_start(); P2 b = {c[k], c[k + 1]}, d = {c[k + 2], c[k + 3]}, e = {c[k + 4], c[k + 5]}; flatten_cubic(pts, cur, b, d, e); cur = e; } k += 6; break;
        case GGCUDA_VERB_CLOSE: if (have) { closed = true; flush(); cur = start; } break;
        }
    }
    flush();
}
