/*
 * oracle/packed.c -- TEST INFRASTRUCTURE ONLY (see twin.c header).
 *
 * CPU restatement of the product's device stages when they start from ggcuda's packed scene
 * (Vello-style path tags with real curves and transforms) instead of pre-flattened PathDefs:
 *   flatten: transform (scene/encoding.go:348-350, f32, no FMA), quad elevation
 *            (internal/gpu/path_convert.go:60-72), zero-length LineTo drop (:55), FlattenFill
 *            (tilecompute/flatten.go:32-184) -- in tag order, the order the device emits;
 *   then ot_coarse_run (tilecompute/coarse.go) and fine (tilecompute/fine.go) on host threads.
 */
#include "twin.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { L_N_TAG_BYTES, L_N_TAG_WORDS, L_N_DRAWS, L_N_PATHS, L_N_CLIPS, L_PATH_TAG_BASE, L_PATH_DATA_BASE, L_DRAW_TAG_BASE,
       L_DRAW_DATA_BASE, L_TRANSFORM_BASE, L_STYLE_BASE, L_CLIP_AUX_BASE, L_N_SCENE_WORDS };

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static inline float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* control points of the segment at tag i in cubic form, device space (transform, then quad elevation) */
static int seg_cubic_form(uint8_t tag, const uint32_t *data, uint32_t off, const float *t, float *c) {
    uint32_t seg = tag & 3u;
    float p[8];
    for (uint32_t k = 0; k <= seg; k++) {
        float x = bits_f(data[off - 2 + 2 * k]), y = bits_f(data[off - 2 + 2 * k + 1]);
        p[2 * k] = t[0] * x + t[1] * y + t[2];
        p[2 * k + 1] = t[3] * x + t[4] * y + t[5];
    }
    if (seg == 1) {
        c[0] = p[0]; c[1] = p[1]; c[6] = p[2]; c[7] = p[3];
        c[2] = c[3] = c[4] = c[5] = 0.0f;
        return 1;
    }
    if (seg == 2) {
        const float k23 = (float)(2.0 / 3.0);
        c[0] = p[0]; c[1] = p[1];
        c[2] = p[0] + k23 * (p[2] - p[0]); c[3] = p[1] + k23 * (p[3] - p[1]);
        c[4] = p[4] + k23 * (p[2] - p[4]); c[5] = p[5] + k23 * (p[3] - p[5]);
        c[6] = p[4]; c[7] = p[5];
    } else {
        memcpy(c, p, sizeof(float) * 8);
    }
    return 3;
}

uint32_t ot_flatten_packed(const uint32_t *scene, const uint32_t *L, ot_line_soup *out, uint32_t cap) {
    const uint8_t *tags = (const uint8_t *)(scene + L[L_PATH_TAG_BASE]);
    const uint32_t *data = scene + L[L_PATH_DATA_BASE];
    const uint32_t *tr = scene + L[L_TRANSFORM_BASE];
    const uint32_t *styles = scene + L[L_STYLE_BASE];
    uint32_t n = 0, path_ix = 0, off = 0, trans_ix = 0, style_ix = 0;
    float t[6] = {1, 0, 0, 0, 1, 0};
    for (uint32_t i = 0; i < L[L_N_TAG_BYTES]; i++) {
        uint8_t tag = tags[i];
        if (tag == 0x20) { for (int k = 0; k < 6; k++) t[k] = bits_f(tr[6 * trans_ix + k]); trans_ix++; continue; }
        if (tag == 0x10) { path_ix++; continue; }
        if (tag == 0x40) { style_ix++; continue; }
        if ((tag & 0x7f) == 0x0C) { off += 2; continue; }   /* MoveTo / marker MoveTo */
        uint32_t seg = tag & 3u;
        if (!seg) continue;
        float c[8];
        int kind = seg_cubic_form(tag, data, off, t, c);
        off += 2 * seg;
        uint32_t before = n;
        const uint32_t *sty = style_ix ? styles + 3 * (style_ix - 1) : NULL;
        uint32_t room = n < cap ? cap - n : 0;
        if (sty && (sty[0] & 1u)) {   /* stroked path: gg_b200/csrc/stroke.cuh */
            int role = 0, next_kind = 0;
            float nx[8];
            if (tag & 0x80) role = (i > 0 && tags[i - 1] == 0x8C) ? 2 : 1;
            else if (i + 1 < L[L_N_TAG_BYTES] && (tags[i + 1] & 3u)) next_kind = seg_cubic_form(tags[i + 1], data, off, t, nx);
            n += ot_stroke_segment(c, kind, role, nx, next_kind, bits_f(sty[1]), bits_f(sty[2]), (int)((sty[0] >> 2) & 3u),
                                   (int)((sty[0] >> 4) & 3u), room ? out + n : NULL, room);
        } else if (kind == 1) {
            if (!(c[0] == c[6] && c[1] == c[7])) {
                if (n < cap) { out[n].p0[0] = c[0]; out[n].p0[1] = c[1]; out[n].p1[0] = c[6]; out[n].p1[1] = c[7]; }
                n++;
            }
        } else {
            n += ot_flatten_fill(c, 1, room ? out + n : NULL, room);
        }
        for (uint32_t k = before; k < n && k < cap; k++) out[k].path_ix = path_ix;
    }
    return n;
}

typedef struct { const ot_coarse *c; const float *bg; int w, h, t0, t1; uint8_t *out; } fine_job;
static inline float clamp01(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static void *fine_worker(void *arg) {
    fine_job *j = (fine_job *)arg;
    const ot_coarse *c = j->c;
    float px[256 * 4];
    for (int t = j->t0; t < j->t1; t++) {
        int tx = t % c->width_in_tiles, ty = t / c->width_in_tiles;
        ot_fine_tile_at(c->ptcl_words + c->ptcl_offsets[t], c->ptcl_offsets[t + 1] - c->ptcl_offsets[t], c->segments, c->n_segments, j->bg, px,
                        tx * 16, ty * 16, c->gtab);
        for (int ly = 0; ly < 16; ly++) {
            int py = ty * 16 + ly; if (py >= j->h) break;
            for (int lx = 0; lx < 16; lx++) {
                int pxx = tx * 16 + lx; if (pxx >= j->w) break;
                const float *p = px + 4 * (ly * 16 + lx);
                uint8_t *o = j->out + 4 * ((size_t)py * j->w + pxx);
                for (int k = 0; k < 4; k++) o[k] = (uint8_t)(uint32_t)(int64_t)(clamp01(p[k]) * 255.0f + 0.5f);   /* fine.wgsl:305-323 */
            }
        }
    }
    return NULL;
}
void ot_fine_frame_mt(const ot_coarse *c, const float bg[4], int w, int h, int threads, uint8_t *out) {
    int n = c->width_in_tiles * c->height_in_tiles;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; fine_job jobs[256];
    for (int i = 0; i < threads; i++) {
        jobs[i].c = c; jobs[i].bg = bg; jobs[i].w = w; jobs[i].h = h; jobs[i].out = out;
        jobs[i].t0 = (int)((long long)n * i / threads); jobs[i].t1 = (int)((long long)n * (i + 1) / threads);
        if (threads == 1) fine_worker(&jobs[i]); else pthread_create(&th[i], NULL, fine_worker, &jobs[i]);
    }
    if (threads > 1) for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
}

static ot_coarse *coarse_from_packed(const uint32_t *scene, const uint32_t *L, int w, int h, double *t_flatten, uint32_t *n_lines_out);

ot_coarse *ot_coarse_from_packed(const uint32_t *scene, const uint32_t *L, int w, int h) {
    return coarse_from_packed(scene, L, w, h, NULL, NULL);
}

int ot_render_packed(const uint32_t *scene, const uint32_t *L, int w, int h, const uint8_t bg[4], int threads,
                     uint8_t *out_premul, ot_timing *tm) {
    double t0 = now_s(), tf = 0;
    uint32_t n = 0;
    ot_coarse *c = coarse_from_packed(scene, L, w, h, &tf, &n);
    double t2 = now_s();
    float bgf[4] = {bg[0] / 255.0f, bg[1] / 255.0f, bg[2] / 255.0f, bg[3] / 255.0f};
    if (out_premul) ot_fine_frame_mt(c, bgf, w, h, threads, out_premul);
    double t3 = now_s();
    if (tm) { tm->t_flatten = tf; tm->t_coarse = t2 - t0 - tf; tm->t_fine = t3 - t2; tm->n_lines = n; tm->n_segments = c->n_segments;
              tm->n_ptcl_words = c->ptcl_offsets[c->width_in_tiles * c->height_in_tiles]; }
    ot_coarse_free(c);
    return 0;
}

static ot_coarse *coarse_from_packed(const uint32_t *scene, const uint32_t *L, int w, int h, double *t_flatten, uint32_t *n_lines_out) {
    double t0 = now_s();
    uint32_t cap = 1u << 16, n;
    ot_line_soup *lines = (ot_line_soup *)malloc(sizeof(ot_line_soup) * cap);
    while ((n = ot_flatten_packed(scene, L, lines, cap)) > cap) { cap = n; lines = (ot_line_soup *)realloc(lines, sizeof(ot_line_soup) * cap); }
    /* elements from the draw stream; lines are already grouped by path (tag order) */
    uint32_t n_draws = L[L_N_DRAWS];
    ot_element *el = (ot_element *)calloc(n_draws ? n_draws : 1, sizeof(ot_element));
    uint32_t *start = (uint32_t *)calloc(n_draws + 2, sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++) if (lines[i].path_ix + 1 <= n_draws) start[lines[i].path_ix + 1]++;
    for (uint32_t p = 0; p < n_draws; p++) start[p + 1] += start[p];
    uint32_t dd = 0;
    const uint32_t *dtags = scene + L[L_DRAW_TAG_BASE], *ddata = scene + L[L_DRAW_DATA_BASE], *styles = scene + L[L_STYLE_BASE];
    for (uint32_t d = 0; d < n_draws; d++) {
        el[d].line_start = start[d]; el[d].line_count = start[d + 1] - start[d];
        if (dtags[d] == 0x44 || dtags[d] == 0x444) {
            el[d].type = OT_ELEM_DRAW | OT_ELEM_PACKED | (dtags[d] == 0x444 ? OT_ELEM_GRADIENT : 0u);
            el[d].packed_rgba = ddata[dd]; el[d].even_odd = (styles[3 * d] & 2u) ? 1 : 0; dd += 1;
        }
        else if (dtags[d] == 0x9) { el[d].type = OT_ELEM_BEGIN_CLIP; el[d].blend = ddata[dd]; el[d].alpha = bits_f(ddata[dd + 1]); dd += 2; }
        else { el[d].type = OT_ELEM_END_CLIP; el[d].line_count = 0; }
    }
    double t1 = now_s();
    int saved = ot_style_per_path;
    ot_style_per_path = 1;
    ot_coarse *c = ot_coarse_run(el, n_draws, lines, w, h);
    ot_style_per_path = saved;
    {   /* the gradient table behind the packed scene's 8 tail words: tail[5] = its word offset, tail[6] = number of gradients */
        const uint32_t *tail = scene + L[L_N_SCENE_WORDS];
        uint32_t ng = tail[6];
        if (ng) {
            const uint32_t *gt = scene + tail[5];
            uint32_t words = 16 * ng;
            if (tail[7] > words) words = tail[7];   /* the table's size: records | stops | image pixels */
            for (uint32_t g = 0; g < ng; g++) {
                uint32_t e1 = gt[16 * g] >= 2u ? 0 : gt[16 * g + 3] + 8 * gt[16 * g + 2];   /* 8 floats per stop; SDF records have none */
                if (e1 > words) words = e1;
            }
            c->gtab = (uint32_t *)malloc(4 * (size_t)words);
            memcpy(c->gtab, gt, 4 * (size_t)words);
            c->n_gtab_words = words;
        }
    }
    if (t_flatten) *t_flatten = t1 - t0;
    if (n_lines_out) *n_lines_out = n;
    free(el); free(start); free(lines);
    return c;
}
