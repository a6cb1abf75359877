/*
 * oracle/twin.c -- TEST INFRASTRUCTURE ONLY. Not part of the product; nothing under
 * gg_b200/ may include, link or dlopen this file.
 *
 * CPU restatement of gg's Vello-style tile pipeline (the "CPU twin" the reference itself
 * uses as the oracle for its GPU shaders). Every function cites the reference file:line it
 * restates (paths relative to gogpu/gg). Arithmetic is float32 with the same operation
 * order as the Go source; transcendental calls go through float64 libm and are rounded to
 * float32 exactly where the Go code does (util.go:36-37,104-105). Must be compiled with
 * -ffp-contract=off (Go/amd64 does not fuse multiply-add).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the reference's own
 * golden PNGs (testdata/golden/vello-gpu-pipeline, copied to tests/golden/) at the
 * thresholds of tilecompute/rasterizer_test.go:39-135 and against the known-answer vectors
 * of scene_encode_test.go / coarse_test.go / fine_ptcl_test.go / fine_clip_test.go.
 */
#include "twin.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* 1 = reproduce coarse.go:425 (even-odd tiles with an even non-zero backdrop painted solid); default 0 = fixed */
int ot_evenodd_solid_quirk = 0;
/* 1 = fine quantises the running colour like gg's CPU pixmap does: truncated to 8 bits after every CmdColor */
int ot_truncate_per_draw = 0;

#define TILE_W 16
#define TILE_H 16
static const float TILE_SCALE = 1.0f / 16.0f;        /* types.go:7-11 */
static const float ONE_MINUS_ULP = 0.99999994f;      /* util.go:17 */
static const float ROBUST_EPSILON = 2e-7f;           /* util.go:20 */

/* ------------------------------------------------------------------ util.go */
typedef struct { float x, y; } vec2;
static inline vec2 v2(float x, float y) { vec2 v = {x, y}; return v; }
static inline vec2 vadd(vec2 a, vec2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline vec2 vsub(vec2 a, vec2 b) { return v2(a.x - b.x, a.y - b.y); }
static inline vec2 vmul(vec2 a, float s) { return v2(a.x * s, a.y * s); }
static inline float vdot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
static inline float vlen_sq(vec2 a) { return vdot(a, a); }
static inline int veq(vec2 a, vec2 b) { return a.x == b.x && a.y == b.y; }

/* Go's math.Hypot (math/hypot.go): p*sqrt(1+q*q) after ordering; restated so that the
 * float64 value matches Go's rather than glibc's correctly rounded hypot. */
static double go_hypot(double p, double q) {
    if (isinf(p) || isinf(q)) return INFINITY;
    if (isnan(p) || isnan(q)) return NAN;
    p = fabs(p); q = fabs(q);
    if (p < q) { double t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = q / p;
    return p * sqrt(1 + q * q);
}
static inline float vlength(vec2 v) { return (float)go_hypot((double)v.x, (double)v.y); }       /* util.go:36 */
static inline float vatan2(vec2 v) { return (float)atan2((double)v.y, (double)v.x); }            /* util.go:37 */

static inline float floor32(float x) { return (float)floor((double)x); }
static inline float ceil32(float x) { return (float)ceil((double)x); }
static inline float round32(float x) { return (float)round((double)x); }   /* half away from zero, as math.Round */
static inline float abs32(float x) { return fabsf(x); }
static inline float min32(float a, float b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; } /* util.go:62 */
static inline float max32(float a, float b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; } /* util.go:76 */
static inline float clamp32(float x, float lo, float hi) { if (x < lo) return lo; if (x > hi) return hi; return x; }
static inline float copysign32(float x, float y) { return copysignf(x, y); }
static inline float sin32(float x) { return (float)sin((double)x); }
static inline float cos32(float x) { return (float)cos((double)x); }
static inline float pow32(float base, int e) { float r = 1.0f; for (int i = 0; i < e; i++) r *= base; return r; }
static inline float signum32(float x) {               /* util.go:118-130 */
    if (x > 0) return 1; if (x < 0) return -1;
    return signbit(x) ? -1.0f : 1.0f;
}
/* Go float32->uint32 / int32 conversions on amd64 (CVTTSS2SQ then truncate / CVTTSS2SL). */
static inline uint32_t f2u(float f) { return (uint32_t)(int64_t)f; }
static inline int32_t f2i(float f) { return (int32_t)f; }

static uint32_t span(float a, float b) {              /* util.go:41-55 */
    float mx = a, mn = a;
    if (b > mx) mx = b;
    if (b < mn) mn = b;
    float r = ceil32(mx) - floor32(mn);
    if (r < 1.0f) r = 1.0f;
    return f2u(r);
}

/* growable line buffer with a hard cap (counts past the cap so callers can size) */
typedef struct { ot_line_soup *out; uint32_t n, cap; } line_sink;
static void sink_push(line_sink *s, vec2 a, vec2 b) {
    if (s->n < s->cap) {
        ot_line_soup *l = &s->out[s->n];
        l->path_ix = 0; l->p0[0] = a.x; l->p0[1] = a.y; l->p1[0] = b.x; l->p1[1] = b.y;
    }
    s->n++;
}

/* ------------------------------------------------------------------ euler.go */
static const float TANGENT_THRESH = 1e-6f;            /* euler.go:11 */
typedef struct { float th0, th1, chord_len, err; } cubic_params;
typedef struct { float th0, k0, k1, ch; } euler_params;

static cubic_params cubic_params_from_points_derivs(vec2 p0, vec2 p1, vec2 q0, vec2 q1, float dt) { /* euler.go:39-90 */
    vec2 chord = vsub(p1, p0);
    float chord_sq = vlen_sq(chord);
    float chord_len = (float)sqrt((double)chord_sq);
    cubic_params cp;
    if (chord_sq < TANGENT_THRESH * TANGENT_THRESH) {
        float chord_err = (float)sqrt((double)((float)(9.0 / 32.0) * (vlen_sq(q0) + vlen_sq(q1)))) * dt;
        cp.th0 = 0; cp.th1 = 0; cp.chord_len = TANGENT_THRESH; cp.err = chord_err;
        return cp;
    }
    float scale = dt / chord_sq;
    vec2 h0 = v2(q0.x * chord.x + q0.y * chord.y, q0.y * chord.x - q0.x * chord.y);
    float th0 = vatan2(h0);
    float d0 = vlength(h0) * scale;
    vec2 h1 = v2(q1.x * chord.x + q1.y * chord.y, q1.x * chord.y - q1.y * chord.x);
    float th1 = vatan2(h1);
    float d1 = vlength(h1) * scale;
    float cth0 = cos32(th0), cth1 = cos32(th1);
    float err;
    if (cth0 * cth1 < 0) {
        err = 2.0f;
    } else {
        const float two_thirds = (float)(2.0 / 3.0);
        float e0 = two_thirds / max32(1.0f + cth0, 1e-9f);
        float e1 = two_thirds / max32(1.0f + cth1, 1e-9f);
        float s0 = sin32(th0), s1 = sin32(th1);
        float s01 = cth0 * s1 + cth1 * s0;
        float amin = 0.15f * (2 * e0 * s0 + 2 * e1 * s1 - e0 * e1 * s01);
        float a = 0.15f * (2 * d0 * s0 + 2 * d1 * s1 - d0 * d1 * s01);
        float aerr = abs32(a - amin);
        float symm = abs32(th0 + th1);
        float asymm = abs32(th0 - th1);
        float dist = (float)go_hypot((double)(d0 - e0), (double)(d1 - e1));
        float ctr = 4.625e-6f * pow32(symm, 5) + 7.5e-3f * asymm * symm * symm;
        float halo_symm = 5e-3f * symm * dist;
        float halo_asymm = 7e-2f * asymm * dist;
        err = ctr + 1.55f * aerr + halo_symm + halo_asymm;
    }
    err *= chord_len;
    cp.th0 = th0; cp.th1 = th1; cp.chord_len = chord_len; cp.err = err;
    return cp;
}

static euler_params euler_params_from_angles(float th0, float th1) {   /* euler.go:93-119 */
    float k0 = th0 + th1;
    float dth = th1 - th0;
    float d2 = dth * dth;
    float k2 = k0 * k0;
    float a = 6.0f;
    a -= d2 * (float)(1.0 / 70.0);
    a -= (d2 * d2) * (float)(1.0 / 10780.0);
    a += (d2 * d2 * d2) * (float)2.769178184818219e-07;
    float b = -0.1f + d2 * (float)(1.0 / 4200.0) + d2 * d2 * (float)1.6959677820260655e-05;
    float c = (float)(-1.0 / 1400.0) + d2 * (float)6.84915970574303e-05 - k2 * (float)7.936475029053326e-06;
    a += (b + c * k2) * k2;
    float k1 = dth * a;
    float ch = 1.0f;
    ch -= d2 * (float)(1.0 / 40.0);
    ch += (d2 * d2) * (float)0.00034226190482569864;
    ch -= (d2 * d2 * d2) * (float)1.9349474568904524e-06;
    float b2 = (float)(-1.0 / 24.0) + d2 * (float)0.0024702380951963226 - d2 * d2 * (float)3.7297408997537985e-05;
    float c2 = (float)(1.0 / 1920.0) - d2 * (float)4.87350869747975e-05 - k2 * (float)3.1001936068463107e-06;
    ch += (b2 + c2 * k2) * k2;
    euler_params ep = {th0, k0, k1, ch};
    return ep;
}

static void integ_euler_10(float k0, float k1, float *uo, float *vo) {   /* euler.go:149-186 */
    float t1_1 = k0;
    float t1_2 = 0.5f * k1;
    float t2_2 = t1_1 * t1_1;
    float t2_3 = 2.0f * (t1_1 * t1_2);
    float t2_4 = t1_2 * t1_2;
    float t3_4 = t2_2 * t1_2 + t2_3 * t1_1;
    float t3_6 = t2_4 * t1_2;
    float t4_4 = t2_2 * t2_2;
    float t4_5 = 2.0f * (t2_2 * t2_3);
    float t4_6 = 2.0f * (t2_2 * t2_4) + t2_3 * t2_3;
    float t4_7 = 2.0f * (t2_3 * t2_4);
    float t4_8 = t2_4 * t2_4;
    float t5_6 = t4_4 * t1_2 + t4_5 * t1_1;
    float t5_8 = t4_6 * t1_2 + t4_7 * t1_1;
    float t6_6 = t4_4 * t2_2;
    float t6_7 = t4_4 * t2_3 + t4_5 * t2_2;
    float t6_8 = t4_4 * t2_4 + t4_5 * t2_3 + t4_6 * t2_2;
    float t7_8 = t6_6 * t1_2 + t6_7 * t1_1;
    float t8_8 = t6_6 * t2_2;
    float u = 1.0f;
    u -= (float)(1.0 / 24.0) * t2_2 + (float)(1.0 / 160.0) * t2_4;
    u += (float)(1.0 / 1920.0) * t4_4 + (float)(1.0 / 10752.0) * t4_6 + (float)(1.0 / 55296.0) * t4_8;
    u -= (float)(1.0 / 322560.0) * t6_6 + (float)(1.0 / 1658880.0) * t6_8;
    u += (float)(1.0 / 92897280.0) * t8_8;
    float v = (float)(1.0 / 12.0) * t1_2;
    v -= (float)(1.0 / 480.0) * t3_4 + (float)(1.0 / 2688.0) * t3_6;
    v += (float)(1.0 / 53760.0) * t5_6 + (float)(1.0 / 276480.0) * t5_8;
    v -= (float)(1.0 / 11612160.0) * t7_8;
    *uo = u; *vo = v;
}

static inline float ep_eval_th(const euler_params *ep, float t) {   /* euler.go:121-123 */
    return (ep->k0 + 0.5f * ep->k1 * (t - 1.0f)) * t - ep->th0;
}
static vec2 ep_eval(const euler_params *ep, float t) {              /* euler.go:125-131 */
    float thm = ep_eval_th(ep, t * 0.5f);
    float u, v;
    integ_euler_10((ep->k0 + ep->k1 * (0.5f * t - 0.5f)) * t, ep->k1 * t * t, &u, &v);
    float s = t / ep->ch * sin32(thm);
    float c = t / ep->ch * cos32(thm);
    return v2(u * c - v * s, -v * c - u * s);
}
static vec2 ep_eval_with_offset(const euler_params *ep, float t, float offset) { /* euler.go:133-137 */
    float th = ep_eval_th(ep, t);
    vec2 ov = v2(offset * sin32(th), offset * cos32(th));
    return vadd(ep_eval(ep, t), ov);
}
static vec2 es_eval_with_offset(vec2 p0, vec2 p1, const euler_params *ep, float t, float off) { /* euler.go:139-146 */
    vec2 chord = vsub(p1, p0);
    vec2 pt = ep_eval_with_offset(ep, t, off);
    return v2(p0.x + chord.x * pt.x - chord.y * pt.y, p0.y + chord.x * pt.y + chord.y * pt.x);
}

/* ------------------------------------------------------------------ flatten.go */
static const float DERIV_THRESH = 1e-6f, DERIV_EPS = 1e-6f, SUBDIV_LIMIT = 1.0f / 65536.0f;
float ot_flatten_tol = 0.25f;   /* flatten.go:19; a test may tighten it to measure what the tolerance costs against gg's CPU path */
#define FLATTEN_TOL ot_flatten_tol

static void eval_cubic_and_deriv(vec2 p0, vec2 p1, vec2 p2, vec2 p3, float t, vec2 *po, vec2 *qo) { /* flatten.go:46-56 */
    float m = 1.0f - t;
    float mm = m * m, mt = m * t, tt = t * t;
    *po = vadd(vmul(p0, mm * m), vmul(vadd(vadd(vmul(p1, 3 * mm), vmul(p2, 3 * mt)), vmul(p3, tt)), t));
    *qo = vadd(vadd(vmul(vsub(p1, p0), mm), vmul(vsub(p2, p1), 2 * mt)), vmul(vsub(p3, p2), tt));
}
static inline float cube_signed_sqrt(float x) { return x * (float)sqrt((double)abs32(x)); } /* flatten.go:195 */
static unsigned trailing_zeros32(uint32_t x) { if (x == 0) return 32; unsigned n = 0; while ((x & 1) == 0) { n++; x >>= 1; } return n; }

static void flatten_euler_fill(vec2 p0, vec2 p1, vec2 p2, vec2 p3, line_sink *lines) {   /* flatten.go:60-184 */
    if (veq(p0, p1) && veq(p0, p2) && veq(p0, p3)) return;
    uint32_t t0u = 0;
    float dt = 1.0f;
    vec2 last_p = p0;
    vec2 last_q = vsub(p1, p0);
    if (vlen_sq(last_q) < DERIV_THRESH * DERIV_THRESH) {
        vec2 dummy; eval_cubic_and_deriv(p0, p1, p2, p3, DERIV_EPS, &dummy, &last_q);
    }
    float last_t = 0.0f;
    vec2 lp0 = p0;
    for (;;) {
        float t0 = (float)t0u * dt;
        if (t0 == 1.0f) break;
        float t1 = t0 + dt;
        vec2 this_p0 = last_p, this_q0 = last_q, this_p1, this_q1;
        eval_cubic_and_deriv(p0, p1, p2, p3, t1, &this_p1, &this_q1);
        if (vlen_sq(this_q1) < DERIV_THRESH * DERIV_THRESH) {
            vec2 new_p1, new_q1;
            eval_cubic_and_deriv(p0, p1, p2, p3, t1 - DERIV_EPS, &new_p1, &new_q1);
            this_q1 = new_q1;
            if (t1 < 1.0f) { this_p1 = new_p1; t1 -= DERIV_EPS; }
        }
        float actual_dt = t1 - last_t;
        cubic_params cp = cubic_params_from_points_derivs(this_p0, this_p1, this_q0, this_q1, actual_dt);
        if (cp.err <= FLATTEN_TOL || dt <= SUBDIV_LIMIT) {
            euler_params ep = euler_params_from_angles(cp.th0, cp.th1);
            float k0_minus_half_k1 = ep.k0 - 0.5f * ep.k1;
            float k1 = ep.k1;
            float scale_mul = 0.5f * (float)(M_SQRT2 / 2.0) *
                              (float)sqrt((double)(cp.chord_len / (ep.ch * FLATTEN_TOL)));
            const float k1_thresh = 1e-3f;
            float n_frac;
            int robust; /* 1 = LowK1, 2 = LowDist */
            float a = 0, b = 0, integral = 0, int0 = 0;
            if (abs32(k1) < k1_thresh) {
                float k = k0_minus_half_k1 + 0.5f * k1;
                n_frac = (float)sqrt((double)abs32(k));
                robust = 1;
            } else {
                a = k1;
                b = k0_minus_half_k1;
                int0 = cube_signed_sqrt(b);
                float int1 = cube_signed_sqrt(a + b);
                integral = int1 - int0;
                n_frac = (float)(2.0 / 3.0) * integral / a;
                robust = 2;
            }
            float n = (float)ceil((double)(n_frac * scale_mul));
            if (n < 1) n = 1;
            if (n > 100) n = 100;
            /* NaN n (degenerate input): Go's int(NaN) on amd64 is INT64_MIN -> empty loop. */
            int n_int = (n != n) ? 0 : (int)n;
            for (int i = 0; i < n_int; i++) {
                vec2 lp1;
                if (i == n_int - 1 && t1 == 1.0f) {
                    lp1 = p3;
                } else {
                    float t = (float)(i + 1) / n;
                    float s;
                    if (robust == 1) {
                        s = t;
                    } else {
                        float c = (float)cbrt((double)(integral * t + int0));
                        float inv = c * abs32(c);
                        s = (inv - b) / a;
                    }
                    lp1 = es_eval_with_offset(this_p0, this_p1, &ep, s, 0.0f);
                }
                sink_push(lines, lp0, lp1);
                lp0 = lp1;
            }
            last_p = this_p1; last_q = this_q1; last_t = t1;
            t0u++;
            unsigned shift = trailing_zeros32(t0u);
            t0u >>= shift;
            dt *= (float)((uint32_t)1 << shift);
        } else {
            if (t0u < 0xFFFFFFFFu / 2) t0u *= 2;
            dt *= 0.5f;
        }
    }
}

/* ------------------------------------------------------------------ stroke expansion (ggcuda's own definition)
 * PARITY UNPINNED against the reference: gg expands strokes on the host with internal/stroke/expander.go before its
 * Vello path ever sees them (tilecompute/flatten.go:7-8: "stroke expansion ... is not yet ported"). ggcuda expands
 * them inside flatten on the device (gg_b200/csrc/stroke.cuh); this is the CPU statement of that definition, written
 * for a sequential walk, and what tests/ pin is (a) device lines == these lines bit for bit, (b) areas against closed
 * forms, (c) pixels against the host polyline stroker. Join/cap/miter-limit semantics follow gg (paint.go,
 * expander.go:442-476); the outline is filled NonZero as software.go:1145-1226 does.
 *
 * Construction: per flattened piece of the centre line a quad between its two offset vertices (offsets on the Euler
 * spiral's own normals, euler.go:133-146); a quad with a backward edge becomes rectangle + outer bevels + inner
 * detours through the centre; joins decorate the outer side and pass through the centre on the inner side. Every
 * piece is positively oriented, so winding numbers never cancel. */
typedef struct { float hw, miter_limit; int join, cap; } stroke_style;
typedef struct { vec2 p, n, l, r; } stroke_vtx;

static inline float vcross(vec2 a, vec2 b) { return a.x * b.y - a.y * b.x; }
static void sk_put(line_sink *s, vec2 a, vec2 b) { if (!veq(a, b)) sink_push(s, a, b); }
static vec2 sk_normal(vec2 t, float hw) {
    float len = (float)sqrt((double)(t.x * t.x + t.y * t.y));
    float k = len > 0.0f ? hw / len : 0.0f;
    return v2(-t.y * k, t.x * k);
}
static stroke_vtx sk_vtx(vec2 p, vec2 n) { stroke_vtx v = {p, n, vadd(p, n), vsub(p, n)}; return v; }
static vec2 sk_start_tangent(const vec2 *c, int kind) {
    if (kind == 1) return vsub(c[3], c[0]);
    if (!veq(c[1], c[0])) return vsub(c[1], c[0]);
    if (!veq(c[2], c[0])) return vsub(c[2], c[0]);
    return vsub(c[3], c[0]);
}
static vec2 sk_end_tangent(const vec2 *c, int kind) {
    if (kind == 1) return vsub(c[3], c[0]);
    if (!veq(c[3], c[2])) return vsub(c[3], c[2]);
    if (!veq(c[3], c[1])) return vsub(c[3], c[1]);
    return vsub(c[3], c[0]);
}
static vec2 sk_arc(line_sink *s, vec2 c, vec2 v0, float sweep, float hw, vec2 from) {
    float cc = 1.0f - 0.25f / hw;
    if (cc < -1.0f) cc = -1.0f;
    float step = 2.0f * (float)acos((double)cc);
    float nf = (float)ceil((double)(abs32(sweep) / step));
    if (!(nf >= 1.0f)) nf = 1.0f;
    if (nf > 1024.0f) nf = 1024.0f;
    int n = (int)nf;
    vec2 last = from;
    for (int k = 1; k < n; k++) {
        float a = sweep * (float)k / nf;
        float cs = cos32(a), sn = sin32(a);
        vec2 p = v2(c.x + (v0.x * cs - v0.y * sn), c.y + (v0.x * sn + v0.y * cs));
        sk_put(s, last, p);
        last = p;
    }
    return last;
}
static void sk_piece(line_sink *s, const stroke_vtx *a, const stroke_vtx *b, float hw) {
    vec2 e = vsub(b->p, a->p);
    float dl = vdot(vsub(b->l, a->l), e), dr = vdot(vsub(b->r, a->r), e);
    if ((dl > 0.0f && dr > 0.0f) || (e.x == 0.0f && e.y == 0.0f)) { sk_put(s, a->l, b->l); sk_put(s, b->r, a->r); return; }
    vec2 n = sk_normal(e, hw);
    vec2 a0 = vadd(a->p, n), b0 = vsub(a->p, n), a1 = vadd(b->p, n), b1 = vsub(b->p, n);
    if (vcross(a->n, n) > 0.0f) { sk_put(s, a->l, a->p); sk_put(s, a->p, a0); sk_put(s, b0, a->r); }
    else { sk_put(s, a->l, a0); sk_put(s, b0, a->p); sk_put(s, a->p, a->r); }
    sk_put(s, a0, a1);
    sk_put(s, b1, b0);
    if (vcross(n, b->n) > 0.0f) { sk_put(s, a1, b->p); sk_put(s, b->p, b->l); sk_put(s, b->r, b1); }
    else { sk_put(s, a1, b->l); sk_put(s, b->r, b->p); sk_put(s, b->p, b1); }
}
static void sk_join(line_sink *s, vec2 p, vec2 n0, vec2 n1, const stroke_style *st) {
    if (veq(n0, n1)) return;
    float cr = vcross(n0, n1), dt = vdot(n0, n1);
    vec2 u0, u1;
    if (cr > 0.0f) {   /* turning towards the left side: left is inner, right is outer */
        sk_put(s, vadd(p, n0), p); sk_put(s, p, vadd(p, n1));
        u0 = v2(-n1.x, -n1.y); u1 = v2(-n0.x, -n0.y);
    } else {
        sk_put(s, vsub(p, n1), p); sk_put(s, p, vsub(p, n0));
        u0 = n0; u1 = n1;
    }
    vec2 from = vadd(p, u0), to = vadd(p, u1);
    float hw2 = st->hw * st->hw;
    if (st->join == 1) {
        float sweep = (float)atan2((double)vcross(u0, u1), (double)vdot(u0, u1));
        vec2 last = sk_arc(s, p, u0, sweep, st->hw, from);
        sk_put(s, last, to);
    } else if (st->join == 0 && 2.0f * hw2 < st->miter_limit * st->miter_limit * (hw2 + dt) && hw2 + dt > 0.0f) {   /* expander.go:455-456 */
        float k = hw2 / (hw2 + dt);
        vec2 m = v2(p.x + (u0.x + u1.x) * k, p.y + (u0.y + u1.y) * k);
        sk_put(s, from, m); sk_put(s, m, to);
    } else {
        sk_put(s, from, to);
    }
}
static void sk_cap(line_sink *s, vec2 p, vec2 nf, const stroke_style *st) {
    vec2 a = vadd(p, nf), b = vsub(p, nf);
    if (st->cap == 1) {
        vec2 last = sk_arc(s, p, nf, -3.14159265358979323846f, st->hw, a);
        sk_put(s, last, b);
    } else if (st->cap == 2) {
        vec2 d = v2(nf.y, -nf.x);
        vec2 a2 = vadd(a, d), b2 = vadd(b, d);
        sk_put(s, a, a2); sk_put(s, a2, b2); sk_put(s, b2, b);
    } else {
        sk_put(s, a, b);
    }
}
/* the subdivision loop of flatten_euler_fill with one offset vertex per subdivision point and the line count of each
 * Euler segment raised by sqrt(1 + hw * max curvature) */
static void sk_cubic(line_sink *lines, vec2 p0, vec2 p1, vec2 p2, vec2 p3, vec2 n_start, vec2 n_end, float hw) {
    uint32_t t0u = 0;
    float dt = 1.0f;
    vec2 last_p = p0;
    vec2 last_q = vsub(p1, p0);
    if (vlen_sq(last_q) < DERIV_THRESH * DERIV_THRESH) { vec2 dummy; eval_cubic_and_deriv(p0, p1, p2, p3, DERIV_EPS, &dummy, &last_q); }
    float last_t = 0.0f;
    stroke_vtx va = sk_vtx(p0, n_start);
    for (;;) {
        float t0 = (float)t0u * dt;
        if (t0 == 1.0f) break;
        float t1 = t0 + dt;
        vec2 this_p0 = last_p, this_q0 = last_q, this_p1, this_q1;
        eval_cubic_and_deriv(p0, p1, p2, p3, t1, &this_p1, &this_q1);
        if (vlen_sq(this_q1) < DERIV_THRESH * DERIV_THRESH) {
            vec2 new_p1, new_q1;
            eval_cubic_and_deriv(p0, p1, p2, p3, t1 - DERIV_EPS, &new_p1, &new_q1);
            this_q1 = new_q1;
            if (t1 < 1.0f) { this_p1 = new_p1; t1 -= DERIV_EPS; }
        }
        float actual_dt = t1 - last_t;
        cubic_params cp = cubic_params_from_points_derivs(this_p0, this_p1, this_q0, this_q1, actual_dt);
        if (cp.err <= FLATTEN_TOL || dt <= SUBDIV_LIMIT) {
            euler_params ep = euler_params_from_angles(cp.th0, cp.th1);
            float kb = ep.k0 - 0.5f * ep.k1;
            float k1 = ep.k1;
            float scale_mul = 0.5f * (float)(M_SQRT2 / 2.0) * (float)sqrt((double)(cp.chord_len / (ep.ch * FLATTEN_TOL)));
            float k_abs = max32(abs32(kb), abs32(kb + k1));
            float widen = (float)sqrt((double)(1.0f + hw * k_abs * ep.ch / cp.chord_len));
            float n_frac, a = 0, b = 0, integral = 0, int0 = 0;
            int low_k1;
            if (abs32(k1) < 1e-3f) {
                float k = kb + 0.5f * k1;
                n_frac = (float)sqrt((double)abs32(k));
                low_k1 = 1;
            } else {
                a = k1; b = kb;
                int0 = cube_signed_sqrt(b);
                float int1 = cube_signed_sqrt(a + b);
                integral = int1 - int0;
                n_frac = (float)(2.0 / 3.0) * integral / a;
                low_k1 = 0;
            }
            float n = (float)ceil((double)(n_frac * scale_mul * widen));
            if (n < 1) n = 1;
            if (n > 100) n = 100;
            int n_int = (n != n) ? 0 : (int)n;
            vec2 chord = vsub(this_p1, this_p0);
            float nscale = hw / cp.chord_len;
            int tiny = vlen_sq(chord) < 1e-12f;
            for (int i = 0; i < n_int; i++) {
                stroke_vtx vb;
                if (i == n_int - 1 && t1 == 1.0f) {
                    vb = sk_vtx(p3, n_end);
                } else {
                    float t = (float)(i + 1) / n;
                    float sp;
                    if (low_k1) sp = t;
                    else {
                        float c = (float)cbrt((double)(integral * t + int0));
                        float inv = c * abs32(c);
                        sp = (inv - b) / a;
                    }
                    vec2 pc = es_eval_with_offset(this_p0, this_p1, &ep, sp, 0.0f);
                    vec2 nn = n_start;   /* no usable chord direction (euler.go:44): the curve's start offset */
                    if (!tiny) {
                        float th = ep_eval_th(&ep, sp);
                        float sx = sin32(th), sy = cos32(th);
                        nn = v2((chord.x * sx - chord.y * sy) * nscale, (chord.x * sy + chord.y * sx) * nscale);
                    }
                    vb = sk_vtx(pc, nn);
                }
                sk_piece(lines, &va, &vb, hw);
                va = vb;
            }
            last_p = this_p1; last_q = this_q1; last_t = t1;
            t0u++;
            unsigned shift = trailing_zeros32(t0u);
            t0u >>= shift;
            dt *= (float)((uint32_t)1 << shift);
        } else {
            if (t0u < 0xFFFFFFFFu / 2) t0u *= 2;
            dt *= 0.5f;
        }
    }
}

/* One segment's share of the outline. seg / next: 4 points in cubic form (kind 1: line from [0] to [3]); role 0 = a
 * real segment (next_kind 0: the subpath ends here with a cap), 1 = marker copy after a closed subpath (draws nothing),
 * 2 = marker copy after the marker MoveTo of an open subpath (draws the start cap). */
uint32_t ot_stroke_segment(const float *seg, int kind, int role, const float *next, int next_kind,
                           float width, float miter_limit, int join, int cap, ot_line_soup *out, uint32_t out_cap) {
    line_sink s = {out, 0, out_cap};
    stroke_style st = {0.5f * width, miter_limit, join, cap};
    vec2 c[4] = {v2(seg[0], seg[1]), v2(seg[2], seg[3]), v2(seg[4], seg[5]), v2(seg[6], seg[7])};
    vec2 ns = sk_normal(sk_start_tangent(c, kind), st.hw);
    if (role != 0) {
        if (role == 2) sk_cap(&s, c[0], v2(-ns.x, -ns.y), &st);
        return s.n;
    }
    vec2 ne = sk_normal(sk_end_tangent(c, kind), st.hw);
    if (kind == 1) { stroke_vtx a = sk_vtx(c[0], ns), b = sk_vtx(c[3], ne); sk_piece(&s, &a, &b, st.hw); }
    else sk_cubic(&s, c[0], c[1], c[2], c[3], ns, ne, st.hw);
    if (next_kind != 0) {
        vec2 d[4] = {v2(next[0], next[1]), v2(next[2], next[3]), v2(next[4], next[5]), v2(next[6], next[7])};
        sk_join(&s, c[3], ne, sk_normal(sk_start_tangent(d, next_kind), st.hw), &st);
    } else {
        sk_cap(&s, c[3], ne, &st);
    }
    return s.n;
}

uint32_t ot_flatten_fill(const float *cubics, uint32_t n, ot_line_soup *out, uint32_t cap) {   /* flatten.go:32-43 */
    line_sink s = {out, 0, cap};
    for (uint32_t i = 0; i < n; i++) {
        const float *c = cubics + 8 * (size_t)i;
        flatten_euler_fill(v2(c[0], c[1]), v2(c[2], c[3]), v2(c[4], c[5]), v2(c[6], c[7]), &s);
    }
    return s.n;
}

/* internal/gpu/path_convert.go:29-112 (convertPathToPathDef geometry part).
 * Lines from LineTo are appended immediately; cubics of a subpath are flushed (flattened)
 * at MoveTo / Close / end, exactly as the reference orders them. */
uint32_t ot_flatten_path(const uint8_t *verbs, uint32_t n_verbs, const double *coords,
                         int auto_close, ot_line_soup *out, uint32_t cap) {
    line_sink s = {out, 0, cap};
    float *cub = NULL; uint32_t ncub = 0, capcub = 0;
    vec2 current = v2(0, 0), start = v2(0, 0);
    int has_move = 0;
    const double *c = coords;
#define FLUSH_CUBICS() do { for (uint32_t k_ = 0; k_ < ncub; k_++) { const float *q_ = cub + 8 * (size_t)k_; \
        flatten_euler_fill(v2(q_[0], q_[1]), v2(q_[2], q_[3]), v2(q_[4], q_[5]), v2(q_[6], q_[7]), &s); } ncub = 0; } while (0)
#define PUSH_CUBIC(a0, a1, a2, a3) do { if (ncub == capcub) { capcub = capcub ? capcub * 2 : 16; cub = (float *)realloc(cub, sizeof(float) * 8 * capcub); } \
        float *q_ = cub + 8 * (size_t)ncub++; q_[0] = (a0).x; q_[1] = (a0).y; q_[2] = (a1).x; q_[3] = (a1).y; q_[4] = (a2).x; q_[5] = (a2).y; q_[6] = (a3).x; q_[7] = (a3).y; } while (0)
    for (uint32_t i = 0; i < n_verbs; i++) {
        switch (verbs[i]) {
        case 0: /* MoveTo */
            FLUSH_CUBICS();
            if (auto_close && has_move && !veq(current, start)) sink_push(&s, current, start);
            current = v2((float)c[0], (float)c[1]); start = current; has_move = 1; c += 2;
            break;
        case 1: { /* LineTo */
            vec2 pt = v2((float)c[0], (float)c[1]); c += 2;
            if (!has_move) break;
            if (!veq(pt, current)) sink_push(&s, current, pt);
            current = pt;
        } break;
        case 2: { /* QuadTo: elevated to cubic, path_convert.go:60-72 */
            vec2 ctrl = v2((float)c[0], (float)c[1]), end = v2((float)c[2], (float)c[3]); c += 4;
            if (!has_move) break;
            const float k = (float)(2.0 / 3.0);
            vec2 c1 = v2(current.x + k * (ctrl.x - current.x), current.y + k * (ctrl.y - current.y));
            vec2 c2 = v2(end.x + k * (ctrl.x - end.x), end.y + k * (ctrl.y - end.y));
            PUSH_CUBIC(current, c1, c2, end);
            current = end;
        } break;
        case 3: { /* CubicTo */
            vec2 c1 = v2((float)c[0], (float)c[1]), c2 = v2((float)c[2], (float)c[3]), end = v2((float)c[4], (float)c[5]); c += 6;
            if (!has_move) break;
            PUSH_CUBIC(current, c1, c2, end);
            current = end;
        } break;
        case 4: /* Close */
            FLUSH_CUBICS();
            if (has_move && !veq(current, start)) sink_push(&s, current, start);
            current = start;
            break;
        default: break;
        }
    }
    FLUSH_CUBICS();
    if (auto_close && has_move && !veq(current, start)) sink_push(&s, current, start);
    free(cub);
#undef FLUSH_CUBICS
#undef PUSH_CUBIC
    return s.n;
}

/* ------------------------------------------------------------------ pathtag.go / draw_leaf.go */
void ot_path_monoid_new(uint32_t tag_word, ot_path_monoid *m) {   /* pathtag.go:26-63 */
    uint32_t point_count = tag_word & 0x03030303u;
    m->path_seg_ix = (uint32_t)__builtin_popcount((point_count * 7) & 0x04040404u);
    m->trans_ix = (uint32_t)__builtin_popcount(tag_word & 0x20202020u);
    uint32_t n_points = point_count + ((tag_word >> 2) & 0x01010101u);
    uint32_t a = n_points + (n_points & (((tag_word >> 3) & 0x01010101u) * 15));
    a += a >> 8;
    a += a >> 16;
    m->path_seg_offset = a & 0xff;
    m->path_ix = (uint32_t)__builtin_popcount(tag_word & 0x10101010u);
    m->style_ix = (uint32_t)__builtin_popcount(tag_word & 0x40404040u);
}
static void pm_combine(ot_path_monoid *a, const ot_path_monoid *b) {   /* pathtag.go:66-74 */
    a->trans_ix += b->trans_ix; a->path_seg_ix += b->path_seg_ix; a->path_seg_offset += b->path_seg_offset;
    a->style_ix += b->style_ix; a->path_ix += b->path_ix;
}
void ot_draw_monoid_new(uint32_t tag, ot_draw_monoid *m) {   /* draw_leaf.go:29-41 */
    m->path_ix = tag != 0 ? 1 : 0;
    m->clip_ix = tag & 1;
    m->scene_offset = (tag >> 2) & 0x7;
    m->info_offset = (tag >> 6) & 0xf;
}
static void dm_combine(ot_draw_monoid *a, const ot_draw_monoid *b) {
    a->path_ix += b->path_ix; a->clip_ix += b->clip_ix; a->scene_offset += b->scene_offset; a->info_offset += b->info_offset;
}

enum { DRAWTAG_COLOR = 0x44, DRAWTAG_BEGIN_CLIP = 0x9, DRAWTAG_END_CLIP = 0x21,             /* scene_encode.go:70-76 */
       DRAWTAG_GRADIENT = 0x444 };   /* ggcuda: DrawTagColor's monoid increments, the scene word is a gradient index */
/* The three scan stages on a packed scene (the reference's own layout or ggcuda's, which shares the draw-tag encoding):
 *   pathtagReduce + pathtagScan (pathtag.go:76-121): exclusive PathMonoid per tag word;
 *   drawReduce + drawLeafScan (draw_leaf.go:54-151): exclusive DrawMonoid per draw object, info[], ClipInp[];
 *   clipLeafScan (clip_leaf.go:27-56): every EndClip takes path_ix and scene_offset of its BeginClip.
 * Any output may be NULL. clip_inps: 2 words per clip {ix, path_ix}. Returns the number of info words. */
uint32_t ot_scan_stages(const uint32_t *scene, uint32_t n_scene_words, uint32_t path_tag_base, uint32_t n_tag_words,
                        uint32_t draw_tag_base, uint32_t draw_data_base, uint32_t n_draw, uint32_t n_clips,
                        ot_path_monoid *tag_monoids, ot_draw_monoid *dm, uint32_t *info, int32_t *clip_inps_out) {
    if (tag_monoids) {
        ot_path_monoid m; memset(&m, 0, sizeof m);
        for (uint32_t i = 0; i < n_tag_words; i++) {
            tag_monoids[i] = m;
            ot_path_monoid t; ot_path_monoid_new(scene[path_tag_base + i], &t);
            pm_combine(&m, &t);
        }
    }
    ot_draw_monoid *own = NULL;
    if (!dm) dm = own = (ot_draw_monoid *)calloc(n_draw ? n_draw : 1, sizeof(ot_draw_monoid));
    ot_draw_monoid pre; memset(&pre, 0, sizeof pre);
    for (uint32_t i = 0; i < n_draw; i++) {
        dm[i] = pre;
        ot_draw_monoid t; ot_draw_monoid_new(scene[draw_tag_base + i], &t);
        dm_combine(&pre, &t);
    }
    uint32_t n_info = pre.info_offset;
    typedef struct { uint32_t ix; int32_t path_ix; } clip_inp;
    clip_inp *clip_inps = (clip_inp *)calloc(n_clips ? n_clips : 1, sizeof(clip_inp));
    for (uint32_t i = 0; i < n_draw; i++) {
        uint32_t tag = scene[draw_tag_base + i];
        if (tag == DRAWTAG_COLOR || tag == DRAWTAG_GRADIENT) {
            uint32_t so = draw_data_base + dm[i].scene_offset;
            if (info && so < n_scene_words && dm[i].info_offset < n_info) info[dm[i].info_offset] = scene[so];
        } else if (tag == DRAWTAG_BEGIN_CLIP) {
            if (dm[i].clip_ix < n_clips) { clip_inps[dm[i].clip_ix].ix = i; clip_inps[dm[i].clip_ix].path_ix = (int32_t)dm[i].path_ix; }
        } else if (tag == DRAWTAG_END_CLIP) {
            if (dm[i].clip_ix < n_clips) { clip_inps[dm[i].clip_ix].ix = i; clip_inps[dm[i].clip_ix].path_ix = ~(int32_t)i; }
        }
    }
    if (clip_inps_out) memcpy(clip_inps_out, clip_inps, sizeof(clip_inp) * n_clips);
    if (n_clips > 0) {   /* clipLeafScan */
        int *stack = (int *)malloc(sizeof(int) * n_clips); int sp = 0;
        for (uint32_t i = 0; i < n_clips; i++) {
            if (clip_inps[i].path_ix >= 0) { stack[sp++] = (int)i; }
            else {
                if (sp == 0) continue;
                int parent = stack[--sp];
                uint32_t end_idx = (uint32_t)(~clip_inps[i].path_ix);
                if (end_idx < n_draw && clip_inps[parent].ix < n_draw) {
                    dm[end_idx].path_ix = (uint32_t)clip_inps[parent].path_ix;
                    dm[end_idx].scene_offset = dm[clip_inps[parent].ix].scene_offset;
                }
            }
        }
        free(stack);
    }
    free(clip_inps);
    free(own);
    return n_info;
}

/* ------------------------------------------------------------------ path_count.go */
typedef struct {
    vec2 xy0, xy1, s0, s1;
    int is_down, is_positive_slope;
    uint32_t count_x, count;
    float dx, dy, a, b, sign, x0, y0;
} dda;

/* Shared DDA set-up: path_count.go:19-69 == path_tiling.go:28-71 */
static void dda_setup(const ot_line_soup *line, dda *d) {
    vec2 p0 = v2(line->p0[0], line->p0[1]), p1 = v2(line->p1[0], line->p1[1]);
    d->is_down = p1.y >= p0.y;
    if (d->is_down) { d->xy0 = p0; d->xy1 = p1; } else { d->xy0 = p1; d->xy1 = p0; }
    d->s0 = vmul(d->xy0, TILE_SCALE);
    d->s1 = vmul(d->xy1, TILE_SCALE);
    d->count_x = span(d->s0.x, d->s1.x) - 1;
    d->count = d->count_x + span(d->s0.y, d->s1.y);
    d->dx = abs32(d->s1.x - d->s0.x);
    d->dy = d->s1.y - d->s0.y;
}
static void dda_finish(dda *d) {
    float idxdy = 1.0f / (d->dx + d->dy);
    float a = d->dx * idxdy;
    d->is_positive_slope = d->s1.x >= d->s0.x;
    d->sign = d->is_positive_slope ? 1.0f : -1.0f;
    float xt0 = floor32(d->s0.x * d->sign);
    float c = d->s0.x * d->sign - xt0;
    d->y0 = floor32(d->s0.y);
    float ytop = (d->s0.y == d->s1.y) ? ceil32(d->s0.y) : d->y0 + 1.0f;
    d->b = min32((d->dy * c + d->dx * (ytop - d->s0.y)) * idxdy, ONE_MINUS_ULP);
    float robust_err = floor32(a * (float)(d->count - 1) + d->b) - (float)d->count_x;
    if (robust_err != 0.0f) a -= copysign32(ROBUST_EPSILON, robust_err);
    d->a = a;
    d->x0 = d->is_positive_slope ? xt0 * d->sign : xt0 * d->sign - 1.0f;
}

uint32_t ot_path_count(const ot_line_soup *lines, uint32_t n_lines, const ot_path *paths,
                       ot_tile *tile, ot_segment_count *seg_counts) {   /* path_count.go:11-205 */
    uint32_t bump_seg_counts = 0;
    for (uint32_t line_ix = 0; line_ix < n_lines; line_ix++) {
        const ot_line_soup *line = &lines[line_ix];
        dda d; dda_setup(line, &d);
        if (d.dx + d.dy == 0.0f) continue;
        if (d.dy == 0.0f && floor32(d.s0.y) == d.s0.y) continue;
        dda_finish(&d);
        float a = d.a, b = d.b, sign = d.sign, x0 = d.x0, y0 = d.y0;
        vec2 s0 = d.s0, s1 = d.s1;
        uint32_t count = d.count;
        const ot_path *path = &paths[line->path_ix];
        int32_t bboxi[4] = {(int32_t)path->bbox[0], (int32_t)path->bbox[1], (int32_t)path->bbox[2], (int32_t)path->bbox[3]};
        float xmin = min32(s0.x, s1.x);
        int32_t stride = bboxi[2] - bboxi[0];
        if (s0.y >= (float)bboxi[3] || s1.y < (float)bboxi[1] || xmin >= (float)bboxi[2] || stride == 0) continue;
        uint32_t imin = 0;
        if (s0.y < (float)bboxi[1]) {
            float iminf = round32(((float)bboxi[1] - y0 + b - a) / (1.0f - a)) - 1.0f;
            if (y0 + iminf - floor32(a * iminf + b) < (float)bboxi[1]) iminf += 1.0f;
            imin = f2u(iminf);
        }
        uint32_t imax = count;
        if (s1.y > (float)bboxi[3]) {
            float imaxf = round32(((float)bboxi[3] - y0 + b - a) / (1.0f - a)) - 1.0f;
            if (y0 + imaxf - floor32(a * imaxf + b) < (float)bboxi[3]) imaxf += 1.0f;
            imax = f2u(imaxf);
        }
        int32_t delta = d.is_down ? -1 : 1;
        int32_t ymin = 0, ymax = 0;
        if (max32(s0.x, s1.x) < (float)bboxi[0]) {
            ymin = f2i(ceil32(s0.y));
            ymax = f2i(ceil32(s1.y));
            imax = imin;
        } else {
            float fudge = d.is_positive_slope ? 0.0f : 1.0f;
            if (xmin < (float)bboxi[0]) {
                float f = round32((sign * ((float)bboxi[0] - x0) - b + fudge) / a);
                if ((x0 + sign * floor32(a * f + b) < (float)bboxi[0]) == d.is_positive_slope) f += 1.0f;
                int32_t ynext = f2i(y0 + f - floor32(a * f + b) + 1.0f);
                if (d.is_positive_slope) {
                    if (f2u(f) > imin) {
                        float y_off = (y0 != s0.y) ? 1.0f : 0.0f;
                        ymin = f2i(y0 + y_off);
                        ymax = ynext;
                        imin = f2u(f);
                    }
                } else if (f2u(f) < imax) {
                    ymin = ynext;
                    ymax = f2i(ceil32(s1.y));
                    imax = f2u(f);
                }
            }
            if (max32(s0.x, s1.x) > (float)bboxi[2]) {
                float f = round32((sign * ((float)bboxi[2] - x0) - b + fudge) / a);
                if ((x0 + sign * floor32(a * f + b) < (float)bboxi[2]) == d.is_positive_slope) f += 1.0f;
                if (d.is_positive_slope) { uint32_t fu = f2u(f); if (fu < imax) imax = fu; }
                else { uint32_t fu = f2u(f); if (fu > imin) imin = fu; }
            }
        }
        if (imin > imax) imax = imin;
        if (ymin < bboxi[1]) ymin = bboxi[1];
        if (ymax > bboxi[3]) ymax = bboxi[3];
        for (int32_t y = ymin; y < ymax; y++) {
            int32_t base = (int32_t)path->tiles + (y - bboxi[1]) * stride;
            tile[base].backdrop += delta;
        }
        float last_z = floor32(a * (float)(imin - 1) + b);
        uint32_t seg_base = bump_seg_counts;
        bump_seg_counts += imax - imin;
        for (uint32_t i = imin; i < imax; i++) {
            float zf = a * (float)i + b;
            float z = floor32(zf);
            int32_t y = f2i(y0 + (float)i - z);
            int32_t x = f2i(x0 + sign * z);
            int32_t base = (int32_t)path->tiles + (y - bboxi[1]) * stride - bboxi[0];
            int top_edge = (i == 0) ? (y0 == s0.y) : (last_z == z);
            if (top_edge && x + 1 < bboxi[2]) {
                int32_t x_bump = x + 1 > bboxi[0] ? x + 1 : bboxi[0];
                tile[base + x_bump].backdrop += delta;
            }
            uint32_t seg_within_slice = tile[base + x].seg_count_or_ix;
            tile[base + x].seg_count_or_ix++;
            if (seg_counts) {
                seg_counts[seg_base + i - imin].line_ix = line_ix;
                seg_counts[seg_base + i - imin].counts = (seg_within_slice << 16) | i;
            }
            last_z = z;
        }
    }
    return bump_seg_counts;
}

/* ------------------------------------------------------------------ path_tiling.go */
void ot_path_tiling(const ot_segment_count *seg_counts, uint32_t n_seg_counts,
                    const ot_line_soup *lines, const ot_path *paths, const ot_tile *tiles,
                    ot_path_segment *segments) {   /* path_tiling.go:11-199 */
    for (uint32_t seg_ix = 0; seg_ix < n_seg_counts; seg_ix++) {
        ot_segment_count sc = seg_counts[seg_ix];
        const ot_line_soup *line = &lines[sc.line_ix];
        uint32_t seg_within_slice = sc.counts >> 16;
        uint32_t seg_within_line = sc.counts & 0xffff;
        dda d; dda_setup(line, &d); dda_finish(&d);
        float a = d.a, b = d.b, sign = d.sign, x0 = d.x0, y0 = d.y0;
        vec2 xy0 = d.xy0, xy1 = d.xy1;
        uint32_t count = d.count;
        float z = floor32(a * (float)seg_within_line + b);
        int32_t x = f2i(x0) + f2i(sign * z);                       /* split truncation, path_tiling.go:74 */
        int32_t y = f2i(y0 + (float)seg_within_line - z);
        const ot_path *path = &paths[line->path_ix];
        int32_t bboxi[4] = {(int32_t)path->bbox[0], (int32_t)path->bbox[1], (int32_t)path->bbox[2], (int32_t)path->bbox[3]};
        int32_t stride = bboxi[2] - bboxi[0];
        int32_t tile_ix = (int32_t)path->tiles + (y - bboxi[1]) * stride + x - bboxi[0];
        ot_tile tile = tiles[tile_ix];
        uint32_t seg_start = ~tile.seg_count_or_ix;
        if ((int32_t)seg_start < 0) continue;
        vec2 tile_xy = v2((float)x * (float)TILE_W, (float)y * (float)TILE_H);
        vec2 tile_xy1 = vadd(tile_xy, v2((float)TILE_W, (float)TILE_H));
        if (seg_within_line > 0) {
            float z_prev = floor32(a * (float)(seg_within_line - 1) + b);
            if (z == z_prev) {
                float xt = xy0.x + (xy1.x - xy0.x) * (tile_xy.y - xy0.y) / (xy1.y - xy0.y);
                xt = clamp32(xt, tile_xy.x + 1e-3f, tile_xy1.x);
                xy0 = v2(xt, tile_xy.y);
            } else {
                float x_clip = d.is_positive_slope ? tile_xy.x : tile_xy1.x;
                float yt = xy0.y + (xy1.y - xy0.y) * (x_clip - xy0.x) / (xy1.x - xy0.x);
                yt = clamp32(yt, tile_xy.y + 1e-3f, tile_xy1.y);
                xy0 = v2(x_clip, yt);
            }
        }
        if (seg_within_line < count - 1) {
            float z_next = floor32(a * (float)(seg_within_line + 1) + b);
            if (z == z_next) {
                float xt = xy0.x + (xy1.x - xy0.x) * (tile_xy1.y - xy0.y) / (xy1.y - xy0.y);
                xt = clamp32(xt, tile_xy.x + 1e-3f, tile_xy1.x);
                xy1 = v2(xt, tile_xy1.y);
            } else {
                float x_clip = d.is_positive_slope ? tile_xy1.x : tile_xy.x;
                float yt = xy0.y + (xy1.y - xy0.y) * (x_clip - xy0.x) / (xy1.x - xy0.x);
                yt = clamp32(yt, tile_xy.y + 1e-3f, tile_xy1.y);
                xy1 = v2(x_clip, yt);
            }
        }
        float y_edge = 1e9f;
        vec2 p0o = vsub(xy0, tile_xy), p1o = vsub(xy1, tile_xy);
        const float epsilon = 1e-6f;
        if (p0o.x == 0.0f) {
            if (p1o.x == 0.0f) {
                p0o.x = epsilon;
                if (p0o.y == 0.0f) { p1o.x = epsilon; p1o.y = (float)TILE_H; }
                else { p1o.x = 2.0f * epsilon; p1o.y = p0o.y; }
            } else if (p0o.y == 0.0f) {
                p0o.x = epsilon;
            } else {
                y_edge = p0o.y;
            }
        } else if (p1o.x == 0.0f) {
            if (p1o.y == 0.0f) p1o.x = epsilon; else y_edge = p1o.y;
        }
        if (p0o.x == floor32(p0o.x) && p0o.x != 0.0f) p0o.x -= epsilon;
        if (p1o.x == floor32(p1o.x) && p1o.x != 0.0f) p1o.x -= epsilon;
        if (!d.is_down) { vec2 t = p0o; p0o = p1o; p1o = t; }
        ot_path_segment *o = &segments[seg_start + seg_within_slice];
        o->p0[0] = p0o.x; o->p0[1] = p0o.y; o->p1[0] = p1o.x; o->p1[1] = p1o.y; o->y_edge = y_edge;
    }
}

/* ------------------------------------------------------------------ coarse.go:169-223 / rasterizer.go:425-487 */
void ot_line_bbox(const ot_line_soup *lines, uint32_t n, int w, int h, uint32_t bbox[4]) {
    float min_x = 3.40282346638528859811704183484516925440e+38f, min_y = min_x, max_x = -min_x, max_y = -min_x;
    for (uint32_t i = 0; i < n; i++) {
        const float *pts[2] = {lines[i].p0, lines[i].p1};
        for (int k = 0; k < 2; k++) {
            const float *p = pts[k];
            if (p[0] < min_x) min_x = p[0];
            if (p[0] > max_x) max_x = p[0];
            if (p[1] < min_y) min_y = p[1];
            if (p[1] > max_y) max_y = p[1];
        }
    }
    if (min_x < 0) min_x = 0;
    if (min_y < 0) min_y = 0;
    if (max_x > (float)w) max_x = (float)w;
    if (max_y > (float)h) max_y = (float)h;
    /* Guard (deviation, documented in DESIGN.md): a path entirely off the canvas makes the reference's
     * uint32 conversions wrap (coarse.go:203-216) -- a huge or full-width bbox, i.e. a panic or wasted
     * empty tiles, never a visible pixel. Oracle and product both give such paths an empty bbox. */
    if (n == 0 || max_x < min_x || max_y < min_y) { bbox[0] = bbox[1] = bbox[2] = bbox[3] = 0; return; }
    /* Go: uint32(math.Floor(float64(minX / 16))) -- float64->uint32 goes through int64 on amd64 */
    uint32_t x0 = (uint32_t)(int64_t)floor((double)(min_x / (float)TILE_W));
    uint32_t y0 = (uint32_t)(int64_t)floor((double)(min_y / (float)TILE_H));
    uint32_t x1 = (uint32_t)(int64_t)ceil((double)(max_x / (float)TILE_W));
    uint32_t y1 = (uint32_t)(int64_t)ceil((double)(max_y / (float)TILE_H));
    uint32_t gx = (uint32_t)ceil((double)w / (double)TILE_W), gy = (uint32_t)ceil((double)h / (double)TILE_H);
    if (x1 > gx) x1 = gx;
    if (y1 > gy) y1 = gy;
    bbox[0] = x0; bbox[1] = y0; bbox[2] = x1; bbox[3] = y1;
}

/* ------------------------------------------------------------------ fine.go:219-289 */
static void fill_path(float *area, const ot_path_segment *segs, uint32_t n_segs, int32_t backdrop, int even_odd) {
    float backdrop_f = (float)backdrop;
    for (int i = 0; i < TILE_W * TILE_H; i++) area[i] = backdrop_f;
    for (uint32_t s = 0; s < n_segs; s++) {
        const ot_path_segment *seg = &segs[s];
        float delta0 = seg->p1[0] - seg->p0[0], delta1 = seg->p1[1] - seg->p0[1];
        for (int yi = 0; yi < TILE_H; yi++) {
            float y = seg->p0[1] - (float)yi;
            float y0 = clamp32(y, 0.0f, 1.0f);
            float y1 = clamp32(y + delta1, 0.0f, 1.0f);
            float dy = y0 - y1;
            float y_edge = signum32(delta0) * clamp32((float)yi - seg->y_edge + 1.0f, 0.0f, 1.0f);
            if (dy != 0.0f) {
                float vec_y_recip = 1.0f / delta1;
                float t0 = (y0 - y) * vec_y_recip;
                float t1 = (y1 - y) * vec_y_recip;
                float startx = seg->p0[0];
                float x0 = startx + t0 * delta0;
                float x1 = startx + t1 * delta0;
                float xmin0 = min32(x0, x1);
                float xmax0 = max32(x0, x1);
                for (int i = 0; i < TILE_W; i++) {
                    float i_f = (float)i;
                    float xmin = min32(xmin0 - i_f, 1.0f) - 1.0e-6f;
                    float xmax = xmax0 - i_f;
                    float b = min32(xmax, 1.0f);
                    float c = max32(b, 0.0f);
                    float d = max32(xmin, 0.0f);
                    float a = (b + 0.5f * (d * d - c * c) - xmin) / (xmax - xmin);
                    area[yi * TILE_W + i] += y_edge + a * dy;
                }
            } else if (y_edge != 0.0f) {
                for (int i = 0; i < TILE_W; i++) area[yi * TILE_W + i] += y_edge;
            }
        }
    }
    if (even_odd) {
        for (int i = 0; i < TILE_W * TILE_H; i++) area[i] = abs32(area[i] - 2.0f * round32(0.5f * area[i]));
    } else {
        for (int i = 0; i < TILE_W * TILE_H; i++) area[i] = min32(abs32(area[i]), 1.0f);
    }
}

/* ------------------------------------------------------------------ rasterizer.go:27-170 */
void ot_rasterize(const ot_line_soup *lines, uint32_t n_lines, int even_odd, int w, int h, float *alpha) {
    memset(alpha, 0, sizeof(float) * (size_t)w * h);
    if (n_lines == 0) return;
    ot_path path; path.tiles = 0;
    ot_line_bbox(lines, n_lines, w, h, path.bbox);
    int tiles_x = (int)(path.bbox[2] - path.bbox[0]), tiles_y = (int)(path.bbox[3] - path.bbox[1]);
    int tile_count = tiles_x * tiles_y;
    if (tile_count == 0) return;
    ot_tile *tiles = (ot_tile *)calloc((size_t)tile_count, sizeof(ot_tile));
    ot_line_soup *ll = (ot_line_soup *)malloc(sizeof(ot_line_soup) * n_lines);
    uint32_t max_sc = 0;
    for (uint32_t i = 0; i < n_lines; i++) {
        ll[i] = lines[i]; ll[i].path_ix = 0;
        dda d; dda_setup(&ll[i], &d); max_sc += d.count;
    }
    ot_segment_count *sc = (ot_segment_count *)calloc(max_sc ? max_sc : 1, sizeof(ot_segment_count));
    uint32_t n_sc = ot_path_count(ll, n_lines, &path, tiles, sc);
    uint32_t next = 0;
    for (int i = 0; i < tile_count; i++) {
        uint32_t n = tiles[i].seg_count_or_ix;
        if (n != 0) { tiles[i].seg_count_or_ix = ~next; next += n; }
    }
    uint32_t total = next;
    ot_path_segment *segs = (ot_path_segment *)calloc(total ? total : 1, sizeof(ot_path_segment));
    ot_path_tiling(sc, n_sc, ll, &path, tiles, segs);
    for (int y = 0; y < tiles_y; y++) {
        int32_t sum = 0;
        for (int x = 0; x < tiles_x; x++) { sum += tiles[y * tiles_x + x].backdrop; tiles[y * tiles_x + x].backdrop = sum; }
    }
    float area[TILE_W * TILE_H];
    for (int ty = 0; ty < tiles_y; ty++) for (int tx = 0; tx < tiles_x; tx++) {
        int tix = ty * tiles_x + tx;
        uint32_t seg_start = ~tiles[tix].seg_count_or_ix, n_ts = 0;
        const ot_path_segment *ts = NULL;
        if ((int32_t)seg_start >= 0) {
            uint32_t seg_end = total;
            for (int nx = tix + 1; nx < tile_count; nx++) {
                uint32_t ns = ~tiles[nx].seg_count_or_ix;
                if ((int32_t)ns >= 0) { seg_end = ns; break; }
            }
            if (seg_start < seg_end) { ts = segs + seg_start; n_ts = seg_end - seg_start; }
        }
        fill_path(area, ts, n_ts, tiles[tix].backdrop, even_odd);
        int gx = ((int)path.bbox[0] + tx) * TILE_W, gy = ((int)path.bbox[1] + ty) * TILE_H;
        for (int ly = 0; ly < TILE_H; ly++) {
            int py = gy + ly; if (py >= h) break;
            for (int lx = 0; lx < TILE_W; lx++) {
                int px = gx + lx; if (px >= w) break;
                alpha[py * w + px] = area[ly * TILE_W + lx];
            }
        }
    }
    free(tiles); free(ll); free(sc); free(segs);
}

/* compositor.go:24-59 */
static void blend_source_over(const uint8_t src[4], float alpha, uint8_t dst[4]) {
    if (alpha <= 0) return;
    if (alpha > 1.0f) alpha = 1.0f;
    float src_a = alpha * (float)src[3] / 255.0f;
    float sr = (float)src[0] / 255.0f * src_a, sg = (float)src[1] / 255.0f * src_a, sb = (float)src[2] / 255.0f * src_a;
    float dr = (float)dst[0] / 255.0f, dg = (float)dst[1] / 255.0f, db = (float)dst[2] / 255.0f, da = (float)dst[3] / 255.0f;
    float inv = 1.0f - src_a;
    float or_ = sr + dr * inv, og = sg + dg * inv, ob = sb + db * inv, oa = src_a + da * inv;
    dst[0] = (uint8_t)(or_ * 255.0f + 0.5f); dst[1] = (uint8_t)(og * 255.0f + 0.5f);
    dst[2] = (uint8_t)(ob * 255.0f + 0.5f); dst[3] = (uint8_t)(oa * 255.0f + 0.5f);
}

/* rasterizer.go:176-224 RasterizeScene */
void ot_rasterize_scene(const uint8_t bg[4], const ot_element *elems, uint32_t n_elems,
                        const ot_line_soup *lines, int w, int h, uint8_t *out) {
    for (int i = 0; i < w * h; i++) memcpy(out + 4 * (size_t)i, bg, 4);
    float *alpha = (float *)malloc(sizeof(float) * (size_t)w * h);
    for (uint32_t e = 0; e < n_elems; e++) {
        if (OT_ELEM_TYPE(elems[e].type) != OT_ELEM_DRAW || elems[e].line_count == 0) continue;
        ot_rasterize(lines + elems[e].line_start, elems[e].line_count, (int)elems[e].even_odd, w, h, alpha);
        for (int i = 0; i < w * h; i++) {
            if (alpha[i] <= 0) continue;
            blend_source_over(elems[e].color, alpha[i], out + 4 * (size_t)i);
        }
    }
    free(alpha);
}

/* ------------------------------------------------------------------ scene_encode.go + scans */
typedef struct { uint32_t *d; uint32_t n, cap; } u32vec;
static void u32_push(u32vec *v, uint32_t x) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 64; v->d = (uint32_t *)realloc(v->d, sizeof(uint32_t) * v->cap); }
    v->d[v->n++] = x;
}
static inline uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

enum { PTAG_LINETO = 0x9, PTAG_PATH = 0x10, PTAG_TRANSFORM = 0x20, PTAG_STYLE = 0x40 };      /* scene_encode.go:78-86 */
enum { CMD_END = 0, CMD_FILL = 1, CMD_SOLID = 3, CMD_COLOR = 5, CMD_BEGIN_CLIP = 10, CMD_END_CLIP = 11, /* ptcl.go:17-24 */
       CMD_GRAD = 6 };   /* ggcuda: {tag, gradient index} */

/* scene_encode.go:312-347 encodePath */
static void encode_path(u32vec *raw_tags, u32vec *path_data, u32vec *transforms, u32vec *styles,
                        const ot_line_soup *lines, uint32_t n, int even_odd) {
    u32_push(raw_tags, PTAG_TRANSFORM);
    const float ident[6] = {1, 0, 0, 1, 0, 0};
    for (int i = 0; i < 6; i++) u32_push(transforms, f32_bits(ident[i]));
    u32_push(raw_tags, PTAG_STYLE);
    u32_push(styles, even_odd ? 0x02u : 0u);
    int needs_move = 1; float last[2] = {0, 0};
    for (uint32_t i = 0; i < n; i++) {
        if (needs_move || lines[i].p0[0] != last[0] || lines[i].p0[1] != last[1]) {
            u32_push(raw_tags, PTAG_LINETO);
            u32_push(path_data, f32_bits(lines[i].p0[0])); u32_push(path_data, f32_bits(lines[i].p0[1]));
            needs_move = 0;
        }
        u32_push(raw_tags, PTAG_LINETO);
        u32_push(path_data, f32_bits(lines[i].p1[0])); u32_push(path_data, f32_bits(lines[i].p1[1]));
        last[0] = lines[i].p1[0]; last[1] = lines[i].p1[1];
    }
    u32_push(raw_tags, PTAG_PATH);
}

/* scene_encode.go:162-168 colour packing */
static uint32_t pack_color(const uint8_t c[4]) {
    float a = (float)c[3] / 255.0f;
    uint32_t r = f2u((float)c[0] * a + 0.5f), g = f2u((float)c[1] * a + 0.5f), b = f2u((float)c[2] * a + 0.5f);
    return r | (g << 8) | (b << 16) | ((uint32_t)c[3] << 24);
}

typedef struct { uint32_t *cmds; uint32_t n, cap; } ptcl;
static void ptcl_push(ptcl *p, uint32_t w) {
    if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 64; p->cmds = (uint32_t *)realloc(p->cmds, sizeof(uint32_t) * p->cap); }
    p->cmds[p->n++] = w;
}
typedef struct { uint32_t clip_depth, clip_zero_depth, blend_depth, max_blend_depth;   /* coarse.go:313-320 */
                 uint32_t *begin_pos; uint32_t n_begin, cap_begin; } tile_clip_state;   /* + PTCL position of every open BeginClip (ggcuda layer elision) */
/* ggcuda extension (not in the reference, which never maps layers to PTCL): a BeginClip whose blend word has
 * bit 31 set is removed again, together with its EndClip, when no command was written between them in a tile. */
#define OT_BLEND_ELIDE_EMPTY 0x80000000u
/* Bit 30: "implicit" layer -- no clip geometry, coverage 1 in every tile of the canvas (ggcuda's encoding of a
 * PushLayer without clip shape); BeginClip everywhere (retracted where nothing is drawn), CmdSolid + CmdEndClip. */
#define OT_BLEND_IMPLICIT 0x40000000u

/* coarse.go:651-679 tileSegRange */
static void tile_seg_range(ot_tile tile, int local_idx, int tile_count, const ot_tile *path_tiles, uint32_t total,
                           uint32_t *count, uint32_t *start) {
    uint32_t seg_start = ~tile.seg_count_or_ix;
    if ((int32_t)seg_start < 0) { *count = 0; *start = 0; return; }
    uint32_t seg_end = total;
    for (int nx = local_idx + 1; nx < tile_count; nx++) {
        uint32_t ns = ~path_tiles[nx].seg_count_or_ix;
        if ((int32_t)ns >= 0) { seg_end = ns; break; }
    }
    if (seg_end <= seg_start) { *count = 0; *start = seg_start; return; }
    *count = seg_end - seg_start; *start = seg_start;
}

int ot_style_per_path = 0;

ot_coarse *ot_coarse_run(const ot_element *elems, uint32_t n_elems, const ot_line_soup *lines_in, int w, int h) {
    ot_coarse *out = (ot_coarse *)calloc(1, sizeof(ot_coarse));
    /* --- EncodeSceneDef (scene_encode.go:240-308) --- */
    u32vec raw_tags = {0}, path_data = {0}, draw_tags = {0}, draw_data = {0}, transforms = {0}, styles = {0};
    uint32_t n_paths = 0, n_draw = 0, n_clips = 0;
    for (uint32_t e = 0; e < n_elems; e++) {
        const ot_element *el = &elems[e];
        switch (OT_ELEM_TYPE(el->type)) {
        case OT_ELEM_DRAW:
            encode_path(&raw_tags, &path_data, &transforms, &styles, lines_in + el->line_start, el->line_count, (int)el->even_odd);
            u32_push(&draw_tags, (el->type & OT_ELEM_GRADIENT) ? DRAWTAG_GRADIENT : DRAWTAG_COLOR);
            u32_push(&draw_data, (el->type & OT_ELEM_PACKED) ? el->packed_rgba : pack_color(el->color));
            n_paths++; n_draw++;
            break;
        case OT_ELEM_BEGIN_CLIP:
            encode_path(&raw_tags, &path_data, &transforms, &styles, lines_in + el->line_start, el->line_count, 0);
            u32_push(&draw_tags, DRAWTAG_BEGIN_CLIP);
            u32_push(&draw_data, el->blend); u32_push(&draw_data, f32_bits(el->alpha));
            n_paths++; n_draw++; n_clips++;
            break;
        case OT_ELEM_END_CLIP:
            /* reference quirk: no style word for the dummy path, so styles[path_ix] (coarse.go:683-705) is
             * shifted after the first clip pair. ot_style_per_path != 0 emits one (ggcuda's encoder does). */
            if (ot_style_per_path) { u32_push(&raw_tags, PTAG_STYLE); u32_push(&styles, 0); }
            u32_push(&raw_tags, PTAG_PATH);
            u32_push(&draw_tags, DRAWTAG_END_CLIP);
            n_paths++; n_draw++; n_clips++;
            break;
        }
    }
    /* packPathTags (scene_encode.go:189-198) + PackScene (:280-356) */
    uint32_t n_tag_words = (raw_tags.n + 3) / 4;
    uint32_t padded = ((n_tag_words + 255) / 256) * 256;
    if (padded == 0) padded = 256;
    ot_layout L; memset(&L, 0, sizeof L);
    L.n_draw_objects = n_draw; L.n_paths = n_paths; L.n_clips = n_clips;
    uint32_t off = 0;
    L.path_tag_base = off; off += padded;
    L.path_data_base = off; off += path_data.n;
    L.draw_tag_base = off; off += draw_tags.n;
    L.draw_data_base = off; off += draw_data.n;
    L.transform_base = off; off += transforms.n;
    L.style_base = off; off += styles.n;
    uint32_t *scene = (uint32_t *)calloc(off ? off : 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < raw_tags.n; i++) scene[L.path_tag_base + i / 4] |= (raw_tags.d[i] & 0xff) << ((i % 4) * 8);
    if (path_data.n) memcpy(scene + L.path_data_base, path_data.d, 4 * (size_t)path_data.n);
    if (draw_tags.n) memcpy(scene + L.draw_tag_base, draw_tags.d, 4 * (size_t)draw_tags.n);
    if (draw_data.n) memcpy(scene + L.draw_data_base, draw_data.d, 4 * (size_t)draw_data.n);
    if (transforms.n) memcpy(scene + L.transform_base, transforms.d, 4 * (size_t)transforms.n);
    if (styles.n) memcpy(scene + L.style_base, styles.d, 4 * (size_t)styles.n);
    out->scene = scene; out->n_scene_words = off; out->layout = L;
    free(raw_tags.d); free(path_data.d); free(draw_tags.d); free(draw_data.d); free(transforms.d); free(styles.d);

    /* --- pathtag / draw / clip-leaf scans (pathtag.go:76-121, draw_leaf.go:54-151, clip_leaf.go:27-56) --- */
    out->n_tag_words = padded;
    out->tag_monoids = (ot_path_monoid *)calloc(padded, sizeof(ot_path_monoid));
    ot_draw_monoid *dm = (ot_draw_monoid *)calloc(n_draw ? n_draw : 1, sizeof(ot_draw_monoid));
    uint32_t n_info = ot_scan_stages(scene, off, L.path_tag_base, padded, L.draw_tag_base, L.draw_data_base, n_draw, n_clips,
                                     out->tag_monoids, dm, NULL, NULL);
    uint32_t *info = (uint32_t *)calloc(n_info ? n_info : 1, sizeof(uint32_t));
    ot_scan_stages(scene, off, L.path_tag_base, padded, L.draw_tag_base, L.draw_data_base, n_draw, n_clips, NULL, dm, info, NULL);
    out->draw_monoids = dm; out->info = info; out->n_info = n_info;

    /* --- allLines with PathIx (rasterizer.go:363-384) --- */
    uint32_t n_lines = 0;
    for (uint32_t e = 0; e < n_elems; e++) if (OT_ELEM_TYPE(elems[e].type) != OT_ELEM_END_CLIP) n_lines += elems[e].line_count;
    ot_line_soup *all = (ot_line_soup *)malloc(sizeof(ot_line_soup) * (n_lines ? n_lines : 1));
    uint32_t *path_line_start = (uint32_t *)calloc(n_paths + 1, sizeof(uint32_t));
    {
        uint32_t k = 0, pix = 0;
        for (uint32_t e = 0; e < n_elems; e++) {
            path_line_start[pix] = k;
            if (OT_ELEM_TYPE(elems[e].type) != OT_ELEM_END_CLIP)
                for (uint32_t i = 0; i < elems[e].line_count; i++) { all[k] = lines_in[elems[e].line_start + i]; all[k].path_ix = pix; k++; }
            pix++;
        }
        path_line_start[n_paths] = k;
    }

    /* --- CoarseRasterize (coarse.go:59-154) --- */
    int wt = (w + TILE_W - 1) / TILE_W, ht = (h + TILE_H - 1) / TILE_H;
    out->width_in_tiles = wt; out->height_in_tiles = ht;
    int n_grid = wt * ht;
    ptcl *ptcls = (ptcl *)calloc((size_t)(n_grid ? n_grid : 1), sizeof(ptcl));
    for (int i = 0; i < n_grid; i++) ptcl_push(&ptcls[i], 0);   /* word 0 = blend_offset (ptcl.go:72-78) */
    out->n_paths = n_paths;
    out->paths = (ot_path *)calloc(n_paths ? n_paths : 1, sizeof(ot_path));
    out->path_seg_base = (uint32_t *)calloc(n_paths ? n_paths : 1, sizeof(uint32_t));
    out->path_total_segs = (uint32_t *)calloc(n_paths ? n_paths : 1, sizeof(uint32_t));
    ot_tile *tiles = NULL; uint32_t n_tiles = 0, cap_tiles = 0;
    ot_path_segment *segs = NULL; uint32_t n_segs = 0, cap_segs = 0;

    if (n_draw != 0 && n_paths != 0 && n_lines != 0) {
        uint32_t cur_tile_off = 0, global_seg_off = 0;
        for (uint32_t pix = 0; pix < n_paths; pix++) {
            const ot_line_soup *pl = all + path_line_start[pix];
            uint32_t npl = path_line_start[pix + 1] - path_line_start[pix];
            out->path_seg_base[pix] = global_seg_off;
            if (npl == 0) { out->paths[pix].tiles = cur_tile_off; continue; }
            ot_path path; path.tiles = cur_tile_off;
            ot_line_bbox(pl, npl, w, h, path.bbox);
            int bw = (int)(path.bbox[2] - path.bbox[0]), bh = (int)(path.bbox[3] - path.bbox[1]);
            uint32_t tile_count = (uint32_t)(bw * bh);
            out->paths[pix] = path;
            if (tile_count == 0) continue;
            if (n_tiles + tile_count > cap_tiles) { cap_tiles = (n_tiles + tile_count) * 2; tiles = (ot_tile *)realloc(tiles, sizeof(ot_tile) * cap_tiles); }
            ot_tile *pt = tiles + n_tiles;
            memset(pt, 0, sizeof(ot_tile) * tile_count);
            n_tiles += tile_count; cur_tile_off += tile_count;
            /* runPathStages (coarse.go:229-302) */
            ot_line_soup *ll = (ot_line_soup *)malloc(sizeof(ot_line_soup) * npl);
            uint32_t max_sc = 0;
            for (uint32_t i = 0; i < npl; i++) { ll[i] = pl[i]; ll[i].path_ix = 0; dda d; dda_setup(&ll[i], &d); max_sc += d.count; }
            ot_segment_count *sc = (ot_segment_count *)calloc(max_sc ? max_sc : 1, sizeof(ot_segment_count));
            ot_path local = path; local.tiles = 0;
            uint32_t n_sc = ot_path_count(ll, npl, &local, pt, sc);
            uint32_t next = 0;
            for (uint32_t i = 0; i < tile_count; i++) { uint32_t n = pt[i].seg_count_or_ix; if (n != 0) { pt[i].seg_count_or_ix = ~next; next += n; } }
            if (n_segs + next > cap_segs) { cap_segs = (n_segs + next) * 2 + 16; segs = (ot_path_segment *)realloc(segs, sizeof(ot_path_segment) * cap_segs); }
            memset(segs + n_segs, 0, sizeof(ot_path_segment) * next);
            ot_path_tiling(sc, n_sc, ll, &local, pt, segs + n_segs);
            for (int y = 0; y < bh; y++) { int32_t sum = 0; for (int x = 0; x < bw; x++) { sum += pt[y * bw + x].backdrop; pt[y * bw + x].backdrop = sum; } }
            out->path_total_segs[pix] = next;
            n_segs += next; global_seg_off += next;
            free(ll); free(sc);
        }

        /* generatePTCLs (coarse.go:322-376) */
        tile_clip_state *cs_all = (tile_clip_state *)calloc((size_t)n_grid, sizeof(tile_clip_state));
        /* extractPathFillRules (coarse.go:683-705): style i for path i (reference quirk kept) */
        uint32_t style_count = L.transform_base - L.style_base;   /* uint32 wrap, as in the reference */
        for (uint32_t draw_ix = 0; draw_ix < n_draw; draw_ix++) {
            uint32_t tag = scene[L.draw_tag_base + draw_ix];
            ot_draw_monoid m = dm[draw_ix];
            uint32_t pix = m.path_ix;
            if (pix >= n_paths) continue;
            ot_path path = out->paths[pix];
            int bw = (int)(path.bbox[2] - path.bbox[0]), bh = (int)(path.bbox[3] - path.bbox[1]);
            if (tag == DRAWTAG_COLOR || tag == DRAWTAG_GRADIENT) {
                const uint32_t cmd_paint = tag == DRAWTAG_GRADIENT ? CMD_GRAD : CMD_COLOR;
                uint32_t rgba = m.info_offset < n_info ? info[m.info_offset] : 0;
                int even_odd = 0;
                if (pix < style_count && L.style_base + pix < off) even_odd = (scene[L.style_base + pix] & 0x02) != 0;
                uint32_t gsb = out->path_seg_base[pix], total = out->path_total_segs[pix];
                if (bw == 0 || bh == 0) continue;
                for (int ty = 0; ty < bh; ty++) for (int tx = 0; tx < bw; tx++) {   /* emitDrawToTilesClipAware :380-432 */
                    int gtx = (int)path.bbox[0] + tx, gty = (int)path.bbox[1] + ty;
                    if (gtx < 0 || gtx >= wt || gty < 0 || gty >= ht) continue;
                    int g = gty * wt + gtx;
                    if (cs_all[g].clip_zero_depth > 0) continue;
                    int li = ty * bw + tx;
                    uint32_t tix = path.tiles + (uint32_t)li;
                    if (tix >= n_tiles) continue;
                    ot_tile t = tiles[tix];
                    uint32_t cnt, st;
                    tile_seg_range(t, li, bw * bh, tiles + path.tiles, total, &cnt, &st);
                    if (cnt > 0) {
                        ptcl_push(&ptcls[g], CMD_FILL); ptcl_push(&ptcls[g], (cnt << 1) | (uint32_t)even_odd);
                        ptcl_push(&ptcls[g], gsb + st); ptcl_push(&ptcls[g], (uint32_t)t.backdrop);
                        ptcl_push(&ptcls[g], cmd_paint); ptcl_push(&ptcls[g], rgba);
                    } else if (t.backdrop != 0 && (!even_odd || ot_evenodd_solid_quirk || (t.backdrop & 1))) {
                        /* DEVIATION from coarse.go:425, which paints every tile with backdrop != 0 solid whatever the
                         * fill rule: under even-odd a tile without segments is inside only when its winding is odd
                         * (two nested same-direction contours leave backdrop 2 in the hole). ot_evenodd_solid_quirk = 1
                         * restores the reference's behaviour. */
                        ptcl_push(&ptcls[g], CMD_SOLID);
                        ptcl_push(&ptcls[g], cmd_paint); ptcl_push(&ptcls[g], rgba);
                    }
                }
            } else if (tag == DRAWTAG_BEGIN_CLIP && (scene[L.draw_data_base + m.scene_offset] & OT_BLEND_IMPLICIT)) {
                for (int g = 0; g < n_grid; g++) {   /* implicit layer: every tile is "inside with full coverage" */
                    tile_clip_state *cs = &cs_all[g];
                    if (cs->clip_zero_depth > 0) { cs->clip_depth++; continue; }
                    if (cs->n_begin == cs->cap_begin) { cs->cap_begin = cs->cap_begin ? cs->cap_begin * 2 : 4; cs->begin_pos = (uint32_t *)realloc(cs->begin_pos, 4 * cs->cap_begin); }
                    cs->begin_pos[cs->n_begin++] = ptcls[g].n;
                    ptcl_push(&ptcls[g], CMD_BEGIN_CLIP);
                    cs->blend_depth++;
                    if (cs->blend_depth > cs->max_blend_depth) cs->max_blend_depth = cs->blend_depth;
                    cs->clip_depth++;
                }
            } else if (tag == DRAWTAG_BEGIN_CLIP) {   /* emitBeginClipToTiles :442-507 */
                for (int ty = 0; ty < bh; ty++) for (int tx = 0; tx < bw; tx++) {
                    int gtx = (int)path.bbox[0] + tx, gty = (int)path.bbox[1] + ty;
                    if (gtx < 0 || gtx >= wt || gty < 0 || gty >= ht) continue;
                    int g = gty * wt + gtx;
                    tile_clip_state *cs = &cs_all[g];
                    if (cs->clip_zero_depth > 0) { cs->clip_depth++; continue; }
                    uint32_t tix = path.tiles + (uint32_t)(ty * bw + tx);
                    int has_seg = 0, has_bd = 0;
                    if (tix < n_tiles) { has_seg = tiles[tix].seg_count_or_ix != 0; has_bd = tiles[tix].backdrop != 0; }
                    if (!has_seg && !has_bd) {
                        cs->clip_zero_depth = cs->clip_depth + 1;
                    } else {
                        if (cs->n_begin == cs->cap_begin) { cs->cap_begin = cs->cap_begin ? cs->cap_begin * 2 : 4; cs->begin_pos = (uint32_t *)realloc(cs->begin_pos, 4 * cs->cap_begin); }
                        cs->begin_pos[cs->n_begin++] = ptcls[g].n;
                        ptcl_push(&ptcls[g], CMD_BEGIN_CLIP);
                        cs->blend_depth++;
                        if (cs->blend_depth > cs->max_blend_depth) cs->max_blend_depth = cs->blend_depth;
                    }
                    cs->clip_depth++;
                }
                for (int ty = 0; ty < ht; ty++) for (int tx = 0; tx < wt; tx++) {
                    if ((uint32_t)tx >= path.bbox[0] && (uint32_t)tx < path.bbox[2] && (uint32_t)ty >= path.bbox[1] && (uint32_t)ty < path.bbox[3]) continue;
                    tile_clip_state *cs = &cs_all[ty * wt + tx];
                    if (cs->clip_zero_depth == 0) cs->clip_zero_depth = cs->clip_depth + 1;
                    cs->clip_depth++;
                }
            } else if (tag == DRAWTAG_END_CLIP) {   /* :360-372 + emitEndClipToTiles :517-563 + endClipForTile :566-625 */
                uint32_t blend = 0; float alpha = 0;
                uint32_t so = L.draw_data_base + m.scene_offset;
                if (so + 1 < off) { blend = scene[so]; alpha = bits_f32(scene[so + 1]); }
                if (blend & OT_BLEND_IMPLICIT) {
                    for (int g = 0; g < n_grid; g++) {
                        tile_clip_state *cs = &cs_all[g];
                        cs->clip_depth--;
                        if (cs->clip_zero_depth == cs->clip_depth + 1) { cs->clip_zero_depth = 0; continue; }
                        if (cs->clip_zero_depth > 0) continue;
                        uint32_t bpos = cs->n_begin ? cs->begin_pos[--cs->n_begin] : 0;
                        if ((blend & OT_BLEND_ELIDE_EMPTY) && ptcls[g].n == bpos + 1) { ptcls[g].n = bpos; cs->blend_depth--; continue; }
                        ptcl_push(&ptcls[g], CMD_SOLID);
                        ptcl_push(&ptcls[g], CMD_END_CLIP); ptcl_push(&ptcls[g], blend); ptcl_push(&ptcls[g], f32_bits(alpha));
                        cs->blend_depth--;
                    }
                    continue;
                }
                for (int ty = 0; ty < bh; ty++) for (int tx = 0; tx < bw; tx++) {
                    int gtx = (int)path.bbox[0] + tx, gty = (int)path.bbox[1] + ty;
                    if (gtx < 0 || gtx >= wt || gty < 0 || gty >= ht) continue;
                    int g = gty * wt + gtx;
                    tile_clip_state *cs = &cs_all[g];
                    cs->clip_depth--;
                    if (cs->clip_zero_depth == cs->clip_depth + 1) { cs->clip_zero_depth = 0; continue; }
                    if (cs->clip_zero_depth > 0) continue;
                    int li = ty * bw + tx;
                    uint32_t tix = path.tiles + (uint32_t)li;
                    uint32_t bpos = cs->n_begin ? cs->begin_pos[--cs->n_begin] : 0;
                    if ((blend & OT_BLEND_ELIDE_EMPTY) && ptcls[g].n == bpos + 1) {   /* nothing between Begin and End */
                        ptcls[g].n = bpos;
                        cs->blend_depth--;
                        continue;
                    }
                    if (tix < n_tiles) {
                        ot_tile t = tiles[tix];
                        uint32_t cnt, st;
                        tile_seg_range(t, li, bw * bh, tiles + path.tiles, out->path_total_segs[pix], &cnt, &st);
                        if (cnt > 0) {
                            ptcl_push(&ptcls[g], CMD_FILL); ptcl_push(&ptcls[g], (cnt << 1));
                            ptcl_push(&ptcls[g], out->path_seg_base[pix] + st); ptcl_push(&ptcls[g], (uint32_t)t.backdrop);
                        } else {
                            ptcl_push(&ptcls[g], CMD_SOLID);
                        }
                        ptcl_push(&ptcls[g], CMD_END_CLIP); ptcl_push(&ptcls[g], blend); ptcl_push(&ptcls[g], f32_bits(alpha));
                    }
                    cs->blend_depth--;
                }
                for (int ty = 0; ty < ht; ty++) for (int tx = 0; tx < wt; tx++) {
                    if ((uint32_t)tx >= path.bbox[0] && (uint32_t)tx < path.bbox[2] && (uint32_t)ty >= path.bbox[1] && (uint32_t)ty < path.bbox[3]) continue;
                    tile_clip_state *cs = &cs_all[ty * wt + tx];
                    cs->clip_depth--;
                    if (cs->clip_zero_depth == cs->clip_depth + 1) cs->clip_zero_depth = 0;
                }
            }
        }
        for (int i = 0; i < n_grid; i++) free(cs_all[i].begin_pos);
        free(cs_all);
    }
    /* WriteEnd on every tile (coarse.go:148-151) and flatten to one array */
    out->ptcl_offsets = (uint32_t *)calloc((size_t)n_grid + 1, sizeof(uint32_t));
    uint32_t total_words = 0;
    for (int i = 0; i < n_grid; i++) { ptcl_push(&ptcls[i], CMD_END); out->ptcl_offsets[i] = total_words; total_words += ptcls[i].n; }
    out->ptcl_offsets[n_grid] = total_words;
    out->ptcl_words = (uint32_t *)malloc(sizeof(uint32_t) * (total_words ? total_words : 1));
    for (int i = 0; i < n_grid; i++) { memcpy(out->ptcl_words + out->ptcl_offsets[i], ptcls[i].cmds, 4 * (size_t)ptcls[i].n); free(ptcls[i].cmds); }
    free(ptcls);
    out->tiles = tiles; out->n_tiles = n_tiles; out->segments = segs; out->n_segments = n_segs;
    free(all); free(path_line_start);
    return out;
}

void ot_coarse_free(ot_coarse *c) {
    if (!c) return;
    free(c->paths); free(c->tiles); free(c->segments); free(c->path_seg_base); free(c->path_total_segs);
    free(c->ptcl_offsets); free(c->ptcl_words); free(c->scene); free(c->tag_monoids); free(c->draw_monoids); free(c->info); free(c->gtab);
    free(c);
}

/* ------------------------------------------------------------------ fine.go:40-187 */
/* gg's gradient colour at a point (gradient_linear.go:52-66, gradient_radial.go computeTSimple, gradient.go:42-131), from a
 * record of ggcuda's gradient table: the raw geometry and the sorted stops, evaluated like the Go code (float64 geometry,
 * float32 linear-light interpolation) -- NOT from the device's colour ramp. Straight RGBA out. */
static float og_srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : (float)pow((double)((s + 0.055f) / 1.055f), 2.4); }
static float og_linear_to_srgb(float l) { return l <= 0.0031308f ? l * 12.92f : 1.055f * (float)pow((double)l, 1.0 / 2.4) - 0.055f; }
static void og_color_at(const uint32_t *gtab, uint32_t idx, double x, double y, float out[4]) {
    const uint32_t *g = gtab + 16 * idx;
    uint32_t kind = g[0], extend = g[1], n = g[2];
    const uint32_t *st = gtab + g[3];   /* 8 floats per stop: offset, r g b a, (linear-light r g b: the device's, unused here) */
    double t = 0; int first_only = 0;
    double g0 = bits_f32(g[5]), g1 = bits_f32(g[6]), g2 = bits_f32(g[7]), g3 = bits_f32(g[8]);
    if (kind == 0) {
        double dx = g2 - g0, dy = g3 - g1, l2 = dx * dx + dy * dy;
        if (l2 == 0) first_only = 1; else t = ((x - g0) * dx + (y - g1) * dy) / l2;
    } else if (kind == 1) {
        double dx = x - g0, dy = y - g1, rd = g3 - g2;
        if (rd == 0) first_only = 1; else t = (sqrt(dx * dx + dy * dy) - g2) / rd;
    } else if (kind == 4) {   /* gradient_sweep.go:79-150 */
        double dx = x - g0, dy = y - g1, sweep = g3 - g2;
        if (dx == 0 && dy == 0) first_only = 1;
        else if (sweep != 0) {
            const double two_pi = 2 * 3.14159265358979323846;
            double rel = atan2(dy, dx) - g2;
            if (sweep > 0) { for (int k = 0; k < 64 && rel < 0; k++) rel += two_pi; for (int k = 0; k < 64 && rel >= two_pi; k++) rel -= two_pi; }
            else { for (int k = 0; k < 64 && rel > 0; k++) rel -= two_pi; for (int k = 0; k < 64 && rel <= -two_pi; k++) rel += two_pi; }
            t = rel / sweep;
        }
    } else {                  /* gradient_radial.go:84-196, focus off the centre */
        double fx0 = bits_f32(g[9]), fy0 = bits_f32(g[10]);
        if (g3 - g2 == 0) first_only = 1;
        else {
            double dx = x - fx0, dy = y - fy0, fx = g0 - fx0, fy = g1 - fy0;
            double a = dx * dx + dy * dy, b = -2 * (dx * fx + dy * fy), c = fx * fx + fy * fy - g3 * g3;
            if (a != 0) {
                double disc = b * b - 4 * a * c;
                if (disc < 0) t = 1;
                else {
                    double sq = sqrt(disc), t1 = (-b - sq) / (2 * a), t2 = (-b + sq) / (2 * a), tt = 0;
                    int hit = 1;
                    if (t1 > 0 && t2 > 0) tt = t1 < t2 ? t1 : t2; else if (t1 > 0) tt = t1; else if (t2 > 0) tt = t2; else hit = 0;
                    if (hit) { double pd = sqrt(a), idist = tt * pd; if (idist != 0) t = pd / idist; }
                }
            }
        }
    }
    if (n == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    if (first_only || n == 1) { for (int k = 0; k < 4; k++) out[k] = bits_f32(st[1 + k]); return; }
    if (extend == 1) { t -= floor(t); if (t < 0) t += 1; }
    else if (extend == 2) { t = fabs(t); double per = floor(t); t -= per; if (((long long)per) % 2 == 1) t = 1 - t; }
    else t = t < 0 ? 0 : (t > 1 ? 1 : t);
    uint32_t i = 0;
    while (i < n && !((double)bits_f32(st[8 * i]) >= t)) i++;
    const uint32_t *a = NULL;
    if (i == 0) a = st; else if (i >= n) a = st + 8 * (n - 1); else if (bits_f32(st[8 * i]) == bits_f32(st[8 * (i - 1)])) a = st + 8 * (i - 1);
    if (a) { for (int k = 0; k < 4; k++) out[k] = bits_f32(a[1 + k]); return; }
    const uint32_t *s1 = st + 8 * (i - 1), *s2 = st + 8 * i;
    float lt = (float)((t - (double)bits_f32(s1[0])) / ((double)bits_f32(s2[0]) - (double)bits_f32(s1[0])));
    for (int k = 0; k < 3; k++) {
        float l1 = og_srgb_to_linear(bits_f32(s1[1 + k])), l2v = og_srgb_to_linear(bits_f32(s2[1 + k]));
        out[k] = og_linear_to_srgb(l1 + lt * (l2v - l1));
    }
    out[3] = bits_f32(s1[4]) + lt * (bits_f32(s2[4]) - bits_f32(s1[4]));
}

void ot_fine_tile(const uint32_t *cmds, uint32_t n_words, const ot_path_segment *segs, uint32_t n_segs,
                  const float bg[4], float *rgba_out) {
    ot_fine_tile_at(cmds, n_words, segs, n_segs, bg, rgba_out, 0, 0, NULL);
}
void ot_fine_tile_at(const uint32_t *cmds, uint32_t n_words, const ot_path_segment *segs, uint32_t n_segs,
                     const float bg[4], float *rgba_out, int origin_x, int origin_y, const uint32_t *gtab) {
    enum { PC = TILE_W * TILE_H, SPLIT = 4 };
    float (*rgba)[4] = (float (*)[4])rgba_out;
    for (int i = 0; i < PC; i++) memcpy(rgba[i], bg, 16);
    if (!cmds) return;
    float area[PC];
    for (int i = 0; i < PC; i++) area[i] = 0;
    /* blend stack: first 4 levels "registers", deeper levels spill (fine.go:58-62) */
    float (*stack)[PC][4] = NULL; uint32_t stack_cap = 0;
    uint32_t clip_depth = 0;
    uint32_t off = 1;   /* CmdStartOffset, ptcl.go:98 */
    for (;;) {
        uint32_t tag = off < n_words ? cmds[off] : CMD_END;
        if (off < n_words) off++;
        switch (tag) {
        case CMD_END: goto done;
        case CMD_FILL: {
            uint32_t packed = cmds[off], seg_index = cmds[off + 1]; int32_t backdrop = (int32_t)cmds[off + 2];
            off += 3;
            uint32_t seg_end = seg_index + (packed >> 1);
            if (seg_end > n_segs) seg_end = n_segs;
            uint32_t cnt = seg_end > seg_index ? seg_end - seg_index : 0;
            fill_path(area, segs + seg_index, cnt, backdrop, (int)(packed & 1));
        } break;
        case CMD_SOLID:
            for (int i = 0; i < PC; i++) area[i] = 1.0f;
            break;
        case CMD_COLOR: {
            uint32_t c = cmds[off++];
            float r = (float)(c & 0xff) / 255.0f, g = (float)((c >> 8) & 0xff) / 255.0f;
            float b = (float)((c >> 16) & 0xff) / 255.0f, a = (float)((c >> 24) & 0xff) / 255.0f;
            for (int i = 0; i < PC; i++) {
                float cov = area[i];
                float fr = r * cov, fg = g * cov, fb = b * cov, fa = a * cov;
                float inv = 1.0f - fa;
                rgba[i][0] = rgba[i][0] * inv + fr; rgba[i][1] = rgba[i][1] * inv + fg;
                rgba[i][2] = rgba[i][2] * inv + fb; rgba[i][3] = rgba[i][3] * inv + fa;
                if (ot_truncate_per_draw && cov != 0.0f)   /* gg's CPU pixmap: 8 bits, truncated, after every draw that touches the pixel
                                                             * (software.go:1003-1024, pixmap.go:218-228); 1/512 absorbs the float32 error of k/255*255 */
                    for (int k = 0; k < 4; k++) { float v = rgba[i][k] * 255.0f + (1.0f / 512.0f); v = v < 0 ? 0 : (v > 255.0f ? 255.0f : v); rgba[i][k] = floorf(v) / 255.0f; }
            }
        } break;
        case CMD_GRAD: {   /* CmdColor with the brush's colour at the pixel centre (software.go:1086-1090) */
            uint32_t gi = cmds[off++];
            if (gtab && gtab[16 * gi] == 2u) {
                /* TagFillRoundRect: sdfRoundRectCoverage at the pixel centre (scene/shape.go:246-274, scene/renderer.go:986-1043), float32 */
                const uint32_t *g = gtab + 16 * gi;
                float cx = bits_f32(g[5]), cy = bits_f32(g[6]), hw = bits_f32(g[7]), hh = bits_f32(g[8]), rad = bits_f32(g[9]);
                uint32_t pc = g[10];
                float cr = (float)(pc & 0xff) / 255.0f, cg = (float)((pc >> 8) & 0xff) / 255.0f, cb = (float)((pc >> 16) & 0xff) / 255.0f, ca = (float)(pc >> 24) / 255.0f;
                for (int i = 0; i < PC; i++) {
                    float px = (float)(origin_x + i % TILE_W) + 0.5f, py = (float)(origin_y + i / TILE_W) + 0.5f;
                    float dx = fabsf(px - cx) - hw + rad, dy = fabsf(py - cy) - hh + rad;
                    float mx = dx > 0 ? dx : 0, my = dy > 0 ? dy : 0;
                    float outside = (float)sqrt((double)(mx * mx + my * my));
                    float inside = (dx > dy ? dx : dy); if (inside > 0) inside = 0;
                    float dist = outside + inside - rad, cov;
                    if (dist >= 0.7f) cov = 0;
                    else if (dist <= -0.7f) cov = 1;
                    else { float t = (dist + 0.7f) / (2 * 0.7f); cov = 1 - (t * t * (3 - 2 * t)); }
                    float fa = ca * cov, inv = 1.0f - fa;
                    rgba[i][0] = rgba[i][0] * inv + cr * cov; rgba[i][1] = rgba[i][1] * inv + cg * cov;
                    rgba[i][2] = rgba[i][2] * inv + cb * cov; rgba[i][3] = rgba[i][3] * inv + fa;
                }
                break;
            }
            if (gtab && gtab[16 * gi] == 3u) {
                /* TagImage: blitImageToTile (scene/renderer.go:1093-1243), float32 in its order of operations; the pixmap of the
                 * reference holds bytes and rounds after every draw, this one composites in float32 like every other command */
                const uint32_t *g = gtab + 16 * gi;
                const int iw = (int)g[1], ih = (int)g[2];
                const uint32_t *tex = gtab + g[3];
                const float iA = bits_f32(g[5]), iB = bits_f32(g[6]), iC = bits_f32(g[7]), iD = bits_f32(g[8]), iE = bits_f32(g[9]), iF = bits_f32(g[10]);
                for (int i = 0; i < PC; i++) {
                    const int px = origin_x + i % TILE_W, py = origin_y + i / TILE_W;
                    if (px < (int32_t)g[11] || py < (int32_t)g[12] || px > (int32_t)g[13] || py > (int32_t)g[14]) continue;
                    const float cxp = (float)px + 0.5f, cyp = (float)py + 0.5f;
                    float sx = iA * cxp + iB * cyp + iC, sy = iD * cxp + iE * cyp + iF;
                    sx -= 0.5f; sy -= 0.5f;
                    const float flx = floorf(sx), fly = floorf(sy);
                    if (!(fabsf(flx) < 1.0e9f) || !(fabsf(fly) < 1.0e9f)) continue;
                    const int ix0 = (int)flx, iy0 = (int)fly;
                    if (ix0 + 1 < 0 || iy0 + 1 < 0 || ix0 >= iw || iy0 >= ih) continue;
                    const float wx = sx - flx, wy = sy - fly;
                    float s4[4];
                    if (wx == 0.0f && wy == 0.0f) {
                        if (ix0 < 0 || iy0 < 0) continue;
                        const uint32_t t = tex[iy0 * iw + ix0];
                        if ((t >> 24) == 0u) continue;
                        for (int k = 0; k < 4; k++) s4[k] = (float)((t >> (8 * k)) & 0xffu) / 255.0f;
                    } else {
                        const int cx0 = ix0 < 0 ? 0 : (ix0 > iw - 1 ? iw - 1 : ix0), cx1 = ix0 + 1 < 0 ? 0 : (ix0 + 1 > iw - 1 ? iw - 1 : ix0 + 1);
                        const int cy0 = iy0 < 0 ? 0 : (iy0 > ih - 1 ? ih - 1 : iy0), cy1 = iy0 + 1 < 0 ? 0 : (iy0 + 1 > ih - 1 ? ih - 1 : iy0 + 1);
                        uint32_t t00 = tex[cy0 * iw + cx0], t10 = tex[cy0 * iw + cx1], t01 = tex[cy1 * iw + cx0], t11 = tex[cy1 * iw + cx1];
                        if ((t00 >> 24) == 0u) t00 = 0u;
                        if ((t10 >> 24) == 0u) t10 = 0u;
                        if ((t01 >> 24) == 0u) t01 = 0u;
                        if ((t11 >> 24) == 0u) t11 = 0u;
                        const float ifx = 1.0f - wx, ify = 1.0f - wy;
                        const float w00 = ifx * ify, w10 = wx * ify, w01 = ifx * wy, w11 = wx * wy;
                        for (int k = 0; k < 4; k++)
                            s4[k] = (float)((t00 >> (8 * k)) & 0xffu) * w00 + (float)((t10 >> (8 * k)) & 0xffu) * w10 +
                                    (float)((t01 >> (8 * k)) & 0xffu) * w01 + (float)((t11 >> (8 * k)) & 0xffu) * w11;
                        if (s4[3] < 0.5f / 255.0f) continue;
                        for (int k = 0; k < 4; k++) s4[k] = s4[k] / 255.0f;
                    }
                    const float inv = 1.0f - s4[3];
                    rgba[i][0] = rgba[i][0] * inv + s4[0]; rgba[i][1] = rgba[i][1] * inv + s4[1];
                    rgba[i][2] = rgba[i][2] * inv + s4[2]; rgba[i][3] = rgba[i][3] * inv + s4[3];
                }
                break;
            }
            for (int i = 0; i < PC; i++) {
                float c[4] = {0, 0, 0, 0};
                if (gtab) og_color_at(gtab, gi, (double)(origin_x + i % TILE_W) + 0.5, (double)(origin_y + i / TILE_W) + 0.5, c);
                float cov = area[i];
                float fa = c[3] * cov, inv = 1.0f - fa;
                rgba[i][0] = rgba[i][0] * inv + c[0] * c[3] * cov; rgba[i][1] = rgba[i][1] * inv + c[1] * c[3] * cov;
                rgba[i][2] = rgba[i][2] * inv + c[2] * c[3] * cov; rgba[i][3] = rgba[i][3] * inv + fa;
            }
        } break;
        case CMD_BEGIN_CLIP:
            if (clip_depth >= stack_cap) { stack_cap = stack_cap ? stack_cap * 2 : 8; stack = realloc(stack, sizeof(*stack) * stack_cap); }
            memcpy(stack[clip_depth], rgba, sizeof(*stack));
            clip_depth++;
            for (int i = 0; i < PC; i++) rgba[i][0] = rgba[i][1] = rgba[i][2] = rgba[i][3] = 0;
            break;
        case CMD_END_CLIP: {
            /* cmds[off] = blend word. The reference carries whatever SceneElement.BlendMode holds (scene_encode.go:233) and
             * its fine stage ignores it (fine.go:164: source-over only). ggcuda gives the word a meaning -- (mix << 8) |
             * compose, composite defined by ot_blend_f32 (oracle/blend.c) -- so the reference's own clip vectors, written
             * with blend 0, correspond to 0x8003 (clip) or 3 (Normal / SrcOver) here, for which ot_blend_f32 is exactly the
             * reference's saved*(1-fg.a)+fg; 0 itself is BlendClear. */
            uint32_t blend = cmds[off];
            float alpha = bits_f32(cmds[off + 1]);
            off += 2;
            if (clip_depth == 0) continue;
            clip_depth--;
            float (*saved)[4] = stack[clip_depth];
            /* Coverage: scaling the source by it (fine.go:152-160) is the same as out = D + cov (blend(S, D) - D) for every
             * mode whose backdrop factor is 1 under a transparent source; for the six that wipe their backdrop (Clear, Copy,
             * SrcIn, DestIn, SrcOut, DestAtop) it is not -- a pixel the layer's clip does not cover would be wiped too --
             * so those are blended at full strength and interpolated. */
            uint32_t bmix = (blend >> 8) & 0xffu, bcomp = blend & 0xffu;
            int wipe = bmix == 0 && (bcomp == 0 || bcomp == 1 || bcomp == 5 || bcomp == 6 || bcomp == 7 || bcomp == 10);
            for (int i = 0; i < PC; i++) {
                float scale = wipe ? alpha : area[i] * alpha;
                float fgc[4] = {rgba[i][0] * scale, rgba[i][1] * scale, rgba[i][2] * scale, rgba[i][3] * scale};
                ot_blend_f32(blend, saved[i], fgc, rgba[i]);
                if (wipe) for (int k = 0; k < 4; k++) rgba[i][k] = saved[i][k] + area[i] * (rgba[i][k] - saved[i][k]);
            }
        } break;
        default: goto done;
        }
    }
done:
    free(stack);
    (void)SPLIT;
}

/* fine.go:190-217 */
static void premul_to_straight_u8(const float pm[4], uint8_t out[4]) {
    float a = pm[3];
    if (a <= 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    if (a > 1.0f) a = 1.0f;
    float r = pm[0] / a; if (r > 1.0f) r = 1.0f;
    float g = pm[1] / a; if (g > 1.0f) g = 1.0f;
    float b = pm[2] / a; if (b > 1.0f) b = 1.0f;
    out[0] = (uint8_t)(r * 255.0f + 0.5f); out[1] = (uint8_t)(g * 255.0f + 0.5f);
    out[2] = (uint8_t)(b * 255.0f + 0.5f); out[3] = (uint8_t)(a * 255.0f + 0.5f);
}

void ot_fine_frame(const ot_coarse *c, const uint8_t bg[4], int w, int h, uint8_t *out_straight, uint8_t *out_premul) {
    float bg_a = (float)bg[3] / 255.0f;   /* rasterizer.go:236-243 */
    float bgf[4] = {(float)bg[0] / 255.0f * bg_a, (float)bg[1] / 255.0f * bg_a, (float)bg[2] / 255.0f * bg_a, bg_a};
    float px[TILE_W * TILE_H * 4];
    for (int ty = 0; ty < c->height_in_tiles; ty++) for (int tx = 0; tx < c->width_in_tiles; tx++) {
        int t = ty * c->width_in_tiles + tx;
        ot_fine_tile_at(c->ptcl_words + c->ptcl_offsets[t], c->ptcl_offsets[t + 1] - c->ptcl_offsets[t],
                        c->segments, c->n_segments, bgf, px, tx * TILE_W, ty * TILE_H, c->gtab);
        for (int ly = 0; ly < TILE_H; ly++) {
            int py = ty * TILE_H + ly; if (py >= h) break;
            for (int lx = 0; lx < TILE_W; lx++) {
                int pxx = tx * TILE_W + lx; if (pxx >= w) break;
                const float *p = px + 4 * (ly * TILE_W + lx);
                size_t o = 4 * ((size_t)py * w + pxx);
                if (out_straight) premul_to_straight_u8(p, out_straight + o);
                if (out_premul) {   /* fine.wgsl:305-323 */
                    for (int k = 0; k < 4; k++) { float v = clamp32(p[k], 0.0f, 1.0f); out_premul[o + k] = (uint8_t)f2u(v * 255.0f + 0.5f); }
                }
            }
        }
    }
}
