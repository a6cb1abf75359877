/*
 * oracle/aaa.c -- TEST INFRASTRUCTURE ONLY (never linked into libggcuda.so).
 *
 * CPU restatement of gg's own CPU rasteriser, the pixel oracle of SURVEY.md section 8 row a15: what
 * gg.SoftwareRenderer.Fill does to a pixmap (software.go:485-587) --
 *   EdgeBuilder            internal/raster/edge_builder.go   (lines, canvas clip, native or flattened curves)
 *   LineEdge / QuadraticEdge / CubicEdge   internal/raster/curve_edge.go (FDot6 / FDot16 set-up, forward differencing)
 *   ChopQuad/CubicAtYExtrema               internal/raster/path_geometry.go
 *   CurveAwareAET                          internal/raster/curve_aet.go
 *   AnalyticFiller.Fill    internal/raster/analytic_filler.go:159-1930 (gg's port of Skia's SkScan_AAAPath walker:
 *                          sub-strips, incremental edge X, trapezoid rows, 8-bit additive coverage)
 *   AlphaRuns              internal/raster/alpha_runs.go
 *   source-over of the runs onto the premultiplied RGBA8 pixmap, float64, TRUNCATED to 8 bits per draw
 *                          (software.go:953-1026, pixmap.go:200-228).
 * Integer code: the operation order, shifts and truncations are the reference's (int32 wrap-around like Go: -fwrapv).
 * Pinned (tests/test_cpu_aaa.py) to the goldens the reference's own tests hold for this code: diff == 0 on
 * skia-aaa-{polygon,float-rect-aa,star-aa}-white.png (analytic_filler_golden_test.go:466-560), on
 * multicontour-fill-20x20.png (flattened curves) and multicontour-curve-20x20.png (native quadratic edges,
 * multicontour_golden_test.go:16-131).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ fixed.go */
typedef int32_t fdot6;
typedef int32_t fdot16;
#define FDOT6_ONE 64
#define FDOT6_HALF 32
#define FDOT6_SHIFT 6
#define FDOT16_ONE (1 << 16)
#define FDOT16_SHIFT 16
#define SK_FIXED1 (1 << 16)
#define SK_FIXED_HALF (1 << 15)
#define MAX_S32 0x7FFFFFFF

static int32_t left_shift(int32_t v, int shift) { return shift < 0 ? v >> (-shift) : (int32_t)((uint32_t)v << shift); }   /* fixed.go:253 */
static int32_t saturate_i32(int64_t v) { return v > MAX_S32 ? MAX_S32 : (v < -(int64_t)0x80000000 ? (int32_t)0x80000000 : (int32_t)v); }
static int32_t abs_i32(int32_t v) { return v < 0 ? -v : v; }
static int32_t fdot6_round(fdot6 v) { return (v + FDOT6_HALF) >> FDOT6_SHIFT; }
static fdot16 fdot6_to_fdot16(fdot6 v) {   /* fixed.go:112-122, saturating */
    int64_t r = (int64_t)v << 10;
    if (r > MAX_S32) return MAX_S32;
    if (r < -(int64_t)MAX_S32) return -MAX_S32;
    return (fdot16)r;
}
static fdot16 fdot16_div(int32_t numer, int32_t denom) {   /* fixed.go:213-222 */
    if (denom == 0) return numer >= 0 ? MAX_S32 : -MAX_S32;
    return saturate_i32(((int64_t)numer << 16) / (int64_t)denom);
}
static fdot16 fdot6_div(fdot6 a, fdot6 b) {   /* fixed.go:127-144 */
    if (b == 0) return a >= 0 ? MAX_S32 : -MAX_S32;
    if (a == (int32_t)(int16_t)a) return left_shift(a, 16) / b;
    return fdot16_div(a, b);
}
static fdot16 fdot16_mul(fdot16 a, fdot16 b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 16); }
static int32_t fdot16_round_to_int(fdot16 v) { return (v + (1 << 15)) >> 16; }
static fdot16 fdot6_to_fixed_div2(fdot6 v) { return left_shift(v, 9); }
static int32_t f2i(float f) { return (int32_t)f; }   /* Go int32(float32): truncation */

/* ------------------------------------------------------------------ curve_edge.go */
enum { MAX_COEFF_SHIFT = 6, DEFAULT_ACCURACY = 2 };
typedef struct {
    fdot16 x, dx;
    int32_t first_y, last_y;
    fdot16 upper_y, lower_y;
    int32_t upper_x, pixel_dx, pixel_dy;
    int8_t winding;
} line_edge;
typedef struct {
    int32_t top_y, bottom_y;
    line_edge line;
    int8_t curve_count; uint8_t curve_shift;
    fdot16 qx, qy, qdx, qdy, qddx, qddy, qlast_x, qlast_y, snapped_x, snapped_y;
} quad_edge;
typedef struct {
    int32_t top_y, bottom_y;
    line_edge line;
    int8_t curve_count; uint8_t curve_shift, dshift;
    fdot16 cx, cy, cdx, cdy, cddx, cddy, cdddx, cdddy, clast_x, clast_y, snapped_y;
} cubic_edge;

static fdot16 snap_y(fdot16 y) {   /* curve_edge.go:245-250 */
    const int32_t half = 1 << (16 - 2 - 1);
    const int32_t mask = ~((1 << (16 - 2)) - 1);
    return (y + half) & mask;
}
static fdot16 sk_fixed_round_to_fixed(fdot16 v) { return (v + 0x8000) & ~(fdot16)0xFFFF; }
static fdot16 min_fixed(fdot16 a, fdot16 b) { return a < b ? a : b; }
static fdot6 compute_dy(int32_t top, fdot6 y0) { return left_shift(top, FDOT6_SHIFT) + FDOT6_HALF - y0; }   /* :1019 */

/* curve_edge.go:138-243 NewLineEdge */
static int new_line_edge(float p0x, float p0y, float p1x, float p1y, int shift, line_edge *e) {
    float scale = (float)((int32_t)1 << (shift + FDOT6_SHIFT));
    int32_t x0 = f2i(p0x * scale), y0 = f2i(p0y * scale), x1 = f2i(p1x * scale), y1 = f2i(p1y * scale);
    const float mult = 4.0f;   /* 1 << kDefaultAccuracy */
    int32_t skx0 = f2i(p0x * mult * 64.0f), sky0 = f2i(p0y * mult * 64.0f), skx1 = f2i(p1x * mult * 64.0f), sky1 = f2i(p1y * mult * 64.0f);
    int32_t px_x0 = left_shift(skx0, 10 - 2), px_y0 = snap_y(left_shift(sky0, 10 - 2));
    int32_t px_x1 = left_shift(skx1, 10 - 2), px_y1 = snap_y(left_shift(sky1, 10 - 2));
    int8_t winding = 1;
    if (y0 > y1) {
        int32_t t;
        t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t;
        t = px_x0; px_x0 = px_x1; px_x1 = t; t = px_y0; px_y0 = px_y1; px_y1 = t;
        winding = -1;
    }
    int32_t top = fdot6_round(y0), bottom = fdot6_round(y1);
    if (top == bottom) return 0;
    fdot16 slope = fdot6_div(x1 - x0, y1 - y0);
    fdot6 dy = compute_dy(top, y0);
    int32_t px_dx = (px_x1 - px_x0) >> 10, px_dy = (px_y1 - px_y0) >> 10;
    int32_t pixel_dx = px_dy == 0 ? 0 : fdot6_div(px_dx, px_dy);
    int32_t pixel_dy;
    if (px_dx == 0 || pixel_dx == 0) pixel_dy = MAX_S32;
    else {
        pixel_dy = fdot6_div(abs_i32(px_dy), abs_i32(px_dx));
        if (pixel_dy < 0) pixel_dy = MAX_S32;
    }
    e->x = fdot6_to_fdot16(x0 + fdot16_mul(slope, dy));
    e->dx = slope; e->first_y = top; e->last_y = bottom - 1;
    e->upper_y = px_y0; e->lower_y = px_y1; e->upper_x = px_x0; e->pixel_dx = pixel_dx; e->pixel_dy = pixel_dy; e->winding = winding;
    return 1;
}

/* curve_edge.go:283-353 LineEdge.updateLine (forward-differenced curve segments) */
static int update_line(line_edge *e, fdot16 x0, fdot16 y0, fdot16 x1, fdot16 y1, fdot16 slope) {
    if (y0 > y1) { fdot16 t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; e->winding = (int8_t)-e->winding; }
    int32_t dy = (y1 - y0) >> 10;
    if (dy == 0) return 0;
    int32_t dx = (x1 - x0) >> 10;
    e->x = x0; e->dx = slope; e->upper_x = x0; e->upper_y = y0; e->lower_y = y1;
    int32_t abs_slope6 = abs_i32(slope >> 10);
    if (dx == 0 || slope == 0) e->pixel_dy = MAX_S32;
    else if (abs_slope6 > 0) {
        e->pixel_dy = fdot6_div(abs_i32(dy), abs_i32(dx));
        if (e->pixel_dy < 0) e->pixel_dy = MAX_S32;
    } else e->pixel_dy = MAX_S32;
    e->pixel_dx = slope;
    int32_t top = fdot16_round_to_int(y0), bottom = fdot16_round_to_int(y1);
    if (top == bottom) bottom = top + 1;
    e->first_y = top; e->last_y = bottom - 1;
    return 1;
}

static int32_t cheap_distance(fdot6 dx, fdot6 dy) { dx = abs_i32(dx); dy = abs_i32(dy); return dx > dy ? dx + (dy >> 1) : dy + (dx >> 1); }
static int diff_to_shift(fdot6 dx, fdot6 dy, int shift_aa) {   /* :1038-1057 */
    int32_t dist = cheap_distance(dx, dy);
    dist = (dist + (1 << (2 + shift_aa))) >> (3 + shift_aa);
    if (dist <= 0) return 0;
    return (32 - __builtin_clz((uint32_t)dist)) >> 1;
}
static fdot6 cubic_delta_from_line(fdot6 a, fdot6 b, fdot6 c, fdot6 d) {   /* :1078-1082 */
    int32_t one_third = ((a * 8 - b * 15 + 6 * c + d) * 19) >> 9;
    int32_t two_third = ((a + 6 * b - c * 15 + d * 8) * 19) >> 9;
    int32_t x = abs_i32(one_third), y = abs_i32(two_third);
    return x > y ? x : y;
}
static int curve_pixel_accuracy(int shift) { return shift < DEFAULT_ACCURACY ? shift : DEFAULT_ACCURACY; }

/* curve_edge.go:566-660 QuadraticEdge.Update */
static int quad_update(quad_edge *q) {
    int count = q->curve_count;
    if (count <= 0) return 0;
    fdot16 oldx = q->qx, oldy = q->qy, dx = q->qdx, dy = q->qdy;
    unsigned shift = q->curve_shift;
    fdot16 newx = 0, newy = 0, nsx = 0, nsy = 0;
    int success = 0;
    for (;;) {
        fdot16 slope;
        count--;
        if (count > 0) {
            newx = oldx + (dx >> shift);
            newy = oldy + (dy >> shift);
            int32_t abs_dy_shifted = abs_i32(dy >> shift);
            if (abs_dy_shifted >= FDOT16_ONE * 2 && ((int64_t)abs_i32(dy) << 6) > (int64_t)abs_i32(dx)) {
                int32_t diff_y = (newy - q->snapped_y) >> 10;
                if (diff_y != 0) slope = fdot6_div((newx - q->snapped_x) >> 10, diff_y); else slope = MAX_S32;
                nsy = min_fixed(q->qlast_y, sk_fixed_round_to_fixed(newy));
                nsx = newx - fdot16_mul(slope, newy - nsy);
            } else {
                nsy = min_fixed(q->qlast_y, snap_y(newy));
                nsx = newx;
                int32_t diff_y = (nsy - q->snapped_y) >> 10;
                if (diff_y != 0) slope = fdot6_div((newx - q->snapped_x) >> 10, diff_y); else slope = MAX_S32;
            }
            dx += q->qddx; dy += q->qddy;
        } else {
            newx = q->qlast_x; newy = q->qlast_y; nsx = newx; nsy = newy;
            int32_t diff_y = (newy - q->snapped_y) >> 10;
            if (diff_y != 0) slope = fdot6_div((newx - q->snapped_x) >> 10, diff_y); else slope = MAX_S32;
        }
        if (slope < MAX_S32) success = update_line(&q->line, q->snapped_x, q->snapped_y, nsx, nsy, slope);
        q->snapped_x = nsx; q->snapped_y = nsy;
        oldx = newx; oldy = newy;
        if (count == 0 || success) break;
    }
    q->qx = newx; q->qy = newy; q->qdx = dx; q->qdy = dy; q->curve_count = (int8_t)count;
    return success;
}

/* curve_edge.go:407-563 NewQuadraticEdge */
static int new_quad_edge(const float p[6], int shift, quad_edge *q) {
    float scale = (float)((int32_t)1 << (shift + FDOT6_SHIFT));
    int32_t x0 = f2i(p[0] * scale), y0 = f2i(p[1] * scale), x1 = f2i(p[2] * scale), y1 = f2i(p[3] * scale), x2 = f2i(p[4] * scale), y2 = f2i(p[5] * scale);
    int8_t winding = 1;
    if (y0 > y2) { int32_t t = x0; x0 = x2; x2 = t; t = y0; y0 = y2; y2 = t; winding = -1; }
    int32_t top = fdot6_round(y0), bottom = fdot6_round(y2);
    if (top == bottom) return 0;
    int32_t dx = (left_shift(x1, 1) - x0 - x2) >> 2, dy = (left_shift(y1, 1) - y0 - y2) >> 2;
    int curve_shift = diff_to_shift(dx, dy, shift);
    if (curve_shift < 0) curve_shift = 0;
    if (curve_shift == 0) curve_shift = 1; else if (curve_shift > MAX_COEFF_SHIFT) curve_shift = MAX_COEFF_SHIFT;
    int8_t curve_count = (int8_t)(1 << curve_shift);
    int coeff_shift = curve_shift - 1;
    fdot16 a = fdot6_to_fixed_div2(x0 - x1 - x1 + x2), b = fdot6_to_fdot16(x1 - x0);
    fdot16 qx = fdot6_to_fdot16(x0), qdx = b + (a >> curve_shift);
    fdot16 qddx = coeff_shift >= 1 ? a >> (coeff_shift - 1) : (fdot16)((uint32_t)a << 1);
    a = fdot6_to_fixed_div2(y0 - y1 - y1 + y2); b = fdot6_to_fdot16(y1 - y0);
    fdot16 qy = fdot6_to_fdot16(y0), qdy = b + (a >> curve_shift);
    fdot16 qddy = coeff_shift >= 1 ? a >> (coeff_shift - 1) : (fdot16)((uint32_t)a << 1);
    fdot16 qlx = fdot6_to_fdot16(x2), qly = fdot6_to_fdot16(y2);
    int stored_shift = coeff_shift < 0 ? 0 : coeff_shift;
    int acc = curve_pixel_accuracy(shift);
    qx >>= acc; qy >>= acc; qdx >>= acc; qdy >>= acc; qddx >>= acc; qddy >>= acc; qlx >>= acc; qly >>= acc;
    qy = snap_y(qy); qly = snap_y(qly);
    memset(q, 0, sizeof *q);
    q->top_y = top; q->bottom_y = bottom;
    q->line.first_y = top; q->line.last_y = bottom - 1; q->line.winding = winding;
    q->curve_count = curve_count; q->curve_shift = (uint8_t)stored_shift;
    q->qx = qx; q->qy = qy; q->qdx = qdx; q->qdy = qdy; q->qddx = qddx; q->qddy = qddy; q->qlast_x = qlx; q->qlast_y = qly;
    q->snapped_x = qx; q->snapped_y = qy;
    return quad_update(q);
}

/* curve_edge.go:918-997 CubicEdge.Update */
static int cubic_update(cubic_edge *c) {
    int count = c->curve_count;
    if (count >= 0) return 0;
    fdot16 oldx = c->cx, oldy = c->cy;
    unsigned ddshift = c->curve_shift, dshift = c->dshift;
    fdot16 newx = 0, newy = 0;
    int success = 0;
    for (;;) {
        count++;
        if (count < 0) {
            newx = oldx + (c->cdx >> dshift);
            c->cdx += c->cddx >> ddshift;
            c->cddx += c->cdddx;
            newy = oldy + (c->cdy >> dshift);
            c->cdy += c->cddy >> ddshift;
            c->cddy += c->cdddy;
        } else { newx = c->clast_x; newy = c->clast_y; }
        if (newy < oldy) newy = oldy;
        fdot16 nsy = snap_y(newy);
        if (c->clast_y < nsy) { nsy = c->clast_y; count = 0; }
        fdot16 slope;
        int32_t dy6 = (nsy - c->snapped_y) >> 10;
        if (dy6 == 0) slope = MAX_S32; else slope = fdot6_div((newx - oldx) >> 10, dy6);
        if (slope < MAX_S32) success = update_line(&c->line, oldx, c->snapped_y, newx, nsy, slope);
        c->snapped_y = nsy;
        oldx = newx; oldy = newy;
        if (count == 0 || success) break;
    }
    c->cx = newx; c->cy = newy; c->curve_count = (int8_t)count;
    return success;
}

/* curve_edge.go:748-916 NewCubicEdge */
static int new_cubic_edge(const float p[8], int shift, cubic_edge *c) {
    float scale = (float)((int32_t)1 << (shift + FDOT6_SHIFT));
    int32_t x0 = f2i(p[0] * scale), y0 = f2i(p[1] * scale), x1 = f2i(p[2] * scale), y1 = f2i(p[3] * scale);
    int32_t x2 = f2i(p[4] * scale), y2 = f2i(p[5] * scale), x3 = f2i(p[6] * scale), y3 = f2i(p[7] * scale);
    int8_t winding = 1;
    if (y0 > y3) {
        int32_t t;
        t = x0; x0 = x3; x3 = t; t = x1; x1 = x2; x2 = t; t = y0; y0 = y3; y3 = t; t = y1; y1 = y2; y2 = t;
        winding = -1;
    }
    int32_t top = fdot6_round(y0), bot = fdot6_round(y3);
    if (top == bot) return 0;
    int32_t dx = cubic_delta_from_line(x0, x1, x2, x3), dy = cubic_delta_from_line(y0, y1, y2, y3);
    int curve_shift = diff_to_shift(dx, dy, 2) + 1;
    if (curve_shift < 1) curve_shift = 1;
    if (curve_shift > MAX_COEFF_SHIFT) curve_shift = MAX_COEFF_SHIFT;
    int up_shift = 6, down_shift = curve_shift + up_shift - 10;
    if (down_shift < 0) { down_shift = 0; up_shift = 10 - curve_shift; }
    int8_t curve_count = (int8_t)left_shift(-1, curve_shift);
    int32_t b = left_shift(3 * (x1 - x0), up_shift), cc = left_shift(3 * (x0 - x1 - x1 + x2), up_shift), d = left_shift(x3 + 3 * (x1 - x2) - x0, up_shift);
    fdot16 cx = fdot6_to_fdot16(x0);
    fdot16 cdx = b + (cc >> curve_shift) + (d >> (2 * curve_shift));
    fdot16 cddx = 2 * cc + ((3 * d) >> (curve_shift - 1));
    fdot16 cdddx = (3 * d) >> (curve_shift - 1);
    b = left_shift(3 * (y1 - y0), up_shift); cc = left_shift(3 * (y0 - y1 - y1 + y2), up_shift); d = left_shift(y3 + 3 * (y1 - y2) - y0, up_shift);
    fdot16 cy = fdot6_to_fdot16(y0);
    fdot16 cdy = b + (cc >> curve_shift) + (d >> (2 * curve_shift));
    fdot16 cddy = 2 * cc + ((3 * d) >> (curve_shift - 1));
    fdot16 cdddy = (3 * d) >> (curve_shift - 1);
    fdot16 clx = fdot6_to_fdot16(x3), cly = fdot6_to_fdot16(y3);
    int acc = curve_pixel_accuracy(shift);
    cx >>= acc; cy >>= acc; cdx >>= acc; cdy >>= acc; cddx >>= acc; cddy >>= acc; cdddx >>= acc; cdddy >>= acc; clx >>= acc; cly >>= acc;
    cy = snap_y(cy); cly = snap_y(cly);
    memset(c, 0, sizeof *c);
    c->top_y = top; c->bottom_y = bot;
    c->line.first_y = top; c->line.last_y = bot - 1; c->line.winding = winding;
    c->curve_count = curve_count; c->curve_shift = (uint8_t)curve_shift; c->dshift = (uint8_t)down_shift;
    c->cx = cx; c->cy = cy; c->cdx = cdx; c->cdy = cdy; c->cddx = cddx; c->cddy = cddy; c->cdddx = cdddx; c->cdddy = cdddy;
    c->clast_x = clx; c->clast_y = cly; c->snapped_y = cy;
    return cubic_update(c);
}

/* ------------------------------------------------------------------ path_geometry.go */
typedef struct { float x, y; } gpt;
static gpt lerp_pt(gpt a, gpt b, float t) { gpt r = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)}; return r; }
static float abs_f(float x) { return x < 0 ? -x : x; }
static float min_f(float a, float b) { return a < b ? a : b; }
static float max_f(float a, float b) { return a > b ? a : b; }
static int is_not_monotonic(float a, float b, float c) { float ab = a - b, bc = b - c; if (ab < 0) bc = -bc; return ab == 0 || bc < 0; }
static float valid_unit_divide(float numer, float denom) {
    if (denom == 0) return 0;
    float t = numer / denom;
    if (t > 0 && t < 1) { if (isnan(t) || isinf(t)) return 0; return t; }
    return 0;
}
static int chop_quad_at_y_extrema(const gpt src[3], gpt dst[5]) {   /* :52-105 */
    float a = src[0].y, b = src[1].y, c = src[2].y;
    if (is_not_monotonic(a, b, c)) {
        float t = valid_unit_divide(a - b, a - 2 * b + c);
        if (t > 0 && t < 1) {
            gpt ab = lerp_pt(src[0], src[1], t), bc = lerp_pt(src[1], src[2], t), abbc = lerp_pt(ab, bc, t);
            dst[0] = src[0]; dst[1] = ab; dst[2] = abbc; dst[3] = bc; dst[4] = src[2];
            float mn = min_f(dst[0].y, dst[2].y), mx = max_f(dst[0].y, dst[2].y);
            if (dst[1].y < mn) dst[1].y = mn; else if (dst[1].y > mx) dst[1].y = mx;
            mn = min_f(dst[2].y, dst[4].y); mx = max_f(dst[2].y, dst[4].y);
            if (dst[3].y < mn) dst[3].y = mn; else if (dst[3].y > mx) dst[3].y = mx;
            return 1;
        }
        if (abs_f(a - b) < abs_f(b - c)) b = a; else b = c;
    }
    dst[0].x = src[0].x; dst[0].y = a; dst[1].x = src[1].x; dst[1].y = b; dst[2].x = src[2].x; dst[2].y = c;
    return 0;
}
static void chop_cubic_at_single(const gpt src[4], float t, gpt dst[10]) {
    gpt ab = lerp_pt(src[0], src[1], t), bc = lerp_pt(src[1], src[2], t), cd = lerp_pt(src[2], src[3], t);
    gpt abbc = lerp_pt(ab, bc, t), bccd = lerp_pt(bc, cd, t), mid = lerp_pt(abbc, bccd, t);
    dst[0] = src[0]; dst[1] = ab; dst[2] = abbc; dst[3] = mid; dst[4] = bccd; dst[5] = cd; dst[6] = src[3];
}
static int find_unit_quad_roots(float a, float b, float c, float roots[2]) {   /* :355-398 */
    const float eps = 1e-7f;
    if (abs_f(a) < eps) {
        if (abs_f(b) < eps) return 0;
        float t = -c / b;
        if (t > 0 && t < 1) { roots[0] = t; return 1; }
        return 0;
    }
    float disc = b * b - 4 * a * c;
    if (disc < 0) return 0;
    float sq = (float)sqrt((double)disc);
    float inv2a = (float)(1.0 / (double)(2 * a));   /* Go: 1.0 / (2*a) with an untyped constant: float32 division */
    inv2a = 1.0f / (2 * a);
    float t1 = (-b - sq) * inv2a, t2 = (-b + sq) * inv2a;
    if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
    int n = 0;
    if (t1 > eps && t1 < 1 - eps) roots[n++] = t1;
    if (t2 > eps && t2 < 1 - eps && abs_f(t2 - t1) > eps) roots[n++] = t2;
    return n;
}
static int chop_cubic_at_y_extrema(const gpt src[4], gpt dst[10]) {   /* :122-186 */
    float a = src[0].y, b = src[1].y, c = src[2].y, d = src[3].y;
    float tv[2];
    int n = find_unit_quad_roots(d - a + 3 * (b - c), 2 * (a - 2 * b + c), b - a, tv);
    if (n == 0) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3]; }
    else {
        chop_cubic_at_single(src, tv[0], dst);
        if (n == 2) {
            float nt = valid_unit_divide(tv[1] - tv[0], 1 - tv[0]);
            if (nt <= 0) { dst[7] = src[3]; dst[8] = src[3]; dst[9] = src[3]; }
            else {
                gpt rem[4] = {dst[3], dst[4], dst[5], dst[6]}, sh[10];
                chop_cubic_at_single(rem, nt, sh);
                dst[4] = sh[1]; dst[5] = sh[2]; dst[6] = sh[3]; dst[7] = sh[4]; dst[8] = sh[5]; dst[9] = sh[6];
            }
        }
    }
    for (int k = 0; k <= n; k++) {
        int s = 3 * k;
        float mn = min_f(dst[s].y, dst[s + 3].y), mx = max_f(dst[s].y, dst[s + 3].y);
        if (dst[s + 1].y < mn) dst[s + 1].y = mn; else if (dst[s + 1].y > mx) dst[s + 1].y = mx;
        if (dst[s + 2].y < mn) dst[s + 2].y = mn; else if (dst[s + 2].y > mx) dst[s + 2].y = mx;
    }
    return n;
}

/* ------------------------------------------------------------------ edge_builder.go */
enum { EDGE_LINE = 0, EDGE_QUAD = 1, EDGE_CUBIC = 2 };
typedef struct { int type; int ix; int32_t top_y; } edge_ref;
typedef struct {
    line_edge *lines; int n_lines, cap_lines;
    quad_edge *quads; int n_quads, cap_quads;
    cubic_edge *cubics; int n_cubics, cap_cubics;
    int aa_shift, flatten;
    float flatten_tol;
    int has_clip; float clip[4];   /* MinX, MinY, MaxX, MaxY */
    int b_empty; float bminx, bminy, bmaxx, bmaxy;
} edge_builder;

static void eb_union(edge_builder *eb, float x, float y) {   /* :167-192 */
    if (eb->b_empty) { eb->bminx = eb->bmaxx = x; eb->bminy = eb->bmaxy = y; eb->b_empty = 0; return; }
    if (x < eb->bminx) eb->bminx = x;
    if (x > eb->bmaxx) eb->bmaxx = x;
    if (y < eb->bminy) eb->bminy = y;
    if (y > eb->bmaxy) eb->bmaxy = y;
}
enum { COMBINE_NO, COMBINE_PARTIAL, COMBINE_TOTAL };
static int combine_vertical(const line_edge *edge, line_edge *last) {   /* :1166-1222 */
    if (last->dx != 0 || edge->x != last->x) return COMBINE_NO;
    if (edge->winding == last->winding) {
        if (edge->last_y + 1 == last->first_y) { last->first_y = edge->first_y; last->upper_y = edge->upper_y; return COMBINE_PARTIAL; }
        if (edge->first_y == last->last_y + 1) { last->last_y = edge->last_y; last->lower_y = edge->lower_y; return COMBINE_PARTIAL; }
        return COMBINE_NO;
    }
    if (edge->first_y == last->first_y) {
        if (edge->last_y == last->last_y) return COMBINE_TOTAL;
        if (edge->last_y < last->last_y) { last->first_y = edge->last_y + 1; last->upper_y = edge->lower_y; return COMBINE_PARTIAL; }
        last->first_y = last->last_y + 1; last->upper_y = last->lower_y; last->last_y = edge->last_y; last->lower_y = edge->lower_y; last->winding = edge->winding;
        return COMBINE_PARTIAL;
    }
    if (edge->last_y == last->last_y) {
        if (edge->first_y > last->first_y) { last->last_y = edge->first_y - 1; last->lower_y = edge->upper_y; return COMBINE_PARTIAL; }
        last->last_y = last->first_y - 1; last->lower_y = last->upper_y; last->first_y = edge->first_y; last->upper_y = edge->upper_y; last->winding = edge->winding;
        return COMBINE_PARTIAL;
    }
    return COMBINE_NO;
}
static void eb_add_line_unclipped(edge_builder *eb, float x0, float y0, float x1, float y1) {   /* :393-451 */
    eb_union(eb, x0, y0); eb_union(eb, x1, y1);
    line_edge e;
    if (!new_line_edge(x0, y0, x1, y1, eb->aa_shift, &e)) return;
    if (e.dx == 0 && eb->n_lines > 0) {
        int r = combine_vertical(&e, &eb->lines[eb->n_lines - 1]);
        if (r == COMBINE_TOTAL) { eb->n_lines--; return; }
        if (r == COMBINE_PARTIAL) return;
    }
    if (eb->n_lines == eb->cap_lines) { eb->cap_lines = eb->cap_lines ? eb->cap_lines * 2 : 64; eb->lines = (line_edge *)realloc(eb->lines, sizeof(line_edge) * (size_t)eb->cap_lines); }
    eb->lines[eb->n_lines++] = e;
}
static void eb_emit_segment(edge_builder *eb, float sx0, float sy0, float sx1, float sy1, float left, float right, int preserve) {   /* :604-626 */
    float mid = (sx0 + sx1) * 0.5f, ex0, ey0, ex1, ey1;
    if (mid < left) { ex0 = left; ey0 = sy0; ex1 = left; ey1 = sy1; }
    else if (mid > right) { ex0 = right; ey0 = sy0; ex1 = right; ey1 = sy1; }
    else { ex0 = sx0; ey0 = sy0; ex1 = sx1; ey1 = sy1; }
    if (!preserve) { float t = ex0; ex0 = ex1; ex1 = t; t = ey0; ey0 = ey1; ey1 = t; }
    eb_add_line_unclipped(eb, ex0, ey0, ex1, ey1);
}
static void eb_clip_line_x(edge_builder *eb, float x0, float y0, float x1, float y1) {   /* :501-600 */
    float left = eb->clip[0], right = eb->clip[2], tx, ty, bx, by;
    if (y0 <= y1) { tx = x0; ty = y0; bx = x1; by = y1; } else { tx = x1; ty = y1; bx = x0; by = y0; }
    if (tx >= left && tx <= right && bx >= left && bx <= right) { eb_add_line_unclipped(eb, x0, y0, x1, y1); return; }
    if (tx <= left && bx <= left) { eb_add_line_unclipped(eb, left, y0, left, y1); return; }
    if (tx >= right && bx >= right) { eb_add_line_unclipped(eb, right, y0, right, y1); return; }
    int preserve = y0 <= y1;
    float st[2], sx[2]; int cnt = 0;
    float dx = bx - tx;
    if (dx != 0) {
        float bnd[2] = {left, right};
        for (int k = 0; k < 2; k++) { float t = (bnd[k] - tx) / dx; if (t > 0 && t < 1) { st[cnt] = t; sx[cnt] = bnd[k]; cnt++; } }
        if (cnt == 2 && st[0] > st[1]) { float t = st[0]; st[0] = st[1]; st[1] = t; t = sx[0]; sx[0] = sx[1]; sx[1] = t; }
    }
    float px = tx, py = ty;
    for (int k = 0; k < cnt; k++) {
        float yat = ty + st[k] * (by - ty);
        eb_emit_segment(eb, px, py, sx[k], yat, left, right, preserve);
        px = sx[k]; py = yat;
    }
    eb_emit_segment(eb, px, py, bx, by, left, right, preserve);
}
static void eb_add_line(edge_builder *eb, float x0, float y0, float x1, float y1) {   /* :382-390, :454-498 */
    if (!eb->has_clip) { eb_add_line_unclipped(eb, x0, y0, x1, y1); return; }
    float miny = eb->clip[1], maxy = eb->clip[3];
    int down = 1;
    if (y0 > y1) { float t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; down = 0; }
    if (y1 <= miny || y0 >= maxy) return;
    if (y0 < miny) { float t = (miny - y0) / (y1 - y0); x0 += t * (x1 - x0); y0 = miny; }
    if (y1 > maxy) { float t = (maxy - y0) / (y1 - y0); x1 = x0 + t * (x1 - x0); y1 = maxy; }
    if (y0 >= y1) return;
    if (!down) { float t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
    eb_clip_line_x(eb, x0, y0, x1, y1);
}
static float eb_tol(const edge_builder *eb) { return eb->flatten_tol > 0 ? eb->flatten_tol : 0.1f; }
static void eb_flatten_quad(edge_builder *eb, float x0, float y0, float cx, float cy, float x1, float y1, float tol, int depth) {   /* :887-927 */
    if (depth > 10) { eb_add_line(eb, x0, y0, x1, y1); return; }
    float dx = x1 - x0, dy = y1 - y0, dcx = cx - x0, dcy = cy - y0;
    float cross = dcx * dy - dcy * dx, len_sq = dx * dx + dy * dy;
    if (len_sq < 1e-6f || cross * cross / len_sq < tol * tol) { eb_add_line(eb, x0, y0, x1, y1); return; }
    float q0x = (x0 + cx) * 0.5f, q0y = (y0 + cy) * 0.5f, q1x = (cx + x1) * 0.5f, q1y = (cy + y1) * 0.5f;
    float rx = (q0x + q1x) * 0.5f, ry = (q0y + q1y) * 0.5f;
    eb_flatten_quad(eb, x0, y0, q0x, q0y, rx, ry, tol, depth + 1);
    eb_flatten_quad(eb, rx, ry, q1x, q1y, x1, y1, tol, depth + 1);
}
static void eb_flatten_cubic(edge_builder *eb, float x0, float y0, float c1x, float c1y, float c2x, float c2y, float x1, float y1, float tol, int depth) {   /* :1083-1148 */
    if (depth > 10) { eb_add_line(eb, x0, y0, x1, y1); return; }
    float dx = x1 - x0, dy = y1 - y0, len_sq = dx * dx + dy * dy;
    if (len_sq < 1e-6f) { eb_add_line(eb, x0, y0, x1, y1); return; }
    float cross1 = (c1x - x0) * dy - (c1y - y0) * dx, cross2 = (c2x - x0) * dy - (c2y - y0) * dx;
    float mc = cross1;
    if (cross1 < 0) mc = -cross1;
    if (cross2 > mc) mc = cross2;
    if (cross2 < -mc) mc = -cross2;
    if (mc * mc / len_sq < tol * tol) { eb_add_line(eb, x0, y0, x1, y1); return; }
    float m01x = (x0 + c1x) * 0.5f, m01y = (y0 + c1y) * 0.5f, m12x = (c1x + c2x) * 0.5f, m12y = (c1y + c2y) * 0.5f, m23x = (c2x + x1) * 0.5f, m23y = (c2y + y1) * 0.5f;
    float m012x = (m01x + m12x) * 0.5f, m012y = (m01y + m12y) * 0.5f, m123x = (m12x + m23x) * 0.5f, m123y = (m12y + m23y) * 0.5f;
    float mx = (m012x + m123x) * 0.5f, my = (m012y + m123y) * 0.5f;
    eb_flatten_cubic(eb, x0, y0, m01x, m01y, m012x, m012y, mx, my, tol, depth + 1);
    eb_flatten_cubic(eb, mx, my, m123x, m123y, m23x, m23y, x1, y1, tol, depth + 1);
}
static int eb_inside_clip(const edge_builder *eb, const float *c, int n) {
    for (int i = 0; i < n; i += 2) if (c[i] < eb->clip[0] || c[i] > eb->clip[2] || c[i + 1] < eb->clip[1] || c[i + 1] > eb->clip[3]) return 0;
    return 1;
}
static void eb_add_quad(edge_builder *eb, float x0, float y0, float cx, float cy, float x1, float y1) {   /* :740-825 */
    if (eb->flatten) { eb_flatten_quad(eb, x0, y0, cx, cy, x1, y1, eb_tol(eb), 0); return; }
    float devx = cx - (x0 + x1) * 0.5f, devy = cy - (y0 + y1) * 0.5f;
    float dev_sq = devx * devx + devy * devy;
    const float max_dev_sq = (float)(0.1 * 0.1);
    if (dev_sq > max_dev_sq) {
        float mx01 = (x0 + cx) * 0.5f, my01 = (y0 + cy) * 0.5f, mx12 = (cx + x1) * 0.5f, my12 = (cy + y1) * 0.5f;
        float mx = (mx01 + mx12) * 0.5f, my = (my01 + my12) * 0.5f;
        eb_add_quad(eb, x0, y0, mx01, my01, mx, my);
        eb_add_quad(eb, mx, my, mx12, my12, x1, y1);
        return;
    }
    float pts[6] = {x0, y0, cx, cy, x1, y1};
    if (eb->has_clip && !eb_inside_clip(eb, pts, 6)) { eb_flatten_quad(eb, x0, y0, cx, cy, x1, y1, eb_tol(eb), 0); return; }
    gpt src[3] = {{x0, y0}, {cx, cy}, {x1, y1}}, dst[5];
    int n = chop_quad_at_y_extrema(src, dst);
    eb_union(eb, dst[0].x, dst[0].y); eb_union(eb, dst[2].x, dst[2].y);
    if (n > 0) eb_union(eb, dst[4].x, dst[4].y);
    for (int i = 0; i <= n; i++) {
        float p[6] = {dst[i * 2].x, dst[i * 2].y, dst[i * 2 + 1].x, dst[i * 2 + 1].y, dst[i * 2 + 2].x, dst[i * 2 + 2].y};
        quad_edge q;
        if (new_quad_edge(p, eb->aa_shift, &q)) {
            if (eb->n_quads == eb->cap_quads) { eb->cap_quads = eb->cap_quads ? eb->cap_quads * 2 : 16; eb->quads = (quad_edge *)realloc(eb->quads, sizeof(quad_edge) * (size_t)eb->cap_quads); }
            eb->quads[eb->n_quads++] = q;
        }
    }
}
static void eb_add_cubic(edge_builder *eb, float x0, float y0, float c1x, float c1y, float c2x, float c2y, float x1, float y1) {   /* :931-1019 */
    if (eb->flatten) { eb_flatten_cubic(eb, x0, y0, c1x, c1y, c2x, c2y, x1, y1, eb_tol(eb), 0); return; }
    float d1x = c1x - (x0 * 2 + x1) / 3, d1y = c1y - (y0 * 2 + y1) / 3, d2x = c2x - (x0 + x1 * 2) / 3, d2y = c2y - (y0 + y1 * 2) / 3;
    float dev1 = d1x * d1x + d1y * d1y, dev2 = d2x * d2x + d2y * d2y;
    float dev_sq = dev1;
    if (dev2 > dev_sq) dev_sq = dev2;
    const float max_dev_sq = (float)(0.1 * 0.1);
    if (dev_sq > max_dev_sq) {
        float m01x = (x0 + c1x) * 0.5f, m01y = (y0 + c1y) * 0.5f, m12x = (c1x + c2x) * 0.5f, m12y = (c1y + c2y) * 0.5f, m23x = (c2x + x1) * 0.5f, m23y = (c2y + y1) * 0.5f;
        float m012x = (m01x + m12x) * 0.5f, m012y = (m01y + m12y) * 0.5f, m123x = (m12x + m23x) * 0.5f, m123y = (m12y + m23y) * 0.5f;
        float mx = (m012x + m123x) * 0.5f, my = (m012y + m123y) * 0.5f;
        eb_add_cubic(eb, x0, y0, m01x, m01y, m012x, m012y, mx, my);
        eb_add_cubic(eb, mx, my, m123x, m123y, m23x, m23y, x1, y1);
        return;
    }
    float pts[8] = {x0, y0, c1x, c1y, c2x, c2y, x1, y1};
    if (eb->has_clip && !eb_inside_clip(eb, pts, 8)) { eb_flatten_cubic(eb, x0, y0, c1x, c1y, c2x, c2y, x1, y1, eb_tol(eb), 0); return; }
    gpt src[4] = {{x0, y0}, {c1x, c1y}, {c2x, c2y}, {x1, y1}}, dst[10];
    int n = chop_cubic_at_y_extrema(src, dst);
    eb_union(eb, dst[0].x, dst[0].y); eb_union(eb, dst[3].x, dst[3].y);
    if (n >= 1) eb_union(eb, dst[6].x, dst[6].y);
    if (n >= 2) eb_union(eb, dst[9].x, dst[9].y);
    for (int i = 0; i <= n; i++) {
        float p[8] = {dst[i * 3].x, dst[i * 3].y, dst[i * 3 + 1].x, dst[i * 3 + 1].y, dst[i * 3 + 2].x, dst[i * 3 + 2].y, dst[i * 3 + 3].x, dst[i * 3 + 3].y};
        cubic_edge c;
        if (new_cubic_edge(p, eb->aa_shift, &c)) {
            if (eb->n_cubics == eb->cap_cubics) { eb->cap_cubics = eb->cap_cubics ? eb->cap_cubics * 2 : 16; eb->cubics = (cubic_edge *)realloc(eb->cubics, sizeof(cubic_edge) * (size_t)eb->cap_cubics); }
            eb->cubics[eb->n_cubics++] = c;
        }
    }
}
/* edge_builder.go:318-378 BuildFromPathF64 (verbs: MoveTo 0, LineTo 1, QuadTo 2, CubicTo 3, Close 4) */
static void eb_build(edge_builder *eb, const uint8_t *verbs, uint32_t n_verbs, const double *co) {
    float cx = 0, cy = 0, sx = 0, sy = 0;
    size_t k = 0;
    for (uint32_t i = 0; i < n_verbs; i++) {
        switch (verbs[i]) {
        case 0:
            if (cx != sx || cy != sy) eb_add_line(eb, cx, cy, sx, sy);
            cx = (float)co[k]; cy = (float)co[k + 1]; sx = cx; sy = cy; k += 2; break;
        case 1: { float nx = (float)co[k], ny = (float)co[k + 1]; eb_add_line(eb, cx, cy, nx, ny); cx = nx; cy = ny; k += 2; } break;
        case 2: { float qx = (float)co[k], qy = (float)co[k + 1], x = (float)co[k + 2], y = (float)co[k + 3]; eb_add_quad(eb, cx, cy, qx, qy, x, y); cx = x; cy = y; k += 4; } break;
        case 3: { float ax = (float)co[k], ay = (float)co[k + 1], bx = (float)co[k + 2], by = (float)co[k + 3], x = (float)co[k + 4], y = (float)co[k + 5];
                  eb_add_cubic(eb, cx, cy, ax, ay, bx, by, x, y); cx = x; cy = y; k += 6; } break;
        case 4:
            if (cx != sx || cy != sy) eb_add_line(eb, cx, cy, sx, sy);
            cx = sx; cy = sy; break;
        default: break;
        }
    }
    if (cx != sx || cy != sy) eb_add_line(eb, cx, cy, sx, sy);
}

/* ------------------------------------------------------------------ analytic_filler.go */
typedef struct { int type; line_edge line; quad_edge quad; cubic_edge cubic; } edge_var;   /* CurveEdgeVariant, by value */
static line_edge *ev_line(edge_var *e) { return e->type == EDGE_LINE ? &e->line : (e->type == EDGE_QUAD ? &e->quad.line : &e->cubic.line); }
static int32_t ev_top(edge_var *e) { return e->type == EDGE_LINE ? e->line.first_y : (e->type == EDGE_QUAD ? e->quad.top_y : e->cubic.top_y); }
static int32_t ev_bottom(edge_var *e) { return e->type == EDGE_LINE ? e->line.last_y + 1 : (e->type == EDGE_QUAD ? e->quad.bottom_y : e->cubic.bottom_y); }

typedef struct { int32_t fx, fdx, fupper_x, fupper_y, flower_y, fdy; int8_t winding; int valid; } edge_y_state;   /* :1174-1183 */
typedef struct { int valid; int32_t top_x, bot_x, dy; uint8_t full_alpha; int8_t winding; } edge_line_state;          /* :1188-1195 */
typedef struct { int idx; int32_t upper_y; } deferred_edge;

typedef struct {
    int width, height;
    /* the AET holds indices into edge_buf: the Go code copies CurveEdgeVariant structs whose pointers alias the builder's
     * edges, so every copy sees the same edge state -- an index has the same meaning */
    int *aet; int n_aet, cap_aet;
    uint8_t *coverage;
    int edge_idx;
    edge_var *edge_buf; int n_edges;
    edge_line_state *resolved; int n_resolved, cap_resolved;
    edge_y_state *states;
    int32_t *strip_y; int n_strip, cap_strip;
    deferred_edge *deferred; int n_deferred, cap_deferred;
    int32_t next_next_y;
    int32_t aa_scale;
    /* AlphaRuns */
    uint16_t *runs; uint8_t *alpha; int run_offset;
} filler;

static int32_t sk_fixed_mul(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 16); }
static int32_t sk_floor_to_int(int32_t v) { return v >> 16; }
static int32_t sk_ceil_to_int(int32_t v) { return (v + SK_FIXED1 - 1) >> 16; }
static int32_t sk_floor_to_fixed(int32_t v) { return v & ~(SK_FIXED1 - 1); }
static int32_t sk_ceil_to_fixed(int32_t v) { return sk_floor_to_fixed(v + SK_FIXED1 - 1); }
static int32_t sk32_sat_add(int32_t a, int32_t b) { int64_t s = (int64_t)a + b; return s > MAX_S32 ? MAX_S32 : (s < -(int64_t)0x80000000 ? (int32_t)0x80000000 : (int32_t)s); }
static uint8_t sat_sub8(uint8_t a, uint8_t b) { return b >= a ? 0 : (uint8_t)(a - b); }
static int32_t clamp_alpha32(int32_t v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
static uint8_t fixed_to_alpha(int32_t f) {   /* :1851-1866 */
    if (f <= 0) return 0;
    if (f >= SK_FIXED1) return 255;
    int64_t v = ((int64_t)255 * f + SK_FIXED_HALF) >> 16;
    return v > 255 ? 255 : (v < 0 ? 0 : (uint8_t)v);
}
static int32_t compute_edge_dy(int32_t slope) {   /* :1317-1332 */
    int32_t a = slope < 0 ? -slope : slope;
    int32_t a6 = a >> 10;
    if (a6 == 0) return MAX_S32;
    int32_t fdy = fdot6_div(FDOT6_ONE, a6);
    return fdy < 0 ? MAX_S32 : fdy;
}
static int compute_y_shift(int32_t d) { return d == (SK_FIXED1 >> 2) ? 2 : (d == (SK_FIXED1 >> 1) ? 1 : (d == SK_FIXED1 ? 0 : -1)); }
static uint8_t trapezoid_to_alpha_scaled(int32_t l1, int32_t l2, uint8_t full) {   /* :1660-1692 */
    if (l1 < 0) l1 = 0;
    if (l2 < 0) l2 = 0;
    int64_t area = ((int64_t)l1 + (int64_t)l2) / 2;
    if (full == 255) { int64_t v = area >> 8; return v > 255 ? 255 : (v < 0 ? 0 : (uint8_t)v); }
    int32_t a = (int32_t)(area >> 8);
    if (a > 255) a = 255;
    if (a < 0) a = 0;
    return (uint8_t)(((uint16_t)a * (uint16_t)full) >> 8);
}
static uint8_t trapezoid_to_alpha(int32_t l1, int32_t l2) {   /* :1697-1713 */
    if (l1 < 0) l1 = 0;
    if (l2 < 0) l2 = 0;
    int32_t area = (l1 + l2) / 2, r = area >> 8;
    return r > 255 ? 255 : (r < 0 ? 0 : (uint8_t)r);
}
static uint8_t partial_triangle_to_alpha(int32_t a, int32_t b) {   /* :1718-1738 */
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    if (a > SK_FIXED1) a = SK_FIXED1;
    int32_t a11 = a >> 11, b11 = b >> 11;
    int32_t area = a11 * a11 * b11;
    int32_t r = (area >> 8) & 0xFF;
    return r < 0 ? 0 : (uint8_t)r;
}
static uint8_t get_partial_alpha8(uint8_t a, uint8_t full) { return (uint8_t)(((uint16_t)a * (uint16_t)full) >> 8); }
static void compute_alpha_above_line(uint8_t *al, int len, int32_t l, int32_t r, int32_t dy, uint8_t full) {   /* :1753-1780 */
    if (l < 0) l = 0;
    if (l > r) { int32_t t = l; l = r; r = t; }
    int32_t R = sk_ceil_to_int(r);
    if (R <= 0 || R > len) return;
    if (R == 1) { al[0] = get_partial_alpha8((uint8_t)clamp_alpha32(((R << 17) - l - r) >> 9), full); return; }
    int32_t first = SK_FIXED1 - l, last = r - ((R - 1) << 16);
    int32_t first_h = sk_fixed_mul(first, dy);
    al[0] = (uint8_t)clamp_alpha32(sk_fixed_mul(first, first_h) >> 9);
    int32_t a16 = sk32_sat_add(first_h, dy >> 1);
    for (int32_t i = 1; i < R - 1; i++) { al[i] = (uint8_t)clamp_alpha32(a16 >> 8); a16 = sk32_sat_add(a16, dy); }
    al[R - 1] = sat_sub8(full, partial_triangle_to_alpha(last, dy));
}
static void compute_alpha_below_line(uint8_t *al, int len, int32_t l, int32_t r, int32_t dy, uint8_t full) {   /* :1784-1813 */
    if (l < 0) l = 0;
    if (l > r) { int32_t t = l; l = r; r = t; }
    int32_t R = sk_ceil_to_int(r);
    if (R <= 0 || R > len) return;
    if (R == 1) { al[0] = get_partial_alpha8(trapezoid_to_alpha(l, r), full); return; }
    int32_t last = r - ((R - 1) << 16);
    int32_t last_h = sk_fixed_mul(last, dy);
    al[R - 1] = (uint8_t)clamp_alpha32(sk_fixed_mul(last, last_h) >> 9);
    int32_t a16 = sk32_sat_add(last_h, dy >> 1);
    for (int32_t i = R - 2; i > 0; i--) { al[i] = (uint8_t)clamp_alpha32(a16 >> 8); a16 = sk32_sat_add(a16, dy); }
    int32_t first = SK_FIXED1 - l;
    al[0] = sat_sub8(full, partial_triangle_to_alpha(first, dy));
}
static int32_t approximate_intersection(int32_t l1, int32_t r1, int32_t l2, int32_t r2) {   /* :1817-1833 */
    if (l1 > r1) { int32_t t = l1; l1 = r1; r1 = t; }
    if (l2 > r2) { int32_t t = l2; l2 = r2; r2 = t; }
    int32_t ml = l1 > l2 ? l1 : l2, mr = r1 < r2 ? r1 : r2;
    return (ml + mr) / 2;
}
static void safe_add_alpha(filler *af, int32_t x, uint8_t a) {   /* :1591-1600 */
    if (x < 0 || x >= af->width || a == 0) return;
    uint16_t s = (uint16_t)(af->coverage[x] + a);
    af->coverage[x] = s > 255 ? 255 : (uint8_t)s;
}
static void blit_aaa_trapezoid_row(filler *af, int32_t ul, int32_t ur, int32_t ll, int32_t lr, int32_t ldy, int32_t rdy, uint8_t full) {   /* :1502-1589 */
    int32_t base_x = sk_floor_to_int(ul), end_x = sk_ceil_to_int(lr), length = end_x - base_x;
    if (length <= 0) return;
    if (length == 1) { safe_add_alpha(af, base_x, trapezoid_to_alpha_scaled(ur - ul, lr - ll, full)); return; }
    uint8_t *alphas = (uint8_t *)malloc((size_t)length), *tmp = (uint8_t *)calloc((size_t)length, 1);
    memset(alphas, full, (size_t)length);
    int32_t uL = sk_floor_to_int(ul), lL = sk_ceil_to_int(ll);
    if (uL + 2 == lL) {
        int32_t first = (uL << 16) + SK_FIXED1 - ul, second = ll - ul - first;
        uint8_t a1 = sat_sub8(full, partial_triangle_to_alpha(first, ldy)), a2 = partial_triangle_to_alpha(second, ldy);
        alphas[0] = sat_sub8(alphas[0], a1);
        alphas[1] = sat_sub8(alphas[1], a2);
    } else {
        compute_alpha_below_line(tmp + (uL - base_x), length - (uL - base_x), ul - (uL << 16), ll - (uL << 16), ldy, full);
        for (int32_t i = uL; i < lL && i - base_x < length; i++) { int32_t k = i - base_x; if (k >= 0 && k < length) alphas[k] = sat_sub8(alphas[k], tmp[k]); }
    }
    int32_t uR = sk_floor_to_int(ur), lR = sk_ceil_to_int(lr);
    memset(tmp, 0, (size_t)length);
    if (uR + 2 == lR) {   /* subtractRightExclusion :1562-1589 */
        int32_t first = (uR << 16) + SK_FIXED1 - ur, second = lr - ur - first;
        uint8_t a1 = partial_triangle_to_alpha(first, rdy), a2 = sat_sub8(full, partial_triangle_to_alpha(second, rdy));
        if (length - 2 >= 0) alphas[length - 2] = sat_sub8(alphas[length - 2], a1);
        if (length - 1 >= 0) alphas[length - 1] = sat_sub8(alphas[length - 1], a2);
    } else if (uR - base_x >= 0 && uR - base_x <= length) {
        compute_alpha_above_line(tmp + (uR - base_x), length - (uR - base_x), ur - (uR << 16), lr - (uR << 16), rdy, full);
        for (int32_t i = uR; i < lR && i - base_x < length; i++) { int32_t k = i - base_x; if (k >= 0 && k < length) alphas[k] = sat_sub8(alphas[k], tmp[k]); }
    }
    for (int32_t i = 0; i < length; i++) safe_add_alpha(af, base_x + i, alphas[i]);
    free(alphas); free(tmp);
}
static void blit_trapezoid_row(filler *af, int32_t ul, int32_t ur, int32_t ll, int32_t lr, int32_t ldy, int32_t rdy, uint8_t full) {   /* :1376-1446 */
    if (ldy < 0) ldy = -ldy;
    if (rdy < 0) rdy = -rdy;
    if (ul > ur) return;
    if (ll > lr) { int32_t mid = approximate_intersection(ul, ll, ur, lr); ll = mid; lr = mid; }
    if (ul == ur && ll == lr) return;
    if (ul > ll) { int32_t t = ul; ul = ll; ll = t; }
    if (ur > lr) { int32_t t = ur; ur = lr; lr = t; }
    int32_t join_left = sk_ceil_to_fixed(ll), join_rite = sk_floor_to_fixed(ur);
    if (join_left > join_rite) { blit_aaa_trapezoid_row(af, ul, ur, ll, lr, ldy, rdy, full); return; }
    /* blitLeftPartial :1448-1472 */
    if (ul < join_left) {
        switch (sk_ceil_to_int(join_left - ul)) {
        case 1: safe_add_alpha(af, sk_floor_to_int(ul), trapezoid_to_alpha_scaled(join_left - ul, join_left - ll, full)); break;
        case 2: {
            int32_t first = join_left - SK_FIXED1 - ul, second = ll - ul - first;
            uint8_t a1 = partial_triangle_to_alpha(first, ldy), a2 = sat_sub8(full, partial_triangle_to_alpha(second, ldy));
            safe_add_alpha(af, sk_floor_to_int(ul), a1);
            safe_add_alpha(af, sk_floor_to_int(ul) + 1, a2);
        } break;
        default: blit_aaa_trapezoid_row(af, ul, join_left, ll, join_left, ldy, MAX_S32, full); break;
        }
    }
    if (join_left < join_rite) {
        int32_t start = sk_floor_to_int(join_left), count = sk_floor_to_int(join_rite - join_left);
        for (int32_t i = 0; i < count; i++) safe_add_alpha(af, start + i, full);
    }
    /* blitRightPartial :1475-1499 */
    if (lr > join_rite) {
        switch (sk_ceil_to_int(lr - join_rite)) {
        case 1: safe_add_alpha(af, sk_floor_to_int(join_rite), trapezoid_to_alpha_scaled(ur - join_rite, lr - join_rite, full)); break;
        case 2: {
            int32_t first = join_rite + SK_FIXED1 - ur, second = lr - ur - first;
            uint8_t a1 = sat_sub8(full, partial_triangle_to_alpha(first, rdy)), a2 = partial_triangle_to_alpha(second, rdy);
            safe_add_alpha(af, sk_floor_to_int(join_rite), a1);
            safe_add_alpha(af, sk_floor_to_int(join_rite) + 1, a2);
        } break;
        default: blit_aaa_trapezoid_row(af, join_rite, ur, join_rite, lr, MAX_S32, rdy, full); break;
        }
    }
}
static void blit_between(filler *af, const edge_line_state *l, const edge_line_state *r) {   /* :1340-1365 */
    if (!l->valid || !r->valid) return;
    uint8_t full = l->full_alpha < r->full_alpha ? l->full_alpha : r->full_alpha;
    if (full == 0) return;
    blit_trapezoid_row(af, l->top_x, r->top_x, l->bot_x, r->bot_x, l->dy, r->dy, full);
}
static void update_next_next_y(filler *af, int32_t y, int32_t next_y) { if (y > next_y && y < af->next_next_y) af->next_next_y = y; }
static int step_curve_segment(edge_var *e) {   /* :1634-1646 */
    if (e->type == EDGE_QUAD) { if (e->quad.curve_count > 0) return quad_update(&e->quad); }
    else if (e->type == EDGE_CUBIC) { if (e->cubic.curve_count < 0) return cubic_update(&e->cubic); }
    return 0;
}
/* initSingleEdgeState :421-474 == reinitEdgeState :776-826 */
static void init_edge_state(filler *af, int idx, int32_t y_row) {
    line_edge *line = ev_line(&af->edge_buf[idx]);
    int precise = line->upper_y != 0 || line->lower_y != 0;
    edge_y_state st; memset(&st, 0, sizeof st);
    st.winding = line->winding;
    if (precise) {
        st.fupper_x = line->upper_x; st.fupper_y = line->upper_y; st.flower_y = line->lower_y; st.fdx = line->pixel_dx;
        st.fdy = line->pixel_dy != 0 ? line->pixel_dy : compute_edge_dy(line->pixel_dx);
        int32_t init_y = y_row;
        if (st.fupper_y > y_row) init_y = st.fupper_y;
        st.fx = line->upper_x + sk_fixed_mul(line->pixel_dx, init_y - line->upper_y);
    } else {
        int64_t s = af->aa_scale;
        st.fdx = line->dx;
        int32_t ref_x = (int32_t)((int64_t)line->x / s);
        int32_t ref_y = (int32_t)(((int64_t)line->first_y * SK_FIXED1 + SK_FIXED_HALF) / s);
        st.fupper_x = ref_x; st.fupper_y = ref_y;
        st.flower_y = (int32_t)((int64_t)(line->last_y + 1) * SK_FIXED1 / s);
        st.fdy = compute_edge_dy(line->dx);
        int32_t init_y = y_row;
        if (st.fupper_y > y_row) init_y = st.fupper_y;
        st.fx = ref_x + sk_fixed_mul(line->dx, init_y - ref_y);
    }
    st.valid = 1;
    af->states[idx] = st;
}
static int step_edge_state_to_strip(filler *af, int idx, int32_t top, int32_t bot) {   /* :721-771 */
    edge_var *e = &af->edge_buf[idx];
    if (e->type == EDGE_LINE) return 0;
    for (;;) {
        if (!step_curve_segment(e)) { af->states[idx].valid = 0; return 0; }
        line_edge *line = ev_line(e);
        int32_t seg_top, seg_bot;
        if (line->upper_y != 0 || line->lower_y != 0) { seg_top = line->upper_y; seg_bot = line->lower_y; }
        else {
            seg_top = (int32_t)((int64_t)line->first_y * SK_FIXED1 / af->aa_scale);
            seg_bot = (int32_t)((int64_t)(line->last_y + 1) * SK_FIXED1 / af->aa_scale);
        }
        if (seg_top >= bot) { init_edge_state(af, idx, top); return 0; }
        if (seg_bot <= top) continue;
        init_edge_state(af, idx, top);
        return 1;
    }
}
static void advance_edge_states(filler *af, int32_t top, int32_t bot, int32_t ydiff) {   /* :679-711 */
    int ys = compute_y_shift(ydiff);
    for (int i = 0; i < af->n_aet; i++) {
        int idx = af->aet[i];
        edge_y_state *st = &af->states[idx];
        if (!st->valid || st->fupper_y >= bot) continue;
        if (st->flower_y <= top) { if (!step_edge_state_to_strip(af, idx, top, bot)) continue; st = &af->states[idx]; }
        if (st->fupper_y <= top && st->flower_y >= bot) {
            if (ys >= 0) st->fx += st->fdx >> ys; else st->fx = st->fupper_x + sk_fixed_mul(st->fdx, bot - st->fupper_y);
        } else st->fx = st->fupper_x + sk_fixed_mul(st->fdx, bot - st->fupper_y);
    }
}
static void sort_by_top_x(edge_line_state *e, int n) {   /* :1139-1165 */
    for (int i = 1; i < n; i++) {
        edge_line_state key = e[i];
        int32_t kx = key.top_x, ks = key.bot_x - key.top_x;
        int j = i - 1;
        while (j >= 0) {
            int32_t ex = e[j].top_x;
            if (ex > kx) { e[j + 1] = e[j]; j--; continue; }
            if (ex == kx && (e[j].bot_x - e[j].top_x) > ks) { e[j + 1] = e[j]; j--; continue; }
            break;
        }
        e[j + 1] = key;
    }
}
static void sub_strip_no_split(filler *af, int32_t top, int32_t bot, int even_odd) {   /* :538-672 */
    int32_t ydiff = bot - top;
    uint8_t full = fixed_to_alpha(ydiff);
    if (full == 0) { advance_edge_states(af, top, bot, ydiff); return; }
    if (af->cap_resolved < af->n_aet) { af->cap_resolved = af->n_aet + 16; af->resolved = (edge_line_state *)realloc(af->resolved, sizeof(edge_line_state) * (size_t)af->cap_resolved); }
    af->n_resolved = 0;
    for (int i = 0; i < af->n_aet; i++) {
        int idx = af->aet[i];
        edge_y_state *st = &af->states[idx];
        if (!st->valid) continue;
        if (st->fupper_y >= bot) continue;
        if (st->flower_y <= top) { if (!step_edge_state_to_strip(af, idx, top, bot)) continue; st = &af->states[idx]; }
        update_next_next_y(af, st->flower_y, bot);
        int32_t ct = top, cb = bot;
        if (ct < st->fupper_y) ct = st->fupper_y;
        if (cb > st->flower_y) cb = st->flower_y;
        if (cb <= ct) continue;
        uint8_t ea = full;
        if (ct != top || cb != bot) { ea = fixed_to_alpha(cb - ct); if (ea == 0) continue; }
        int32_t top_x = st->fx, bot_x;
        if (ct == top && cb == bot) {
            int ys = compute_y_shift(ydiff);
            if (ys >= 0) bot_x = top_x + (st->fdx >> ys); else bot_x = st->fupper_x + sk_fixed_mul(st->fdx, bot - st->fupper_y);
            st->fx = bot_x;
        } else {
            top_x = st->fupper_x + sk_fixed_mul(st->fdx, ct - st->fupper_y);
            bot_x = st->fupper_x + sk_fixed_mul(st->fdx, cb - st->fupper_y);
            st->fx = st->fupper_x + sk_fixed_mul(st->fdx, bot - st->fupper_y);
        }
        edge_line_state r = {1, top_x, bot_x, st->fdy, ea, st->winding};
        af->resolved[af->n_resolved++] = r;
    }
    sort_by_top_x(af->resolved, af->n_resolved);
    int32_t winding = 0; int in = 0;
    edge_line_state left; memset(&left, 0, sizeof left);
    for (int i = 0; i < af->n_resolved; i++) {
        edge_line_state ls = af->resolved[i];
        winding += ls.winding;
        int prev = in;
        in = even_odd ? (winding & 1) != 0 : winding != 0;
        if (!in && prev) blit_between(af, &left, &ls);
        if (in && !prev) left = ls;
    }
}
static void sub_strip(filler *af, int32_t top, int32_t bot, int even_odd) {   /* :486-534 */
    int32_t ydiff = bot - top;
    if (fixed_to_alpha(ydiff) == 0) { advance_edge_states(af, top, bot, ydiff); return; }
    int32_t earliest = bot;
    for (int i = 0; i < af->n_aet; i++) {
        int idx = af->aet[i];
        edge_y_state *st = &af->states[idx];
        if (!st->valid || st->fupper_y >= bot) continue;
        if (st->flower_y <= top) { if (step_edge_state_to_strip(af, idx, top, bot)) st = &af->states[idx]; else continue; }
        if (st->flower_y > top && st->flower_y < earliest) earliest = st->flower_y;
    }
    if (earliest < bot) { sub_strip_no_split(af, top, earliest, even_odd); sub_strip(af, earliest, bot, even_odd); return; }
    sub_strip_no_split(af, top, bot, even_odd);
}
static void strip_push(filler *af, int32_t v) {
    if (af->n_strip == af->cap_strip) { af->cap_strip = af->cap_strip ? af->cap_strip * 2 : 32; af->strip_y = (int32_t *)realloc(af->strip_y, 4 * (size_t)af->cap_strip); }
    af->strip_y[af->n_strip++] = v;
}
static void sort_dedup_strips(filler *af) {   /* sortInt32s :1097, deduplicateInt32s :1111 */
    int32_t *s = af->strip_y; int n = af->n_strip;
    for (int i = 1; i < n; i++) { int32_t key = s[i]; int j = i - 1; while (j >= 0 && s[j] > key) { s[j + 1] = s[j]; j--; } s[j + 1] = key; }
    if (n <= 1) return;
    int m = 1;
    for (int i = 1; i < n; i++) if (s[i] - s[m - 1] > 128) s[m++] = s[i];
    af->n_strip = m;
}
static void compute_edge_x(const line_edge *line, int32_t aa_scale, int precise, int32_t ct, int32_t cb, int32_t *tx, int32_t *bx) {   /* :1296-1314 */
    if (precise) {
        *tx = line->upper_x + sk_fixed_mul(line->pixel_dx, ct - line->upper_y);
        *bx = line->upper_x + sk_fixed_mul(line->pixel_dx, cb - line->upper_y);
        return;
    }
    int32_t ref_x = (int32_t)((int64_t)line->x / aa_scale);
    int32_t ref_y = (int32_t)(((int64_t)line->first_y * SK_FIXED1 + SK_FIXED_HALF) / aa_scale);
    *tx = ref_x + sk_fixed_mul(line->dx, ct - ref_y);
    *bx = ref_x + sk_fixed_mul(line->dx, cb - ref_y);
}
static int has_edge_crossing(filler *af, int32_t ytop, int32_t ybot) {   /* :989-1022 */
    int n = af->n_aet;
    if (n < 2) return 0;
    int32_t *tx = (int32_t *)malloc(8 * (size_t)n), *bx = tx + n;
    for (int i = 0; i < n; i++) {
        line_edge *line = ev_line(&af->edge_buf[af->aet[i]]);
        compute_edge_x(line, af->aa_scale, line->upper_y != 0 || line->lower_y != 0, ytop, ybot, &tx[i], &bx[i]);
    }
    int hit = 0;
    for (int i = 0; i < n && !hit; i++) for (int j = i + 1; j < n; j++) {
        int64_t dt = (int64_t)tx[i] - tx[j], db = (int64_t)bx[i] - bx[j];
        if ((dt > 0 && db < 0) || (dt < 0 && db > 0)) { hit = 1; break; }
    }
    free(tx);
    return hit;
}
static void collect_strip_boundaries(filler *af, int32_t ytop, int32_t ybot) {   /* :854-987 */
    int64_t s = af->aa_scale;
    af->n_strip = 0;
    strip_push(af, ytop); strip_push(af, ybot);
    if (af->next_next_y > ytop && af->next_next_y < ybot) strip_push(af, af->next_next_y);
    {
        int32_t quarter = SK_FIXED1 / 4, y = ytop;
        while (y < ybot) {
            int32_t next = ybot;
            if (has_edge_crossing(af, y, next)) { next = y + quarter; if (next > ybot) next = ybot; }
            if (next > y && next < ybot) strip_push(af, next);
            y = next;
        }
    }
    for (int i = 0; i < af->n_aet; i++) {
        edge_var *e = &af->edge_buf[af->aet[i]];
        line_edge *line = ev_line(e);
        int32_t seg_top, seg_bot;
        if (line->upper_y != 0 || line->lower_y != 0) { seg_top = line->upper_y; seg_bot = line->lower_y; }
        else { seg_top = (int32_t)((int64_t)line->first_y * SK_FIXED1 / s); seg_bot = (int32_t)((int64_t)(line->last_y + 1) * SK_FIXED1 / s); }
        int32_t e_top = (int32_t)((int64_t)ev_top(e) * SK_FIXED1 / s), e_bot = (int32_t)((int64_t)ev_bottom(e) * SK_FIXED1 / s);
        int32_t ys[4] = {seg_top, seg_bot, e_top, e_bot};
        for (int k = 0; k < 4; k++) if (ys[k] > ytop && ys[k] < ybot) strip_push(af, ys[k]);
    }
    for (int idx = af->edge_idx; idx < af->n_edges; idx++) {
        int32_t t = (int32_t)((int64_t)ev_top(&af->edge_buf[idx]) * SK_FIXED1 / s);
        if (t >= ybot) break;
        if (t > ytop) strip_push(af, t);
    }
    for (int i = 0; i < af->n_aet; i++) {
        int idx = af->aet[i];
        if (af->edge_buf[idx].type == EDGE_LINE) continue;
        edge_y_state *st = &af->states[idx];
        if (!st->valid) continue;
        if (st->flower_y > ytop && st->flower_y < ybot) strip_push(af, st->flower_y);
    }
    sort_dedup_strips(af);
    for (;;) {
        int added = 0;
        for (int i = 0; i < af->n_strip - 1; i++) {
            int32_t d = af->strip_y[i + 1] - af->strip_y[i];
            if (d > 0 && (d & (SK_FIXED1 >> 2)) != 0 && d != (SK_FIXED1 >> 2)) { strip_push(af, af->strip_y[i] + (SK_FIXED1 >> 2)); added = 1; break; }
        }
        if (!added) break;
        sort_dedup_strips(af);
    }
}
static void aet_insert(filler *af, int idx) {
    if (af->n_aet == af->cap_aet) { af->cap_aet = af->cap_aet ? af->cap_aet * 2 : 64; af->aet = (int *)realloc(af->aet, sizeof(int) * (size_t)af->cap_aet); }
    af->aet[af->n_aet++] = idx;
}
/* alpha_runs.go */
static uint8_t catch_overflow(uint16_t a) { if (a > 256) a = 256; return (uint8_t)(a - (a >> 8)); }
static void runs_reset(filler *af) { af->run_offset = 0; af->runs[0] = af->width > 65535 ? 65535 : (uint16_t)af->width; af->runs[af->width] = 0; af->alpha[0] = 0; }
static void break_run(filler *af, int ro0, int x, int count) {   /* alpha_runs.go:208-263 */
    if (count <= 0) return;
    int orig = x, ro = ro0, ao = ro0;
    while (x > 0) {
        int n = af->runs[ro];
        if (n <= 0) return;
        if (x < n) { af->alpha[ao + x] = af->alpha[ao]; af->runs[ro] = (uint16_t)x; af->runs[ro + x] = (uint16_t)(n - x); break; }
        ro += n; ao += n; x -= n;
    }
    ro = ro0 + orig; ao = ro0 + orig; x = count;
    for (;;) {
        int n = af->runs[ro];
        if (n <= 0) break;
        if (x < n) { af->alpha[ao + x] = af->alpha[ao]; af->runs[ro] = (uint16_t)x; af->runs[ro + x] = (uint16_t)(n - x); break; }
        x -= n;
        if (x == 0) break;
        ro += n; ao += n;
    }
}
static void runs_add(filler *af, int x, uint8_t start, int middle, uint8_t end, uint8_t maxv) {   /* alpha_runs.go:133-204 */
    if (x < 0 || x >= af->width) return;
    int ro = af->run_offset, ao = af->run_offset, last = af->run_offset;
    x -= af->run_offset;
    if (start != 0) {
        break_run(af, ro, x, 1);
        af->alpha[ao + x] = catch_overflow((uint16_t)(af->alpha[ao + x] + start));
        ro += x + 1; ao += x + 1; x = 0;
    }
    if (middle > 0) {
        break_run(af, ro, x, middle);
        ao += x; ro += x; x = 0;
        int rem = middle;
        while (rem > 0) {
            af->alpha[ao] = catch_overflow((uint16_t)(af->alpha[ao] + maxv));
            int n = af->runs[ro];
            if (n <= 0) break;
            if (n > rem) n = rem;
            ao += n; ro += n; rem -= n;
        }
        last = ao;
    }
    if (end != 0) {
        break_run(af, ro, x, 1);
        ao += x;
        af->alpha[ao] = catch_overflow((uint16_t)(af->alpha[ao] + end));
        last = ao;
    }
    af->run_offset = last;
}
static void coverage_to_runs(filler *af) {   /* analytic_filler.go:1603-1631 */
    runs_reset(af);
    uint8_t cur = 0; int start = 0;
    for (int i = 0; i < af->width; i++) {
        uint8_t a = af->coverage[i];
        if (i == 0) { cur = a; continue; }
        if (a != cur) {
            if (cur > 0) runs_add(af, start, cur, i - start - 1, 0, cur);
            cur = a; start = i;
        }
    }
    if (cur > 0) runs_add(af, start, cur, af->width - start - 1, 0, cur);
}
static void runs_copy_to(const filler *af, uint8_t *dst) {   /* alpha_runs.go:318-336 */
    int x = 0;
    while (x < af->width) {
        int n = af->runs[x];
        if (n <= 0) break;
        uint8_t a = af->alpha[x];
        for (int i = 0; i < n && x + i < af->width; i++) dst[x + i] = a;
        x += n;
    }
}
static void process_scanline(filler *af, int y, int even_odd) {   /* :260-409 */
    memset(af->coverage, 0, (size_t)af->width);
    int32_t aa = af->aa_scale;
    int32_t y_sub = y * aa, y_sub_next = y_sub + aa;
    {   /* RemoveExpiredSubpixel, curve_aet.go:125-138 */
        int n = 0;
        for (int i = 0; i < af->n_aet; i++) if (ev_bottom(&af->edge_buf[af->aet[i]]) > y_sub) af->aet[n++] = af->aet[i];
        af->n_aet = n;
    }
    int32_t y_fixed = y << 16, y_fixed_end = (y + 1) << 16;
    while (af->edge_idx < af->n_edges) {
        edge_var *e = &af->edge_buf[af->edge_idx];
        if (ev_top(e) >= y_sub_next) break;
        aet_insert(af, af->edge_idx);
        init_edge_state(af, af->edge_idx, y_fixed);
        line_edge *line = ev_line(e);
        if (line->upper_y != 0 || line->lower_y != 0) update_next_next_y(af, line->lower_y, y_fixed);
        af->edge_idx++;
    }
    for (int idx = af->edge_idx; idx < af->n_edges; idx++) {
        line_edge *line = ev_line(&af->edge_buf[idx]);
        if (!(line->upper_y != 0 || line->lower_y != 0)) break;
        int32_t upper = line->upper_y;
        if (upper >= y_fixed_end) break;
        if (af->n_deferred == af->cap_deferred) { af->cap_deferred = af->cap_deferred ? af->cap_deferred * 2 : 16; af->deferred = (deferred_edge *)realloc(af->deferred, sizeof(deferred_edge) * (size_t)af->cap_deferred); }
        af->deferred[af->n_deferred].idx = idx; af->deferred[af->n_deferred].upper_y = upper; af->n_deferred++;
        if (line->lower_y != 0) update_next_next_y(af, line->lower_y, y_fixed);
        af->edge_idx = idx + 1;
    }
    if (af->edge_idx < af->n_edges) {
        line_edge *line = ev_line(&af->edge_buf[af->edge_idx]);
        if (line->upper_y != 0 || line->lower_y != 0) update_next_next_y(af, line->upper_y, y_fixed);
    }
    collect_strip_boundaries(af, y_fixed, y_fixed_end);
    /* the strip list is re-used by nested calls? no: collect runs once per row; copy it, sub-strips do not touch it */
    int ns = af->n_strip;
    int32_t *ys = (int32_t *)malloc(4 * (size_t)(ns > 0 ? ns : 1));
    memcpy(ys, af->strip_y, 4 * (size_t)ns);
    for (int si = 0; si < ns - 1; si++) {
        int32_t top = ys[si], bot = ys[si + 1];
        if (bot <= top) continue;
        for (int i = 0; i < af->n_deferred;) {
            deferred_edge d = af->deferred[i];
            if (d.upper_y <= top) {
                aet_insert(af, d.idx);
                init_edge_state(af, d.idx, y_fixed);
                af->edge_idx = d.idx + 1;
                memmove(af->deferred + i, af->deferred + i + 1, sizeof(deferred_edge) * (size_t)(af->n_deferred - i - 1));
                af->n_deferred--;
            } else i++;
        }
        sub_strip(af, top, bot, even_odd);
    }
    free(ys);
    for (int i = 0; i < af->n_deferred; i++) { aet_insert(af, af->deferred[i].idx); init_edge_state(af, af->deferred[i].idx, y_fixed); af->edge_idx = af->deferred[i].idx + 1; }
    af->n_deferred = 0;
    coverage_to_runs(af);
}

/* EdgeBuilder.sortedEdgesSlice (:1290-1339): lines, quadratics, cubics, stable-sorted by top Y -- then AnalyticFiller.Fill (:159-243).
 * row_cb(y, coverage row) is called for every scanline of the path's bounds. */
typedef void (*aaa_row_fn)(void *ud, int y, const uint8_t *row);
static void aaa_fill(edge_builder *eb, int width, int height, int even_odd, aaa_row_fn cb, void *ud) {
    int n = eb->n_lines + eb->n_quads + eb->n_cubics;
    if (n == 0) return;
    filler af; memset(&af, 0, sizeof af);
    af.width = width; af.height = height;
    af.aa_scale = 1 << eb->aa_shift;
    af.coverage = (uint8_t *)calloc((size_t)width, 1);
    af.runs = (uint16_t *)calloc((size_t)width + 1, 2); af.alpha = (uint8_t *)calloc((size_t)width + 1, 1);
    af.edge_buf = (edge_var *)calloc((size_t)n, sizeof(edge_var)); af.n_edges = n;
    {
        edge_var *tmp = (edge_var *)calloc((size_t)n, sizeof(edge_var));
        int k = 0;
        for (int i = 0; i < eb->n_lines; i++) { tmp[k].type = EDGE_LINE; tmp[k].line = eb->lines[i]; k++; }
        for (int i = 0; i < eb->n_quads; i++) { tmp[k].type = EDGE_QUAD; tmp[k].quad = eb->quads[i]; k++; }
        for (int i = 0; i < eb->n_cubics; i++) { tmp[k].type = EDGE_CUBIC; tmp[k].cubic = eb->cubics[i]; k++; }
        /* stable insertion by top Y (slices.SortStableFunc) -- merge sort for large paths */
        int *order = (int *)malloc(sizeof(int) * (size_t)n), *aux = (int *)malloc(sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) order[i] = i;
        for (int w = 1; w < n; w *= 2) {
            for (int lo = 0; lo < n; lo += 2 * w) {
                int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, o = lo;
                while (a < mid && b < hi) { if (ev_top(&tmp[order[b]]) < ev_top(&tmp[order[a]])) aux[o++] = order[b++]; else aux[o++] = order[a++]; }
                while (a < mid) aux[o++] = order[a++];
                while (b < hi) aux[o++] = order[b++];
            }
            int *t = order; order = aux; aux = t;
        }
        for (int i = 0; i < n; i++) af.edge_buf[i] = tmp[order[i]];
        free(order); free(aux); free(tmp);
    }
    af.states = (edge_y_state *)calloc((size_t)n, sizeof(edge_y_state));
    af.next_next_y = MAX_S32;
    int y_min = (int)floor((double)eb->bminy), y_max = (int)ceil((double)eb->bmaxy);
    if (eb->b_empty) { y_min = 0; y_max = 0; }
    if (y_min < 0) y_min = 0;
    if (y_max > height) y_max = height;
    uint8_t *row = (uint8_t *)malloc((size_t)width);
    for (int y = y_min; y < y_max; y++) {
        process_scanline(&af, y, even_odd);
        memset(row, 0, (size_t)width);
        runs_copy_to(&af, row);
        cb(ud, y, row);
    }
    free(row); free(af.coverage); free(af.runs); free(af.alpha); free(af.edge_buf); free(af.states); free(af.aet); free(af.resolved); free(af.strip_y); free(af.deferred);
}
static void eb_init(edge_builder *eb, int aa_shift, int flatten) { memset(eb, 0, sizeof *eb); eb->aa_shift = aa_shift; eb->flatten = flatten; eb->b_empty = 1; }
static void eb_free(edge_builder *eb) { free(eb->lines); free(eb->quads); free(eb->cubics); }

/* ------------------------------------------------------------------ exported entry points */
typedef struct { uint8_t *buf; int width; } cov_sink;
static void cov_row(void *ud, int y, const uint8_t *row) { cov_sink *s = (cov_sink *)ud; memcpy(s->buf + (size_t)y * s->width, row, (size_t)s->width); }

/* raster.FillToBuffer (analytic_filler.go:2587-2611) on a path: 8-bit coverage, width*height. clip_margin < 0: no clip rectangle
 * (the raster package's own tests); >= 0: SoftwareRenderer.Fill's canvas clip (software.go:530-540). */
void oa_coverage(const uint8_t *verbs, uint32_t n_verbs, const double *coords, int width, int height, int even_odd,
                 int aa_shift, int flatten, float clip_margin, uint8_t *out) {
    edge_builder eb; eb_init(&eb, aa_shift, flatten);
    if (clip_margin >= 0) { eb.has_clip = 1; eb.clip[0] = -clip_margin; eb.clip[1] = -clip_margin; eb.clip[2] = (float)width + clip_margin; eb.clip[3] = (float)height + clip_margin; }
    eb_build(&eb, verbs, n_verbs, coords);
    memset(out, 0, (size_t)width * height);
    cov_sink s = {out, width};
    aaa_fill(&eb, width, height, even_odd, cov_row, &s);
    eb_free(&eb);
}

/* SoftwareRenderer.Fill with a solid colour in source-over mode (software.go:485-587, 953-1026): the pixmap holds
 * premultiplied RGBA8, every draw reads it back and TRUNCATES the float64 result to 8 bits (pixmap.go:218-228). */
typedef struct { uint8_t *pix; int width, height; double r, g, b, a; } blend_sink;
static double clamp255(double x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }
static void blend_row(void *ud, int y, const uint8_t *row) {
    blend_sink *s = (blend_sink *)ud;
    if (y < 0 || y >= s->height) return;
    for (int x = 0; x < s->width; x++) {
        uint8_t alpha = row[x];
        if (alpha == 0) continue;
        uint8_t *p = s->pix + 4 * ((size_t)y * s->width + x);
        if (alpha == 255 && s->a == 1.0) {   /* Pixmap.SetPixel: premultiply, truncate (pixmap.go:118-128) */
            p[0] = (uint8_t)clamp255(s->r * s->a * 255); p[1] = (uint8_t)clamp255(s->g * s->a * 255);
            p[2] = (uint8_t)clamp255(s->b * s->a * 255); p[3] = (uint8_t)clamp255(s->a * 255);
            continue;
        }
        double sa = s->a * (double)alpha / 255.0, inv = 1.0 - sa;
        double sr = s->r * sa, sg = s->g * sa, sb = s->b * sa;
        double dr = (double)p[0] / 255, dg = (double)p[1] / 255, db = (double)p[2] / 255, da = (double)p[3] / 255;
        p[0] = (uint8_t)clamp255((sr + dr * inv) * 255); p[1] = (uint8_t)clamp255((sg + dg * inv) * 255);
        p[2] = (uint8_t)clamp255((sb + db * inv) * 255); p[3] = (uint8_t)clamp255((sa + da * inv) * 255);
    }
}
void oa_fill(uint8_t *pixmap_premul, int width, int height, const uint8_t *verbs, uint32_t n_verbs, const double *coords,
             const double rgba_straight[4], int even_odd) {
    edge_builder eb; eb_init(&eb, 2, 0);   /* NewEdgeBuilder(2), SetFlattenCurves(false) */
    eb.has_clip = 1; eb.clip[0] = -2; eb.clip[1] = -2; eb.clip[2] = (float)width + 2; eb.clip[3] = (float)height + 2;
    eb_build(&eb, verbs, n_verbs, coords);
    blend_sink s = {pixmap_premul, width, height, rgba_straight[0], rgba_straight[1], rgba_straight[2], rgba_straight[3]};
    aaa_fill(&eb, width, height, even_odd, blend_row, &s);
    eb_free(&eb);
}
