/*
 * oracle/blend.c -- TEST INFRASTRUCTURE ONLY (see twin.c header).
 *
 * 1. ot_blend_bytes: bit-exact restatement of gg's 29 byte blend functions
 *    (internal/blend/porter_duff.go:117-247, advanced.go:52-270, hsl.go:15-289, math.go:9-84;
 *    the same formulas as blend_funcs.go), selected by scene.BlendMode through
 *    scene/blend_integration.go:12-80. Pinned by the known-answer vectors of the reference's
 *    porter_duff_test.go / advanced_test.go / hsl_test.go (tests/test_oracle_blend.py).
 * 2. ot_blend_f32: the float32 layer-composite ggcuda's fine stage applies at CmdEndClip.
 *    The reference has NO end-to-end definition of layer blending (scene/renderer.go:716-721 are
 *    TODOs, tilecompute/fine.go:164 ignores the blend word): this function DEFINES it as the
 *    W3C / internal/blend formula  Co = (1-Da) S + (1-Sa) D + Sa Da B(Cs, Cd)  evaluated in float32
 *    on premultiplied inputs, with B() and the Porter-Duff operators taken from (1). "parity unpinned"
 *    for layer pixels: only (1) is pinned; tests bound |f32 - bytes| <= 2/255.
 *    Blend word: (mix << 8) | compose, Vello/peniko numbering; 0x8003 = clip; bit 31 = elide flag.
 */
#include "twin.h"

#include <math.h>

typedef uint8_t u8;
static inline u8 mul_div255(u8 a, u8 b) { return (u8)(((uint16_t)a * (uint16_t)b + 255) >> 8); }       /* math.go:9-30 */
static inline u8 add_div255(u8 a, u8 b) { uint16_t s = (uint16_t)a + b; return s > 255 ? 255 : (u8)s; } /* porter_duff.go:217-223 */
static inline u8 min_b(u8 a, u8 b) { return a < b ? a : b; }
static inline u8 max_b(u8 a, u8 b) { return a > b ? a : b; }

/* ---- separable channel functions, advanced.go:90-256 ---- */
static u8 ch_multiply(u8 s, u8 d) { return mul_div255(s, d); }
static u8 ch_screen(u8 s, u8 d) { return (u8)(255 - mul_div255((u8)(255 - s), (u8)(255 - d))); }
static u8 ch_overlay(u8 s, u8 d) {
    if (d <= 128) return mul_div255((u8)(2 * d), s);                      /* byte wrap of 2*d kept (d == 128 -> 0) */
    return (u8)(255 - mul_div255((u8)(2 * (u8)(255 - d)), (u8)(255 - s)));
}
static u8 ch_darken(u8 s, u8 d) { return min_b(s, d); }
static u8 ch_lighten(u8 s, u8 d) { return max_b(s, d); }
static u8 ch_color_dodge(u8 s, u8 d) {
    if (s == 255) return 255;
    uint16_t r = (uint16_t)(((uint16_t)d * 255) / (uint16_t)(255 - s));
    return r > 255 ? 255 : (u8)r;
}
static u8 ch_color_burn(u8 s, u8 d) {
    if (s == 0) return 0;
    uint16_t r = (uint16_t)(((uint16_t)(255 - d) * 255) / (uint16_t)s);
    return r > 255 ? 0 : (u8)(255 - (u8)r);
}
static u8 ch_hard_light(u8 s, u8 d) {
    if (s <= 128) return mul_div255((u8)(2 * s), d);
    return (u8)(255 - mul_div255((u8)(2 * (u8)(255 - s)), (u8)(255 - d)));
}
static u8 ch_soft_light(u8 s, u8 d) {
    double sf = (double)s / 255.0, df = (double)d / 255.0, r;
    if (sf <= 0.5) r = df - (1 - 2 * sf) * df * (1 - df);
    else { double dx = df <= 0.25 ? ((16 * df - 12) * df + 4) * df : sqrt(df); r = df + (2 * sf - 1) * (dx - df); }
    if (r < 0) return 0;
    if (r > 1) return 255;
    return (u8)(r * 255);
}
static u8 ch_difference(u8 s, u8 d) { return s > d ? (u8)(s - d) : (u8)(d - s); }
static u8 ch_exclusion(u8 s, u8 d) {
    uint16_t sum = (uint16_t)s + d, diff = (uint16_t)(sum - 2 * (uint16_t)mul_div255(s, d));
    return diff > 255 ? 255 : (u8)diff;
}

static void separable(const u8 s[4], const u8 d[4], u8 (*f)(u8, u8), u8 o[4]) {   /* advanced.go:52-100 */
    u8 sa = s[3], da = d[3];
    if (sa == 0) { o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; o[3] = d[3]; return; }
    if (da == 0) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; return; }
    u8 inv_sa = (u8)(255 - sa), inv_da = (u8)(255 - da), sa_da = mul_div255(sa, da);
    for (int k = 0; k < 3; k++) {
        u8 su = (u8)(((uint16_t)s[k] * 255) / sa), du = (u8)(((uint16_t)d[k] * 255) / da);
        u8 b = f(su, du);
        u8 r = add_div255(mul_div255(d[k], inv_sa), mul_div255(s[k], inv_da));
        o[k] = add_div255(r, mul_div255(sa_da, b));
    }
    o[3] = add_div255(sa, mul_div255(da, inv_sa));
}

/* ---- non-separable, hsl.go ---- */
static float lum(float r, float g, float b) { return 0.30f * r + 0.59f * g + 0.11f * b; }
static float min3(float a, float b, float c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
static float max3(float a, float b, float c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
static float sat(float r, float g, float b) { return max3(r, g, b) - min3(r, g, b); }
static void clip_color(float *r, float *g, float *b) {   /* hsl.go:24-43 */
    float l = lum(*r, *g, *b), n = min3(*r, *g, *b), x = max3(*r, *g, *b);
    if (n < 0) { *r = l + (*r - l) * l / (l - n); *g = l + (*g - l) * l / (l - n); *b = l + (*b - l) * l / (l - n); }
    if (x > 1) { *r = l + (*r - l) * (1 - l) / (x - l); *g = l + (*g - l) * (1 - l) / (x - l); *b = l + (*b - l) * (1 - l) / (x - l); }
}
static void set_lum(float *r, float *g, float *b, float l) { float dd = l - lum(*r, *g, *b); *r += dd; *g += dd; *b += dd; clip_color(r, g, b); }
static void set_sat(float *r, float *g, float *b, float s) {   /* hsl.go:53-88 */
    float *mn, *md, *mx;
    if (*r <= *g && *g <= *b) { mn = r; md = g; mx = b; }
    else if (*r <= *b && *b <= *g) { mn = r; md = b; mx = g; }
    else if (*b <= *r && *r <= *g) { mn = b; md = r; mx = g; }
    else if (*g <= *r && *r <= *b) { mn = g; md = r; mx = b; }
    else if (*g <= *b && *b <= *r) { mn = g; md = b; mx = r; }
    else { mn = b; md = g; mx = r; }
    float lo = *mn, mi = *md, hi = *mx;
    if (hi > lo) { *md = ((mi - lo) * s) / (hi - lo); *mx = s; *mn = 0; }
}
/* which: 0 hue, 1 saturation, 2 color, 3 luminosity (hsl.go:90-121) */
static void hsl_blend(int which, float sr, float sg, float sb, float dr, float dg, float db, float *r, float *g, float *b) {
    switch (which) {
    case 0: *r = sr; *g = sg; *b = sb; set_sat(r, g, b, sat(dr, dg, db)); set_lum(r, g, b, lum(dr, dg, db)); break;
    case 1: *r = dr; *g = dg; *b = db; set_sat(r, g, b, sat(sr, sg, sb)); set_lum(r, g, b, lum(dr, dg, db)); break;
    case 2: *r = sr; *g = sg; *b = sb; set_lum(r, g, b, lum(dr, dg, db)); break;
    default: *r = dr; *g = dg; *b = db; set_lum(r, g, b, lum(sr, sg, sb)); break;
    }
}
static void non_separable(const u8 s[4], const u8 d[4], int which, u8 o[4]) {   /* hsl.go:236-289 */
    u8 sa = s[3], da = d[3];
    if (sa == 0) { o[0] = d[0]; o[1] = d[1]; o[2] = d[2]; o[3] = d[3]; return; }
    if (da == 0) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; return; }
    float sur = (float)s[0] / (float)sa, sug = (float)s[1] / (float)sa, sub = (float)s[2] / (float)sa;
    float dur = (float)d[0] / (float)da, dug = (float)d[1] / (float)da, dub = (float)d[2] / (float)da;
    float br, bg, bb;
    hsl_blend(which, sur, sug, sub, dur, dug, dub, &br, &bg, &bb);
    u8 inv_sa = (u8)(255 - sa), inv_da = (u8)(255 - da);
    float sa_da = ((float)sa / 255.0f) * ((float)da / 255.0f);
    float bl[3] = {br, bg, bb};
    for (int k = 0; k < 3; k++) {
        u8 r = add_div255(mul_div255(d[k], inv_sa), mul_div255(s[k], inv_da));
        /* Go: byte(math.Round(float64(blend*saDa*255))) -- float64->byte conversion of an in-range value */
        double c = round((double)(bl[k] * sa_da * 255.0f));
        o[k] = add_div255(r, (u8)(int64_t)c);
    }
    o[3] = add_div255(sa, mul_div255(da, inv_sa));
}

/* scene.BlendMode (scene/encoding.go:17-48) -> byte blend, premultiplied RGBA8 in and out */
void ot_blend_bytes(uint32_t scene_mode, const uint8_t s[4], const uint8_t d[4], uint8_t o[4]) {
    u8 sa = s[3], da = d[3], inv_sa = (u8)(255 - sa), inv_da = (u8)(255 - da);
    switch (scene_mode) {
    case 1: separable(s, d, ch_multiply, o); return;
    case 2: separable(s, d, ch_screen, o); return;
    case 3: separable(s, d, ch_overlay, o); return;
    case 4: separable(s, d, ch_darken, o); return;
    case 5: separable(s, d, ch_lighten, o); return;
    case 6: separable(s, d, ch_color_dodge, o); return;
    case 7: separable(s, d, ch_color_burn, o); return;
    case 8: separable(s, d, ch_hard_light, o); return;
    case 9: separable(s, d, ch_soft_light, o); return;
    case 10: separable(s, d, ch_difference, o); return;
    case 11: separable(s, d, ch_exclusion, o); return;
    case 12: non_separable(s, d, 0, o); return;
    case 13: non_separable(s, d, 1, o); return;
    case 14: non_separable(s, d, 2, o); return;
    case 15: non_separable(s, d, 3, o); return;
    case 16: o[0] = o[1] = o[2] = o[3] = 0; return;                                              /* Clear */
    case 17: for (int k = 0; k < 4; k++) o[k] = s[k]; return;                                      /* Copy */
    case 18: for (int k = 0; k < 4; k++) o[k] = d[k]; return;                                      /* Destination */
    case 20: for (int k = 0; k < 4; k++) o[k] = add_div255(mul_div255(s[k], inv_da), d[k]); return;/* DestinationOver */
    case 21: for (int k = 0; k < 4; k++) o[k] = mul_div255(s[k], da); return;                      /* SourceIn */
    case 22: for (int k = 0; k < 4; k++) o[k] = mul_div255(d[k], sa); return;                      /* DestinationIn */
    case 23: for (int k = 0; k < 4; k++) o[k] = mul_div255(s[k], inv_da); return;                  /* SourceOut */
    case 24: for (int k = 0; k < 4; k++) o[k] = mul_div255(d[k], inv_sa); return;                  /* DestinationOut */
    case 25: for (int k = 0; k < 3; k++) o[k] = add_div255(mul_div255(s[k], da), mul_div255(d[k], inv_sa)); o[3] = da; return;  /* SourceAtop */
    case 26: for (int k = 0; k < 3; k++) o[k] = add_div255(mul_div255(s[k], inv_da), mul_div255(d[k], sa)); o[3] = sa; return;  /* DestinationAtop */
    case 27: for (int k = 0; k < 4; k++) o[k] = add_div255(mul_div255(s[k], inv_da), mul_div255(d[k], inv_sa)); return;         /* Xor */
    case 28: for (int k = 0; k < 4; k++) o[k] = add_div255(s[k], d[k]); return;                    /* Plus */
    default: for (int k = 0; k < 4; k++) o[k] = add_div255(s[k], mul_div255(d[k], inv_sa)); return;/* Normal / SourceOver */
    }
}

/* ------------------------------------------------------------------ float32 layer composite (definition) */
static float mixf(float a, float b, float t) { return a + (b - a) * t; }
static float sep_f32(uint32_t mix, float s, float d) {
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

/* bg, fg: premultiplied RGBA float32. out = bg (blend) fg. */
void ot_blend_f32(uint32_t blend, const float bg[4], const float fg[4], float out[4]) {
    blend &= 0x3fffffffu;   /* bits 30-31: coarse's layer flags */
    uint32_t mix = (blend >> 8) & 0xffu, compose = blend & 0xffu;
    float sa = fg[3], da = bg[3];
    if ((mix == 0 || mix == 0x80u) && compose == 3u) {   /* Normal or Clip with SrcOver: fine.go:168-179 */
        float inv = 1.0f - sa;
        for (int k = 0; k < 4; k++) out[k] = bg[k] * inv + fg[k];
        return;
    }
    if (mix != 0 && mix < 16) {   /* mix modes always compose SrcOver: Co = (1-Da) S + (1-Sa) D + Sa Da B */
        if (sa <= 0.0f) { for (int k = 0; k < 4; k++) out[k] = bg[k]; return; }
        if (da <= 0.0f) { for (int k = 0; k < 4; k++) out[k] = fg[k]; return; }
        float isa = 1.0f / sa, ida = 1.0f / da;
        float cs[3] = {fg[0] * isa, fg[1] * isa, fg[2] * isa}, cd[3] = {bg[0] * ida, bg[1] * ida, bg[2] * ida}, b[3];
        if (mix >= 12) hsl_blend((int)mix - 12, cs[0], cs[1], cs[2], cd[0], cd[1], cd[2], &b[0], &b[1], &b[2]);
        else for (int k = 0; k < 3; k++) b[k] = sep_f32(mix, cs[k], cd[k]);
        float sada = sa * da;
        for (int k = 0; k < 3; k++) out[k] = (1.0f - da) * fg[k] + (1.0f - sa) * bg[k] + sada * b[k];
        out[3] = sa + da * (1.0f - sa);
        return;
    }
    /* Porter-Duff compose (mix Normal): result = Fa * S + Fb * D */
    float fa, fb;
    switch (compose) {
    case 0: fa = 0; fb = 0; break;                    /* Clear */
    case 1: fa = 1; fb = 0; break;                    /* Copy */
    case 2: fa = 0; fb = 1; break;                    /* Dest */
    case 4: fa = 1 - da; fb = 1; break;               /* DestOver */
    case 5: fa = da; fb = 0; break;                   /* SrcIn */
    case 6: fa = 0; fb = sa; break;                   /* DestIn */
    case 7: fa = 1 - da; fb = 0; break;               /* SrcOut */
    case 8: fa = 0; fb = 1 - sa; break;               /* DestOut */
    case 9: fa = da; fb = 1 - sa; break;              /* SrcAtop */
    case 10: fa = 1 - da; fb = sa; break;             /* DestAtop */
    case 11: fa = 1 - da; fb = 1 - sa; break;         /* Xor */
    case 12: fa = 1; fb = 1; break;                   /* Plus (clamped) */
    default: fa = 1; fb = 1 - sa; break;              /* SrcOver */
    }
    for (int k = 0; k < 4; k++) { float v = fa * fg[k] + fb * bg[k]; out[k] = compose == 12 && v > 1.0f ? 1.0f : v; }
    (void)mixf;
}
