// gg_b200/csrc/fine.cu -- fine rasterisation: per-tile PTCL replay with analytic area coverage.
//
// Behavioural spec: gg internal/gpu/tilecompute/fine.go:40-289 (fineRasterizeTile, fillPath), the CPU twin of
// tilecompute/shaders/fine.wgsl. One warp owns a PAIR of horizontally adjacent 16x16 tiles (so that every row it
// finally stores is one full 128-byte line); inside a tile a lane owns 8 consecutive pixels of one row (2 lanes per
// row), colour in registers. Both input streams of a tile reach shared memory through the bulk-copy engine
// (cp.async.bulk completing on mbarriers, SASS UBLKCP): the command list in 512-byte chunks through a two-slot ring,
// and the PathSegment slice of every CmdFill in chunks of 32 segments through a second two-slot ring, the slice of the
// NEXT fill (found by decoding ahead in the command list) in flight while the current one is evaluated.
//
// Coverage (fillPath, fine.go:219-289): the 32 segments of a chunk are first looked at in parallel (lane = segment: rows
// crossed, 1/dy); a prefix sum of the row counts then deals the chunk's (segment, row) pairs to the lanes, 32 at a time.
// A pair evaluates the reference's trapezoid formula for the few columns its segment passes through in that row and
// records ONE difference for everything to the right (where the formula gives exactly dy) in the row's table, with
// integer shared-memory atomics (native ATOMS.ADD; 2^-20 fixed point: sums independent of the order the segments arrive
// in, frames reproducible bit for bit). A running sum over a lane's 8 entries at the end of the fill yields its areas.
// Round 1 accumulated with shared-memory FLOAT atomics -- 15 ATOMS.CAST.SPIN compare-and-swap loops in the innermost
// loop (profiles/r1c_fine_kernel_ncu.txt) -- and its results depended on the order the atomics landed in.
//
// Everything that is not on the path of a plain fill -- layer blending (29 modes), brushes (gradients, SDF round rects,
// images) -- lives in out-of-line functions working on four pixels at a time through shared memory: the command loop has
// to fit the 32 KB instruction cache of an SM (profiles/r2_summary.md).
#include "pipeline.cuh"

#define FINE_WARPS 4
#define PX 8

__device__ __forceinline__ float signum32(float x) {   // util.go:118-130
    return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : ((__float_as_uint(x) >> 31) ? -1.0f : 1.0f));
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// two float32 FMAs in one instruction (FFMA2, sm_100): each half rounds exactly like fmaf
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ uint32_t pack_rgba8(float4 c) {   // fine.wgsl:305-323
    uint32_t r = (uint32_t)(clamp01(c.x) * 255.0f + 0.5f);
    uint32_t g = (uint32_t)(clamp01(c.y) * 255.0f + 0.5f);
    uint32_t b = (uint32_t)(clamp01(c.z) * 255.0f + 0.5f);
    uint32_t a = (uint32_t)(clamp01(c.w) * 255.0f + 0.5f);
    return r | (g << 8) | (b << 16) | (a << 24);
}
// (float)byte / 255.0f (fine.go:104-108) as a table: the IEEE division the reference does, done by the compiler; four
// of them per CmdColor were 8 % of fine's instructions on the 1 M-path scene (ncu r2b). The colour word is the same in
// every lane, so the look-ups are uniform constant loads.
__constant__ float U8_TO_UNIT[256] = {
    0.0f / 255.0f, 1.0f / 255.0f, 2.0f / 255.0f, 3.0f / 255.0f, 4.0f / 255.0f, 5.0f / 255.0f, 6.0f / 255.0f, 7.0f / 255.0f,
    8.0f / 255.0f, 9.0f / 255.0f, 10.0f / 255.0f, 11.0f / 255.0f, 12.0f / 255.0f, 13.0f / 255.0f, 14.0f / 255.0f, 15.0f / 255.0f,
    16.0f / 255.0f, 17.0f / 255.0f, 18.0f / 255.0f, 19.0f / 255.0f, 20.0f / 255.0f, 21.0f / 255.0f, 22.0f / 255.0f, 23.0f / 255.0f,
    24.0f / 255.0f, 25.0f / 255.0f, 26.0f / 255.0f, 27.0f / 255.0f, 28.0f / 255.0f, 29.0f / 255.0f, 30.0f / 255.0f, 31.0f / 255.0f,
    32.0f / 255.0f, 33.0f / 255.0f, 34.0f / 255.0f, 35.0f / 255.0f, 36.0f / 255.0f, 37.0f / 255.0f, 38.0f / 255.0f, 39.0f / 255.0f,
    40.0f / 255.0f, 41.0f / 255.0f, 42.0f / 255.0f, 43.0f / 255.0f, 44.0f / 255.0f, 45.0f / 255.0f, 46.0f / 255.0f, 47.0f / 255.0f,
    48.0f / 255.0f, 49.0f / 255.0f, 50.0f / 255.0f, 51.0f / 255.0f, 52.0f / 255.0f, 53.0f / 255.0f, 54.0f / 255.0f, 55.0f / 255.0f,
    56.0f / 255.0f, 57.0f / 255.0f, 58.0f / 255.0f, 59.0f / 255.0f, 60.0f / 255.0f, 61.0f / 255.0f, 62.0f / 255.0f, 63.0f / 255.0f,
    64.0f / 255.0f, 65.0f / 255.0f, 66.0f / 255.0f, 67.0f / 255.0f, 68.0f / 255.0f, 69.0f / 255.0f, 70.0f / 255.0f, 71.0f / 255.0f,
    72.0f / 255.0f, 73.0f / 255.0f, 74.0f / 255.0f, 75.0f / 255.0f, 76.0f / 255.0f, 77.0f / 255.0f, 78.0f / 255.0f, 79.0f / 255.0f,
    80.0f / 255.0f, 81.0f / 255.0f, 82.0f / 255.0f, 83.0f / 255.0f, 84.0f / 255.0f, 85.0f / 255.0f, 86.0f / 255.0f, 87.0f / 255.0f,
    88.0f / 255.0f, 89.0f / 255.0f, 90.0f / 255.0f, 91.0f / 255.0f, 92.0f / 255.0f, 93.0f / 255.0f, 94.0f / 255.0f, 95.0f / 255.0f,
    96.0f / 255.0f, 97.0f / 255.0f, 98.0f / 255.0f, 99.0f / 255.0f, 100.0f / 255.0f, 101.0f / 255.0f, 102.0f / 255.0f, 103.0f / 255.0f,
    104.0f / 255.0f, 105.0f / 255.0f, 106.0f / 255.0f, 107.0f / 255.0f, 108.0f / 255.0f, 109.0f / 255.0f, 110.0f / 255.0f, 111.0f / 255.0f,
    112.0f / 255.0f, 113.0f / 255.0f, 114.0f / 255.0f, 115.0f / 255.0f, 116.0f / 255.0f, 117.0f / 255.0f, 118.0f / 255.0f, 119.0f / 255.0f,
    120.0f / 255.0f, 121.0f / 255.0f, 122.0f / 255.0f, 123.0f / 255.0f, 124.0f / 255.0f, 125.0f / 255.0f, 126.0f / 255.0f, 127.0f / 255.0f,
    128.0f / 255.0f, 129.0f / 255.0f, 130.0f / 255.0f, 131.0f / 255.0f, 132.0f / 255.0f, 133.0f / 255.0f, 134.0f / 255.0f, 135.0f / 255.0f,
    136.0f / 255.0f, 137.0f / 255.0f, 138.0f / 255.0f, 139.0f / 255.0f, 140.0f / 255.0f, 141.0f / 255.0f, 142.0f / 255.0f, 143.0f / 255.0f,
    144.0f / 255.0f, 145.0f / 255.0f, 146.0f / 255.0f, 147.0f / 255.0f, 148.0f / 255.0f, 149.0f / 255.0f, 150.0f / 255.0f, 151.0f / 255.0f,
    152.0f / 255.0f, 153.0f / 255.0f, 154.0f / 255.0f, 155.0f / 255.0f, 156.0f / 255.0f, 157.0f / 255.0f, 158.0f / 255.0f, 159.0f / 255.0f,
    160.0f / 255.0f, 161.0f / 255.0f, 162.0f / 255.0f, 163.0f / 255.0f, 164.0f / 255.0f, 165.0f / 255.0f, 166.0f / 255.0f, 167.0f / 255.0f,
    168.0f / 255.0f, 169.0f / 255.0f, 170.0f / 255.0f, 171.0f / 255.0f, 172.0f / 255.0f, 173.0f / 255.0f, 174.0f / 255.0f, 175.0f / 255.0f,
    176.0f / 255.0f, 177.0f / 255.0f, 178.0f / 255.0f, 179.0f / 255.0f, 180.0f / 255.0f, 181.0f / 255.0f, 182.0f / 255.0f, 183.0f / 255.0f,
    184.0f / 255.0f, 185.0f / 255.0f, 186.0f / 255.0f, 187.0f / 255.0f, 188.0f / 255.0f, 189.0f / 255.0f, 190.0f / 255.0f, 191.0f / 255.0f,
    192.0f / 255.0f, 193.0f / 255.0f, 194.0f / 255.0f, 195.0f / 255.0f, 196.0f / 255.0f, 197.0f / 255.0f, 198.0f / 255.0f, 199.0f / 255.0f,
    200.0f / 255.0f, 201.0f / 255.0f, 202.0f / 255.0f, 203.0f / 255.0f, 204.0f / 255.0f, 205.0f / 255.0f, 206.0f / 255.0f, 207.0f / 255.0f,
    208.0f / 255.0f, 209.0f / 255.0f, 210.0f / 255.0f, 211.0f / 255.0f, 212.0f / 255.0f, 213.0f / 255.0f, 214.0f / 255.0f, 215.0f / 255.0f,
    216.0f / 255.0f, 217.0f / 255.0f, 218.0f / 255.0f, 219.0f / 255.0f, 220.0f / 255.0f, 221.0f / 255.0f, 222.0f / 255.0f, 223.0f / 255.0f,
    224.0f / 255.0f, 225.0f / 255.0f, 226.0f / 255.0f, 227.0f / 255.0f, 228.0f / 255.0f, 229.0f / 255.0f, 230.0f / 255.0f, 231.0f / 255.0f,
    232.0f / 255.0f, 233.0f / 255.0f, 234.0f / 255.0f, 235.0f / 255.0f, 236.0f / 255.0f, 237.0f / 255.0f, 238.0f / 255.0f, 239.0f / 255.0f,
    240.0f / 255.0f, 241.0f / 255.0f, 242.0f / 255.0f, 243.0f / 255.0f, 244.0f / 255.0f, 245.0f / 255.0f, 246.0f / 255.0f, 247.0f / 255.0f,
    248.0f / 255.0f, 249.0f / 255.0f, 250.0f / 255.0f, 251.0f / 255.0f, 252.0f / 255.0f, 253.0f / 255.0f, 254.0f / 255.0f, 255.0f / 255.0f,
};
__device__ __forceinline__ float4 unpack_rgba8(uint32_t c) {
    return make_float4(U8_TO_UNIT[c & 0xffu], U8_TO_UNIT[(c >> 8) & 0xffu], U8_TO_UNIT[(c >> 16) & 0xffu], U8_TO_UNIT[c >> 24]);
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
    if (n < 0) { const float k = l / (l - n); r = l + (r - l) * k; g = l + (g - l) * k; b = l + (b - l) * k; }
    if (x > 1) { const float k = (1 - l) / (x - l); r = l + (r - l) * k; g = l + (g - l) * k; b = l + (b - l) * k; }
}
__device__ __forceinline__ void set_lum(float& r, float& g, float& b, float l) { float d = l - lum3(r, g, b); r += d; g += d; b += d; clip_color(r, g, b); }
// hsl.go:53-88 without the pointer sort: min -> 0, max -> s, mid -> ((mid - min) * s) / (max - min); a component tied with
// the maximum takes s itself (the reference's quotient there is s up to one rounding)
__device__ __forceinline__ void set_sat(float& r, float& g, float& b, float s) {
    const float lo = min3f(r, g, b), hi = max3f(r, g, b);
    if (hi > lo) {
        const float d = hi - lo;
        r = r == hi ? s : ((r - lo) * s) / d;
        g = g == hi ? s : ((g - lo) * s) / d;
        b = b == hi ? s : ((b - lo) * s) / d;
    }
}
// out of line: one copy for the three channels (the kernel has to fit the instruction cache, see blend_mix4)
__device__ __noinline__ float sep_mix(uint32_t mix, float s, float d) {
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
// bg (blend) fg for a mix mode 1..15 (always composed SrcOver); both visible (sa > 0, da > 0)
__device__ __forceinline__ float4 blend_mix_px(uint32_t mix, float4 bg, float4 fg) {
    const float sa = fg.w, da = bg.w;
    const float isa = 1.0f / sa, ida = 1.0f / da;
    const float csr = fg.x * isa, csg = fg.y * isa, csb = fg.z * isa, cdr = bg.x * ida, cdg = bg.y * ida, cdb = bg.z * ida;
    float br, bgc, bb;
    if (mix >= 12) {
        // Hue 12: SetLum(SetSat(Cs, Sat(Cd)), Lum(Cd)); Saturation 13: SetLum(SetSat(Cd, Sat(Cs)), Lum(Cd));
        // Color 14: SetLum(Cs, Lum(Cd)); Luminosity 15: SetLum(Cd, Lum(Cs))  -- one copy of SetSat / SetLum for all four
        const bool from_src = mix == 12u || mix == 14u;
        br = from_src ? csr : cdr; bgc = from_src ? csg : cdg; bb = from_src ? csb : cdb;
        const float o_r = from_src ? cdr : csr, o_g = from_src ? cdg : csg, o_b = from_src ? cdb : csb;   // the other colour
        if (mix <= 13u) set_sat(br, bgc, bb, max3f(o_r, o_g, o_b) - min3f(o_r, o_g, o_b));
        set_lum(br, bgc, bb, mix == 15u ? lum3(csr, csg, csb) : lum3(cdr, cdg, cdb));
    } else {
        br = sep_mix(mix, csr, cdr); bgc = sep_mix(mix, csg, cdg); bb = sep_mix(mix, csb, cdb);
    }
    const float sada = sa * da;
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
// CmdEndClip (fine.go:140-180) for FOUR of the lane's pixels, held in shared memory: scr[i * 32] = the layer's pixel
// (premultiplied), replaced by the result; cov[i * 32] = the clip's coverage there; slot[i * 32] = the backdrop saved by
// CmdBeginClip. One rolled loop, one copy of the blend code, two calls per CmdEndClip -- with the eight pixels in
// registers the code was unrolled eight times and the kernel did not fit the 32 KB instruction cache of an SM (ncu r2b /
// r2f: sm__icc hit rate 69-74 %, 9-13 cycles of no_instruction stall per issue, the GPC's instruction cache 95 % busy).
// Scaling the source by the coverage (fine.go:152-160) equals out = D + cov (blend(S, D) - D) for every mode whose
// backdrop factor is 1 under a transparent source; the six compose modes that wipe their backdrop are blended at full
// strength and interpolated instead (a pixel the layer's clip does not cover stays).
__device__ __noinline__ void end_clip4(uint32_t blend, float alpha, float4* scr, const float* cov, const float4* slot) {
    const uint32_t mix = (blend >> 8) & 0xffu, compose = blend & 0xffu;
    const bool wipe = mix == 0u && (compose == 0u || compose == 1u || compose == 5u || compose == 6u || compose == 7u || compose == 10u);
    const bool mixed = mix != 0u && mix < 16u;
    const float4 k = COMPOSE_COEF[min(compose, 13u)];   // Normal / clip SrcOver (fine.go:168-179) is the row {1, 0, 1, -1}
    const bool plus = compose == 12u;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        float4 fg = scr[i * 32];
        const float4 bg = slot[i * 32];
        const float cv = cov[i * 32];
        const float scale = wipe ? alpha : cv * alpha;
        fg.x *= scale; fg.y *= scale; fg.z *= scale; fg.w *= scale;
        float4 o;
        if (mixed) {
            // trivial pixels first (transparent source -> backdrop, transparent backdrop -> source)
            o = fg;
            if (fg.w <= 0.0f) o = bg;
            else if (bg.w > 0.0f) o = blend_mix_px(mix, bg, fg);
        } else {
            const float fa = k.x + k.y * bg.w, fb = k.z + k.w * fg.w;
            o.x = fmaf(fb, bg.x, fa * fg.x); o.y = fmaf(fb, bg.y, fa * fg.y);
            o.z = fmaf(fb, bg.z, fa * fg.z); o.w = fmaf(fb, bg.w, fa * fg.w);
            if (plus) { o.x = fminf(o.x, 1.0f); o.y = fminf(o.y, 1.0f); o.z = fminf(o.z, 1.0f); o.w = fminf(o.w, 1.0f); }
            if (wipe) { o.x = bg.x + cv * (o.x - bg.x); o.y = bg.y + cv * (o.y - bg.y); o.z = bg.z + cv * (o.z - bg.z); o.w = bg.w + cv * (o.w - bg.w); }
        }
        scr[i * 32] = o;
    }
}

// ---------------------------------------------------------------- shared memory per warp, bulk copies
#define PTCL_CHUNK 128   // words per command-ring slot (512 B)
#define SEG_CHUNK 32     // segments per segment-ring slot (640 B)
// byte offsets inside a warp's slice of dynamic shared memory
#define SM_STACK 0       // blend stack level 0: float4 [PX][32]                                    4096
#define SM_PTCL 4096     // command ring: u32 [2][PTCL_CHUNK]                                        1024
#define SM_SEGS 5120     // segment ring: GGSegment [2][SEG_CHUNK]                                   1280
#define SM_DER 6400      // 1/dy of the segments of the chunk being evaluated: float [32] (384 reserved)      384
#define SM_D 6784        // coverage difference table: int [8][32] (then: RGBA8 image of tile B)     1152
#define SM_SCR 6400      // CmdEndClip scratch, four pixels per lane: float4 [4][32] (over DER and D) 2048
#define SM_IMG_A 8448    // RGBA8 image of the pair's left tile: u32 [16][16]                        1024
#define SM_BARS 9472     // 4 mbarriers                                                              32
#define SM_COV 9600      // CmdEndClip scratch: coverage of the four pixels, float [4][32]           512
#define FINE_SMEM_PER_WARP 10112

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

// A tile's command list, streamed through a two-slot ring.
struct PtclStream {
    const uint32_t* src;   // tile's PTCL in global memory (16-byte aligned)
    uint32_t* ring;        // [2][PTCL_CHUNK] in shared memory
    uint64_t* bars;        // [2]
    uint32_t len;          // words in this tile's list
    uint32_t loaded_end;   // first word index not yet available
    uint32_t parity;       // bit s = phase parity to wait for on slot s
    uint32_t issued;       // chunks of the current tile handed to the copy engine
    uint32_t limit;        // ensure(cmd) has nothing to do while cmd < limit
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
        limit = 0;
    }
    // Called once per loop iteration with the index of a command's first word; makes that command AND the one after it
    // readable (a coverage command -- CmdFill, 4 words -- is dispatched together with the command that uses it, at most
    // 3 words: 7 words). The ring holds the chunk the command starts in and the next one; the slot of the chunk BEHIND the
    // command start is refilled with the chunk after next. `limit` = first command index at which there is something to do
    // (one compare per iteration).
    __device__ __forceinline__ void ensure(uint32_t cmd) {
        if (cmd >= limit) ensure_slow(cmd);
    }
    __device__ __forceinline__ void ensure_slow(uint32_t cmd) {
        const uint32_t cur = cmd / PTCL_CHUNK;
        while (issued < cur + 2 && issued * PTCL_CHUNK < len) { __syncwarp(); issue(issued); }
        while (cmd + 7 >= loaded_end && loaded_end < issued * PTCL_CHUNK) {
            uint32_t slot = (loaded_end / PTCL_CHUNK) & 1u;
            mbar_wait(bars + slot, (parity >> slot) & 1u);
            parity ^= 1u << slot;
            loaded_end += PTCL_CHUNK;
        }
        const uint32_t a = issued * PTCL_CHUNK >= len ? 0xffffffffu : (issued - 1u) * PTCL_CHUNK;   // next refill
        const uint32_t b = loaded_end >= issued * PTCL_CHUNK ? 0xffffffffu : loaded_end - 7u;          // next wait
        limit = min(a, b);
    }
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return ring[i & (2 * PTCL_CHUNK - 1)]; }
    // The CmdFill that follows the command at word c (the user of the fill being evaluated: CmdColor / CmdGrad, 2 words, or
    // CmdEndClip, 3 words), if both have already arrived. Other shapes are simply not prefetched.
    __device__ __forceinline__ bool next_fill(uint32_t c, uint32_t* seg_ix, uint32_t* n) const {
        if (c + 6 >= min(loaded_end, len)) return false;
        const uint32_t t0 = word(c);
        if (t0 != GG_CMD_COLOR && t0 != GG_CMD_END_CLIP && t0 != GG_CMD_GRAD) return false;
        c += t0 == GG_CMD_END_CLIP ? 3u : 2u;
        if (word(c) != GG_CMD_FILL) return false;
        *n = word(c + 1) >> 1; *seg_ix = word(c + 2);
        return true;
    }
};

// PathSegment slices, streamed through a second two-slot ring in chunks of SEG_CHUNK segments. Chunks start at
// multiples of four segments (80 bytes) so that every copy is 16-byte aligned; the up to three segments before a
// slice's first one are skipped by the consumer. One chunk may be in flight ahead of the one being evaluated.
struct SegStream {
    const GGSegment* segs;
    float* ring;           // [2][SEG_CHUNK * 5]
    uint64_t* bars;        // [2]
    uint32_t issue_seq, wait_seq;   // chunks handed to the copy engine / consumed (slot = seq & 1, parity = (seq >> 1) & 1)
    uint32_t ahead_first;  // first segment of the newest chunk in flight (meaningful while issue_seq > wait_seq)
    uint32_t lane;
    __device__ __forceinline__ void issue(uint32_t first, uint32_t count) {   // count: segments, rounded up to a multiple of 4
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_load(ring + (issue_seq & 1u) * (SEG_CHUNK * 5), segs + first, count * (uint32_t)sizeof(GGSegment), bars + (issue_seq & 1u));
        }
        ahead_first = first;
        issue_seq++;
    }
    __device__ __forceinline__ const float* wait() {
        const uint32_t slot = wait_seq & 1u;
        mbar_wait(bars + slot, (wait_seq >> 1) & 1u);
        wait_seq++;
        return ring + slot * (SEG_CHUNK * 5);
    }
    __device__ __forceinline__ void drain() { while (wait_seq < issue_seq) wait(); }
};

// fine.go:233-275 for ONE (segment, row) pair: trapezoid areas of the columns the segment passes through in this row of
// the tile, and the constant winding step dy for everything to the right, recorded as differences in the row's table.
// The table is laid out for its readers -- lane (row, half) owns entries D[i * 32 + lane], i = column inside its 8-pixel
// half -- and written with shared-memory INTEGER atomics (native ATOMS.ADD, not the compare-and-swap loops of float
// atomics): 2^-20 fixed point, so the sums do not depend on the order the segments of a tile arrive in (path_tiling claims
// their slots with atomics) and a frame is reproducible bit for bit; the quantisation, 5e-7 per term, is the size of
// float32's own rounding at these magnitudes. (Windings beyond +-2047 would wrap.) Operation order of the per-pixel formula
// as in the reference (its numerator cancels for near-vertical segments; this TU is compiled with -fmad=false).
#define AREA_FIX 1048576.0f
__device__ __forceinline__ int d_slot(int row, int col) { return (col & 7) * 32 + row * 2 + (col >> 3); }   // col 0..15
__device__ __forceinline__ void seg_item(int* D, int row, float p0x, float p0y, float dx, float dys, float rc) {
    const float y = p0y - (float)row;
    const float y0 = clamp01(y);
    const float y1 = clamp01(y + dys);
    const float dy = y0 - y1;
    if (dy != 0.0f) {
        const float t0 = (y0 - y) * rc;
        const float t1 = (y1 - y) * rc;
        const float x0 = p0x + t0 * dx;
        const float x1 = p0x + t1 * dx;
        const float xmin0 = fminf(x0, x1);
        const float xmax0 = fmaxf(x0, x1);
        const int c0 = max((int)floorf(xmin0), 0);
        const int c1 = min(max((int)ceilf(xmax0), 0), GG_TILE_W);
        int vprev = 0;
        for (int col = c0; col < c1; col++) {
            const float i_f = (float)col;   // column in the tile, as fine.go:259
            const float xmin = fminf(xmin0 - i_f, 1.0f) - 1.0e-6f;
            const float xmax = xmax0 - i_f;
            const float b = fminf(xmax, 1.0f);
            const float c = fmaxf(b, 0.0f);
            const float d = fmaxf(xmin, 0.0f);
            // the quotient may be 2 ulp off the reference's IEEE division: far below what survives the 8-bit quantisation
            const float a = __fdividef(b + 0.5f * (d * d - c * c) - xmin, xmax - xmin);
            const int v = __float2int_rn(a * dy * AREA_FIX);
            atomicAdd(&D[d_slot(row, col)], v - vprev);
            vprev = v;
        }
        if (c1 < GG_TILE_W) atomicAdd(&D[d_slot(row, c1)], __float2int_rn(dy * AREA_FIX) - vprev);
    }
}

// Area of one CmdFill (fine.go:219-276 fillPath) for the 8 pixels of this lane.
// The segments of a chunk are first looked at in parallel (lane = segment: rows crossed, 1/dy); the (segment, row) pairs
// of the chunk -- the unit of work of fillPath's two loops -- are then dealt to the lanes 32 at a time through a prefix
// sum of the row counts, whatever the shape of the slice: one tall edge (16 pairs of one segment) and a dozen short
// segments of a small path both fill one pass.
__device__ __forceinline__ void fill_area(float* area, int* D, float* der, SegStream& ss, const PtclStream& ps, uint32_t next_cmd,
                                          uint32_t seg_ix, uint32_t n, float backdrop, uint32_t lane) {
    const uint32_t row = lane >> 1;
    const float yi = (float)row;
    __syncwarp();   // the table and the derived values share their bytes with CmdEndClip's scratch and tile B's image
#pragma unroll
    for (int i = 0; i < PX; i++) D[i * 32 + lane] = 0;
    int base = 0;   // the y_edge terms of this row (fine.go:244: the same for every pixel of the row)
    const uint32_t lead = seg_ix & 3u, first = seg_ix - lead, total = lead + n;
    for (uint32_t c0s = 0; c0s < total; c0s += SEG_CHUNK) {
        const uint32_t cnt = min((uint32_t)SEG_CHUNK, total - c0s);
        // this chunk: already in flight if the previous fill (or chunk) found it by looking ahead
        if (!(ss.issue_seq > ss.wait_seq && ss.ahead_first == first + c0s)) {
            ss.drain();
            __syncwarp();
            ss.issue(first + c0s, (cnt + 3u) & ~3u);
        }
        // next chunk: of this slice, or the first one of the next CmdFill in the command list
        if (ss.issue_seq == ss.wait_seq + 1) {
            uint32_t nfirst = first + c0s + SEG_CHUNK, ncnt = total - c0s - SEG_CHUNK;
            bool more = c0s + SEG_CHUNK < total;
            if (!more) {
                uint32_t nseg_ix, nn;
                if (ps.next_fill(next_cmd, &nseg_ix, &nn)) { more = true; nfirst = nseg_ix & ~3u; ncnt = (nseg_ix & 3u) + nn; }
            }
            if (more) {
                __syncwarp();
                ss.issue(nfirst, (min((uint32_t)SEG_CHUNK, ncnt) + 3u) & ~3u);
            }
        }
        const float* sg = ss.wait();
        // ---- lane = segment: rows crossed, derived values
        const uint32_t k = c0s + lane;
        const bool valid = k >= lead && k < total && lane < cnt;
        uint32_t r0 = 0, nrows = 0;
        bool edge = false;
        float ye = 0.0f, sgn = 0.0f;
        if (valid) {
            const float p0x = sg[lane * 5 + 0], p0y = sg[lane * 5 + 1], p1x = sg[lane * 5 + 2], p1y = sg[lane * 5 + 3];
            ye = sg[lane * 5 + 4];
            const float dys = p1y - p0y;
            der[lane] = __frcp_rn(dys);
            if (dys != 0.0f) {
                const float ymin = fminf(p0y, p1y), ymax = fmaxf(p0y, p1y);
                r0 = (uint32_t)max(0, min(15, (int)floorf(ymin)));
                nrows = (uint32_t)max((int)r0 + 1, min(16, (int)ceilf(ymax))) - r0;
            }
            edge = ye < 16.0f;   // touches the tile's left edge: winding step for the rows from y_edge down (fine.go:244)
            sgn = signum32(p1x - p0x);
        }
        uint32_t em = __ballot_sync(0xffffffffu, edge);
        while (em) {
            const int j = __ffs((int)em) - 1;
            em &= em - 1u;
            const float yej = __shfl_sync(0xffffffffu, ye, j), sj = __shfl_sync(0xffffffffu, sgn, j);
            base += __float2int_rn(sj * clamp01(yi - yej + 1.0f) * AREA_FIX);
        }
        // ---- inclusive prefix sum of the row counts: pair t of the chunk belongs to the first segment with incl > t
        uint32_t incl = nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
        const uint32_t n_pairs = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t start = (incl - nrows) | (r0 << 16);   // first pair of my segment | its first row
        __syncwarp();   // der[] written above, the table zeroed: visible to every lane
        for (uint32_t t = lane; t - lane < n_pairs; t += 32) {
            uint32_t j = 0;
#pragma unroll
            for (uint32_t s2 = 16; s2; s2 >>= 1) { if (__shfl_sync(0xffffffffu, incl, j + s2 - 1u) <= t) j += s2; }
            const uint32_t st = __shfl_sync(0xffffffffu, start, j & 31u);
            if (t < n_pairs) {
                const float p0x = sg[j * 5 + 0], p0y = sg[j * 5 + 1];
                seg_item(D, (int)((st >> 16) + t - (st & 0xffffu)), p0x, p0y, sg[j * 5 + 2] - p0x, sg[j * 5 + 3] - p0y, der[j]);
            }
        }
        __syncwarp();
    }
    int run = 0;
    int dv[PX];
#pragma unroll
    for (int i = 0; i < PX; i++) { run += D[i * 32 + lane]; dv[i] = run; }
    const int left = __shfl_sync(0xffffffffu, run, lane & ~1u);   // the row's left half carries into its right half
    base += (lane & 1u) ? left : 0;
#pragma unroll
    for (int i = 0; i < PX; i++) area[i] = backdrop + (float)(dv[i] + base) * (1.0f / AREA_FIX);
}

// VelloAccelerator.compositeOver (vello_accelerator.go:388-442) for one pixel: the scene was rasterised on transparent,
// s = its premultiplied RGBA8 pixel, d = the target's; bytes, (d * inv + 127) / 255, uint8 wrap-around as in Go.
__device__ __forceinline__ uint32_t composite_over_u8(uint32_t s, uint32_t d) {
    const uint32_t sa = s >> 24;
    if (sa == 0u) return d;
    if (sa == 255u) return s;
    const uint32_t inv = 255u - sa;
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t sc = (s >> (8 * k)) & 0xffu, dc = (d >> (8 * k)) & 0xffu;
        o |= ((sc + ((dc * inv + 127u) / 255u)) & 0xffu) << (8 * k);
    }
    return o;
}

// TagImage at a pixel centre (scene/renderer.go:1093-1243 blitImageToTile + blitBilinearPixel / blitNearestPixel), float32
// in the reference's order of operations: map the centre back into the image (the record holds the inverse affine), move
// by half a texel, skip unless the 2 x 2 neighbourhood touches the image; on a texel centre take that texel (and only if
// it is inside), otherwise blend four clamp-to-edge texels in premultiplied space. Returns the premultiplied colour in
// [0, 1]; *cov = 1 where the reference draws, 0 where it leaves the pixel alone. Out of line like grad_color.
__device__ __noinline__ float4 image_color(const uint32_t* __restrict__ gtab, const uint32_t* __restrict__ g, float fx, float fy, int px, int py, float* cov) {
    *cov = 0.0f;
    const float4 none = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (px < (int)g[11] || py < (int)g[12] || px > (int)g[13] || py > (int)g[14]) return none;   // the box the reference iterates over
    const int w = (int)g[1], h = (int)g[2];
    const uint32_t* tex = gtab + g[3];
    float sx = __uint_as_float(g[5]) * fx + __uint_as_float(g[6]) * fy + __uint_as_float(g[7]);
    float sy = __uint_as_float(g[8]) * fx + __uint_as_float(g[9]) * fy + __uint_as_float(g[10]);
    sx -= 0.5f; sy -= 0.5f;
    const float flx = floorf(sx), fly = floorf(sy);
    if (!(fabsf(flx) < 1.0e9f) || !(fabsf(fly) < 1.0e9f)) return none;
    const int ix0 = (int)flx, iy0 = (int)fly;
    if (ix0 + 1 < 0 || iy0 + 1 < 0 || ix0 >= w || iy0 >= h) return none;
    const float wx = sx - flx, wy = sy - fly;
    if (wx == 0.0f && wy == 0.0f) {
        if (ix0 < 0 || iy0 < 0) return none;
        const uint32_t t = tex[iy0 * w + ix0];
        if ((t >> 24) == 0u) return none;
        *cov = 1.0f;
        return unpack_rgba8(t);
    }
    const int cx0 = min(max(ix0, 0), w - 1), cx1 = min(max(ix0 + 1, 0), w - 1), cy0 = min(max(iy0, 0), h - 1), cy1 = min(max(iy0 + 1, 0), h - 1);
    uint32_t t00 = tex[cy0 * w + cx0], t10 = tex[cy0 * w + cx1], t01 = tex[cy1 * w + cx0], t11 = tex[cy1 * w + cx1];
    if ((t00 >> 24) == 0u) t00 = 0u;   // fetchPremul: a texel without alpha counts as (0, 0, 0, 0)
    if ((t10 >> 24) == 0u) t10 = 0u;
    if ((t01 >> 24) == 0u) t01 = 0u;
    if ((t11 >> 24) == 0u) t11 = 0u;
    const float ifx = 1.0f - wx, ify = 1.0f - wy;
    const float w00 = ifx * ify, w10 = wx * ify, w01 = ifx * wy, w11 = wx * wy;
    float s[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float p00 = (float)((t00 >> (8 * k)) & 0xffu), p10 = (float)((t10 >> (8 * k)) & 0xffu);
        const float p01 = (float)((t01 >> (8 * k)) & 0xffu), p11 = (float)((t11 >> (8 * k)) & 0xffu);
        s[k] = p00 * w00 + p10 * w10 + p01 * w01 + p11 * w11;
    }
    if (s[3] < 0.5f / 255.0f) return none;
    *cov = 1.0f;
    return make_float4(s[0] / 255.0f, s[1] / 255.0f, s[2] / 255.0f, s[3] / 255.0f);
}

// gg's brush colour at a pixel centre (gradient_linear.go:52-66, gradient_radial.go computeTSimple, gradient.go:42-131), in gg's
// own arithmetic: float64 geometry and stop search, float32 interpolation in linear light, sRGB on the way out; only the last
// step differs from the Go code (powf instead of float64 math.Pow: < 1e-6 on a colour in [0, 1]). The stops carry their
// linear-light values from the host. Out of line: gradients are rare commands and this is a lot of code.
// Returns the premultiplied colour.
__device__ __noinline__ float4 grad_color(const uint32_t* __restrict__ gtab, const uint32_t* __restrict__ g, float fx, float fy) {
    const uint32_t kind = g[0], extend = g[1], n = g[2];
    const float* st = reinterpret_cast<const float*>(gtab + g[3]);   // 8 floats per stop: offset, r g b a, linear-light r g b
    if (n == 0u) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const double g0 = (double)__uint_as_float(g[5]), g1 = (double)__uint_as_float(g[6]);
    const double g2 = (double)__uint_as_float(g[7]), g3 = (double)__uint_as_float(g[8]);
    const double x = (double)fx, y = (double)fy;
    double t = 0.0;
    bool first_only = n == 1u;
    if (kind == 0u) {
        const double dx = g2 - g0, dy = g3 - g1, l2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (l2 == 0.0) first_only = true;
        else t = __ddiv_rn(__dadd_rn(__dmul_rn(x - g0, dx), __dmul_rn(y - g1, dy)), l2);
    } else if (kind == 1u) {
        const double dx = x - g0, dy = y - g1, rd = g3 - g2;
        if (rd == 0.0) first_only = true;
        else t = __ddiv_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) - g2, rd);
    } else if (kind == 4u) {   // sweep (gradient_sweep.go:79-150): the angle from the centre, relative to the start angle, over the sweep
        const double dx = x - g0, dy = y - g1, sweep = g3 - g2;
        if (dx == 0.0 && dy == 0.0) first_only = true;
        else if (sweep != 0.0) {
            const double two_pi = 6.283185307179586;
            double rel = atan2(dy, dx) - g2;
            if (sweep > 0.0) { for (int k = 0; k < 64 && rel < 0.0; k++) rel += two_pi; for (int k = 0; k < 64 && rel >= two_pi; k++) rel -= two_pi; }
            else { for (int k = 0; k < 64 && rel > 0.0; k++) rel -= two_pi; for (int k = 0; k < 64 && rel <= -two_pi; k++) rel += two_pi; }
            t = __ddiv_rn(rel, sweep);
        }
    } else {                   // radial with the focus off the centre (gradient_radial.go:131-196): ray from the focus against the end circle
        const double fx0 = (double)__uint_as_float(g[9]), fy0 = (double)__uint_as_float(g[10]);
        if (g3 - g2 == 0.0) first_only = true;
        else {
            const double dx = x - fx0, dy = y - fy0, fx = g0 - fx0, fy = g1 - fy0;
            const double a = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const double b = __dmul_rn(-2.0, __dadd_rn(__dmul_rn(dx, fx), __dmul_rn(dy, fy)));
            const double c = __dadd_rn(__dmul_rn(fx, fx), __dmul_rn(fy, fy)) - __dmul_rn(g3, g3);
            if (a != 0.0) {
                const double disc = __dmul_rn(b, b) - __dmul_rn(__dmul_rn(4.0, a), c);
                if (disc < 0.0) t = 1.0;
                else {
                    const double sq = __dsqrt_rn(disc), t1 = __ddiv_rn(-b - sq, __dmul_rn(2.0, a)), t2 = __ddiv_rn(-b + sq, __dmul_rn(2.0, a));
                    double tt = 0.0;
                    bool hit = true;
                    if (t1 > 0.0 && t2 > 0.0) tt = fmin(t1, t2); else if (t1 > 0.0) tt = t1; else if (t2 > 0.0) tt = t2; else hit = false;
                    if (hit) {
                        const double pd = __dsqrt_rn(a), idist = __dmul_rn(tt, pd);
                        if (idist != 0.0) t = __ddiv_rn(pd, idist);
                    }
                }
            }
        }
    }
    const float* a = nullptr;
    uint32_t i = 0;
    if (first_only) a = st;
    else {
        if (extend == 1u) { t -= floor(t); if (t < 0.0) t += 1.0; }
        else if (extend == 2u) { t = fabs(t); const double per = floor(t); t -= per; if (((long long)per) & 1ll) t = 1.0 - t; }
        else t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        while (i < n && !((double)st[8u * i] >= t)) i++;   // sort.Search: first stop with offset >= t
        if (i == 0u) a = st;
        else if (i >= n) a = st + 8u * (n - 1u);
        else if (st[8u * i] == st[8u * (i - 1u)]) a = st + 8u * (i - 1u);
    }
    float r, gg, b, al;
    if (a) { r = a[1]; gg = a[2]; b = a[3]; al = a[4]; }
    else {
        const float* s1 = st + 8u * (i - 1u);
        const float* s2 = st + 8u * i;
        const float lt = (float)__ddiv_rn(t - (double)s1[0], (double)s2[0] - (double)s1[0]);
        float c[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float l = s1[5 + k] + lt * (s2[5 + k] - s1[5 + k]);
            c[k] = l <= 0.0031308f ? l * 12.92f : 1.055f * powf(l, 1.0f / 2.4f) - 0.055f;
        }
        r = c[0]; gg = c[1]; b = c[2];
        al = s1[4] + lt * (s2[4] - s1[4]);
    }
    return make_float4(r * al, gg * al, b * al, al);
}

// CmdGrad for FOUR of the lane's pixels held in shared memory (scr[i * 32], coverage of the fill in cov[i * 32]), starting at
// pixel (px0, py): the brush's premultiplied colour and coverage at each pixel centre, source-over.
//   kind 0 / 1 / 4 / 5  gradients: gg's ColorAt (grad_color), coverage = the fill's area
//   kind 2  TagFillRoundRect the way gg's CPU renderer draws it (scene/renderer.go:986-1043, scene/shape.go:246-274): coverage =
//           Hermite smoothstep over +-0.7 px of the signed distance to the rounded rectangle; the path only binned the tiles
//   kind 3  TagImage (image_color): its own coverage too
__device__ __noinline__ void brush4(const uint32_t* __restrict__ gtab, const uint32_t* __restrict__ g, float4* scr, const float* cov, int px0, int py) {
    const uint32_t kind = g[0];
    const float fy = (float)py + 0.5f;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const int ipx = px0 + i;
        const float fx = (float)ipx + 0.5f;
        float cv = cov[i * 32];
        float4 c;
        if (kind == 2u) {
            const float gcx = __uint_as_float(g[5]), gcy = __uint_as_float(g[6]);
            const float hw = __uint_as_float(g[7]), hh = __uint_as_float(g[8]), rad = __uint_as_float(g[9]);
            c = unpack_rgba8(g[10]);
            const float ddx = fabsf(fx - gcx) - hw + rad, ddy = fabsf(fy - gcy) - hh + rad;
            const float ox = fmaxf(ddx, 0.0f), oy = fmaxf(ddy, 0.0f);
            const float dist = sqrtf(ox * ox + oy * oy) + fminf(fmaxf(ddx, ddy), 0.0f) - rad;
            if (dist >= 0.7f) cv = 0.0f;
            else if (dist <= -0.7f) cv = 1.0f;
            else { const float tt = (dist + 0.7f) / 1.4f; cv = 1.0f - (tt * tt * (3.0f - 2.0f * tt)); }
        } else if (kind == 3u) {
            c = image_color(gtab, g, fx, fy, ipx, py, &cv);
        } else {
            c = grad_color(gtab, g, fx, fy);
        }
        float4 d = scr[i * 32];
        d.x = fmaf(cv, fmaf(-c.w, d.x, c.x), d.x);
        d.y = fmaf(cv, fmaf(-c.w, d.y, c.y), d.y);
        d.z = fmaf(cv, fmaf(-c.w, d.z, c.z), d.z);
        d.w = fmaf(cv, fmaf(-c.w, d.w, c.w), d.w);
        scr[i * 32] = d;
    }
}

__global__ void __launch_bounds__(FINE_WARPS * 32, 4) fine_kernel(GGConfig cfg, const uint32_t* __restrict__ ptcl_off, const uint32_t* __restrict__ ptcl_len,
                                                               const uint32_t* __restrict__ ptcl, const uint32_t* __restrict__ restart_pt,
                                                               const GGSegment* __restrict__ segments, const uint32_t* __restrict__ spill_off,
                                                               float4* spill, GGBump* bump, uint8_t* dst, size_t stride, GGFineRange rg, uint32_t part, GGFineMirrors mir,
                                                               const uint32_t* __restrict__ gtab) {
    // a stage overflowed its buffer: PTCL / segments are incomplete, the host re-runs the pass with larger buffers
    if (bump->failed || bump->hits > cfg.hits_cap || bump->ptcl_words > cfg.ptcl_cap || bump->segments > cfg.segments_cap) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t row = lane >> 1;
    const uint32_t xb = (lane & 1u) * PX;
    extern __shared__ __align__(128) unsigned char fine_smem[];
    unsigned char* wsm = fine_smem + (size_t)(threadIdx.x >> 5) * FINE_SMEM_PER_WARP;
    float4 (*sstk)[PX][32] = reinterpret_cast<float4 (*)[PX][32]>(wsm + SM_STACK);
    int* D = reinterpret_cast<int*>(wsm + SM_D);
    float* der = reinterpret_cast<float*>(wsm + SM_DER);
    uint32_t* img_a = reinterpret_cast<uint32_t*>(wsm + SM_IMG_A);
    uint32_t* img_b = reinterpret_cast<uint32_t*>(wsm + SM_D);   // the difference table is idle between tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + SM_BARS);
    PtclStream ps;
    ps.ring = reinterpret_cast<uint32_t*>(wsm + SM_PTCL); ps.bars = bars; ps.parity = 0; ps.issued = 0; ps.loaded_end = 0; ps.limit = 0; ps.lane = lane;
    ps.src = ptcl; ps.len = 0;
    SegStream ss;
    ss.segs = segments; ss.ring = reinterpret_cast<float*>(wsm + SM_SEGS); ss.bars = bars + 2; ss.issue_seq = 0; ss.wait_seq = 0; ss.ahead_first = 0; ss.lane = lane;
    if (lane == 0) {
        mbar_init(bars + 0, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const bool over = (cfg.flags & GG_FLAG_BG_FROM_DST) != 0u;    // composite the scene over the target's pixels at the end
    const bool f32_out = (cfg.flags & GG_FLAG_TARGET_F32) != 0u;  // premultiplied float4 per pixel instead of RGBA8
    const float4 bgc = over ? make_float4(0, 0, 0, 0) : make_float4(cfg.bg[0], cfg.bg[1], cfg.bg[2], cfg.bg[3]);
    const uint32_t npx = rg.px1 - rg.px0;                         // tile pairs per row of this launch
    const uint32_t n_pairs = npx * (rg.row1 - rg.row0);
    // Work is handed out from a shared cursor: costs vary by orders of magnitude (restart points make most tiles of a scene
    // with opaque content trivial; a static stride left warps idle behind the heavy ones). First the tiles coarse listed as
    // heavy, ONE per warp -- the frame ends when its slowest warp does, so the long lists start first and are not chained
    // two to a warp --, then the pairs, skipping the heavy tiles in them. (If an eighth of the tiles are heavy there is no
    // tail to speak of: everything goes by pairs.)
    const uint32_t band_tiles = (cfg.band_y1 - cfg.band_y0) * cfg.width_in_tiles;
    const uint32_t* heavy_list = spill_off + band_tiles;
    const uint32_t n_heavy = bump->heavy * 8u <= band_tiles ? bump->heavy : 0u;
    for (;;) {
        uint32_t P = 0;
        if (lane == 0) P = atomicAdd(&bump->fine_cursor[part], 1u);
        P = __shfl_sync(0xffffffffu, P, 0);
        uint32_t trow, pcol, only = 2u;   // only: the tile of the pair to render (2 = both)
        uint32_t psel = 2u;               // heavy tiles: which four of a lane's eight pixels this warp is responsible for (2 = all)
        if (P < 2u * n_heavy) {
            // A listed tile goes to TWO warps: both replay the whole list, each blends the layers (the bulk of such a list) for
            // four of every lane's eight pixels only and stores those -- the frame ends when the slowest warp does.
            const uint32_t T = heavy_list[P >> 1];
            psel = P & 1u;
            trow = T / cfg.width_in_tiles;
            const uint32_t tx = T - trow * cfg.width_in_tiles;
            pcol = tx >> 1; only = tx & 1u;
            if (trow < rg.row0 || trow >= rg.row1 || pcol < rg.px0 || pcol >= rg.px1) continue;   // another launch's tile
        } else {
            P -= 2u * n_heavy;
            if (P >= n_pairs) break;
            trow = rg.row0 + P / npx;                  // tile row, relative to the band
            pcol = rg.px0 + P % npx;
        }
        const uint32_t py = (trow + cfg.band_y0) * GG_TILE_H + row;
        const bool row_in = py < cfg.height;
        uint32_t done = 0;   // bit h: half h was rendered by this warp
        for (uint32_t half = 0; half < 2; half++) {
            const uint32_t tx = pcol * 2 + half;
            if (tx >= cfg.width_in_tiles) break;
            if (only != 2u && only != half) continue;
            const uint32_t T = trow * cfg.width_in_tiles + tx;
            // coarse found the last command of this tile that overwrites every pixel whatever came before (an opaque
            // solid colour or a backdrop-wiping layer at clip depth 0): start right after it, from that colour
            const uint32_t restart = restart_pt[2 * T];
            const uint32_t tlen = ptcl_len[T];
            if (only == 2u && n_heavy && tlen - max(restart, 1u) > GG_FINE_HEAVY_WORDS) continue;   // listed: rendered on its own
            done |= 1u << half;
            float4 rgba[PX];
            float area[PX];
            {
                const float4 c0 = restart ? unpack_rgba8(restart_pt[2 * T + 1]) : bgc;
#pragma unroll
                for (int i = 0; i < PX; i++) { rgba[i] = c0; area[i] = 0.0f; }
            }
            uint32_t clip_depth = 0;
            uint32_t cmd = restart ? restart : 1u;   // word 0 = blend offset (ptcl.go:98)
            ps.begin(ptcl + ptcl_off[T], tlen, cmd / PTCL_CHUNK);
            const uint32_t sp_off = spill_off[T];
            for (;;) {
                ps.ensure(cmd);
                uint32_t tag = ps.word(cmd);
                // a coverage command (CmdFill / CmdSolid) is followed by the command that uses it: fall through to it
                if (tag == GG_CMD_FILL) {
                    const uint32_t packed = ps.word(cmd + 1), seg_ix = ps.word(cmd + 2);
                    const float backdrop = (float)(int32_t)ps.word(cmd + 3);
                    cmd += 4;
                    fill_area(area, D, der, ss, ps, cmd, seg_ix, packed >> 1, backdrop, lane);
                    if (packed & 1u) {
#pragma unroll
                        for (int i = 0; i < PX; i++) { const float a = area[i]; area[i] = fabsf(a - 2.0f * roundf(0.5f * a)); }   // fine.go:281
                    } else {
#pragma unroll
                        for (int i = 0; i < PX; i++) area[i] = fminf(fabsf(area[i]), 1.0f);                                      // fine.go:286
                    }
                    tag = ps.word(cmd);
                } else if (tag == GG_CMD_SOLID) {
                    cmd += 1;
#pragma unroll
                    for (int i = 0; i < PX; i++) area[i] = 1.0f;
                    tag = ps.word(cmd);
                }
                if (tag == GG_CMD_COLOR) {
                    const float4 c = unpack_rgba8(ps.word(cmd + 1));
                    cmd += 2;
                    // fine.go:104-123: rgba * (1 - c.a * cov) + c * cov, two FMAs per channel, two channels per instruction
                    const uint64_t nw2 = pack2(-c.w, -c.w), cxy = pack2(c.x, c.y), czw = pack2(c.z, c.w);
#pragma unroll
                    for (int i = 0; i < PX; i++) {
                        const uint64_t cov2 = pack2(area[i], area[i]);
                        const uint64_t xy = pack2(rgba[i].x, rgba[i].y), zw = pack2(rgba[i].z, rgba[i].w);
                        unpack2(fma2(cov2, fma2(nw2, xy, cxy), xy), rgba[i].x, rgba[i].y);
                        unpack2(fma2(cov2, fma2(nw2, zw, czw), zw), rgba[i].z, rgba[i].w);
                    }
                } else if (tag == GG_CMD_GRAD) {
                    // A brush evaluated per pixel (gradient, SDF round rect, image): out of line, over the same shared scratch
                    // as CmdEndClip -- rare commands with a lot of code, kept out of the command loop's instruction footprint.
                    const uint32_t* g = gtab + 16u * ps.word(cmd + 1);
                    cmd += 2;
                    float4* scr = reinterpret_cast<float4*>(wsm + SM_SCR) + lane;
                    float* cvs = reinterpret_cast<float*>(wsm + SM_COV) + lane;
                    __syncwarp();   // the scratch lies over the coverage table of the fill just evaluated
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        if (psel != 2u && psel != (uint32_t)hf) continue;
#pragma unroll
                        for (int i = 0; i < 4; i++) { scr[i * 32] = rgba[hf * 4 + i]; cvs[i * 32] = area[hf * 4 + i]; }
                        brush4(gtab, g, scr, cvs, (int)(tx * GG_TILE_W + xb) + hf * 4, (int)py);
#pragma unroll
                        for (int i = 0; i < 4; i++) rgba[hf * 4 + i] = scr[i * 32];
                    }
                } else if (tag == GG_CMD_BEGIN_CLIP) {   // fine.go:125-138
                    cmd += 1;
                    // Blend stack (fine.go:58-62 keeps 4 levels "in registers" and spills deeper ones; where a level lives is
                    // invisible in the output): level 0 in shared memory ([pixel][lane]: conflict-free 128-bit accesses),
                    // deeper levels in the global spill buffer coarse sized for this tile.
                    // Both as [pixel][lane]: a warp-wide access is 512 contiguous bytes either way.
                    float4* slot = clip_depth < GG_BLEND_STACK_SPLIT ? &sstk[clip_depth][0][lane]
                                                                     : spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane);
                    if (clip_depth < GG_BLEND_STACK_SPLIT || sp_off != 0xffffffffu) {
#pragma unroll
                        for (int i = 0; i < PX; i++) {   // (a heavy tile's two warps share its spill slots: each saves its own four pixels)
                            if (psel == 2u || psel == (uint32_t)(i >> 2)) slot[i * 32] = rgba[i];
                        }
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
                    const float4* slot = clip_depth < GG_BLEND_STACK_SPLIT ? &sstk[clip_depth][0][lane]
                                                                           : spill + ((size_t)(sp_off + clip_depth - GG_BLEND_STACK_SPLIT) * 256 + lane);
                    float4* scr = reinterpret_cast<float4*>(wsm + SM_SCR) + lane;
                    float* cvs = reinterpret_cast<float*>(wsm + SM_COV) + lane;
                    __syncwarp();   // the scratch lies over the coverage table of the fill just evaluated
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
                        if (psel != 2u && psel != (uint32_t)hf) continue;   // the other warp of this heavy tile does those
#pragma unroll
                        for (int i = 0; i < 4; i++) { scr[i * 32] = rgba[hf * 4 + i]; cvs[i * 32] = area[hf * 4 + i]; }
                        end_clip4(blend, alpha, scr, cvs, slot + hf * 4 * 32);
#pragma unroll
                        for (int i = 0; i < 4; i++) rgba[hf * 4 + i] = scr[i * 32];
                    }
                } else if (tag != GG_CMD_FILL && tag != GG_CMD_SOLID) {
                    break;   // CmdEnd, or an unknown command: stop (fine.go:182-185)
                }
            }
            ss.drain();   // a slice fetched ahead for a fill that never came (cannot happen with a well-formed list)
            if (f32_out) {
                // premultiplied float4 per pixel: 8 x 16 bytes, one full 128-byte line per lane
                const uint32_t px = tx * GG_TILE_W + xb;
                if (row_in) {
                    float4* o = reinterpret_cast<float4*>(dst + (size_t)(py - cfg.band_y0 * GG_TILE_H) * stride + (size_t)px * 16);
#pragma unroll
                    for (int i = 0; i < PX; i++)
                        if (px + i < cfg.width && (psel == 2u || psel == (uint32_t)(i >> 2))) o[i] = make_float4(clamp01(rgba[i].x), clamp01(rgba[i].y), clamp01(rgba[i].z), clamp01(rgba[i].w));
                }
            } else {
                // RGBA8 image of the tile in shared memory: rows of 64 bytes (tile B's lies over the coverage table)
                __syncwarp();
                uint32_t o[PX];
#pragma unroll
                for (int i = 0; i < PX; i++) o[i] = pack_rgba8(rgba[i]);
                uint4* im = reinterpret_cast<uint4*>((half ? img_b : img_a) + row * 16 + xb);
                im[0] = make_uint4(o[0], o[1], o[2], o[3]);
                im[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        if (f32_out) continue;
        __syncwarp();
        // ---- store the pair: a lane takes 16 bytes, eight lanes one 128-byte row of the pair, the warp four rows per pass.
        //      Multi-GPU: the same 128-byte lines also go straight into the other devices' frames (peer memory over NVLink, or
        //      one multimem store that the NVSwitch replicates to every device of the group) while the rest of the band is
        //      still being rasterised -- the all-gather of bands, fused into the kernel that makes them.
#pragma unroll 1
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t r = k * 4 + (lane >> 3), chunk = lane & 7u, half = chunk >> 2, cq = chunk & 3u;
            if (!((done >> half) & 1u)) continue;
            if (psel != 2u && (cq & 1u) != psel) continue;   // a heavy tile's other warp stores the other four pixels of every lane
            uint4 v = *reinterpret_cast<const uint4*>((half ? img_b : img_a) + r * 16 + cq * 4);
            const uint32_t x = (pcol * 2 + half) * GG_TILE_W + cq * 4;
            const uint32_t y = (trow + cfg.band_y0) * GG_TILE_H + r;
            if (y >= cfg.height || x >= cfg.width) continue;
            const size_t off = (size_t)(y - cfg.band_y0 * GG_TILE_H) * stride + (size_t)x * 4;
            uint8_t* out = dst + off;
            const bool full = x + 4 <= cfg.width;
            if (full && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0)) {
                if (over) {
                    const uint4 d = *reinterpret_cast<const uint4*>(out);
                    v.x = composite_over_u8(v.x, d.x); v.y = composite_over_u8(v.y, d.y); v.z = composite_over_u8(v.z, d.z); v.w = composite_over_u8(v.w, d.w);
                }
                *reinterpret_cast<uint4*>(out) = v;
            } else {
                uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; i++) if (x + i < cfg.width) {
                    uint32_t* po = reinterpret_cast<uint32_t*>(out) + i;
                    if (over) vv[i] = composite_over_u8(vv[i], *po);
                    *po = vv[i];
                }
                v = make_uint4(vv[0], vv[1], vv[2], vv[3]);
            }
            if (mir.multicast) {
                uint8_t* m = mir.p[0] + off;
                if (full && ((reinterpret_cast<uintptr_t>(m) & 15u) == 0)) {
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(m), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
                } else {
                    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) if (x + i < cfg.width) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(m + 4 * i), "f"(__uint_as_float(vv[i])) : "memory");
                }
            } else {
                for (uint32_t q = 0; q < mir.n; q++) {
                    uint8_t* m = mir.p[q] + off;
                    if (full && ((reinterpret_cast<uintptr_t>(m) & 15u) == 0)) {
                        *reinterpret_cast<uint4*>(m) = v;
                    } else {
                        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) if (x + i < cfg.width) reinterpret_cast<uint32_t*>(m)[i] = vv[i];
                    }
                }
            }
        }
        __syncwarp();
    }
}

void gg_launch_fine(const GGConfig& cfg, const GGBuffers& b, uint8_t* dst, size_t stride, cudaStream_t s, const GGFineRange& rg, uint32_t part,
                    const GGFineMirrors& mir) {
    // tile rows [row0, row1) relative to the band, tile-pair columns [px0, px1); `part` selects the work cursor (each
    // launch of a frame needs its own)
    if (rg.row1 <= rg.row0 || rg.px1 <= rg.px0) return;
    uint32_t n_pairs = (rg.px1 - rg.px0) * (rg.row1 - rg.row0);
    uint32_t blocks = (n_pairs + FINE_WARPS - 1) / FINE_WARPS;
    // resident CTAs: 4 per SM at 128 registers. (5 at 96 registers spilled in the command loop: 10 % more instructions; the
    // benchmark scene's fine stage, bound by single-warp latency, went 0.258 -> 0.215 ms with the wider budget.)
    uint32_t max_blocks = cfg.sm_count * 4;
    if (blocks > max_blocks) blocks = max_blocks;
    const int smem = FINE_WARPS * FINE_SMEM_PER_WARP;
    cudaFuncSetAttribute(fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device; cheap
    fine_kernel<<<blocks, FINE_WARPS * 32, smem, s>>>(cfg, b.ptcl_off, b.ptcl_len, b.ptcl, b.restart_pt, b.segments, b.spill_off, b.spill, b.bump, dst, stride,
                                                      rg, part, mir, b.scene + cfg.grad_base);
}

// ---------------------------------------------------------------- deferred band broadcast (ggcuda_broadcast_band)
// Copies a finished band into the frames of the other devices: 16 bytes per thread and store, four loads in flight per
// thread, a grid of a few CTAs -- the copy is bound by the NVLink ingress of the receivers (every device takes in N - 1
// bands), not by this device, and it runs beside the next frame's kernels.
__global__ void __launch_bounds__(256) band_bcast_kernel(const uint4* __restrict__ src, GGFineMirrors mir, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n16; i0 += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const size_t i = i0 + k * stride; if (i < n16) v[k] = __ldg(src + i); }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t i = i0 + k * stride;
            if (i >= n16) continue;
            if (mir.multicast) {
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mir.p[0] + 16 * i), "f"(__uint_as_float(v[k].x)), "f"(__uint_as_float(v[k].y)),
                             "f"(__uint_as_float(v[k].z)), "f"(__uint_as_float(v[k].w)) : "memory");
            } else {
                for (uint32_t q = 0; q < mir.n; q++) reinterpret_cast<uint4*>(mir.p[q])[i] = v[k];
            }
        }
    }
}
void gg_launch_band_bcast(const void* band, const GGFineMirrors& mir, size_t bytes, uint32_t sm_count, cudaStream_t s) {
    const size_t n16 = bytes / 16;
    if (!n16) return;
    const uint32_t blocks = (uint32_t)std::min<size_t>((n16 + 1023) / 1024, std::max(8u, sm_count / 4));
    band_bcast_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(band), mir, n16);
}
