// gg_b200/csrc/stroke.cuh -- stroke expansion on the device, inside flatten.
//
// The reference expands strokes on the host, one path at a time, before the Vello pipeline sees them
// (internal/stroke/expander.go:264 via scene/renderer.go:655-713; tilecompute/flatten.go:7-8: "stroke expansion
// ... is not yet ported"). Here a stroked path travels as its centre line plus a 3-word style and the thread
// that owns a path segment writes that segment's share of the outline as LineSoup, to be filled NonZero
// (software.go:1145-1226 fills the expanded stroke with NonZero too):
//   * the two parallel sides of the segment (curves: flattened along the Euler-spiral subdivision of
//     flatten.go, one offset point per subdivision on the spiral's own normal, euler.go:133-146),
//   * the join to the next segment of the subpath (miter / round / bevel, paint.go LineJoin), or the end cap,
//   * the start cap, drawn by the marker segment the host appends to every open subpath.
// The outline is built as a union of positively oriented pieces only, so NonZero winding can never cancel:
// a side quad whose inner edge would run backwards (offset larger than the radius of curvature) is replaced by
// the piece's rectangle plus bevels, and the inner side of every join goes through the centre point.
#pragma once
#include "flatten.cuh"

#define GG_STYLE_STROKE 0x01u      // style word 0: bit 0 stroke, bit 1 even-odd, bits 2-3 join, bits 4-5 cap
#define GG_STYLE_WORDS 3           // {flags, width (f32 bits), miter limit (f32 bits)}
#define GG_PTAG_MARKER 0x80u       // on a segment tag: copy of the subpath's first segment (stroke marker)
#define GG_PTAG_MARKER_MOVE 0x8Cu  // MoveTo back to the first point of an OPEN subpath, before its marker

struct StrokeStyle { float hw, miter_limit; uint32_t join, cap; };
struct StrokeVertex { V2 p, n, l, r; };   // centre point, offset vector, left (p + n) and right (p - n) outline points

struct LineOut { GGLine* out; uint32_t n, cap, path_ix; float bb[4]; };

template <bool EMIT>
__device__ __forceinline__ void put_line(LineOut& o, V2 a, V2 b) {
    if (veq(a, b)) return;
    if (EMIT) {
        if (o.n < o.cap) { GGLine l; l.path_ix = o.path_ix; l.p0x = a.x; l.p0y = a.y; l.p1x = b.x; l.p1y = b.y; o.out[o.n] = l; }
        o.bb[0] = fminf(o.bb[0], fminf(a.x, b.x)); o.bb[1] = fminf(o.bb[1], fminf(a.y, b.y));
        o.bb[2] = fmaxf(o.bb[2], fmaxf(a.x, b.x)); o.bb[3] = fmaxf(o.bb[3], fmaxf(a.y, b.y));
    }
    o.n++;
}

__device__ __forceinline__ float vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float atan2_32(float y, float x) { return (float)atan2((double)y, (double)x); }

// Offset vector of a tangent: the tangent turned by +90 degrees, scaled to the half width.
__device__ __forceinline__ V2 stroke_normal(V2 t, float hw) {
    float len = sqrt32(t.x * t.x + t.y * t.y);
    float s = len > 0.0f ? hw / len : 0.0f;
    return mk(-t.y * s, t.x * s);
}
__device__ __forceinline__ StrokeVertex stroke_vertex(V2 p, V2 n) {
    StrokeVertex v; v.p = p; v.n = n; v.l = vadd(p, n); v.r = vsub(p, n);
    return v;
}
// Tangents of a segment in cubic form (kind 1: line p0 -> p3). The host drops segments whose points all coincide.
__device__ __forceinline__ V2 seg_start_tangent(V2 p0, V2 p1, V2 p2, V2 p3, uint32_t kind) {
    if (kind == 1) return vsub(p3, p0);
    if (!veq(p1, p0)) return vsub(p1, p0);
    if (!veq(p2, p0)) return vsub(p2, p0);
    return vsub(p3, p0);
}
__device__ __forceinline__ V2 seg_end_tangent(V2 p0, V2 p1, V2 p2, V2 p3, uint32_t kind) {
    if (kind == 1) return vsub(p3, p0);
    if (!veq(p3, p2)) return vsub(p3, p2);
    if (!veq(p3, p1)) return vsub(p3, p1);
    return vsub(p3, p0);
}

__device__ __forceinline__ float arc_step(float hw) {   // angle whose chord stays within the flatten tolerance of the circle
    float c = 1.0f - 0.25f / hw;
    if (c < -1.0f) c = -1.0f;
    return 2.0f * (float)acos((double)c);
}
// Points of the arc from c + v0 turning by `sweep` (exclusive of both ends; the caller closes it on the exact end point).
template <bool EMIT>
__device__ inline V2 put_arc(LineOut& o, V2 c, V2 v0, float sweep, float hw, V2 from) {
    float nf = ceilf(fabsf(sweep) / arc_step(hw));
    if (!(nf >= 1.0f)) nf = 1.0f;
    if (nf > 1024.0f) nf = 1024.0f;
    int n = (int)nf;
    V2 last = from;
    for (int k = 1; k < n; k++) {
        float a = sweep * (float)k / nf;
        float cs = cos32(a), sn = sin32(a);
        V2 p = mk(c.x + (v0.x * cs - v0.y * sn), c.y + (v0.x * sn + v0.y * cs));
        put_line<EMIT>(o, last, p);
        last = p;
    }
    return last;
}

// One side quad between two outline vertices of the same segment: exactly two lines in the curve's own slots
// (main[0], main[1]; possibly zero length, which path_count ignores). A quad with an edge that would run backwards
// becomes the piece's rectangle + a bevel on the outer side of each end with the inner side through the centre
// point: up to eight lines, the surplus appended behind the scan-allocated lines (slots claimed from bump->lines).
__device__ __forceinline__ void write_line(GGLine* dst, uint32_t path_ix, V2 a, V2 b, float* bb) {
    GGLine l; l.path_ix = path_ix; l.p0x = a.x; l.p0y = a.y; l.p1x = b.x; l.p1y = b.y;
    *dst = l;
    bb[0] = fminf(bb[0], fminf(a.x, b.x)); bb[1] = fminf(bb[1], fminf(a.y, b.y));
    bb[2] = fmaxf(bb[2], fmaxf(a.x, b.x)); bb[3] = fmaxf(bb[3], fmaxf(a.y, b.y));
}
__device__ inline void stroke_piece_emit(GGLine* lines, uint32_t main_ix, uint32_t lines_cap, uint32_t* bump_lines, uint32_t* bump_failed,
                                         uint32_t path_ix, const StrokeVertex& a, const StrokeVertex& b, float hw, float* bb) {
    V2 e = vsub(b.p, a.p);
    float dl = vdot(vsub(b.l, a.l), e), dr = vdot(vsub(b.r, a.r), e);
    if ((dl > 0.0f && dr > 0.0f) || (e.x == 0.0f && e.y == 0.0f)) {
        write_line(lines + main_ix, path_ix, a.l, b.l, bb);
        write_line(lines + main_ix + 1, path_ix, b.r, a.r, bb);
        return;
    }
    V2 n = stroke_normal(e, hw);
    V2 a0 = vadd(a.p, n), b0 = vsub(a.p, n), a1 = vadd(b.p, n), b1 = vsub(b.p, n);
    V2 q[16];
    int k = 0;
#define GG_Q(u, v) do { if (!veq(u, v)) { q[k++] = u; q[k++] = v; } } while (0)
    if (vcross(a.n, n) > 0.0f) { GG_Q(a.l, a.p); GG_Q(a.p, a0); GG_Q(b0, a.r); }
    else { GG_Q(a.l, a0); GG_Q(b0, a.p); GG_Q(a.p, a.r); }
    GG_Q(a0, a1);
    GG_Q(b1, b0);
    if (vcross(n, b.n) > 0.0f) { GG_Q(a1, b.p); GG_Q(b.p, b.l); GG_Q(b.r, b1); }
    else { GG_Q(a1, b.l); GG_Q(b.r, b.p); GG_Q(b.p, b1); }
#undef GG_Q
    int nl = k / 2;
    for (int i = 0; i < 2; i++) {
        if (i < nl) write_line(lines + main_ix + i, path_ix, q[2 * i], q[2 * i + 1], bb);
        else write_line(lines + main_ix + i, path_ix, a.p, a.p, bb);
    }
    if (nl > 2) {
        uint32_t extra = (uint32_t)(nl - 2);
        uint32_t slot = atomicAdd(bump_lines, extra);
        if ((uint64_t)slot + extra > lines_cap) { atomicOr(bump_failed, GG_FAIL_LINES); return; }
        for (int i = 2; i < nl; i++) write_line(lines + slot + (i - 2), path_ix, q[2 * i], q[2 * i + 1], bb);
    }
}

// Join at p between the end of one segment (offset n0) and the start of the next (offset n1).
template <bool EMIT>
__device__ inline void stroke_join(LineOut& o, V2 p, V2 n0, V2 n1, const StrokeStyle& st) {
    if (veq(n0, n1)) return;
    float cr = vcross(n0, n1), dt = vdot(n0, n1);
    // side = +1: the left side is the outer one (the path turns right), -1: the right side
    float side = cr > 0.0f ? -1.0f : 1.0f;
    V2 from, to, v0, v1;
    if (side > 0.0f) {   // left chain runs forward: p + n0 -> p + n1; right chain (backward) through the centre
        put_line<EMIT>(o, vsub(p, n1), p); put_line<EMIT>(o, p, vsub(p, n0));
        v0 = n0; v1 = n1;
    } else {             // right chain runs backward: p - n1 -> p - n0
        put_line<EMIT>(o, vadd(p, n0), p); put_line<EMIT>(o, p, vadd(p, n1));
        v0 = mk(-n1.x, -n1.y); v1 = mk(-n0.x, -n0.y);
    }
    from = vadd(p, v0); to = vadd(p, v1);
    float hw2 = st.hw * st.hw;
    if (st.join == 1u) {            // round
        float sweep = atan2_32(vcross(v0, v1), vdot(v0, v1));
        V2 last = put_arc<EMIT>(o, p, v0, sweep, st.hw, from);
        put_line<EMIT>(o, last, to);
    } else if (st.join == 0u && 2.0f * hw2 < st.miter_limit * st.miter_limit * (hw2 + dt) && hw2 + dt > 0.0f) {   // miter within the limit (expander.go:455-456)
        float k = hw2 / (hw2 + dt);
        V2 m = mk(p.x + (v0.x + v1.x) * k, p.y + (v0.y + v1.y) * k);
        put_line<EMIT>(o, from, m);
        put_line<EMIT>(o, m, to);
    } else {                        // bevel
        put_line<EMIT>(o, from, to);
    }
}

// Cap from p + nf round to p - nf, bulging towards nf turned by -90 degrees (end cap: nf = n; start cap: nf = -n).
template <bool EMIT>
__device__ inline void stroke_cap(LineOut& o, V2 p, V2 nf, const StrokeStyle& st) {
    V2 a = vadd(p, nf), b = vsub(p, nf);
    if (st.cap == 1u) {             // round
        V2 last = put_arc<EMIT>(o, p, nf, -3.14159265358979323846f, st.hw, a);
        put_line<EMIT>(o, last, b);
    } else if (st.cap == 2u) {      // square
        V2 d = mk(nf.y, -nf.x);
        V2 a2 = vadd(a, d), b2 = vadd(b, d);
        put_line<EMIT>(o, a, a2); put_line<EMIT>(o, a2, b2); put_line<EMIT>(o, b2, b);
    } else {
        put_line<EMIT>(o, a, b);
    }
}

// Outline vertex at subdivision point j of an Euler segment: the point and the spiral's own normal there
// (euler.go:121-137), scaled to the half width.
__device__ inline StrokeVertex eseg_vertex(const GGESeg& r, int j, float hw, V2 n_end, V2 n_fallback) {
    if ((r.flags & GG_ESEG_LAST) && j == (int)r.n - 1) return stroke_vertex(r.p1, n_end);
    EulerParams ep; ep.th0 = r.th0; ep.k0 = r.k0; ep.k1 = r.k1; ep.ch = r.ch;
    float s = eseg_param(r, j);
    V2 pc = euler_seg_eval(r.p0, r.p1, ep, s);
    V2 chord = vsub(r.p1, r.p0);
    if (vlen_sq(chord) < 1e-12f) return stroke_vertex(pc, n_fallback);   // euler.go:44: no usable chord direction
    float th = (ep.k0 + 0.5f * ep.k1 * (s - 1.0f)) * s - ep.th0;
    float sx = sin32(th), sy = cos32(th);
    float nscale = hw / r.chord_len;
    return stroke_vertex(pc, mk((chord.x * sx - chord.y * sy) * nscale, (chord.x * sy + chord.y * sx) * nscale));
}
