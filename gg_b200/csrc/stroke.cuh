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

// One side quad between two outline vertices of the same segment.
template <bool EMIT>
__device__ inline void stroke_piece(LineOut& o, const StrokeVertex& a, const StrokeVertex& b, float hw) {
    V2 e = vsub(b.p, a.p);
    float dl = vdot(vsub(b.l, a.l), e), dr = vdot(vsub(b.r, a.r), e);
    if ((dl > 0.0f && dr > 0.0f) || (e.x == 0.0f && e.y == 0.0f)) {
        put_line<EMIT>(o, a.l, b.l);
        put_line<EMIT>(o, b.r, a.r);
        return;
    }
    // An edge of the quad runs backwards: rectangle of the piece + a bevel on the outer side of each end,
    // the inner side through the centre point.
    V2 n = stroke_normal(e, hw);
    V2 a0 = vadd(a.p, n), b0 = vsub(a.p, n), a1 = vadd(b.p, n), b1 = vsub(b.p, n);
    if (vcross(a.n, n) > 0.0f) { put_line<EMIT>(o, a.l, a.p); put_line<EMIT>(o, a.p, a0); put_line<EMIT>(o, b0, a.r); }
    else { put_line<EMIT>(o, a.l, a0); put_line<EMIT>(o, b0, a.p); put_line<EMIT>(o, a.p, a.r); }
    put_line<EMIT>(o, a0, a1);
    put_line<EMIT>(o, b1, b0);
    if (vcross(n, b.n) > 0.0f) { put_line<EMIT>(o, a1, b.p); put_line<EMIT>(o, b.p, b.l); put_line<EMIT>(o, b.r, b1); }
    else { put_line<EMIT>(o, a1, b.l); put_line<EMIT>(o, b.r, b.p); put_line<EMIT>(o, b.p, b1); }
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

// Both sides of a cubic: the subdivision loop of flatten_cubic (flatten.go:60-184) with one offset vertex per
// subdivision point. The line count of an Euler segment is raised by sqrt(1 + hw * max curvature) so the outer
// parallel curve stays within the tolerance as well.
template <bool EMIT>
__device__ inline void stroke_cubic(LineOut& o, V2 p0, V2 p1, V2 p2, V2 p3, V2 n_start, V2 n_end, float hw) {
    const float DERIV_THRESH = 1e-6f, DERIV_EPS = 1e-6f, SUBDIV_LIMIT = 1.0f / 65536.0f, FLATTEN_TOL = 0.25f;
    uint32_t t0u = 0;
    float dt = 1.0f;
    V2 last_p = p0;
    V2 last_q = vsub(p1, p0);
    if (vlen_sq(last_q) < DERIV_THRESH * DERIV_THRESH) {
        V2 dummy; eval_cubic_and_deriv(p0, p1, p2, p3, DERIV_EPS, &dummy, &last_q);
    }
    float last_t = 0.0f;
    StrokeVertex v0 = stroke_vertex(p0, n_start);
    for (;;) {
        float t0 = (float)t0u * dt;
        if (t0 == 1.0f) break;
        float t1 = t0 + dt;
        V2 this_p0 = last_p, this_q0 = last_q, this_p1, this_q1;
        eval_cubic_and_deriv(p0, p1, p2, p3, t1, &this_p1, &this_q1);
        if (vlen_sq(this_q1) < DERIV_THRESH * DERIV_THRESH) {
            V2 new_p1, new_q1;
            eval_cubic_and_deriv(p0, p1, p2, p3, t1 - DERIV_EPS, &new_p1, &new_q1);
            this_q1 = new_q1;
            if (t1 < 1.0f) { this_p1 = new_p1; t1 -= DERIV_EPS; }
        }
        float actual_dt = t1 - last_t;
        CubicParams cp = cubic_params_from_points_derivs(this_p0, this_p1, this_q0, this_q1, actual_dt);
        if (cp.err <= FLATTEN_TOL || dt <= SUBDIV_LIMIT) {
            EulerParams ep = euler_params_from_angles(cp.th0, cp.th1);
            float k0_minus_half_k1 = ep.k0 - 0.5f * ep.k1;
            float k1 = ep.k1;
            float scale_mul = 0.5f * (float)(1.41421356237309504880168872420969808 / 2.0) * sqrt32(cp.chord_len / (ep.ch * FLATTEN_TOL));
            float k_abs = f_max(fabsf(k0_minus_half_k1), fabsf(k0_minus_half_k1 + k1));
            float widen = sqrt32(1.0f + hw * k_abs * ep.ch / cp.chord_len);
            float n_frac;
            bool low_k1;
            float a = 0, b = 0, integral = 0, int0 = 0;
            if (fabsf(k1) < 1e-3f) {
                float k = k0_minus_half_k1 + 0.5f * k1;
                n_frac = sqrt32(fabsf(k));
                low_k1 = true;
            } else {
                a = k1;
                b = k0_minus_half_k1;
                int0 = cube_signed_sqrt(b);
                float int1 = cube_signed_sqrt(a + b);
                integral = int1 - int0;
                n_frac = (float)(2.0 / 3.0) * integral / a;
                low_k1 = false;
            }
            float n = ceilf(n_frac * scale_mul * widen);
            if (n < 1) n = 1;
            if (n > 100) n = 100;
            int n_int = (n != n) ? 0 : (int)n;
            V2 chord = vsub(this_p1, this_p0);
            float nscale = hw / cp.chord_len;
            bool tiny = vlen_sq(chord) < 1e-12f;   // euler.go:44: no usable chord direction, keep the previous offset
            for (int i = 0; i < n_int; i++) {
                StrokeVertex v1;
                if (i == n_int - 1 && t1 == 1.0f) {
                    v1 = stroke_vertex(p3, n_end);
                } else {
                    float t = (float)(i + 1) / n;
                    float s;
                    if (low_k1) {
                        s = t;
                    } else {
                        float c = (float)cbrt((double)(integral * t + int0));
                        float inv = c * fabsf(c);
                        s = (inv - b) / a;
                    }
                    V2 pc = euler_seg_eval(this_p0, this_p1, ep, s);
                    V2 nn = v0.n;
                    if (!tiny) {
                        float th = (ep.k0 + 0.5f * ep.k1 * (s - 1.0f)) * s - ep.th0;   // euler.go:121-123
                        float sx = sin32(th), sy = cos32(th);                           // euler.go:133-137: offset direction
                        nn = mk((chord.x * sx - chord.y * sy) * nscale, (chord.x * sy + chord.y * sx) * nscale);
                    }
                    v1 = stroke_vertex(pc, nn);
                }
                stroke_piece<EMIT>(o, v0, v1, hw);
                v0 = v1;
            }
            last_p = this_p1; last_q = this_q1; last_t = t1;
            t0u++;
            uint32_t shift = (uint32_t)(__ffs((int)t0u) - 1);
            t0u >>= shift;
            dt *= (float)(1u << shift);
        } else {
            if (t0u < 0xFFFFFFFFu / 2) t0u *= 2;
            dt *= 0.5f;
        }
    }
    // t == 1 is reached through an exact end point only if the last subdivision ended there; close the chain otherwise
    if (!veq(v0.p, p3) || !veq(v0.n, n_end)) stroke_piece<EMIT>(o, v0, stroke_vertex(p3, n_end), hw);
}
