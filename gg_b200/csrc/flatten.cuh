// gg_b200/csrc/flatten.cuh -- Euler-spiral cubic flattening on the device.
//
// Behavioural spec: gg internal/gpu/tilecompute/flatten.go:60-184 (flattenEulerFill),
// euler.go:39-224 and the quad elevation of internal/gpu/path_convert.go:60-72.
// One thread owns one curve; the subdivision loop is run twice (count, then emit) so that
// LineSoup lands in tag order without atomics. float32 arithmetic in the reference's
// operation order (TU compiled with -fmad=false); transcendental calls go through float64
// and are rounded to float32 at the same places the Go code does (util.go:36-37,104-105).
#pragma once
#include "common.cuh"

struct V2 { float x, y; };
__device__ __forceinline__ V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ V2 vadd(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 vmul(V2 a, float s) { return mk(a.x * s, a.y * s); }
__device__ __forceinline__ float vlen_sq(V2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ bool veq(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }

// Go math.Hypot: p*sqrt(1+q*q) on ordered magnitudes (not the correctly rounded libm hypot).
__device__ __forceinline__ double go_hypot(double p, double q) {
    if (isinf(p) || isinf(q)) return CUDART_INF;
    if (isnan(p) || isnan(q)) return CUDART_NAN;
    p = fabs(p); q = fabs(q);
    if (p < q) { double t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = q / p;
    return p * sqrt(1 + q * q);
}
__device__ __forceinline__ float sin32(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos32(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float sqrt32(float x) { return (float)sqrt((double)x); }
__device__ __forceinline__ float pow5(float b) { float r = 1.0f; r *= b; r *= b; r *= b; r *= b; r *= b; return r; }

struct CubicParams { float th0, th1, chord_len, err; };
struct EulerParams { float th0, k0, k1, ch; };

__device__ inline CubicParams cubic_params_from_points_derivs(V2 p0, V2 p1, V2 q0, V2 q1, float dt) {   // euler.go:39-90
    const float TANGENT_THRESH = 1e-6f;
    V2 chord = vsub(p1, p0);
    float chord_sq = vlen_sq(chord);
    float chord_len = sqrt32(chord_sq);
    CubicParams cp;
    if (chord_sq < TANGENT_THRESH * TANGENT_THRESH) {
        float chord_err = sqrt32((float)(9.0 / 32.0) * (vlen_sq(q0) + vlen_sq(q1))) * dt;
        cp.th0 = 0; cp.th1 = 0; cp.chord_len = TANGENT_THRESH; cp.err = chord_err;
        return cp;
    }
    float scale = dt / chord_sq;
    V2 h0 = mk(q0.x * chord.x + q0.y * chord.y, q0.y * chord.x - q0.x * chord.y);
    float th0 = (float)atan2((double)h0.y, (double)h0.x);
    float d0 = (float)go_hypot((double)h0.x, (double)h0.y) * scale;
    V2 h1 = mk(q1.x * chord.x + q1.y * chord.y, q1.x * chord.y - q1.y * chord.x);
    float th1 = (float)atan2((double)h1.y, (double)h1.x);
    float d1 = (float)go_hypot((double)h1.x, (double)h1.y) * scale;
    float cth0 = cos32(th0), cth1 = cos32(th1);
    float err;
    if (cth0 * cth1 < 0) {
        err = 2.0f;
    } else {
        const float two_thirds = (float)(2.0 / 3.0);
        float e0 = two_thirds / f_max(1.0f + cth0, 1e-9f);
        float e1 = two_thirds / f_max(1.0f + cth1, 1e-9f);
        float s0 = sin32(th0), s1 = sin32(th1);
        float s01 = cth0 * s1 + cth1 * s0;
        float amin = 0.15f * (2 * e0 * s0 + 2 * e1 * s1 - e0 * e1 * s01);
        float a = 0.15f * (2 * d0 * s0 + 2 * d1 * s1 - d0 * d1 * s01);
        float aerr = fabsf(a - amin);
        float symm = fabsf(th0 + th1);
        float asymm = fabsf(th0 - th1);
        float dist = (float)go_hypot((double)(d0 - e0), (double)(d1 - e1));
        float ctr = 4.625e-6f * pow5(symm) + 7.5e-3f * asymm * symm * symm;
        float halo_symm = 5e-3f * symm * dist;
        float halo_asymm = 7e-2f * asymm * dist;
        err = ctr + 1.55f * aerr + halo_symm + halo_asymm;
    }
    err *= chord_len;
    cp.th0 = th0; cp.th1 = th1; cp.chord_len = chord_len; cp.err = err;
    return cp;
}

__device__ inline EulerParams euler_params_from_angles(float th0, float th1) {   // euler.go:93-119
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
    EulerParams ep; ep.th0 = th0; ep.k0 = k0; ep.k1 = k1; ep.ch = ch;
    return ep;
}

__device__ inline void integ_euler_10(float k0, float k1, float* uo, float* vo) {   // euler.go:149-186
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

// euler.go:121-146 with offset == 0 (fill): the offset vector is (0*sin, 0*cos) == +/-0 and
// cannot change a finite sum, so it is not evaluated.
__device__ inline V2 euler_seg_eval(V2 p0, V2 p1, const EulerParams& ep, float t) {
    float thm = (ep.k0 + 0.5f * ep.k1 * (t * 0.5f - 1.0f)) * (t * 0.5f) - ep.th0;
    float u, v;
    integ_euler_10((ep.k0 + ep.k1 * (0.5f * t - 0.5f)) * t, ep.k1 * t * t, &u, &v);
    float s = t / ep.ch * sin32(thm);
    float c = t / ep.ch * cos32(thm);
    V2 pt = mk(u * c - v * s, -v * c - u * s);
    V2 chord = vsub(p1, p0);
    return mk(p0.x + chord.x * pt.x - chord.y * pt.y, p0.y + chord.x * pt.y + chord.y * pt.x);
}

__device__ __forceinline__ void eval_cubic_and_deriv(V2 p0, V2 p1, V2 p2, V2 p3, float t, V2* po, V2* qo) {   // flatten.go:46-56
    float m = 1.0f - t;
    float mm = m * m, mt = m * t, tt = t * t;
    *po = vadd(vmul(p0, mm * m), vmul(vadd(vadd(vmul(p1, 3 * mm), vmul(p2, 3 * mt)), vmul(p3, tt)), t));
    *qo = vadd(vadd(vmul(vsub(p1, p0), mm), vmul(vsub(p2, p1), 2 * mt)), vmul(vsub(p3, p2), tt));
}
__device__ __forceinline__ float cube_signed_sqrt(float x) { return x * sqrt32(fabsf(x)); }   // flatten.go:195

// One accepted Euler segment of a curve on the flatten work list (or, for stroked paths, one straight piece / one cap
// marker). The subdivision loop (serial per curve, cheap) only decides these records and how many lines each one
// owns; the lines themselves -- one Euler-spiral evaluation with float64 transcendentals per point -- are written by
// one thread per record.
struct GGESeg {
    V2 p0, p1;                    // end points of the segment's chord (this_p0, this_p1 of flatten.go:88-96)
    float th0, k0, k1, ch;        // EulerParams (euler.go:93-119)
    float chord_len, n;           // CubicParams.chord_len; line count n as the float the reference divides by
    float a, b, integral, int0;   // inverse-integral parameters (flatten.go:119-129)
    uint32_t tag_ix;              // tag byte of the owning segment
    uint32_t path_ix;
    uint32_t line_rel;            // first line of this record, relative to line_off[tag_ix]
    uint32_t prev;                // record holding the previous Euler segment of the same curve (0xffffffff: first)
    uint32_t flags;               // GG_ESEG_*
};
#define GG_ESEG_LOW_K1 1u         // s = t (flatten.go:148-149)
#define GG_ESEG_LAST 2u           // t1 == 1: the last point is the curve's end point itself
#define GG_ESEG_STROKE 4u         // stroked path: two outline lines per subdivision (stroke.cuh)
#define GG_ESEG_LINE 8u           // stroked straight segment: one piece, no Euler parameters
#define GG_ESEG_CAP 16u           // stroke marker after an open subpath: the start cap

// The adaptive subdivision of flatten.go:76-183 without the emission loop, as a resumable state machine: one call of
// subdiv_step() tries one interval [t0, t0 + dt] and either halves it (returns 0), accepts it (returns 1 with the
// Euler segment's record filled in, bookkeeping fields aside) or finds the curve finished (returns 2). A lane of
// flatten_subdivide_kernel holds one curve at a time and picks up the next one as soon as its own is finished, so the
// lanes of a warp stay in the same instruction stream whatever the depth their curves subdivide to.
// widen_hw > 0 raises n by sqrt(1 + hw * max curvature) so that the outer parallel curve of a stroke stays within the
// tolerance too.
struct SubdivState {
    V2 p0, p1, p2, p3, last_p, last_q;
    uint32_t t0u;
    float dt, last_t, widen_hw;
};
__device__ __forceinline__ bool subdiv_begin(SubdivState& s, V2 p0, V2 p1, V2 p2, V2 p3, float widen_hw) {
    const float DERIV_THRESH = 1e-6f, DERIV_EPS = 1e-6f;
    if (veq(p0, p1) && veq(p0, p2) && veq(p0, p3)) return false;
    s.p0 = p0; s.p1 = p1; s.p2 = p2; s.p3 = p3; s.widen_hw = widen_hw;
    s.t0u = 0; s.dt = 1.0f; s.last_t = 0.0f;
    s.last_p = p0;
    s.last_q = vsub(p1, p0);
    if (vlen_sq(s.last_q) < DERIV_THRESH * DERIV_THRESH) {
        V2 dummy; eval_cubic_and_deriv(p0, p1, p2, p3, DERIV_EPS, &dummy, &s.last_q);
    }
    return true;
}
__device__ inline int subdiv_step(SubdivState& s, GGESeg& r) {
    const float DERIV_THRESH = 1e-6f, DERIV_EPS = 1e-6f, SUBDIV_LIMIT = 1.0f / 65536.0f, FLATTEN_TOL = 0.25f;
    float t0 = (float)s.t0u * s.dt;
    if (t0 == 1.0f) return 2;
    float t1 = t0 + s.dt;
    V2 this_p0 = s.last_p, this_q0 = s.last_q, this_p1, this_q1;
    eval_cubic_and_deriv(s.p0, s.p1, s.p2, s.p3, t1, &this_p1, &this_q1);
    if (vlen_sq(this_q1) < DERIV_THRESH * DERIV_THRESH) {
        V2 new_p1, new_q1;
        eval_cubic_and_deriv(s.p0, s.p1, s.p2, s.p3, t1 - DERIV_EPS, &new_p1, &new_q1);
        this_q1 = new_q1;
        if (t1 < 1.0f) { this_p1 = new_p1; t1 -= DERIV_EPS; }
    }
    float actual_dt = t1 - s.last_t;
    CubicParams cp = cubic_params_from_points_derivs(this_p0, this_p1, this_q0, this_q1, actual_dt);
    if (!(cp.err <= FLATTEN_TOL || s.dt <= SUBDIV_LIMIT)) {
        if (s.t0u < 0xFFFFFFFFu / 2) s.t0u *= 2;
        s.dt *= 0.5f;
        return 0;
    }
    EulerParams ep = euler_params_from_angles(cp.th0, cp.th1);
    float k0_minus_half_k1 = ep.k0 - 0.5f * ep.k1;
    float k1 = ep.k1;
    float scale_mul = 0.5f * (float)(1.41421356237309504880168872420969808 / 2.0) * sqrt32(cp.chord_len / (ep.ch * FLATTEN_TOL));
    r.a = 0; r.b = 0; r.integral = 0; r.int0 = 0; r.flags = 0;
    float n_frac;
    if (fabsf(k1) < 1e-3f) {
        float k = k0_minus_half_k1 + 0.5f * k1;
        n_frac = sqrt32(fabsf(k));
        r.flags |= GG_ESEG_LOW_K1;
    } else {
        r.a = k1;
        r.b = k0_minus_half_k1;
        r.int0 = cube_signed_sqrt(r.b);
        float int1 = cube_signed_sqrt(r.a + r.b);
        r.integral = int1 - r.int0;
        n_frac = (float)(2.0 / 3.0) * r.integral / r.a;
    }
    float nn = n_frac * scale_mul;
    if (s.widen_hw > 0.0f) {
        float k_abs = f_max(fabsf(k0_minus_half_k1), fabsf(k0_minus_half_k1 + k1));
        nn = nn * sqrt32(1.0f + s.widen_hw * k_abs * ep.ch / cp.chord_len);
    }
    float n = ceilf(nn);
    if (n < 1) n = 1;
    if (n > 100) n = 100;
    if (n != n) n = 0;   // NaN (degenerate input): Go's int(NaN) gives an empty loop
    r.p0 = this_p0; r.p1 = this_p1;
    r.th0 = ep.th0; r.k0 = ep.k0; r.k1 = ep.k1; r.ch = ep.ch;
    r.chord_len = cp.chord_len; r.n = n;
    if (t1 == 1.0f) r.flags |= GG_ESEG_LAST;
    s.last_p = this_p1; s.last_q = this_q1; s.last_t = t1;
    s.t0u++;
    uint32_t shift = (uint32_t)(__ffs((int)s.t0u) - 1);   // trailing zeros; t0u != 0 here
    s.t0u >>= shift;
    s.dt *= (float)(1u << shift);
    return 1;
}

// Arc-length parameter of subdivision point j (0-based: the point that ends line j), flatten.go:144-157.
__device__ __forceinline__ float eseg_param(const GGESeg& r, int j) {
    float t = (float)(j + 1) / r.n;
    if (r.flags & GG_ESEG_LOW_K1) return t;
    float c = (float)cbrt((double)(r.integral * t + r.int0));
    float inv = c * fabsf(c);
    return (inv - r.b) / r.a;
}
__device__ __forceinline__ V2 eseg_point(const GGESeg& r, int j) {
    if ((r.flags & GG_ESEG_LAST) && j == (int)r.n - 1) return r.p1;   // exact end point (flatten.go:141-142)
    EulerParams ep; ep.th0 = r.th0; ep.k0 = r.k0; ep.k1 = r.k1; ep.ch = r.ch;
    return euler_seg_eval(r.p0, r.p1, ep, eseg_param(r, j));
}
