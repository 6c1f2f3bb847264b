// Ray set-up shared by the forward projectors.
//
// Every operation here is an explicitly rounded IEEE float32 intrinsic so that nvcc cannot
// contract anything into an FMA: oracle/radon.py::ray_setup_f32 performs the same sequence in
// numpy float32 and must take the same discrete decisions (clip interval, step count).
// Convention restated ([RECALL] torch_radon radon_forward_kernel): the ray runs from the source
// side to the detector side, is clipped to the image square (or inscribed circle), cut into
// n = ceil(length) equal steps and sampled at j = 0 .. n.
#pragma once
#include "common.cuh"

namespace pdu {

struct Ray {
    float xc0, yc0;   // first sample, pixel-centre coordinates (texture coordinate - 0.5)
    float vx, vy;     // step vector
    float step;       // |v|
    int n_steps;      // samples are j = 0 .. n_steps ; -1 == misses the volume
};

__device__ __forceinline__ float guard_nonzero(float d) {
    return d >= 0.f ? fmaxf(d, 1e-6f) : fminf(d, -1e-6f);
}

__device__ __forceinline__ Ray ray_setup(const pdu_radon_geom_t& g, float cs, float sn, int d) {
    Ray r;
    const float v = __fmul_rn((float)g.n, 0.5f);
    const float u = __fmul_rn(__fadd_rn(__fsub_rn((float)d, __fmul_rn((float)g.det_count, 0.5f)), 0.5f),
                              g.det_spacing);
    float sx, sy, ex, ey;
    if (g.geom == PDU_GEOM_PARALLEL) {
        sx = u; sy = (float)g.n; ex = u; ey = -(float)g.n;
    } else {
        sx = 0.f; sy = g.s_dist; ex = u; ey = -g.d_dist;
    }
    const float rsx = __fsub_rn(__fmul_rn(sx, cs), __fmul_rn(sy, sn));
    const float rsy = __fadd_rn(__fmul_rn(sx, sn), __fmul_rn(sy, cs));
    const float rex = __fsub_rn(__fmul_rn(ex, cs), __fmul_rn(ey, sn));
    const float rey = __fadd_rn(__fmul_rn(ex, sn), __fmul_rn(ey, cs));
    const float dx = guard_nonzero(__fsub_rn(rex, rsx));
    const float dy = guard_nonzero(__fsub_rn(rey, rsy));
    float a_s, a_e;
    bool hit = true;
    if (!g.clip_to_circle) {
        const float ax0 = __fdiv_rn(__fsub_rn(-v, rsx), dx);
        const float ax1 = __fdiv_rn(__fsub_rn(v, rsx), dx);
        const float ay0 = __fdiv_rn(__fsub_rn(-v, rsy), dy);
        const float ay1 = __fdiv_rn(__fsub_rn(v, rsy), dy);
        a_s = fmaxf(fminf(ax0, ax1), fminf(ay0, ay1));
        a_e = fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1));
    } else {
        const float a = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float b = __fadd_rn(__fmul_rn(rsx, dx), __fmul_rn(rsy, dy));
        const float c = __fsub_rn(__fadd_rn(__fmul_rn(rsx, rsx), __fmul_rn(rsy, rsy)), __fmul_rn(v, v));
        const float delta = __fsub_rn(__fmul_rn(b, b), __fmul_rn(a, c));
        hit = delta > 0.f;
        const float sq = __fsqrt_rn(hit ? delta : 0.f);
        a_s = __fdiv_rn(__fsub_rn(-b, sq), a);
        a_e = __fdiv_rn(__fadd_rn(-b, sq), a);
    }
    a_s = fmaxf(a_s, 0.f);
    a_e = fminf(a_e, 1.f);
    hit = hit && (a_s < a_e);
    const float x0 = __fadd_rn(__fadd_rn(rsx, __fmul_rn(dx, a_s)), v);
    const float y0 = __fadd_rn(__fadd_rn(rsy, __fmul_rn(dy, a_s)), v);
    const float x1 = __fadd_rn(__fadd_rn(rsx, __fmul_rn(dx, a_e)), v);
    const float y1 = __fadd_rn(__fadd_rn(rsy, __fmul_rn(dy, a_e)), v);
    const float lx = __fsub_rn(x1, x0);
    const float ly = __fsub_rn(y1, y0);
    const float len = __fsqrt_rn(__fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly)));
    int n = (int)ceilf(len);
    hit = hit && (n > 0);
    if (!hit) {
        r.xc0 = r.yc0 = r.vx = r.vy = r.step = 0.f;
        r.n_steps = -1;
        return r;
    }
    const float nf = (float)n;
    r.vx = __fdiv_rn(lx, nf);
    r.vy = __fdiv_rn(ly, nf);
    r.step = __fsqrt_rn(__fadd_rn(__fmul_rn(r.vx, r.vx), __fmul_rn(r.vy, r.vy)));
    r.xc0 = __fsub_rn(x0, 0.5f);
    r.yc0 = __fsub_rn(y0, 0.5f);
    r.n_steps = n;
    return r;
}

// Texture-unit weight emulation (option "tex_weights"): [RECALL] torch_radon samples through CUDA texture filtering,
// whose interpolation weights are 9-bit fixed point with 8 fractional bits (CUDA C programming guide, "Linear
// filtering").  Rounding a fraction in [0, 1] to a multiple of 2^-8: adding 2^15 makes the float32 ulp 2^-8, round to
// nearest even does the rest.  The oracle twin is oracle.radon_forward(..., tex_weights=True).
constexpr float TEXQ_MAGIC = 32768.f;
__device__ __forceinline__ float texq(float f) { return __fsub_rn(__fadd_rn(f, TEXQ_MAGIC), TEXQ_MAGIC); }

// Bilinear sample with a zero border straight from global memory (any coordinates).
__device__ __forceinline__ float bilinear_global(const float* __restrict__ img, int n, float xc, float yc, bool tq = false) {
    const float xf = floorf(xc), yf = floorf(yc);
    float fx = xc - xf, fy = yc - yf;
    if (tq) {
        fx = texq(fx);
        fy = texq(fy);
    }
    const int ix = (int)xf, iy = (int)yf;
    const bool x0 = (unsigned)ix < (unsigned)n, x1 = (unsigned)(ix + 1) < (unsigned)n;
    const bool y0 = (unsigned)iy < (unsigned)n, y1 = (unsigned)(iy + 1) < (unsigned)n;
    const float* p = img + (long)iy * n + ix;
    const float v00 = (x0 && y0) ? __ldg(p) : 0.f;
    const float v01 = (x1 && y0) ? __ldg(p + 1) : 0.f;
    const float v10 = (x0 && y1) ? __ldg(p + n) : 0.f;
    const float v11 = (x1 && y1) ? __ldg(p + n + 1) : 0.f;
    const float top = fmaf(fx, v01 - v00, v00);
    const float bot = fmaf(fx, v11 - v10, v10);
    return fmaf(fy, bot - top, top);
}

}  // namespace pdu
