// Backprojection (pixel driven), parallel and fan beam.  Replaces [RECALL] torch_radon
// radon_backward_kernel, which binds the sinogram as a layered texture and reads it with hardware
// linear filtering; here the detector rows are staged in shared memory and interpolated in exact
// float32.
//
// variant 1 (default) -- a CTA owns a TX x TY pixel tile of one slice (each thread PY pixels in a
//   column of the tile).  Views are consumed in chunks of AC: for every view of the chunk the CTA
//   computes the detector interval its tile projects onto (from the four tile corners) and copies
//   that interval, zero filled outside [0, D) (the texture "border" mode), into a SEG-float shared
//   row; then every pixel takes its two taps per view from shared memory with no bounds test.
//   A chunk in which some view's interval does not fit SEG floats is served from global memory.
// variant 0 -- the same pixel loop with both taps read through L1 (__ldg) and tested against the
//   detector range.  A/B baseline and the fallback above.
#include "common.cuh"

namespace pdu {

struct AdjGeom {
    int n, n_angles, det_count;
    int fan, clip;
    float ids;      // 1 / det_spacing
    float s_dist;   // fan: source -> centre
    float k;        // fan: s_dist + d_dist
    float cr;       // det_count / 2 - 0.5  (detector coordinate of u = 0, in tap units)
    float half;     // n / 2 - 0.5
};

// detector coordinate (in tap units: value = (1-fr) s[i0] + fr s[i0+1], i0 = floor(t)) and weight
__device__ __forceinline__ void project(const AdjGeom& g, float cs, float sn, float dx, float dy, float& t, float& w) {
    const float p = fmaf(cs, dx, sn * dy);
    if (g.fan) {
        const float den = fmaf(sn, dx, g.s_dist) - cs * dy;
        const float iden = __fdividef(g.k, den);
        // one Newton step makes the fast reciprocal exact to ~1 ulp
        const float iden2 = fmaf(iden, fmaf(-den, iden, g.k) * __fdividef(1.f, g.k), iden);
        w = iden2;
        t = fmaf(p * g.ids, iden2, g.cr);
    } else {
        w = 1.f;
        t = fmaf(p, g.ids, g.cr);
    }
}

__device__ __forceinline__ float tap_global(const float* __restrict__ row, int D, float t) {
    const float tf = floorf(t);
    const float fr = t - tf;
    const int i0 = (int)tf;
    const float s0 = (unsigned)i0 < (unsigned)D ? __ldg(row + i0) : 0.f;
    const float s1 = (unsigned)(i0 + 1) < (unsigned)D ? __ldg(row + i0 + 1) : 0.f;
    return fmaf(fr, s1 - s0, s0);
}

template <int TX, int TY, int PY>
__global__ void __launch_bounds__(TX*(TY / PY))
    radon_adj_gather_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ trig,
                            const AdjGeom g) {
    const int x = blockIdx.x * TX + threadIdx.x;
    const int y0 = blockIdx.y * TY + threadIdx.y;
    const int b = blockIdx.z;
    const float dx = (float)x - g.half;
    float acc[PY];
#pragma unroll
    for (int k = 0; k < PY; ++k) acc[k] = 0.f;
    const float* sb = sino + (long)b * g.n_angles * g.det_count;
    if (x < g.n) {
        for (int a = 0; a < g.n_angles; ++a) {
            const float cs = __ldg(trig + 2 * a), sn = __ldg(trig + 2 * a + 1);
            const float* row = sb + (long)a * g.det_count;
#pragma unroll
            for (int k = 0; k < PY; ++k) {
                const float dy = (float)(y0 + k * (TY / PY)) - g.half;
                float t, w;
                project(g, cs, sn, dx, dy, t, w);
                acc[k] = fmaf(w, tap_global(row, g.det_count, t), acc[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PY; ++k) {
        const int y = y0 + k * (TY / PY);
        if (x < g.n && y < g.n) {
            const float dy = (float)y - g.half;
            const float r2 = dx * dx + dy * dy, lim = 0.25f * (float)g.n * (float)g.n;
            const float v = (g.clip && r2 > lim) ? 0.f : acc[k] * g.ids;
            img[((long)b * g.n + y) * g.n + x] = v;
        }
    }
}

template <int TX, int TY, int PY, int AC, int SEG>
__global__ void __launch_bounds__(TX*(TY / PY))
    radon_adj_tile_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ trig,
                          const AdjGeom g) {
    constexpr int THREADS = TX * (TY / PY);
    constexpr int RY = TY / PY;
    __shared__ float s_seg[AC][SEG];
    __shared__ float2 s_trig[AC];
    __shared__ int s_lo[AC];
    __shared__ int s_big;

    const int tid = threadIdx.y * TX + threadIdx.x;
    const int x = blockIdx.x * TX + threadIdx.x;
    const int y0 = blockIdx.y * TY + threadIdx.y;
    const int b = blockIdx.z;
    const float dx = (float)x - g.half;
    const float* sb = sino + (long)b * g.n_angles * g.det_count;

    // tile corners (clamped to the image so the interval is not wasted on padding pixels)
    const float cx0 = (float)(blockIdx.x * TX) - g.half;
    const float cx1 = (float)min(blockIdx.x * TX + TX - 1, g.n - 1) - g.half;
    const float cy0 = (float)(blockIdx.y * TY) - g.half;
    const float cy1 = (float)min(blockIdx.y * TY + TY - 1, g.n - 1) - g.half;

    float acc[PY];
#pragma unroll
    for (int k = 0; k < PY; ++k) acc[k] = 0.f;

    for (int a0 = 0; a0 < g.n_angles; a0 += AC) {
        const int na = min(AC, g.n_angles - a0);
        __syncthreads();   // previous chunk fully consumed
        if (tid == 0) s_big = 0;
        __syncthreads();
        if (tid < na) {
            const float cs = __ldg(trig + 2 * (a0 + tid)), sn = __ldg(trig + 2 * (a0 + tid) + 1);
            s_trig[tid] = make_float2(cs, sn);
            float t, w, tmin, tmax;
            project(g, cs, sn, cx0, cy0, t, w); tmin = t; tmax = t;
            project(g, cs, sn, cx1, cy0, t, w); tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
            project(g, cs, sn, cx0, cy1, t, w); tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
            project(g, cs, sn, cx1, cy1, t, w); tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
            // the projection is monotone along each tile edge (parallel: affine; fan: projective with
            // the source outside the tile), so the corners bound it; one tap of slack for rounding
            const int lo = (int)floorf(tmin) - 1;
            const int hi = (int)floorf(tmax) + 2;
            s_lo[tid] = lo;
            if (hi - lo + 1 > SEG || !(tmin > -1e8f) || !(tmax < 1e8f)) s_big = 1;
        }
        __syncthreads();
        const bool big = s_big != 0;
        if (!big) {
            for (int i = tid; i < na * SEG; i += THREADS) {
                const int al = i / SEG, c = i - al * SEG;
                const int d = s_lo[al] + c;
                s_seg[al][c] = (unsigned)d < (unsigned)g.det_count ? __ldg(sb + (long)(a0 + al) * g.det_count + d) : 0.f;
            }
            __syncthreads();
            if (x < g.n) {
#pragma unroll 4
                for (int al = 0; al < na; ++al) {
                    const float2 tr = s_trig[al];
                    const float lo = (float)s_lo[al];
                    const float* seg = s_seg[al];
#pragma unroll
                    for (int k = 0; k < PY; ++k) {
                        const float dy = (float)(y0 + k * RY) - g.half;
                        float t, w;
                        project(g, tr.x, tr.y, dx, dy, t, w);
                        const float tl = t - lo;               // >= 1 by construction
                        const float tf = floorf(tl);
                        const float fr = tl - tf;
                        int i0 = (int)tf;
                        i0 = min(max(i0, 0), SEG - 2);         // padding pixels of edge tiles only
                        const float s0 = seg[i0], s1 = seg[i0 + 1];
                        acc[k] = fmaf(w, fmaf(fr, s1 - s0, s0), acc[k]);
                    }
                }
            }
        } else if (x < g.n) {
            for (int al = 0; al < na; ++al) {
                const float2 tr = s_trig[al];
                const float* row = sb + (long)(a0 + al) * g.det_count;
#pragma unroll
                for (int k = 0; k < PY; ++k) {
                    const float dy = (float)(y0 + k * RY) - g.half;
                    float t, w;
                    project(g, tr.x, tr.y, dx, dy, t, w);
                    acc[k] = fmaf(w, tap_global(row, g.det_count, t), acc[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PY; ++k) {
        const int y = y0 + k * RY;
        if (x < g.n && y < g.n) {
            const float dy = (float)y - g.half;
            const float r2 = dx * dx + dy * dy, lim = 0.25f * (float)g.n * (float)g.n;
            const float v = (g.clip && r2 > lim) ? 0.f : acc[k] * g.ids;
            img[((long)b * g.n + y) * g.n + x] = v;
        }
    }
}

}  // namespace pdu

using namespace pdu;

extern "C" int pdu_radon_adj_f32(const float* sino, float* img, const float* trig, int batch,
                                 const pdu_radon_geom_t* g, void* workspace, size_t workspace_bytes,
                                 pdu_stream_t stream) {
    (void)workspace;
    (void)workspace_bytes;
    PDU_REQUIRE(g != nullptr, "pdu_radon_adj_f32: geom is null");
    PDU_REQUIRE(g->geom == PDU_GEOM_PARALLEL || g->geom == PDU_GEOM_FAN, "pdu_radon_adj_f32: unknown geom %d", g->geom);
    PDU_REQUIRE(g->n > 0 && g->n_angles > 0 && g->det_count > 0 && batch > 0,
                "pdu_radon_adj_f32: sizes must be positive (n=%d angles=%d det=%d batch=%d)", g->n, g->n_angles,
                g->det_count, batch);
    PDU_REQUIRE(g->det_spacing > 0.f, "pdu_radon_adj_f32: det_spacing must be > 0");
    PDU_REQUIRE(batch <= 65535, "pdu_radon_adj_f32: batch %d exceeds 65535 (split the call)", batch);
    if (g->geom == PDU_GEOM_FAN)
        PDU_REQUIRE(g->s_dist > 0.f && g->d_dist >= 0.f, "pdu_radon_adj_f32: fan beam needs s_dist > 0, d_dist >= 0");
    PDU_REQUIRE(sino && img && trig, "pdu_radon_adj_f32: null pointer");

    AdjGeom ag;
    ag.n = g->n;
    ag.n_angles = g->n_angles;
    ag.det_count = g->det_count;
    ag.fan = g->geom == PDU_GEOM_FAN;
    ag.clip = g->clip_to_circle;
    ag.ids = 1.f / g->det_spacing;
    ag.s_dist = g->s_dist;
    ag.k = g->s_dist + g->d_dist;
    ag.cr = 0.5f * (float)g->det_count - 0.5f;
    ag.half = 0.5f * (float)g->n - 0.5f;

    cudaStream_t st = (cudaStream_t)stream;
    int variant = option(OPT_RADON_ADJ);
    if (variant < 0) variant = 1;
    constexpr int TX = 32, TY = 32, PY = 4;
    dim3 block(TX, TY / PY);
    dim3 grid((unsigned)cdiv(g->n, TX), (unsigned)cdiv(g->n, TY), (unsigned)batch);
    if (variant == 0) {
        radon_adj_gather_kernel<TX, TY, PY><<<grid, block, 0, st>>>(sino, img, trig, ag);
    } else {
        radon_adj_tile_kernel<TX, TY, PY, 64, 96><<<grid, block, 0, st>>>(sino, img, trig, ag);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}
