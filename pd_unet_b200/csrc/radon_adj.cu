// Backprojection (pixel driven), parallel and fan beam.  Replaces [RECALL] torch_radon
// radon_backward_kernel, which binds the sinogram as a layered texture and reads it with hardware
// linear filtering; here the detector rows are staged in shared memory and interpolated in exact
// float32.
//
// variant 3 (default) -- a CTA owns a TX x TY pixel tile of one slice (each thread PY = 8 pixels in a
//   column of the tile).  Views are consumed in chunks of AC: for every view of the chunk the CTA
//   computes, in float64 and relative to the tile, the detector interval the tile projects onto and
//   copies it -- zero filled outside [0, D) (the texture "border" mode), in LINE FORM (A, B) with
//   tap(t) = A + t B -- into a SEG-entry shared row; then every pixel takes its tap per view with one
//   8-byte load and one FMA, two pixels per packed-FP32 instruction, no bounds test.
//   A chunk in which some view's interval does not fit SEG entries is served from global memory.
// variant 0 -- every tap from global memory with float64 coordinates (slow; the reference form of
//   the fallback above).
// (r01 also carried the (value, difference) tap form with 8 and 4 pixels per thread and a forced 96-entry
//  segment -- 254 / 281 / 194 us against 182 -- deleted in r02.)
#include "radon_common.cuh"

namespace pdu {

struct AdjGeom {
    int n, n_angles, det_count;
    int fan, clip;
    float ids;      // 1 / det_spacing
    float det_spacing, d_dist;
    float s_dist;   // fan: source -> centre
    float k;        // fan: s_dist + d_dist
    float cr;       // det_count / 2 - 0.5  (detector coordinate of u = 0, in tap units)
    float half;     // n / 2 - 0.5
    uint32_t koff;  // 0x4B000000 * 8 mod 2^32, passed at run time so that ptxas keeps `base - koff` in one register
    int texq;       // tex_weights option: the interpolation fraction rounded to 8 bits (radon_common.cuh)
    float out_scale;  // tile kernel: what the accumulators are multiplied by on the way out -- 1 / det_spacing, and for fan
                      // beams the constant factor of the tap weights (k, or k s with fbp), which the taps leave out
    float inv_w;      // tile kernel, fan beam: 1 / that constant factor (the float64 fallback tap returns full weights)
    int fbp;        // fan beam only: weight every tap by (s / den) once more -- the 1 / U^2 of fan-beam FBP (Kak & Slaney 3.4.2)
};

// detector coordinate (in tap units: value = (1-fr) s[i0] + fr s[i0+1], i0 = floor(t)) and weight
// The global-memory path (variant 0, and the chunks of the tile kernel whose detector interval does not
// fit its shared segment: wide magnification near the source, coarse detectors) evaluates the detector
// coordinate in float64 and splits it into integer tap and float32 fraction, so it is as accurate as the
// tile-relative shared-memory path no matter how large the coordinate gets.
__device__ __forceinline__ float tap_global_inl(const AdjGeom& g, const float* __restrict__ row, float cs, float sn, float dx,
                                                float dy) {
    const double p = (double)cs * dx + (double)sn * dy;
    const double ids = 1.0 / (double)g.det_spacing;
    double t, w = 1.0;
    if (g.fan) {
        const double den = (double)g.s_dist + (double)sn * dx - (double)cs * dy;
        w = ((double)g.s_dist + (double)g.d_dist) / den;
        t = p * ids * w + (double)g.cr;
        if (g.fbp) w *= (double)g.s_dist / den;
    } else {
        t = p * ids + (double)g.cr;
    }
    const double tf = floor(t);
    if (!(tf > -2.0 && tf < (double)g.det_count)) return 0.f;       // both taps outside the detector (or NaN)
    float fr = (float)(t - tf);
    if (g.texq) fr = texq(fr);
    const int i0 = (int)tf, D = g.det_count;
    const float s0 = (unsigned)i0 < (unsigned)D ? __ldg(row + i0) : 0.f;
    const float s1 = (unsigned)(i0 + 1) < (unsigned)D ? __ldg(row + i0 + 1) : 0.f;
    return (float)w * fmaf(fr, s1 - s0, s0);
}
// Inlined into the parallel-beam tile kernel (no stack frame, no spill: ptxas -v), called from the fan-beam one
// (there the inlined float64 sequence would push the tap loop's registers over the 5-CTAs-per-SM cap).
__device__ __noinline__ float tap_global_call(const AdjGeom& g, const float* __restrict__ row, float cs, float sn, float dx,
                                              float dy) {
    return tap_global_inl(g, row, cs, sn, dx, dy);
}
template <bool CALL>
__device__ __forceinline__ float tap_global(const AdjGeom& g, const float* __restrict__ row, float cs, float sn, float dx, float dy) {
    return CALL ? tap_global_call(g, row, cs, sn, dx, dy) : tap_global_inl(g, row, cs, sn, dx, dy);
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
                acc[k] += tap_global<true>(g, row, cs, sn, dx, dy);
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

// Per-view constants of one CTA, tile-relative and set up in float64 so that the per-tap float32
// arithmetic only ever rounds at the magnitude of the tile (< SEG), not of the detector (< D):
//   parallel:  t - lo = base + a lx + b ly
//   fan:       t - lo = (n0 + nx lx + ny ly) / den,   den = d0 + sn lx - cs ly,   weight = k / den
// (lx, ly) = pixel offset inside the tile, lo = first detector bin of the staged segment.
// parallel beam: cap the registers so that 7 (PY = 8) / 3 (PY = 4) CTAs fit an SM -- B N^2 / PY threads then
// make one balanced wave; the rarely taken float64 fallback is what would otherwise raise the count.
// fan beam: also 7 CTAs (72 registers; ptxas spills 40 bytes in the per-chunk float64 set-up, none in the tap loop).
// At the 96 registers of 5 CTAs the kernel ran at 28 % warps active and 66 % issue: measured on the configs[2] share
// (2048 CTAs), 5 / 6 / 7 / 8 CTAs per SM: 1033 / 985 / 955 / 969 us -- 7 also makes it two even waves
// The segment holds the line through the two samples in the segment's own coordinate, (A, B) with
//   tap(t) = A + t B, A = value - c B for entry c.  No fraction is needed, only floor(t) for the index, and the byte
//   address comes from the magic-number bit pattern with one IMAD (the constant exponent part is folded into the
//   per-view base, mod 2^32): 4.5 instructions per tap.  A is rounded at the magnitude of c |B| (c < SEG = 96),
//   i.e. an error of <= 6e-6 |B| per tap, independent from tap to tap.
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// Fan-beam taps in line form (MODE 1) for the PY pixels of one thread and one view.
// SAFE: every corner of the tile is at least 8 pixels in front of the source (checked per chunk in float64), so
// MUFU.RCP's 1 ulp is enough (t < 96: < 1.2e-5 bins) and the quotient cannot leave the staged interval -- no
// Newton step, no clamps.  Otherwise one Newton step and a clamp onto the segment, as in MODE 0.
// MUFU.RCP alone: __fdividef(1, d) wraps it in range scaling (FSETP + 2 FSEL + 2 FMUL per call) that the
// denominators here (distance from the source, 1e-3 .. 1e5) never need
__device__ __forceinline__ float rcp_approx(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// Segments wider than 96 entries (fine detectors) hold the lines in a coordinate CENTRED on the segment, t' = t - SEG / 2:
// A = value - (c - SEG / 2) B is then rounded at the magnitude of 96 |B| as in the 96-entry segment, not of 192 |B|
// (uncentred, a 16-view white-noise backprojection at det_spacing 0.25 measured 1.003e-5 against the 1e-5 budget).  The
// shift is folded into the per-view constants (float64 set-up) and into the address constant; the floor trick uses
// 1.5 x 2^23 so that negative coordinates stay in the binade with ulp 1.
// (r02 random sweep, tools/fuzz_parity.py: fan beams through the 192-entry segment sit at 7e-6 median / 1.7e-5 worst on
//  few-view white-noise sinograms.  Keeping (value, difference) pairs and taking the fraction from the fused residual
//  (num - floor(c) den) / den removes the line offset's and the quotient's rounding: 7.3e-6 -> 5.9e-6 median,
//  1.3e-5 -> 1.2e-5 worst, for 293 -> 367 us at s_dist = n -- not kept: what remains is float32 evaluating num and den
//  themselves at the magnitude |c| den, which only float64 coordinates (radon_adj_variant 0) remove.)
__host__ __device__ constexpr int adj_seg_centre(int seg) { return seg > 96 ? seg / 2 : 0; }

// FBP (fan-beam FBP's second 1 / den, Kak & Slaney 3.4.2) is a template parameter: as a run-time uniform branch it was
// compiled to 16 predicated instructions per view and thread that the plain adjoint issued for nothing (19 % of its
// tap loop, ncu r02)
template <int PY, int SEG, bool SAFE, bool TQ, bool FBP>
__device__ __forceinline__ void fan_taps_line(const float* __restrict__ view, uint32_t cbase, float lx,
                                              const ull* __restrict__ ly_pk, float* __restrict__ acc) {
    constexpr int CEN = adj_seg_centre(SEG);
    constexpr float MAGIC = CEN ? 12582912.f : 8388608.f;       // 1.5 x 2^23 when the coordinate can be negative (centred)
    const float4 v = *reinterpret_cast<const float4*>(view);
    const float2 tr = *reinterpret_cast<const float2*>(view + 4);
    const float nx_ = fmaf(v.y, lx, v.x), dx_ = fmaf(tr.x, lx, v.w);
    const ull p_nx = pk2(nx_, nx_), p_dx = pk2(dx_, dx_);
    const ull p_ny = pk2(v.z, v.z), p_dy = pk2(-tr.y, -tr.y);
    const ull p_one = pk2(1.f, 1.f), p_m = pk2(MAGIC, MAGIC);
#pragma unroll
    for (int k = 0; k < PY; k += 2) {
        const ull p_num = fma2(p_ny, ly_pk[k / 2], p_nx);
        const ull p_den = fma2(p_dy, ly_pk[k / 2], p_dx);
        float d0, d1;
        upk2(p_den, d0, d1);
        ull p_r = pk2(rcp_approx(d0), rcp_approx(d1));
        if (!SAFE) p_r = fma2(p_r, sub2(p_one, mul2(p_den, p_r)), p_r);
        ull p_c = mul2(p_num, p_r);
        if (TQ) p_c = sub2(add2(p_c, pk2(TEXQ_MAGIC, TEXQ_MAGIC)), pk2(TEXQ_MAGIC, TEXQ_MAGIC));
        float c0, c1;
        upk2(p_c, c0, c1);
        if (!SAFE) {
            c0 = fminf(fmaxf(c0, (float)-CEN), (float)(SEG - 1 - CEN));
            c1 = fminf(fmaxf(c1, (float)-CEN), (float)(SEG - 1 - CEN));
            p_c = pk2(c0, c1);
        }
        float t0, t1, w0, w1;
        upk2(add2_rm(p_c, p_m), t0, t1);
        // tap weight k / den (fan-beam FBP: k s / den^2) without its constant factor: the kernel multiplies the
        // accumulators by it once on the way out (AdjGeom::out_scale) instead of once (twice) per pixel pair and view
        const ull p_w = FBP ? mul2(p_r, p_r) : p_r;
        upk2(p_w, w0, w1);
        const float2 s0 = lds64((uint32_t)__float_as_int(t0) * 8u + cbase);
        const float2 s1 = lds64((uint32_t)__float_as_int(t1) * 8u + cbase);
        acc[k] = fmaf(w0, fmaf(c0, s0.y, s0.x), acc[k]);
        acc[k + 1] = fmaf(w1, fmaf(c1, s1.y, s1.x), acc[k + 1]);
    }
}

template <int TX, int TY, int PY, int AC, int SEG, bool FAN, bool TQ, bool FBP>
__global__ void __launch_bounds__(TX*(TY / PY), FAN ? (PY == 8 ? 7 : 2) : (PY == 8 ? 7 : 3))
    radon_adj_tile_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ trig,
                          const AdjGeom g) {
    constexpr int THREADS = TX * (TY / PY);
    constexpr int RY = TY / PY;
    // staged detector rows as (value, next - value) pairs: one 8-byte load per tap
    __shared__ __align__(16) float2 s_seg[AC][SEG];
    __shared__ __align__(16) float s_view[AC][FAN ? 8 : 4];
    __shared__ int s_lo[AC];
    __shared__ int s_big;
    __shared__ int s_close;   // fan: some view of the chunk has the source within 8 pixels of the tile

    const int tid = threadIdx.y * TX + threadIdx.x;
    const int x = blockIdx.x * TX + threadIdx.x;
    const int y0 = blockIdx.y * TY + threadIdx.y;
    const int b = blockIdx.z;
    const float lx = (float)threadIdx.x, ly0 = (float)threadIdx.y;
    const float* sb = sino + (long)b * g.n_angles * g.det_count;

    // tile origin and far corner (clamped to the image so that padding pixels do not widen the interval)
    const double ox = (double)(blockIdx.x * TX) - (double)g.half, oy = (double)(blockIdx.y * TY) - (double)g.half;
    const double ex = (double)(min(blockIdx.x * TX + TX, g.n) - 1 - blockIdx.x * TX);
    const double ey = (double)(min(blockIdx.y * TY + TY, g.n) - 1 - blockIdx.y * TY);

    float acc[PY];
#pragma unroll
    for (int k = 0; k < PY; ++k) acc[k] = 0.f;
    ull acc2[PY / 2];        // parallel beam: packed accumulators of the pixel pairs
#pragma unroll
    for (int k = 0; k < PY / 2; ++k) acc2[k] = pk2(0.f, 0.f);
    constexpr int CEN = adj_seg_centre(SEG);
    constexpr float MAGIC = CEN ? 12582912.f : 8388608.f;
    static_assert(PY % 2 == 0, "pixels are processed in packed pairs");
    ull ly_pk[PY / 2];      // row offsets inside the tile of this thread's pixel pairs, clamped to the image
#pragma unroll
    for (int k = 0; k < PY; k += 2)
        ly_pk[k / 2] = pk2(fminf(ly0 + (float)(k * RY), (float)ey), fminf(ly0 + (float)((k + 1) * RY), (float)ey));

    for (int a0 = 0; a0 < g.n_angles; a0 += AC) {
        const int na = min(AC, g.n_angles - a0);
        __syncthreads();   // previous chunk fully consumed
        if (tid == 0) { s_big = 0; s_close = 0; }
        __syncthreads();
        if (tid < na) {
            const double cs = (double)__ldg(trig + 2 * (a0 + tid)), sn = (double)__ldg(trig + 2 * (a0 + tid) + 1);
            const double ids = 1.0 / (double)g.det_spacing, cr = (double)g.cr;
            double tmin, tmax;
            if (!FAN) {
                const double a = cs * ids, bb = sn * ids;
                const double t00 = a * ox + bb * oy + cr;
                const double c1 = a * ex, c2 = bb * ey;
                tmin = t00 + fmin(c1, 0.0) + fmin(c2, 0.0);
                tmax = t00 + fmax(c1, 0.0) + fmax(c2, 0.0);
                const int lo = (int)floor(tmin) - 1;
                s_lo[tid] = lo;
                float* v = s_view[tid];
                v[0] = (float)a; v[1] = (float)bb; v[2] = (float)(t00 - (double)(lo + CEN)); v[3] = 0.f;
            } else {
                const double K = ids * ((double)g.s_dist + (double)g.d_dist);
                const double p0 = cs * ox + sn * oy, d0 = (double)g.s_dist + sn * ox - cs * oy;
                tmin = 1e300; tmax = -1e300;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double cx = (c & 1) ? ex : 0.0, cy = (c & 2) ? ey : 0.0;
                    const double den = d0 + sn * cx - cs * cy;
                    const double t = cr + K * (p0 + cs * cx + sn * cy) / den;
                    // the source must stay outside the tile, otherwise the projection is unbounded
                    tmin = den > 1e-3 ? fmin(tmin, t) : -1e300;
                    tmax = den > 1e-3 ? fmax(tmax, t) : 1e300;
                    if (!(den >= 8.0)) s_close = 1;
                }
                const bool sane = tmin > -1e8 && tmax < 1e8;
                const int lo = sane ? (int)floor(tmin) - 1 : 0;
                s_lo[tid] = lo;
                const double L = (double)(lo + CEN) - cr;
                float* v = s_view[tid];
                v[0] = (float)(K * p0 - L * d0); v[1] = (float)(K * cs - L * sn); v[2] = (float)(K * sn + L * cs);
                v[3] = (float)d0; v[4] = (float)sn; v[5] = (float)cs; v[6] = v[7] = 0.f;
            }
            // one tap of slack on either side for rounding; SEG - 2 is the last legal tap index
            if (!(tmin > -1e8 && tmax < 1e8) || (floor(tmax) + 2.0) - (floor(tmin) - 1.0) + 1.0 > (double)SEG) s_big = 1;
        }
        __syncthreads();
        const bool big = s_big != 0;
        const bool close = s_close != 0;
        if (!big) {
            {
                // one warp per view, lanes along the segment: each bin is loaded once and its right neighbour comes
                // from the next lane (SHFL) -- half the loads and none of the index arithmetic of the generic loop
                static_assert(SEG % 32 == 0, "segment = whole warps of bins");
                constexpr int NB = SEG / 32;
                const int lane = tid & 31;
                for (int al = tid >> 5; al < na; al += THREADS / 32) {
                    const int d0 = s_lo[al] + lane;
                    // one 64-bit address per view; the NB + 1 loads are that address + i * 128 bytes (never dereferenced
                    // outside the row: the bounds predicate guards each load)
                    const float* rp = sb + ((long)(a0 + al) * g.det_count + d0);
                    float v[NB + 1];
#pragma unroll
                    for (int i = 0; i <= NB; ++i) {
                        const int d = d0 + i * 32;
                        v[i] = ((unsigned)d < (unsigned)g.det_count && (i < NB || lane == 0)) ? __ldg(rp + i * 32) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        const float nxt = __shfl_sync(0xffffffffu, v[i + 1], 0);
                        const float dn = __shfl_down_sync(0xffffffffu, v[i], 1);
                        const float dv = (lane == 31 ? nxt : dn) - v[i];
                        const int c = i * 32 + lane;
                        s_seg[al][c] = make_float2(fmaf(-(float)(c - CEN), dv, v[i]), dv);
                    }
                }
            }
            __syncthreads();
            if (x < g.n) {
                // Packed FP32 (FFMA2 / FADD2.RM): two of the thread's PY pixels per instruction for the
                // coordinate, its floor and its fraction.  ly_pk holds the row offsets of the pixel pairs,
                // clamped to the image so that padding rows of an edge tile stay inside the staged interval
                // (no per-tap clamp).
#pragma unroll 2
                for (int al = 0; al < na; ++al) {
                    const float2* seg = s_seg[al];
                    if (!FAN) {
                        const float4 v = *reinterpret_cast<const float4*>(s_view[al]);
                        const float tx = fmaf(v.x, lx, v.z);
                        const ull p_tx = pk2(tx, tx), p_b = pk2(v.y, v.y), p_m = pk2(MAGIC, MAGIC);
                        {
                            // bits(t + 2^23) = 0x4B000000 + floor(t): fold the constant into the base (mod 2^32)
                            const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(seg) - g.koff;
#pragma unroll
                            for (int k = 0; k < PY; k += 2) {
                                ull p_tl = fma2(p_b, ly_pk[k / 2], p_tx);              // >= 1 by construction
                                if (TQ) p_tl = sub2(add2(p_tl, pk2(TEXQ_MAGIC, TEXQ_MAGIC)), pk2(TEXQ_MAGIC, TEXQ_MAGIC));
                                const ull p_t = add2_rm(p_tl, p_m);
                                float t0, t1, c0, c1;
                                upk2(p_t, t0, t1);
                                upk2(p_tl, c0, c1);
                                const float2 s0 = lds64((uint32_t)__float_as_int(t0) * 8u + cbase);
                                const float2 s1 = lds64((uint32_t)__float_as_int(t1) * 8u + cbase);
                                acc2[k / 2] = add2(acc2[k / 2], pk2(fmaf(c0, s0.y, s0.x), fmaf(c1, s1.y, s1.x)));
                            }
                        }
                    } else {
                        const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(seg) - g.koff;
                        if (close) fan_taps_line<PY, SEG, false, TQ, FBP>(s_view[al], cbase, lx, ly_pk, acc);
                        else fan_taps_line<PY, SEG, true, TQ, FBP>(s_view[al], cbase, lx, ly_pk, acc);
                    }
                }
            }
        } else if (x < g.n) {
            const float dx = (float)x - g.half;
            for (int al = 0; al < na; ++al) {
                const float cs = __ldg(trig + 2 * (a0 + al)), sn = __ldg(trig + 2 * (a0 + al) + 1);
                const float* row = sb + (long)(a0 + al) * g.det_count;
#pragma unroll
                for (int k = 0; k < PY; ++k) {
                    const float dy = (float)(y0 + k * RY) - g.half;
                    acc[k] += FAN ? tap_global<FAN>(g, row, cs, sn, dx, dy) * g.inv_w : tap_global<FAN>(g, row, cs, sn, dx, dy);
                }
            }
        }
    }
    const float dx = (float)x - g.half;
    if (!FAN) {
#pragma unroll
        for (int k = 0; k < PY; k += 2) {
            float a0, a1;
            upk2(acc2[k / 2], a0, a1);
            acc[k] += a0;
            acc[k + 1] += a1;
        }
    }
#pragma unroll
    for (int k = 0; k < PY; ++k) {
        const int y = y0 + k * RY;
        if (x < g.n && y < g.n) {
            const float dy = (float)y - g.half;
            const float r2 = dx * dx + dy * dy, lim = 0.25f * (float)g.n * (float)g.n;
            const float v = (g.clip && r2 > lim) ? 0.f : acc[k] * g.out_scale;
            img[((long)b * g.n + y) * g.n + x] = v;
        }
    }
}

}  // namespace pdu

using namespace pdu;

extern "C" int pdu_radon_adj_weighted_f32(const float* sino, float* img, const float* trig, int batch,
                                          const pdu_radon_geom_t* g, int fbp_weight, void* workspace, size_t workspace_bytes,
                                          pdu_stream_t stream);

extern "C" int pdu_radon_adj_f32(const float* sino, float* img, const float* trig, int batch,
                                 const pdu_radon_geom_t* g, void* workspace, size_t workspace_bytes,
                                 pdu_stream_t stream) {
    return pdu_radon_adj_weighted_f32(sino, img, trig, batch, g, 0, workspace, workspace_bytes, stream);
}

extern "C" int pdu_radon_adj_weighted_f32(const float* sino, float* img, const float* trig, int batch,
                                          const pdu_radon_geom_t* g, int fbp_weight, void* workspace, size_t workspace_bytes,
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
    PDU_CHECK_DEVICE("pdu_radon_adj_f32");

    AdjGeom ag;
    ag.n = g->n;
    ag.n_angles = g->n_angles;
    ag.det_count = g->det_count;
    ag.fan = g->geom == PDU_GEOM_FAN;
    ag.clip = g->clip_to_circle;
    ag.ids = 1.f / g->det_spacing;
    ag.det_spacing = g->det_spacing;
    ag.d_dist = g->d_dist;
    ag.s_dist = g->s_dist;
    ag.k = g->s_dist + g->d_dist;
    ag.cr = 0.5f * (float)g->det_count - 0.5f;
    ag.half = 0.5f * (float)g->n - 0.5f;
    ag.koff = 0x4B000000u * 8u;
    ag.fbp = (fbp_weight && ag.fan) ? 1 : 0;
    {
        const double wc = ag.fan ? (double)ag.k * (ag.fbp ? (double)ag.s_dist : 1.0) : 1.0;
        ag.out_scale = (float)((double)ag.ids * wc);
        ag.inv_w = (float)(1.0 / wc);
    }
    ag.texq = option(OPT_TEX_WEIGHTS) > 0 ? 1 : 0;

    cudaStream_t st = (cudaStream_t)stream;
    int variant = option(OPT_RADON_ADJ);
    if (variant != 0) variant = 3;
    constexpr int TX = 32, TY = 32;
    dim3 grid((unsigned)cdiv(g->n, TX), (unsigned)cdiv(g->n, TY), (unsigned)batch);
    if (variant == 0) {
        radon_adj_gather_kernel<TX, TY, 4><<<grid, dim3(TX, TY / 4), 0, st>>>(sino, img, trig, ag);
        note_kernel(OP_RADON_ADJ, "radon_adj_gather_kernel grid %ux%ux%u (float64 taps through L1)", grid.x, grid.y, grid.z);
    } else {
        // line-form taps, 8 pixels per thread, 128 threads: B N^2 / 8 threads fit one balanced wave.
        // a 32 x 32 tile projects onto at most 32 sqrt(2) / det_spacing bins: a 64-entry segment is enough for
        // parallel beams with det_spacing >= 0.8 (a third less staging work)
        const bool seg64 = !ag.fan && 46.f * ag.ids + 5.f <= 64.f;
        // fine detectors (det_spacing < 0.5, or a fan beam's magnification): the tile projects onto more than 96 bins and
        // every chunk would take the float64 global-load path (measured, 16 x 256^2 x 512 views: det_spacing 0.4 839 us,
        // 0.25 1106 us against 180 us at 1.0).  A 192-entry segment with 16 views per chunk (same shared memory) keeps
        // those geometries on the shared-memory taps.
        const float mag = ag.fan ? ag.k / fmaxf(ag.s_dist - 0.7072f * (float)ag.n, 1.f) : 1.f;
        const bool seg192 = 45.3f * ag.ids * mag + 4.f > 96.f;     // (32 sqrt 2 bins per unit spacing + the two slack taps and rounding)
        const dim3 blk(TX, TY / 8);
#define PDU_ADJ_LAUNCH(AC_, SEG_, FAN_, TQ_, FBP_) \
    radon_adj_tile_kernel<TX, TY, 8, AC_, SEG_, FAN_, TQ_, FBP_><<<grid, blk, 0, st>>>(sino, img, trig, ag)
#define PDU_ADJ_DISPATCH(TQ_)                                                   \
    do {                                                                        \
        if (ag.fan && ag.fbp) {                                                 \
            if (seg192) PDU_ADJ_LAUNCH(16, 192, true, TQ_, true);               \
            else PDU_ADJ_LAUNCH(32, 96, true, TQ_, true);                       \
        } else if (ag.fan) {                                                    \
            if (seg192) PDU_ADJ_LAUNCH(16, 192, true, TQ_, false);              \
            else PDU_ADJ_LAUNCH(32, 96, true, TQ_, false);                      \
        } else if (seg64) PDU_ADJ_LAUNCH(32, 64, false, TQ_, false);            \
        else if (seg192) PDU_ADJ_LAUNCH(16, 192, false, TQ_, false);            \
        else PDU_ADJ_LAUNCH(32, 96, false, TQ_, false);                         \
    } while (0)
        if (seg192 && !seg64)      // centred coordinate: bits(t' + 1.5 x 2^23) * 8 + base - koff = base + 8 (floor(t') + 96)
            ag.koff = 0x4B400000u * 8u - 8u * (uint32_t)adj_seg_centre(192);
        if (ag.texq) PDU_ADJ_DISPATCH(true);
        else PDU_ADJ_DISPATCH(false);
#undef PDU_ADJ_DISPATCH
#undef PDU_ADJ_LAUNCH
        note_kernel(OP_RADON_ADJ, "radon_adj_tile_kernel<32,32,8,%d,%d,%s> grid %ux%ux%u (line-form taps in shared memory, packed FP32)",
                    seg192 && !seg64 ? 16 : 32, seg64 ? 64 : (seg192 ? 192 : 96), ag.fan ? "fan" : "parallel", grid.x, grid.y, grid.z);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}
