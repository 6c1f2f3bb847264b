// Forward projection (ray driven), parallel and fan beam.  Replaces [RECALL] torch_radon
// radon_forward_kernel, which samples a layered CUDA texture; here the image is staged in shared
// memory by TMA and interpolated in exact float32.
//
// Strip marching (both tile formats).  A CTA owns DB detectors x AG neighbouring views of one slice.  Rays are
//   walked in the direction of increasing major coordinate (rows for mostly-vertical rays; for
//   mostly-horizontal views the CTA reads a transposed copy of the slice so the same code applies) through
//   horizontal strips of TH rows.  Before marching, every ray records its column extent in every strip it
//   crosses (warp-reduced, then shared-memory atomicMin/Max by one lane), which fixes one box per strip; a
//   producer streams those boxes with cp.async.bulk.tensor (zero fill outside the image = the texture
//   "border" mode) through an NBUF-deep mbarrier ring while the warps sample the current box.  A strip whose
//   extent does not fit the box falls back to global loads for that strip only.
// variants 9 / 11 / 13 -- cell ("quad") tiles: one float4 per bilinear cell, one LDS.128 + four FMAs per sample
//   (radon_fwd_quad_kernel; shape 13 is the default for dense view sets, 9 / 11 serve sparser ones when the float
//   tile cannot be used).
// variant 1 -- float tiles: (TH+1) x W floats, four LDS.32 per sample, boxes up to 248 columns wide
//   (radon_fwd_strip_kernel with its own shape heuristic, the default for sparse view sets).
// variant 0 -- one thread per ray, bilinear taps through L1 (__ldg).  Kept as the A/B baseline and for
//   shapes neither tensor map can describe.
// Every mbarrier wait is bounded; a time-out is reported through the device error word (common.cuh) and the CTA gives
// up instead of sampling an unfilled tile -- the next library call returns PDU_ECUDA.
// The dispatch and the measurements behind it: pdu_radon_fwd_f32 at the end of this file, DESIGN.md 3.1.
#include <cuda.h>

#include "radon_common.cuh"

namespace pdu {

// ------------------------------------------------------------------ small kernels
__global__ void trig_kernel(const float* __restrict__ angles, float* __restrict__ trig, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double s, c;
        sincos((double)angles[i], &s, &c);
        trig[2 * i] = (float)c;
        trig[2 * i + 1] = (float)s;
    }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
    __shared__ float t[32][33];
    const long base = (long)blockIdx.z * n * n;
    int x = blockIdx.x * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int y = blockIdx.y * 32 + r;
        if (x < n && y < n) t[r][threadIdx.x] = in[base + (long)y * n + x];
    }
    __syncthreads();
    int ox = blockIdx.y * 32 + threadIdx.x;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int oy = blockIdx.x * 32 + r;
        if (ox < n && oy < n) out[base + (long)oy * n + ox] = t[threadIdx.x][r];
    }
}

// ------------------------------------------------------------------ variant 0
__global__ void __launch_bounds__(256) radon_fwd_gather_kernel(const float* __restrict__ img, float* __restrict__ sino,
                                                               const float* __restrict__ trig, pdu_radon_geom_t g, int tq) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y * blockDim.y + threadIdx.y;
    const int b = blockIdx.z;
    if (d >= g.det_count || a >= g.n_angles) return;
    const float cs = __ldg(trig + 2 * a), sn = __ldg(trig + 2 * a + 1);
    const Ray r = ray_setup(g, cs, sn, d);
    float acc = 0.f;
    const float* src = img + (long)b * g.n * g.n;
    for (int j = 0; j <= r.n_steps; ++j) {
        const float jf = (float)j;
        acc += bilinear_global(src, g.n, fmaf(jf, r.vx, r.xc0), fmaf(jf, r.vy, r.yc0), tq != 0);
    }
    sino[((long)b * g.n_angles + a) * g.det_count + d] = acc * r.step;
}

// ------------------------------------------------------------------ variant 1
constexpr int MAX_STRIPS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait (common.cuh): false on time-out -- the caller reports and abandons the tile.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, unsigned long long timeout_ns) {
    return mbar_wait_bounded(smem_u32(bar), parity, timeout_ns);
}
// what the kernels need to report a pipeline failure (and, for the tests, to provoke one)
struct FaultCtl {
    int* err_word;                    // device_error_word(), may be null
    unsigned long long timeout_ns;
    int fault;                        // debug_fault option: the producer skips its TMA loads
    int texq;                         // tex_weights option: interpolation fractions rounded to 8 bits (radon_common.cuh)
};
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

template <int DB, int AG, int TH, int W, int NBUF>
struct FwdCfg {
    static constexpr int THREADS = DB * AG;
    static constexpr int ROWS = TH + 1;
    static constexpr int TILE_BYTES = ROWS * W * 4;
    static constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;
    static constexpr int SMEM = NBUF * TILE_STRIDE + 2 * NBUF * 8 + 2 * MAX_STRIPS * 4;
    // W % 32 != 0 on purpose: the lanes of a warp sit in neighbouring rows of the tile, and a row
    // pitch that is a multiple of 32 floats would put the same column of every row in one bank
    static_assert(W % 4 == 0 && W % 32 != 0 && W <= 256, "box width: 16-byte multiple, not a multiple of 32 banks");
};

template <int DB, int AG, int TH, int W, int NBUF, int LD, bool TEXQ>
__global__ void __launch_bounds__(DB* AG)
    radon_fwd_strip_kernel(const __grid_constant__ CUtensorMap tm_img, const __grid_constant__ CUtensorMap tm_imgT,
                           const float* __restrict__ img, const float* __restrict__ imgT, float* __restrict__ sino,
                           const float* __restrict__ trig, const pdu_radon_geom_t g, const FaultCtl fc) {
    using C = FwdCfg<DB, AG, TH, W, NBUF>;
    // indexed directly (no re-aligned generic pointer) so that the tile reads compile to LDS
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    uint64_t* full = (uint64_t*)(smem_dyn + NBUF * C::TILE_STRIDE);
    uint64_t* empty = full + NBUF;          // one arrival per warp: the buffer may be refilled
    int* s_umin = (int*)(empty + NBUF);
    int* s_umax = s_umin + MAX_STRIPS;

    const int tid = threadIdx.x;
    // A warp is LD neighbouring detectors x 32/LD neighbouring views.  With LD = 32 the lanes of one
    // load span up to 45 tile columns (detector pitch 1/cos) and wrap the 32 banks: a 2-way conflict on
    // every load.  Narrower runs of several views keep the footprint inside 32 columns; the views'
    // samples fall on the same or neighbouring addresses (broadcast, not conflict).
    static_assert(32 % LD == 0 && DB % LD == 0 && AG % (32 / LD) == 0, "warp shape must tile the CTA");
    constexpr int AGW = 32 / LD;
    const int lane = tid & 31, warp = tid >> 5;
    const int dl = (warp % (DB / LD)) * LD + lane % LD;
    const int al = (warp / (DB / LD)) * AGW + lane / LD;
    const int d = blockIdx.x * DB + dl;
    const int a = blockIdx.y * AG + al;
    const int b = blockIdx.z;
    const int N = g.n;
    const int n_strips = (N + 1 + TH - 1) / TH;
    const bool valid = d < g.det_count && a < g.n_angles;
    // TMA wants a 128-byte aligned destination; if the runtime ever places dynamic shared memory
    // otherwise every strip takes the global-load path (still correct)
    const bool tma_ok = (smem_u32(smem_dyn) & 127u) == 0;

    for (int k = tid; k < n_strips; k += C::THREADS) {
        s_umin[k] = INT_MAX;
        s_umax[k] = INT_MIN;
    }
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, C::THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // orientation of the whole CTA from its first view: rays run along (sn, -cs)
    const int a_ref = min(blockIdx.y * AG, g.n_angles - 1);
    const bool use_t = fabsf(__ldg(trig + 2 * a_ref + 1)) > fabsf(__ldg(trig + 2 * a_ref));

    Ray r;
    r.n_steps = -1;
    r.xc0 = r.yc0 = r.vx = r.vy = r.step = 0.f;
    if (valid) r = ray_setup(g, __ldg(trig + 2 * a), __ldg(trig + 2 * a + 1), d);
    // (u, w) = (minor, major) coordinates; w indexes the rows of the source the CTA reads
    const float u0 = use_t ? r.yc0 : r.xc0, w0 = use_t ? r.xc0 : r.yc0;
    const float vu = use_t ? r.vy : r.vx, vw = use_t ? r.vx : r.vy;
    const int n = r.n_steps;
    // walk in the direction of increasing w: the s-th visited sample is j = jf, jf += dj
    const bool rev = vw < 0.f;
    const float dj = rev ? -1.f : 1.f;
    float jf = rev ? (float)n : 0.f;
    const float aw = fabsf(vw);
    const float inv_aw = aw > 1e-3f ? __fdividef(1.f, aw) : 0.f;   // 0: no estimate, count by stepping

    __syncthreads();
    {
        // (moving this registration into a once-per-call kernel, as quad_boxes_kernel does for the cell kernels, was
        //  measured here too: 529 -> 531 us at configs[1], so the float-tile kernel keeps it in place)
        // every ray registers its column extent in every strip it crosses.  The lanes of a warp mostly
        // cross the same strips, so the extents are first reduced across the warp (REDUX) and one lane
        // updates the shared box: per-lane atomics on one address serialise 32-way (ncu r02: a third of
        // all shared-memory wavefronts of the kernel)
        const float wa = fmaf(jf, vw, w0), ua = fmaf(jf, vu, u0);
        const float jend = rev ? 0.f : (float)n;
        const float wb = fmaf(jend, vw, w0), ub = fmaf(jend, vu, u0);
        // strips are [k TH - 1, (k+1) TH - 1); widen by eps so that a sample the marching loop puts on
        // the other side of a boundary (it rounds differently) still finds its ray registered there
        constexpr float EPS = 1e-3f;
        const int kA = n >= 0 ? max(((int)floorf(wa - EPS) + 1) / TH, 0) : INT_MAX;
        const int kB = n >= 0 ? min(((int)floorf(wb + EPS) + 1) / TH, n_strips - 1) : -1;
        const float dw = wb - wa;
        const float slope = dw > 0.f ? (ub - ua) / dw : 0.f;
        const int kA_w = __reduce_min_sync(0xffffffffu, kA), kB_w = __reduce_max_sync(0xffffffffu, kB);
        for (int k = kA_w; k <= kB_w; ++k) {
            int lo_k = INT_MAX, hi_k = INT_MIN;
            if (k >= kA && k <= kB) {
                const float wlo = fminf(fmaxf(wa, (float)(k * TH - 1)), wb);
                const float whi = fmaxf(fminf(wb, (float)((k + 1) * TH - 1)), wa);
                const float ulo = dw > 0.f ? fmaf(wlo - wa, slope, ua) : fminf(ua, ub);
                const float uhi = dw > 0.f ? fmaf(whi - wa, slope, ua) : fmaxf(ua, ub);
                lo_k = (int)floorf(fminf(ulo, uhi)) - 1;
                hi_k = (int)floorf(fmaxf(ulo, uhi)) + 2;
            }
            lo_k = __reduce_min_sync(0xffffffffu, lo_k);
            hi_k = __reduce_max_sync(0xffffffffu, hi_k);
            if (lane == 0 && lo_k <= hi_k) {
                atomicMin(&s_umin[k], lo_k);
                atomicMax(&s_umax[k], hi_k);
            }
        }
    }
    __syncthreads();

    const CUtensorMap* tm = use_t ? &tm_imgT : &tm_img;
    const float* src = (use_t ? imgT : img) + (long)b * N * N;

    int k_issue = 0, seq_issue = 0;   // producer state (thread 0)
    auto issue = [&]() -> bool {
        while (k_issue < n_strips && s_umin[k_issue] > s_umax[k_issue]) ++k_issue;
        if (k_issue < n_strips) {
            const int buf = seq_issue % NBUF;
            // the buffer's previous strip must have been read by every warp (they run ahead of each other
            // by up to NBUF - 1 strips: there is no block-wide barrier in the marching loop)
            if (seq_issue >= NBUF && !mbar_wait(empty + buf, ((seq_issue / NBUF) - 1) & 1, fc.timeout_ns)) {
                report_device_error(fc.err_word, DEV_ERR_RADON_FWD);
                return false;                // stop producing: the consumers time out and give up as well
            }
            // TMA needs the innermost start coordinate on a 16-byte boundary (measured: any c0 % 4 != 0
            // raises "illegal instruction" on sm_100a, negative values are fine) -> round down to 4 floats
            const int lo = s_umin[k_issue] & ~3;
            if (tma_ok && s_umax[k_issue] - lo + 1 <= W) {
                mbar_expect_tx(full + buf, C::TILE_BYTES);
                if (!fc.fault) tma_load_3d(smem_dyn + buf * C::TILE_STRIDE, tm, lo, k_issue * TH - 1, b, full + buf);
            } else {
                mbar_arrive(full + buf);
            }
            ++seq_issue;
            ++k_issue;
        }
        return true;
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NBUF - 1; ++i) issue();
    }

    constexpr float MAGIC = 8388608.f;   // 2^23: x + MAGIC rounded down == floor(x) + MAGIC for 0 <= x < 2^23
    float acc = 0.f;
    int s = 0;          // samples of this ray already taken
    int seq = 0;
    for (int k = 0; k < n_strips; ++k) {
        const int hi = s_umax[k];
        if (s_umin[k] > hi) continue;
        const int lo = s_umin[k] & ~3;
        if (tid == 0 && !issue()) return;
        const int buf = seq % NBUF;
        if (!mbar_wait(full + buf, (seq / NBUF) & 1, fc.timeout_ns)) {
            report_device_error(fc.err_word, DEV_ERR_RADON_FWD);     // never sample an unfilled tile
            return;
        }
        // strip-local coordinates: both subtractions are exact (result is a multiple of the operands' ulp)
        const float w0l = w0 - (float)(k * TH - 1);
        const float u0l = u0 - (float)lo;
        // how many of the remaining samples fall in this strip: estimate, then settle with the very
        // predicate the addressing relies on (wl < TH)
        int cnt = 0;
        const int left = n - s + 1;
        if (left > 0) {
            const float wl = fmaf(jf, vw, w0l);
            if (wl < (float)TH) {
                int m = inv_aw > 0.f ? (int)(((float)TH - wl) * inv_aw) + 1 : 1;
                m = max(1, min(m, left));
                while (m > 1 && fmaf(jf + (float)(m - 1) * dj, vw, w0l) >= (float)TH) --m;
                while (m < left && fmaf(jf + (float)m * dj, vw, w0l) < (float)TH) ++m;
                cnt = m;
            }
        }
        if (tma_ok && hi - lo + 1 <= W) {
            const float* tile = (const float*)(smem_dyn + buf * C::TILE_STRIDE);
            int i = 0;
            {
                // Packed-FP32 inner loop (Blackwell FFMA2 / FADD2): the (w, u) coordinate pair, its floor
                // and fraction, and the two horizontal lerps each take one instruction for both lanes of
                // the pair; the tile is addressed straight from the magic-number bit patterns.
                // A first sample that lands an ulp before the strip (wl < 0) goes through the scalar
                // path below, which clamps it onto the first row.
                if (cnt > 0 && fmaf(jf, vw, w0l) < 0.f) i = -1;
                if (i == 0) {
                    const ull p_v = pk2(vw, vu), p_0 = pk2(w0l, u0l), p_m = pk2(MAGIC, MAGIC), p_dj = pk2(dj, dj);
                    ull p_j = pk2(jf, jf);
#pragma unroll 2
                    for (; i < cnt; ++i) {
                        const ull p_c = fma2(p_j, p_v, p_0);            // (wl, ul)
                        const ull p_t = add2_rm(p_c, p_m);              // (floor + 2^23) each
                        ull p_f = sub2(p_c, sub2(p_t, p_m));            // (fw, fu)
                        if (TEXQ) p_f = sub2(add2(p_f, pk2(TEXQ_MAGIC, TEXQ_MAGIC)), pk2(TEXQ_MAGIC, TEXQ_MAGIC));
                        float tw, tu, fw, fu;
                        upk2(p_t, tw, tu);
                        upk2(p_f, fw, fu);
                        // bits(x + 2^23) = 0x4B000000 | floor(x): the exponent parts of both terms only reach bit 24
                        // and above, so the low 24 bits of the product-sum are exactly iw * W + iu
                        const uint32_t idx = ((uint32_t)__float_as_int(tw) * (uint32_t)W + (uint32_t)__float_as_int(tu)) & 0xFFFFFFu;
                        const float* p = tile + idx;
                        const float v00 = p[0], v01 = p[1], v10 = p[W], v11 = p[W + 1];
                        const ull p_lo = pk2(v00, v10);
                        const ull p_tb = fma2(pk2(fu, fu), sub2(pk2(v01, v11), p_lo), p_lo);   // (top, bot)
                        float top, bot;
                        upk2(p_tb, top, bot);
                        acc += fmaf(fw, bot - top, top);
                        p_j = add2(p_j, p_dj);
                    }
                    float j_hi;
                    upk2(p_j, jf, j_hi);
                } else {
                    i = 0;
                }
            }
#pragma unroll 2
            for (; i < cnt; ++i) {
                const float wl = fmaxf(fmaf(jf, vw, w0l), 0.f);   // a sample an ulp before the strip clamps onto its first row
                const float ul = fmaf(jf, vu, u0l);
                const float tw = __fadd_rd(wl, MAGIC), tu = __fadd_rd(ul, MAGIC);
                const int iw = __float_as_int(tw) & 0x7fffff, iu = __float_as_int(tu) & 0x7fffff;
                float fw = wl - (tw - MAGIC), fu = ul - (tu - MAGIC);
                if (TEXQ) {
                    fw = texq(fw);
                    fu = texq(fu);
                }
                const float* p = tile + iw * W + iu;
                const float v00 = p[0], v01 = p[1], v10 = p[W], v11 = p[W + 1];
                const float top = fmaf(fu, v01 - v00, v00);
                const float bot = fmaf(fu, v11 - v10, v10);
                acc += fmaf(fw, bot - top, top);
                jf += dj;
            }
        } else {
            for (int i = 0; i < cnt; ++i) {
                acc += bilinear_global(src, N, fmaf(jf, vu, u0), fmaf(jf, vw, w0), TEXQ);
                jf += dj;
            }
        }
        s += cnt;
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + buf);     // this warp is done with the buffer
        ++seq;
    }
    if (valid) sino[((long)b * g.n_angles + a) * g.det_count + d] = acc * r.step;
}

// ------------------------------------------------------------------ variant 7+: "quad" strips
// Same strip marching, different tile format.  A pre-pass turns the slice (and its transpose) into
// one float4 per bilinear CELL -- (v00, d_w, d_u, d_uw) with d_w = v10 - v00, d_u = v01 - v00,
// d_uw = (v11 - v10) - (v01 - v00) -- so a sample is ONE 16-byte shared-memory load and four FMAs,
//   v00 + fu d_u + fw (d_w + fu d_uw),
// instead of four 4-byte loads, two subtractions and three FMAs, and the tile address comes straight
// from the magic-number bit patterns with two integer instructions (no mask: the constant high part
// is folded into the per-strip base, all arithmetic mod 2^32).  Cell (R, C) has its corners at pixel
// rows R-1, R and columns C-1, C, so the cell tensor is (n+1) x (n+1) and TMA's zero fill outside
// it is again the texture "border" mode.  The 16-byte cells also remove the 16-byte start alignment
// constraint of the float tile (any cell column is a legal TMA start).
__global__ void __launch_bounds__(256) quad_build_kernel(const float* __restrict__ img, float4* __restrict__ q,
                                                         float4* __restrict__ qt, int n) {
    __shared__ float t[33][34];
    const int n1 = n + 1;
    const long ibase = (long)blockIdx.z * n * n;
    const long qbase = (long)blockIdx.z * n1 * n1;
    const int R0 = blockIdx.y * 32, C0 = blockIdx.x * 32;   // cell tile; pixel rows R0-1 .. R0+31
    for (int i = threadIdx.y * 32 + threadIdx.x; i < 33 * 33; i += 256) {
        const int rr = i / 33, cc = i - rr * 33;
        const int y = R0 - 1 + rr, x = C0 - 1 + cc;
        t[rr][cc] = ((unsigned)y < (unsigned)n && (unsigned)x < (unsigned)n) ? __ldg(img + ibase + (long)y * n + x) : 0.f;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const int cc = threadIdx.x;
        {   // Q(R0 + rr, C0 + cc): rows = image rows
            const int R = R0 + rr, Cc = C0 + cc;
            if (R < n1 && Cc < n1) {
                const float v00 = t[rr][cc], v01 = t[rr][cc + 1], v10 = t[rr + 1][cc], v11 = t[rr + 1][cc + 1];
                q[qbase + (long)R * n1 + Cc] = make_float4(v00, v10 - v00, v01 - v00, (v11 - v10) - (v01 - v00));
            }
        }
        {   // QT(C0 + rr, R0 + cc): the same cells of the transposed image (rows = image columns)
            const int R = C0 + rr, Cc = R0 + cc;
            if (R < n1 && Cc < n1) {
                const float v00 = t[cc][rr], v01 = t[cc + 1][rr], v10 = t[cc][rr + 1], v11 = t[cc + 1][rr + 1];
                qt[qbase + (long)R * n1 + Cc] = make_float4(v00, v10 - v00, v01 - v00, (v11 - v10) - (v01 - v00));
            }
        }
    }
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// one cell straight from global memory (oversize strips only); (wc, uc) are cell coordinates
__device__ __forceinline__ float quad_global(const float4* __restrict__ q, int n1, float wc, float uc, bool tq) {
    const float wf = floorf(wc), uf = floorf(uc);
    const int R = (int)wf, Cc = (int)uf;
    if ((unsigned)R >= (unsigned)n1 || (unsigned)Cc >= (unsigned)n1) return 0.f;
    const float4 c = __ldg(q + (long)R * n1 + Cc);
    float fw = wc - wf, fu = uc - uf;
    if (tq) {
        fw = texq(fw);
        fu = texq(fu);
    }
    return fmaf(fu, c.z, c.x) + fw * fmaf(fu, c.w, c.y);
}

template <int DB, int AG, int TH, int W, int NBUF>
struct QuadCfg {
    static constexpr int THREADS = DB * AG;
    // TH + 1 cell rows: the last one is the first row of the next strip again.  With that slack row the number of
    // samples a ray takes in a strip is ONE estimate -- "those with local row < TH + 1/2" -- and rounding in the estimate
    // only moves a sample between a strip's slack row and the next strip's first row (the same cells); the r01 kernel
    // settled the count exactly with two fix-up loops per ray and strip and clamped samples an ulp before a strip.
    static constexpr int ROWS = TH + 1;
    static constexpr int TILE_BYTES = ROWS * W * 16;
    static constexpr int TILE_STRIDE = (TILE_BYTES + 127) / 128 * 128;
    static constexpr int SMEM = NBUF * TILE_STRIDE + 2 * NBUF * 8 + 2 * MAX_STRIPS * 4;
    static_assert(W <= 128, "one TMA box of 2 W doubles");
};

// the per-ray quantities of the cell-tile kernels: cell coordinates (pixel-centre coordinate + 1), (u, w) = (column, row)
// of the source the CTA reads, walked in the direction of increasing w
struct QuadRay {
    float u0, w0, vu, vw, jf, dj, step;
    int n;
    bool rev;
};
__device__ __forceinline__ QuadRay quad_ray(const pdu_radon_geom_t& g, const float* __restrict__ trig, int a, int d, bool valid, bool use_t) {
    Ray r;
    r.n_steps = -1;
    r.xc0 = r.yc0 = r.vx = r.vy = r.step = 0.f;
    if (valid) r = ray_setup(g, __ldg(trig + 2 * a), __ldg(trig + 2 * a + 1), d);
    QuadRay q;
    q.u0 = (use_t ? r.yc0 : r.xc0) + 1.f;
    q.w0 = (use_t ? r.xc0 : r.yc0) + 1.f;
    q.vu = use_t ? r.vy : r.vx;
    q.vw = use_t ? r.vx : r.vy;
    q.n = r.n_steps;
    q.rev = q.vw < 0.f;
    q.dj = q.rev ? -1.f : 1.f;
    q.jf = q.rev ? (float)q.n : 0.f;
    q.step = r.step;
    return q;
}
constexpr float QUAD_EPS = 1e-3f;
// first strip a ray takes samples from (INT_MAX: the ray misses the slice)
template <int TH>
__device__ __forceinline__ int quad_first_strip(const QuadRay& q) {
    const float wa = fmaf(q.jf, q.vw, q.w0);
    return q.n >= 0 ? max((int)floorf(wa - QUAD_EPS), 0) / TH : INT_MAX;
}

// Strip boxes of one CTA (DB detectors x AG views): for every TH-row strip the range of cell columns its rays touch.
// They depend on the geometry only, not on the slice: quad_boxes_kernel computes them once per call for the
// (detector block, view group) grid and every slice's CTA reads its row of the table (r01 / early r02: every CTA
// registered its own boxes -- 9.5 % of the kernel's instructions and two more CTA barriers, repeated for each slice).
// Warp-reduced extents, one lane per warp touches the shared box.  u is monotone along the ray, so its range inside
// strip k is [U_k, U_k+1] clamped to the ray's own range, with U_k = u at the strip boundary: one FMA, one clamp and
// one floor per boundary, shared by the two strips it separates.
template <int TH>
__device__ __forceinline__ void quad_register_boxes(const QuadRay& q, int n_strips, int lane, int* s_umin, int* s_umax) {
    const int n = q.n;
    const float wa = fmaf(q.jf, q.vw, q.w0), ua = fmaf(q.jf, q.vu, q.u0);
    const float jend = q.rev ? 0.f : (float)n;
    const float wb = fmaf(jend, q.vw, q.w0), ub = fmaf(jend, q.vu, q.u0);
    const int kA = quad_first_strip<TH>(q);
    const int kB = n >= 0 ? min(max((int)floorf(wb + QUAD_EPS), 0) / TH, n_strips - 1) : -1;
    const float dw = wb - wa;
    const bool lin = dw > 0.f;
    const float slope = lin ? (ub - ua) / dw : 0.f;
    const float umin = fminf(ua, ub), umax = fmaxf(ua, ub);
    const int kA_w = __reduce_min_sync(0xffffffffu, kA), kB_w = __reduce_max_sync(0xffffffffu, kB);
    // a strip takes this ray's samples with row coordinate in [k TH - 1/2, (k + 1) TH + 1/2) (slack row, see QuadCfg):
    // Us = u at k TH - 1/2, Ue = u at (k + 1) TH + 1/2 = (next strip's Us) + slope
    auto u_at = [&](float w) { return lin ? fminf(fmaxf(fmaf(w - wa, slope, ua), umin), umax) : umin; };
    float wk = (float)((kA_w <= kB_w ? kA_w : 0) * TH) - 0.5f;      // (no valid ray in the warp: kA_w == INT_MAX, loop empty)
    int Fs = (int)floorf(u_at(wk));
    for (int k = kA_w; k <= kB_w; ++k) {
        wk += (float)TH;
        const int Fs1 = (int)floorf(u_at(wk));
        const int Fe = (int)floorf(lin ? u_at(wk + 1.f) : umax);
        int lo_k = INT_MAX, hi_k = INT_MIN;
        if (k >= kA && k <= kB) {
            lo_k = min(Fs, Fe) - 1;
            hi_k = max(Fs, Fe) + 1;
        }
        lo_k = __reduce_min_sync(0xffffffffu, lo_k);
        hi_k = __reduce_max_sync(0xffffffffu, hi_k);
        if (lane == 0 && lo_k <= hi_k) {
            atomicMin(&s_umin[k], lo_k);
            atomicMax(&s_umax[k], hi_k);
        }
        Fs = Fs1;
    }
}

// thread -> (detector, view) of a DB x AG CTA whose warps are LD detectors x 32 / LD views
template <int DB, int AG, int LD>
__device__ __forceinline__ void quad_thread_ray(int tid, int& dl, int& al) {
    static_assert(32 % LD == 0 && DB % LD == 0 && AG % (32 / LD) == 0, "warp shape must tile the CTA");
    constexpr int AGW = 32 / LD;
    const int lane = tid & 31, warp = tid >> 5;
    dl = (warp % (DB / LD)) * LD + lane % LD;
    al = (warp / (DB / LD)) * AGW + lane / LD;
}

template <int DB, int AG, int TH, int LD>
__global__ void __launch_bounds__(DB* AG)
    quad_boxes_kernel(const float* __restrict__ trig, const pdu_radon_geom_t g, int2* __restrict__ boxes) {
    __shared__ int s_umin[MAX_STRIPS], s_umax[MAX_STRIPS];
    const int tid = threadIdx.x;
    int dl, al;
    quad_thread_ray<DB, AG, LD>(tid, dl, al);
    const int d = blockIdx.x * DB + dl, a = blockIdx.y * AG + al;
    const int n_strips = (g.n + 1 + TH - 1) / TH;
    for (int k = tid; k < n_strips; k += DB * AG) {
        s_umin[k] = INT_MAX;
        s_umax[k] = INT_MIN;
    }
    const int a_ref = min(blockIdx.y * AG, g.n_angles - 1);
    const bool use_t = fabsf(__ldg(trig + 2 * a_ref + 1)) > fabsf(__ldg(trig + 2 * a_ref));
    const QuadRay q = quad_ray(g, trig, a, d, d < g.det_count && a < g.n_angles, use_t);
    __syncthreads();
    quad_register_boxes<TH>(q, n_strips, tid & 31, s_umin, s_umax);
    __syncthreads();
    int2* dst = boxes + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * n_strips;
    for (int k = tid; k < n_strips; k += DB * AG) dst[k] = make_int2(s_umin[k], s_umax[k]);
}

template <int DB, int AG, int TH, int W, int NBUF, int LD, bool TEXQ>
__global__ void __launch_bounds__(DB* AG + 32, 5)      // 5: ptxas settles at 40 registers without a spill (4: 56 registers WITH one)
    radon_fwd_quad_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_qt,
                          const float4* __restrict__ q, const float4* __restrict__ qt, float* __restrict__ sino,
                          const float* __restrict__ trig, const int2* __restrict__ boxes, const pdu_radon_geom_t g,
                          const FaultCtl fc) {
    using C = QuadCfg<DB, AG, TH, W, NBUF>;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    uint64_t* full = (uint64_t*)(smem_dyn + NBUF * C::TILE_STRIDE);
    uint64_t* empty = full + NBUF;
    int* s_umin = (int*)(empty + NBUF);
    int* s_umax = s_umin + MAX_STRIPS;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // warp THREADS/32 is the TMA producer: it owns no rays, so no compute warp ever waits for the other
    // warps of the CTA (the `empty` barriers) before starting its own strip
    const bool producer = warp == C::THREADS / 32;
    int dl, al;
    quad_thread_ray<DB, AG, LD>(tid, dl, al);
    const int d = blockIdx.x * DB + dl;
    const int a = blockIdx.y * AG + al;
    const int b = blockIdx.z;
    const int N1 = g.n + 1;
    const int n_strips = (N1 + TH - 1) / TH;
    const bool valid = !producer && d < g.det_count && a < g.n_angles;
    const uint32_t smem_base = smem_u32(smem_dyn);
    const bool tma_ok = (smem_base & 127u) == 0;

    {   // this CTA's strip boxes (quad_boxes_kernel)
        const int2* bx = boxes + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * n_strips;
        for (int k = tid; k < n_strips; k += C::THREADS + 32) {
            const int2 v = __ldg(bx + k);
            s_umin[k] = v.x;
            s_umax[k] = v.y;
        }
    }
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, C::THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const int a_ref = min(blockIdx.y * AG, g.n_angles - 1);
    const bool use_t = fabsf(__ldg(trig + 2 * a_ref + 1)) > fabsf(__ldg(trig + 2 * a_ref));

    const QuadRay qr = quad_ray(g, trig, a, d, valid, use_t);
    const float u0 = qr.u0, w0 = qr.w0, vu = qr.vu, vw = qr.vw, dj = qr.dj;
    const int n = qr.n;
    float jf = qr.jf;
    const float aw = fabsf(vw);
    const float inv_aw = aw > 1e-3f ? __fdividef(1.f, aw) : 0.f;
    const int k_first = quad_first_strip<TH>(qr);     // the first strip this ray registered in: it takes no sample from an earlier one
    __syncthreads();

    const CUtensorMap* tm = use_t ? &tm_qt : &tm_q;
    const float4* src = (use_t ? qt : q) + (long)b * N1 * N1;

    int k_issue = 0, seq_issue = 0;   // producer state
    auto issue = [&]() -> bool {      // false: a consumer never released the buffer (time-out) -- stop producing
        while (k_issue < n_strips && s_umin[k_issue] > s_umax[k_issue]) ++k_issue;
        if (k_issue < n_strips) {
            const int buf = seq_issue % NBUF;
            if (seq_issue >= NBUF && !mbar_wait(empty + buf, ((seq_issue / NBUF) - 1) & 1, fc.timeout_ns)) {
                report_device_error(fc.err_word, DEV_ERR_RADON_FWD);
                return false;
            }
            const int lo = s_umin[k_issue];
            if (tma_ok && s_umax[k_issue] - lo + 1 <= W) {
                mbar_expect_tx(full + buf, C::TILE_BYTES);
                // the map describes the cells as pairs of doubles: start = 2 lo (always 16-byte aligned)
                if (!fc.fault) tma_load_3d(smem_dyn + buf * C::TILE_STRIDE, tm, 2 * lo, k_issue * TH, b, full + buf);
            } else {
                mbar_arrive(full + buf);
            }
            ++seq_issue;
            ++k_issue;
        }
        return true;
    };
    if (producer) {
        if (lane == 0)
            while (k_issue < n_strips && issue()) {}
        return;
    }

    constexpr float MAGIC = 8388608.f;
    // bits(x + 2^23) = 0x4B000000 + floor(x):  bits(tw) * 16 W + bits(tu) * 16 = 16 (iw W + iu) + KOFF  (mod 2^32)
    constexpr uint32_t KOFF = 0x4B000000u * (16u * (uint32_t)W) + 0x4B000000u * 16u;
    float acc0 = 0.f, acc1 = 0.f;
    int s = 0;
    int seq = 0;
    for (int k = 0; k < n_strips; ++k) {
        const int hi = s_umax[k];
        const int lo = s_umin[k];
        if (lo > hi) continue;
        const int buf = seq % NBUF;
        if (!mbar_wait(full + buf, (seq / NBUF) & 1, fc.timeout_ns)) {
            report_device_error(fc.err_word, DEV_ERR_RADON_FWD);     // never sample an unfilled tile
            return;
        }
        const float w0l = w0 - (float)(k * TH);
        const float u0l = u0 - (float)lo;
        // samples of this ray in this strip: those whose local row is below TH + 1/2 (the tile has the slack row)
        int cnt = 0;
        const int left = n - s + 1;
        if (left > 0 && k >= k_first) {
            const float room = ((float)TH + 0.5f) - fmaf(jf, vw, w0l);
            if (room > 0.f) cnt = inv_aw > 0.f ? min((int)(room * inv_aw) + 1, left) : left;
        }
        if (tma_ok && hi - lo + 1 <= W) {
            const uint32_t cbase = smem_base + (uint32_t)(buf * C::TILE_STRIDE) - KOFF;
            int i = 0;
            const ull p_v = pk2(vw, vu), p_0 = pk2(w0l, u0l), p_m = pk2(MAGIC, MAGIC), p_dj = pk2(dj, dj);
            ull p_j = pk2(jf, jf);
#pragma unroll 4
            for (; i < cnt; ++i) {
                const ull p_c = fma2(p_j, p_v, p_0);            // (wl, ul)
                const ull p_t = add2_rm(p_c, p_m);              // (floor + 2^23) each
                ull p_f = sub2(p_c, sub2(p_t, p_m));            // (fw, fu)
                if (TEXQ) p_f = sub2(add2(p_f, pk2(TEXQ_MAGIC, TEXQ_MAGIC)), pk2(TEXQ_MAGIC, TEXQ_MAGIC));
                float tw, tu, fw, fu;
                upk2(p_t, tw, tu);
                upk2(p_f, fw, fu);
                const uint32_t addr = (uint32_t)__float_as_int(tw) * (16u * (uint32_t)W) + cbase + ((uint32_t)__float_as_int(tu) << 4);
                const float4 c = lds128(addr);
                acc0 += fmaf(fu, c.z, c.x);
                acc1 = fmaf(fw, fmaf(fu, c.w, c.y), acc1);
                p_j = add2(p_j, p_dj);
            }
            float j_hi;
            upk2(p_j, jf, j_hi);
        } else {
            for (int i = 0; i < cnt; ++i) {
                acc0 += quad_global(src, N1, fmaf(jf, vw, w0), fmaf(jf, vu, u0), TEXQ);
                jf += dj;
            }
        }
        s += cnt;
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + buf);
        ++seq;
    }
    if (valid) sino[((long)b * g.n_angles + a) * g.det_count + d] = (acc0 + acc1) * qr.step;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = []() -> encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (encode_tiled_fn)p;
    }();
    return fn;
}

// [batch, n, n] float32, box (w, rows, 1), zero fill outside.
static int make_image_map(CUtensorMap* tm, const float* ptr, int batch, int n, int box_w, int box_rows) {
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return PDU_EUNSUPPORTED;
    }
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (n=%d batch=%d box=%dx%d)", (int)rc, n, batch, box_w,
                  box_rows);
        return PDU_ECUDA;
    }
    return PDU_OK;
}

static FaultCtl fault_ctl() {
    const int fault = option(OPT_DEBUG_FAULT) == 1 ? 1 : 0;
    return FaultCtl{device_error_word(), fault ? MBAR_TIMEOUT_FAULT_NS : MBAR_TIMEOUT_NS, fault, option(OPT_TEX_WEIGHTS) > 0 ? 1 : 0};
}

template <int DB, int AG, int TH, int W, int NBUF, int LD = 32>
static int launch_strip(const float* img, const float* imgT, float* sino, const float* trig, int batch,
                        const pdu_radon_geom_t& g, cudaStream_t st) {
    using C = FwdCfg<DB, AG, TH, W, NBUF>;
    CUtensorMap tm, tmT;
    int rc = make_image_map(&tm, img, batch, g.n, W, C::ROWS);
    if (rc) return rc;
    rc = make_image_map(&tmT, imgT, batch, g.n, W, C::ROWS);
    if (rc) return rc;
    const FaultCtl fc = fault_ctl();
    dim3 grid((unsigned)cdiv(g.det_count, DB), (unsigned)cdiv(g.n_angles, AG), (unsigned)batch);
    if (fc.texq) {
        PDU_CUDA((ensure_dyn_smem<radon_fwd_strip_kernel<DB, AG, TH, W, NBUF, LD, true>>(C::SMEM)));
        radon_fwd_strip_kernel<DB, AG, TH, W, NBUF, LD, true><<<grid, C::THREADS, C::SMEM, st>>>(tm, tmT, img, imgT, sino, trig, g, fc);
    } else {
        PDU_CUDA((ensure_dyn_smem<radon_fwd_strip_kernel<DB, AG, TH, W, NBUF, LD, false>>(C::SMEM)));
        radon_fwd_strip_kernel<DB, AG, TH, W, NBUF, LD, false><<<grid, C::THREADS, C::SMEM, st>>>(tm, tmT, img, imgT, sino, trig, g, fc);
    }
    PDU_LAUNCHED();
    note_kernel(OP_RADON_FWD, "transpose_kernel + radon_fwd_strip_kernel<%d,%d,%d,%d,%d,%d> grid %ux%ux%u (float tiles, TMA ring)", DB, AG,
                TH, W, NBUF, LD, grid.x, grid.y, grid.z);
    return PDU_OK;
}

// cells [batch, n+1, n+1] x 16 bytes, described as pairs of doubles: box (2 w, rows, 1), zero fill outside.
static int make_quad_map(CUtensorMap* tm, const float4* ptr, int batch, int n1, int box_w, int box_rows) {
    encode_tiled_fn enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return PDU_EUNSUPPORTED;
    }
    cuuint64_t dims[3] = {(cuuint64_t)n1 * 2, (cuuint64_t)n1, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)n1 * 16, (cuuint64_t)n1 * n1 * 16};
    cuuint32_t box[3] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)ptr, dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (cells) failed with CUresult %d (n=%d batch=%d box=%dx%d)", (int)rc, n1 - 1,
                  batch, box_w, box_rows);
        return PDU_ECUDA;
    }
    return PDU_OK;
}

static size_t quad_bytes(int n, int batch) { return (size_t)2 * batch * (n + 1) * (n + 1) * sizeof(float4); }

// strip-box table of the cell-tile kernels: one row of n_strips (lo, hi) pairs per (detector block, view group);
// sized for the smallest CTA shape the dispatcher uses (32 detectors x 4 views, 16-row strips)
static size_t quad_boxes_bytes(const pdu_radon_geom_t& g) {
    return (size_t)cdiv(g.det_count, 32) * cdiv(g.n_angles, 4) * ((g.n + 1 + 15) / 16) * sizeof(int2);
}

template <int DB, int AG, int TH, int W, int NBUF, int LD>
static int launch_quad(const float4* q, const float4* qt, float* sino, const float* trig, int2* boxes, int batch,
                       const pdu_radon_geom_t& g, cudaStream_t st) {
    using C = QuadCfg<DB, AG, TH, W, NBUF>;
    static_assert(DB >= 32 && AG >= 4 && TH >= 16, "quad_boxes_bytes assumes no smaller CTA shape");
    CUtensorMap tm, tmT;
    int rc = make_quad_map(&tm, q, batch, g.n + 1, W, C::ROWS);
    if (rc) return rc;
    rc = make_quad_map(&tmT, qt, batch, g.n + 1, W, C::ROWS);
    if (rc) return rc;
    const FaultCtl fc = fault_ctl();
    dim3 grid((unsigned)cdiv(g.det_count, DB), (unsigned)cdiv(g.n_angles, AG), (unsigned)batch);
    quad_boxes_kernel<DB, AG, TH, LD><<<dim3(grid.x, grid.y), C::THREADS, 0, st>>>(trig, g, boxes);
    PDU_LAUNCHED();
    if (fc.texq) {
        PDU_CUDA((ensure_dyn_smem<radon_fwd_quad_kernel<DB, AG, TH, W, NBUF, LD, true>>(C::SMEM)));
        radon_fwd_quad_kernel<DB, AG, TH, W, NBUF, LD, true><<<grid, C::THREADS + 32, C::SMEM, st>>>(tm, tmT, q, qt, sino, trig, boxes, g, fc);
    } else {
        PDU_CUDA((ensure_dyn_smem<radon_fwd_quad_kernel<DB, AG, TH, W, NBUF, LD, false>>(C::SMEM)));
        radon_fwd_quad_kernel<DB, AG, TH, W, NBUF, LD, false><<<grid, C::THREADS + 32, C::SMEM, st>>>(tm, tmT, q, qt, sino, trig, boxes, g, fc);
    }
    PDU_LAUNCHED();
    note_kernel(OP_RADON_FWD, "quad_build_kernel + quad_boxes_kernel + radon_fwd_quad_kernel<%d,%d,%d,%d,%d,%d> grid %ux%ux%u (bilinear-cell tiles, TMA ring)",
                DB, AG, TH, W, NBUF, LD, grid.x, grid.y, grid.z);
    return PDU_OK;
}

}  // namespace pdu

using namespace pdu;

static int check_geom(const pdu_radon_geom_t* g, int batch, const char* who) {
    PDU_REQUIRE(g != nullptr, "%s: geom is null", who);
    PDU_REQUIRE(g->geom == PDU_GEOM_PARALLEL || g->geom == PDU_GEOM_FAN, "%s: unknown geom %d", who, g->geom);
    PDU_REQUIRE(g->n > 0 && g->n_angles > 0 && g->det_count > 0 && batch > 0,
                "%s: sizes must be positive (n=%d angles=%d det=%d batch=%d)", who, g->n, g->n_angles, g->det_count,
                batch);
    PDU_REQUIRE(g->det_spacing > 0.f, "%s: det_spacing must be > 0", who);
    PDU_REQUIRE(batch <= 65535, "%s: batch %d exceeds 65535 (split the call)", who, batch);
    if (g->geom == PDU_GEOM_FAN)
        PDU_REQUIRE(g->s_dist > 0.f && g->d_dist >= 0.f, "%s: fan beam needs s_dist > 0, d_dist >= 0", who);
    return PDU_OK;
}

extern "C" {

int pdu_radon_trig_f32(const float* angles, float* trig, int n_angles, pdu_stream_t stream) {
    PDU_REQUIRE(angles && trig && n_angles > 0, "pdu_radon_trig_f32: null pointer or n_angles <= 0");
    trig_kernel<<<(unsigned)cdiv(n_angles, 128), 128, 0, (cudaStream_t)stream>>>(angles, trig, n_angles);
    PDU_LAUNCHED();
    return PDU_OK;
}

size_t pdu_radon_workspace_bytes(const pdu_radon_geom_t* g, int batch) {
    if (!g || g->n <= 0 || batch <= 0) return 0;
    // cell tensors of the slice and of its transpose (quad variants); the float-tile variants use the
    // first batch * n * n floats of the same scratch for the transposed copy
    return quad_bytes(g->n, batch) + quad_boxes_bytes(*g);
}

int pdu_radon_fwd_f32(const float* img, float* sino, const float* trig, int batch, const pdu_radon_geom_t* g,
                      void* workspace, size_t workspace_bytes, pdu_stream_t stream) {
    int rc = check_geom(g, batch, "pdu_radon_fwd_f32");
    if (rc) return rc;
    PDU_CHECK_DEVICE("pdu_radon_fwd_f32");
    PDU_REQUIRE(img && sino && trig, "pdu_radon_fwd_f32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int variant = option(OPT_RADON_FWD);
    if (variant < 0) variant = -1;
    if (variant == 0) {
        dim3 block(64, 4);
        dim3 grid((unsigned)cdiv(g->det_count, 64), (unsigned)cdiv(g->n_angles, 4), (unsigned)batch);
        radon_fwd_gather_kernel<<<grid, block, 0, st>>>(img, sino, trig, *g, option(OPT_TEX_WEIGHTS) > 0 ? 1 : 0);
        PDU_LAUNCHED();
        note_kernel(OP_RADON_FWD, "radon_fwd_gather_kernel grid %ux%ux%u (one thread per ray, L1 gather)", grid.x, grid.y, grid.z);
        return PDU_OK;
    }
    // columns one view step moves a ray across the slice (assumes evenly spread views; a wrong guess costs
    // speed, never correctness -- oversized strips take the global-load path)
    const float span = g->geom == PDU_GEOM_PARALLEL ? 3.14159265f : 6.2831853f;
    const float drift = (span / g->n_angles) * 0.7072f * g->n;
    const bool quad_ok = (g->n + 1 + 15) / 16 <= MAX_STRIPS && g->n_angles <= 65535 * 4;
    const bool tile_ok = (g->n % 4 == 0) && (((uintptr_t)img & 15) == 0) && (g->n + 1 + 31) / 32 <= MAX_STRIPS &&
                         g->n_angles <= 65535 * 2;
    // default: cell ("quad") strips for dense view sets (8 neighbouring views share an 88-cell box); sparser view
    // sets go to the float-tile kernel, whose wider boxes (136 .. 248 columns) keep several views per CTA --
    // measured on the sweep's sparse shapes (tools/prof_fwd_shapes.py): tile kernel 1.2 - 1.4x faster there,
    // cell kernel 1.03 - 1.15x faster at drift <= 3.3.  Shape 13 (a quarter-warp = 4 detectors x 2 neighbouring views:
    // the 8 lanes of one LDS.128 phase then span < 8 cell columns) is 2-3 % faster than shape 7 (8 detectors of one view)
    // r02 re-tuning after the strip-box table (tools/prof_fwd_shapes.py, batch 8): from 512^2 up the cell kernels also win
    // on sparser view sets -- drift 4.4: 462 vs 534 us (512^2 x 256 views), 3073 vs 3826 (1024^2 x 512); drift 8.9:
    // the 128-cell box 1799 vs 2014 us (1024^2 x 256); drift 17.8: four views per CTA 179 vs 218 us (512^2 x 64);
    // below 512^2 the float-tile kernel stays equal or better on those view sets
    if (variant < 0) {
        const bool large = g->n >= 512;
        if (quad_ok && (7.f * drift <= 27.f || (large && 7.f * drift <= 35.f))) variant = 13;
        else if (quad_ok && large) variant = 7.f * drift <= 84.f ? 11 : 9;
        else if (tile_ok) variant = 1;
        else if (quad_ok) variant = 7.f * drift <= 63.f ? 11 : 9;
        else variant = 0;
    }
    if (variant >= 2 && variant != 9 && variant != 11) variant = 13;      // the only explicit choices: 0, 1, 9, 11, 13
    if (variant >= 7 && !quad_ok) variant = tile_ok ? 1 : 0;
    if (variant == 1 && !tile_ok) variant = quad_ok ? 9 : 0;
    if (variant == 0) {
        dim3 block(64, 4);
        dim3 grid((unsigned)cdiv(g->det_count, 64), (unsigned)cdiv(g->n_angles, 4), (unsigned)batch);
        radon_fwd_gather_kernel<<<grid, block, 0, st>>>(img, sino, trig, *g, option(OPT_TEX_WEIGHTS) > 0 ? 1 : 0);
        PDU_LAUNCHED();
        note_kernel(OP_RADON_FWD, "radon_fwd_gather_kernel grid %ux%ux%u (one thread per ray, L1 gather)", grid.x, grid.y, grid.z);
        return PDU_OK;
    }
    const size_t need = variant >= 7 ? quad_bytes(g->n, batch) + quad_boxes_bytes(*g) : (size_t)batch * g->n * g->n * sizeof(float);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15)) {
        set_error("pdu_radon_fwd_f32: workspace of %zu bytes (16-byte aligned) required, got %zu", need,
                  workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    if (variant >= 7) {
        const int n1 = g->n + 1;
        float4* q = (float4*)workspace;
        float4* qt = q + (size_t)batch * n1 * n1;
        int2* boxes = (int2*)(qt + (size_t)batch * n1 * n1);
        {
            dim3 block(32, 8);
            dim3 grid((unsigned)cdiv(n1, 32), (unsigned)cdiv(n1, 32), (unsigned)batch);
            quad_build_kernel<<<grid, block, 0, st>>>(img, q, qt, g->n);
            PDU_LAUNCHED();
        }
        // shapes the dispatcher uses (every other shape measured in r01 -- 32-row strips, 512-thread CTAs, 3-deep rings,
        // 84/76/88/89-cell pitches, 8- and 16-detector quarter-warps -- lost by 2-20 % and was deleted; r02 with the
        // slack row, configs[1]: 12 / 16 / 20 / 24-row strips 474 / 451 / 458 / 445 us, 3-deep ring 589 us, coordinates of
        // four samples from per-strip origins (one FMA2 per sample, no per-sample counter): 454 us at 56 registers --
        // the loop is paced by the LDS.128 wavefronts, not by those two instructions)
        switch (variant) {
            case 9: return launch_quad<32, 4, 16, 120, 2, 8>(q, qt, sino, trig, boxes, batch, *g, st);    // 4 views / CTA (sparser views)
            case 11: return launch_quad<32, 8, 16, 128, 2, 8>(q, qt, sino, trig, boxes, batch, *g, st);   // widest cell box (sparser views)
            default: return launch_quad<32, 8, 16, 92, 2, 4>(q, qt, sino, trig, boxes, batch, *g, st);    // 13: quarter-warp = 4 detectors x 2 views
        }
    }
    float* imgT = (float*)workspace;
    {
        dim3 block(32, 8);
        dim3 grid((unsigned)cdiv(g->n, 32), (unsigned)cdiv(g->n, 32), (unsigned)batch);
        transpose_kernel<<<grid, block, 0, st>>>(img, imgT, g->n);
        PDU_LAUNCHED();
    }
    // variant 1: the r01 default -- as many neighbouring views per CTA as keep the strip box inside W
    if (7.f * drift <= 40.f) return launch_strip<32, 8, 32, 136, 2, 8>(img, imgT, sino, trig, batch, *g, st);
    if (3.f * drift <= 28.f) return launch_strip<64, 4, 32, 168, 2, 16>(img, imgT, sino, trig, batch, *g, st);
    return launch_strip<128, 2, 32, 248, 3, 32>(img, imgT, sino, trig, batch, *g, st);
}

}  // extern "C"
