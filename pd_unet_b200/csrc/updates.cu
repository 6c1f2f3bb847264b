// The elementwise steps of one unrolled primal-dual iteration that are not convolutions:
//   cat(...) feeding each block, state + delta with the channel slice the next operator consumes,
//   a x + b y, and PD-UNet's angular sinogram upsampling with its exact transpose.
// All are HBM streaming kernels: 128-bit accesses when the plane size and pointers allow, grid sized
// in multiples of the SM count, one pass over the data (the slice is written in the same pass that
// writes the sum, so the operator input is not re-read from the state).
#include "common.cuh"

namespace pdu {

static inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static inline unsigned stream_grid(long work_items, int threads) {
    const long want = cdiv(work_items, threads);
    const long cap = (long)sm_count() * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

// ------------------------------------------------------------------ concat
template <typename V>
__device__ __forceinline__ V vzero();
template <>
__device__ __forceinline__ float vzero<float>() { return 0.f; }
template <>
__device__ __forceinline__ float4 vzero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <typename V>
__device__ __forceinline__ V vscale(V a, float s);
template <>
__device__ __forceinline__ float vscale<float>(float a, float s) { return a * s; }
template <>
__device__ __forceinline__ float4 vscale<float4>(float4 a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}

// planar (NCHW): per batch element the output row is [a-row (la) | b-row (lb) | c-row (lc) | zeros (lz)] in units of V
template <typename V>
__global__ void __launch_bounds__(256)
    concat_kernel(V* __restrict__ out, const V* __restrict__ a, const V* __restrict__ b, const V* __restrict__ c,
                  long la, long lb, long lc, long lz, float scale_b, long total) {
    const long lo = la + lb + lc + lz;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long n = i / lo, r = i - n * lo;
        V v;
        if (r < la) v = a[n * la + r];
        else if (r < la + lb) v = vscale<V>(b[n * lb + (r - la)], scale_b);
        else if (r < la + lb + lc) v = c[n * lc + (r - la - lb)];
        else v = vzero<V>();
        out[i] = v;
    }
}

// channels-last (NHWC): every pixel's output vector is [a-channels | b-channels | c-channels | zeros]
__global__ void __launch_bounds__(256)
    concat_nhwc_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                       const float* __restrict__ c, int ca, int cb, int cc, int ct, float scale_b, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long pix = i / ct;
        const int ch = (int)(i - pix * ct);
        float v;
        if (ch < ca) v = a[pix * ca + ch];
        else if (ch < ca + cb) v = b[pix * cb + (ch - ca)] * scale_b;
        else if (ch < ca + cb + cc) v = c[pix * cc + (ch - ca - cb)];
        else v = 0.f;
        out[i] = v;
    }
}

// channels-last output and state, but b and / or c still PLANAR ([batch, channels, plane]: what the operators write):
// the layout change happens in the concatenation instead of in a separate copy pass per operand.  A warp covers a few
// pixels x all channels; the planar reads of one channel are 4-byte pieces of a line that the following warps reuse
// from L1.
__global__ void __launch_bounds__(256)
    concat_nhwc_mixed_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                             const float* __restrict__ c, int ca, int cb, int cc, int ct, long plane, float scale_b,
                             int b_planar, int c_planar, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long pix = i / ct;
        const int ch = (int)(i - pix * ct);
        const long n = pix / plane, p = pix - n * plane;
        float v;
        if (ch < ca) v = a[pix * ca + ch];
        else if (ch < ca + cb) v = (b_planar ? __ldg(b + (n * cb + (ch - ca)) * plane + p) : b[pix * cb + (ch - ca)]) * scale_b;
        else if (ch < ca + cb + cc) v = c_planar ? __ldg(c + (n * cc + (ch - ca - cb)) * plane + p) : c[pix * cc + (ch - ca - cb)];
        else v = 0.f;
        out[i] = v;
    }
}

// ------------------------------------------------------------------ residual + slice
template <typename V>
__device__ __forceinline__ V vadd(V a, V b);
template <>
__device__ __forceinline__ float vadd<float>(float a, float b) { return a + b; }
template <>
__device__ __forceinline__ float4 vadd<float4>(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

template <typename V>
__global__ void __launch_bounds__(256)
    residual_slice_kernel(V* out, V* __restrict__ slice, const V* state, const V* delta,
                          long plane, int channels, int k, int kn, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const V v = vadd<V>(state[i], delta[i]);
        out[i] = v;
        if (slice) {
            const long nc = i / plane, p = i - nc * plane;
            const long n = nc / channels;
            const int ch = (int)(nc - n * channels) - k;
            if (ch >= 0 && ch < kn) slice[(n * kn + ch) * plane + p] = v;
        }
    }
}

// channels-last state [batch, plane, channels]; the slice stays planar [batch, kn, plane] (operator input)
__global__ void __launch_bounds__(256)
    residual_slice_nhwc_kernel(float* out, float* __restrict__ slice, const float* state, const float* delta, long plane,
                               int channels, int k, int kn, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const float v = state[i] + delta[i];
        out[i] = v;
        if (slice) {
            const long pix = i / channels;
            const int ch = (int)(i - pix * channels) - k;
            if (ch >= 0 && ch < kn) {
                const long n = pix / plane, p = pix - n * plane;
                slice[(n * kn + ch) * plane + p] = v;
            }
        }
    }
}

// ------------------------------------------------------------------ axpby
template <typename V>
__device__ __forceinline__ V vaxpby(float a, V x, float b, V y);
template <>
__device__ __forceinline__ float vaxpby<float>(float a, float x, float b, float y) { return fmaf(a, x, b * y); }
template <>
__device__ __forceinline__ float4 vaxpby<float4>(float a, float4 x, float b, float4 y) {
    return make_float4(fmaf(a, x.x, b * y.x), fmaf(a, x.y, b * y.y), fmaf(a, x.z, b * y.z), fmaf(a, x.w, b * y.w));
}

template <typename V>
__global__ void __launch_bounds__(256)
    axpby_kernel(V* out, float alpha, const V* x, float beta, const V* y, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        out[i] = vaxpby<V>(alpha, x[i], beta, y[i]);
}

// ------------------------------------------------------------------ angular upsampling
// full[b, i*f + r, d] = (1 - r/f) s[b, i, d] + (r/f) s[b, i+1, d];  the row after the last one is
// s[b, 0, D-1-d] (mode 0), s[b, 0, d] (mode 1) or s[b, As-1, d] (mode 2).
// value of full-view pixel `pix` (= (b * As * f + view) * D + d) of the upsampled sinogram
__device__ __forceinline__ float upsampled_at(const float* __restrict__ sparse, long pix, int As, int f, int D, int mode, float inv_f) {
    const int d = (int)(pix % D);
    const long row = pix / D;
    const int av = (int)(row % ((long)As * f));
    const long b = row / ((long)As * f);
    const int ia = av / f, r = av - ia * f;
    const float* sb = sparse + b * (long)As * D;
    const float lo = __ldg(sb + (long)ia * D + d);
    float hi;
    if (ia + 1 < As) hi = __ldg(sb + (long)(ia + 1) * D + d);
    else if (mode == PDU_WRAP_FLIP) hi = __ldg(sb + (D - 1 - d));
    else if (mode == PDU_WRAP_PERIODIC) hi = __ldg(sb + d);
    else hi = lo;
    return fmaf((float)r * inv_f, hi - lo, lo);
}

__global__ void __launch_bounds__(256)
    upsample_kernel(const float* __restrict__ sparse, float* __restrict__ full, int As, int f, int D, int mode, float scale,
                    long total) {
    const float inv_f = 1.f / (float)f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        full[i] = scale * upsampled_at(sparse, i, As, f, D, mode, inv_f);
}

// The dual update's concatenation with the measured data interpolated on the fly: out = [a | scale_b b | scale_c up(sparse) | 0].
// The full-view sinogram g is never materialised -- the sparse views (1 / factor of it) are read instead, from L2.
// nhwc: out is [batch, plane, ct], a [batch, plane, ca], b [batch, plane]; else planar.
__global__ void __launch_bounds__(256)
    concat_upsample_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                           const float* __restrict__ sparse, int ca, int ct, long plane, int As, int f, int D, int mode,
                           float scale_b, float scale_c, int nhwc, long total) {
    const float inv_f = 1.f / (float)f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long pix;
        int ch;
        if (nhwc) {
            pix = i / ct;
            ch = (int)(i - pix * ct);
        } else {
            const long n = i / ((long)ct * plane), r = i - n * (long)ct * plane;
            ch = (int)(r / plane);
            pix = n * plane + (r - (long)ch * plane);
        }
        float v = 0.f;
        if (ch < ca) v = nhwc ? a[pix * ca + ch] : a[(pix / plane * ca + ch) * plane + pix % plane];
        else if (ch == ca) v = b[pix] * scale_b;
        else if (ch == ca + 1) v = scale_c * upsampled_at(sparse, pix, As, f, D, mode, inv_f);
        out[i] = v;
    }
}

// channels-last fast path (4 state channels, output padded to 8): one thread per pixel
__global__ void __launch_bounds__(256)
    concat_upsample_nhwc_4_to8_kernel(float4* __restrict__ out, const float4* __restrict__ a, const float* __restrict__ b,
                                      const float* __restrict__ sparse, int As, int f, int D, int mode, float scale_b,
                                      float scale_c, long pixels) {
    const float inv_f = 1.f / (float)f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long)gridDim.x * blockDim.x) {
        out[2 * i] = a[i];
        out[2 * i + 1] = make_float4(b[i] * scale_b, scale_c * upsampled_at(sparse, i, As, f, D, mode, inv_f), 0.f, 0.f);
    }
}

// Transpose: sparse[b, i, d] = sum_r (1 - r/f) full[b, i f + r, d] + sum_r (r/f) full[b, (i-1) f + r, d]
// (+ the wrapped contribution of the last group).  Gather form, deterministic, no atomics.
__global__ void __launch_bounds__(256)
    upsample_adj_kernel(const float* __restrict__ full, float* __restrict__ sparse, int As, int f, int D, int mode,
                        long total) {
    const float inv_f = 1.f / (float)f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long row = i / D;
        const int ia = (int)(row % As);
        const long b = row / As;
        const float* fb = full + b * (long)As * f * D;
        float acc = 0.f;
        for (int r = 0; r < f; ++r) acc = fmaf(1.f - (float)r * inv_f, __ldg(fb + ((long)ia * f + r) * D + d), acc);
        if (ia > 0)
            for (int r = 1; r < f; ++r) acc = fmaf((float)r * inv_f, __ldg(fb + ((long)(ia - 1) * f + r) * D + d), acc);
        // the last group's upper neighbour
        const long last = (long)(As - 1) * f;
        if (mode == PDU_WRAP_FLIP && ia == 0) {
            for (int r = 1; r < f; ++r) acc = fmaf((float)r * inv_f, __ldg(fb + (last + r) * D + (D - 1 - d)), acc);
        } else if (mode == PDU_WRAP_PERIODIC && ia == 0) {
            for (int r = 1; r < f; ++r) acc = fmaf((float)r * inv_f, __ldg(fb + (last + r) * D + d), acc);
        } else if (mode == PDU_WRAP_CLAMP && ia == As - 1) {
            for (int r = 1; r < f; ++r) acc = fmaf((float)r * inv_f, __ldg(fb + (last + r) * D + d), acc);
        }
        sparse[i] = acc;
    }
}

}  // namespace pdu

using namespace pdu;

// channels-last fast path for the shapes the unrolled iteration produces: a has 4 channels, b (and c) one,
// output padded to 8 -- one thread per pixel, one 16-byte load and two 16-byte stores
namespace pdu {
__global__ void __launch_bounds__(256)
    concat_nhwc_4_1_1_to8_kernel(float4* __restrict__ out, const float4* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ c, float scale_b, long pixels) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long)gridDim.x * blockDim.x) {
        out[2 * i] = a[i];
        out[2 * i + 1] = make_float4(b[i] * scale_b, c ? c[i] : 0.f, 0.f, 0.f);
    }
}
}  // namespace pdu

extern "C" {

int pdu_concat_f32(float* out, const float* a, const float* b, const float* c, int batch, int ca, int cb, int cc,
                   int c_out, long plane, float scale_b, int layout, pdu_stream_t stream) {
    PDU_REQUIRE(out && a && b, "pdu_concat_f32: null pointer");
    PDU_REQUIRE(batch > 0 && ca > 0 && cb > 0 && cc >= 0 && plane > 0, "pdu_concat_f32: sizes must be positive");
    PDU_REQUIRE((c != nullptr) == (cc > 0), "pdu_concat_f32: c must be given exactly when cc > 0");
    PDU_REQUIRE(c_out >= ca + cb + cc, "pdu_concat_f32: c_out %d is smaller than the %d input channels", c_out, ca + cb + cc);
    const long la = ca * plane, lb = cb * plane, lc = cc * plane, lz = (long)(c_out - ca - cb - cc) * plane;
    const long total = (long)batch * (la + lb + lc + lz);
    cudaStream_t st = (cudaStream_t)stream;
    PDU_REQUIRE(layout == PDU_LAYOUT_NCHW || layout == PDU_LAYOUT_NHWC, "pdu_concat_f32: unknown layout %d", layout);
    if (layout == PDU_LAYOUT_NHWC) {
        if (ca == 4 && cb == 1 && cc <= 1 && c_out == 8 && al16(out) && al16(a)) {
            const long pixels = (long)batch * plane;
            concat_nhwc_4_1_1_to8_kernel<<<stream_grid(pixels, 256), 256, 0, st>>>((float4*)out, (const float4*)a, b, c, scale_b,
                                                                                   pixels);
            PDU_LAUNCHED();
            return PDU_OK;
        }
        concat_nhwc_kernel<<<stream_grid(total, 256), 256, 0, st>>>(out, a, b, c, ca, cb, cc, c_out, scale_b, total);
        PDU_LAUNCHED();
        return PDU_OK;
    }
    const bool vec = plane % 4 == 0 && al16(out) && al16(a) && al16(b) && (c == nullptr || al16(c));
    if (vec) {
        concat_kernel<float4><<<stream_grid(total / 4, 256), 256, 0, st>>>((float4*)out, (const float4*)a, (const float4*)b,
                                                                           (const float4*)c, la / 4, lb / 4, lc / 4, lz / 4, scale_b, total / 4);
    } else {
        concat_kernel<float><<<stream_grid(total, 256), 256, 0, st>>>(out, a, b, c, la, lb, lc, lz, scale_b, total);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_concat_mixed_f32(float* out, const float* a, const float* b, const float* c, int batch, int ca, int cb, int cc,
                         int c_out, long plane, float scale_b, int b_planar, int c_planar, pdu_stream_t stream) {
    PDU_REQUIRE(out && a && b, "pdu_concat_mixed_f32: null pointer");
    PDU_REQUIRE(batch > 0 && ca > 0 && cb > 0 && cc >= 0 && plane > 0, "pdu_concat_mixed_f32: sizes must be positive");
    PDU_REQUIRE((c != nullptr) == (cc > 0), "pdu_concat_mixed_f32: c must be given exactly when cc > 0");
    PDU_REQUIRE(c_out >= ca + cb + cc, "pdu_concat_mixed_f32: c_out %d is smaller than the %d input channels", c_out, ca + cb + cc);
    const long total = (long)batch * plane * c_out;
    concat_nhwc_mixed_kernel<<<stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(out, a, b, c, ca, cb, cc, c_out, plane, scale_b,
                                                                                        b_planar ? 1 : 0, c_planar ? 1 : 0, total);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_residual_slice_f32(float* out, float* slice, const float* state, const float* delta, int batch, int channels,
                           long plane, int k, int kn, int layout, pdu_stream_t stream) {
    PDU_REQUIRE(out && state && delta, "pdu_residual_slice_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && plane > 0, "pdu_residual_slice_f32: sizes must be positive");
    PDU_REQUIRE(slice == nullptr || (k >= 0 && kn >= 1 && k + kn <= channels),
                "pdu_residual_slice_f32: slice channels [%d, %d) out of range", k, k + kn);
    const long total = (long)batch * channels * plane;
    cudaStream_t st = (cudaStream_t)stream;
    PDU_REQUIRE(layout == PDU_LAYOUT_NCHW || layout == PDU_LAYOUT_NHWC, "pdu_residual_slice_f32: unknown layout %d", layout);
    if (layout == PDU_LAYOUT_NHWC) {
        residual_slice_nhwc_kernel<<<stream_grid(total, 256), 256, 0, st>>>(out, slice, state, delta, plane, channels, k, kn,
                                                                            total);
        PDU_LAUNCHED();
        return PDU_OK;
    }
    const bool vec = plane % 4 == 0 && al16(out) && al16(state) && al16(delta) && (slice == nullptr || al16(slice));
    if (vec) {
        residual_slice_kernel<float4><<<stream_grid(total / 4, 256), 256, 0, st>>>(
            (float4*)out, (float4*)slice, (const float4*)state, (const float4*)delta, plane / 4, channels, k, kn, total / 4);
    } else {
        residual_slice_kernel<float><<<stream_grid(total, 256), 256, 0, st>>>(out, slice, state, delta, plane, channels, k,
                                                                              kn, total);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_axpby_f32(float* out, float alpha, const float* x, float beta, const float* y, long n, pdu_stream_t stream) {
    PDU_REQUIRE(out && x && y && n > 0, "pdu_axpby_f32: null pointer or n <= 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (n % 4 == 0 && al16(out) && al16(x) && al16(y)) {
        axpby_kernel<float4><<<stream_grid(n / 4, 256), 256, 0, st>>>((float4*)out, alpha, (const float4*)x, beta,
                                                                      (const float4*)y, n / 4);
    } else {
        axpby_kernel<float><<<stream_grid(n, 256), 256, 0, st>>>(out, alpha, x, beta, y, n);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_angular_upsample_scaled_f32(const float* sparse, float* full, int batch, int a_sparse, int factor, int det_count,
                                    int mode, float scale, pdu_stream_t stream) {
    PDU_REQUIRE(sparse && full, "pdu_angular_upsample_f32: null pointer");
    PDU_REQUIRE(batch > 0 && a_sparse > 0 && factor > 0 && det_count > 0, "pdu_angular_upsample_f32: sizes must be positive");
    PDU_REQUIRE(mode >= 0 && mode <= 2, "pdu_angular_upsample_f32: unknown mode %d", mode);
    const long total = (long)batch * a_sparse * factor * det_count;
    upsample_kernel<<<stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(sparse, full, a_sparse, factor, det_count,
                                                                               mode, scale, total);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_angular_upsample_f32(const float* sparse, float* full, int batch, int a_sparse, int factor, int det_count,
                             int mode, pdu_stream_t stream) {
    return pdu_angular_upsample_scaled_f32(sparse, full, batch, a_sparse, factor, det_count, mode, 1.f, stream);
}

int pdu_concat_upsample_f32(float* out, const float* a, const float* b, const float* sparse, int batch, int ca, int c_out,
                            int a_sparse, int factor, int det_count, int mode, float scale_b, float scale_c, int layout,
                            pdu_stream_t stream) {
    PDU_REQUIRE(out && a && b && sparse, "pdu_concat_upsample_f32: null pointer");
    PDU_REQUIRE(batch > 0 && ca > 0 && a_sparse > 0 && factor > 0 && det_count > 0, "pdu_concat_upsample_f32: sizes must be positive");
    PDU_REQUIRE(c_out >= ca + 2, "pdu_concat_upsample_f32: c_out %d is smaller than the %d input channels", c_out, ca + 2);
    PDU_REQUIRE(mode >= 0 && mode <= 2, "pdu_concat_upsample_f32: unknown mode %d", mode);
    PDU_REQUIRE(layout == PDU_LAYOUT_NCHW || layout == PDU_LAYOUT_NHWC, "pdu_concat_upsample_f32: unknown layout %d", layout);
    cudaStream_t st = (cudaStream_t)stream;
    const long plane = (long)a_sparse * factor * det_count;
    const long pixels = (long)batch * plane;
    if (layout == PDU_LAYOUT_NHWC && ca == 4 && c_out == 8 && al16(out) && al16(a)) {
        concat_upsample_nhwc_4_to8_kernel<<<stream_grid(pixels, 256), 256, 0, st>>>((float4*)out, (const float4*)a, b, sparse, a_sparse,
                                                                                    factor, det_count, mode, scale_b, scale_c, pixels);
        PDU_LAUNCHED();
        return PDU_OK;
    }
    const long total = pixels * c_out;
    concat_upsample_kernel<<<stream_grid(total, 256), 256, 0, st>>>(out, a, b, sparse, ca, c_out, plane, a_sparse, factor, det_count, mode,
                                                                    scale_b, scale_c, layout == PDU_LAYOUT_NHWC ? 1 : 0, total);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_angular_upsample_adj_f32(const float* full, float* sparse, int batch, int a_sparse, int factor, int det_count,
                                 int mode, pdu_stream_t stream) {
    PDU_REQUIRE(sparse && full, "pdu_angular_upsample_adj_f32: null pointer");
    PDU_REQUIRE(batch > 0 && a_sparse > 0 && factor > 0 && det_count > 0,
                "pdu_angular_upsample_adj_f32: sizes must be positive");
    PDU_REQUIRE(mode >= 0 && mode <= 2, "pdu_angular_upsample_adj_f32: unknown mode %d", mode);
    const long total = (long)batch * a_sparse * det_count;
    upsample_adj_kernel<<<stream_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(full, sparse, a_sparse, factor,
                                                                                   det_count, mode, total);
    PDU_LAUNCHED();
    return PDU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ bias + PReLU epilogue
// y <- prelu(y + bias[c], slope[c]) in place, one pass (ATen runs the bias add and the activation of a
// convolution as two full passes over the activation map).  slope may be NULL (bias only) and may hold
// one value for all channels (n_slope == 1).
namespace pdu {

__device__ __forceinline__ float prelu1(float v, float s) { return v >= 0.f ? v : v * s; }

template <bool HAS_SLOPE>
__global__ void __launch_bounds__(256)
    bias_prelu_nhwc4_kernel(float4* y, const float4* __restrict__ bias, const float4* __restrict__ slope, int c4,
                            int n_slope, long total4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        const float4 b = __ldg(bias + c);
        float4 v = y[i];
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (HAS_SLOPE) {
            float4 s;
            if (n_slope == 1) { const float s1 = __ldg((const float*)slope); s = make_float4(s1, s1, s1, s1); }
            else s = __ldg(slope + c);
            v.x = prelu1(v.x, s.x); v.y = prelu1(v.y, s.y); v.z = prelu1(v.z, s.z); v.w = prelu1(v.w, s.w);
        }
        y[i] = v;
    }
}

// general: channel = (i / inner) % channels with inner = 1 (NHWC) or plane (NCHW)
template <bool HAS_SLOPE>
__global__ void __launch_bounds__(256)
    bias_prelu_kernel(float* y, const float* __restrict__ bias, const float* __restrict__ slope, int channels, long inner,
                      int n_slope, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)((i / inner) % channels);
        float v = y[i] + __ldg(bias + c);
        if (HAS_SLOPE) v = prelu1(v, __ldg(slope + (n_slope == 1 ? 0 : c)));
        y[i] = v;
    }
}

}  // namespace pdu

extern "C" int pdu_bias_prelu_f32(float* y, const float* bias, const float* slope, int n_slope, int batch, int channels,
                                  long plane, int layout, pdu_stream_t stream) {
    PDU_REQUIRE(y && bias, "pdu_bias_prelu_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && plane > 0, "pdu_bias_prelu_f32: sizes must be positive");
    PDU_REQUIRE(slope == nullptr || n_slope == 1 || n_slope == channels, "pdu_bias_prelu_f32: slope must have 1 or %d values",
                channels);
    PDU_REQUIRE(layout == PDU_LAYOUT_NCHW || layout == PDU_LAYOUT_NHWC, "pdu_bias_prelu_f32: unknown layout %d", layout);
    const long total = (long)batch * channels * plane;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = layout == PDU_LAYOUT_NHWC && channels % 4 == 0 && al16(y) && al16(bias) &&
                     (slope == nullptr || n_slope == 1 || al16(slope));
    if (vec) {
        if (slope) bias_prelu_nhwc4_kernel<true><<<stream_grid(total / 4, 256), 256, 0, st>>>(
                (float4*)y, (const float4*)bias, (const float4*)slope, channels / 4, n_slope, total / 4);
        else bias_prelu_nhwc4_kernel<false><<<stream_grid(total / 4, 256), 256, 0, st>>>(
                (float4*)y, (const float4*)bias, nullptr, channels / 4, 1, total / 4);
    } else {
        const long inner = layout == PDU_LAYOUT_NHWC ? 1 : plane;
        if (slope) bias_prelu_kernel<true><<<stream_grid(total, 256), 256, 0, st>>>(y, bias, slope, channels, inner, n_slope, total);
        else bias_prelu_kernel<false><<<stream_grid(total, 256), 256, 0, st>>>(y, bias, nullptr, channels, inner, 1, total);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}

// ------------------------------------------------------------------ bias + PReLU for training: out of place + backward
// Forward keeps the convolution output y (the backward needs the sign and value of z = y + bias) and writes
// out = prelu(z).  Backward, ONE pass over (g, y):  gz = g * (z > 0 ? 1 : slope)  and per-channel
//   gbias[c] = sum gz,   gslope[c] = sum g * min(z, 0)
// accumulated in registers (a thread always sees the same four channels: the grid stride is a multiple of C / 4),
// reduced per block in shared memory, and summed over the blocks by a second tiny kernel in a fixed order
// (reproducible: no atomics).  ATen runs bias add, PReLU, PReLU backward (which writes a full-size slope-gradient
// tensor) and two full-size sum reductions: 41 % of a PD-UNet training step (profiles/r01_training.md).
namespace pdu {

template <bool HAS_SLOPE>
__global__ void __launch_bounds__(256)
    bias_prelu_out_nhwc4_kernel(const float4* __restrict__ y, float4* __restrict__ out, const float4* __restrict__ bias,
                                const float4* __restrict__ slope, int c4, int n_slope, long total4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        const float4 b = __ldg(bias + c);
        float4 v = __ldg(y + i);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (HAS_SLOPE) {
            float4 s;
            if (n_slope == 1) { const float s1 = __ldg((const float*)slope); s = make_float4(s1, s1, s1, s1); }
            else s = __ldg(slope + c);
            v.x = prelu1(v.x, s.x); v.y = prelu1(v.y, s.y); v.z = prelu1(v.z, s.z); v.w = prelu1(v.w, s.w);
        }
        out[i] = v;
    }
}

// d(prelu)/dz with ATen's convention at z == 0 (the negative branch)
__device__ __forceinline__ void prelu_bwd1(float g, float z, float s, float& gz, float& sb, float& sa) {
    const bool pos = z > 0.f;
    gz = pos ? g : g * s;
    sb += gz;
    sa += pos ? 0.f : g * z;
}

__global__ void __launch_bounds__(256)
    bias_prelu_bwd_nhwc4_kernel(const float4* __restrict__ g, const float4* __restrict__ y, const float4* __restrict__ bias,
                                const float4* __restrict__ slope, int n_slope, float4* __restrict__ gz,
                                float* __restrict__ partial, int c4, long total4) {
    // gridDim.x * 256 is a multiple of c4 (host checks 256 % c4 == 0): the channel group of a thread never changes
    const int c = threadIdx.x % c4;
    const float4 b = __ldg(bias + c);
    float4 s;
    if (n_slope == 1) { const float s1 = __ldg((const float*)slope); s = make_float4(s1, s1, s1, s1); }
    else s = __ldg(slope + c);
    float sb[4] = {0.f, 0.f, 0.f, 0.f}, sa[4] = {0.f, 0.f, 0.f, 0.f};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const float4 gi = __ldg(g + i), yi = __ldg(y + i);
        float4 o;
        prelu_bwd1(gi.x, yi.x + b.x, s.x, o.x, sb[0], sa[0]);
        prelu_bwd1(gi.y, yi.y + b.y, s.y, o.y, sb[1], sa[1]);
        prelu_bwd1(gi.z, yi.z + b.z, s.z, o.z, sb[2], sa[2]);
        prelu_bwd1(gi.w, yi.w + b.w, s.w, o.w, sb[3], sa[3]);
        gz[i] = o;
    }
    __shared__ float red[8][256];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        red[k][threadIdx.x] = sb[k];
        red[4 + k][threadIdx.x] = sa[k];
    }
    __syncthreads();
    // thread t < 8 c4 sums quantity q = t / c4 of channel group cc = t % c4 over the 256 / c4 threads that own it
    for (int idx = threadIdx.x; idx < 8 * c4; idx += 256) {
        const int q = idx / c4, cc = idx - q * c4;
        float acc = 0.f;
        for (int t = cc; t < 256; t += c4) acc += red[q][t];
        // partial [block][2][C]: q < 4 -> gbias of channel 4 cc + q, q >= 4 -> gslope of channel 4 cc + q - 4
        const int C = 4 * c4;
        partial[((long)blockIdx.x * 2 + (q >> 2)) * C + 4 * cc + (q & 3)] = acc;
    }
}

// second stage: block idx = (quantity, channel); its 256 threads add that entry of every first-stage block and a
// fixed-shape tree adds the threads (the first version -- one block, each thread walking all 592 partial rows -- took
// 54 us per layer, latency bound: 7.6 % of a training step)
__global__ void __launch_bounds__(256)
    bias_prelu_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int C, int n_slope, float* __restrict__ gbias,
                                 float* __restrict__ gslope, float* __restrict__ slope_tmp) {
    __shared__ float red[256];
    const int idx = blockIdx.x, t = threadIdx.x;
    float acc = 0.f;
    for (int b = t; b < nblocks; b += 256) acc += partial[(long)b * 2 * C + idx];
    red[t] = acc;
    __syncthreads();
#pragma unroll
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) red[t] += red[t + w];
        __syncthreads();
    }
    if (t == 0) {
        if (idx < C) gbias[idx] = red[0];
        else if (n_slope == C) gslope[idx - C] = red[0];
        else slope_tmp[idx - C] = red[0];
    }
}

// one shared slope: add the per-channel sums in a fixed order
__global__ void bias_prelu_bwd_scalar_slope_kernel(const float* __restrict__ slope_tmp, int C, float* __restrict__ gslope) {
    float v = 0.f;
    for (int c = 0; c < C; ++c) v += slope_tmp[c];
    gslope[0] = v;
}

// per-channel sum of a channels-last tensor (the bias gradient of a convolution without activation): same two stages
__global__ void __launch_bounds__(256)
    channel_sum_nhwc4_kernel(const float4* __restrict__ g, float* __restrict__ partial, int c4, long total4) {
    float sb[4] = {0.f, 0.f, 0.f, 0.f};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const float4 gi = __ldg(g + i);
        sb[0] += gi.x; sb[1] += gi.y; sb[2] += gi.z; sb[3] += gi.w;
    }
    __shared__ float red[4][256];
#pragma unroll
    for (int k = 0; k < 4; ++k) red[k][threadIdx.x] = sb[k];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * c4; idx += 256) {
        const int q = idx / c4, cc = idx - q * c4;
        float acc = 0.f;
        for (int t = cc; t < 256; t += c4) acc += red[q][t];
        partial[(long)blockIdx.x * 4 * c4 + 4 * cc + q] = acc;
    }
}

__global__ void __launch_bounds__(256)
    partial_rows_sum_kernel(const float* __restrict__ partial, int nblocks, int row_len, float* __restrict__ out) {
    __shared__ float red[256];
    const int idx = blockIdx.x, t = threadIdx.x;
    float acc = 0.f;
    for (int b = t; b < nblocks; b += 256) acc += partial[(long)b * row_len + idx];
    red[t] = acc;
    __syncthreads();
#pragma unroll
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w) red[t] += red[t + w];
        __syncthreads();
    }
    if (t == 0) out[idx] = red[0];
}

}  // namespace pdu

static bool bias_prelu_train_ok(int channels, int layout, const void* a, const void* b, const void* c, const void* d) {
    const int c4 = channels / 4;
    return layout == PDU_LAYOUT_NHWC && channels % 4 == 0 && c4 >= 1 && c4 <= 64 && 256 % c4 == 0 && pdu::al16(a) && pdu::al16(b) &&
           pdu::al16(c) && (d == nullptr || pdu::al16(d));
}

extern "C" int pdu_bias_prelu_fwd_f32(const float* y, float* out, const float* bias, const float* slope, int n_slope, int batch,
                                      int channels, long plane, int layout, pdu_stream_t stream) {
    using namespace pdu;
    PDU_REQUIRE(y && out && bias && slope, "pdu_bias_prelu_fwd_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && plane > 0, "pdu_bias_prelu_fwd_f32: sizes must be positive");
    PDU_REQUIRE(n_slope == 1 || n_slope == channels, "pdu_bias_prelu_fwd_f32: slope must have 1 or %d values", channels);
    if (!bias_prelu_train_ok(channels, layout, y, out, bias, n_slope == 1 ? nullptr : slope)) {
        set_error("pdu_bias_prelu_fwd_f32: needs a 16-byte aligned channels-last tensor with channels in {4,8,16,32,64,128,256}");
        return PDU_EUNSUPPORTED;
    }
    const long total4 = (long)batch * channels * plane / 4;
    bias_prelu_out_nhwc4_kernel<true><<<stream_grid(total4, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)y, (float4*)out, (const float4*)bias, (const float4*)slope, channels / 4, n_slope, total4);
    PDU_LAUNCHED();
    return PDU_OK;
}

extern "C" int pdu_channel_sum_f32(const float* g, float* out, void* workspace, size_t workspace_bytes, int batch, int channels,
                                   long plane, int layout, pdu_stream_t stream) {
    using namespace pdu;
    PDU_REQUIRE(g && out, "pdu_channel_sum_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && plane > 0, "pdu_channel_sum_f32: sizes must be positive");
    const int c4 = channels / 4;
    if (!(layout == PDU_LAYOUT_NHWC && channels % 4 == 0 && c4 <= 64 && 256 % c4 == 0 && al16(g))) {
        set_error("pdu_channel_sum_f32: needs a 16-byte aligned channels-last tensor with channels in {4,8,16,32,64,128,256}");
        return PDU_EUNSUPPORTED;
    }
    const size_t need = (size_t)4 * sm_count() * channels * sizeof(float);
    if (!workspace || workspace_bytes < need) {
        set_error("pdu_channel_sum_f32: workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    const long total4 = (long)batch * channels * plane / 4;
    int nblocks = 4 * sm_count();
    if ((long)nblocks > (total4 + 255) / 256) nblocks = (int)((total4 + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    channel_sum_nhwc4_kernel<<<nblocks, 256, 0, st>>>((const float4*)g, (float*)workspace, c4, total4);
    PDU_LAUNCHED();
    partial_rows_sum_kernel<<<channels, 256, 0, st>>>((const float*)workspace, nblocks, channels, out);
    PDU_LAUNCHED();
    return PDU_OK;
}

extern "C" size_t pdu_bias_prelu_bwd_workspace_bytes(int channels) {
    // first-stage partial sums [4 SMs][2][C] + C floats for the shared-slope case
    return channels > 0 ? ((size_t)4 * pdu::sm_count() * 2 + 1) * channels * sizeof(float) : 0;
}

extern "C" int pdu_bias_prelu_bwd_f32(const float* g, const float* y, const float* bias, const float* slope, int n_slope,
                                      float* gz, float* gbias, float* gslope, void* workspace, size_t workspace_bytes,
                                      int batch, int channels, long plane, int layout, pdu_stream_t stream) {
    using namespace pdu;
    PDU_REQUIRE(g && y && bias && slope && gz && gbias && gslope, "pdu_bias_prelu_bwd_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && plane > 0, "pdu_bias_prelu_bwd_f32: sizes must be positive");
    PDU_REQUIRE(n_slope == 1 || n_slope == channels, "pdu_bias_prelu_bwd_f32: slope must have 1 or %d values", channels);
    if (!bias_prelu_train_ok(channels, layout, g, y, gz, bias) || (n_slope != 1 && !al16(slope))) {
        set_error("pdu_bias_prelu_bwd_f32: needs 16-byte aligned channels-last tensors with channels in {4,8,16,32,64,128,256}");
        return PDU_EUNSUPPORTED;
    }
    const size_t need = pdu_bias_prelu_bwd_workspace_bytes(channels);
    if (!workspace || workspace_bytes < need) {
        set_error("pdu_bias_prelu_bwd_f32: workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    const long total4 = (long)batch * channels * plane / 4;
    int nblocks = 4 * sm_count();
    if ((long)nblocks > (total4 + 255) / 256) nblocks = (int)((total4 + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    bias_prelu_bwd_nhwc4_kernel<<<nblocks, 256, 0, st>>>((const float4*)g, (const float4*)y, (const float4*)bias,
                                                         (const float4*)slope, n_slope, (float4*)gz, (float*)workspace,
                                                         channels / 4, total4);
    PDU_LAUNCHED();
    float* slope_tmp = (float*)workspace + (size_t)4 * sm_count() * 2 * channels;
    bias_prelu_bwd_reduce_kernel<<<2 * channels, 256, 0, st>>>((const float*)workspace, nblocks, channels, n_slope, gbias, gslope,
                                                               slope_tmp);
    PDU_LAUNCHED();
    if (n_slope == 1) {
        bias_prelu_bwd_scalar_slope_kernel<<<1, 1, 0, st>>>(slope_tmp, channels, gslope);
        PDU_LAUNCHED();
    }
    return PDU_OK;
}

// ------------------------------------------------------------------ bias + PReLU + skip placement + 2x2 max pool
// The epilogue of a UNet encoder block in one pass over the convolution output y (channels-last):
//   v = prelu(y + bias, slope)  ->  written into its slot of the decoder's concatenation buffer
//                                   (pixel stride skip_pix, channel offset already applied to `skip`)
//   max over each 2x2 window    ->  pooled [batch, H/2, W/2, C]     (pooled == NULL: no pooling)
// ATen runs bias add, PReLU, max-pool and torch.cat as four passes.
namespace pdu {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 bp4(float4 v, float4 b, float4 s, bool has_slope) {
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (has_slope) { v.x = prelu1(v.x, s.x); v.y = prelu1(v.y, s.y); v.z = prelu1(v.z, s.z); v.w = prelu1(v.w, s.w); }
    return v;
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

template <bool POOL>
__global__ void __launch_bounds__(256)
    bias_prelu_place_kernel(const float* __restrict__ y, const float* __restrict__ bias, const float* __restrict__ slope,
                            int n_slope, float* __restrict__ skip, long skip_pix, float* __restrict__ pooled, int H, int W,
                            int C, long total) {
    const int c4n = C >> 2;
    const int Ho = POOL ? H >> 1 : H, Wo = POOL ? W >> 1 : W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        long t = i / c4n;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const long b = t / Ho;
        const float4 bv = ld4(bias + c);
        float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slope) sv = n_slope == 1 ? make_float4(__ldg(slope), __ldg(slope), __ldg(slope), __ldg(slope)) : ld4(slope + c);
        if (POOL) {
            float4 m;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long pix = (b * H + (2 * ho + (q >> 1))) * W + (2 * wo + (q & 1));
                const float4 v = bp4(ld4(y + pix * C + c), bv, sv, slope != nullptr);
                st4(skip + pix * skip_pix + c, v);
                m = q == 0 ? v : max4(m, v);
            }
            st4(pooled + ((b * Ho + ho) * Wo + wo) * (long)C + c, m);
        } else {
            const long pix = (b * H + ho) * W + wo;
            st4(skip + pix * skip_pix + c, bp4(ld4(y + pix * C + c), bv, sv, slope != nullptr));
        }
    }
}

}  // namespace pdu

extern "C" int pdu_bias_prelu_place_f32(const float* y, const float* bias, const float* slope, int n_slope, float* skip,
                                        long skip_pixel_stride, float* pooled, int batch, int channels, int height, int width,
                                        pdu_stream_t stream) {
    PDU_REQUIRE(y && bias && skip, "pdu_bias_prelu_place_f32: null pointer");
    PDU_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0, "pdu_bias_prelu_place_f32: sizes must be positive");
    PDU_REQUIRE(channels % 4 == 0 && skip_pixel_stride % 4 == 0 && skip_pixel_stride >= channels,
                "pdu_bias_prelu_place_f32: channels (%d) and the destination pixel stride (%ld) must be multiples of 4", channels,
                skip_pixel_stride);
    PDU_REQUIRE(al16(y) && al16(bias) && al16(skip) && (!slope || n_slope == 1 || al16(slope)) && (!pooled || al16(pooled)),
                "pdu_bias_prelu_place_f32: pointers must be 16-byte aligned");
    PDU_REQUIRE(slope == nullptr || n_slope == 1 || n_slope == channels, "pdu_bias_prelu_place_f32: slope must have 1 or %d values",
                channels);
    PDU_REQUIRE(!pooled || (height % 2 == 0 && width % 2 == 0), "pdu_bias_prelu_place_f32: pooling needs even height and width");
    cudaStream_t st = (cudaStream_t)stream;
    if (pooled) {
        const long total = (long)batch * (height / 2) * (width / 2) * (channels / 4);
        bias_prelu_place_kernel<true><<<stream_grid(total, 256), 256, 0, st>>>(y, bias, slope, n_slope, skip, skip_pixel_stride,
                                                                               pooled, height, width, channels, total);
    } else {
        const long total = (long)batch * height * width * (channels / 4);
        bias_prelu_place_kernel<false><<<stream_grid(total, 256), 256, 0, st>>>(y, bias, slope, n_slope, skip, skip_pixel_stride,
                                                                                nullptr, height, width, channels, total);
    }
    PDU_LAUNCHED();
    return PDU_OK;
}
