// Register-resident pruned FFT for the two oversampled-grid sizes of the BASELINE MRI shapes
// (K = 512 for 256^2 images, K = 640 for 320^2): compile-time radices 8 x 8 x (K / 64), three
// butterflies per thread with two shared-memory exchanges in between.  The first butterfly reads
// its inputs straight from global memory and the last one writes its outputs straight back, so a
// sequence crosses shared memory twice (pfft.cuh's generic run-time-radix passes: five times) and
// there is no index arithmetic that is not a shift or a compile-time constant.
//
// Stockham autosort, natural order in and out; pass with radix R, Ns = product of earlier radices:
//   butterfly j (0 <= j < K/R):  k = j mod Ns, hi = j / Ns
//     v[r] = in[j + r K/R] * W_K^(r k K/(Ns R)),  v = DFT_R(v),  out[hi Ns R + k + r Ns] = v[r]
// Pruning: HALF_IN  -- inputs K/2 .. K-1 are zero (zero padding of the image), never loaded;
//          HALF_OUT -- outputs K/2 .. K-1 are not wanted (the crop of the adjoint), never stored.
#pragma once
#include "pfft.cuh"

namespace pdu {

template <int K>
struct FastFft {
    static constexpr bool ok = (K == 512 || K == 640);
    static constexpr int TPS = K / 8;     // threads per sequence (radix-8 butterflies per pass)
    static constexpr int R3 = K / 64;     // radix of the last pass: 8 or 10
    // Padding: one slot every 2^PS elements, chosen per thread layout (measured, ncu r02):
    //  PS = 3 (column kernels: a half-warp is 8 neighbouring sequences x 2 butterflies) -- element 8 t + r sits at
    //    9 t + r; with the pitch == 2 (mod 16) float2 every access of the three passes is conflict free;
    //  PS = 4 (row kernels: a half-warp is 16 consecutive butterflies of one sequence) -- the unit-stride loads of
    //    passes 2 and 3 then never straddle a padding slot (PS = 3 made 16 lanes span 18 slots: a 2-way conflict on
    //    every load), the stride-8 stores of pass 1 stay conflict free, only the pass-2 stores keep a 2-way conflict.
    template <int PS>
    __host__ __device__ static constexpr int pitch() { return (K + (K >> PS) + 13) / 16 * 16 + 2; }
};
template <int PS>
__device__ __forceinline__ int ff_pos(int e) { return e + (e >> PS); }

template <bool INV>
__device__ __forceinline__ float2 ff_tw(const float2* __restrict__ tw, int i) {
    float2 w = tw[i];
    if (INV) w.y = -w.y;
    return w;
}

template <bool INV>
__device__ __forceinline__ void ff_r5(float2* x) {
    constexpr float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;
    constexpr float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;
    const float2 a1 = pf_add(x[1], x[4]), a2 = pf_add(x[2], x[3]);
    const float2 b1 = pf_sub(x[1], x[4]), b2 = pf_sub(x[2], x[3]);
    const float2 t1 = make_float2(x[0].x + C1 * a1.x + C2 * a2.x, x[0].y + C1 * a1.y + C2 * a2.y);
    const float2 t2 = make_float2(x[0].x + C2 * a1.x + C1 * a2.x, x[0].y + C2 * a1.y + C1 * a2.y);
    const float2 u1 = pf_rot<INV>(make_float2(S1 * b1.x + S2 * b2.x, S1 * b1.y + S2 * b2.y));
    const float2 u2 = pf_rot<INV>(make_float2(S2 * b1.x - S1 * b2.x, S2 * b1.y - S1 * b2.y));
    x[0] = pf_add(x[0], pf_add(a1, a2));
    x[1] = pf_add(t1, u1);
    x[4] = pf_sub(t1, u1);
    x[2] = pf_add(t2, u2);
    x[3] = pf_sub(t2, u2);
}

template <bool INV>
__device__ __forceinline__ void ff_r10(float2* v) {
    float2 e[5] = {v[0], v[2], v[4], v[6], v[8]}, o[5] = {v[1], v[3], v[5], v[7], v[9]};
    ff_r5<INV>(e);
    ff_r5<INV>(o);
    // W10^q = exp(-+ 2 pi i q / 10)
    constexpr float C[5] = {1.f, 0.80901699437494742f, 0.30901699437494742f, -0.30901699437494742f, -0.80901699437494742f};
    constexpr float S[5] = {0.f, 0.58778525229247313f, 0.95105651629515357f, 0.95105651629515357f, 0.58778525229247313f};
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const float2 w = make_float2(C[q], INV ? S[q] : -S[q]);
        const float2 t = q == 0 ? o[0] : pf_mul(o[q], w);
        v[q] = pf_add(e[q], t);
        v[q + 5] = pf_sub(e[q], t);
    }
}

// One sequence of length K held by the TPS threads t = 0 .. TPS-1 of a CTA (every thread of the CTA must call
// this: it synchronises).  buf = this sequence's pitch<PS>() float2 of shared memory, tw = the K-entry table
// exp(-2 pi i m / K) in shared memory.  ld(e) returns input element e, st(e, v) consumes output element e.
template <int K, int PS, bool INV, bool HALF_IN, bool HALF_OUT, class LD, class ST>
__device__ __forceinline__ void ff_transform(float2* __restrict__ buf, const float2* __restrict__ tw, int t, LD ld, ST st) {
    using F = FastFft<K>;
    constexpr int TPS = F::TPS, R3 = F::R3;
    float2 v[8];
    // pass 1: radix 8, Ns = 1
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (HALF_IN && r >= 4) ? make_float2(0.f, 0.f) : ld(t + r * TPS);
    pf_r8<INV>(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[ff_pos<PS>(t * 8 + r)] = v[r];
    __syncthreads();
    // pass 2: radix 8, Ns = 8
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = buf[ff_pos<PS>(t + r * TPS)];
    __syncthreads();
    {
        const int k = t & 7;
#pragma unroll
        for (int r = 1; r < 8; ++r) v[r] = pf_mul(v[r], ff_tw<INV>(tw, r * k * (K / 64)));
        pf_r8<INV>(v);
        const int o0 = (t >> 3) * 64 + k;
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[ff_pos<PS>(o0 + r * 8)] = v[r];
    }
    __syncthreads();
    // pass 3: radix K / 64, Ns = 64; 64 butterflies
    if (t < 64) {
        float2 u[R3];
#pragma unroll
        for (int r = 0; r < R3; ++r) u[r] = buf[ff_pos<PS>(t + r * 64)];
#pragma unroll
        for (int r = 1; r < R3; ++r) u[r] = pf_mul(u[r], ff_tw<INV>(tw, r * t));
        if constexpr (R3 == 8) pf_r8<INV>(u);
        else ff_r10<INV>(u);
#pragma unroll
        for (int r = 0; r < R3; ++r)
            if (!HALF_OUT || r < R3 / 2) st(t + r * 64, u[r]);
    }
}

}  // namespace pdu
