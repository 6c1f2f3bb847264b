// Pruned 2-D FFT for the NUFFT's oversampled grid, in shared memory (replaces the cuFFT calls when the
// grid sizes factor into 2, 3, 5).
//
// Why not cuFFT: the image occupies N of the K = 2N rows and columns of the grid, so the library path
// writes a zero-padded K x K grid (one full pass) and then moves all of it twice more.  Here
//   forward:  row pass    only the N non-zero rows; apodisation (and coil maps) applied while loading,
//                         zero padding happens in shared memory                     -> T [N0][K1]
//             column pass all K1 columns, reading only the N0 rows that exist       -> grid [K0][K1]
//   adjoint:  row pass    all K0 rows, keeping only the first N1 outputs            -> T [K0][N1]
//             column pass the N1 kept columns, keeping only the first N0 outputs    -> U [N0][N1]
// which is 2.5x less traffic than pad + two full passes.
//
// A sequence of length K is transformed by Stockham autosort passes of radix 8 / 4 / 2 / 5 / 3 between
// two shared-memory buffers (natural order in, natural order out); twiddles come from a table computed
// in float64 on the host.  A CTA transforms SEQ sequences at once; for the column pass these are SEQ
// neighbouring columns so that global accesses are SEQ * 8 contiguous bytes.
#pragma once
#include "common.cuh"

namespace pdu {

constexpr int PFFT_MAX_STAGES = 8;

struct PfftPlan {
    int K;
    int n_stages;
    int radix[PFFT_MAX_STAGES];
};

static inline bool pfft_factor(int K, PfftPlan* out) {
    PfftPlan p;
    p.K = K;
    p.n_stages = 0;
    int r = K;
    const int cand[5] = {8, 4, 2, 5, 3};
    for (int c = 0; c < 5; ++c)
        while (r % cand[c] == 0 && r > 1) {
            if (p.n_stages == PFFT_MAX_STAGES) return false;
            p.radix[p.n_stages++] = cand[c];
            r /= cand[c];
        }
    if (r != 1 || K < 2 || K > 4096) return false;
    for (int i = p.n_stages; i < PFFT_MAX_STAGES; ++i) p.radix[i] = 1;
    if (out) *out = p;
    return true;
}

// position of element e inside a sequence: one padding slot every 16 elements, so that the stride-R
// writes of the first pass (and stride-8 reads of later ones) spread over the banks
__host__ __device__ __forceinline__ int pf_pos(int e) { return e + (e >> 4); }
// row pitch of a [SEQ][.] buffer: == 2 (mod 16) float2, so SEQ neighbouring sequences at the same element
// land in different banks (column pass loads / stores)
static inline int pfft_pitch(int K) { return (pf_pos(K) + 15) / 16 * 16 + 2; }

__device__ __forceinline__ float2 pf_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 pf_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 pf_mul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 pf_rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

template <bool INV>
__device__ __forceinline__ void pf_r2(float2* v) {
    const float2 a = v[0], b = v[1];
    v[0] = pf_add(a, b);
    v[1] = pf_sub(a, b);
}
template <bool INV>
__device__ __forceinline__ void pf_r4(float2* v) {
    const float2 a = pf_add(v[0], v[2]), b = pf_sub(v[0], v[2]);
    const float2 c = pf_add(v[1], v[3]), d = pf_rot<INV>(pf_sub(v[1], v[3]));
    v[0] = pf_add(a, c);
    v[1] = pf_add(b, d);
    v[2] = pf_sub(a, c);
    v[3] = pf_sub(b, d);
}
template <bool INV>
__device__ __forceinline__ void pf_r8(float2* v) {
    constexpr float H = 0.70710678118654752f;
    // two radix-4 transforms of the even and odd inputs, then the W8 twiddles
    float2 e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    pf_r4<INV>(e);
    pf_r4<INV>(o);
    const float2 w1 = INV ? make_float2(H, H) : make_float2(H, -H);
    const float2 w3 = INV ? make_float2(-H, H) : make_float2(-H, -H);
    const float2 o1 = pf_mul(o[1], w1), o2 = pf_rot<INV>(o[2]), o3 = pf_mul(o[3], w3);
    v[0] = pf_add(e[0], o[0]); v[4] = pf_sub(e[0], o[0]);
    v[1] = pf_add(e[1], o1);   v[5] = pf_sub(e[1], o1);
    v[2] = pf_add(e[2], o2);   v[6] = pf_sub(e[2], o2);
    v[3] = pf_add(e[3], o3);   v[7] = pf_sub(e[3], o3);
}
// small odd radices: direct DFT with the table (R * R complex multiplies; only one such stage per transform)
template <bool INV, int R>
__device__ __forceinline__ void pf_rodd(float2* v, const float2* __restrict__ tw, int K) {
    float2 out[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
        float2 acc = v[0];
#pragma unroll
        for (int r = 1; r < R; ++r) {
            float2 w = tw[((q * r) % R) * (K / R)];
            if (INV) w.y = -w.y;
            acc = pf_add(acc, pf_mul(v[r], w));
        }
        out[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = out[q];
}

// one Stockham pass of radix R over the CTA's sequences: src/dst are [SEQ][pitch].  Thread t works on
// sequence t / tps with tps = nthreads / SEQ threads striding its butterflies, so the loop has no
// division by run-time values; Ns is a power of two except possibly in the last odd-radix pass.
template <bool INV, int R>
__device__ __forceinline__ void pf_pass(const float2* __restrict__ src, float2* __restrict__ dst, const float2* __restrict__ tw,
                                        int K, int Ns, int pitch, int s, int t_in_seq, int tps) {
    const int nb = K / R;                  // butterflies per sequence
    const int tstep = K / (Ns * R);        // table stride of this pass
    const bool pow2 = (Ns & (Ns - 1)) == 0;
    const int sh = 31 - __clz(Ns);
    const float2* in = src + s * pitch;
    float2* out = dst + s * pitch;
    for (int j = t_in_seq; j < nb; j += tps) {
        const int hi = pow2 ? (j >> sh) : (j / Ns);
        const int k = j - hi * Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[pf_pos(j + r * nb)];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                float2 w = tw[r * k * tstep];          // r k tstep < R Ns tstep = K: no wrap
                if (INV) w.y = -w.y;
                v[r] = pf_mul(v[r], w);
            }
        }
        if constexpr (R == 2) pf_r2<INV>(v);
        else if constexpr (R == 4) pf_r4<INV>(v);
        else if constexpr (R == 8) pf_r8<INV>(v);
        else pf_rodd<INV, R>(v, tw, K);
        const int o0 = hi * Ns * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[pf_pos(o0 + r * Ns)] = v[r];
    }
}

// transforms the CTA's SEQ sequences held in buf0 ([SEQ][pitch]); returns the buffer that holds the result.
// nthreads must be a multiple of SEQ.
template <bool INV>
__device__ __forceinline__ float2* pf_transform(float2* buf0, float2* buf1, const float2* __restrict__ tw, const PfftPlan& pl,
                                                int seq, int pitch, int tid, int nthreads) {
    float2* src = buf0;
    float2* dst = buf1;
    const int tps = nthreads / seq;
    const int s = tid / tps, t = tid - s * tps;
    int Ns = 1;
    for (int st = 0; st < pl.n_stages; ++st) {
        const int R = pl.radix[st];
        if (R == 8) pf_pass<INV, 8>(src, dst, tw, pl.K, Ns, pitch, s, t, tps);
        else if (R == 4) pf_pass<INV, 4>(src, dst, tw, pl.K, Ns, pitch, s, t, tps);
        else if (R == 2) pf_pass<INV, 2>(src, dst, tw, pl.K, Ns, pitch, s, t, tps);
        else if (R == 5) pf_pass<INV, 5>(src, dst, tw, pl.K, Ns, pitch, s, t, tps);
        else pf_pass<INV, 3>(src, dst, tw, pl.K, Ns, pitch, s, t, tps);
        Ns *= R;
        __syncthreads();
        float2* x = src; src = dst; dst = x;
    }
    return src;
}

}  // namespace pdu
