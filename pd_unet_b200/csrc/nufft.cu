// Radial-MRI non-uniform FFT: Kaiser-Bessel table interpolation around an oversampled FFT.
// Replaces [RECALL] torchkbnufft KbNufft / KbNufftAdjoint / KbInterp / KbInterpAdjoint, which are
// compositions of ATen ops (complex multiply, pad, torch.fft, index arithmetic, index_add_).
//
//   forward :  grids 512 / 640 (the BASELINE shapes): ff_rows_fwd_kernel (apodise, * smaps, pruned row FFT of the
//                N non-zero rows) ; ff_cols_fwd_kernel (pruned column FFT)        -- pfft_fast.cuh, 2 passes
//              other grids: apod_pad_kernel + cuFFT, or the generic pruned FFT of pfft.cuh (variant 1)
//              interp_fwd_kernel J x J table-weighted gather per k-space sample, * phase
//   adjoint :  trajectory used more than once: sorted (CSR) gather interp_adj_csrT_kernel on plane-interleaved
//                kdata -- no atomics, no memset, reproducible; else memset + interp_adj_kernel (float2 atomics)
//              inverse FFT: ff_rows_adj_kernel / ff_cols_adj_kernel (only the N kept outputs per axis), or cuFFT
//              crop_apod_kernel  crop * scaling_coef (* conj(smaps), summed over coils)
//
// Grid offsets and table indices are computed with explicitly rounded float32 operations, the same
// sequence oracle/nufft.py::_tap_indices_f32 performs, so both read the same table entries.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "nufft_common.cuh"

namespace pdu {

// ------------------------------------------------------------------ apodise + zero pad
// one thread = two neighbouring grid cells (one 16-byte store); three quarters of the stores are zeros
__global__ void __launch_bounds__(256)
    apod_pad_kernel(const float2* __restrict__ image, const float2* __restrict__ smaps, float4* __restrict__ grid,
                    const float* __restrict__ s0, const float* __restrict__ s1, NufftDims d, int coils, int smaps_batch,
                    long total2) {
    const int k1h = d.k1 >> 1;
    const long plane = (long)d.n0 * d.n1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (long)gridDim.x * blockDim.x) {
        const int c1 = (int)(i % k1h) * 2;
        const long t = i / k1h;
        const int c0 = (int)(t % d.k0);
        const long p = t / d.k0;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 < d.n0 && c1 < d.n1) {
            const long b = p / coils, c = p - b * coils;
            const float w0 = __ldg(s0 + c0);
            float2 v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                v[e] = make_float2(0.f, 0.f);
                if (c1 + e < d.n1) {
                    const long pix = (long)c0 * d.n1 + c1 + e;
                    if (smaps) {
                        const long sb = smaps_batch == 1 ? 0 : b;
                        v[e] = cmul(__ldg(image + b * plane + pix), __ldg(smaps + (sb * coils + c) * plane + pix));
                    } else {
                        v[e] = __ldg(image + p * plane + pix);
                    }
                    const float w = w0 * __ldg(s1 + c1 + e);
                    v[e].x *= w;
                    v[e].y *= w;
                }
            }
            out = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        }
        grid[i] = out;
    }
}

// ------------------------------------------------------------------ crop + apodise (+ coil combine)
__global__ void __launch_bounds__(256)
    crop_apod_kernel(const float2* __restrict__ grid, const float2* __restrict__ smaps, float2* __restrict__ image,
                     const float* __restrict__ s0, const float* __restrict__ s1, NufftDims d, int coils, int smaps_batch,
                     float scale, long total, int split = 0) {
    const long plane = (long)d.n0 * d.n1;
    const long gplane = (long)d.k0 * d.k1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c1 = (int)(i % d.n1);
        const long t = i / d.n1;
        const int c0 = (int)(t % d.n0);
        const long q = t / d.n0;          // output plane: b (with smaps) or b * coils + c
        const float w = __ldg(s0 + c0) * __ldg(s1 + c1) * scale;
        const long gp = (long)c0 * d.k1 + c1;
        float2 v;
        if (smaps) {
            const long sb = smaps_batch == 1 ? 0 : q;
            // coils in groups of four: eight independent loads in flight per thread (the serial loop was
            // latency bound: ncu long-scoreboard 19 cycles per issue)
            float2 acc = make_float2(0.f, 0.f);
            const float2* gq = grid + q * coils * gplane + gp;
            const float2* sq = smaps + sb * coils * plane + (long)c0 * d.n1 + c1;
            int c = 0;
            for (; c + 4 <= coils; c += 4) {
                float2 gv[4], sv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    gv[u] = __ldg(gq + (c + u) * gplane);
                    sv[u] = __ldg(sq + (c + u) * plane);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 z = cmul_conj(gv[u], sv[u]);
                    acc.x += z.x;
                    acc.y += z.y;
                }
            }
            for (; c < coils; ++c) {
                const float2 z = cmul_conj(__ldg(gq + c * gplane), __ldg(sq + c * plane));
                acc.x += z.x;
                acc.y += z.y;
            }
            v = acc;
        } else {
            v = __ldg(grid + q * gplane + gp);
        }
        if (split) {              // [planes][2][n0][n1] float32: real plane, imaginary plane
            float* o = reinterpret_cast<float*>(image) + 2 * q * plane + (long)c0 * d.n1 + c1;
            o[0] = v.x * w;
            o[plane] = v.y * w;
        } else {
            image[i] = make_float2(v.x * w, v.y * w);
        }
    }
}

// ------------------------------------------------------------------ interpolation (gather)
// thread = one sample m; loops over the PC planes of its plane chunk (blockIdx.y)
template <int JT>
__global__ void __launch_bounds__(128)    // (r02: capping at 64 registers for 32 warps / SM spills and is not faster)
    interp_fwd_kernel(const float2* __restrict__ grid, float2* __restrict__ kdata, const float* __restrict__ omega,
                      const float2* __restrict__ t0, const float2* __restrict__ t1, NufftDims d, int planes, int pc,
                      long M, float scale) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int J = JT > 0 ? JT : d.J;
    constexpr int N = JT > 0 ? JT : MAXJ;
    const float om0 = __ldg(omega + m), om1 = __ldg(omega + M + m);
    int g0[N], g1[N];
    float2 c0[N], c1[N];
    axis_taps<JT>(om0, d.gam0, d.k0, d.J, d.L, t0, g0, c0);
    axis_taps<JT>(om1, d.gam1, d.k1, d.J, d.L, t1, g1, c1);
    float2 ph = shift_phase(om0, om1, d.shift0, d.shift1);
    ph.x *= scale;
    ph.y *= scale;
    const long gplane = (long)d.k0 * d.k1;
    const int p_end = min(planes, ((int)blockIdx.y + 1) * pc);
    for (int p = blockIdx.y * pc; p < p_end; ++p) {
        const float2* gp = grid + p * gplane;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (a < J) {
                const float2* row = gp + (long)g0[a] * d.k1;
                float2 racc = make_float2(0.f, 0.f);
#pragma unroll
                for (int b = 0; b < N; ++b) {
                    if (b < J) {
                        const float2 z = cmul(__ldg(row + g1[b]), c1[b]);
                        racc.x += z.x;
                        racc.y += z.y;
                    }
                }
                const float2 z = cmul(racc, c0[a]);
                acc.x += z.x;
                acc.y += z.y;
            }
        }
        kdata[(long)p * M + m] = cmul(acc, ph);
    }
}

// ------------------------------------------------------------------ interpolation adjoint (scatter)
template <int JT>
__global__ void __launch_bounds__(128)
    interp_adj_kernel(const float2* __restrict__ kdata, float2* __restrict__ grid, const float* __restrict__ omega,
                      const float2* __restrict__ t0, const float2* __restrict__ t1, NufftDims d, int planes, int pc,
                      long M) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int J = JT > 0 ? JT : d.J;
    constexpr int N = JT > 0 ? JT : MAXJ;
    const float om0 = __ldg(omega + m), om1 = __ldg(omega + M + m);
    int g0[N], g1[N];
    float2 c0[N], c1[N];
    axis_taps<JT>(om0, d.gam0, d.k0, d.J, d.L, t0, g0, c0);
    axis_taps<JT>(om1, d.gam1, d.k1, d.J, d.L, t1, g1, c1);
    const float2 ph = shift_phase(om0, om1, d.shift0, d.shift1);
    const long gplane = (long)d.k0 * d.k1;
    const int p_end = min(planes, ((int)blockIdx.y + 1) * pc);
    for (int p = blockIdx.y * pc; p < p_end; ++p) {
        const float2 y = cmul_conj(__ldg(kdata + (long)p * M + m), ph);
        float2* gp = grid + p * gplane;
#pragma unroll
        for (int a = 0; a < N; ++a) {
            if (a < J) {
                const float2 ya = cmul_conj(y, c0[a]);
                float2* row = gp + (long)g0[a] * d.k1;
#pragma unroll
                for (int b = 0; b < N; ++b) {
                    if (b < J) atomicAdd(row + g1[b], cmul_conj(ya, c1[b]));
                }
            }
        }
    }
}

static unsigned stream_grid(long items) {
    const long want = cdiv(items, 256);
    const long cap = (long)sm_count() * 16;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

static int plane_chunk(int planes, long M) {
    // enough CTAs for two waves when the problem allows it, otherwise reuse the taps across planes
    // one plane per thread until there are more than ~16 CTAs per SM slot (measured: the gather is latency
    // bound, 2.7 us / plane at 4 planes per thread against 1.9 us at 1)
    const long ctas_m = cdiv(M, 128);
    int pc = 1;
    while (pc < planes && ctas_m * cdiv(planes, pc) > 16L * 16 * sm_count()) pc *= 2;
    return pc;
}

// The cuFFT handle of a plan is shared state (stream binding, work area): the plan's mutex is held from the
// cufftSetStream to the end of the exec call, so two host threads that share a plan cannot retarget each other's
// transform.  (Kernels of one handle still run in the order they were enqueued; use one plan per stream for overlap.)
static int run_fft(pdu_nufft_plan* p, float2* grid, int planes, int dir, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(p->mu);
    auto it = p->fft.find(planes);
    if (it == p->fft.end()) {
        cufftHandle h;
        int n[2] = {p->k0, p->k1};
        cufftResult r = cufftPlanMany(&h, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2C, planes);
        if (r != CUFFT_SUCCESS) {
            set_error("cufftPlanMany(%d x %d, batch %d) failed with %d", p->k0, p->k1, planes, (int)r);
            return PDU_EFFT;
        }
        it = p->fft.emplace(planes, h).first;
    }
    cufftResult r = cufftSetStream(it->second, st);
    if (r != CUFFT_SUCCESS) {
        set_error("cufftSetStream failed with %d", (int)r);
        return PDU_EFFT;
    }
    r = cufftExecC2C(it->second, (cufftComplex*)grid, (cufftComplex*)grid, dir);
    if (r != CUFFT_SUCCESS) {
        set_error("cufftExecC2C failed with %d", (int)r);
        return PDU_EFFT;
    }
    count_launch(2);   // a 2-D C2C transform is at least two library kernels
    return PDU_OK;
}

static int launch_interp_fwd(pdu_nufft_plan* p, const float2* grid, float2* kdata, const float* omega, int planes,
                             long M, float scale, cudaStream_t st) {
    const int pc = plane_chunk(planes, M);
    dim3 g((unsigned)cdiv(M, 128), (unsigned)cdiv(planes, pc));
    if (p->J == 6) interp_fwd_kernel<6><<<g, 128, 0, st>>>(grid, kdata, omega, p->d_t0, p->d_t1, dims_of(p), planes, pc, M, scale);
    else interp_fwd_kernel<0><<<g, 128, 0, st>>>(grid, kdata, omega, p->d_t0, p->d_t1, dims_of(p), planes, pc, M, scale);
    PDU_LAUNCHED();
    return PDU_OK;
}

static int launch_interp_adj(pdu_nufft_plan* p, const float2* kdata, float2* grid, const float* omega, int planes,
                             long M, cudaStream_t st) {
    const int pc = plane_chunk(planes, M);
    dim3 g((unsigned)cdiv(M, 128), (unsigned)cdiv(planes, pc));
    if (p->J == 6) interp_adj_kernel<6><<<g, 128, 0, st>>>(kdata, grid, omega, p->d_t0, p->d_t1, dims_of(p), planes, pc, M);
    else interp_adj_kernel<0><<<g, 128, 0, st>>>(kdata, grid, omega, p->d_t0, p->d_t1, dims_of(p), planes, pc, M);
    PDU_LAUNCHED();
    return PDU_OK;
}

// ------------------------------------------------------------------ pruned FFT passes (pfft.cuh)
template <typename T>
__device__ __forceinline__ T* pf_smem() {
    extern __shared__ __align__(16) unsigned char pf_dyn[];
    return reinterpret_cast<T*>(pf_dyn);
}

// forward, along the last axis: SEQ image rows of one plane -> T [planes][n0][k1]
template <int SEQ>
__global__ void __launch_bounds__(256)
    pfft_rows_fwd_kernel(const float2* __restrict__ image, const float2* __restrict__ smaps, float2* __restrict__ T,
                         const float* __restrict__ s0, const float* __restrict__ s1, const float2* __restrict__ tw_g,
                         PfftPlan pl, NufftDims d, int coils, int smaps_batch, int pitch) {
    float2* buf0 = pf_smem<float2>();
    float2* buf1 = buf0 + SEQ * pitch;
    float2* tw = buf1 + SEQ * pitch;
    const int tid = threadIdx.x, K = pl.K;
    const long p = blockIdx.y;
    const int r0 = blockIdx.x * SEQ;
    for (int i = tid; i < K; i += blockDim.x) tw[i] = __ldg(tw_g + i);
    const long plane = (long)d.n0 * d.n1;
    const long b = p / coils, c = p - b * coils;
    const int tps = blockDim.x / SEQ, s = tid / tps, row = r0 + s;      // a fixed row per thread: no run-time division below
    for (int col = tid - s * tps; col < K; col += tps) {
        float2 v = make_float2(0.f, 0.f);
        if (row < d.n0 && col < d.n1) {
            const long pix = (long)row * d.n1 + col;
            if (smaps) {
                const long sb = smaps_batch == 1 ? 0 : b;
                v = cmul(__ldg(image + b * plane + pix), __ldg(smaps + (sb * coils + c) * plane + pix));
            } else {
                v = __ldg(image + p * plane + pix);
            }
            const float w = __ldg(s0 + row) * __ldg(s1 + col);
            v.x *= w;
            v.y *= w;
        }
        buf0[s * pitch + pf_pos(col)] = v;
    }
    __syncthreads();
    const float2* res = pf_transform<false>(buf0, buf1, tw, pl, SEQ, pitch, tid, blockDim.x);
    if (row < d.n0)
        for (int col = tid - s * tps; col < K; col += tps) T[(p * d.n0 + row) * K + col] = res[s * pitch + pf_pos(col)];
}

// forward, along the first axis: SEQ neighbouring columns of T [planes][n0][k1] -> grid [planes][k0][k1]
template <int SEQ>
__global__ void __launch_bounds__(256)
    pfft_cols_fwd_kernel(const float2* __restrict__ T, float2* __restrict__ grid, const float2* __restrict__ tw_g, PfftPlan pl,
                         NufftDims d, int pitch) {
    float2* buf0 = pf_smem<float2>();
    float2* buf1 = buf0 + SEQ * pitch;
    float2* tw = buf1 + SEQ * pitch;
    const int tid = threadIdx.x, K = pl.K;           // K == k0
    const long p = blockIdx.y;
    const int c0 = blockIdx.x * SEQ;
    for (int i = tid; i < K; i += blockDim.x) tw[i] = __ldg(tw_g + i);
    for (int i = tid; i < SEQ * K; i += blockDim.x) {
        const int row = i / SEQ, s = i - row * SEQ, col = c0 + s;
        float2 v = make_float2(0.f, 0.f);
        if (row < d.n0 && col < d.k1) v = __ldg(T + (p * d.n0 + row) * d.k1 + col);
        buf0[s * pitch + pf_pos(row)] = v;
    }
    __syncthreads();
    const float2* res = pf_transform<false>(buf0, buf1, tw, pl, SEQ, pitch, tid, blockDim.x);
    for (int i = tid; i < SEQ * K; i += blockDim.x) {
        const int row = i / SEQ, s = i - row * SEQ, col = c0 + s;
        if (col < d.k1) grid[(p * K + row) * d.k1 + col] = res[s * pitch + pf_pos(row)];
    }
}

// adjoint (unnormalised inverse), along the last axis: SEQ grid rows -> T [planes][k0][n1] (first n1 outputs kept)
template <int SEQ>
__global__ void __launch_bounds__(256)
    pfft_rows_adj_kernel(const float2* __restrict__ grid, float2* __restrict__ T, const float2* __restrict__ tw_g, PfftPlan pl,
                         NufftDims d, int pitch) {
    float2* buf0 = pf_smem<float2>();
    float2* buf1 = buf0 + SEQ * pitch;
    float2* tw = buf1 + SEQ * pitch;
    const int tid = threadIdx.x, K = pl.K;           // K == k1
    const long p = blockIdx.y;
    const int r0 = blockIdx.x * SEQ;
    for (int i = tid; i < K; i += blockDim.x) tw[i] = __ldg(tw_g + i);
    const int tps = blockDim.x / SEQ, s = tid / tps, row = r0 + s;
    for (int col = tid - s * tps; col < K; col += tps)
        buf0[s * pitch + pf_pos(col)] = row < d.k0 ? __ldg(grid + (p * d.k0 + row) * K + col) : make_float2(0.f, 0.f);
    __syncthreads();
    const float2* res = pf_transform<true>(buf0, buf1, tw, pl, SEQ, pitch, tid, blockDim.x);
    if (row < d.k0)
        for (int col = tid - s * tps; col < d.n1; col += tps) T[(p * d.k0 + row) * d.n1 + col] = res[s * pitch + pf_pos(col)];
}

// adjoint, along the first axis: SEQ neighbouring columns of T [planes][k0][n1] -> U [planes][n0][n1]
template <int SEQ>
__global__ void __launch_bounds__(256)
    pfft_cols_adj_kernel(const float2* __restrict__ T, float2* __restrict__ U, const float2* __restrict__ tw_g, PfftPlan pl,
                         NufftDims d, int pitch) {
    float2* buf0 = pf_smem<float2>();
    float2* buf1 = buf0 + SEQ * pitch;
    float2* tw = buf1 + SEQ * pitch;
    const int tid = threadIdx.x, K = pl.K;           // K == k0
    const long p = blockIdx.y;
    const int c0 = blockIdx.x * SEQ;
    for (int i = tid; i < K; i += blockDim.x) tw[i] = __ldg(tw_g + i);
    for (int i = tid; i < SEQ * K; i += blockDim.x) {
        const int row = i / SEQ, s = i - row * SEQ, col = c0 + s;
        buf0[s * pitch + pf_pos(row)] = col < d.n1 ? __ldg(T + (p * K + row) * d.n1 + col) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const float2* res = pf_transform<true>(buf0, buf1, tw, pl, SEQ, pitch, tid, blockDim.x);
    for (int i = tid; i < SEQ * d.n0; i += blockDim.x) {
        const int row = i / SEQ, s = i - row * SEQ, col = c0 + s;
        if (col < d.n1) U[(p * d.n0 + row) * d.n1 + col] = res[s * pitch + pf_pos(row)];
    }
}

// ------------------------------------------------------------------ register-resident pruned FFT (pfft_fast.cuh)
// Same four passes as above for K = 512 / 640 (grid = 2 x image).  Row kernels: thread = (row s, butterfly t),
// consecutive threads along the row; column kernels: consecutive threads across the SEQ neighbouring columns so
// that a warp touches whole 64-byte runs of every grid row.
// These kernels run only for grid == 2 x image (ff_supported): N = K / 2 is a compile-time constant, every element
// address is base(thread) + r * constant.
template <int K, int SEQ>
__global__ void __launch_bounds__(SEQ* FastFft<K>::TPS)
    ff_rows_fwd_kernel(const float2* __restrict__ image, const float2* __restrict__ smaps, float2* __restrict__ T,
                       const float* __restrict__ s0, const float* __restrict__ s1, const float2* __restrict__ tw_g, NufftDims d,
                       int coils, int smaps_batch) {
    using F = FastFft<K>;
    constexpr int N = K / 2, TPS = F::TPS, NS3 = F::NS3;
    float2* buf = pf_smem<float2>();
    float2* tw = buf + SEQ * F::template pitch<4>();
    const int tid = threadIdx.x, s = tid / TPS, t = tid - s * TPS;
    const int p = blockIdx.y;
    const int row = blockIdx.x * SEQ + s;
    for (int i = tid; i < K; i += SEQ * TPS) tw[i] = __ldg(tw_g + i);
    const int b = p / coils, c = p - b * coils;
    const bool live = row < N;
    const float w0 = live ? __ldg(s0 + row) : 0.f;
    const float2* src = image + ((long)(smaps ? b : p) * N + row) * N + t;
    const float2* sm = smaps ? smaps + ((long)((smaps_batch == 1 ? 0 : b) * coils + c) * N + row) * N + t : nullptr;
    const float* s1t = s1 + t;
    auto ld = [&](int r) {                     // element t + r TPS < N
        if (!live) return make_float2(0.f, 0.f);
        float2 v = __ldg(src + r * TPS);
        if (sm) v = cmul(v, __ldg(sm + r * TPS));
        const float w = w0 * __ldg(s1t + r * TPS);
        return make_float2(v.x * w, v.y * w);
    };
    float2* dst = T + ((long)p * N + row) * K;
    auto st = [&](int j, int r, float2 v) {
        if (live) dst[j + r * NS3] = v;
    };
    ff_transform<K, 4, false, true, false, false, false, true>(buf + s * F::template pitch<4>(), tw, t, ld, st);
}

template <int K, int SEQ>
__global__ void __launch_bounds__(SEQ* FastFft<K>::TPS, K <= 640 ? 3 : 1)
    ff_cols_fwd_kernel(const float2* __restrict__ T, float2* __restrict__ grid, const float2* __restrict__ tw_g, NufftDims d) {
    using F = FastFft<K>;
    constexpr int N = K / 2, TPS = F::TPS, NS3 = F::NS3;
    float2* buf = pf_smem<float2>();
    float2* tw = buf + SEQ * F::template pitch<3>();
    const int tid = threadIdx.x, t = tid / SEQ, s = tid - t * SEQ;
    const int p = blockIdx.y;
    const int col = blockIdx.x * SEQ + s;
    for (int i = tid; i < K; i += SEQ * TPS) tw[i] = __ldg(tw_g + i);
    const bool live = col < K;
    const float2* src = T + ((long)p * N + t) * K + col;
    float2* dst = grid + (long)p * K * K + col;
    auto ld = [&](int r) { return live ? __ldcs(src + r * (TPS * K)) : make_float2(0.f, 0.f); };
    auto st = [&](int j, int r, float2 v) {
        if (live) __stcs(dst + j * K + r * (NS3 * K), v);
    };
    ff_transform<K, 3, false, true, false>(buf + s * F::template pitch<3>(), tw, t, ld, st);
}

template <int K, int SEQ>
__global__ void __launch_bounds__(SEQ* FastFft<K>::TPS)
    ff_rows_adj_kernel(const float2* __restrict__ grid, float2* __restrict__ T, const float2* __restrict__ tw_g, NufftDims d) {
    using F = FastFft<K>;
    constexpr int N = K / 2, TPS = F::TPS, NS3 = F::NS3;
    float2* buf = pf_smem<float2>();
    float2* tw = buf + SEQ * F::template pitch<4>();
    const int tid = threadIdx.x, s = tid / TPS, t = tid - s * TPS;
    const int p = blockIdx.y;
    const int row = blockIdx.x * SEQ + s;
    for (int i = tid; i < K; i += SEQ * TPS) tw[i] = __ldg(tw_g + i);
    const bool live = row < K;
    const float2* src = grid + ((long)p * K + row) * K + t;
    float2* dst = T + ((long)p * K + row) * N;
    auto ld = [&](int r) { return live ? __ldcs(src + r * TPS) : make_float2(0.f, 0.f); };
    auto st = [&](int j, int r, float2 v) {
        if (live) __stcs(dst + j + r * NS3, v);
    };
    ff_transform<K, 4, true, false, true, false, false, true>(buf + s * F::template pitch<4>(), tw, t, ld, st);
}

// The row pass on the COMPACT gridded samples (interp_adj_csrT_kernel<.., true>): a thread looks up the compact index of
// each of its eight inputs (nz_idx, 4 bytes per cell, coalesced and L2 resident) and loads the value only where the cell
// is non-empty.  Against the dense form the gather writes and this pass reads one value per non-empty cell instead of K
// per row.  Branch-free: an empty cell reads the plane's first value (one shared sector) and discards it, so that the
// eight indices and then the eight values are in flight together -- with a branch per input the loads went out one at
// a time (95 us for this pass at 64 planes of 640^2).  Other forms measured there: the non-empty cells spread over a
// zeroed shared-memory row and the transform started from it (three more CTA barriers: 86 us against 72 for the dense
// pass in spite of 94 instead of 210 MB); occupancy words + population count instead of the index array (two loads and
// ~6 more instructions per input: 4 us more for the call; as two separate arrays 12 us more); six CTAs per SM (32
// registers, maximum shared-memory carve-out) instead of four: 4 us more.
template <int K, int SEQ>
__global__ void __launch_bounds__(SEQ* FastFft<K>::TPS)
    ff_rows_adj_compact_kernel(const float2* __restrict__ gridc, float2* __restrict__ T, const float2* __restrict__ tw_g,
                               const int* __restrict__ nz_idx, NufftDims d) {
    using F = FastFft<K>;
    constexpr int N = K / 2, TPS = F::TPS, NS3 = F::NS3;
    float2* buf = pf_smem<float2>();
    float2* tw = buf + SEQ * F::template pitch<4>();
    const int tid = threadIdx.x, s = tid / TPS, t = tid - s * TPS;
    const int p = blockIdx.y;
    const int row = blockIdx.x * SEQ + s;
    for (int i = tid; i < K; i += SEQ * TPS) tw[i] = __ldg(tw_g + i);
    const bool live = row < K;
    const float2* src = gridc + (long)p * K * K;
    float2* dst = T + ((long)p * K + row) * N;
    float2 val[8];
    {
        const int* irow = nz_idx + (long)row * K + t;
        int ix[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) ix[r] = live ? __ldg(irow + r * TPS) : -1;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 x = __ldcs(src + (ix[r] >= 0 ? ix[r] : 0));
            val[r] = ix[r] >= 0 ? x : make_float2(0.f, 0.f);
        }
    }
    auto ld = [&](int r) { return val[r]; };
    auto st = [&](int j, int r, float2 v) {
        if (live) __stcs(dst + j + r * NS3, v);
    };
    ff_transform<K, 4, true, false, true, false, false, true>(buf + s * F::template pitch<4>(), tw, t, ld, st);
}

template <int K, int SEQ>
__global__ void __launch_bounds__(SEQ* FastFft<K>::TPS, K <= 640 ? 3 : 1)
    ff_cols_adj_kernel(const float2* __restrict__ T, float2* __restrict__ U, const float2* __restrict__ tw_g, NufftDims d) {
    using F = FastFft<K>;
    constexpr int N = K / 2, TPS = F::TPS, NS3 = F::NS3;
    float2* buf = pf_smem<float2>();
    float2* tw = buf + SEQ * F::template pitch<3>();
    const int tid = threadIdx.x, t = tid / SEQ, s = tid - t * SEQ;
    const int p = blockIdx.y;
    const int col = blockIdx.x * SEQ + s;
    for (int i = tid; i < K; i += SEQ * TPS) tw[i] = __ldg(tw_g + i);
    const bool live = col < N;
    const float2* src = T + ((long)p * K + t) * N + col;
    float2* dst = U + (long)p * N * N + col;
    auto ld = [&](int r) { return live ? __ldcs(src + r * (TPS * N)) : make_float2(0.f, 0.f); };
    auto st = [&](int j, int r, float2 v) {
        if (live) __stcs(dst + j * N + r * (NS3 * N), v);
    };
    ff_transform<K, 3, true, false, true>(buf + s * F::template pitch<3>(), tw, t, ld, st);
}

constexpr int FF_SEQ_ROWS = 4;
template <int K>
static constexpr int ff_seq_cols() { return K <= 1024 ? 8 : 4; }     // SEQ * K / 8 threads <= 1024
template <int K>
static constexpr size_t ff_smem_bytes(int seq) { return ((size_t)seq * FastFft<K>::template pitch<3>() + K) * sizeof(float2); }   // PS = 3 is the larger pitch

static bool ff_supported(const pdu_nufft_plan* p) {
    return p->k0 == p->k1 && p->k0 == 2 * p->n0 && p->k1 == 2 * p->n1 && fast_fft_size(p->k0) && p->d_w0 && p->d_w1;
}

template <int K>
static int ff_forward(pdu_nufft_plan* p, const float2* image, const float2* smaps, float2* T, float2* grid, int planes,
                      int coils, int smaps_batch, cudaStream_t st) {
    const NufftDims d = dims_of(p);
    constexpr int FF_SEQ_COLS = ff_seq_cols<K>();
    auto rows = ff_rows_fwd_kernel<K, FF_SEQ_ROWS>;
    auto cols = ff_cols_fwd_kernel<K, FF_SEQ_COLS>;
    PDU_CUDA((ensure_dyn_smem<ff_rows_fwd_kernel<K, FF_SEQ_ROWS>>((int)ff_smem_bytes<K>(FF_SEQ_ROWS))));
    PDU_CUDA((ensure_dyn_smem<ff_cols_fwd_kernel<K, FF_SEQ_COLS>>((int)ff_smem_bytes<K>(FF_SEQ_COLS))));
    rows<<<dim3((unsigned)cdiv(p->n0, FF_SEQ_ROWS), (unsigned)planes), FF_SEQ_ROWS * FastFft<K>::TPS, ff_smem_bytes<K>(FF_SEQ_ROWS), st>>>(
        image, smaps, T, p->d_s0, p->d_s1, p->d_w1, d, coils, smaps_batch);
    PDU_LAUNCHED();
    cols<<<dim3((unsigned)cdiv(p->k1, FF_SEQ_COLS), (unsigned)planes), FF_SEQ_COLS * FastFft<K>::TPS, ff_smem_bytes<K>(FF_SEQ_COLS), st>>>(
        T, grid, p->d_w0, d);
    PDU_LAUNCHED();
    return PDU_OK;
}

// nz_idx: the compact form of the gridded samples (CsrView), nullptr for dense K x K planes
template <int K>
static int ff_adjoint(pdu_nufft_plan* p, const float2* grid, float2* T, float2* U, int planes, cudaStream_t st,
                      const int* nz_idx = nullptr) {
    const NufftDims d = dims_of(p);
    constexpr int FF_SEQ_COLS = ff_seq_cols<K>();
    auto rows = ff_rows_adj_kernel<K, FF_SEQ_ROWS>;
    auto rows_c = ff_rows_adj_compact_kernel<K, FF_SEQ_ROWS>;
    auto cols = ff_cols_adj_kernel<K, FF_SEQ_COLS>;
    PDU_CUDA((ensure_dyn_smem<ff_rows_adj_kernel<K, FF_SEQ_ROWS>>((int)ff_smem_bytes<K>(FF_SEQ_ROWS))));
    PDU_CUDA((ensure_dyn_smem<ff_rows_adj_compact_kernel<K, FF_SEQ_ROWS>>((int)ff_smem_bytes<K>(FF_SEQ_ROWS))));
    PDU_CUDA((ensure_dyn_smem<ff_cols_adj_kernel<K, FF_SEQ_COLS>>((int)ff_smem_bytes<K>(FF_SEQ_COLS))));
    const dim3 gr((unsigned)cdiv(p->k0, FF_SEQ_ROWS), (unsigned)planes);
    if (nz_idx) rows_c<<<gr, FF_SEQ_ROWS * FastFft<K>::TPS, ff_smem_bytes<K>(FF_SEQ_ROWS), st>>>(grid, T, p->d_w1, nz_idx, d);
    else rows<<<gr, FF_SEQ_ROWS * FastFft<K>::TPS, ff_smem_bytes<K>(FF_SEQ_ROWS), st>>>(grid, T, p->d_w1, d);
    PDU_LAUNCHED();
    cols<<<dim3((unsigned)cdiv(p->n1, FF_SEQ_COLS), (unsigned)planes), FF_SEQ_COLS * FastFft<K>::TPS, ff_smem_bytes<K>(FF_SEQ_COLS), st>>>(
        T, U, p->d_w0, d);
    PDU_LAUNCHED();
    return PDU_OK;
}

constexpr int PF_SEQ_ROWS = 4, PF_SEQ_COLS = 8;

static size_t pf_smem_bytes(int seq, int K) { return ((size_t)2 * seq * pfft_pitch(K) + K) * sizeof(float2); }

template <typename Kern>
static int pf_set_smem(Kern kern, size_t bytes) {
    PDU_REQUIRE(bytes <= 200 * 1024, "pruned FFT: %zu bytes of shared memory needed (grid too large)", bytes);
    PDU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return PDU_OK;
}

// ------------------------------------------------------------------ adjoint interpolation as a sorted gather
// The atomic scatter above sits on the L2 atomic unit's rate (measured 0.23 T float2 atomics/s on B200
// whatever the occupancy), and its summation order changes from run to run.  For a trajectory that is
// used more than once -- every unrolled iteration, every DCF iteration, every training step -- the
// adjoint interpolator is built once as a sparse matrix in CSR form (one row per grid cell: the samples
// that touch it and their conjugated Kaiser-Bessel weights, phase included), sorted by cell with a
// stable radix sort, and applied as a gather: no atomics, no memset, bit-reproducible.
//   csr buffer:  row_ptr int32[cells + 1] | samp int32[n] | w float2[n] | build scratch      (n = M J^2)
// Rows with more entries than this are summed by a whole warp (lanes stride the entries: coalesced index and weight
// loads).  A bound that follows the trajectory's average, 4 x entries per cell, so that dense trajectories keep their
// cells on the lane-group loop, was measured and lost (128^2 x 256 spokes, 8 planes: 218 against 159 us): the warp form
// is the better one for every row of 32 entries or more; what dense trajectories need is enough warp-per-row CTAs.
static int csr_long_threshold(const pdu_nufft_plan*, long) { return 32; }

struct CsrView {
    int* row_ptr;
    int* n_long;       // [1] number of long rows
    int* long_rows;    // [cells] their indices in nz_cell (first n_long valid)
    int* nz_cell;      // [cells] the non-empty cells in increasing order (first nz_ptr[k0] valid)
    int* nz_ptr;       // [k0 + 1] grid row r owns nz_cell[nz_ptr[r] .. nz_ptr[r + 1])
    int* nz_idx;       // [cells] compact index of a non-empty cell, -1 for an empty one
    int* nz_rptr;      // [cells + 1] row_ptr of the non-empty cells: entries of compact cell i are [nz_rptr[i], nz_rptr[i + 1])
    int* samp;
    float2* w;
    // build scratch
    unsigned* key_in;
    unsigned* id_in;
    unsigned* key_out;
    unsigned* id_out;
    float2* w_unsorted;
    int* rank;         // [cells + 1] number of non-empty cells before cell c
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static CsrView csr_layout(const pdu_nufft_plan* p, long M, void* base, bool with_scratch = true) {
    const size_t n = (size_t)M * p->J * p->J, cells = (size_t)p->k0 * p->k1;
    CsrView v;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return (char*)base + o; };
    v.row_ptr = (int*)take((cells + 1) * 4);
    v.n_long = (int*)take(4);
    v.long_rows = (int*)take(cells * 4);
    v.nz_cell = (int*)take(cells * 4);
    v.nz_ptr = (int*)take(((size_t)p->k0 + 1) * 4);
    v.nz_idx = (int*)take(cells * 4);
    v.nz_rptr = (int*)take((cells + 1) * 4);
    v.samp = (int*)take(n * 4);
    v.w = (float2*)take(n * 8);
    v.key_in = (unsigned*)take(n * 4);
    v.id_in = (unsigned*)take(n * 4);
    v.key_out = (unsigned*)take(n * 4);
    v.id_out = (unsigned*)take(n * 4);
    v.w_unsorted = (float2*)take(n * 8);
    v.rank = (int*)take((cells + 1) * 4);
    v.cub_bytes = 0;
    v.cub_tmp = nullptr;
    v.total = off;
    if (!with_scratch) return v;      // applying the matrix only needs the arrays before key_in
    cub::DeviceRadixSort::SortPairs(nullptr, v.cub_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr,
                                    (unsigned*)nullptr, (int)n);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int*)nullptr, (int*)nullptr, (int)(cells + 1));
    if (scan_bytes > v.cub_bytes) v.cub_bytes = scan_bytes;
    v.cub_tmp = take(v.cub_bytes);
    v.total = off;
    return v;
}

// one thread per sample: its J^2 (cell, weight) entries, weight = conj(phase * c0 * c1)
__global__ void __launch_bounds__(128)
    csr_entries_kernel(const float* __restrict__ omega, const float2* __restrict__ t0, const float2* __restrict__ t1,
                       NufftDims d, long M, unsigned* __restrict__ key, unsigned* __restrict__ id, float2* __restrict__ w) {
    const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float om0 = __ldg(omega + m), om1 = __ldg(omega + M + m);
    int g0[MAXJ], g1[MAXJ];
    float2 c0[MAXJ], c1[MAXJ];
    axis_taps(om0, d.gam0, d.k0, d.J, d.L, t0, g0, c0);
    axis_taps(om1, d.gam1, d.k1, d.J, d.L, t1, g1, c1);
    const float2 ph = shift_phase(om0, om1, d.shift0, d.shift1);
    const float2 one = make_float2(1.f, 0.f);
    const float2 cph = cmul_conj(one, ph);                       // conj(phase)
    long e = m * d.J * d.J;
#pragma unroll
    for (int a = 0; a < MAXJ; ++a) {
        if (a < d.J) {
            const float2 wa = cmul_conj(cph, c0[a]);            // the same association as the scatter: ((y conj ph) conj c0) conj c1
#pragma unroll
            for (int b = 0; b < MAXJ; ++b) {
                if (b < d.J) {
                    key[e] = (unsigned)(g0[a] * d.k1 + g1[b]);
                    id[e] = (unsigned)e;
                    w[e] = cmul_conj(wa, c1[b]);
                    ++e;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
    csr_finish_kernel(const unsigned* __restrict__ key_sorted, const unsigned* __restrict__ id_sorted,
                      const float2* __restrict__ w_unsorted, int* __restrict__ row_ptr, int* __restrict__ samp,
                      float2* __restrict__ w, long n, long cells, int taps) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned e = id_sorted[i];
        samp[i] = (int)(e / (unsigned)taps);
        w[i] = w_unsorted[e];
    }
    if (i <= cells) {          // row_ptr[c] = first sorted position whose key is >= c
        long lo = 0, hi = n;
        while (lo < hi) {
            const long mid = (lo + hi) >> 1;
            if ((long)key_sorted[mid] < i) lo = mid + 1; else hi = mid;
        }
        row_ptr[i] = (int)lo;
    }
}

// A sparse trajectory leaves most of the oversampled grid empty (48 radial spokes on 640^2: 61 % of the cells), so the
// NUFFT adjoint keeps the gridded samples COMPACT between the gather and the row pass: one value per non-empty cell, in
// cell order.  flag[c] = 1 for a non-empty cell (c < cells), 0 for c == cells; its exclusive sum is the compact index.
__global__ void __launch_bounds__(256) csr_flag_kernel(const int* __restrict__ row_ptr, int* __restrict__ flag, long cells) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= cells) flag[c] = (c < cells && row_ptr[c + 1] > row_ptr[c]) ? 1 : 0;
}
// the compact list itself, the first compact index of every grid row, and the rows too long for one thread (the k-space
// centre of a radial trajectory collects every spoke) by their compact index
__global__ void __launch_bounds__(256)
    csr_compact_kernel(const int* __restrict__ row_ptr, const int* __restrict__ rank, int* __restrict__ nz_cell,
                       int* __restrict__ nz_ptr, int* __restrict__ nz_idx, int* __restrict__ nz_rptr, int* __restrict__ n_long,
                       int* __restrict__ long_rows, long cells, int k0, int k1, int long_thresh) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > cells) return;
    if (c % k1 == 0) nz_ptr[c / k1] = rank[c];          // c == cells: nz_ptr[k0] = the number of non-empty cells
    if (c == cells) {
        nz_rptr[rank[c]] = row_ptr[c];                  // closes the last non-empty cell
        return;
    }
    const int len = row_ptr[c + 1] - row_ptr[c];
    if (len > 0) {
        nz_cell[rank[c]] = (int)c;
        nz_rptr[rank[c]] = row_ptr[c];
    }
    nz_idx[c] = len > 0 ? rank[c] : -1;
    if (len > long_thresh) long_rows[atomicAdd(n_long, 1)] = rank[c];
}

// one thread per grid cell, PG planes per thread (an entry is loaded once for all of them)
template <int PG>
__global__ void __launch_bounds__(256)
    interp_adj_csr_kernel(const float2* __restrict__ kdata, float2* __restrict__ grid, const int* __restrict__ row_ptr,
                          const int* __restrict__ samp, const float2* __restrict__ w, long cells, long M, int planes,
                          int long_thresh) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int p0 = blockIdx.y * PG;
    float2 acc[PG];
#pragma unroll
    for (int q = 0; q < PG; ++q) acc[q] = make_float2(0.f, 0.f);
    const int beg = __ldg(row_ptr + c), end = __ldg(row_ptr + c + 1);
    if (end - beg > long_thresh) return;                  // interp_adj_csr_long_kernel owns this cell
    for (int i = beg; i < end; ++i) {
        const long m = __ldg(samp + i);
        const float2 wi = __ldg(w + i);
#pragma unroll
        for (int q = 0; q < PG; ++q) {
            if (p0 + q < planes) {
                const float2 z = cmul(__ldg(kdata + (long)(p0 + q) * M + m), wi);
                acc[q].x += z.x;
                acc[q].y += z.y;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < PG; ++q)
        if (p0 + q < planes) grid[(long)(p0 + q) * cells + c] = acc[q];
}

// kdata [planes][M] -> kT [M][planes4] (planes4 = planes rounded up to 4, padding zeroed): an entry of the sparse
// matrix then finds the samples of PG neighbouring planes in PG * 8 contiguous bytes (one or two 16-byte loads
// from one sector) instead of PG sectors M * 8 bytes apart
__global__ void __launch_bounds__(256)
    transpose_kdata_kernel(const float2* __restrict__ kdata, float2* __restrict__ kT, long M, int planes, int planes4) {
    __shared__ float2 t[32][33];
    const long m0 = (long)blockIdx.x * 32;
    const int p0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int pl = p0 + r;
        const long m = m0 + threadIdx.x;
        t[r][threadIdx.x] = (pl < planes && m < M) ? __ldg(kdata + (long)pl * M + m) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const long m = m0 + r;
        const int pl = p0 + threadIdx.x;
        if (m < M && pl < planes4) kT[m * planes4 + pl] = t[threadIdx.x][r];
    }
}

// Sorted gather on the plane-interleaved samples.  LPC neighbouring lanes share a cell; a lane owns two planes in each of
// G plane groups 2 LPC apart, i.e. one 16-byte load per entry and group, and the LPC lanes of a cell read LPC * 16
// contiguous bytes: a warp instruction touches 32 / LPC cache lines, only 32 / LPC cells share a warp's trip count, per
// plane a warp writes 32 / LPC neighbouring cells, and an entry's index and weight are loaded once for 2 LPC G planes
// with G independent sample loads in flight.  r01's form -- one thread = one cell x 8 planes = four 16-byte loads, every
// one of them from 32 different lines (ncu: L1 74 % busy at 18 % of the DRAM rate, 40 % warps active) -- against
// (LPC, G), whole adjoint at 64 planes of 640^2, one box: (4, 1) 280, (4, 2) 253, (4, 4) 259, (4, 8) 290, (8, 1) 284,
// (8, 2) 259, (8, 4) 263, (2, 4) 266, (2, 8) 296 us, r01's form with the long rows as a kernel of their own ~ 300;
// two or four neighbouring cells per thread with 16-byte stores lost to the merged loop's divergence (310 / 361 us).
// Entries are summed in the same order per (cell, plane) whatever the shape: bit-identical results.
// The first long_blocks CTAs of every plane group take the long rows (the k-space centre: more than csr_long_threshold entries),
// one warp per row, lanes striding the entries and a fixed-order shuffle tree adding them up, exactly as
// interp_adj_csr_long_kernel does from the planar samples -- inside this launch, and scheduled first, the serial walk
// over a centre cell's ~1700 entries (18 us as a kernel of its own) is hidden behind the short cells.
// COMPACT: one value per NON-EMPTY cell (nz_cell, csr_compact_kernel) at grid[plane * cells + compact index] -- the row
// pass (ff_rows_adj_compact_kernel) spreads them over its shared-memory row; otherwise the dense K x K planes.
template <int LPC, int G, bool COMPACT>
__global__ void __launch_bounds__(256)
    interp_adj_csrT_kernel(const float2* __restrict__ kT, float2* __restrict__ grid, const int* __restrict__ row_ptr,
                           const int* __restrict__ samp, const float2* __restrict__ w, const int* __restrict__ n_long,
                           const int* __restrict__ long_rows, const int* __restrict__ nz_cell, const int* __restrict__ nz_rptr,
                           const int* __restrict__ n_nz, long cells, int planes, int planes4, int long_blocks, int long_thresh) {
    constexpr int CPB = 256 / LPC;
    if ((int)blockIdx.x < long_blocks) {
        constexpr int PG = 2 * LPC * G;                   // planes of this plane group
        const int lane = threadIdx.x & 31, p0 = blockIdx.y * PG;
        const int nl = __ldg(n_long);
        for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < nl; r += long_blocks * 8) {
            const int ci = __ldg(long_rows + r);
            const long slot = COMPACT ? (long)ci : (long)__ldg(nz_cell + ci);     // where the cell's value goes inside a plane
            const int beg = __ldg(nz_rptr + ci), end = __ldg(nz_rptr + ci + 1);
            // four planes at a time: eight accumulators, so that this rare path does not set the kernel's register count
#pragma unroll 1
            for (int h = 0; h < PG && p0 + h < planes4; h += 4) {
                float2 acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
                for (int i = beg + lane; i < end; i += 32) {
                    const long m = __ldg(samp + i);
                    const float2 wi = __ldg(w + i);
                    const float4* src = reinterpret_cast<const float4*>(kT + m * planes4 + p0 + h);
#pragma unroll
                    for (int q = 0; q < 4; q += 2) {
                        const float4 v = __ldg(src + q / 2);
                        const float2 z0 = cmul(make_float2(v.x, v.y), wi), z1 = cmul(make_float2(v.z, v.w), wi);
                        acc[q].x += z0.x;
                        acc[q].y += z0.y;
                        acc[q + 1].x += z1.x;
                        acc[q + 1].y += z1.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
                        acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
                    }
                    if (lane == 0 && p0 + h + q < planes) grid[(long)(p0 + h + q) * cells + slot] = acc[q];
                }
            }
        }
        return;
    }
    const int sub = threadIdx.x % LPC;
    const long o = (long)(blockIdx.x - long_blocks) * CPB + threadIdx.x / LPC;
    const int p = blockIdx.y * (2 * LPC * G) + 2 * sub;   // first of this lane's planes; the others follow 2 LPC apart
    if (o >= (COMPACT ? (long)__ldg(n_nz) : cells) || p >= planes4) return;
    // compact: the row pointer of the non-empty cells (one load level less than nz_cell -> row_ptr)
    const int* rp = COMPACT ? nz_rptr + o : row_ptr + o;
    const int beg = __ldg(rp), end = __ldg(rp + 1);
    if (end - beg > long_thresh) return;                  // the long-row CTAs own this cell
    float2 a0[G], a1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) a0[g] = a1[g] = make_float2(0.f, 0.f);
    const float2* src = kT + p;
    for (int i = beg; i < end; ++i) {
        const long m = __ldg(samp + i);
        const float2 wi = __ldg(w + i);
        const float4* sp = reinterpret_cast<const float4*>(src + m * planes4);
        float4 v[G];
#pragma unroll
        for (int g = 0; g < G; ++g)
            v[g] = p + g * 2 * LPC < planes4 ? __ldg(sp + g * LPC) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float2 z0 = cmul(make_float2(v[g].x, v[g].y), wi), z1 = cmul(make_float2(v[g].z, v[g].w), wi);
            a0[g].x += z0.x;
            a0[g].y += z0.y;
            a1[g].x += z1.x;
            a1[g].y += z1.y;
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const int pg = p + g * 2 * LPC;
        if (pg < planes) grid[(long)pg * cells + o] = a0[g];
        if (pg + 1 < planes) grid[(long)(pg + 1) * cells + o] = a1[g];
    }
}

// one warp per long row; lanes stride the entries, a fixed-order shuffle tree adds them up (reproducible)
template <int PG>
__global__ void __launch_bounds__(256)
    interp_adj_csr_long_kernel(const float2* __restrict__ kdata, float2* __restrict__ grid, const int* __restrict__ row_ptr,
                               const int* __restrict__ n_long, const int* __restrict__ long_rows, const int* __restrict__ nz_cell,
                               const int* __restrict__ samp, const float2* __restrict__ w, long cells, long M, int planes) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int p0 = blockIdx.y * PG;
    const int nl = __ldg(n_long);
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nl; r += warps) {
        const long c = __ldg(nz_cell + __ldg(long_rows + r));
        const int beg = __ldg(row_ptr + c), end = __ldg(row_ptr + c + 1);
        float2 acc[PG];
#pragma unroll
        for (int q = 0; q < PG; ++q) acc[q] = make_float2(0.f, 0.f);
        for (int i = beg + lane; i < end; i += 32) {
            const long m = __ldg(samp + i);
            const float2 wi = __ldg(w + i);
#pragma unroll
            for (int q = 0; q < PG; ++q) {
                if (p0 + q < planes) {
                    const float2 z = cmul(__ldg(kdata + (long)(p0 + q) * M + m), wi);
                    acc[q].x += z.x;
                    acc[q].y += z.y;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PG; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
                acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
            }
            if (lane == 0 && p0 + q < planes) grid[(long)(p0 + q) * cells + c] = acc[q];
        }
    }
}

static int csr_build(pdu_nufft_plan* p, const float* omega, long M, void* buf, size_t bytes, cudaStream_t st) {
    const long n = M * p->J * p->J, cells = (long)p->k0 * p->k1;
    PDU_REQUIRE(n < 2147483647L && cells < 2147483647L, "pdu_nufft_csr_build: %ld entries exceed 32-bit indexing", n);
    CsrView v = csr_layout(p, M, buf);
    if (!buf || bytes < v.total || ((uintptr_t)buf & 255)) {
        set_error("pdu_nufft_csr_build: buffer of %zu bytes (256-byte aligned) required, got %zu", v.total, buf ? bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    csr_entries_kernel<<<(unsigned)cdiv(M, 128), 128, 0, st>>>(omega, p->d_t0, p->d_t1, dims_of(p), M, v.key_in, v.id_in,
                                                              v.w_unsorted);
    PDU_LAUNCHED();
    int bits = 1;
    while ((1L << bits) < cells) ++bits;
    PDU_CUDA(cub::DeviceRadixSort::SortPairs(v.cub_tmp, v.cub_bytes, v.key_in, v.key_out, v.id_in, v.id_out, (int)n, 0, bits, st));
    count_launch(4);
    const long work = n > cells + 1 ? n : cells + 1;
    csr_finish_kernel<<<(unsigned)cdiv(work, 256), 256, 0, st>>>(v.key_out, v.id_out, v.w_unsorted, v.row_ptr, v.samp, v.w, n,
                                                                 cells, p->J * p->J);
    PDU_LAUNCHED();
    PDU_CUDA(cudaMemsetAsync(v.n_long, 0, 4, st));
    int* flag = (int*)v.key_out;                        // the sorted keys have been consumed by csr_finish_kernel
    int* flag_buf = n >= cells + 1 ? flag : v.rank;     // (a trajectory with fewer entries than cells: scan in place)
    csr_flag_kernel<<<(unsigned)cdiv(cells + 1, 256), 256, 0, st>>>(v.row_ptr, flag_buf, cells);
    PDU_LAUNCHED();
    PDU_CUDA(cub::DeviceScan::ExclusiveSum(v.cub_tmp, v.cub_bytes, flag_buf, v.rank, (int)(cells + 1), st));
    count_launch(2);
    csr_compact_kernel<<<(unsigned)cdiv(cells + 1, 256), 256, 0, st>>>(v.row_ptr, v.rank, v.nz_cell, v.nz_ptr, v.nz_idx, v.nz_rptr, v.n_long, v.long_rows,
                                                                      cells, p->k0, p->k1, csr_long_threshold(p, M));
    PDU_LAUNCHED();
    return PDU_OK;
}

// kT_scratch (optional): room for M * round_up(planes, 4) float2 -- the plane-interleaved copy of kdata
// compact (with kT_scratch only): see interp_adj_csrT_kernel; *compact is cleared when the dense form was written
static int launch_interp_adj_csr(pdu_nufft_plan* p, const float2* kdata, float2* grid, const void* csr, int planes, long M,
                                 cudaStream_t st, float2* kT_scratch = nullptr, bool* compact = nullptr) {
    const CsrView v = csr_layout(p, M, const_cast<void*>(csr), false);
    const long cells = (long)p->k0 * p->k1;
    constexpr int PG = 4;
    dim3 g((unsigned)cdiv(cells, 256), (unsigned)cdiv(planes, PG));
    if (kT_scratch && planes >= 4) {
        const int planes4 = (planes + 3) & ~3;
        transpose_kdata_kernel<<<dim3((unsigned)cdiv(M, 32), (unsigned)cdiv(planes4, 32)), dim3(32, 8), 0, st>>>(
            kdata, kT_scratch, M, planes, planes4);
        PDU_LAUNCHED();
        const int long_blocks = sm_count();               // 8 warps each, grid-stride over the long rows
        const bool cpt = compact && *compact;
        // compact: the grid is sized for the non-empty cells the trajectory can have at most (min(cells, entries)); threads
        // past the actual count (nz_ptr[k0], on the device) leave at once
        const long slots = cpt ? std::min(cells, M * p->J * p->J) : cells;
#define PDU_CSRT(L, G_)                                                                                                     \
    do {                                                                                                                    \
        const int lt_ = csr_long_threshold(p, M);                                                                           \
        dim3 gt_((unsigned)(long_blocks + cdiv(slots, 256 / L)), (unsigned)cdiv(planes4, 2 * L * G_));                       \
        if (cpt)                                                                                                            \
            interp_adj_csrT_kernel<L, G_, true><<<gt_, 256, 0, st>>>(kT_scratch, grid, v.row_ptr, v.samp, v.w, v.n_long,      \
                                                                    v.long_rows, v.nz_cell, v.nz_rptr, v.nz_ptr + p->k0, cells, \
                                                                    planes, planes4, long_blocks, lt_);                     \
        else                                                                                                                \
            interp_adj_csrT_kernel<L, G_, false><<<gt_, 256, 0, st>>>(kT_scratch, grid, v.row_ptr, v.samp, v.w, v.n_long,     \
                                                                     v.long_rows, v.nz_cell, v.nz_rptr, v.nz_ptr + p->k0, cells,\
                                                                     planes, planes4, long_blocks, lt_);                    \
    } while (0)
        // 16 planes per plane group for sparse trajectories (few entries per cell: the index loads are a large part of the
        // loop) and from 64 planes on; 8 otherwise (measured, 16 / 32 planes: 320^2 x 48 spokes 91 / 151 against 97 / 165 us,
        // 512^2 x 96 226 / 425 against 253 / 485; but 320^2 x 160 spokes 228 / 325 against 184 / 312)
        const double per_cell = (double)M * p->J * p->J / (double)cells;
        // (re-measured in the compact form at 64 planes: (4, 1) 251, (4, 2) 228, (4, 4) 226, (8, 1) 249, (8, 2) 234 us)
        if (planes4 >= 16 && (per_cell <= 5.0 || planes4 >= 64)) PDU_CSRT(4, 2);
        else if (planes4 >= 8) PDU_CSRT(4, 1);
        else PDU_CSRT(2, 1);
#undef PDU_CSRT
        PDU_LAUNCHED();
        return PDU_OK;
    }
    if (compact) *compact = false;
    interp_adj_csr_kernel<PG><<<g, 256, 0, st>>>(kdata, grid, v.row_ptr, v.samp, v.w, cells, M, planes, csr_long_threshold(p, M));
    PDU_LAUNCHED();
    dim3 gl((unsigned)(2 * sm_count()), (unsigned)cdiv(planes, PG));      // 8 warps per CTA, grid-stride over the long rows
    interp_adj_csr_long_kernel<PG><<<gl, 256, 0, st>>>(kdata, grid, v.row_ptr, v.n_long, v.long_rows, v.nz_cell, v.samp, v.w,
                                                       cells, M, planes);
    PDU_LAUNCHED();
    return PDU_OK;
}

// [planes][2][n] float32 (real plane, imaginary plane) <-> [planes][n] complex64: the layout change between the CNN
// side of PD-UNet and the generic (complex64) operator path, one pass, optionally times a real weight per element index
// (density compensation).  The fused path reads and writes the split layout directly and does not need these.
__global__ void __launch_bounds__(256)
    layout_to_complex_kernel(const float* __restrict__ split, float2* __restrict__ out, const float* __restrict__ weight, long n,
                             long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long p = i / n, e = i - p * n;
        const float w = weight ? __ldg(weight + e) : 1.f;
        out[i] = make_float2(w * __ldg(split + (2 * p) * n + e), w * __ldg(split + (2 * p + 1) * n + e));
    }
}
__global__ void __launch_bounds__(256)
    layout_to_split_kernel(const float2* __restrict__ in, float* __restrict__ split, long n, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long p = i / n, e = i - p * n;
        const float2 v = __ldg(in + i);
        split[(2 * p) * n + e] = v.x;
        split[(2 * p + 1) * n + e] = v.y;
    }
}

// cropped planes U [planes][n0][n1] -> image (x apodisation x scale, x conj(smaps) summed over coils); nufft_fused.cu
int launch_crop_apod(pdu_nufft_plan* p, const float2* U, const float2* smaps, float2* image, int out_planes, int coils,
                     int smaps_batch, float scale, int split, cudaStream_t st) {
    NufftDims dc = dims_of(p);          // the cropped result is a dense [n0][n1] "grid"
    dc.k0 = dc.n0;
    dc.k1 = dc.n1;
    const long total = (long)out_planes * p->n0 * p->n1;
    crop_apod_kernel<<<stream_grid(total), 256, 0, st>>>(U, smaps, image, p->d_s0, p->d_s1, dc, coils, smaps_batch, scale, total,
                                                         split);
    PDU_LAUNCHED();
    return PDU_OK;
}

static int check_call(const pdu_nufft_plan* p, const void* a, const void* b, const void* omega, int batch, int coils,
                      int smaps_batch, const void* smaps, long m, const char* who) {
    PDU_REQUIRE(p != nullptr, "%s: plan is null", who);
    PDU_REQUIRE(a && b && omega, "%s: null pointer", who);
    PDU_REQUIRE(batch > 0 && coils > 0 && m > 0, "%s: batch, coils and m must be > 0", who);
    PDU_REQUIRE((long)batch * coils <= 65535L * 64, "%s: too many planes", who);
    if (smaps) PDU_REQUIRE(smaps_batch == 1 || smaps_batch == batch, "%s: smaps batch must be 1 or batch", who);
    int dev = -1;
    PDU_CUDA(cudaGetDevice(&dev));
    PDU_REQUIRE(dev == p->device, "%s: plan was created on device %d, current device is %d", who, p->device, dev);
    return PDU_OK;
}

}  // namespace pdu

using namespace pdu;

extern "C" {

int pdu_nufft_plan_create(pdu_nufft_plan_t** plan, int n0, int n1, int k0, int k1, int numpoints, int table_oversamp,
                          int shift0, int shift1, const float* table0, const float* table1, const float* scal0,
                          const float* scal1) {
    PDU_REQUIRE(plan != nullptr, "pdu_nufft_plan_create: plan is null");
    *plan = nullptr;
    PDU_REQUIRE(n0 > 0 && n1 > 0 && k0 >= n0 && k1 >= n1, "pdu_nufft_plan_create: need 0 < n <= k per axis");
    PDU_REQUIRE(k1 % 2 == 0, "pdu_nufft_plan_create: the last grid dimension must be even (got %d)", k1);
    PDU_REQUIRE(numpoints >= 1 && numpoints <= MAXJ, "pdu_nufft_plan_create: numpoints must be in 1..%d", MAXJ);
    PDU_REQUIRE(table_oversamp >= 1 && (numpoints * table_oversamp) % 2 == 0,
                "pdu_nufft_plan_create: numpoints * table_oversamp must be even");
    PDU_REQUIRE(table0 && table1 && scal0 && scal1, "pdu_nufft_plan_create: null table");
    pdu_nufft_plan* p = new (std::nothrow) pdu_nufft_plan();
    if (!p) {
        set_error("pdu_nufft_plan_create: out of host memory");
        return PDU_ENOMEM;
    }
    p->n0 = n0; p->n1 = n1; p->k0 = k0; p->k1 = k1; p->J = numpoints; p->L = table_oversamp;
    p->shift0 = shift0; p->shift1 = shift1;
    p->d_t0 = p->d_t1 = nullptr;
    p->d_s0 = p->d_s1 = nullptr;
    p->d_w0 = p->d_w1 = nullptr;
    p->pfft_ok = pfft_factor(k0, &p->pf0) && pfft_factor(k1, &p->pf1) &&
                 pf_smem_bytes(PF_SEQ_ROWS, k1) <= 200 * 1024 && pf_smem_bytes(PF_SEQ_COLS, k0) <= 200 * 1024;
    const size_t tl = (size_t)numpoints * table_oversamp + 1;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_t0, tl * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_t1, tl * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_s0, (size_t)n0 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_s1, (size_t)n1 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_t0, table0, tl * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_t1, table1, tl * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_s0, scal0, (size_t)n0 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_s1, scal1, (size_t)n1 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && (p->pfft_ok || (fast_fft_size(k0) && fast_fft_size(k1)))) {
        for (int ax = 0; ax < 2 && e == cudaSuccess; ++ax) {
            const int K = ax == 0 ? k0 : k1;
            std::vector<float2> w((size_t)K);
            for (int i = 0; i < K; ++i) {
                const double a = -2.0 * 3.14159265358979323846 * (double)i / (double)K;
                w[i] = make_float2((float)cos(a), (float)sin(a));
            }
            float2** dst = ax == 0 ? &p->d_w0 : &p->d_w1;
            e = cudaMalloc(dst, (size_t)K * sizeof(float2));
            if (e == cudaSuccess) e = cudaMemcpy(*dst, w.data(), (size_t)K * sizeof(float2), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        set_error("pdu_nufft_plan_create: %s", cudaGetErrorString(e));
        pdu_nufft_plan_destroy(p);
        return PDU_ECUDA;
    }
    *plan = p;
    return PDU_OK;
}

int pdu_nufft_plan_destroy(pdu_nufft_plan_t* p) {
    if (!p) return PDU_OK;
    for (auto& kv : p->fft) cufftDestroy(kv.second);
    cudaFree(p->d_t0);
    cudaFree(p->d_t1);
    cudaFree(p->d_s0);
    cudaFree(p->d_s1);
    cudaFree(p->d_w0);
    cudaFree(p->d_w1);
    delete p;
    return PDU_OK;
}

size_t pdu_nufft_workspace_bytes(const pdu_nufft_plan_t* p, int planes) {
    if (!p || planes <= 0) return 0;
    // oversampled grid | intermediate of the pruned FFT (one axis transformed) | cropped adjoint result
    const size_t grid = (size_t)p->k0 * p->k1;
    const size_t mid = (size_t)std::max((long)p->n0 * p->k1, (long)p->k0 * p->n1);
    const size_t crop = (size_t)p->n0 * p->n1;
    return (size_t)planes * (grid + mid + crop) * sizeof(float2);
}

// A call is cut into batch chunks so that the scratch grids stay bounded (512 MB).  Measured on B200:
// chunking down to L2-resident grids (56 MB) does NOT pay -- cuFFT runs at ~3 TB/s either way and the
// gather / scatter lose parallelism -- so the budget only caps the workspace.
static int batch_chunk(const pdu_nufft_plan* p, int batch, int coils) {
    const size_t plane_bytes = (size_t)p->k0 * p->k1 * sizeof(float2);
    const size_t budget = (size_t)512 << 20;
    long planes = (long)(budget / plane_bytes);
    long cb = planes / coils;
    if (cb < 1) cb = 1;
    return (int)(cb < batch ? cb : batch);
}

static int nufft_fwd_chunk(pdu_nufft_plan_t* p, const float2* image, float2* kdata, const float* omega, const float2* smaps,
                           int batch, int coils, int smaps_batch, long m, float scale, float2* grid, cudaStream_t st) {
    const int planes = batch * coils;
    const long total = (long)planes * p->k0 * p->k1;
    // variant 1 = own pruned shared-memory FFT (pfft.cuh), variant 0 = pad + cuFFT.  Measured on B200 (64 planes of
    // 640^2): the pruned passes move 2.5x fewer bytes but are still ~1.7x slower than cuFFT's register-resident
    // radix kernels (row+column 407 us vs 238 us), so cuFFT stays the default until they are.
    int variant = option(OPT_NUFFT_FWD);
    // default: the register-resident pruned FFT on the BASELINE grids, the generic pruned FFT on every other grid
    // that factors into 2, 3, 5 (slower than cuFFT, but the library's own); cuFFT only for other prime factors
    if (variant < 0) variant = ff_supported(p) ? 2 : (p->pfft_ok ? 1 : 0);
    int rc;
    if (variant == 2 && !ff_supported(p)) variant = 0;
    if (variant == 2) {
        float2* T = grid + total;
        switch (p->k0) {
            case 256: rc = ff_forward<256>(p, image, smaps, T, grid, planes, coils, smaps_batch, st); break;
            case 512: rc = ff_forward<512>(p, image, smaps, T, grid, planes, coils, smaps_batch, st); break;
            case 640: rc = ff_forward<640>(p, image, smaps, T, grid, planes, coils, smaps_batch, st); break;
            case 1024: rc = ff_forward<1024>(p, image, smaps, T, grid, planes, coils, smaps_batch, st); break;
            default: rc = ff_forward<2048>(p, image, smaps, T, grid, planes, coils, smaps_batch, st); break;
        }
        if (rc) return rc;
    } else if (variant == 1 && p->pfft_ok) {
        // pruned FFT: apodise + pad + transform the n0 non-zero rows, then every column
        float2* T = grid + total;
        const NufftDims d = dims_of(p);
        const int pitch1 = pfft_pitch(p->k1), pitch0 = pfft_pitch(p->k0);
        rc = pf_set_smem(pfft_rows_fwd_kernel<PF_SEQ_ROWS>, pf_smem_bytes(PF_SEQ_ROWS, p->k1));
        if (rc) return rc;
        rc = pf_set_smem(pfft_cols_fwd_kernel<PF_SEQ_COLS>, pf_smem_bytes(PF_SEQ_COLS, p->k0));
        if (rc) return rc;
        pfft_rows_fwd_kernel<PF_SEQ_ROWS><<<dim3((unsigned)cdiv(p->n0, PF_SEQ_ROWS), (unsigned)planes), 256,
                                            pf_smem_bytes(PF_SEQ_ROWS, p->k1), st>>>(image, smaps, T, p->d_s0, p->d_s1, p->d_w1,
                                                                                     p->pf1, d, coils, smaps_batch, pitch1);
        PDU_LAUNCHED();
        pfft_cols_fwd_kernel<PF_SEQ_COLS><<<dim3((unsigned)cdiv(p->k1, PF_SEQ_COLS), (unsigned)planes), 256,
                                            pf_smem_bytes(PF_SEQ_COLS, p->k0), st>>>(T, grid, p->d_w0, p->pf0, d, pitch0);
        PDU_LAUNCHED();
    } else {
        apod_pad_kernel<<<stream_grid(total / 2), 256, 0, st>>>(image, smaps, (float4*)grid, p->d_s0, p->d_s1, dims_of(p), coils,
                                                                smaps_batch, total / 2);
        PDU_LAUNCHED();
        rc = run_fft(p, grid, planes, CUFFT_FORWARD, st);
        if (rc) return rc;
    }
    note_kernel(OP_NUFFT_FWD, "%s + interp_fwd_kernel<%d> (%d planes of %dx%d, M=%ld)",
                variant == 2 ? "ff_rows_fwd_kernel + ff_cols_fwd_kernel (register-resident pruned FFT)"
                             : (variant == 1 && p->pfft_ok ? "pfft_rows_fwd_kernel + pfft_cols_fwd_kernel (generic pruned FFT)"
                                                           : "apod_pad_kernel + cuFFT C2C"),
                p->J == 6 ? 6 : 0, planes, p->k0, p->k1, m);
    return launch_interp_fwd(p, grid, kdata, omega, planes, m, scale, st);
}

static int nufft_adj_chunk(pdu_nufft_plan_t* p, const float2* kdata, float2* image, const float* omega, const float2* smaps,
                           int batch, int coils, int smaps_batch, long m, float scale, float2* grid, const void* csr,
                           cudaStream_t st) {
    const int planes = batch * coils;
    int rc;
    int variant = option(OPT_NUFFT_ADJ);
    if (variant < 0) variant = ff_supported(p) ? 2 : (p->pfft_ok ? 1 : 0);      // see nufft_fwd_chunk
    if (variant == 2 && !ff_supported(p)) variant = 0;
    // the gridded samples stay compact (non-empty cells only) between the sorted gather and the register FFT's row pass
    // (only for sparse trajectories, at most 8 interpolation entries per cell on average: 48 radial spokes on 640^2 have
    //  2.7 and leave 61 % of the cells empty; a dense trajectory fills the grid and would only pay for the indirection)
    bool compact = csr && variant == 2 && (double)m * p->J * p->J <= 8.0 * (double)p->k0 * p->k1;
    if (csr) {
        // the FFT's intermediate buffer and the cropped result behind it are free until the transform starts: they hold the
        // plane-interleaved kdata
        float2* mid = grid + (long)planes * p->k0 * p->k1;
        const long mid_elems = (long)planes * (std::max((long)p->n0 * p->k1, (long)p->k0 * p->n1) + (long)p->n0 * p->n1);
        const bool fits = m * (long)((planes + 3) & ~3) <= mid_elems;
        rc = launch_interp_adj_csr(p, kdata, grid, csr, planes, m, st, fits ? mid : nullptr, &compact);   // writes every cell: no memset
    } else {
        PDU_CUDA(cudaMemsetAsync(grid, 0, (size_t)planes * p->k0 * p->k1 * sizeof(float2), st));
        rc = launch_interp_adj(p, kdata, grid, omega, planes, m, st);
    }
    if (rc) return rc;
    const int out_planes = smaps ? batch : planes;
    const long total = (long)out_planes * p->n0 * p->n1;
    note_kernel(OP_NUFFT_ADJ, "%s + %s + crop_apod_kernel (%d planes of %dx%d, M=%ld)",
                csr ? (compact ? "transpose_kdata_kernel + interp_adj_csrT_kernel (sorted gather: 4 lanes per cell x 16 planes, long rows by warps, non-empty cells only)"
                               : "transpose_kdata_kernel + interp_adj_csrT_kernel (sorted gather: 4 lanes per cell x 16 planes, long rows by warps)")
                    : "interp_adj_kernel (float2 atomics)",
                variant == 2 ? (compact ? "ff_rows_adj_compact_kernel + ff_cols_adj_kernel (register-resident pruned FFT)"
                                        : "ff_rows_adj_kernel + ff_cols_adj_kernel (register-resident pruned FFT)")
                             : (variant == 1 && p->pfft_ok ? "pfft_rows_adj_kernel + pfft_cols_adj_kernel (generic pruned FFT)"
                                                           : "cuFFT C2C"),
                planes, p->k0, p->k1, m);
    if (variant == 2) {
        float2* T = grid + (long)planes * p->k0 * p->k1;
        float2* U = T + (long)planes * std::max((long)p->n0 * p->k1, (long)p->k0 * p->n1);
        const CsrView cv = compact ? csr_layout(p, m, const_cast<void*>(csr), false) : CsrView{};
        const int* nzi = compact ? cv.nz_idx : nullptr;
        switch (p->k0) {
            case 256: rc = ff_adjoint<256>(p, grid, T, U, planes, st, nzi); break;
            case 512: rc = ff_adjoint<512>(p, grid, T, U, planes, st, nzi); break;
            case 640: rc = ff_adjoint<640>(p, grid, T, U, planes, st, nzi); break;
            case 1024: rc = ff_adjoint<1024>(p, grid, T, U, planes, st, nzi); break;
            default: rc = ff_adjoint<2048>(p, grid, T, U, planes, st, nzi); break;
        }
        if (rc) return rc;
        NufftDims dc = dims_of(p);          // the cropped result is a dense [n0][n1] "grid"
        dc.k0 = dc.n0;
        dc.k1 = dc.n1;
        crop_apod_kernel<<<stream_grid(total), 256, 0, st>>>(U, smaps, image, p->d_s0, p->d_s1, dc, coils, smaps_batch, scale,
                                                             total);
        PDU_LAUNCHED();
        return PDU_OK;
    }
    if (variant == 1 && p->pfft_ok) {
        // pruned inverse FFT: every row but only the n1 kept outputs, then the n1 kept columns and n0 kept outputs
        float2* T = grid + (long)planes * p->k0 * p->k1;
        float2* U = T + (long)planes * std::max((long)p->n0 * p->k1, (long)p->k0 * p->n1);
        const NufftDims d = dims_of(p);
        rc = pf_set_smem(pfft_rows_adj_kernel<PF_SEQ_ROWS>, pf_smem_bytes(PF_SEQ_ROWS, p->k1));
        if (rc) return rc;
        rc = pf_set_smem(pfft_cols_adj_kernel<PF_SEQ_COLS>, pf_smem_bytes(PF_SEQ_COLS, p->k0));
        if (rc) return rc;
        pfft_rows_adj_kernel<PF_SEQ_ROWS><<<dim3((unsigned)cdiv(p->k0, PF_SEQ_ROWS), (unsigned)planes), 256,
                                            pf_smem_bytes(PF_SEQ_ROWS, p->k1), st>>>(grid, T, p->d_w1, p->pf1, d,
                                                                                     pfft_pitch(p->k1));
        PDU_LAUNCHED();
        pfft_cols_adj_kernel<PF_SEQ_COLS><<<dim3((unsigned)cdiv(p->n1, PF_SEQ_COLS), (unsigned)planes), 256,
                                            pf_smem_bytes(PF_SEQ_COLS, p->k0), st>>>(T, U, p->d_w0, p->pf0, d, pfft_pitch(p->k0));
        PDU_LAUNCHED();
        NufftDims dc = d;          // the cropped result is a dense [n0][n1] "grid"
        dc.k0 = d.n0;
        dc.k1 = d.n1;
        crop_apod_kernel<<<stream_grid(total), 256, 0, st>>>(U, smaps, image, p->d_s0, p->d_s1, dc, coils, smaps_batch, scale,
                                                             total);
        PDU_LAUNCHED();
        return PDU_OK;
    }
    rc = run_fft(p, grid, planes, CUFFT_INVERSE, st);
    if (rc) return rc;
    crop_apod_kernel<<<stream_grid(total), 256, 0, st>>>(grid, smaps, image, p->d_s0, p->d_s1, dims_of(p), coils, smaps_batch,
                                                         scale, total);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_nufft_fwd_c64(pdu_nufft_plan_t* p, const float* image, float* kdata, const float* omega, const float* smaps,
                      int batch, int coils, int smaps_batch, long m, float scale, void* workspace, size_t workspace_bytes,
                      pdu_stream_t stream) {
    int rc = check_call(p, image, kdata, omega, batch, coils, smaps_batch, smaps, m, "pdu_nufft_fwd_c64");
    if (rc) return rc;
    PDU_CHECK_DEVICE("pdu_nufft_fwd_c64");
    const int cb = batch_chunk(p, batch, coils);
    const size_t need = pdu_nufft_workspace_bytes(p, cb * coils);
    if (!workspace || workspace_bytes < need) {
        set_error("pdu_nufft_fwd_c64: workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    const long plane = (long)p->n0 * p->n1;
    const long img_b = (smaps ? 1 : coils) * plane, smap_b = smaps_batch == 1 ? 0 : coils * plane;
    for (int b0 = 0; b0 < batch; b0 += cb) {
        const int nb = b0 + cb <= batch ? cb : batch - b0;
        rc = nufft_fwd_chunk(p, (const float2*)image + b0 * img_b, (float2*)kdata + (long)b0 * coils * m, omega,
                             smaps ? (const float2*)smaps + b0 * smap_b : nullptr, nb, coils, smaps_batch == 1 ? 1 : nb, m, scale,
                             (float2*)workspace, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return PDU_OK;
}

size_t pdu_nufft_csr_bytes(const pdu_nufft_plan_t* p, long m) {
    if (!p || m <= 0) return 0;
    return csr_layout(p, m, nullptr).total;
}

size_t pdu_nufft_csr_bytes2(const pdu_nufft_plan_t* p, long m, size_t* persist_bytes) {
    if (persist_bytes) *persist_bytes = 0;
    if (!p || m <= 0) return 0;
    const CsrView v = csr_layout(p, m, nullptr);
    // row_ptr | n_long | long_rows | samp | w are what applying the matrix reads; everything behind is sort scratch
    if (persist_bytes) *persist_bytes = (size_t)((char*)v.key_in - (char*)nullptr);
    return v.total;
}

int pdu_nufft_csr_build(pdu_nufft_plan_t* p, const float* omega, long m, void* csr, size_t csr_bytes, pdu_stream_t stream) {
    PDU_REQUIRE(p && omega && m > 0, "pdu_nufft_csr_build: null pointer or m <= 0");
    return csr_build(p, omega, m, csr, csr_bytes, (cudaStream_t)stream);
}

int pdu_nufft_interp_adj_csr_c64(pdu_nufft_plan_t* p, const float* kdata, float* grid, const void* csr, int planes, long m,
                                 pdu_stream_t stream) {
    PDU_REQUIRE(p && kdata && grid && csr && planes > 0 && m > 0, "pdu_nufft_interp_adj_csr_c64: null pointer or empty problem");
    return launch_interp_adj_csr(p, (const float2*)kdata, (float2*)grid, csr, planes, m, (cudaStream_t)stream);
}

int pdu_nufft_adj_csr_c64(pdu_nufft_plan_t* p, const float* kdata, float* image, const float* omega, const float* smaps,
                          int batch, int coils, int smaps_batch, long m, float scale, const void* csr, void* workspace,
                          size_t workspace_bytes, pdu_stream_t stream);

int pdu_nufft_adj_c64(pdu_nufft_plan_t* p, const float* kdata, float* image, const float* omega, const float* smaps,
                      int batch, int coils, int smaps_batch, long m, float scale, void* workspace, size_t workspace_bytes,
                      pdu_stream_t stream) {
    return pdu_nufft_adj_csr_c64(p, kdata, image, omega, smaps, batch, coils, smaps_batch, m, scale, nullptr, workspace,
                                 workspace_bytes, stream);
}

int pdu_nufft_adj_csr_c64(pdu_nufft_plan_t* p, const float* kdata, float* image, const float* omega, const float* smaps,
                          int batch, int coils, int smaps_batch, long m, float scale, const void* csr, void* workspace,
                          size_t workspace_bytes, pdu_stream_t stream) {
    int rc = check_call(p, kdata, image, omega, batch, coils, smaps_batch, smaps, m, "pdu_nufft_adj_c64");
    if (rc) return rc;
    PDU_CHECK_DEVICE("pdu_nufft_adj_c64");
    const int cb = batch_chunk(p, batch, coils);
    const size_t need = pdu_nufft_workspace_bytes(p, cb * coils);
    if (!workspace || workspace_bytes < need) {
        set_error("pdu_nufft_adj_c64: workspace of %zu bytes required, got %zu", need, workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    const long plane = (long)p->n0 * p->n1;
    const long img_b = (smaps ? 1 : coils) * plane, smap_b = smaps_batch == 1 ? 0 : coils * plane;
    for (int b0 = 0; b0 < batch; b0 += cb) {
        const int nb = b0 + cb <= batch ? cb : batch - b0;
        rc = nufft_adj_chunk(p, (const float2*)kdata + (long)b0 * coils * m, (float2*)image + b0 * img_b, omega,
                             smaps ? (const float2*)smaps + b0 * smap_b : nullptr, nb, coils, smaps_batch == 1 ? 1 : nb, m, scale,
                             (float2*)workspace, csr, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return PDU_OK;
}

int pdu_nufft_fwd_binned_c64(pdu_nufft_plan_t* p, const float* image, float* kdata, const float* smaps, int batch, int coils,
                             int smaps_batch, long m, float scale, const void* bins, int flags, void* workspace,
                             size_t workspace_bytes, pdu_stream_t stream) {
    int rc = check_call(p, image, kdata, bins, batch, coils, smaps_batch, smaps, m, "pdu_nufft_fwd_binned_c64");
    if (rc) return rc;
    PDU_CHECK_DEVICE("pdu_nufft_fwd_binned_c64");
    if (!fused_supported(p)) {
        set_error("pdu_nufft_fwd_binned_c64: grid %d x %d (J = %d) has no fused path; use pdu_nufft_fwd_c64", p->k0, p->k1, p->J);
        return PDU_EUNSUPPORTED;
    }
    const size_t need = fused_workspace_bytes(p, batch * coils, m);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
        set_error("pdu_nufft_fwd_binned_c64: workspace of %zu bytes (256-byte aligned) required, got %zu", need,
                  workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    return fused_forward(p, image, kdata, smaps, batch, coils, smaps_batch, m, scale, bins, flags, workspace, (cudaStream_t)stream);
}

int pdu_nufft_adj_binned_c64(pdu_nufft_plan_t* p, const float* kdata, float* image, const float* smaps, const float* kweight,
                             int batch, int coils, int smaps_batch, long m, float scale, const void* bins, int flags,
                             void* workspace, size_t workspace_bytes, pdu_stream_t stream) {
    int rc = check_call(p, kdata, image, bins, batch, coils, smaps_batch, smaps, m, "pdu_nufft_adj_binned_c64");
    if (rc) return rc;
    PDU_CHECK_DEVICE("pdu_nufft_adj_binned_c64");
    if (!fused_supported(p)) {
        set_error("pdu_nufft_adj_binned_c64: grid %d x %d (J = %d) has no fused path; use pdu_nufft_adj_c64", p->k0, p->k1, p->J);
        return PDU_EUNSUPPORTED;
    }
    const size_t need = fused_workspace_bytes(p, batch * coils, m);
    if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 255)) {
        set_error("pdu_nufft_adj_binned_c64: workspace of %zu bytes (256-byte aligned) required, got %zu", need,
                  workspace ? workspace_bytes : (size_t)0);
        return PDU_ENOMEM;
    }
    return fused_adjoint(p, kdata, image, smaps, kweight, batch, coils, smaps_batch, m, scale, bins, flags, workspace,
                         (cudaStream_t)stream);
}

int pdu_complex_from_split_f32(const float* split, float* out, const float* weight, long planes, long n, pdu_stream_t stream) {
    PDU_REQUIRE(split && out && planes > 0 && n > 0, "pdu_complex_from_split_f32: null pointer or empty tensor");
    layout_to_complex_kernel<<<stream_grid(planes * n), 256, 0, (cudaStream_t)stream>>>(split, (float2*)out, weight, n, planes * n);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_split_from_complex_f32(const float* in, float* split, long planes, long n, pdu_stream_t stream) {
    PDU_REQUIRE(split && in && planes > 0 && n > 0, "pdu_split_from_complex_f32: null pointer or empty tensor");
    layout_to_split_kernel<<<stream_grid(planes * n), 256, 0, (cudaStream_t)stream>>>((const float2*)in, split, n, planes * n);
    PDU_LAUNCHED();
    return PDU_OK;
}

int pdu_nufft_interp_fwd_c64(pdu_nufft_plan_t* p, const float* grid, float* kdata, const float* omega, int planes,
                             long m, pdu_stream_t stream) {
    int rc = check_call(p, grid, kdata, omega, planes, 1, 1, nullptr, m, "pdu_nufft_interp_fwd_c64");
    if (rc) return rc;
    return launch_interp_fwd(p, (const float2*)grid, (float2*)kdata, omega, planes, m, 1.f, (cudaStream_t)stream);
}

int pdu_nufft_interp_adj_c64(pdu_nufft_plan_t* p, const float* kdata, float* grid, const float* omega, int planes,
                             long m, pdu_stream_t stream) {
    int rc = check_call(p, kdata, grid, omega, planes, 1, 1, nullptr, m, "pdu_nufft_interp_adj_c64");
    if (rc) return rc;
    return launch_interp_adj(p, (const float2*)kdata, (float2*)grid, omega, planes, m, (cudaStream_t)stream);
}

}  // extern "C"
