// Sinogram filtering along the detector axis: out[r, i] = sum_j sino[r, j] * taps[(i - j) + D - 1].
// Replaces [RECALL] torch_radon Radon.filter_sinogram (zero-pad to a power of two, rfft, multiply by
// the ramp response, irfft, crop, scale): the padded circular product IS this linear convolution,
// so it is done directly as a Toeplitz contraction and the sinogram makes one round trip through HBM
// instead of the FFT route's three.
//
// variant 1 (default when det_count % 4 == 0 and >= 128; the contraction is zero-padded to the 128-wide tile) -- exact
//   split-TF32 GEMM on tcgen05 tensor cores: filter_tc.cu.
//   Measured on B200 (8192 x 256 rows): 20.5 us against 47.1 us for variant 0.
// variant 0 -- register-tiled FP32 contraction on the CUDA cores.  A CTA owns RB rows; the rows
//   and the 2D-1 taps sit in shared memory; a thread produces a 4 (rows) x 4 (adjacent outputs)
//   patch, sliding a 4-tap window so every inner step costs 4 broadcast row loads + 1 tap load for
//   16 FMAs.
#include "common.cuh"

namespace pdu {

// filter_tc.cu
bool filter_tc_supported(int D);
size_t filter_tc_workspace_bytes(int D);
int filter_tc_prepare(const float* taps, void* ws, size_t ws_bytes, int D, cudaStream_t st);
int filter_tc_launch(const float* sino, float* out, const void* ws, const float* col_weight, long rows, int D, cudaStream_t st);

constexpr int FILT_RB = 16;   // rows per CTA

// dynamic smem: taps[2D+6] (zero padded by 3 in front, 4 behind) | rows[RB][D]
__global__ void __launch_bounds__(256)
    filter_direct_kernel(const float* __restrict__ sino, float* __restrict__ out, const float* __restrict__ taps,
                         const float* __restrict__ col_weight, long rows, int D) {
    extern __shared__ float s_f[];
    const int TL = 2 * D - 1;
    float* s_taps = s_f;                 // s_taps[3 + k] = taps[k]
    float* s_rows = s_f + (TL + 8 + 3) / 4 * 4;
    const long r0 = (long)blockIdx.x * FILT_RB;
    const int nr = (int)min((long)FILT_RB, rows - r0);
    for (int i = threadIdx.x; i < TL + 7; i += blockDim.x) {
        const int k = i - 3;
        s_taps[i] = (k >= 0 && k < TL) ? __ldg(taps + k) : 0.f;
    }
    for (int i = threadIdx.x; i < FILT_RB * D; i += blockDim.x) {
        const int r = i / D;
        float v = r < nr ? __ldg(sino + r0 * D + i) : 0.f;
        if (col_weight) v *= __ldg(col_weight + (i - r * D));
        s_rows[i] = v;
    }
    __syncthreads();
    // patches: (FILT_RB / 4) row groups x ceil(D / 4) column groups
    const int cg = (D + 3) / 4;
    for (int p = threadIdx.x; p < (FILT_RB / 4) * cg; p += blockDim.x) {
        const int rg = p / cg, c = (p - rg * cg) * 4;
        const float* ra = s_rows + (rg * 4 + 0) * D;
        const float* rb = s_rows + (rg * 4 + 1) * D;
        const float* rc = s_rows + (rg * 4 + 2) * D;
        const float* rd = s_rows + (rg * 4 + 3) * D;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[a][q] = 0.f;
        // output i = c + q needs taps[(c + q - j) + D - 1]; window w[q] at step j
        const float* tp = s_taps + 3 + c + D - 1;     // tp[q - j]
        float w0 = tp[0], w1 = tp[1], w2 = tp[2], w3 = tp[3];
#pragma unroll 4
        for (int j = 0; j < D; ++j) {
            const float xa = ra[j], xb = rb[j], xc = rc[j], xd = rd[j];
            acc[0][0] = fmaf(xa, w0, acc[0][0]); acc[0][1] = fmaf(xa, w1, acc[0][1]);
            acc[0][2] = fmaf(xa, w2, acc[0][2]); acc[0][3] = fmaf(xa, w3, acc[0][3]);
            acc[1][0] = fmaf(xb, w0, acc[1][0]); acc[1][1] = fmaf(xb, w1, acc[1][1]);
            acc[1][2] = fmaf(xb, w2, acc[1][2]); acc[1][3] = fmaf(xb, w3, acc[1][3]);
            acc[2][0] = fmaf(xc, w0, acc[2][0]); acc[2][1] = fmaf(xc, w1, acc[2][1]);
            acc[2][2] = fmaf(xc, w2, acc[2][2]); acc[2][3] = fmaf(xc, w3, acc[2][3]);
            acc[3][0] = fmaf(xd, w0, acc[3][0]); acc[3][1] = fmaf(xd, w1, acc[3][1]);
            acc[3][2] = fmaf(xd, w2, acc[3][2]); acc[3][3] = fmaf(xd, w3, acc[3][3]);
            w3 = w2; w2 = w1; w1 = w0;
            w0 = tp[-(j + 1)];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = rg * 4 + a;
            if (r < nr) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (c + q < D) out[(r0 + r) * D + c + q] = acc[a][q];
            }
        }
    }
}

}  // namespace pdu

using namespace pdu;

extern "C" {

size_t pdu_filter_workspace_bytes(int det_count) { return det_count > 0 ? filter_tc_workspace_bytes(det_count) : 0; }

int pdu_filter_prepare_f32(const float* taps, void* workspace, size_t workspace_bytes, int det_count,
                           pdu_stream_t stream) {
    PDU_REQUIRE(det_count > 0, "pdu_filter_prepare_f32: det_count must be > 0");
    if (!filter_tc_supported(det_count)) return PDU_OK;
    PDU_REQUIRE(taps != nullptr, "pdu_filter_prepare_f32: taps is null");
    return filter_tc_prepare(taps, workspace, workspace_bytes, det_count, (cudaStream_t)stream);
}

int pdu_filter_sinogram_weighted_f32(const float* sino, float* out, const float* taps, const float* col_weight,
                                     const void* workspace, size_t workspace_bytes, long rows, int det_count,
                                     pdu_stream_t stream) {
    PDU_CHECK_DEVICE("pdu_filter_sinogram_f32");
    PDU_REQUIRE(sino && out && taps, "pdu_filter_sinogram_f32: null pointer");
    PDU_REQUIRE(rows > 0 && det_count > 0, "pdu_filter_sinogram_f32: rows and det_count must be > 0");
    PDU_REQUIRE(sino != out, "pdu_filter_sinogram_f32: in-place filtering is not supported");
    const int D = det_count;
    int variant = option(OPT_FILTER);
    if (variant < 0) variant = 1;      // 1 = tcgen05 split-TF32 GEMM, exact products (filter_tc.cu); 0 = CUDA cores
    if (variant >= 1 && filter_tc_supported(D) && workspace && workspace_bytes >= filter_tc_workspace_bytes(D) &&
        (((uintptr_t)sino | (uintptr_t)out | (uintptr_t)workspace | (uintptr_t)col_weight) & 15) == 0)
        return filter_tc_launch(sino, out, workspace, col_weight, rows, D, (cudaStream_t)stream);
    const size_t smem = ((size_t)(2 * D - 1 + 8 + 3) / 4 * 4 + (size_t)FILT_RB * D) * sizeof(float);
    PDU_REQUIRE(smem <= 200 * 1024, "pdu_filter_sinogram_f32: det_count %d too large for the shared-memory tile", D);
    PDU_CUDA(ensure_dyn_smem<filter_direct_kernel>(200 * 1024));
    const long blocks = cdiv(rows, FILT_RB);
    PDU_REQUIRE(blocks <= 2147483647L, "pdu_filter_sinogram_f32: too many rows");
    filter_direct_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(sino, out, taps, col_weight, rows, D);
    PDU_LAUNCHED();
    note_kernel(OP_FILTER, "filter_direct_kernel grid %ld (CUDA-core Toeplitz contraction: det_count %% 4 != 0 or < 128)%s", blocks,
                col_weight ? " + fused detector weight" : "");
    return PDU_OK;
}

int pdu_filter_sinogram_f32(const float* sino, float* out, const float* taps, const void* workspace,
                            size_t workspace_bytes, long rows, int det_count, pdu_stream_t stream) {
    return pdu_filter_sinogram_weighted_f32(sino, out, taps, nullptr, workspace, workspace_bytes, rows, det_count, stream);
}

}  // extern "C"
