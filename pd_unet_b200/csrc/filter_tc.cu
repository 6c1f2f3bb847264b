// Sinogram filter on the 5th-generation tensor cores: Out[R, D] = X[R, D] . H[D, D] with the
// Toeplitz matrix H[k][n] = taps[(n - k) + D - 1], as a split-TF32 GEMM accumulated in float32 in
// tensor memory.  Every operand is cut into pieces of 11 significant bits (what kind::tf32 keeps), so
// each tensor-core product of two pieces is EXACT and only the float32 accumulation rounds:
//     X = X1 + X2 + X3, H = H1 + H2 + H3,   Out = X1H1 + X1H2 + X2H1 + X2H2 + X1H3 + X3H1   (dropped terms < 2^-33)
// (The usual 3-product "3xTF32" form is ~4e-7 of sum|x||h| -- not enough for ramp-filtered object sinograms, whose
// output is ~50x smaller than sum|x||h|: measured 1.5e-5 rel-L2 at 512 bins against the 1e-5 budget.  It, the
// variants that synthesised the Toeplitz tiles in shared memory and the 4-stage ring of 16-wide K blocks were
// measured in r01 (16-25 us against 20.5) and deleted in r02: DESIGN.md 3.3.)
// The sinogram makes a single round trip through HBM.  This is the one step of the hot path that is a
// dense contraction.
//
// One CTA = one 128 (rows) x 128 (outputs) tile.  Six warps:
//   warp 0      TMA producer: per 32-wide K block loads the raw X tile and the pre-split H pieces
//               (cp.async.bulk.tensor, SWIZZLE_128B) into a 2-stage ring
//   warp 1      MMA issuer: one elected lane issues 24 tcgen05.mma.kind::tf32 (M128 N128 K8) per K
//               block, tcgen05.commit frees the stage and finally signals the epilogue
//   warps 2..5  splitters, then epilogue: cut the raw X tile into its pieces (X1 in place) --
//               elementwise, so independent of the swizzle -- fence to the async proxy, and at the end
//               read the accumulator with tcgen05.ld (each warp its own 32 TMEM lanes) and store it.
// H is split once by pdu_filter_prepare_f32 into the workspace ([3][D][D], K-major: B[n][k]).
// Fan-beam FBP: an optional per-detector weight (the cosine pre-weight) is applied to X by the splitter warps
// before the cut, so the weighting costs no extra pass over the sinogram.
// Every mbarrier wait is bounded; after a time-out the CTA stores nothing and reports through the device error word.
#include <cuda.h>

#include "common.cuh"

namespace pdu {

constexpr int TC_BM = 128, TC_BN = 128;
constexpr int TC_SPLITTERS = 256;                          // warps 2..9 cut X; warps 2..5 also run the epilogue
constexpr int TC_THREADS = 64 + TC_SPLITTERS;
constexpr int TC_BK = 32;                                  // K extent of one pipeline stage: 128-byte rows, SWIZZLE_128B
constexpr int TC_SPLIT = 3;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 4;           // every operand tile (BM == BN)
constexpr int TC_STAGE_BYTES = 2 * TC_SPLIT * TC_TILE_BYTES;   // X pieces (piece 0 = raw X, split in place), H pieces
constexpr int TC_STAGES = (192 * 1024) / TC_STAGE_BYTES;   // 2
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t tc_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded (common.cuh): false on time-out
__device__ __forceinline__ bool tc_mbar_wait(uint32_t bar, uint32_t parity, unsigned long long timeout_ns) {
    return mbar_wait_bounded(bar, parity, timeout_ns);
}
__device__ __forceinline__ void tc_tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile rows are 128 bytes, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)((8 * TC_BK * 4) >> 4) << 32;     // stride byte offset: 8 rows x 128 bytes
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // layout: SWIZZLE_128B
    return d;
}
// kind::tf32, float32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_head(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// col_weight (nullable): float[D] multiplied into X[:, j] before the split (fan-beam cosine pre-weight).
__global__ void __launch_bounds__(TC_THREADS, 1)
    filter_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                     const float* __restrict__ col_weight, float* __restrict__ out, long rows, int D,
                     int* __restrict__ err_word, unsigned long long timeout_ns, int fault) {
    constexpr int CHUNKS = TC_BK / 4;                // 16-byte chunks per tile row
    extern __shared__ unsigned char tc_dyn[];
    const uint32_t dyn = tc_s32(tc_dyn);
    const uint32_t base = (dyn + 1023u) & ~1023u;                  // SWIZZLE_128B tiles want 1024-byte alignment
    unsigned char* base_ptr = tc_dyn + (base - dyn);
    const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;       // full[], split[], empty[], accum, tmem slot
    auto full = [&](int s) { return bars + 8u * s; };
    auto split = [&](int s) { return bars + 8u * (TC_STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };
    const uint32_t accum = bars + 8u * (3 * TC_STAGES);
    const uint32_t tmem_slot = accum + 8u;
    volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(base_ptr + TC_STAGES * TC_STAGE_BYTES + 8 * (3 * TC_STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BN;
    const int Dp = (D + TC_BN - 1) / TC_BN * TC_BN;         // the contraction runs on D rounded up to the tile: TMA zero-fills
    const int n_kb = Dp / TC_BK;                              // the columns of X beyond D, H is built zero-padded

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            tc_mbar_init(full(s), 1);
            tc_mbar_init(split(s), TC_SPLITTERS);
            tc_mbar_init(empty(s), 1);
        }
        tc_mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: 128 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TC_BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot_ptr;

    // `ok` turns false at the first wait that times out; the role loops stop there (every later wait would time out
    // as well), nothing is stored, and the CTA still reaches the barrier below to release its tensor memory
    bool ok = true;
    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int kb = 0; ok && kb < n_kb; ++kb) {
                const int s = kb % TC_STAGES;
                if (kb >= TC_STAGES) ok = tc_mbar_wait(empty(s), ((kb / TC_STAGES) - 1) & 1, timeout_ns);
                if (!ok) break;
                const uint32_t st = base + s * TC_STAGE_BYTES;
                tc_mbar_expect_tx(full(s), (1 + TC_SPLIT) * TC_TILE_BYTES);
                if (fault) continue;                                                   // debug_fault: the loads never come
                tc_tma_2d(st, &tm_x, kb * TC_BK, m0, full(s));                         // raw X -> split in place into X1
#pragma unroll
                for (int p = 0; p < TC_SPLIT; ++p)                                     // pre-split H pieces, stacked by rows
                    tc_tma_2d(st + (TC_SPLIT + p) * TC_TILE_BYTES, &tm_h, kb * TC_BK, p * Dp + n0, full(s));
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            for (int kb = 0; ok && kb < n_kb; ++kb) {
                const int s = kb % TC_STAGES;
                ok = tc_mbar_wait(split(s), (kb / TC_STAGES) & 1, timeout_ns);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * TC_STAGE_BYTES;
                uint64_t xd[TC_SPLIT], hd[TC_SPLIT];
#pragma unroll
                for (int p = 0; p < TC_SPLIT; ++p) {
                    xd[p] = tc_smem_desc(st + p * TC_TILE_BYTES);
                    hd[p] = tc_smem_desc(st + (TC_SPLIT + p) * TC_TILE_BYTES);
                }
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // 8 tf32 = 32 bytes along K inside the swizzle atom
                    tc_mma(tmem_d, xd[0] + adv, hd[0] + adv, (kb | k) != 0);
                    tc_mma(tmem_d, xd[1] + adv, hd[0] + adv, 1);
                    tc_mma(tmem_d, xd[0] + adv, hd[1] + adv, 1);
                    tc_mma(tmem_d, xd[1] + adv, hd[1] + adv, 1);
                    tc_mma(tmem_d, xd[2] + adv, hd[0] + adv, 1);
                    tc_mma(tmem_d, xd[0] + adv, hd[2] + adv, 1);
                }
                tc_commit(empty(s));                  // stage reusable once these MMAs have read it
            }
            if (ok) tc_commit(accum);                 // accumulator complete
        }
    } else {
        // ------------------------------------------------------------ splitters (256 threads)
        const int t = threadIdx.x - 64;
        for (int kb = 0; ok && kb < n_kb; ++kb) {
            const int s = kb % TC_STAGES;
            ok = tc_mbar_wait(full(s), (kb / TC_STAGES) & 1, timeout_ns);
            if (!ok) break;
            float4* x1 = (float4*)(base_ptr + s * TC_STAGE_BYTES);
            float4* x2 = (float4*)(base_ptr + s * TC_STAGE_BYTES + TC_TILE_BYTES);
            float4* x3 = (float4*)(base_ptr + s * TC_STAGE_BYTES + 2 * TC_TILE_BYTES);
#pragma unroll
            for (int i = 0; i < TC_TILE_BYTES / 16 / TC_SPLITTERS; ++i) {
                const int idx = t + i * TC_SPLITTERS;
                float4 v = x1[idx];
                if (col_weight) {
                    // SWIZZLE_128B: physical chunk = logical chunk ^ (row & 7)
                    const int n = idx / CHUNKS, c = (idx % CHUNKS) ^ (n & 7);
                    const float4 w = kb * TC_BK + 4 * c < D ? __ldg((const float4*)(col_weight + kb * TC_BK) + c)
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);   // (D % 4 == 0)
                    v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w;
                }
                float4 a, b, c, d;
                a.x = tf32_head(v.x); b.x = v.x - a.x;      // every subtraction here is exact
                a.y = tf32_head(v.y); b.y = v.y - a.y;
                a.z = tf32_head(v.z); b.z = v.z - a.z;
                a.w = tf32_head(v.w); b.w = v.w - a.w;
                c.x = tf32_head(b.x); d.x = b.x - c.x;
                c.y = tf32_head(b.y); d.y = b.y - c.y;
                c.z = tf32_head(b.z); d.z = b.z - c.z;
                c.w = tf32_head(b.w); d.w = b.w - c.w;
                x1[idx] = a;
                x2[idx] = c;
                x3[idx] = d;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
            tc_mbar_arrive(split(s));
        }
        // ------------------------------------------------------------ epilogue (warps 2..5: one TMEM lane quarter each)
        if (warp < 6) {
            if (ok) ok = tc_mbar_wait(accum, 0, timeout_ns);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = warp & 3;                                        // this warp's TMEM lane quarter
            const long row = (long)m0 + q * 32 + lane;
#pragma unroll
            for (int c = 0; c < TC_BN; c += 32) {
                if (!ok) break;                                            // warp-uniform: every lane waited on `accum`
                uint32_t v[32];
                const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < rows) {
                    float4* dst = (float4*)(out + row * D + n0 + c);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (n0 + c + 4 * i < D)            // the last column tile of a padded width is partly outside
                            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                 __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                }
            }
        }
    }
    if (!ok) report_device_error(err_word, DEV_ERR_FILTER_TC);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_BN));
    }
}

// workspace layout: [H1 | H2 | H3] each D*D floats (three 11-bit pieces, H1 + H2 + H3 == H exactly),
// K-major: B[n][k] = taps[(n - k) + D - 1].
__global__ void __launch_bounds__(256) filter_tc_prepare_kernel(const float* __restrict__ taps, float* __restrict__ ws, int D, int Dp) {
    const long total = (long)Dp * Dp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int n = (int)(i / Dp), k = (int)(i - (long)n * Dp);
        const float v = (n < D && k < D) ? __ldg(taps + (n - k) + D - 1) : 0.f;
        const float h1 = tf32_head(v), r = v - h1;
        const float h2 = tf32_head(r);
        ws[i] = h1;
        ws[total + i] = h2;
        ws[2 * total + i] = r - h2;
    }
}

typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static tc_encode_fn tc_get_encode() {
    static tc_encode_fn fn = []() -> tc_encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (tc_encode_fn)p;
    }();
    return fn;
}

// [n_rows, D] float32 row-major, box = 32 floats (128 bytes, one swizzle span) x 128 rows
static int tc_make_map(CUtensorMap* tm, const float* ptr, long n_rows, int D) {
    tc_encode_fn enc = tc_get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return PDU_EUNSUPPORTED;
    }
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)D * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (filter) failed with CUresult %d (rows=%ld D=%d)", (int)rc, n_rows, D);
        return PDU_ECUDA;
    }
    return PDU_OK;
}

// any detector count whose rows are 16-byte multiples (the TMA row pitch); the contraction is padded to the 128-wide tile
bool filter_tc_supported(int D) { return D % 4 == 0 && D >= TC_BN && D <= 4096; }
static int tc_padded(int D) { return (D + TC_BN - 1) / TC_BN * TC_BN; }

size_t filter_tc_workspace_bytes(int D) {
    return filter_tc_supported(D) ? (size_t)3 * tc_padded(D) * tc_padded(D) * sizeof(float) : 0;
}

int filter_tc_prepare(const float* taps, void* ws, size_t ws_bytes, int D, cudaStream_t st) {
    if (!filter_tc_supported(D)) return PDU_OK;
    if (!ws || ws_bytes < filter_tc_workspace_bytes(D) || ((uintptr_t)ws & 15)) {
        set_error("pdu_filter_prepare_f32: workspace of %zu bytes (16-byte aligned) required", filter_tc_workspace_bytes(D));
        return PDU_ENOMEM;
    }
    const int Dp = tc_padded(D);
    const long total = (long)Dp * Dp;
    filter_tc_prepare_kernel<<<(unsigned)std::min<long>(cdiv(total, 256), 148L * 8), 256, 0, st>>>(taps, (float*)ws, D, Dp);
    PDU_LAUNCHED();
    return PDU_OK;
}

int filter_tc_launch(const float* sino, float* out, const void* ws, const float* col_weight, long rows, int D, cudaStream_t st) {
    CUtensorMap tx, th;
    int rc = tc_make_map(&tx, sino, rows, D);
    if (rc) return rc;
    const int Dp = tc_padded(D);
    rc = tc_make_map(&th, (const float*)ws, 3L * Dp, Dp);   // the three pieces stacked by rows
    if (rc) return rc;
    PDU_CUDA((ensure_dyn_smem<filter_tc_kernel>(TC_SMEM)));
    dim3 grid((unsigned)cdiv(rows, TC_BM), (unsigned)(Dp / TC_BN));
    const int fault = option(OPT_DEBUG_FAULT) == 1 ? 1 : 0;
    filter_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(tx, th, col_weight, out, rows, D, device_error_word(),
                                                        fault ? MBAR_TIMEOUT_FAULT_NS : MBAR_TIMEOUT_NS, fault);
    PDU_LAUNCHED();
    note_kernel(OP_FILTER, "filter_tc_kernel grid %ux%u (tcgen05 kind::tf32, 6-product split, M128 N128 K8, TMEM accumulator)%s", grid.x,
                grid.y, col_weight ? " + fused detector weight" : "");
    return PDU_OK;
}

}  // namespace pdu
