// Stand-alone TMA diagnostic: loads one (W x ROWS x 1) box of a [B, n, n] float tensor into shared
// memory and copies it out.  argv[1] selects the sub-test so that a fault stays in its own process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W, int ROWS>
__global__ void box_kernel(const __grid_constant__ CUtensorMap tm, const CUtensorMap* tm_g, int use_global, int c0, int c1,
                           int c2, float* out) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = (unsigned char*)(((uintptr_t)dyn + 127) & ~(uintptr_t)127);
    uint64_t* bar = (uint64_t*)(base + ((W * ROWS * 4 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap* p = use_global ? tm_g : &tm;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(W * ROWS * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(s32(base)), "l"((uint64_t)p), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
    }
    bool ok = false;
    for (int spin = 0; spin < (1 << 20) && !ok; ++spin) {
        uint32_t r;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(r) : "r"(s32(bar)), "r"(0) : "memory");
        ok = r != 0;
    }
    const float* t = (const float*)base;
    for (int i = threadIdx.x; i < W * ROWS; i += blockDim.x) out[i] = ok ? t[i] : -777.f;
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int ROWS>
int run(int n, int B, int c0, int c1, int c2, int use_global) {
    std::vector<float> h((size_t)B * n * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.5f + 1.f;
    float *d, *o;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMalloc(&o, W * ROWS * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    if (!fp) { printf("  no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
    cuuint32_t box[3] = {W, ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = ((enc_fn)fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("  encode rc=%d  n=%d B=%d box=%dx%d coords=(%d,%d,%d) desc_in_%s\n", (int)rc, n, B, W, ROWS, c0, c1, c2,
           use_global ? "global" : "param");
    if (rc != CUDA_SUCCESS) return 1;
    CUtensorMap* tmg;
    CK(cudaMalloc(&tmg, sizeof(tm)));
    CK(cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    const int smem = ((W * ROWS * 4 + 127) / 128) * 128 + 8 + 128;
    CK(cudaFuncSetAttribute(box_kernel<W, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    box_kernel<W, ROWS><<<1, 128, smem>>>(tm, tmg, use_global, c0, c1, c2, o);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> r(W * ROWS);
    CK(cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < ROWS; ++y)
        for (int x = 0; x < W; ++x) {
            const int gx = c0 + x, gy = c1 + y;
            const float want = (gx >= 0 && gx < n && gy >= 0 && gy < n) ? h[((size_t)c2 * n + gy) * n + gx] : 0.f;
            if (r[y * W + x] != want) ++bad;
        }
    printf("  mismatches: %ld of %d (first value %g)\n", bad, W * ROWS, r[0]);
    return bad != 0;
}

int main(int argc, char** argv) {
    const int t = argc > 1 ? atoi(argv[1]) : 0;
    printf("diag_tma test %d\n", t);
    switch (t) {
        case 0: return run<32, 8>(64, 2, 0, 0, 0, 0);          // small box, in range, descriptor in param
        case 1: return run<32, 8>(64, 2, 0, 0, 0, 1);          // descriptor in global memory
        case 2: return run<32, 8>(64, 2, -3, -1, 1, 0);        // negative coordinates (zero fill)
        case 3: return run<160, 33>(256, 2, 40, 31, 1, 0);     // the projector's box
        case 4: return run<160, 33>(64, 2, -5, -1, 1, 0);      // box wider than the tensor
        case 5: return run<240, 33>(256, 2, 200, 250, 0, 0);   // past the far edges
        case 6: return run<160, 33>(64, 2, -5, -1, 1, 1);
        case 7: return run<32, 8>(64, 2, 3, 0, 0, 0);          // misaligned positive x
        case 8: return run<32, 8>(64, 2, -4, 0, 0, 0);         // aligned negative x
        case 9: return run<32, 8>(64, 2, 0, -1, 0, 0);         // negative y only
        case 10: return run<32, 8>(64, 2, 4, 2, 1, 0);
        case 11: return run<32, 8>(64, 2, -4, -1, 1, 0);
        case 12: return run<32, 8>(64, 2, -3, 0, 0, 0);        // misaligned negative x
        case 13: return run<32, 8>(64, 2, 61, 60, 1, 0);       // misaligned, past both far edges
    }
    return 0;
}
