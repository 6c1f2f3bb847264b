// Shared plumbing of libpdu_b200: error reporting, launch accounting, option switches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pdu.h"

namespace pdu {

void set_error(const char* fmt, ...);
int option(int which);                 // see enum Opt; -1 == default
void count_launch(int n = 1);
int sm_count();                        // cached multiProcessorCount of the current device

enum Opt { OPT_RADON_FWD = 0, OPT_RADON_ADJ = 1, OPT_FILTER = 2, OPT_NUFFT_ADJ = 3, OPT_NUFFT_FWD = 4, OPT_COUNT };

#define PDU_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            pdu::set_error(__VA_ARGS__);               \
            return PDU_EINVAL;                         \
        }                                              \
    } while (0)

#define PDU_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            pdu::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PDU_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

// after a <<<>>> launch
#define PDU_LAUNCHED()                                                                      \
    do {                                                                                    \
        pdu::count_launch();                                                                \
        cudaError_t e__ = cudaPeekAtLastError();                                            \
        if (e__ != cudaSuccess) {                                                           \
            (void)cudaGetLastError();                                                       \
            pdu::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return PDU_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

static inline long cdiv(long a, long b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device, size): the attribute belongs to the device's
// context, so a process that drives several GPUs must set it on each of them (a plain `static bool` did it once).
// Kern is the kernel itself (non-type template parameter): one record per kernel instantiation.
template <auto Kern>
inline cudaError_t ensure_dyn_smem(int bytes) {
    static int set_bytes[64] = {};             // per device ordinal; racing first calls repeat an idempotent call
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int& rec = set_bytes[dev & 63];
    if (rec >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) rec = bytes;
    return e;
}
#endif

#ifdef __CUDACC__
// packed-FP32 helpers (sm_100 FFMA2 / FADD2): a 64-bit register holds (lo, hi) floats
typedef unsigned long long ull;
__device__ __forceinline__ ull pk2(float lo, float hi) { ull r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(ull v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ ull fma2(ull a, ull b, ull c) { ull d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ ull add2(ull a, ull b) { ull d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ull add2_rm(ull a, ull b) { ull d; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ull sub2(ull a, ull b) { ull d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ ull mul2(ull a, ull b) { ull d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#endif

}  // namespace pdu
