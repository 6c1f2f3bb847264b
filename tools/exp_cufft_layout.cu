// Experiment: batched 2-D C2C cuFFT, plane-contiguous vs plane-interleaved (batch innermost) layouts.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/exp_cufft_layout.cu -lcufft -o tools/exp_cufft_layout
#include <cstdio>
#include <cuda_runtime.h>
#include <cufft.h>
static float time_plan(cufftHandle h, cufftComplex* d, void* flush, size_t fb, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9f;
    for (int i = 0; i < reps + 2; ++i) {
        cudaMemsetAsync(flush, 0, fb);
        cudaEventRecord(a);
        cufftExecC2C(h, d, d, CUFFT_FORWARD);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (i >= 2 && ms < best) best = ms;
    }
    return best * 1e3f;
}
int main(int argc, char** argv) {
    int K = argc > 1 ? atoi(argv[1]) : 640, P = argc > 2 ? atoi(argv[2]) : 64;
    size_t n = (size_t)K * K * P;
    cufftComplex* d; cudaMalloc(&d, n * sizeof(cufftComplex)); cudaMemset(d, 0, n * sizeof(cufftComplex));
    void* flush; size_t fb = 256u << 20; cudaMalloc(&flush, fb);
    int dims[2] = {K, K};
    cufftHandle h1, h2, h3, h4;
    cufftPlanMany(&h1, 2, dims, dims, 1, K * K, dims, 1, K * K, CUFFT_C2C, P);
    printf("contiguous planes   : %8.1f us\n", time_plan(h1, d, flush, fb, 5));
    cufftResult r = cufftPlanMany(&h2, 2, dims, dims, P, 1, dims, P, 1, CUFFT_C2C, P);
    if (r == CUFFT_SUCCESS) printf("interleaved planes  : %8.1f us\n", time_plan(h2, d, flush, fb, 5));
    else printf("interleaved plan failed %d\n", (int)r);
    // 1-D passes alone: rows (contiguous) and columns (stride K) of contiguous planes
    int d1[1] = {K};
    cufftPlanMany(&h3, 1, d1, d1, 1, K, d1, 1, K, CUFFT_C2C, K * P);
    printf("1-D rows, all planes: %8.1f us\n", time_plan(h3, d, flush, fb, 5));
    r = cufftPlanMany(&h4, 1, d1, d1, K, 1, d1, K, 1, CUFFT_C2C, K);   // one plane's columns; loop planes
    if (r == CUFFT_SUCCESS) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 3; ++rep) {
            cudaMemsetAsync(flush, 0, fb);
            cudaEventRecord(a);
            for (int p = 0; p < P; ++p) cufftExecC2C(h4, d + (size_t)p * K * K, d + (size_t)p * K * K, CUFFT_FORWARD);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep == 2) printf("1-D columns, %d calls : %8.1f us\n", P, ms * 1e3f);
        }
    }
    return 0;
}
