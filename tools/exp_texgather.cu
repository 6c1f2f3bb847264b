// Experiment (r02, VERDICT item 4 "also try"): the ceiling of a tex2Dgather-based forward projector.
// One thread per ray, unit steps clipped to the image square, the four bilinear corners of every sample from ONE
// tex2Dgather on the raw float image (exact fp32 corners; weights in fp32 by the thread) -- against the same loop with
// hardware bilinear filtering (tex2D, 9-bit weights: what torch_radon does) and with plain __ldg corner loads.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/exp_texgather tools/exp_texgather.cu && tools/exp_texgather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct RaySeg { float x0, y0, vx, vy; int n; };

__device__ __forceinline__ RaySeg ray(int N, float cs, float sn, int d, int D) {
    // detector coordinate s, ray direction (-sn, cs), clipped to [-N/2, N/2]^2; pixel-centre coordinates
    const float h = 0.5f * N, s = (float)d - 0.5f * D + 0.5f;
    const float px = s * cs, py = s * sn, dx = -sn, dy = cs;
    float t0 = -1e9f, t1 = 1e9f;
    if (fabsf(dx) > 1e-6f) { float a = (-h - px) / dx, b = (h - px) / dx; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    else if (fabsf(px) > h) t1 = -1e9f;
    if (fabsf(dy) > 1e-6f) { float a = (-h - py) / dy, b = (h - py) / dy; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    else if (fabsf(py) > h) t1 = -1e9f;
    RaySeg r;
    r.n = t1 > t0 ? (int)ceilf(t1 - t0) : -1;
    r.x0 = px + t0 * dx + h - 0.5f; r.y0 = py + t0 * dy + h - 0.5f; r.vx = dx; r.vy = dy;
    return r;
}

template <int MODE>   // 0: tex2Dgather + fp32 weights, 1: tex2D linear (9-bit weights), 2: four __ldg
__global__ void __launch_bounds__(256) proj(cudaTextureObject_t tex, const float* __restrict__ img, float* __restrict__ sino,
                                            const float2* __restrict__ trig, int N, int A, int D, int rows_per_slice) {
    // warp = 4 detectors x 8 views (like the library's cell kernel); block = 32 detectors x 8 views
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 32 + warp * 4 + (lane & 3), a = blockIdx.y * 8 + (lane >> 2), b = blockIdx.z;
    if (d >= D || a >= A) return;
    const float2 t = trig[a];
    const RaySeg r = ray(N, t.x, t.y, d, D);
    const float yoff = (float)(b * rows_per_slice);
    float acc = 0.f;
    for (int j = 0; j <= r.n; ++j) {
        const float x = fmaf((float)j, r.vx, r.x0), y = fmaf((float)j, r.vy, r.y0);
        if (MODE == 0) {
            const float xf = floorf(x), yf = floorf(y), fx = x - xf, fy = y - yf;
            // footprint of (xf + 1, yf + 1) in texel-centre coordinates = texels (xf, yf) .. (xf + 1, yf + 1)
            const float4 g = tex2Dgather<float4>(tex, xf + 1.0f, yf + 1.0f + yoff, 0);
            // g.w = (i, j), g.z = (i+1, j), g.x = (i, j+1), g.y = (i+1, j+1)
            const float top = fmaf(fx, g.z - g.w, g.w), bot = fmaf(fx, g.y - g.x, g.x);
            acc += fmaf(fy, bot - top, top);
        } else if (MODE == 1) {
            acc += tex2D<float>(tex, x + 0.5f, y + 0.5f + yoff);
        } else {
            const float xf = floorf(x), yf = floorf(y), fx = x - xf, fy = y - yf;
            const int ix = (int)xf, iy = (int)yf;
            const bool x0 = (unsigned)ix < (unsigned)N, x1 = (unsigned)(ix + 1) < (unsigned)N;
            const bool y0 = (unsigned)iy < (unsigned)N, y1 = (unsigned)(iy + 1) < (unsigned)N;
            const float* p = img + ((long)b * rows_per_slice + iy) * N + ix;
            const float v00 = (x0 && y0) ? __ldg(p) : 0.f, v01 = (x1 && y0) ? __ldg(p + 1) : 0.f;
            const float v10 = (x0 && y1) ? __ldg(p + N) : 0.f, v11 = (x1 && y1) ? __ldg(p + N + 1) : 0.f;
            const float top = fmaf(fx, v01 - v00, v00), bot = fmaf(fx, v11 - v10, v10);
            acc += fmaf(fy, bot - top, top);
        }
    }
    sino[((long)b * A + a) * D + d] = acc;
}

__global__ void count_samples(const float2* trig, int N, int A, int D, unsigned long long* total) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (d >= D) return;
    const RaySeg r = ray(N, trig[a].x, trig[a].y, d, D);
    if (r.n >= 0) atomicAdd(total, (unsigned long long)(r.n + 1));
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256, A = argc > 2 ? atoi(argv[2]) : 512, B = argc > 3 ? atoi(argv[3]) : 16, D = N;
    const int RPS = N + 1;                                  // one zero row between slices: no bleeding across the stack
    std::vector<float> h((size_t)B * RPS * N, 0.f);
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N * N; ++i) h[(size_t)b * RPS * N + i] = (float)rand() / RAND_MAX;
    std::vector<float2> ht(A);
    for (int a = 0; a < A; ++a) ht[a] = make_float2((float)cos(M_PI * a / A), (float)sin(M_PI * a / A));
    float *img, *s0, *s1, *s2; float2* trig; unsigned long long* total;
    size_t pitch;
    CK(cudaMallocPitch(&img, &pitch, (size_t)N * 4, (size_t)B * RPS));
    if (pitch != (size_t)N * 4) { printf("pitch %zu != %d\n", pitch, N * 4); }
    CK(cudaMemcpy2D(img, pitch, h.data(), (size_t)N * 4, (size_t)N * 4, (size_t)B * RPS, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&s0, (size_t)B * A * D * 4)); CK(cudaMalloc(&s1, (size_t)B * A * D * 4)); CK(cudaMalloc(&s2, (size_t)B * A * D * 4));
    CK(cudaMalloc(&trig, A * sizeof(float2))); CK(cudaMemcpy(trig, ht.data(), A * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&total, 8)); CK(cudaMemset(total, 0, 8));
    count_samples<<<dim3((D + 127) / 128, A), 128>>>(trig, N, A, D, total);
    unsigned long long nsamp; CK(cudaMemcpy(&nsamp, total, 8, cudaMemcpyDeviceToHost));
    const double samples = (double)nsamp * B;

    cudaTextureObject_t tex[2];
    for (int m = 0; m < 2; ++m) {
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = img;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<float>(); rd.res.pitch2D.width = N; rd.res.pitch2D.height = (size_t)B * RPS;
        rd.res.pitch2D.pitchInBytes = pitch;
        cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
        td.filterMode = m == 0 ? cudaFilterModePoint : cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        CK(cudaCreateTextureObject(&tex[m], &rd, &td, nullptr));
    }
    void* flush; CK(cudaMalloc(&flush, 256 << 20));
    dim3 grid((D + 31) / 32, (A + 7) / 8, B);
    auto run = [&](int mode, float* out) {
        if (mode == 0) proj<0><<<grid, 256>>>(tex[0], img, out, trig, N, A, D, RPS);
        else if (mode == 1) proj<1><<<grid, 256>>>(tex[1], img, out, trig, N, A, D, RPS);
        else proj<2><<<grid, 256>>>(tex[0], img, out, trig, N, A, D, RPS);
    };
    const char* names[3] = {"tex2Dgather + fp32 weights", "tex2D linear (9-bit weights)", "four __ldg corners"};
    float* outs[3] = {s0, s1, s2};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        run(mode, outs[mode]); CK(cudaDeviceSynchronize());
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(flush, 0, 256 << 20));
            cudaEventRecord(e0); run(mode, outs[mode]); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
        }
        printf("%-32s %8.1f us  %6.3f T samples/s\n", names[mode], best * 1e3, samples / (best * 1e-3) / 1e12);
    }
    // agreement of the gather path with the __ldg path (both exact fp32 weights)
    std::vector<float> r0((size_t)B * A * D), r2((size_t)B * A * D), r1((size_t)B * A * D);
    CK(cudaMemcpy(r0.data(), s0, r0.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), s2, r2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r1.data(), s1, r1.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0, num1 = 0;
    for (size_t i = 0; i < r0.size(); ++i) { num += (double)(r0[i] - r2[i]) * (r0[i] - r2[i]); num1 += (double)(r1[i] - r2[i]) * (r1[i] - r2[i]); den += (double)r2[i] * r2[i]; }
    printf("samples %.3f G;  rel-L2 gather vs ldg %.3e;  hardware-linear vs ldg %.3e\n", samples / 1e9, sqrt(num / den), sqrt(num1 / den));
    // cost of making the texture object (it would be per call, or cached per image pointer)
    {
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = img;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<float>(); rd.res.pitch2D.width = N; rd.res.pitch2D.height = (size_t)B * RPS;
        rd.res.pitch2D.pitchInBytes = pitch;
        cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint;
        timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 100; ++i) { cudaTextureObject_t t; CK(cudaCreateTextureObject(&t, &rd, &td, nullptr)); CK(cudaDestroyTextureObject(t)); }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("cudaCreateTextureObject + destroy: %.1f us each (host)\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 100 / 1e3);
    }
    return 0;
}
