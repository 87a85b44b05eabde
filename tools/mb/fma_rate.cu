// fma_rate.cu — issue rate of FFMA, FFMA2 and mixes of the two per SM sub-partition (sm_100a), no memory traffic:
// `wps` warps per sub-partition, 16 independent accumulator chains per lane, cycles per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float x, float2 w, float2 acc) {
    unsigned long long xx, ww, aa, dd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w.x), "f"(w.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(acc.x), "f"(acc.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(xx), "l"(ww), "l"(aa));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(dd));
    return r;
}

// MODE 0: 16 FFMA per round; 1: 16 FFMA2; 2: 8 FFMA2 + 16 FFMA (same FMA count on both pipes); 3: 8 FFMA2 + 8 FFMA
template <int MODE>
__global__ void k(long long* out, float* sink, float x, int iters) {
    float a[16];
    float2 b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = i + threadIdx.x; b[i] = make_float2(i, threadIdx.x); }
    const float2 w = make_float2(x, x * 0.5f);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], x, w.y);
        } else if constexpr (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) b[i] = ffma2(x, w, b[i]);
        } else if constexpr (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                b[i] = ffma2(x, w, b[i]);
                a[2 * i] = __fmaf_rn(a[2 * i], x, w.y);
                a[2 * i + 1] = __fmaf_rn(a[2 * i + 1], x, w.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                b[i] = ffma2(x, w, b[i]);
                a[i] = __fmaf_rn(a[i], x, w.y);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + b[i].x + b[i].y;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 0.123f) sink[0] = s;
}

template <int MODE>
void run(const char* name, int per_round, long long* out, float* sink) {
    const int iters = 4000;
    for (int wps : {1, 2, 4}) {
        k<MODE><<<148, 128 * wps>>>(out, sink, 1.0001f, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("%-40s %d warps per sub-partition: %6.2f cycles per warp instruction per sub-partition\n", name, wps,
               s / 148.0 / iters / per_round / wps);
    }
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 148 * 8);
    cudaMalloc(&sink, 4);
    run<0>("16 FFMA", 16, out, sink);
    run<1>("16 FFMA2", 16, out, sink);
    run<2>("8 FFMA2 + 16 FFMA", 24, out, sink);
    run<3>("8 FFMA2 + 8 FFMA", 16, out, sink);
    return 0;
}
