// microbenchmark: lane = tree, x[64] in registers, weights from __constant__ memory (warp-uniform operand)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ float Wc[2 * 64 * 64];

template <int BASE>
__device__ __forceinline__ void fixed_layer(const float (&x)[64], float (&acc)[16]) {
#pragma unroll
    for (int k = 0; k < 64; ++k) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(x[k], Wc[BASE + k * 64 + j], acc[j]);
    }
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) layer_kernel(const float* __restrict__ xin, float* __restrict__ out, long long* clk, int iters) {
    __shared__ float xs[64][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = xin[lane * 64 + k];
    int base;
    if (MODE == 0) base = __shfl_sync(0xffffffffu, warp * 16, 0);
    else base = warp * 16;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int lb = base + (it & 1) * 4096;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        if (MODE == 2) {
            const int sel = warp * 2 + (it & 1);
            switch (sel) {
                case 0: fixed_layer<0>(x, acc); break;
                case 1: fixed_layer<4096>(x, acc); break;
                case 2: fixed_layer<16>(x, acc); break;
                case 3: fixed_layer<4096 + 16>(x, acc); break;
                case 4: fixed_layer<32>(x, acc); break;
                case 5: fixed_layer<4096 + 32>(x, acc); break;
                case 6: fixed_layer<48>(x, acc); break;
                default: fixed_layer<4096 + 48>(x, acc); break;
            }
        } else {
#pragma unroll
        for (int k = 0; k < 64; ++k) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(x[k], Wc[lb + k * 64 + j], acc[j]);
        }
        }
        // exchange: every warp needs all 64 outputs of every tree
#pragma unroll
        for (int j = 0; j < 16; ++j) xs[warp * 16 + j][lane] = fmaxf(acc[j], 0.f);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 64; ++k) x[k] = xs[k][lane];
        __syncthreads();
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) s += x[k];
    out[blockIdx.x * 128 + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    float* h = new float[8192];
    for (int i = 0; i < 8192; ++i) h[i] = ((i * 2654435761u) >> 8 & 0xffff) / 65536.0f * 0.2f - 0.1f;
    cudaMemcpyToSymbol(Wc, h, sizeof(float) * 8192);
    float *xin, *out; long long* clk;
    cudaMalloc(&xin, 32 * 64 * 4); cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&clk, 148 * 8);
    cudaMemcpy(xin, h, 32 * 64 * 4, cudaMemcpyHostToDevice);
    const int iters = 100;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) layer_kernel<0><<<148, 128>>>(xin, out, clk, iters);
            else if (mode == 1) layer_kernel<1><<<148, 128>>>(xin, out, clk, iters);
            else layer_kernel<2><<<148, 128>>>(xin, out, clk, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long c[148];
            cudaMemcpy(c, clk, sizeof(c), cudaMemcpyDeviceToHost);
            printf("mode %d rep %d: %s  cycles per layer (4 warps, 32 trees x 64 neurons): %.1f\n", mode, rep, cudaGetErrorString(e), (double)c[0] / iters);
        }
    }
    return 0;
}
