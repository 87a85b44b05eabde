// microbench_lds.cu — shared-memory load throughput by access pattern (sm_100a): cycles per warp-wide
// LDS instruction per SM when 16 warps issue independent loads back to back.
#include <cstdio>
#include <cuda_runtime.h>

template <int BYTES>
__device__ __forceinline__ float lds(const float* p) {
    if constexpr (BYTES == 16) { float4 v = *reinterpret_cast<const float4*>(p); return v.x + v.y + v.z + v.w; }
    else if constexpr (BYTES == 8) { float2 v = *reinterpret_cast<const float2*>(p); return v.x + v.y; }
    else return *p;
}

// pattern: lane -> float offset = ((lane / div) % mod) * stride
template <int BYTES>
__global__ void k_lds(long long* out, int div, int mod, int stride) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 12288; i += blockDim.x) sm[i] = 1.0f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float* p = sm + ((lane / div) % mod) * stride;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += lds<BYTES>(p + u * 1024 + (it & 7) * 16);
    }
    long long t1 = clock64();
    float s = 0;
    for (int u = 0; u < 8; ++u) s += acc[u];
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (s == 0.12345f) out[1] = 1;
}

int main() {
    long long* o;
    cudaMalloc(&o, 64);
    struct P { const char* name; int bytes, div, mod, stride; } ps[] = {
        {"LDS.128 all lanes same address", 16, 32, 1, 0},
        {"LDS.128 4 distinct, one per quarter-warp (x tile 4t)", 16, 8, 4, 4},
        {"LDS.128 4 distinct, lane%4", 16, 1, 4, 4},
        {"LDS.128 8 distinct contiguous, lane%8", 16, 1, 8, 4},
        {"LDS.128 8 distinct, lane/4", 16, 4, 8, 4},
        {"LDS.128 32 distinct contiguous", 16, 1, 32, 4},
        {"LDS.64 all same", 8, 32, 1, 0},
        {"LDS.64 8 distinct contiguous, lane%8 (w tile 2n)", 8, 1, 8, 2},
        {"LDS.64 4 distinct, lane/8", 8, 8, 4, 2},
        {"LDS.64 32 distinct contiguous", 8, 1, 32, 2},
        {"LDS.32 all same", 4, 32, 1, 0},
        {"LDS.32 32 distinct contiguous", 4, 1, 32, 1},
        {"LDS.32 16 distinct, lane%16", 4, 1, 16, 1},
        {"LDS.32 2 distinct, lane/16", 4, 16, 2, 1},
    };
    cudaFuncSetAttribute(k_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (auto& p : ps) {
        for (int warps : {1, 4, 16}) {
            cudaMemset(o, 0, 64);
            if (p.bytes == 16) k_lds<16><<<1, 32 * warps, 65536>>>(o, p.div, p.mod, p.stride);
            else if (p.bytes == 8) k_lds<8><<<1, 32 * warps, 65536>>>(o, p.div, p.mod, p.stride);
            else k_lds<4><<<1, 32 * warps, 65536>>>(o, p.div, p.mod, p.stride);
            cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
            printf("%-56s %2d warps: %7.2f cycles per warp-LDS per SM\n", p.name, warps, (double)h[0] / (512.0 * 8 * warps));
        }
    }
    return 0;
}
