// lds_rate.cu — how fast ONE warp can issue shared-memory loads (sm_100a): U independent LDS per round (U = 4, 8, 16),
// widths 32 / 64 / 128 bits, broadcast (all lanes one address) and row-per-quarter-warp patterns, 1 / 2 / 4 / 16 warps per SM.
// Reports cycles per warp LDS instruction as seen by one warp (latency-bound if U is small) and per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int W, int U, int PAT>
__global__ void k(long long* out, float* sink, int iters) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // PAT 0: broadcast; PAT 1: lane/8 selects one of 4 rows 68 floats apart (the heads' x pattern); PAT 2: 32 distinct contiguous
    const int base = PAT == 0 ? 0 : PAT == 1 ? (lane >> 3) * 68 : lane * (W / 32);
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int o = base + (it & 7) * 4;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (W == 128) {
                const float4 v = *reinterpret_cast<const float4*>(buf + o + u * 8 * (PAT == 2 ? 16 : 1) * 4 % 4096);
                acc[u] += v.x + v.w;
            } else if constexpr (W == 64) {
                const float2 v = *reinterpret_cast<const float2*>(buf + o + u * 8 * (PAT == 2 ? 8 : 1) * 2 % 4096);
                acc[u] += v.x + v.y;
            } else {
                acc[u] += buf[o + u * 8 * (PAT == 2 ? 4 : 1) % 4096];
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) s += acc[u];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 0.123f) sink[0] = s;
}

template <int W, int U, int PAT>
void run(long long* out, float* sink) {
    const int iters = 2000;
    const char* pat = PAT == 0 ? "broadcast" : PAT == 1 ? "4 rows (lane/8)" : "32 distinct";
    for (int warps : {1, 2, 4, 16}) {
        k<W, U, PAT><<<148, 32 * warps>>>(out, sink, iters);
        cudaDeviceSynchronize();
        k<W, U, PAT><<<148, 32 * warps>>>(out, sink, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < 148; ++i) m += h[i];
        m /= 148;
        printf("LDS.%-3d %-16s %2d independent per round, %2d warps/SM: %6.2f cycles per LDS for a warp, %5.2f per SM\n", W, pat, U, warps,
               m / (iters * (double)U), m / (iters * (double)U * warps));
    }
}

int main() {
    long long* out;
    float* sink;
    cudaMalloc(&out, 148 * sizeof(long long));
    cudaMalloc(&sink, 4);
    run<128, 4, 0>(out, sink);
    run<128, 16, 0>(out, sink);
    run<128, 16, 1>(out, sink);
    run<128, 16, 2>(out, sink);
    run<64, 16, 0>(out, sink);
    run<32, 16, 0>(out, sink);
    run<32, 16, 2>(out, sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
