// fp64_rate.cu — latency and issue rate of DFMA / DADD / DMUL per SM sub-partition (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(long long* out, double* sink, double x, int iters) {
    double a[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = 1.0 + i + threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) a[i] = __fma_rn(a[i], x, 0.5);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 0.123) sink[0] = s;
}
template <int CH>
void run(long long* out, double* sink) {
    const int iters = 500;
    for (int wps : {1, 2, 4}) {
        k<CH><<<148, 128 * wps>>>(out, sink, 1.0000001, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 148; ++i) s += (double)h[i];
        printf("DFMA, %2d independent chains per lane, %d warps per sub-partition: %6.2f cycles per warp DFMA per sub-partition, %6.2f per step of one chain\n",
               CH, wps, s / 148.0 / iters / (8 * CH) / wps, s / 148.0 / iters / 8);
    }
}
int main() {
    long long* out; double* sink;
    cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 8);
    run<1>(out, sink); run<2>(out, sink); run<4>(out, sink); run<8>(out, sink);
    return 0;
}
