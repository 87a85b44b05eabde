// microbench.cu — latency / issue-rate probes for the instructions the search kernels are made of
// (sm_100a).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
// Each probe runs ITER dependent (or independent) instructions in one warp (or W warps) of one CTA and
// reports SM cycles per instruction from clock64().
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

__device__ __forceinline__ unsigned long long ffma2_raw(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// NCH independent FFMA2 chains per lane
template <int NCH>
__global__ void k_ffma2(long long* out, unsigned long long seed) {
    unsigned long long acc[NCH];
    for (int i = 0; i < NCH; ++i) acc[i] = seed + i;
    unsigned long long a = seed * 3, b = seed * 5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < NCH; ++i) acc[i] = ffma2_raw(a, b, acc[i]);
    }
    long long t1 = clock64();
    unsigned long long s = 0;
    for (int i = 0; i < NCH; ++i) s += acc[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 0x1234567) out[1] = s;
}

template <int NCH>
__global__ void k_ffma(long long* out, float seed) {
    float acc[NCH];
    for (int i = 0; i < NCH; ++i) acc[i] = seed + i;
    float a = seed * 3, b = seed * 5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(a), "f"(b));
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < NCH; ++i) s += acc[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 0.1234567f) out[1] = 1;
}

template <int NCH>
__global__ void k_dfma(long long* out, double seed) {
    double acc[NCH];
    for (int i = 0; i < NCH; ++i) acc[i] = seed + i;
    double a = seed * 1.0000001, b = seed * 1e-9;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < NCH; ++i) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < NCH; ++i) s += acc[i];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (s == 0.1234567) out[1] = 1;
}

__global__ void k_dadd(long long* out, double seed) {
    double acc = seed;
    double b = seed * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc) : "d"(b));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0.1234567) out[1] = 1;
}

__global__ void k_dmul(long long* out, double seed) {
    double acc = seed;
    double b = 1.0000000001;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(acc) : "d"(b));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0.1234567) out[1] = 1;
}

// int -> double -> int conversion chain (I2F.F64 + F2I.F64)
__global__ void k_i2d(long long* out, int seed) {
    int v = seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double d;
            asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(v));
            asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(v) : "d"(d));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v == 1234567) out[1] = 1;
}

// float -> double -> float chain
__global__ void k_f2d(long long* out, float seed) {
    float v = seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(v) : "d"(d));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v == 0.1234567f) out[1] = 1;
}

// double compare + select chain (DSETP + SEL)
__global__ void k_dsetp(long long* out, double seed) {
    double a = seed, b = seed * 0.5;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(a) : "d"(b));
            asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(b) : "d"(a));
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 2;
    if (a + b == 0.1234567) out[1] = 1;
}

// pointer chase through shared memory (LDS latency), 32-bit and 128-bit
__global__ void k_lds(long long* out) {
    __shared__ int4 buf[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = make_int4((i * 7 + 1) & 255, 0, 0, 0);
    __syncthreads();
    int p = threadIdx.x & 255;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) p = buf[p].x;
    }
    long long t1 = clock64();
    int4 q = buf[p];
    long long t2 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { q = buf[(q.x + q.w) & 255]; }
    }
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[2] = t3 - t2; }
    if (p + q.y == 1234567) out[1] = 1;
}

// pointer chase through global memory that stays in L1 (LDG latency, 128-bit records)
__global__ void k_ldg(long long* out, int4* g, int n) {
    int p = threadIdx.x % n;
    // warm
    for (int i = 0; i < n; ++i) p = g[p].x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) p = g[p].x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (p == 1234567) out[1] = 1;
}

// same chase with ld.global.cg (L2 hit latency)
__global__ void k_ldg_cg(long long* out, int4* g, int n) {
    int p = threadIdx.x % n;
    for (int i = 0; i < n; ++i) p = g[p].x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            int4 v;
            asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g + p));
            p = v.x;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (p == 1234567) out[1] = 1;
}

// store then load of the same global address from the same thread (RMW round trip as in backup -> select)
__global__ void k_stld(long long* out, int4* g) {
    int p = threadIdx.x;
    int4 v = g[p];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v.x += 1;
            asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(g + p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            asm volatile("ld.global.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g + p) : "memory");
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v.x == 1234567) out[1] = 1;
}

__global__ void k_shfl(long long* out, int seed) {
    int v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v == 1234567) out[1] = 1;
}

__global__ void k_vote(long long* out, int seed) {
    int v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v += __any_sync(0xffffffffu, v & 1);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (v == 1234567) out[1] = 1;
}

// named barrier ping-pong between 2 groups of warps (cost of one bar.sync round)
__global__ void k_bar(long long* out, int count) {
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

// the team kernel's layer loop in isolation: W warps, each 64 k-steps x (2 LDS.128 + 4 FFMA2), REP times
__global__ void k_layer(long long* out, int XS, int wstride, int H) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1.0f / (1 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane >> 3, np = lane & 7;
    const float* xp = sm + 4 * q;
    const float* wp = sm + 2048 + 4 * np + 32 * (warp & 3);
    float2 a00 = make_float2(0, 0), a01 = a00, a10 = a00, a11 = a00;
    long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 16; ++rep) {
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float4 xv = *reinterpret_cast<const float4*>(xp + (size_t)k * XS);
            const float4 wv = *reinterpret_cast<const float4*>(wp + (size_t)k * wstride);
            unsigned long long x01, x23, w0, w1, c;
            asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "f"(xv.x), "f"(xv.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "f"(xv.z), "f"(xv.w));
            asm("mov.b64 %0, {%1, %2};" : "=l"(w0) : "f"(wv.x), "f"(wv.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(w1) : "f"(wv.z), "f"(wv.w));
#define STEP(acc, xx, ww)                                                         \
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));              \
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(xx), "l"(ww));             \
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
            STEP(a00, x01, w0) STEP(a01, x23, w0) STEP(a10, x01, w1) STEP(a11, x23, w1)
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 16;
    if (a00.x + a01.x + a10.y + a11.y == 0.1234567f) out[1] = 1;
}

// scalar-FFMA layer loops with compile-time strides: lane tile TQ trees x NV neurons.
// x[k][XS] (XS floats per k), w[k][NB] for the warp's neuron block (NB = NV * lanes-per-quad-row)
template <int TQ, int NV, int XS, int NB, int PIPE>
__global__ void k_layer_s(long long* out, int H) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1.0f / (1 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NPL = NB / NV;           // lanes along the neuron axis
    const int q = lane / NPL, np = lane % NPL;
    const float* xp = sm + TQ * q;
    const float* wp = sm + 4096 + NV * np + (warp & 3) * 64 * NB;
    float acc[NV][TQ];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[v][t] = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 16; ++rep) {
        const float* x = xp;
        const float* w = wp;
        if constexpr (PIPE == 0) {
#pragma unroll 1
            for (int k0 = 0; k0 < H; k0 += 8) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    float xv[TQ], wv[NV];
                    if constexpr (TQ == 4) { const float4 t = *reinterpret_cast<const float4*>(x + kk * XS); xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w; }
                    else if constexpr (TQ == 2) { const float2 t = *reinterpret_cast<const float2*>(x + kk * XS); xv[0] = t.x; xv[1] = t.y; }
                    else { xv[0] = x[kk * XS]; }
                    if constexpr (NV == 4) { const float4 t = *reinterpret_cast<const float4*>(w + kk * NB); wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w; }
                    else if constexpr (NV == 2) { const float2 t = *reinterpret_cast<const float2*>(w + kk * NB); wv[0] = t.x; wv[1] = t.y; }
                    else { wv[0] = w[kk * NB]; }
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = __fmaf_rn(xv[t], wv[v], acc[v][t]);
                }
                x += 8 * XS;
                w += 8 * NB;
            }
        } else {
            // register double buffering: loads of block i+1 are issued before the FMAs of block i
            float xb[2][4][TQ], wb[2][4][NV];
            auto load = [&](int buf, const float* xx, const float* ww) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if constexpr (TQ == 4) { const float4 t = *reinterpret_cast<const float4*>(xx + kk * XS); xb[buf][kk][0] = t.x; xb[buf][kk][1] = t.y; xb[buf][kk][2] = t.z; xb[buf][kk][3] = t.w; }
                    else if constexpr (TQ == 2) { const float2 t = *reinterpret_cast<const float2*>(xx + kk * XS); xb[buf][kk][0] = t.x; xb[buf][kk][1] = t.y; }
                    else { xb[buf][kk][0] = xx[kk * XS]; }
                    if constexpr (NV == 4) { const float4 t = *reinterpret_cast<const float4*>(ww + kk * NB); wb[buf][kk][0] = t.x; wb[buf][kk][1] = t.y; wb[buf][kk][2] = t.z; wb[buf][kk][3] = t.w; }
                    else if constexpr (NV == 2) { const float2 t = *reinterpret_cast<const float2*>(ww + kk * NB); wb[buf][kk][0] = t.x; wb[buf][kk][1] = t.y; }
                    else { wb[buf][kk][0] = ww[kk * NB]; }
                }
            };
            load(0, x, w);
#pragma unroll 1
            for (int k0 = 0; k0 < H; k0 += 8) {
                load(1, x + 4 * XS, w + 4 * NB);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = __fmaf_rn(xb[0][kk][t], wb[0][kk][v], acc[v][t]);
                x += 8 * XS;
                w += 8 * NB;
                load(0, x, w);   // (reads past the end on the last block: harmless here)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = __fmaf_rn(xb[1][kk][t], wb[1][kk][v], acc[v][t]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int t = 0; t < TQ; ++t) s += acc[v][t];
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 16;
    if (s == 0.1234567f) out[1] = 1;
}

// heads: one 64-long dependent FMA chain per lane, operands from shared memory (x stride XS, w stride HP)
template <int XS, int HP, int PIPE>
__global__ void k_heads(long long* out, int H) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1.0f / (1 + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const float* xp = sm + (lane & 15);
    const float* wp = sm + 4096 + (lane >> 4);
    float acc = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < 16; ++rep) {
        if constexpr (PIPE == 0) {
#pragma unroll 8
            for (int k = 0; k < H; ++k) acc = __fmaf_rn(xp[k * XS], wp[k * HP], acc);
        } else {
            float xb[2][8], wb[2][8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) { xb[0][kk] = xp[kk * XS]; wb[0][kk] = wp[kk * HP]; }
#pragma unroll 1
            for (int k0 = 0; k0 < H; k0 += 16) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) { xb[1][kk] = xp[(k0 + 8 + kk) * XS]; wb[1][kk] = wp[(k0 + 8 + kk) * HP]; }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) acc = __fmaf_rn(xb[0][kk], wb[0][kk], acc);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) { xb[0][kk] = xp[(k0 + 16 + kk) * XS]; wb[0][kk] = wp[(k0 + 16 + kk) * HP]; }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) acc = __fmaf_rn(xb[1][kk], wb[1][kk], acc);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 16;
    if (acc == 0.1234567f) out[1] = 1;
}

static long long run(long long* d_out, const char* name, double per, void (*launch)()) {
    cudaMemset(d_out, 0, 64);
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
    printf("%-44s %10lld cycles  %8.2f per op   (%s)\n", name, h[0], (double)h[0] / per, cudaGetErrorString(e));
    return h[0];
}

static long long* g_out;
static int4* g_buf;

int main() {
    cudaMalloc(&g_out, 64);
    cudaMalloc(&g_buf, 1 << 20);
    // chase table: 64 entries (fits L1), stride pattern
    {
        int4 h[4096];
        for (int i = 0; i < 4096; ++i) h[i] = make_int4((i * 37 + 11) % 64, 0, 0, 0);
        cudaMemcpy(g_buf, h, sizeof(h), cudaMemcpyHostToDevice);
    }
    long long* o = g_out;
#define RUN(name, per, ...)                                          \
    {                                                                \
        cudaMemset(o, 0, 64);                                        \
        __VA_ARGS__;                                                 \
        cudaError_t e = cudaDeviceSynchronize();                     \
        long long h[4];                                              \
        cudaMemcpy(h, o, 32, cudaMemcpyDeviceToHost);                \
        printf("%-52s %9lld cyc %8.2f /op  aux %8.2f (%s)\n", name, h[0], (double)h[0] / (per), (double)h[2] / (per), cudaGetErrorString(e)); \
    }
    RUN("FFMA2 1 chain, 1 warp (latency)", ITER, k_ffma2<1><<<1, 32>>>(o, 3))
    RUN("FFMA2 2 chains, 1 warp", ITER * 2, k_ffma2<2><<<1, 32>>>(o, 3))
    RUN("FFMA2 4 chains, 1 warp", ITER * 4, k_ffma2<4><<<1, 32>>>(o, 3))
    RUN("FFMA2 8 chains, 1 warp", ITER * 8, k_ffma2<8><<<1, 32>>>(o, 3))
    RUN("FFMA2 4 chains, 4 warps (1/SMSP)", ITER * 4, k_ffma2<4><<<1, 128>>>(o, 3))
    RUN("FFMA2 4 chains, 8 warps (2/SMSP)", ITER * 4, k_ffma2<4><<<1, 256>>>(o, 3))
    RUN("FFMA2 4 chains, 16 warps (4/SMSP)", ITER * 4, k_ffma2<4><<<1, 512>>>(o, 3))
    RUN("FFMA 1 chain, 1 warp (latency)", ITER, k_ffma<1><<<1, 32>>>(o, 3.f))
    RUN("FFMA 8 chains, 1 warp", ITER * 8, k_ffma<8><<<1, 32>>>(o, 3.f))
    RUN("FFMA 8 chains, 8 warps (2/SMSP)", ITER * 8, k_ffma<8><<<1, 256>>>(o, 3.f))
    RUN("FFMA 8 chains, 16 warps (4/SMSP)", ITER * 8, k_ffma<8><<<1, 512>>>(o, 3.f))
    RUN("DFMA 1 chain (latency)", ITER, k_dfma<1><<<1, 32>>>(o, 1.5))
    RUN("DFMA 4 chains, 1 warp", ITER * 4, k_dfma<4><<<1, 32>>>(o, 1.5))
    RUN("DFMA 1 chain, 1 lane active", ITER, k_dfma<1><<<1, 1>>>(o, 1.5))
    RUN("DFMA 4 chains, 8 warps", ITER * 4, k_dfma<4><<<1, 256>>>(o, 1.5))
    RUN("DADD chain (latency)", ITER, k_dadd<<<1, 32>>>(o, 1.5))
    RUN("DMUL chain (latency)", ITER, k_dmul<<<1, 32>>>(o, 1.5))
    RUN("I2D+D2I chain (pair)", ITER, k_i2d<<<1, 32>>>(o, 5))
    RUN("F2D+D2F chain (pair)", ITER, k_f2d<<<1, 32>>>(o, 1.5f))
    RUN("DSETP+SEL chain (per pair)", ITER, k_dsetp<<<1, 32>>>(o, 1.5))
    RUN("LDS chase 32-bit (aux: 128-bit + 2 ALU)", ITER, k_lds<<<1, 32>>>(o))
    RUN("LDG.128 chase, L1 hit", ITER, k_ldg<<<1, 32>>>(o, g_buf, 64))
    RUN("LDG.128 chase, 1 lane", ITER, k_ldg<<<1, 1>>>(o, g_buf, 64))
    RUN("LDG.128.cg chase (L2 hit)", ITER, k_ldg_cg<<<1, 32>>>(o, g_buf, 64))
    RUN("STG.128 -> LDG.128 same address (volatile)", ITER, k_stld<<<1, 32>>>(o, g_buf + 1024))
    RUN("SHFL.BFLY + IADD chain", ITER, k_shfl<<<1, 32>>>(o, 1))
    RUN("VOTE.ANY + IADD chain", ITER, k_vote<<<1, 32>>>(o, 1))
    RUN("bar.sync 1, 160 (5 warps)", ITER, k_bar<<<1, 160>>>(o, 160))
    RUN("bar.sync 1, 128 (4 warps)", ITER, k_bar<<<1, 128>>>(o, 128))
    cudaFuncSetAttribute(k_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    RUN("team layer loop H=64: 1 warp (per layer)", 64, k_layer<<<1, 32, 65536>>>(o, 20, 128, 64))
    RUN("team layer loop H=64: 4 warps (1/SMSP)", 64, k_layer<<<1, 128, 65536>>>(o, 20, 128, 64))
    RUN("team layer loop H=64: 8 warps (2/SMSP)", 64, k_layer<<<1, 256, 65536>>>(o, 20, 128, 64))
    RUN("team layer loop H=64: 16 warps (4/SMSP)", 64, k_layer<<<1, 512, 65536>>>(o, 20, 128, 64))

#define LAYER(TQ, NV, NB, PIPE, W)                                                                               \
    {                                                                                                             \
        cudaFuncSetAttribute(k_layer_s<TQ, NV, 20, NB, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); \
        char nm[128];                                                                                             \
        snprintf(nm, sizeof nm, "scalar layer %dt x %dn NB=%d pipe=%d: %d warps", TQ, NV, NB, PIPE, W);           \
        RUN(nm, 64, k_layer_s<TQ, NV, 20, NB, PIPE><<<1, 32 * W, 65536>>>(o, 64))                               \
    }
    LAYER(4, 2, 16, 0, 1) LAYER(4, 2, 16, 0, 4) LAYER(4, 2, 16, 0, 8) LAYER(4, 2, 16, 0, 16)
    LAYER(4, 2, 16, 1, 1) LAYER(4, 2, 16, 1, 4) LAYER(4, 2, 16, 1, 8) LAYER(4, 2, 16, 1, 16)
    LAYER(4, 4, 32, 0, 1) LAYER(4, 4, 32, 0, 4) LAYER(4, 4, 32, 0, 8)
    LAYER(4, 4, 32, 1, 1) LAYER(4, 4, 32, 1, 4) LAYER(4, 4, 32, 1, 8)
    LAYER(2, 2, 8, 0, 1) LAYER(2, 2, 8, 0, 8) LAYER(2, 2, 8, 1, 1) LAYER(2, 2, 8, 1, 8) LAYER(2, 2, 8, 1, 16)
    LAYER(4, 1, 8, 1, 1) LAYER(4, 1, 8, 1, 8) LAYER(4, 1, 8, 1, 16)
    LAYER(2, 4, 16, 1, 1) LAYER(2, 4, 16, 1, 8)
    cudaFuncSetAttribute(k_heads<20, 4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_heads<20, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    RUN("heads chain 64 (unroll 8): 1 warp, per chain", 1, k_heads<20, 4, 0><<<1, 32, 65536>>>(o, 64))
    RUN("heads chain 64 (double buffered): 1 warp", 1, k_heads<20, 4, 1><<<1, 32, 65536>>>(o, 64))
    RUN("heads chain 64 (double buffered): 4 warps", 1, k_heads<20, 4, 1><<<1, 128, 65536>>>(o, 64))
    int dev = 0, clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("SM clock (attr) %d kHz\n", clk);
    return 0;
}
