// conv_fp32.cu — image encoder in the fp32-exact (parity) mode.
//
//   ConvEncoder.__call__   network.py:131-147   relu(conv3x3 s2) -> L x ResidualBlock -> mean(H,W) -> Linear
//   ResidualBlock.__call__ network.py:29-45     conv -> BN -> relu -> conv -> BN -> (+x) -> relu
// NHWC, channels = obs_shape[-1] (network.py:127).  Pinned arithmetic (oracle/cumind_oracle.c):
// taps in (ky,kx,ci) order, taps in the zero padding skipped, fp32 fused multiply-add chain, then
// + bias, then the BatchNorm / residual / ReLU epilogue element-wise.  Runs once per search.
//
// Tiling: one CTA = 8x8 output pixels of one image, 256 threads = 64 pixels x 4 channel groups of
// 8 output channels; the input patch is staged channel-major in shared memory (conflict-free
// pixel-contiguous reads), weights are warp-uniform float4 loads served by L1.
#include "kernels.h"
#include "search_core.cuh"

namespace cm {

static constexpr int TILE = 8;

struct ConvArgs {
    const float* in;    // [n][Hi][Wi][Cin]
    float* out;         // [n][Ho][Wo][Cout]
    const float* K;     // [3][3][Cin][Cout]
    const float* bias;  // [Cout]
    const float* bn;    // scale,bias,mean,var [4][Cout] or nullptr
    const float* res;   // residual [n][Ho][Wo][Cout] or nullptr
    int Hi, Wi, Cin, Ho, Wo, Cout, stride, relu;
};

__global__ void __launch_bounds__(256) conv3x3_kernel(const ConvArgs a) {
    extern __shared__ __align__(16) float tile[];  // [Cin][TH][TW]
    const int TIN = (TILE - 1) * a.stride + 3;
    const int TW = TIN + 1;  // pad the row to break bank periodicity
    const int tiles_x = (a.Wo + TILE - 1) / TILE;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x % tiles_x;
    const int img = blockIdx.y;
    const int oy0 = ty * TILE, ox0 = tx * TILE;
    const int iy0 = oy0 * a.stride - 1, ix0 = ox0 * a.stride - 1;
    const float* in = a.in + (size_t)img * a.Hi * a.Wi * a.Cin;

    // stage the input patch: global reads are channel-contiguous (coalesced), smem writes channel-major
    const int n_in = TIN * TIN * a.Cin;
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
        const int ci = i % a.Cin;
        const int pix = i / a.Cin;
        const int py = pix / TIN, px = pix % TIN;
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.0f;
        if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) v = in[((size_t)iy * a.Wi + ix) * a.Cin + ci];
        tile[((size_t)ci * TIN + py) * TW + px] = v;
    }
    __syncthreads();

    const int p = threadIdx.x & 63;
    const int py = p >> 3, px = p & 7;
    const int oy = oy0 + py, ox = ox0 + px;
    const bool pix_ok = oy < a.Ho && ox < a.Wo;
    for (int co0 = (threadIdx.x >> 6) * 8; co0 < a.Cout; co0 += 32) {
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
        const bool full8 = (co0 + 8 <= a.Cout) && ((a.Cout & 3) == 0);
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * a.stride - 1 + ky;
            if (iy < 0 || iy >= a.Hi) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * a.stride - 1 + kx;
                if (ix < 0 || ix >= a.Wi) continue;
                const float* tp = tile + (size_t)(py * a.stride + ky) * TW + (px * a.stride + kx);
                const float* kp = a.K + (size_t)((ky * 3 + kx) * a.Cin) * a.Cout + co0;
                if (full8) {
                    for (int ci = 0; ci < a.Cin; ++ci) {
                        const float v = tp[(size_t)ci * TIN * TW];
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(kp + (size_t)ci * a.Cout));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(kp + (size_t)ci * a.Cout + 4));
                        acc[0] = ffma(v, w0.x, acc[0]); acc[1] = ffma(v, w0.y, acc[1]);
                        acc[2] = ffma(v, w0.z, acc[2]); acc[3] = ffma(v, w0.w, acc[3]);
                        acc[4] = ffma(v, w1.x, acc[4]); acc[5] = ffma(v, w1.y, acc[5]);
                        acc[6] = ffma(v, w1.z, acc[6]); acc[7] = ffma(v, w1.w, acc[7]);
                    }
                } else {
                    for (int ci = 0; ci < a.Cin; ++ci) {
                        const float v = tp[(size_t)ci * TIN * TW];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (co0 + c < a.Cout) acc[c] = ffma(v, __ldg(kp + (size_t)ci * a.Cout + c), acc[c]);
                    }
                }
            }
        }
        if (!pix_ok) continue;
        const size_t obase = (((size_t)img * a.Ho + oy) * a.Wo + ox) * a.Cout + co0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (co0 + c >= a.Cout) break;
            float v = fadd(acc[c], a.bias[co0 + c]);
            if (a.bn) {
                const int ch = co0 + c;
                const float scale = a.bn[ch], beta = a.bn[a.Cout + ch], mean = a.bn[2 * a.Cout + ch], var = a.bn[3 * a.Cout + ch];
                const float gain = fmul(scale, __fdiv_rn(1.0f, __fsqrt_rn(fadd(var, 1e-5f))));
                v = fadd(fmul(__fsub_rn(v, mean), gain), beta);
            }
            if (a.res) v = fadd(v, a.res[obase + c]);
            if (a.relu) v = frelu(v);
            a.out[obase + c] = v;
        }
    }
}

// jnp.mean over (H,W) (sequential row-major sum, / count) then final_dense (network.py:146-147).
__global__ void pool_dense_kernel(const float* __restrict__ x, int npix, int Cc, const float* __restrict__ K,
                                  const float* __restrict__ b, int H, float* __restrict__ hidden) {
    extern __shared__ __align__(16) float pooled[];
    const int img = blockIdx.x;
    const float* xi = x + (size_t)img * npix * Cc;
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
        float s = 0.0f;
        for (int i = 0; i < npix; ++i) s = fadd(s, xi[(size_t)i * Cc + c]);
        pooled[c] = __fdiv_rn(s, (float)npix);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < Cc; ++k) acc = ffma(pooled[k], K[(size_t)k * H + j], acc);
        hidden[(size_t)img * H + j] = fadd(acc, b[j]);
    }
}

static inline int out_dim(int n) { return (n - 1) / 2 + 1; }

size_t conv_encoder_work_bytes(const cumind_net_desc& d, int chunk) {
    const size_t npix = (size_t)out_dim(d.obs_shape[0]) * out_dim(d.obs_shape[1]);
    return 3 * (size_t)chunk * npix * d.conv_channels * sizeof(float);
}

static cudaError_t run_conv(const ConvArgs& a, int n, cudaStream_t st) {
    const int TIN = (TILE - 1) * a.stride + 3;
    const size_t smem = (size_t)a.Cin * TIN * (TIN + 1) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int tiles = ((a.Ho + TILE - 1) / TILE) * ((a.Wo + TILE - 1) / TILE);
    conv3x3_kernel<<<dim3(tiles, n), 256, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_conv_encoder(const NetView& nv, const float* obs, int B, float* hidden, float* work, int chunk, cudaStream_t st) {
    const cumind_net_desc& d = nv.desc;
    const int Hi = d.obs_shape[0], Wi = d.obs_shape[1], C = d.obs_shape[2], Cc = d.conv_channels, L = d.num_blocks, H = d.hidden_dim;
    const int Ho = out_dim(Hi), Wo = out_dim(Wi);
    const size_t npix = (size_t)Ho * Wo;
    const size_t buf = (size_t)chunk * npix * Cc;
    float* x = work;
    float* t = work + buf;
    float* u = work + 2 * buf;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int n = B - b0 < chunk ? B - b0 : chunk;
        const float* p = nv.blob + nv.lo.enc;
        ConvArgs a{};
        a.in = obs + (size_t)b0 * Hi * Wi * C; a.out = x; a.K = p; a.bias = p + (size_t)9 * C * Cc; a.bn = nullptr; a.res = nullptr;
        a.Hi = Hi; a.Wi = Wi; a.Cin = C; a.Ho = Ho; a.Wo = Wo; a.Cout = Cc; a.stride = 2; a.relu = 1;
        cudaError_t e = run_conv(a, n, st);
        if (e != cudaSuccess) return e;
        p += (size_t)9 * C * Cc + Cc;
        for (int l = 0; l < L; ++l) {
            const float* k1 = p;           const float* b1 = k1 + (size_t)9 * Cc * Cc;
            const float* bn1 = b1 + Cc;    const float* k2 = bn1 + 4 * Cc;
            const float* b2 = k2 + (size_t)9 * Cc * Cc;
            const float* bn2 = b2 + Cc;
            ConvArgs c1{};
            c1.in = x; c1.out = t; c1.K = k1; c1.bias = b1; c1.bn = bn1; c1.res = nullptr;
            c1.Hi = Ho; c1.Wi = Wo; c1.Cin = Cc; c1.Ho = Ho; c1.Wo = Wo; c1.Cout = Cc; c1.stride = 1; c1.relu = 1;
            if ((e = run_conv(c1, n, st)) != cudaSuccess) return e;
            ConvArgs c2 = c1;
            c2.in = t; c2.out = u; c2.K = k2; c2.bias = b2; c2.bn = bn2; c2.res = x;
            if ((e = run_conv(c2, n, st)) != cudaSuccess) return e;
            float* tmp = x; x = u; u = tmp;
            p = bn2 + 4 * Cc;
        }
        pool_dense_kernel<<<n, 128, Cc * sizeof(float), st>>>(x, (int)npix, Cc, p, p + (size_t)Cc * H, H, hidden + (size_t)b0 * H);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace cm
