// train.cu — optimiser kernel of the train step.
//
// optax.adamw(learning_rate, weight_decay) as the reference builds it (src/cumind/agent/agent.py:48;
// b1=0.9, b2=0.999, eps=1e-8, decoupled weight decay on every trainable leaf):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  m^ = m/(1-b1^t);  v^ = v/(1-b2^t)
//   p = p - lr * ( m^ / (sqrt(v^) + eps) + wd * p )
// over the flat fp32 parameter blob (layout of include/cumind_b200.h), fused with the 1/world scaling
// of the all-reduced gradient.  `mask` is 0 for leaves the reference never trains (BatchNorm running
// mean / var are nnx.BatchStat, trainer.py:92 only takes nnx.Param) and 1 elsewhere.
#include <cuda_bf16.h>

#include "kernels.h"

namespace cm {

// ---- kernel-side parameter layouts from the canonical blob (maps built once on the host, abi.cu) ----
__global__ void repack_f32_kernel(const float* __restrict__ blob, const int* __restrict__ map, float* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int m = map[i];
        if (m >= 0) dst[i] = blob[m];
        else if (m == -1) dst[i] = 0.0f;
    }
}

__global__ void repack_bf16_kernel(const float* __restrict__ blob, const int* __restrict__ map, __nv_bfloat16* __restrict__ dst,
                                   size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int m = map[i];
        if (m >= 0) dst[i] = __float2bfloat16_rn(blob[m]);
        else if (m == -1) dst[i] = __float2bfloat16_rn(0.0f);
    }
}

// conv bias + inference BatchNorm -> y = acc*gain + shift (network.py:25-27; eps = 1e-5)
__global__ void bn_fold_kernel(const float* __restrict__ blob, const BnFold* __restrict__ folds, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BnFold f = folds[i];
    if (f.scale >= 0) {
        const float g = blob[f.scale] / sqrtf(blob[f.var] + 1e-5f);
        out[f.gain_dst] = g;
        out[f.shift_dst] = (blob[f.bias] - blob[f.mean]) * g + blob[f.beta];
    } else {
        out[f.gain_dst] = 1.0f;
        out[f.shift_dst] = blob[f.bias];
    }
}

static inline int grid_for(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}
cudaError_t launch_repack_f32(const float* blob, const int* map, float* dst, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    repack_f32_kernel<<<grid_for(n), 256, 0, st>>>(blob, map, dst, n);
    return cudaGetLastError();
}
cudaError_t launch_repack_bf16(const float* blob, const int* map, unsigned char* dst, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    repack_bf16_kernel<<<grid_for(n), 256, 0, st>>>(blob, map, reinterpret_cast<__nv_bfloat16*>(dst), n);
    return cudaGetLastError();
}
cudaError_t launch_bn_fold(const float* blob, const BnFold* folds, int n, float* cpack_f32, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    bn_fold_kernel<<<(n + 127) / 128, 128, 0, st>>>(blob, folds, n, cpack_f32);
    return cudaGetLastError();
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             const unsigned char* __restrict__ mask, size_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2, float grad_scale) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask && !mask[i]) continue;
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float mh = mi / bc1, vh = vi / bc2;
        const float pi = p[i];
        p[i] = pi - lr * (mh / (sqrtf(vh) + eps) + wd * pi);
    }
}

// graph-friendly variant: the step counter lives on the device (incremented by a 1-thread kernel first), so
// the same captured launch computes the right bias corrections on every replay
__global__ void incr_step_kernel(long long* step) { *step += 1; }

__global__ void adamw_dev_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                      const unsigned char* __restrict__ mask, size_t n, float lr, float b1, float b2, float eps,
                                      float wd, const long long* __restrict__ step, float grad_scale) {
    const float t = (float)(*step);
    const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask && !mask[i]) continue;
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pi = p[i];
        p[i] = pi - lr * ((mi / bc1) / (sqrtf(vi / bc2) + eps) + wd * pi);
    }
}

cudaError_t launch_adamw_dev_step(float* p, const float* g, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                                  float b2, float eps, float wd, long long* d_step, float grad_scale, cudaStream_t st) {
    incr_step_kernel<<<1, 1, 0, st>>>(d_step);
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_dev_step_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, mask, n, lr, b1, b2, eps, wd, d_step, grad_scale);
    return cudaGetLastError();
}

cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                         float b2, float eps, float wd, long long step, float grad_scale, cudaStream_t st) {
    const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
    const int block = 256;
    size_t blocks = (n + block - 1) / block;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adamw_kernel<<<(int)blocks, block, 0, st>>>(p, g, m, v, mask, n, lr, b1, b2, eps, wd, bc1, bc2, grad_scale);
    return cudaGetLastError();
}

}  // namespace cm
