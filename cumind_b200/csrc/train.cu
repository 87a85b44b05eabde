// train.cu — optimiser kernel of the train step.
//
// optax.adamw(learning_rate, weight_decay) as the reference builds it (src/cumind/agent/agent.py:48;
// b1=0.9, b2=0.999, eps=1e-8, decoupled weight decay on every trainable leaf):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  m^ = m/(1-b1^t);  v^ = v/(1-b2^t)
//   p = p - lr * ( m^ / (sqrt(v^) + eps) + wd * p )
// over the flat fp32 parameter blob (layout of include/cumind_b200.h), fused with the 1/world scaling
// of the all-reduced gradient.  `mask` is 0 for leaves the reference never trains (BatchNorm running
// mean / var are nnx.BatchStat, trainer.py:92 only takes nnx.Param) and 1 elsewhere.
#include "kernels.h"

namespace cm {

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
