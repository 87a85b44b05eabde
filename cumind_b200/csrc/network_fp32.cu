// network_fp32.cu — batched network entry points in the fp32-exact (parity) mode, one warp per row.
//
//   vector_encoder_kernel <- VectorEncoder.__call__      network.py:94-108
//   recurrent_kernel      <- CuMindNetwork.recurrent_inference  network.py:302-315
//                            (DynamicsNetwork.__call__ 217-236 + PredictionNetwork.__call__ 254-266)
//   prediction_kernel     <- PredictionNetwork.__call__  network.py:254-266
//   softmax_kernel        <- jax.nn.softmax              mcts.py:128,191
// Arithmetic: sequential-k fp32 FMA-chain dot products (see common.cuh); bit-identical to oracle/.
#include "kernels.h"
#include "search_core.cuh"

namespace cm {

static constexpr int kWarps = 4;

CM_DEVICE void linear_any(const float* __restrict__ K, const float* __restrict__ b, const float* __restrict__ x, int in,
                          int out, float* __restrict__ y, bool relu, int lane) {
    for (int j = lane; j < out; j += 32) {
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < in; ++k) acc = ffma(x[k], K[(size_t)k * out + j], acc);
        const float v = fadd(acc, b[j]);
        y[j] = relu ? frelu(v) : v;
    }
}

// VectorEncoder for 4 rows per warp with the fused kernel's FFMA2 tile: all layer parameters staged in
// shared memory once per CTA (row-major [in][H]), activations k-major with the 4 rows interleaved, lane
// owns R consecutive neurons of every row; persistent CTAs loop over row quads.  Same sequential-k FMA
// chain per (row, neuron) as cmo_linear, so the result is bit-identical to the warp-per-row kernel.
template <int R>
__global__ void __launch_bounds__(256) vector_encoder_tile_kernel(const float* __restrict__ enc, int n_enc, int D, int H, int L,
                                                                   const float* __restrict__ obs, int B, float* __restrict__ hidden) {
    constexpr int TW = 4;
    extern __shared__ __align__(16) float sm[];
    float* w_s = sm;                                           // encoder parameters
    const int M = D > H ? D : H;
    float* xbuf = sm + ((n_enc + 3) & ~3) + (size_t)(threadIdx.x >> 5) * 2 * M * TW;
    pdl_launch_dependents();  // the search kernel behind this one may stage its parameters meanwhile
    for (int i = threadIdx.x; i < n_enc; i += blockDim.x) w_s[i] = enc[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int j0 = lane * R;
    const int n_quads = (B + TW - 1) / TW;
    for (int q = blockIdx.x * nwarps + warp; q < n_quads; q += gridDim.x * nwarps) {
        float* a = xbuf;
        float* c = xbuf + (size_t)M * TW;
        for (int idx = lane; idx < D * TW; idx += 32) {          // a[k][t] = obs[row t][k]
            const int k = idx / TW, t = idx - k * TW;
            const int r = q * TW + t;
            a[idx] = r < B ? obs[(size_t)r * D + k] : 0.0f;
        }
        __syncwarp();
        const float* p = w_s;
        int in = D;
        for (int i = 0; i < L; ++i) {
            if (j0 < H) {
                WarpAcc<TW, R> acc;
                acc.zero();
#pragma unroll 4
                for (int k = 0; k < in; ++k) {
                    float xk[TW], w[R];
                    load_vec<TW>(a + (size_t)k * TW, xk);
                    load_vec<R>(p + (size_t)k * H + j0, w);
                    if constexpr (R >= 2) {
#pragma unroll
                        for (int t = 0; t < TW; ++t)
#pragma unroll
                            for (int v = 0; v < R; v += 2)
                                acc.v[t * (R / 2) + v / 2] = ffma2(xk[t], make_float2(w[v], w[v + 1]), acc.v[t * (R / 2) + v / 2]);
                    } else {
#pragma unroll
                        for (int t = 0; t < TW; t += 2) acc.v[t / 2] = ffma2(w[0], make_float2(xk[t], xk[t + 1]), acc.v[t / 2]);
                    }
                }
                const float* bias = p + (size_t)in * H;
#pragma unroll
                for (int v = 0; v < R; ++v) {
                    float y[TW];
#pragma unroll
                    for (int t = 0; t < TW; ++t) {
                        const float z = fadd(acc.get(t, v), bias[j0 + v]);
                        y[t] = (i < L - 1) ? frelu(z) : z;
                    }
                    store_vec<TW>(c + (size_t)(j0 + v) * TW, y);
                }
            }
            __syncwarp();
            p += (size_t)in * H + H;
            in = H;
            float* tmp = a; a = c; c = tmp;
        }
        for (int idx = lane; idx < H * TW; idx += 32) {          // hidden[row t][j] = a[j][t]
            const int t = idx / H, j = idx - t * H;
            const int r = q * TW + t;
            if (r < B) hidden[(size_t)r * H + j] = a[(size_t)j * TW + t];
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kWarps * 32) vector_encoder_kernel(const float* __restrict__ enc, int D, int H, int L,
                                                                    const float* __restrict__ obs, int B,
                                                                    float* __restrict__ hidden) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int M = D > H ? D : H;
    float* a = sm + (size_t)w * 2 * M;
    float* c = a + M;
    pdl_launch_dependents();
    const int r = blockIdx.x * kWarps + w;
    const bool active = r < B;
    if (active)
        for (int j = lane; j < D; j += 32) a[j] = obs[(size_t)r * D + j];
    __syncwarp();
    const float* p = enc;
    int in = D;
    for (int i = 0; i < L; ++i) {
        if (active) linear_any(p, p + (size_t)in * H, a, in, H, c, i < L - 1, lane);
        __syncwarp();
        p += (size_t)in * H + H;
        in = H;
        float* t = a; a = c; c = t;
    }
    if (active)
        for (int j = lane; j < H; j += 32) hidden[(size_t)r * H + j] = a[j];
}

__global__ void __launch_bounds__(kWarps * 32) recurrent_kernel(const float* __restrict__ pack, const PackLayout pk,
                                                               const float* __restrict__ hidden,
                                                               const int32_t* __restrict__ action, int B,
                                                               float* __restrict__ next, float* __restrict__ reward,
                                                               float* __restrict__ logits, float* __restrict__ value,
                                                               int with_dynamics) {
    extern __shared__ __align__(16) float sm[];
    const int H = pk.H, L = pk.L, HP = pk.HP, A = pk.A;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* x0 = sm + (size_t)w * (2 * H + HP);
    float* x1 = x0 + H;
    float* lg = x1 + H;
    const int r = blockIdx.x * kWarps + w;
    const bool active = r < B;
    if (active) {
        const float* h = hidden + (size_t)r * H;
        if (with_dynamics) {
            int a = action[r];
            a = a < 0 ? 0 : (a >= A ? A - 1 : a);  // jnp.take clamps out-of-range indices
            const float* emb = pack + pk.emb + (size_t)a * H;
            for (int j = lane; j < H; j += 32) x0[j] = fadd(h[j], emb[j]);
        } else {
            for (int j = lane; j < H; j += 32) x0[j] = h[j];
        }
    }
    __syncwarp();
    float* cur = x0;
    float* nxt = x1;
    if (with_dynamics) {
        for (int l = 0; l < L; ++l) {
            if (active) {
                const float* K = pack + pk.layer_k(l);
                const float* b = pack + pk.layer_b(l);
                for (int j = lane; j < H; j += 32) {
                    float acc = 0.0f;
#pragma unroll 8
                    for (int k = 0; k < H; ++k) acc = ffma(cur[k], K[(size_t)k * H + j], acc);
                    nxt[j] = fadd(frelu(fadd(acc, b[j])), cur[j]);
                }
            }
            __syncwarp();
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    if (active) {
        if (next)
            for (int j = lane; j < H; j += 32) next[(size_t)r * H + j] = cur[j];
        heads<32>(pack + pk.heads_w, pack + pk.heads_b, cur, H, HP, A + 2, lg, lane);
    }
    __syncwarp();
    if (active) {
        if (logits)
            for (int a = lane; a < A; a += 32) logits[(size_t)r * A + a] = lg[a];
        if (lane == 0) {
            if (value) value[r] = lg[A];
            if (reward) reward[r] = lg[A + 1];
        }
    }
}

__global__ void softmax_kernel(const float* __restrict__ logits, int B, int A, float* __restrict__ probs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    const float* l = logits + (size_t)r * A;
    float* o = probs + (size_t)r * A;
    float m = l[0];
    for (int a = 1; a < A; ++a) if (l[a] > m) m = l[a];
    float s = 0.0f;
    for (int a = 0; a < A; ++a) {
        const float e = __double2float_rn(pexp((double)__fsub_rn(l[a], m)));
        o[a] = e;
        s = fadd(s, e);
    }
    for (int a = 0; a < A; ++a) o[a] = __fdiv_rn(o[a], s);
}

static cudaError_t set_smem(const void* fn, size_t smem) {
    if (smem > 48 * 1024) return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaSuccess;
}

template <int R>
static cudaError_t launch_encoder_tile(const NetView& nv, const float* obs, int B, float* hidden, size_t smem, int n_enc, cudaStream_t st) {
    const int D = nv.desc.obs_shape[0], H = nv.desc.hidden_dim, L = nv.desc.num_blocks;
    cudaError_t e = cudaFuncSetAttribute(vector_encoder_tile_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int quads = (B + 3) / 4;
    int grid = (quads + 7) / 8;
    if (grid > 148 * 2) grid = 148 * 2;
    vector_encoder_tile_kernel<R><<<grid, 256, smem, st>>>(nv.blob + nv.lo.enc, n_enc, D, H, L, obs, B, hidden);
    return cudaGetLastError();
}

cudaError_t launch_vector_encoder(const NetView& nv, const float* obs, int B, float* hidden, cudaStream_t st) {
    const int D = nv.desc.obs_shape[0], H = nv.desc.hidden_dim, L = nv.desc.num_blocks;
    const int M = D > H ? D : H;
    {   // tile kernel when the parameters fit shared memory and H maps onto R neurons per lane
        const int n_enc = (int)((size_t)D * H + H + (size_t)(L - 1) * ((size_t)H * H + H));
        const size_t smem_t = ((size_t)((n_enc + 3) & ~3) + (size_t)8 * 2 * M * 4) * sizeof(float);
        const int R = pick_neurons_per_lane(H);
        if (R > 0 && smem_t <= 200 * 1024 && B >= 64) {
            switch (R) {
                case 1: return launch_encoder_tile<1>(nv, obs, B, hidden, smem_t, n_enc, st);
                case 2: return launch_encoder_tile<2>(nv, obs, B, hidden, smem_t, n_enc, st);
                case 4: return launch_encoder_tile<4>(nv, obs, B, hidden, smem_t, n_enc, st);
                case 8: return launch_encoder_tile<8>(nv, obs, B, hidden, smem_t, n_enc, st);
            }
        }
    }
    const size_t smem = (size_t)kWarps * 2 * M * sizeof(float);
    cudaError_t e = set_smem((const void*)vector_encoder_kernel, smem);
    if (e != cudaSuccess) return e;
    vector_encoder_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(nv.blob + nv.lo.enc, D, H, L, obs, B, hidden);
    return cudaGetLastError();
}

cudaError_t launch_recurrent(const NetView& nv, const float* hidden, const int32_t* action, int B, float* next, float* reward,
                             float* logits, float* value, cudaStream_t st) {
    const size_t smem = (size_t)kWarps * (2 * nv.pk.H + nv.pk.HP) * sizeof(float);
    cudaError_t e = set_smem((const void*)recurrent_kernel, smem);
    if (e != cudaSuccess) return e;
    recurrent_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(nv.pack, nv.pk, hidden, action, B, next, reward, logits,
                                                                        value, 1);
    return cudaGetLastError();
}

cudaError_t launch_prediction(const NetView& nv, const float* hidden, int B, float* logits, float* value, cudaStream_t st) {
    const size_t smem = (size_t)kWarps * (2 * nv.pk.H + nv.pk.HP) * sizeof(float);
    cudaError_t e = set_smem((const void*)recurrent_kernel, smem);
    if (e != cudaSuccess) return e;
    recurrent_kernel<<<(B + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(nv.pack, nv.pk, hidden, nullptr, B, nullptr, nullptr,
                                                                        logits, value, 0);
    return cudaGetLastError();
}

cudaError_t launch_softmax(const float* logits, int B, int A, float* probs, cudaStream_t st) {
    softmax_kernel<<<(B + 127) / 128, 128, 0, st>>>(logits, B, A, probs);
    return cudaGetLastError();
}

}  // namespace cm
