// train_fused.cu — loss and gradient of the MuZero train step in ONE kernel (vector observations).
//
//   Trainer.compute_losses   src/cumind/agent/trainer.py:172-205   K-step unroll of initial_inference / recurrent_inference
//                            (network.py:288-315), value MSE + policy cross-entropy at every one of the K+1 states against the
//                            SAME targets, reward MSE at the K unrolled states; losses divided by K+1 / K+1 / K
//   nnx.value_and_grad       trainer.py:88-95                       the gradient of value + policy + reward loss w.r.t. every parameter
//
// A CTA owns R = 8 trajectories of the replay batch and keeps everything the backward pass needs on chip: all parameters
// (row stride H+1 so that both `x @ W` — lanes along the output neuron — and `dy @ W^T` — lanes along the input — read
// shared memory without bank conflicts), the activations of every layer of every unroll step, the ReLU masks and the head
// outputs.  Thread (j, g) owns output neuron j for the two trajectories of row group g (4 groups, 4 H threads).  The weight
// gradients of the CTA's 8 trajectories are accumulated in registers over the whole unroll (thread (j, g) owns dW[k][j] for a
// quarter of the k), written once as a partial gradient per CTA, and summed over the CTAs in a fixed order by a second
// kernel — the result does not depend on scheduling (no atomics).  fp32 arithmetic; checked against the float64 restatement
// oracle/train_np.py within 1e-4 (tests/test_train.py).
#include "kernels.h"

namespace cm {

namespace {

constexpr int TR = 8;      // trajectories per CTA
constexpr int TG = 4;      // row groups
constexpr int TRG = TR / TG;
constexpr int TKP = 16;    // weight-gradient entries per thread and H x H layer (H / TG for H = 64)

struct TrainSmem {
    // float offsets into dynamic shared memory
    int enc_w[8], enc_b[8], emb, dyn_w[8], dyn_b[8], heads_w, heads_b;
    int obs, ye[8], hst, x, outs, dA, dB, dout, demb, red;
    int mask_bytes_off;   // byte offset of the mask region
    int total_floats;
    int nop;              // padded row stride of the heads matrix (odd)
};

__host__ __device__ inline TrainSmem train_smem_layout(int D, int H, int L, int A, int K) {
    TrainSmem s;
    const int NO = A + 2;
    int o = 0;
    auto take = [&](int n) { const int r = o; o += (n + 3) & ~3; return r; };
    int in = D;
    for (int i = 0; i < L; ++i) { s.enc_w[i] = take(in * (H + 1)); s.enc_b[i] = take(H); in = H; }
    s.emb = take(A * H);
    for (int l = 0; l < L; ++l) { s.dyn_w[l] = take(H * (H + 1)); s.dyn_b[l] = take(H); }
    s.nop = NO | 1;
    s.heads_w = take(H * s.nop);
    s.heads_b = take(NO);
    s.obs = take(TR * D);
    for (int i = 0; i < L; ++i) s.ye[i] = take(TR * H);           // encoder layer outputs; ye[L-1] is state 0
    s.hst = take(K * TR * H);                                      // states 1..K
    s.x = take(K * L * TR * H);                                    // x[s][l], l < L: input of dynamics layer l at step s
    s.outs = take((K + 1) * TR * NO);
    s.dA = take(TR * H);
    s.dB = take(TR * H);
    s.dout = take(TR * NO);
    s.demb = take(A * H);
    s.red = take(32);
    s.mask_bytes_off = o * 4;
    o += ((L - 1 + K * L) * TR * H + 3) / 4;
    s.total_floats = o;
    return s;
}

struct TrainArgs {
    const float* blob;
    BlobLayout lo;
    int D, H, L, A, K, B;
    const float* obs;        // [B][D]
    const int32_t* actions;  // [B][K]
    const float* tv;         // [B]
    const float* tp;         // [B][A]
    const float* tr;         // [B][K]
    float* partials;         // [n_cta][n_params + 4]
    int n_params;
};

// y[r][j] = act(sum_k x[r][k] W[k][j] + b[j]) for the thread's TRG rows; mode 0: linear, 1: ReLU (mask kept), 2: ReLU + residual
template <int H>
__device__ __forceinline__ void fwd_linear(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ xin, int ldx,
                                           int in, float* __restrict__ y, unsigned char* __restrict__ mask, int mode, int j, int g) {
    float acc[TRG];
#pragma unroll
    for (int rr = 0; rr < TRG; ++rr) acc[rr] = 0.0f;
    const float* x0 = xin + (size_t)(g * TRG) * ldx;
#pragma unroll 4
    for (int k = 0; k < in; ++k) {
        const float w = W[k * (H + 1) + j];
#pragma unroll
        for (int rr = 0; rr < TRG; ++rr) acc[rr] = fmaf(x0[rr * ldx + k], w, acc[rr]);
    }
    const float bj = b[j];
#pragma unroll
    for (int rr = 0; rr < TRG; ++rr) {
        const int r = g * TRG + rr;
        float v = acc[rr] + bj;
        if (mode != 0) {
            mask[r * H + j] = v > 0.0f ? 1 : 0;
            v = v > 0.0f ? v : 0.0f;
            if (mode == 2) v += xin[r * ldx + j];
        }
        y[r * H + j] = v;
    }
}

// dx[r][k] (+)= sum_j dt[r][j] W[k][j]  (thread (k, g): lanes along k, row stride H+1: conflict-free)
template <int H>
__device__ __forceinline__ void bwd_input(const float* __restrict__ W, const float* __restrict__ dt, float* __restrict__ dx, bool accumulate,
                                          int k, int g) {
    float acc[TRG];
#pragma unroll
    for (int rr = 0; rr < TRG; ++rr) acc[rr] = 0.0f;
    const float* d0 = dt + (size_t)(g * TRG) * H;
#pragma unroll 4
    for (int j = 0; j < H; ++j) {
        const float w = W[k * (H + 1) + j];
#pragma unroll
        for (int rr = 0; rr < TRG; ++rr) acc[rr] = fmaf(d0[rr * H + j], w, acc[rr]);
    }
#pragma unroll
    for (int rr = 0; rr < TRG; ++rr) {
        const int r = g * TRG + rr;
        dx[r * H + k] = accumulate ? dx[r * H + k] + acc[rr] : acc[rr];
    }
}

// dW[k][j] += sum_r x[r][k] dt[r][j] for the thread's k = k0 + kk * kstep (kk < nk); db[j] += sum_r dt[r][j] (g == 0)
template <int H>
__device__ __forceinline__ void bwd_weight(const float* __restrict__ xin, int ldx, const float* __restrict__ dt, float (&accW)[TKP], float& accB,
                                           int k0, int kstep, int nk, int j, int g) {
    float d[TR];
#pragma unroll
    for (int r = 0; r < TR; ++r) d[r] = dt[r * H + j];
    if (g == 0) {
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < TR; ++r) s += d[r];
        accB += s;
    }
#pragma unroll
    for (int kk = 0; kk < TKP; ++kk) {
        if (kk < nk) {
            const int k = k0 + kk * kstep;
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < TR; ++r) s = fmaf(xin[r * ldx + k], d[r], s);
            accW[kk] += s;
        }
    }
}

template <int H, int L>
__global__ void __launch_bounds__(TG * H) train_fwd_bwd_kernel(const TrainArgs p) {
    extern __shared__ __align__(16) float sm[];
    const TrainSmem S = train_smem_layout(p.D, H, L, p.A, p.K);
    unsigned char* masks = reinterpret_cast<unsigned char*>(sm) + S.mask_bytes_off;
    const int t = threadIdx.x, j = t % H, g = t / H;
    const int D = p.D, A = p.A, K = p.K, NO = A + 2, NOP = S.nop;
    const int row0 = blockIdx.x * TR;
    constexpr int NT = TG * H;

    // ---- stage parameters (padded row stride) and the CTA's inputs -----------------------------------------------------
    {
        size_t src = p.lo.enc;
        int in = D;
        for (int i = 0; i < L; ++i) {
            for (int e = t; e < in * H; e += NT) sm[S.enc_w[i] + (e / H) * (H + 1) + (e % H)] = p.blob[src + e];
            for (int e = t; e < H; e += NT) sm[S.enc_b[i] + e] = p.blob[src + (size_t)in * H + e];
            src += (size_t)in * H + H;
            in = H;
        }
        for (int e = t; e < A * H; e += NT) sm[S.emb + e] = p.blob[p.lo.emb + e];
        for (int l = 0; l < L; ++l) {
            const size_t base = p.lo.dyn + (size_t)l * ((size_t)H * H + H);
            for (int e = t; e < H * H; e += NT) sm[S.dyn_w[l] + (e / H) * (H + 1) + (e % H)] = p.blob[base + e];
            for (int e = t; e < H; e += NT) sm[S.dyn_b[l] + e] = p.blob[base + (size_t)H * H + e];
        }
        for (int e = t; e < H * NO; e += NT) {   // heads columns: policy 0..A-1 | value | reward
            const int k = e / NO, o = e % NO;
            sm[S.heads_w + k * NOP + o] = o < A ? p.blob[p.lo.pol_k + (size_t)k * A + o] : (o == A ? p.blob[p.lo.val_k + k] : p.blob[p.lo.rew_k + k]);
        }
        for (int e = t; e < NO; e += NT) sm[S.heads_b + e] = e < A ? p.blob[p.lo.pol_b + e] : (e == A ? p.blob[p.lo.val_b] : p.blob[p.lo.rew_b]);
        for (int e = t; e < TR * D; e += NT) {
            const int r = row0 + e / D;
            sm[S.obs + e] = r < p.B ? p.obs[(size_t)r * D + e % D] : 0.0f;
        }
        for (int e = t; e < A * H; e += NT) sm[S.demb + e] = 0.0f;
    }
    __syncthreads();

    // heads of state `h` -> outs[s] (thread per (row, output))
    auto heads_fwd = [&](const float* h, int s) {
        if (t < TR * NO) {
            const int r = t / NO, o = t % NO;
            float acc = 0.0f;
            for (int k = 0; k < H; ++k) acc = fmaf(h[r * H + k], sm[S.heads_w + k * NOP + o], acc);
            sm[S.outs + (s * TR + r) * NO + o] = acc + sm[S.heads_b + o];
        }
    };

    // ---- forward ------------------------------------------------------------------------------------------------------
    {
        const float* xin = sm + S.obs;
        int ldx = D, in = D;
        for (int i = 0; i < L; ++i) {
            fwd_linear<H>(sm + S.enc_w[i], sm + S.enc_b[i], xin, ldx, in, sm + S.ye[i], masks + (size_t)i * TR * H, i < L - 1 ? 1 : 0, j, g);
            __syncthreads();
            xin = sm + S.ye[i]; ldx = H; in = H;
        }
    }
    const float* state0 = sm + S.ye[L - 1];
    heads_fwd(state0, 0);
    for (int s = 0; s < K; ++s) {
        const float* hprev = s == 0 ? state0 : sm + S.hst + (size_t)(s - 1) * TR * H;
        float* x0 = sm + S.x + (size_t)(s * L) * TR * H;
#pragma unroll
        for (int rr = 0; rr < TRG; ++rr) {
            const int r = g * TRG + rr;
            int a = row0 + r < p.B ? p.actions[(size_t)(row0 + r) * K + s] : 0;
            a = a < 0 ? 0 : (a >= A ? A - 1 : a);   // jnp.take clamps out-of-range indices
            x0[r * H + j] = hprev[r * H + j] + sm[S.emb + a * H + j];
        }
        __syncthreads();
        for (int l = 0; l < L; ++l) {
            const float* xin = sm + S.x + (size_t)(s * L + l) * TR * H;
            float* y = l == L - 1 ? sm + S.hst + (size_t)s * TR * H : sm + S.x + (size_t)(s * L + l + 1) * TR * H;
            fwd_linear<H>(sm + S.dyn_w[l], sm + S.dyn_b[l], xin, H, H, y, masks + (size_t)(L - 1 + s * L + l) * TR * H, 2, j, g);
            __syncthreads();
        }
        heads_fwd(sm + S.hst + (size_t)s * TR * H, s + 1);
    }
    __syncthreads();

    // ---- losses and their gradients w.r.t. the head outputs (thread per (state, row)) ------------------------------------
    const float inv_nv = 1.0f / ((float)p.B * (float)(K + 1)), inv_nr = 1.0f / ((float)p.B * (float)K);
    float lv = 0.0f, lp = 0.0f, lr = 0.0f;
    if (t < (K + 1) * TR) {
        const int s = t / TR, r = t % TR;
        float* o = sm + S.outs + (s * TR + r) * NO;
        if (row0 + r < p.B) {
            const size_t gr = (size_t)(row0 + r);
            float m = o[0];
            for (int a = 1; a < A; ++a) m = fmaxf(m, o[a]);
            float se = 0.0f;
            for (int a = 0; a < A; ++a) se += expf(o[a] - m);
            const float lse = m + logf(se);
            float tsum = 0.0f, ce = 0.0f;
            for (int a = 0; a < A; ++a) { const float tpa = p.tp[gr * A + a]; tsum += tpa; ce -= tpa * (o[a] - lse); }
            lp = ce;
            for (int a = 0; a < A; ++a) o[a] = (expf(o[a] - lse) * tsum - p.tp[gr * A + a]) * inv_nv;
            const float dv = o[A] - p.tv[gr];
            lv = dv * dv;
            o[A] = 2.0f * dv * inv_nv;
            if (s >= 1) {
                const float dr = o[A + 1] - p.tr[gr * K + (s - 1)];
                lr = dr * dr;
                o[A + 1] = 2.0f * dr * inv_nr;
            } else {
                o[A + 1] = 0.0f;
            }
        } else {
            for (int a = 0; a < NO; ++a) o[a] = 0.0f;
        }
    }
    // CTA sums of the three losses (warp shuffle, then one value per warp through shared memory)
    {
        for (int off = 16; off > 0; off >>= 1) {
            lv += __shfl_xor_sync(0xffffffffu, lv, off);
            lp += __shfl_xor_sync(0xffffffffu, lp, off);
            lr += __shfl_xor_sync(0xffffffffu, lr, off);
        }
        __syncthreads();
        if ((t & 31) == 0) { sm[S.red + (t >> 5)] = lv; sm[S.red + 8 + (t >> 5)] = lp; sm[S.red + 16 + (t >> 5)] = lr; }
        __syncthreads();
    }

    // ---- backward -----------------------------------------------------------------------------------------------------
    float gWd[L][TKP], gBd[L], gWe[L][TKP], gBe[L], gWh[4], gBh = 0.0f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        gBd[l] = 0.0f; gBe[l] = 0.0f;
#pragma unroll
        for (int kk = 0; kk < TKP; ++kk) { gWd[l][kk] = 0.0f; gWe[l][kk] = 0.0f; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) gWh[i] = 0.0f;
    float* dA = sm + S.dA;
    float* dB = sm + S.dB;
    // heads backward of state s: dstate (+)= dout @ Wh^T; dWh += state^T dout; dbh += sum dout
    auto heads_bwd = [&](const float* h, int s, float* dstate, bool accumulate) {
        const float* dout = sm + S.outs + (size_t)s * TR * NO;
#pragma unroll
        for (int rr = 0; rr < TRG; ++rr) {
            const int r = g * TRG + rr;
            float acc = 0.0f;
            for (int o = 0; o < NO; ++o) acc = fmaf(dout[r * NO + o], sm[S.heads_w + j * NOP + o], acc);
            dstate[r * H + j] = accumulate ? dstate[r * H + j] + acc : acc;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = t + i * NT;
            if (e < H * NO) {
                const int k = e / NO, o = e % NO;
                float sacc = 0.0f;
#pragma unroll
                for (int r = 0; r < TR; ++r) sacc = fmaf(h[r * H + k], dout[r * NO + o], sacc);
                gWh[i] += sacc;
            }
        }
        if (t < NO) {
            float sacc = 0.0f;
#pragma unroll
            for (int r = 0; r < TR; ++r) sacc += dout[r * NO + t];
            gBh += sacc;
        }
    };
    // d(state K)
    heads_bwd(K > 0 ? sm + S.hst + (size_t)(K - 1) * TR * H : state0, K, dA, false);
    __syncthreads();
    for (int s = K - 1; s >= 0; --s) {
        // dA = d x[s][L] (= d state s+1); walk the dynamics layers backwards
        for (int l = L - 1; l >= 0; --l) {
            const unsigned char* mk = masks + (size_t)(L - 1 + s * L + l) * TR * H;
            // dt = mask * dA  (in dB)
#pragma unroll
            for (int rr = 0; rr < TRG; ++rr) {
                const int r = g * TRG + rr;
                dB[r * H + j] = mk[r * H + j] ? dA[r * H + j] : 0.0f;
            }
            __syncthreads();
            const float* xin = sm + S.x + (size_t)(s * L + l) * TR * H;
            bwd_weight<H>(xin, H, dB, gWd[l], gBd[l], g * (H / TG), 1, H / TG, j, g);
            bwd_input<H>(sm + S.dyn_w[l], dB, dA, true, j, g);   // dA += dt @ W^T  (residual path keeps dA)
            __syncthreads();
        }
        // dA = d x[s][0] = d(state s + Emb[a]): embedding gradient, then add the heads' contribution of state s
        if (g == 0) {
            for (int r = 0; r < TR; ++r) {
                if (row0 + r < p.B) {
                    int a = p.actions[(size_t)(row0 + r) * K + s];
                    a = a < 0 ? 0 : (a >= A ? A - 1 : a);
                    sm[S.demb + a * H + j] += dA[r * H + j];
                }
            }
        }
        __syncthreads();   // the embedding gradient has read dA before the heads add their share of d(state s)
        heads_bwd(s == 0 ? state0 : sm + S.hst + (size_t)(s - 1) * TR * H, s, dA, true);
        __syncthreads();
    }
    if (K == 0) { /* dA already holds d(state 0) */ }
    // encoder, last layer first: dA = d ye[L-1]
    for (int i = L - 1; i >= 0; --i) {
        if (i < L - 1) {
            const unsigned char* mk = masks + (size_t)i * TR * H;
#pragma unroll
            for (int rr = 0; rr < TRG; ++rr) {
                const int r = g * TRG + rr;
                dA[r * H + j] = mk[r * H + j] ? dA[r * H + j] : 0.0f;
            }
            __syncthreads();
        }
        if (i > 0) {
            bwd_weight<H>(sm + S.ye[i - 1], H, dA, gWe[i], gBe[i], g * (H / TG), 1, H / TG, j, g);
            bwd_input<H>(sm + S.enc_w[i], dA, dB, false, j, g);
            __syncthreads();
            float* tmp = dA; dA = dB; dB = tmp;
        } else {
            bwd_weight<H>(sm + S.obs, D, dA, gWe[0], gBe[0], g, TG, (D - g + TG - 1) / TG, j, g);
        }
    }
    __syncthreads();

    // ---- the CTA's partial gradient, blob order -------------------------------------------------------------------------
    float* out = p.partials + (size_t)blockIdx.x * (p.n_params + 4);
    {
        size_t dst = p.lo.enc;
        // encoder layer 0: k = g + TG * kk
#pragma unroll
        for (int kk = 0; kk < TKP; ++kk) { const int k = g + TG * kk; if (k < D) out[dst + (size_t)k * H + j] = gWe[0][kk]; }
        if (g == 0) out[dst + (size_t)D * H + j] = gBe[0];
        dst += (size_t)D * H + H;
#pragma unroll
        for (int i = 1; i < L; ++i) {
#pragma unroll
            for (int kk = 0; kk < TKP; ++kk) { const int k = g * (H / TG) + kk; if (kk < H / TG) out[dst + (size_t)k * H + j] = gWe[i][kk]; }
            if (g == 0) out[dst + (size_t)H * H + j] = gBe[i];
            dst += (size_t)H * H + H;
        }
        for (int e = t; e < A * H; e += NT) out[p.lo.emb + e] = sm[S.demb + e];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const size_t base = p.lo.dyn + (size_t)l * ((size_t)H * H + H);
#pragma unroll
            for (int kk = 0; kk < TKP; ++kk) { const int k = g * (H / TG) + kk; if (kk < H / TG) out[base + (size_t)k * H + j] = gWd[l][kk]; }
            if (g == 0) out[base + (size_t)H * H + j] = gBd[l];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = t + i * NT;
            if (e < H * NO) {
                const int k = e / NO, o = e % NO;
                if (o < A) out[p.lo.pol_k + (size_t)k * A + o] = gWh[i];
                else if (o == A) out[p.lo.val_k + k] = gWh[i];
                else out[p.lo.rew_k + k] = gWh[i];
            }
        }
        if (t < NO) {
            if (t < A) out[p.lo.pol_b + t] = gBh;
            else if (t == A) out[p.lo.val_b] = gBh;
            else out[p.lo.rew_b] = gBh;
        }
        if (t < 3) {
            float sacc = 0.0f;
            for (int w = 0; w < NT / 32; ++w) sacc += sm[S.red + 8 * t + w];
            out[p.n_params + t] = sacc;
        }
    }
}

// grad[i] = sum over the CTAs of partial[c][i], in CTA order; losses[0..2] = value / policy / reward loss, [3] = total
__global__ void train_reduce_kernel(const float* __restrict__ partials, int n_cta, int n_params, int B, int K, float* __restrict__ grad,
                                    float* __restrict__ losses) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)n_params + 4;
    if (i < n_params) {
        float s = 0.0f;
        for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * stride + i];
        grad[i] = s;
    } else if (i == n_params) {
        float v = 0.0f, pl = 0.0f, r = 0.0f;
        for (int c = 0; c < n_cta; ++c) {
            v += partials[(size_t)c * stride + n_params];
            pl += partials[(size_t)c * stride + n_params + 1];
            r += partials[(size_t)c * stride + n_params + 2];
        }
        v /= (float)B * (float)(K + 1);
        pl /= (float)B * (float)(K + 1);
        r /= (float)B * (float)K;
        losses[0] = v; losses[1] = pl; losses[2] = r; losses[3] = v + pl + r;
    }
}

}  // namespace

bool train_fused_supported(const cumind_net_desc& d, int K) {
    const int H = d.hidden_dim, L = d.num_blocks;
    return d.obs_ndim == 1 && (H == 32 || H == 64) && L >= 1 && L <= 3 && d.obs_shape[0] <= TG * TKP && d.action_space_size + 2 <= 16 &&
           H * (d.action_space_size + 2) <= 4 * TG * H && K >= 1 && K <= 16 && (K + 1) * TR <= TG * H;
}

size_t train_fused_work_bytes(const cumind_net_desc& d, int B) {
    const size_t n_params = make_blob_layout(&d).total;
    return (size_t)((B + TR - 1) / TR) * (n_params + 4) * sizeof(float);
}

template <int H, int L>
static cudaError_t launch_train(const TrainArgs& a, float* grad, float* losses, cudaStream_t st) {
    const TrainSmem S = train_smem_layout(a.D, H, L, a.A, a.K);
    const size_t smem = (size_t)S.total_floats * 4;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(train_fwd_bwd_kernel<H, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int n_cta = (a.B + TR - 1) / TR;
    train_fwd_bwd_kernel<H, L><<<n_cta, TG * H, smem, st>>>(a);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    train_reduce_kernel<<<(a.n_params + 1 + 255) / 256, 256, 0, st>>>(a.partials, n_cta, a.n_params, a.B, a.K, grad, losses);
    return cudaGetLastError();
}

cudaError_t launch_train_fwd_bwd(const cumind_net_desc& d, const float* blob, const float* obs, const int32_t* actions, const float* tv,
                                 const float* tp, const float* tr, int B, int K, float* work, float* grad, float* losses, cudaStream_t st) {
    if (!train_fused_supported(d, K)) return cudaErrorInvalidValue;
    TrainArgs a;
    a.blob = blob; a.lo = make_blob_layout(&d);
    a.D = d.obs_shape[0]; a.H = d.hidden_dim; a.L = d.num_blocks; a.A = d.action_space_size; a.K = K; a.B = B;
    a.obs = obs; a.actions = actions; a.tv = tv; a.tp = tp; a.tr = tr; a.partials = work; a.n_params = (int)a.lo.total;
    const int H = a.H, L = a.L;
#define CM_TRAIN_CASE(HH, LL) if (H == HH && L == LL) return launch_train<HH, LL>(a, grad, losses, st);
    CM_TRAIN_CASE(64, 1) CM_TRAIN_CASE(64, 2) CM_TRAIN_CASE(64, 3) CM_TRAIN_CASE(32, 1) CM_TRAIN_CASE(32, 2) CM_TRAIN_CASE(32, 3)
#undef CM_TRAIN_CASE
    return cudaErrorInvalidValue;
}

}  // namespace cm
