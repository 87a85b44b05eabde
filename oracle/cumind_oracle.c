/*
 * cumind_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement of CuMind's MuZero search hot path.  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg have an independent checker for the
 * CUDA kernels in cumind_b200/csrc.  Nothing under cumind_b200/ may import, link or call it.
 *
 * What it restates (reference = /root/reference, carletonai/CuMind v0.1.7):
 *   tree policy   src/cumind/core/mcts.py:18-98   (Node: value, ucb_score, select_child, expand, backup)
 *   search loop   src/cumind/core/mcts.py:115-222 (MCTS.search, _simulate, _add_exploration_noise)
 *   networks      src/cumind/core/network.py:13-45 (ResidualBlock), 70-108 (VectorEncoder),
 *                 111-147 (ConvEncoder), 194-236 (DynamicsNetwork), 239-266 (PredictionNetwork),
 *                 288-315 (initial_inference / recurrent_inference)
 * The flax/jax layer arithmetic is third-party (flax 0.10.6 / jax 0.6.2, not vendored in the
 * reference): Linear y = x@K[in,out]+b; Embed = row lookup; Conv 3x3 NHWC, symmetric pad 1,
 * cross-correlation; BatchNorm inference (x-mean)*rsqrt(var+1e-5)*scale+bias; softmax =
 * exp(x-max)/sum.  Those libraries do not define a floating-point summation order, so this file
 * PINS one (DESIGN.md "Canonical numerics"):
 *   - every dot product is a sequential chain of IEEE fused multiply-adds over k = 0..K-1:
 *     acc = fmaf(x_k, w_k, acc) (ONE rounding per step), acc0 = +0, then fl32(acc + bias).
 *     Every other operation is a single correctly rounded fp32 op and is never contracted
 *     (compile with -ffp-contract=off);
 *   - conv taps in (ky,kx,ci) order, taps that fall in the zero padding are skipped;
 *   - exp for softmax is cm_pexp() below: fp64, plain mul/add only, rounded once to fp32;
 *   - tree statistics are float64 in the reference's operation order (mcts.py:64-65, 97-98, 222).
 *
 * Parity pinning: the tests/golden npz files were produced by executing the reference's own mcts.py
 * verbatim (oracle/reference_exec.py) over the NumPy network of oracle/network_np.py;
 * tests/test_oracle_golden.py checks this file against those vectors bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cumind_b200.h"

#define CMO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* parameter blob layout (shared with the product through the header comment only)            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    size_t enc;        /* start of encoder params                       */
    size_t emb;        /* dynamics embedding [A,H]                      */
    size_t dyn;        /* dynamics layers: L x (kernel[H,H], bias[H])   */
    size_t rew_k, rew_b;
    size_t pol_k, pol_b;
    size_t val_k, val_b;
    size_t total;
} cmo_layout;

static int cmo_valid(const cumind_net_desc* d) {
    if (!d) return 0;
    if (d->hidden_dim <= 0 || d->num_blocks <= 0 || d->action_space_size <= 0) return 0;
    if (d->obs_ndim == 1) return d->obs_shape[0] > 0;
    if (d->obs_ndim == 3)
        return d->obs_shape[0] > 0 && d->obs_shape[1] > 0 && d->obs_shape[2] > 0 && d->conv_channels > 0;
    return 0;
}

static size_t cmo_encoder_count(const cumind_net_desc* d) {
    size_t H = d->hidden_dim, L = d->num_blocks;
    if (d->obs_ndim == 1) {
        size_t D = d->obs_shape[0];
        return D * H + H + (L - 1) * (H * H + H);
    }
    size_t C = d->obs_shape[2], Cc = d->conv_channels;
    size_t n = 9 * C * Cc + Cc;                         /* initial_conv */
    n += L * (2 * (9 * Cc * Cc + Cc) + 2 * 4 * Cc);     /* residual blocks */
    n += Cc * H + H;                                    /* final_dense */
    return n;
}

static void cmo_make_layout(const cumind_net_desc* d, cmo_layout* lo) {
    size_t H = d->hidden_dim, L = d->num_blocks, A = d->action_space_size;
    size_t o = 0;
    lo->enc = o;   o += cmo_encoder_count(d);
    lo->emb = o;   o += A * H;
    lo->dyn = o;   o += L * (H * H + H);
    lo->rew_k = o; o += H;
    lo->rew_b = o; o += 1;
    lo->pol_k = o; o += H * A;
    lo->pol_b = o; o += A;
    lo->val_k = o; o += H;
    lo->val_b = o; o += 1;
    lo->total = o;
}

CMO_EXPORT size_t cmo_param_count(const cumind_net_desc* d) {
    if (!cmo_valid(d)) return 0;
    cmo_layout lo;
    cmo_make_layout(d, &lo);
    return lo.total;
}

/* ------------------------------------------------------------------------------------------ */
/* pinned scalar numerics                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* exp(d) for d <= 0 in float64 using only IEEE mul/add (bit-reproducible on any IEEE machine,
 * and on the GPU with __dmul_rn/__dadd_rn).  Cody-Waite reduction by ln2 (hi part has 21
 * trailing zero bits so k*LN2_HI is exact for |k| < 2^21), degree-13 Taylor polynomial in
 * Horner form on |r| <= 0.3466, scaling by 2^k through the exponent field. */
static const double CM_LOG2E  = 1.4426950408889634;          /* 0x3FF71547652B82FE */
static const double CM_LN2_HI = 6.93147180369123816490e-01;  /* 0x3FE62E42FEE00000 */
static const double CM_LN2_LO = 1.90821492927058770002e-10;  /* 0x3DEA39EF35793C76 */
static const double CM_EXP_C[14] = {
    1.0, 1.0, 0.5,
    1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
    1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};

CMO_EXPORT double cmo_pexp(double d) {
    if (d != d) return d;            /* NaN propagates */
    if (!(d >= -200.0)) return 0.0;  /* exp(-200) rounds to 0 in fp32 */
    if (d > 0.0) d = 0.0;            /* never happens for x - max(x) */
    double kf = rint(d * CM_LOG2E);  /* round half to even (default rounding mode) */
    double r = (d - kf * CM_LN2_HI) - kf * CM_LN2_LO;
    double p = CM_EXP_C[13];
    for (int i = 12; i >= 0; --i) p = p * r + CM_EXP_C[i];
    int64_t k = (int64_t)kf;         /* -289 <= k <= 0 */
    uint64_t bits = (uint64_t)(k + 1023) << 52;
    double scale;
    memcpy(&scale, &bits, 8);
    return p * scale;
}

/* fmaf() exposed so the tests can pin network_np.fma32 (the NumPy emulation) against it. */
CMO_EXPORT void cmo_fmaf_array(const float* a, const float* b, const float* c, float* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}

/* jax.nn.softmax(logits, axis=-1) in fp32 (mcts.py:128,191) with the pinned exp. */
CMO_EXPORT void cmo_softmax(const float* logits, int A, float* out) {
    float m = logits[0];
    for (int a = 1; a < A; ++a) if (logits[a] > m) m = logits[a];
    float s = 0.0f;
    for (int a = 0; a < A; ++a) {
        float d = logits[a] - m;
        float e = (float)cmo_pexp((double)d);
        out[a] = e;
        s = s + e;
    }
    for (int a = 0; a < A; ++a) out[a] = out[a] / s;
}

/* nnx.Linear for one row: y[j] = fl(fma-chain_k(x[k], K[k,j]) + b[j]); K is [in,out] row-major. */
static void cmo_linear(const float* x, const float* K, const float* b, int in, int out, float* y) {
    for (int j = 0; j < out; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < in; ++k) acc = fmaf(x[k], K[(size_t)k * out + j], acc);
        y[j] = acc + b[j];
    }
}

static inline float cmo_relu(float v) { return v > 0.0f ? v : 0.0f; }

/* PredictionNetwork.__call__ (network.py:254-266) for one row. */
static void cmo_prediction_row(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                               const float* h, float* logits, float* value) {
    int H = d->hidden_dim, A = d->action_space_size;
    if (logits) cmo_linear(h, P + lo->pol_k, P + lo->pol_b, H, A, logits);
    if (value) cmo_linear(h, P + lo->val_k, P + lo->val_b, H, 1, value);
}

/* DynamicsNetwork.__call__ (network.py:217-236) for one row. scratch: 2*H floats. */
static void cmo_dynamics_row(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                             const float* h, int action, float* next, float* reward, float* scratch) {
    int H = d->hidden_dim, L = d->num_blocks;
    float* x = scratch;
    float* t = scratch + H;
    const float* emb = P + lo->emb + (size_t)action * H;
    for (int j = 0; j < H; ++j) x[j] = h[j] + emb[j];
    for (int l = 0; l < L; ++l) {
        const float* K = P + lo->dyn + (size_t)l * ((size_t)H * H + H);
        const float* b = K + (size_t)H * H;
        cmo_linear(x, K, b, H, H, t);
        for (int j = 0; j < H; ++j) x[j] = cmo_relu(t[j]) + x[j];
    }
    if (reward) cmo_linear(x, P + lo->rew_k, P + lo->rew_b, H, 1, reward);
    memcpy(next, x, sizeof(float) * H);
}

/* VectorEncoder.__call__ (network.py:94-108): L Linear layers, ReLU between, none after the last. */
static void cmo_vector_encoder_row(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                                   const float* obs, float* h, float* scratch) {
    int H = d->hidden_dim, L = d->num_blocks, D = d->obs_shape[0];
    const float* p = P + lo->enc;
    const float* x = obs;
    int in = D;
    float* bufs[2] = {scratch, scratch + H};
    for (int i = 0; i < L; ++i) {
        float* y = (i == L - 1) ? h : bufs[i & 1];
        cmo_linear(x, p, p + (size_t)in * H, in, H, y);
        if (i < L - 1) for (int j = 0; j < H; ++j) y[j] = cmo_relu(y[j]);
        p += (size_t)in * H + H;
        x = y;
        in = H;
    }
}

/* nnx.Conv 3x3, NHWC, padding=1, given stride; kernel [3,3,Cin,Cout]; taps in the zero padding
 * are skipped; sequential (ky,kx,ci) accumulation, then + bias. */
static void cmo_conv3x3(const float* in, int Hh, int Ww, int Cin, const float* K, const float* b,
                        int Cout, int stride, float* out, int Ho, int Wo) {
    for (int oy = 0; oy < Ho; ++oy)
        for (int ox = 0; ox < Wo; ++ox)
            for (int co = 0; co < Cout; ++co) {
                float acc = 0.0f;
                for (int ky = 0; ky < 3; ++ky) {
                    int iy = oy * stride - 1 + ky;
                    if (iy < 0 || iy >= Hh) continue;
                    for (int kx = 0; kx < 3; ++kx) {
                        int ix = ox * stride - 1 + kx;
                        if (ix < 0 || ix >= Ww) continue;
                        const float* ip = in + ((size_t)iy * Ww + ix) * Cin;
                        const float* kp = K + ((size_t)(ky * 3 + kx) * Cin) * Cout + co;
                        for (int ci = 0; ci < Cin; ++ci) acc = fmaf(ip[ci], kp[(size_t)ci * Cout], acc);
                    }
                }
                out[((size_t)oy * Wo + ox) * Cout + co] = acc + b[co];
            }
}

/* nnx.BatchNorm(use_running_average=True): y = fl(fl(fl(x-mean)*g)+beta),
 * g = fl(scale * fl(1/fl(sqrt(fl(var+1e-5))))) computed once per channel. */
CMO_EXPORT float cmo_bn_gain(float scale, float var) {
    float s = sqrtf(var + 1e-5f);
    float inv = 1.0f / s;
    return scale * inv;
}

static void cmo_bn(float* x, size_t n_pix, int C, const float* bn /* scale,bias,mean,var */) {
    const float* scale = bn;
    const float* beta = bn + C;
    const float* mean = bn + 2 * C;
    const float* var = bn + 3 * C;
    for (int c = 0; c < C; ++c) {
        float g = cmo_bn_gain(scale[c], var[c]);
        for (size_t p = 0; p < n_pix; ++p) {
            float v = x[p * C + c] - mean[c];
            v = v * g;
            x[p * C + c] = v + beta[c];
        }
    }
}

/* ConvEncoder.__call__ (network.py:131-147) + ResidualBlock.__call__ (29-45) for one image. */
static void cmo_conv_encoder_row(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                                 const float* obs, float* h) {
    int Hh = d->obs_shape[0], Ww = d->obs_shape[1], C = d->obs_shape[2];
    int Cc = d->conv_channels, L = d->num_blocks, H = d->hidden_dim;
    int Ho = (Hh - 1) / 2 + 1, Wo = (Ww - 1) / 2 + 1;
    size_t npix = (size_t)Ho * Wo;
    float* x = (float*)malloc(sizeof(float) * npix * Cc);
    float* t = (float*)malloc(sizeof(float) * npix * Cc);
    float* u = (float*)malloc(sizeof(float) * npix * Cc);
    const float* p = P + lo->enc;
    cmo_conv3x3(obs, Hh, Ww, C, p, p + (size_t)9 * C * Cc, Cc, 2, x, Ho, Wo);
    for (size_t i = 0; i < npix * Cc; ++i) x[i] = cmo_relu(x[i]);
    p += (size_t)9 * C * Cc + Cc;
    for (int l = 0; l < L; ++l) {
        const float* k1 = p;              const float* b1 = k1 + (size_t)9 * Cc * Cc;
        const float* bn1 = b1 + Cc;       const float* k2 = bn1 + 4 * Cc;
        const float* b2 = k2 + (size_t)9 * Cc * Cc;
        const float* bn2 = b2 + Cc;
        cmo_conv3x3(x, Ho, Wo, Cc, k1, b1, Cc, 1, t, Ho, Wo);
        cmo_bn(t, npix, Cc, bn1);
        for (size_t i = 0; i < npix * Cc; ++i) t[i] = cmo_relu(t[i]);
        cmo_conv3x3(t, Ho, Wo, Cc, k2, b2, Cc, 1, u, Ho, Wo);
        cmo_bn(u, npix, Cc, bn2);
        for (size_t i = 0; i < npix * Cc; ++i) x[i] = cmo_relu(u[i] + x[i]);
        p = bn2 + 4 * Cc;
    }
    /* jnp.mean over (H,W): sequential row-major sum, then / count */
    float* pooled = (float*)malloc(sizeof(float) * Cc);
    for (int c = 0; c < Cc; ++c) {
        float s = 0.0f;
        for (size_t i = 0; i < npix; ++i) s = s + x[i * Cc + c];
        pooled[c] = s / (float)npix;
    }
    cmo_linear(pooled, p, p + (size_t)Cc * H, Cc, H, h);
    free(pooled); free(u); free(t); free(x);
}

static size_t cmo_obs_size(const cumind_net_desc* d) {
    return d->obs_ndim == 1 ? (size_t)d->obs_shape[0]
                            : (size_t)d->obs_shape[0] * d->obs_shape[1] * d->obs_shape[2];
}

/* CuMindNetwork.initial_inference (network.py:288-300), batch of rows evaluated one at a time. */
CMO_EXPORT int cmo_initial_inference(const cumind_net_desc* d, const float* P, const float* obs, int B,
                                     float* hidden, float* logits, float* value) {
    if (!cmo_valid(d)) return 1;
    cmo_layout lo; cmo_make_layout(d, &lo);
    int H = d->hidden_dim, A = d->action_space_size;
    size_t osz = cmo_obs_size(d);
    float* scratch = (float*)malloc(sizeof(float) * 2 * H);
    for (int b = 0; b < B; ++b) {
        float* h = hidden + (size_t)b * H;
        if (d->obs_ndim == 1) cmo_vector_encoder_row(d, &lo, P, obs + b * osz, h, scratch);
        else cmo_conv_encoder_row(d, &lo, P, obs + b * osz, h);
        cmo_prediction_row(d, &lo, P, h, logits ? logits + (size_t)b * A : NULL, value ? value + b : NULL);
    }
    free(scratch);
    return 0;
}

/* CuMindNetwork.recurrent_inference (network.py:302-315). */
CMO_EXPORT int cmo_recurrent_inference(const cumind_net_desc* d, const float* P, const float* hidden,
                                       const int32_t* action, int B, float* next, float* reward,
                                       float* logits, float* value) {
    if (!cmo_valid(d)) return 1;
    cmo_layout lo; cmo_make_layout(d, &lo);
    int H = d->hidden_dim, A = d->action_space_size;
    float* scratch = (float*)malloc(sizeof(float) * 3 * H);
    for (int b = 0; b < B; ++b) {
        if (action[b] < 0 || action[b] >= A) { free(scratch); return 2; }
        float* nx = next ? next + (size_t)b * H : scratch + 2 * H;
        cmo_dynamics_row(d, &lo, P, hidden + (size_t)b * H, action[b], nx, reward ? reward + b : NULL, scratch);
        cmo_prediction_row(d, &lo, P, nx, logits ? logits + (size_t)b * A : NULL, value ? value + b : NULL);
    }
    free(scratch);
    return 0;
}

CMO_EXPORT int cmo_prediction(const cumind_net_desc* d, const float* P, const float* hidden, int B,
                              float* logits, float* value) {
    if (!cmo_valid(d)) return 1;
    cmo_layout lo; cmo_make_layout(d, &lo);
    int H = d->hidden_dim, A = d->action_space_size;
    for (int b = 0; b < B; ++b)
        cmo_prediction_row(d, &lo, P, hidden + (size_t)b * H, logits ? logits + (size_t)b * A : NULL,
                           value ? value + b : NULL);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* the search                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* Node.ucb_score (mcts.py:51-65) in float64, the reference's operation order:
 *   inf if n == 0 else W/n + ((c*p)*sqrt(N_parent))/(1+n)                                   */
static inline double cmo_ucb(int32_t n, double W, double prior, int32_t n_parent, double c_puct) {
    if (n == 0) return INFINITY;
    double explo = ((c_puct * prior) * sqrt((double)n_parent)) / (double)(1 + n);
    return W / (double)n + explo;
}

/* One tree.  Arrays are the rows of the SoA described in include/cumind_b200.h.
 * Returns the sum of path lengths (parents per simulation) for the d-bar statistic. */
static int64_t cmo_search_one(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                              const float* root_h, int S, int S_max, int A_tree, double c_puct,
                              double frac, const double* noise /* [A_tree] or NULL */,
                              int32_t* visit, double* value_sum, float* prior, double* root_prior,
                              int32_t* child_eidx, int32_t* exp_node, float* hidden,
                              int32_t* path, float* scratch /* 3H + 2A */) {
    int H = d->hidden_dim, A_net = d->action_space_size;
    int N = 1 + A_tree * (S_max + 1);
    float* logits = scratch + 3 * H;
    float* probs = logits + A_net;
    for (int i = 0; i < N; ++i) { visit[i] = 0; value_sum[i] = 0.0; prior[i] = 0.0f; child_eidx[i] = -1; }
    for (int i = 0; i <= S_max; ++i) exp_node[i] = -1;

    /* MCTS.search, mcts.py:126-136: root priors, root = Node(1.0), expand, optional noise */
    cmo_prediction_row(d, lo, P, root_h, logits, NULL);
    cmo_softmax(logits, A_net, probs);
    prior[0] = 1.0f;
    child_eidx[0] = 0;
    exp_node[0] = 0;
    memcpy(hidden, root_h, sizeof(float) * H);
    for (int a = 0; a < A_tree; ++a) {
        prior[1 + a] = probs[a];
        double p = (double)probs[a];
        if (noise) p = p * (1.0 - frac) + noise[a] * frac;   /* mcts.py:222 */
        root_prior[a] = p;
    }
    int n_exp = 1;
    int64_t depth_sum = 0;

    for (int s = 0; s < S; ++s) {
        /* selection (mcts.py:165-171) */
        int node = 0, plen = 0;
        while (child_eidx[node] >= 0) {
            int base = 1 + child_eidx[node] * A_tree;
            int best = 0;
            double best_score = 0.0;
            for (int a = 0; a < A_tree; ++a) {
                int c = base + a;
                double pr = (node == 0) ? root_prior[a] : (double)prior[c];
                double sc = cmo_ucb(visit[c], value_sum[c], pr, visit[node], c_puct);
                if (a == 0 || sc > best_score) { best = a; best_score = sc; }   /* first maximum */
            }
            path[plen++] = node;
            node = base + best;
        }
        depth_sum += plen;
        /* expansion + evaluation (mcts.py:181-196); the leaf always has a parent here */
        int parent_e = (node - 1) / A_tree;
        int action = (node - 1) % A_tree;
        float* leaf_h = hidden + (size_t)n_exp * H;
        cmo_dynamics_row(d, lo, P, hidden + (size_t)parent_e * H, action, leaf_h, NULL, scratch);
        float value;
        cmo_prediction_row(d, lo, P, leaf_h, logits, &value);
        cmo_softmax(logits, A_net, probs);
        child_eidx[node] = n_exp;
        exp_node[n_exp] = node;
        for (int a = 0; a < A_tree; ++a) prior[1 + n_exp * A_tree + a] = probs[a];
        n_exp++;
        /* backup (mcts.py:199-203): every parent on the path, then the root once more */
        double v = (double)value;
        for (int i = plen - 1; i >= 0; --i) {
            visit[path[i]] += 1;
            value_sum[path[i]] = value_sum[path[i]] + v;
        }
        visit[0] += 1;
        value_sum[0] = value_sum[0] + v;
    }
    return depth_sum;
}

/* MCTS.search for B trees, one after the other (mcts.py:115-155).
 * Tree arrays may be NULL except visit/value_sum/... which are required (sized as in the header).
 * out_probs follows mcts.py:143-153 (float64; uniform when the root children have no visits). */
CMO_EXPORT int cmo_search(const cumind_net_desc* d, const float* P, const float* root_hidden, int B,
                          int S, int S_max, int A_cfg, double c_puct, double frac,
                          const double* noise /* [B,A_tree] or NULL */,
                          int32_t* visit, double* value_sum, float* prior, double* root_prior,
                          int32_t* child_eidx, int32_t* exp_node, float* hidden,
                          int32_t* out_visits, double* out_probs, double* out_root_value,
                          int64_t* depth_sum) {
    if (!cmo_valid(d) || S < 0 || S > S_max || A_cfg <= 0) return 1;
    cmo_layout lo; cmo_make_layout(d, &lo);
    int H = d->hidden_dim, A_net = d->action_space_size;
    int A_tree = A_cfg < A_net ? A_cfg : A_net;
    int N = 1 + A_tree * (S_max + 1);
    int32_t* path = (int32_t*)malloc(sizeof(int32_t) * (S_max + 2));
    float* scratch = (float*)malloc(sizeof(float) * (3 * H + 2 * A_net));
    for (int b = 0; b < B; ++b) {
        int64_t ds = cmo_search_one(d, &lo, P, root_hidden + (size_t)b * H, S, S_max, A_tree, c_puct, frac,
                                    noise ? noise + (size_t)b * A_tree : NULL,
                                    visit + (size_t)b * N, value_sum + (size_t)b * N, prior + (size_t)b * N,
                                    root_prior + (size_t)b * A_tree, child_eidx + (size_t)b * N,
                                    exp_node + (size_t)b * (S_max + 1),
                                    hidden + (size_t)b * (S_max + 1) * H, path, scratch);
        if (depth_sum) depth_sum[b] = ds;
        const int32_t* v = visit + (size_t)b * N;
        int64_t total = 0;
        for (int a = 0; a < A_tree; ++a) total += v[1 + a];
        for (int a = 0; a < A_cfg; ++a) {
            int32_t c = a < A_tree ? v[1 + a] : 0;
            if (out_visits) out_visits[(size_t)b * A_cfg + a] = c;
            if (out_probs)
                out_probs[(size_t)b * A_cfg + a] = total == 0 ? 1.0 / (double)A_cfg : (double)c / (double)total;
        }
        if (out_root_value)
            out_root_value[b] = v[0] == 0 ? 0.0 : value_sum[(size_t)b * N] / (double)v[0];
    }
    free(scratch); free(path);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* OPTIONAL tree mode CUMIND_TREE_MINMAX_DISCOUNT (include/cumind_b200.h) — NOT the reference.    */
/* The reference's MCTS (mcts.py:51-65, 157-203) has no min-max normalisation, no discount, no   */
/* reward in the backup and never counts the leaf; BASELINE.json's north_star names the textbook */
/* MuZero rules as a mode, and this is its definition (parity for this mode is against THIS      */
/* restatement and oracle/mcts_minmax.py; there is no reference implementation to pin it to).    */
/*   score(child)  = vs + ((c_puct*prior)*sqrt(N_parent))/(1+n),  N_parent = parent.visit_count   */
/*                   vs = 0 if n == 0 else normalise(reward + discount*(W/n))                     */
/*   normalise(x)  = (x - mn)/(mx - mn) if mx > mn else x      (mn, mx: per-tree bounds)          */
/*   select        = first maximum in action order (Python max), from the root down to a leaf    */
/*   expand        = hidden, reward = dynamics(parent.hidden, action); priors, value = prediction */
/*   backup        = for node in reversed(path incl. the leaf): W += G; n += 1;                   */
/*                   q = reward + discount*(W/n); mn = q if q < mn; mx = q if q > mx;             */
/*                   G = reward + discount*G         (the root's reward is 0, it is counted once) */
/* All statistics float64, one correctly rounded operation each, in the order written.          */
/* ------------------------------------------------------------------------------------------ */
static inline double cmo_mm_norm(double x, double mn, double mx) {
    return mx > mn ? (x - mn) / (mx - mn) : x;
}
static inline double cmo_mm_score(int32_t n, double W, double prior, double reward, int32_t n_parent, double c_puct,
                                  double discount, double mn, double mx) {
    double explo = ((c_puct * prior) * sqrt((double)n_parent)) / (double)(1 + n);
    double vs = 0.0;
    if (n != 0) vs = cmo_mm_norm(reward + discount * (W / (double)n), mn, mx);
    return vs + explo;
}

static int64_t cmo_search_one_mm(const cumind_net_desc* d, const cmo_layout* lo, const float* P,
                                 const float* root_h, int S, int S_max, int A_tree, double c_puct, double discount,
                                 double frac, const double* noise,
                                 int32_t* visit, double* value_sum, float* prior, double* root_prior, float* reward,
                                 double* minmax, int32_t* child_eidx, int32_t* exp_node, float* hidden,
                                 int32_t* path, float* scratch) {
    int H = d->hidden_dim, A_net = d->action_space_size;
    int N = 1 + A_tree * (S_max + 1);
    float* logits = scratch + 3 * H;
    float* probs = logits + A_net;
    for (int i = 0; i < N; ++i) { visit[i] = 0; value_sum[i] = 0.0; prior[i] = 0.0f; reward[i] = 0.0f; child_eidx[i] = -1; }
    for (int i = 0; i <= S_max; ++i) exp_node[i] = -1;
    double mn = INFINITY, mx = -INFINITY;
    cmo_prediction_row(d, lo, P, root_h, logits, NULL);
    cmo_softmax(logits, A_net, probs);
    prior[0] = 1.0f;
    child_eidx[0] = 0;
    exp_node[0] = 0;
    memcpy(hidden, root_h, sizeof(float) * H);
    for (int a = 0; a < A_tree; ++a) {
        prior[1 + a] = probs[a];
        double p = (double)probs[a];
        if (noise) p = p * (1.0 - frac) + noise[a] * frac;
        root_prior[a] = p;
    }
    int n_exp = 1;
    int64_t depth_sum = 0;
    for (int s = 0; s < S; ++s) {
        int node = 0, plen = 0;
        while (child_eidx[node] >= 0) {
            int base = 1 + child_eidx[node] * A_tree;
            int best = 0;
            double best_score = 0.0;
            for (int a = 0; a < A_tree; ++a) {
                int c = base + a;
                double pr = (node == 0) ? root_prior[a] : (double)prior[c];
                double sc = cmo_mm_score(visit[c], value_sum[c], pr, (double)reward[c], visit[node], c_puct, discount, mn, mx);
                if (a == 0 || sc > best_score) { best = a; best_score = sc; }
            }
            path[plen++] = node;
            node = base + best;
        }
        depth_sum += plen;
        int parent_e = (node - 1) / A_tree;
        int action = (node - 1) % A_tree;
        float* leaf_h = hidden + (size_t)n_exp * H;
        float rew;
        cmo_dynamics_row(d, lo, P, hidden + (size_t)parent_e * H, action, leaf_h, &rew, scratch);
        float value;
        cmo_prediction_row(d, lo, P, leaf_h, logits, &value);
        cmo_softmax(logits, A_net, probs);
        reward[node] = rew;
        child_eidx[node] = n_exp;
        exp_node[n_exp] = node;
        for (int a = 0; a < A_tree; ++a) prior[1 + n_exp * A_tree + a] = probs[a];
        n_exp++;
        double G = (double)value;
        for (int i = plen; i >= 0; --i) {          /* the leaf first, then its ancestors up to the root */
            int nd = i == plen ? node : path[i];
            value_sum[nd] = value_sum[nd] + G;
            visit[nd] += 1;
            double r = (double)reward[nd];
            double q = r + discount * (value_sum[nd] / (double)visit[nd]);
            if (q < mn) mn = q;
            if (q > mx) mx = q;
            G = r + discount * G;
        }
    }
    minmax[0] = mn;
    minmax[1] = mx;
    return depth_sum;
}

CMO_EXPORT int cmo_search_mm(const cumind_net_desc* d, const float* P, const float* root_hidden, int B,
                             int S, int S_max, int A_cfg, double c_puct, double discount, double frac,
                             const double* noise, int32_t* visit, double* value_sum, float* prior, double* root_prior,
                             float* reward, double* minmax, int32_t* child_eidx, int32_t* exp_node, float* hidden,
                             int32_t* out_visits, double* out_probs, double* out_root_value, int64_t* depth_sum) {
    if (!cmo_valid(d) || S < 0 || S > S_max || A_cfg <= 0) return 1;
    cmo_layout lo; cmo_make_layout(d, &lo);
    int H = d->hidden_dim, A_net = d->action_space_size;
    int A_tree = A_cfg < A_net ? A_cfg : A_net;
    int N = 1 + A_tree * (S_max + 1);
    int32_t* path = (int32_t*)malloc(sizeof(int32_t) * (S_max + 2));
    float* scratch = (float*)malloc(sizeof(float) * (3 * H + 2 * A_net));
    for (int b = 0; b < B; ++b) {
        int64_t ds = cmo_search_one_mm(d, &lo, P, root_hidden + (size_t)b * H, S, S_max, A_tree, c_puct, discount, frac,
                                       noise ? noise + (size_t)b * A_tree : NULL,
                                       visit + (size_t)b * N, value_sum + (size_t)b * N, prior + (size_t)b * N,
                                       root_prior + (size_t)b * A_tree, reward + (size_t)b * N, minmax + (size_t)b * 2,
                                       child_eidx + (size_t)b * N, exp_node + (size_t)b * (S_max + 1),
                                       hidden + (size_t)b * (S_max + 1) * H, path, scratch);
        if (depth_sum) depth_sum[b] = ds;
        const int32_t* v = visit + (size_t)b * N;
        int64_t total = 0;
        for (int a = 0; a < A_tree; ++a) total += v[1 + a];
        for (int a = 0; a < A_cfg; ++a) {
            int32_t c = a < A_tree ? v[1 + a] : 0;
            if (out_visits) out_visits[(size_t)b * A_cfg + a] = c;
            if (out_probs)
                out_probs[(size_t)b * A_cfg + a] = total == 0 ? 1.0 / (double)A_cfg : (double)c / (double)total;
        }
        if (out_root_value)
            out_root_value[b] = v[0] == 0 ? 0.0 : value_sum[(size_t)b * N] / (double)v[0];
    }
    free(scratch); free(path);
    return 0;
}
