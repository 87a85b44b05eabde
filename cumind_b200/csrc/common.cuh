// common.cuh — shared declarations of the CuMind-B200 kernels (sm_100a).
//
// Canonical numerics: every value that feeds the bit-exactness gate is produced with
// explicitly rounded intrinsics (__fmaf_rn / fma.rn.f32x2 for the dot-product chains,
// __fmul_rn/__fadd_rn/__dmul_rn/__dadd_rn/__ddiv_rn/__dsqrt_rn elsewhere), which nvcc never
// re-associates or contracts, in the operation order of the reference
// (src/cumind/core/mcts.py:51-65, 91-98, 222) and of the pinned network arithmetic
// (dot product = sequential-k chain of IEEE fused multiply-adds, then + bias).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cumind_b200.h"

#define CM_HOST_DEVICE __host__ __device__ __forceinline__
#define CM_DEVICE __device__ __forceinline__

namespace cm {

// ------------------------------------------------------------------------------------------------
// parameter blob layout (include/cumind_b200.h) and the packed "search" layout kept on device
// ------------------------------------------------------------------------------------------------
struct BlobLayout {
    size_t enc, emb, dyn, rew_k, rew_b, pol_k, pol_b, val_k, val_b, total;
};

inline bool desc_valid(const cumind_net_desc* d) {
    if (!d) return false;
    if (d->hidden_dim <= 0 || d->num_blocks <= 0 || d->action_space_size <= 0) return false;
    if (d->obs_ndim == 1) return d->obs_shape[0] > 0;
    if (d->obs_ndim == 3)
        return d->obs_shape[0] > 0 && d->obs_shape[1] > 0 && d->obs_shape[2] > 0 && d->conv_channels > 0;
    return false;
}

inline size_t encoder_count(const cumind_net_desc* d) {
    size_t H = d->hidden_dim, L = d->num_blocks;
    if (d->obs_ndim == 1) {
        size_t D = d->obs_shape[0];
        return D * H + H + (L - 1) * (H * H + H);
    }
    size_t C = d->obs_shape[2], Cc = d->conv_channels;
    return (9 * C * Cc + Cc) + L * (2 * (9 * Cc * Cc + Cc) + 8 * Cc) + (Cc * H + H);
}

inline BlobLayout make_blob_layout(const cumind_net_desc* d) {
    size_t H = d->hidden_dim, L = d->num_blocks, A = d->action_space_size;
    BlobLayout lo;
    size_t o = 0;
    lo.enc = o;   o += encoder_count(d);
    lo.emb = o;   o += A * H;
    lo.dyn = o;   o += L * (H * H + H);
    lo.rew_k = o; o += H;
    lo.rew_b = o; o += 1;
    lo.pol_k = o; o += H * A;
    lo.pol_b = o; o += A;
    lo.val_k = o; o += H;
    lo.val_b = o; o += 1;
    lo.total = o;
    return lo;
}

// Packed recurrent-path parameters ("pack"): one 16-byte-aligned fp32 buffer the search kernel
// bulk-copies into shared memory:
//   emb [A][H] | L x ( K_l [H][H] | b_l [H] ) | Wh [H][HP] | bh [HP]
// Head columns: 0..A-1 policy, A value, A+1 reward; HP = round_up(A+2, 4).
struct PackLayout {
    int H, L, A, HP;
    size_t emb, layers, heads_w, heads_b, total;  // float offsets
    CM_HOST_DEVICE size_t layer_k(int l) const { return layers + (size_t)l * ((size_t)H * H + H); }
    CM_HOST_DEVICE size_t layer_b(int l) const { return layer_k(l) + (size_t)H * H; }
};

inline PackLayout make_pack_layout(int H, int L, int A) {
    PackLayout p;
    p.H = H; p.L = L; p.A = A;
    p.HP = (A + 2 + 3) & ~3;
    size_t o = 0;
    auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };
    p.emb = o;     o = al(o + (size_t)A * H);
    p.layers = o;  o = al(o + (size_t)L * ((size_t)H * H + H));
    p.heads_w = o; o = al(o + (size_t)H * p.HP);
    p.heads_b = o; o = al(o + (size_t)p.HP);
    p.total = o;
    return p;
}

// R = neurons per lane of the fused kernel's warp-cooperative network tile: H <= 32 -> 1, else the
// smallest of {2,4,8} with 32*R >= H and H % R == 0.  0 = no mapping (H % 4 != 0 or H > 256).
inline int pick_neurons_per_lane(int H) {
    if (H % 4 != 0) return 0;
    if (H <= 32) return 1;
    const int cand[3] = {2, 4, 8};
    for (int i = 0; i < 3; ++i)
        if (32 * cand[i] >= H && H % cand[i] == 0) return cand[i];
    return 0;
}

// Parameters as the FUSED search kernel wants them in shared memory ("fpack"), one bulk copy:
//   emb [A][H] | L x ( Wb_l | b_l [H] ) | Wht [HP][HS] | bh [HP]
// Wb_l is the layer kernel in lane-blocked order: the lane that owns neurons j0 = lane*R .. j0+R-1
// finds, for every block of 4 consecutive k, its 4*R weights in R 16-byte chunks laid out
// [k/4][chunk][lane][4] (float index f = (k%4)*R + v -> chunk f/4, slot f%4), so each LDS.128 is
// conflict-free and all offsets from the lane's base pointer are compile-time constants.
// Lanes with j0 >= H hold zeros.  Wh is the head matrix row-major [H][HP] (HP = round_up(A+2, 4)):
// a lane that owns an output pair reads its two weights of row k with one LDS.64.
struct FusedPackLayout {
    int H, L, A, HP, R, HS;
    size_t emb, layers, layer_stride, heads_w, heads_b, total;  // float offsets
    CM_HOST_DEVICE size_t layer_w(int l) const { return layers + (size_t)l * layer_stride; }
    CM_HOST_DEVICE size_t layer_b(int l) const { return layer_w(l) + (size_t)32 * R * H; }
    CM_HOST_DEVICE static size_t w_index(int k, int lane, int v, int R) {
        const int f = (k & 3) * R + v;
        return ((((size_t)(k >> 2) * R + (f >> 2)) * 32 + lane) << 2) + (f & 3);
    }
};

inline FusedPackLayout make_fused_pack_layout(int H, int L, int A) {
    FusedPackLayout p;
    p.H = H; p.L = L; p.A = A;
    p.HP = (A + 2 + 3) & ~3;
    p.R = pick_neurons_per_lane(H);
    p.HS = p.HP;
    auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };
    const int R = p.R > 0 ? p.R : 1;
    size_t o = 0;
    p.emb = o;     o = al(o + (size_t)A * H);
    p.layer_stride = al((size_t)32 * R * H + H);
    p.layers = o;  o = al(o + (size_t)L * p.layer_stride);
    p.heads_w = o; o = al(o + (size_t)H * p.HP);
    p.heads_b = o; o = al(o + (size_t)p.HP);
    p.total = o;
    return p;
}

// Parameters as the TEAM search kernel (search_team.cu) wants them in shared memory ("tpack"):
//   emb [A][H] | L x ( W_l [H][Hp] | b_l [Hp] ) | Wh [H][HP] | bh [HP] | WhT [HP][H]
// (WhT = the head matrix transposed: a head chain reads four consecutive k of its column with one LDS.128.)
// W_l is the layer kernel k-major with the output neurons padded to Hp = round_up(H, 64) (zeros): the
// lane that owns neurons j..j+NV-1 reads them with one LDS.64/128 per k, conflict-free across the warp.
// (dup = 1 stores every weight twice, {w,w}: packed-FFMA2 operand form, measured slower on B200 and unused.)
struct TeamPackLayout {
    int H, L, A, HP, Hp, dup;
    size_t emb, layers, layer_stride, heads_w, heads_b, heads_wt, total;  // float offsets
    CM_HOST_DEVICE size_t layer_w(int l) const { return layers + (size_t)l * layer_stride; }
    CM_HOST_DEVICE size_t layer_b(int l) const { return layer_w(l) + (size_t)H * Hp * (dup ? 2 : 1); }
};

inline TeamPackLayout make_team_pack_layout(int H, int L, int A, int dup) {
    TeamPackLayout p;
    p.H = H; p.L = L; p.A = A; p.dup = dup;
    p.HP = (A + 2 + 3) & ~3;
    p.Hp = (H + 63) & ~63;
    auto al = [](size_t x) { return (x + 3) & ~(size_t)3; };
    size_t o = 0;
    p.emb = o;     o = al(o + (size_t)A * H);
    p.layer_stride = al((size_t)H * p.Hp * (dup ? 2 : 1) + p.Hp);
    p.layers = o;  o = al(o + (size_t)L * p.layer_stride);
    p.heads_w = o; o = al(o + (size_t)H * p.HP);
    p.heads_b = o; o = al(o + (size_t)p.HP);
    p.heads_wt = o; o = al(o + (size_t)p.HP * H);
    p.total = o;
    return p;
}

// Parameters for the tensor-core (bf16) network kernel: bf16 weight tiles in the UMMA K-major
// core-matrix layout ([k/8][n/8][n%8][k%8], one tile per layer with rows n = output neuron, then the
// head tile with NP = 16 rows: policy 0..A-1, value A, reward A+1, zero padded), followed by the
// fp32 biases and the fp32 embedding.  Byte offsets, every section 128-byte aligned.
struct TcPackLayout {
    int H, L, A, NP;
    size_t layer_off, heads_off, bias_off, bh_off, emb_off, total_bytes;
};

inline TcPackLayout make_tc_pack_layout(int H, int L, int A) {
    TcPackLayout p;
    p.H = H; p.L = L; p.A = A; p.NP = 16;
    auto al = [](size_t x) { return (x + 127) & ~(size_t)127; };
    size_t o = 0;
    p.layer_off = o; o = al(o + (size_t)L * H * H * 2);
    p.heads_off = o; o = al(o + (size_t)p.NP * H * 2);
    p.bias_off = o;  o = al(o + (size_t)L * H * 4);
    p.bh_off = o;    o = al(o + (size_t)p.NP * 4);
    p.emb_off = o;   o = al(o + (size_t)A * H * 4);
    p.total_bytes = o;
    return p;
}

// Per-layer sections of the tensor-core conv pack: bf16 weight tile [Kpad/8][Cout/8][8][8]
// (row = output channel, k = (ky*3+kx)*Cin + ci, zero padded to Kpad), then fp32 gain[Cout] and
// shift[Cout] (conv bias and inference BatchNorm folded: y = acc*gain + shift).
// one record per conv output channel: where gain / shift go (4-byte units inside the conv pack) and
// which blob entries they are folded from (scale = -1: layer without BatchNorm)
struct BnFold {
    int gain_dst, shift_dst, bias, scale, beta, mean, var;
};

struct ConvTcPackLayout {
    static constexpr int kMaxLayers = 17;  // 1 + 2*L, L <= 8
    int n_layers;
    int K[kMaxLayers], Kpad[kMaxLayers], cin[kMaxLayers];
    size_t off[kMaxLayers], bytes[kMaxLayers], gain_off[kMaxLayers];
    size_t total_bytes;
};

// ------------------------------------------------------------------------------------------------
// tree storage
// ------------------------------------------------------------------------------------------------
// One 16-byte record per node: a node's A children are contiguous, so one 32-byte sector serves
// two children (all of them when A == 2) and one LDG.128 fetches everything ucb_score needs.
struct __align__(16) Node {
    double value_sum;  // Node.value_sum  (float64, mcts.py:29)
    int32_t visit;     // Node.visit_count
    float prior;       // Node.prior as the fp32 softmax produced it (root children: see root_prior)
};
static_assert(sizeof(Node) == 16, "Node record must be 16 bytes");

struct TreeView {
    int32_t B, S_max, A, H, N;
    Node* nodes;          // [B][N]
    int32_t* child_eidx;  // [B][N]   -1 = not expanded, else expansion index e
    double* root_prior;   // [B][A]   root children priors after noise mixing (float64)
    int32_t* exp_node;    // [B][S_max+1]
    float* hidden;        // [B][S_max+1][H]
    int32_t* path;        // [B][S_max+2]
    int32_t* path_len;    // [B]
    int32_t* leaf;        // [B]
    int32_t* n_expanded;  // [B]
    long long* depth_sum; // [B]
    float* reward;        // [B][N]   CUMIND_TREE_MINMAX_DISCOUNT only
    double* minmax;       // [B][2]   CUMIND_TREE_MINMAX_DISCOUNT only
};

struct NetView {
    cumind_net_desc desc;
    const float* blob;   // canonical fp32 blob (device)
    BlobLayout lo;
    const float* pack;   // packed recurrent-path parameters (device, 16B aligned)
    PackLayout pk;
};

// ------------------------------------------------------------------------------------------------
// pinned scalar numerics (device)
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
CM_DEVICE float fmul(float a, float b) { return __fmul_rn(a, b); }
CM_DEVICE float fadd(float a, float b) { return __fadd_rn(a, b); }
CM_DEVICE float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// Two independent IEEE fmas in one issue slot (SASS FFMA2): acc.{x,y} = fma(x, w.{x,y}, acc.{x,y}).
CM_DEVICE float2 ffma2(float x, float2 w, float2 acc) {
    unsigned long long xx, ww, aa, dd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ww) : "f"(w.x), "f"(w.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(acc.x), "f"(acc.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(xx), "l"(ww), "l"(aa));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(dd));
    return r;
}
// Fully packed form: d.{x,y} = fma(a.{x,y}, b.{x,y}, c.{x,y}).
CM_DEVICE float2 ffma2v(float2 a, float2 b, float2 c) {
    unsigned long long aa, bb, cc, dd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(dd) : "l"(aa), "l"(bb), "l"(cc));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(dd));
    return r;
}
CM_DEVICE float frelu(float v) { return v > 0.0f ? v : 0.0f; }

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may start while the previous kernel of the stream is still running; pdl_wait() blocks until that kernel has
// completed and its writes are visible (a no-op for a normal launch), pdl_launch_dependents() lets the next
// kernel of the stream start its prologue early.
CM_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
CM_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// exp(d), d <= 0, float64, mul/add only — the same operation sequence as cmo_pexp (oracle) so the
// fp32-rounded result is bit-identical on CPU and GPU.
CM_DEVICE double pexp(double d) {
    if (d != d) return d;
    if (!(d >= -200.0)) return 0.0;
    if (d > 0.0) d = 0.0;
    const double LOG2E = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    double kf = rint(__dmul_rn(d, LOG2E));
    double r = __dsub_rn(__dsub_rn(d, __dmul_rn(kf, LN2_HI)), __dmul_rn(kf, LN2_LO));
    const double C[14] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                          1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0,
                          1.0 / 479001600.0, 1.0 / 6227020800.0};
    double p = C[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), C[i]);
    long long k = (long long)kf;
    double scale = __longlong_as_double((k + 1023) << 52);
    return __dmul_rn(p, scale);
}

// pexp without branches (same operations on the same values, the special cases selected at the end): keeps
// the instruction stream of a latency-bound caller straight.
CM_DEVICE double pexp_nb(double d) {
    const bool is_nan = d != d;
    const bool vanishes = !(d >= -200.0);  // also true for NaN, which is restored below
    double x = d > 0.0 ? 0.0 : d;
    x = vanishes ? 0.0 : x;
    const double LOG2E = 1.4426950408889634;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double kf = rint(__dmul_rn(x, LOG2E));
    const double r = __dsub_rn(__dsub_rn(x, __dmul_rn(kf, LN2_HI)), __dmul_rn(kf, LN2_LO));
    const double C[14] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0,
                          1.0 / 40320.0, 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0,
                          1.0 / 479001600.0, 1.0 / 6227020800.0};
    double p = C[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = __dadd_rn(__dmul_rn(p, r), C[i]);
    const int k = __double2int_rn(kf);  // integral, -289 <= k <= 0
    const double scale = __hiloint2double((k + 1023) << 20, 0);
    double res = __dmul_rn(p, scale);
    res = vanishes ? 0.0 : res;
    return is_nan ? d : res;
}

// Node.ucb_score (mcts.py:51-65): inf if n == 0 else W/n + ((c*p)*sqrt(N))/(1+n), float64.
// `sqrtN` is __dsqrt_rn((double)N_parent) (correctly rounded, same as math.sqrt).
// a / d for a small positive integer d, correctly rounded, from r = RN(1/d):
//   q0 = RN(a*r); two FMA residual corrections (Markstein): after the first q1 is faithful, and a
//   faithful quotient corrected once more with an exact residual and r = RN(1/d) is RN(a/d)
//   (the exception, a divisor significand of all ones, cannot occur for d < 2^53 - 1).
// Outside the range where residuals are exact (|a| not in [2^-900, 2^900], a == 0, inf, NaN) it
// falls back to the IEEE division.  tests/test_gpu_parity.py::test_small_divisor_division_is_ieee
// checks it against __ddiv_rn on the device.
CM_DEVICE double div_small(double a, double d, double r) {
    double q = __dmul_rn(a, r);
    q = __fma_rn(__fma_rn(-d, q, a), r, q);
    q = __fma_rn(__fma_rn(-d, q, a), r, q);
    const double m = fabs(a);
    if (!(m >= 1.1830521861667747e-271 && m <= 8.452712498170644e+270))  // 2^-900 .. 2^900
        q = (a == 0.0) ? a : __ddiv_rn(a, d);
    return q;
}

// Node.ucb_score with the reciprocal table: recip[i] = RN(1/i).
CM_DEVICE double ucb_score_tab(int32_t n, double W, double prior, double sqrtN, double c_puct,
                               const double* __restrict__ recip) {
    if (n == 0) return __longlong_as_double(0x7FF0000000000000LL);
    const double explo = div_small(__dmul_rn(__dmul_rn(c_puct, prior), sqrtN), (double)(1 + n), recip[1 + n]);
    return __dadd_rn(div_small(W, (double)n, recip[n]), explo);
}

// Branch-free variant for the fused select: an unvisited child divides by 1 (recip[1] == 1.0 exactly)
// and the +inf is selected at the end, so the level has no divergent branch.
CM_DEVICE double ucb_score_tab_nobranch(int32_t n, double W, double prior, double sqrtN, double c_puct,
                                        const double* __restrict__ recip) {
    const bool unvisited = n == 0;
    const int nn = unvisited ? 1 : n;
    const double num = __dmul_rn(__dmul_rn(c_puct, prior), sqrtN);
    const double explo = div_small(num, (double)(1 + n), recip[1 + n]);
    const double q = div_small(unvisited ? 0.0 : W, (double)nn, recip[nn]);
    const double s = __dadd_rn(q, explo);
    return unvisited ? __longlong_as_double(0x7FF0000000000000LL) : s;
}

// The compiler if-converts the n == 0 case, so the divisions always execute: feed them 1/1 then
// (0/0 would take the slow special-case path of the IEEE division routine on every unvisited child).
CM_DEVICE double ucb_score(int32_t n, double W, double prior, double sqrtN, double c_puct) {
    const bool unvisited = n == 0;
    const double nn = unvisited ? 1.0 : (double)n;
    const double ww = unvisited ? 1.0 : W;
    const double num = unvisited ? 1.0 : __dmul_rn(__dmul_rn(c_puct, prior), sqrtN);  // sqrt(0) = 0 under a fresh parent
    const double explo = __ddiv_rn(num, (double)(1 + n));
    const double s = __dadd_rn(__ddiv_rn(ww, nn), explo);
    return unvisited ? __longlong_as_double(0x7FF0000000000000LL) : s;
}
#endif

}  // namespace cm
