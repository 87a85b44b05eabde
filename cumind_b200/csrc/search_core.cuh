// search_core.cuh — per-tree device routines of the MuZero search, shared by the fused whole-search
// kernel (search_fused.cu) and the one-phase-per-launch kernels (tree_steps.cu).
//
// Thread mapping: a tree is owned by a group of G consecutive lanes of one warp (G = 4..32, a
// launch-time tuning knob: G = 32 is the "warp per tree" of the design brief, smaller G packs
// several trees into a warp when B is large).  All cross-lane traffic is warp shuffles inside the
// group; phases are separated by __syncwarp(), never by a block or grid barrier, so every tree
// advances through its S simulations independently of all others.
//
// Reference semantics reproduced here (src/cumind/core/mcts.py):
//   select_leaf*            <- MCTS._simulate selection loop 165-171, Node.select_child 67-76, ucb_score 51-65
//   warp_layer / warp_heads <- DynamicsNetwork.__call__ network.py:217-236, PredictionNetwork.__call__ 254-266
//   softmax_prepare         <- jax.nn.softmax mcts.py:128,191 (children get the fp32 priors, Node.expand 78-89)
//   backup_path             <- mcts.py:199-203 (parents only; the root a second time)
#pragma once

#include "common.cuh"

namespace cm {

CM_DEVICE Node load_node(const Node* p) {
    int4 raw = *reinterpret_cast<const int4*>(p);
    Node n;
    n.value_sum = __hiloint2double(raw.y, raw.x);
    n.visit = raw.z;
    n.prior = __int_as_float(raw.w);
    return n;
}

CM_DEVICE void store_node(Node* p, double value_sum, int32_t visit, float prior) {
    int4 raw;
    raw.x = __double2loint(value_sum);
    raw.y = __double2hiint(value_sum);
    raw.z = visit;
    raw.w = __float_as_int(prior);
    *reinterpret_cast<int4*>(p) = raw;
}

CM_DEVICE double dinf() { return __longlong_as_double(0x7FF0000000000000LL); }

// ------------------------------------------------------------------------------------------------
// Selection: walk from the root while the node is expanded.  Each lane scores the children
// a = g, g+G, ... of the current node; the group arg-max is a shuffle butterfly carrying
// (score, action, visit count, expansion index) so the next level needs no dependent load.
// Tie-break = lowest action (Python max() keeps the first maximum, mcts.py:76).  A NaN score
// never wins unless it is child 0 (Python: `nan > x` is False), mapped here to -inf / +inf.
// Returns the leaf slot; `parent_e`/`action` identify the edge that reaches it; the parents on the
// path are written to path[0..plen).
// ------------------------------------------------------------------------------------------------
// `recip` (RN(1/i) table) selects the table-driven division of ucb_score_tab; nullptr = IEEE division.
template <int G, bool ONE, bool TAB>
CM_DEVICE int select_leaf_impl(const Node* __restrict__ nodes, const int32_t* __restrict__ eidx,
                               const double* __restrict__ root_prior, const double* __restrict__ sqrt_tab,
                               const double* __restrict__ recip,
                               int A, double c_puct, int root_visits, int* __restrict__ path, int g, bool active,
                               int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, n_parent = root_visits, plen = 0, action = 0, parent_e = 0;
    bool walking = active;
    // lanes g >= A hold no child: only the first `span` lanes of the group take part in the butterfly
    int span = 1;
    while (span < A && span < G) span <<= 1;
    // ONE: A <= G, exactly one child per lane -> no inner loop, per-lane base pointers
    const bool has_child = g < A;
    const Node* my_nodes = nodes + 1 + g;
    const int32_t* my_eidx = eidx + 1 + g;
    double my_root_prior = 0.0;
    if (ONE && has_child) my_root_prior = root_prior[g];
    while (__any_sync(0xffffffffu, walking)) {
        double best_s = -dinf();
        // payload: action in the high half, expansion index + 1 in the low half (both < 65536)
        unsigned best_ae = 0xffffffffu;
        int best_n = 0;
        if constexpr (ONE) {
            if (walking && has_child) {
                const int off = e * A;
                const Node nd = load_node(my_nodes + off);
                const int ce = my_eidx[off];
                const double p = (node == 0) ? my_root_prior : (double)nd.prior;
                double s = TAB ? ucb_score_tab(nd.visit, nd.value_sum, p, sqrt_tab[n_parent], c_puct, recip)
                               : ucb_score(nd.visit, nd.value_sum, p, sqrt_tab[n_parent], c_puct);
                if (s != s) s = (g == 0) ? dinf() : -dinf();
                best_s = s; best_ae = ((unsigned)g << 16) | (unsigned)(ce + 1); best_n = nd.visit;
            }
        } else {
            if (walking) {
                const int base = 1 + e * A;
                const double sq = sqrt_tab[n_parent];
                for (int a = g; a < A; a += G) {
                    const int c = base + a;
                    const Node nd = load_node(nodes + c);
                    const int ce = eidx[c];
                    const double p = (node == 0) ? root_prior[a] : (double)nd.prior;
                    double s = TAB ? ucb_score_tab(nd.visit, nd.value_sum, p, sq, c_puct, recip)
                                   : ucb_score(nd.visit, nd.value_sum, p, sq, c_puct);
                    if (s != s) s = (a == 0) ? dinf() : -dinf();
                    if (best_ae == 0xffffffffu || s > best_s) {
                        best_s = s; best_ae = ((unsigned)a << 16) | (unsigned)(ce + 1); best_n = nd.visit;
                    }
                }
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            if (off < span) {
                const double os = __shfl_xor_sync(0xffffffffu, best_s, off);
                const unsigned oae = __shfl_xor_sync(0xffffffffu, best_ae, off);
                const int on = __shfl_xor_sync(0xffffffffu, best_n, off);
                if (os > best_s || (os == best_s && oae < best_ae)) { best_s = os; best_ae = oae; best_n = on; }
            }
        }
        if (span < G) {  // lanes beyond the butterfly take the result from the group's lane 0
            best_ae = __shfl_sync(0xffffffffu, best_ae, 0, G);
            best_n = __shfl_sync(0xffffffffu, best_n, 0, G);
        }
        if (walking) {
            if (g == 0) path[plen] = node;
            ++plen;
            parent_e = e;
            action = (int)(best_ae >> 16);
            node = 1 + e * A + action;
            n_parent = best_n;
            e = (int)(best_ae & 0xffffu) - 1;
            walking = e >= 0;
        }
    }
    plen_out = plen;
    parent_e_out = parent_e;
    action_out = action;
    return node;
}

template <int G>
CM_DEVICE int select_leaf(const Node* __restrict__ nodes, const int32_t* __restrict__ eidx,
                          const double* __restrict__ root_prior, const double* __restrict__ sqrt_tab,
                          int A, double c_puct, int root_visits, int* __restrict__ path, int g, bool active,
                          int& plen_out, int& parent_e_out, int& action_out) {
    if (A <= G)
        return select_leaf_impl<G, true, false>(nodes, eidx, root_prior, sqrt_tab, nullptr, A, c_puct, root_visits, path, g,
                                                active, plen_out, parent_e_out, action_out);
    return select_leaf_impl<G, false, false>(nodes, eidx, root_prior, sqrt_tab, nullptr, A, c_puct, root_visits, path, g,
                                             active, plen_out, parent_e_out, action_out);
}

// Fused-kernel selection for A <= G: one child per lane, SPAN = pow2 >= A lanes in the butterfly,
// no divergent branch inside a level (lanes without a child and trees that already reached their
// leaf alias valid records and are masked out of the arg-max), table-driven exact division.
template <int G, int SPAN>
CM_DEVICE int select_leaf_fast(const Node* __restrict__ nodes, const int32_t* __restrict__ eidx,
                               const double* __restrict__ root_prior, const double* __restrict__ sqrt_tab,
                               const double* __restrict__ recip, int A, double c_puct, int root_visits,
                               int* __restrict__ path, int g, bool active, int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, n_parent = root_visits, plen = 0, action = 0, parent_e = 0;
    int walking = active ? 1 : 0;
    const bool has_child = g < A;
    const int cg = has_child ? g : 0;
    const Node* my_nodes = nodes + 1 + cg;
    const int32_t* my_eidx = eidx + 1 + cg;
    const double my_root_prior = root_prior[cg];
    while (__any_sync(0xffffffffu, walking != 0)) {
        const int off = walking ? e * A : 0;
        const Node nd = load_node(my_nodes + off);
        const int ce = my_eidx[off];
        const double p = (node == 0) ? my_root_prior : (double)nd.prior;
        double s = ucb_score_tab_nobranch(nd.visit, nd.value_sum, p, sqrt_tab[n_parent], c_puct, recip);
        if (s != s) s = (g == 0) ? dinf() : -dinf();
        const bool valid = has_child && walking;
        double best_s = valid ? s : -dinf();
        unsigned best_ae = valid ? (((unsigned)g << 16) | (unsigned)(ce + 1)) : 0xffffffffu;
        int best_n = nd.visit;
#pragma unroll
        for (int o = SPAN / 2; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
            const unsigned oae = __shfl_xor_sync(0xffffffffu, best_ae, o);
            const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
            const bool take = os > best_s || (os == best_s && oae < best_ae);
            best_s = take ? os : best_s;
            best_ae = take ? oae : best_ae;
            best_n = take ? on : best_n;
        }
        if constexpr (SPAN < G) {
            best_ae = __shfl_sync(0xffffffffu, best_ae, 0, G);
            best_n = __shfl_sync(0xffffffffu, best_n, 0, G);
        }
        if (walking && g == 0) path[plen] = node;
        const int act = (int)(best_ae >> 16);
        const int ne = (int)(best_ae & 0xffffu) - 1;
        plen += walking;
        parent_e = walking ? e : parent_e;
        action = walking ? act : action;
        node = walking ? 1 + e * A + act : node;
        n_parent = walking ? best_n : n_parent;
        e = walking ? ne : e;
        walking = (walking && ne >= 0) ? 1 : 0;
    }
    plen_out = plen;
    parent_e_out = parent_e;
    action_out = action;
    return node;
}

// fused kernel: reciprocal-table division
template <int G>
CM_DEVICE int select_leaf_tab(const Node* __restrict__ nodes, const int32_t* __restrict__ eidx,
                              const double* __restrict__ root_prior, const double* __restrict__ sqrt_tab,
                              const double* __restrict__ recip, int A, double c_puct, int root_visits,
                              int* __restrict__ path, int g, bool active, int& plen_out, int& parent_e_out, int& action_out) {
#define CM_SELECT_FAST(SP)                                                                                                   \
    if constexpr (SP <= G)                                                                                                   \
        if (A <= SP)                                                                                                         \
            return select_leaf_fast<G, SP>(nodes, eidx, root_prior, sqrt_tab, recip, A, c_puct, root_visits, path, g, active, \
                                           plen_out, parent_e_out, action_out);
    CM_SELECT_FAST(1)
    CM_SELECT_FAST(2)
    CM_SELECT_FAST(4)
    CM_SELECT_FAST(8)
    CM_SELECT_FAST(16)
    CM_SELECT_FAST(32)
#undef CM_SELECT_FAST
    return select_leaf_impl<G, false, true>(nodes, eidx, root_prior, sqrt_tab, recip, A, c_puct, root_visits, path, g, active,
                                            plen_out, parent_e_out, action_out);
}

// Backup (mcts.py:199-203): every parent on the path gets n += 1, W += v; the root (always
// path[0]) is backed up a second time, sequentially: W = (W + v) + v.
template <int G>
CM_DEVICE void backup_path(Node* __restrict__ nodes, const int* __restrict__ path, int plen, float leaf_value, int g) {
    const double v = (double)leaf_value;
#pragma unroll 1
    for (int i = g; i < plen; i += G) {
        const int nd = path[i];
        Node r = load_node(nodes + nd);
        double w = __dadd_rn(r.value_sum, v);
        int n = r.visit + 1;
        if (nd == 0) { w = __dadd_rn(w, v); n += 1; }
        store_node(nodes + nd, w, n, r.prior);
    }
}

// ------------------------------------------------------------------------------------------------
// Network, fp32-exact mode: every dot product is a sequential-k chain of IEEE fused multiply-adds,
// exactly as oracle/cumind_oracle.c::cmo_linear.  Per-tree helpers (stepwise kernels) first, then the
// warp-cooperative tile of the fused kernel.
// ------------------------------------------------------------------------------------------------
template <int V> CM_DEVICE void load_vec(const float* p, float (&w)[V]);
template <> CM_DEVICE void load_vec<1>(const float* p, float (&w)[1]) { w[0] = p[0]; }
template <> CM_DEVICE void load_vec<2>(const float* p, float (&w)[2]) {
    float2 t = *reinterpret_cast<const float2*>(p); w[0] = t.x; w[1] = t.y;
}
template <> CM_DEVICE void load_vec<4>(const float* p, float (&w)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
}
template <> CM_DEVICE void load_vec<8>(const float* p, float (&w)[8]) {
    float4 t = *reinterpret_cast<const float4*>(p), u = *reinterpret_cast<const float4*>(p + 4);
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w; w[4] = u.x; w[5] = u.y; w[6] = u.z; w[7] = u.w;
}
template <int V> CM_DEVICE void store_vec(float* p, const float (&w)[V]);
template <> CM_DEVICE void store_vec<8>(float* p, const float (&w)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(w[4], w[5], w[6], w[7]);
}
template <> CM_DEVICE void store_vec<1>(float* p, const float (&w)[1]) { p[0] = w[0]; }
template <> CM_DEVICE void store_vec<2>(float* p, const float (&w)[2]) { *reinterpret_cast<float2*>(p) = make_float2(w[0], w[1]); }
template <> CM_DEVICE void store_vec<4>(float* p, const float (&w)[4]) { *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]); }

// Heads: out[o] = x @ Wh[:,o] + bh[o] for o < n_out (columns: policy 0..A-1, value A, reward A+1).
template <int G>
CM_DEVICE void heads(const float* __restrict__ Wh /*[H][HP]*/, const float* __restrict__ bh, const float* __restrict__ x,
                     int H, int HP, int n_out, float* __restrict__ out, int g) {
    for (int o = g; o < n_out; o += G) {
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < H; ++k) acc = ffma(x[k], Wh[(size_t)k * HP + o], acc);
        out[o] = fadd(acc, bh[o]);
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-cooperative network tile (fused kernel).  The TW trees of a warp share every weight load:
// lane owns R consecutive neurons j0 = lane*R.. for ALL TW trees, the activations live in shared
// memory k-major with the trees interleaved (x[k][t]) so one vector load fetches x_k of every tree.
// Per k: one LDS of R weights + one broadcast LDS of TW activations feed TW*R fused multiply-adds,
// issued two per instruction (FFMA2: packed over neuron pairs, or over tree pairs when R == 1).
// Same sequential-k FMA chain per (tree, neuron) as cmo_linear.
// ------------------------------------------------------------------------------------------------
template <int TW, int R>
struct WarpAcc {
    static constexpr int NP = (R >= 2) ? TW * (R / 2) : (TW >= 2 ? TW / 2 : 1);
    float2 v[NP];
    CM_DEVICE void zero() {
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] = make_float2(0.0f, 0.0f);
    }
    // value for tree t, neuron v_
    CM_DEVICE float get(int t, int v_) const {
        if constexpr (R >= 2) { const float2 q = v[t * (R / 2) + v_ / 2]; return (v_ & 1) ? q.y : q.x; }
        else if constexpr (TW >= 2) { const float2 q = v[t / 2]; return (t & 1) ? q.y : q.x; }
        else return v[0].x;
    }
};

// Wb: lane-blocked layer kernel (FusedPackLayout): chunk (kb, c) of this lane is at
// Wb + ((kb*R + c)*32 + lane)*4; float f = (k%4)*R + v of a k-block sits in chunk f/4, slot f%4.
template <int TW, int R>
CM_DEVICE void warp_dot(const float* __restrict__ Wb, const float* __restrict__ x /*[H][TW]*/, int H, int lane,
                        WarpAcc<TW, R>& acc) {
    acc.zero();
    const float* wp = Wb + lane * 4;
#pragma unroll 2
    for (int kb = 0; kb < H; kb += 4) {
        float wq[4 * R];
#pragma unroll
        for (int c = 0; c < R; ++c) {
            float t4[4];
            load_vec<4>(wp + (size_t)((kb >> 2) * R + c) * 128, t4);
#pragma unroll
            for (int i = 0; i < 4; ++i) wq[c * 4 + i] = t4[i];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float xk[TW];
            load_vec<TW>(x + (size_t)(kb + kk) * TW, xk);
            if constexpr (R >= 2) {
#pragma unroll
                for (int t = 0; t < TW; ++t)
#pragma unroll
                    for (int v = 0; v < R; v += 2)
                        acc.v[t * (R / 2) + v / 2] =
                            ffma2(xk[t], make_float2(wq[kk * R + v], wq[kk * R + v + 1]), acc.v[t * (R / 2) + v / 2]);
            } else if constexpr (TW >= 2) {
#pragma unroll
                for (int t = 0; t < TW; t += 2) acc.v[t / 2] = ffma2(wq[kk], make_float2(xk[t], xk[t + 1]), acc.v[t / 2]);
            } else {
                acc.v[0].x = ffma(xk[0], wq[kk], acc.v[0].x);
            }
        }
    }
}

// one dynamics layer for the warp's TW trees: xout[j][t] = relu(dot + b[j]) + xin[j][t]
template <int TW, int R>
CM_DEVICE void warp_layer(const float* __restrict__ Wb, const float* __restrict__ b, const float* __restrict__ xin,
                          float* __restrict__ xout, int H, int lane) {
    WarpAcc<TW, R> acc;
    warp_dot<TW, R>(Wb, xin, H, lane, acc);
    const int j0 = lane * R;
    if (j0 >= H) return;
#pragma unroll
    for (int v = 0; v < R; ++v) {
        float xr[TW], y[TW];
        load_vec<TW>(xin + (size_t)(j0 + v) * TW, xr);
        const float bb = b[j0 + v];
#pragma unroll
        for (int t = 0; t < TW; ++t) y[t] = fadd(frelu(fadd(acc.get(t, v), bb)), xr[t]);
        store_vec<TW>(xout + (size_t)(j0 + v) * TW, y);
    }
}

// heads for the warp's TW trees: lane <-> (tree t, output pair q); out[t*HP + 2q..2q+1] = x[:,t] @ Wh[:,2q..2q+1] + bh
// Wh is row-major [H][HP]: one LDS.64 fetches the pair's weights of row k, one FFMA2 advances both chains.
template <int TW>
CM_DEVICE void warp_heads(const float* __restrict__ Wh /*[H][HP]*/, const float* __restrict__ bh, const float* __restrict__ x,
                          int H, int HP, int n_out, float* __restrict__ out, int lane) {
    const int n_pairs = (n_out + 1) >> 1;  // HP is a multiple of 4 >= n_out, so a pair never leaves the row
    for (int idx = lane; idx < TW * n_pairs; idx += 32) {
        const int t = idx / n_pairs, q = idx - t * n_pairs;
        const float* xp = x + t;
        const float* wp = Wh + 2 * q;
        float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 1
        for (int kb = 0; kb < H; kb += 8) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                if (kk < 4 || kb + kk < H) {
                    const float2 w = *reinterpret_cast<const float2*>(wp + (size_t)(kb + kk) * HP);
                    acc = ffma2(xp[(size_t)(kb + kk) * TW], w, acc);
                }
            }
        }
        const float2 bb = *reinterpret_cast<const float2*>(bh + 2 * q);
        out[t * HP + 2 * q] = fadd(acc.x, bb.x);
        out[t * HP + 2 * q + 1] = fadd(acc.y, bb.y);
    }
}

// softmax over lg[0..A_net) with the pinned exp; e_buf is scratch [A_net].  Every lane of the group
// ends up with max and sum; own probabilities are read back with softmax_prob().
template <int G>
CM_DEVICE float softmax_prepare(const float* __restrict__ lg, float* __restrict__ e_buf, int A_net, int g, bool active) {
    float s = 0.0f;
    if (active) {
        float m = lg[0];
        for (int a = 1; a < A_net; ++a) { const float l = lg[a]; if (l > m) m = l; }
        for (int a = g; a < A_net; a += G) e_buf[a] = __double2float_rn(pexp((double)__fsub_rn(lg[a], m)));
    }
    __syncwarp();
    if (active) {
        for (int a = 0; a < A_net; ++a) s = fadd(s, e_buf[a]);
    }
    return s;
}

// e / s of the softmax; a saturated softmax produces exact zeros, which would send the IEEE
// division down its slow special-case path: 0 / s == 0 exactly for the finite s >= 1 of a softmax.
CM_DEVICE float prior_div(float e, float s) { return (e == 0.0f && s >= 1.0f) ? 0.0f : __fdiv_rn(e, s); }

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) and a Dirichlet(alpha) draw for the root noise
// (replaces np.random.dirichlet at mcts.py:218; statistical parity only).
// ------------------------------------------------------------------------------------------------
struct Philox {
    uint32_t ctr[4], key[2];
    uint32_t out[4];
    int have;
    CM_DEVICE void init(uint64_t seed, uint64_t stream) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0; ctr[1] = 0; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    static CM_DEVICE void block(const uint32_t (&c)[4], const uint32_t (&k)[2], uint32_t (&o)[4]) {
        uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
    }
    CM_DEVICE uint32_t next() {
        if (have == 0) {
            block(ctr, key, out);
            if (++ctr[0] == 0) ++ctr[1];
            have = 4;
        }
        return out[4 - have--];
    }
    // uniform in (0,1), 53 bits
    CM_DEVICE double uniform() {
        const uint64_t hi = next(), lo = next();
        const uint64_t x = ((hi << 32) | lo) >> 11;
        return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
    }
    CM_DEVICE double normal() {
        const double u1 = uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    // Gamma(alpha, 1): Marsaglia-Tsang for alpha' = alpha (+1 when alpha < 1), then the U^(1/alpha) boost.
    CM_DEVICE double gamma(double alpha) {
        const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        double g = d;
        for (int it = 0; it < 64; ++it) {
            const double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = uniform();
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
        }
        if (alpha < 1.0) g *= pow(uniform(), 1.0 / alpha);
        return g;
    }
};

// Dirichlet(alpha,...,alpha) of dimension A for one tree, written to out[0..A); serial (A is small).
CM_DEVICE void dirichlet_row(double* out, int A, double alpha, uint64_t seed, uint64_t tree_index) {
    Philox rng;
    rng.init(seed, tree_index);
    double s = 0.0;
    for (int a = 0; a < A; ++a) { const double gm = rng.gamma(alpha); out[a] = gm; s += gm; }
    if (!(s > 0.0)) { for (int a = 0; a < A; ++a) out[a] = 1.0 / (double)A; return; }
    for (int a = 0; a < A; ++a) out[a] = out[a] / s;
}

}  // namespace cm
