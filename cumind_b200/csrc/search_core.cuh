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
//   select_leaf   <- MCTS._simulate selection loop 165-171, Node.select_child 67-76, ucb_score 51-65
//   dyn_layer ... <- DynamicsNetwork.__call__ network.py:217-236, PredictionNetwork.__call__ 254-266
//   expand_leaf   <- Node.expand 78-89 (children get the fp32 softmax priors)
//   backup_path   <- mcts.py:199-203 (parents only; the root a second time)
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
template <int G>
CM_DEVICE int select_leaf(const Node* __restrict__ nodes, const int32_t* __restrict__ eidx,
                          const double* __restrict__ root_prior, const double* __restrict__ sqrt_tab,
                          int A, double c_puct, int root_visits, int* __restrict__ path, int g, bool active,
                          int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, n_parent = root_visits, plen = 0, action = 0, parent_e = 0;
    bool walking = active;
    while (__any_sync(0xffffffffu, walking)) {
        double best_s = -dinf();
        int best_a = 0x7fffffff, best_n = 0, best_e = -1;
        if (walking) {
            const int base = 1 + e * A;
            const double sq = sqrt_tab[n_parent];
            for (int a = g; a < A; a += G) {
                const int c = base + a;
                const Node nd = load_node(nodes + c);
                const int ce = eidx[c];
                const double p = (node == 0) ? root_prior[a] : (double)nd.prior;
                double s = ucb_score(nd.visit, nd.value_sum, p, sq, c_puct);
                if (s != s) s = (a == 0) ? dinf() : -dinf();
                if (best_a == 0x7fffffff || s > best_s) { best_s = s; best_a = a; best_n = nd.visit; best_e = ce; }
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best_s, off);
            const int oa = __shfl_xor_sync(0xffffffffu, best_a, off);
            const int on = __shfl_xor_sync(0xffffffffu, best_n, off);
            const int oe = __shfl_xor_sync(0xffffffffu, best_e, off);
            if (os > best_s || (os == best_s && oa < best_a)) { best_s = os; best_a = oa; best_n = on; best_e = oe; }
        }
        if (walking) {
            if (g == 0) path[plen] = node;
            ++plen;
            parent_e = e;
            action = best_a;
            node = 1 + e * A + best_a;
            n_parent = best_n;
            e = best_e;
            walking = e >= 0;
        }
    }
    plen_out = plen;
    parent_e_out = parent_e;
    action_out = action;
    return node;
}

// Backup (mcts.py:199-203): every parent on the path gets n += 1, W += v; the root (always
// path[0]) is backed up a second time, sequentially: W = (W + v) + v.
template <int G>
CM_DEVICE void backup_path(Node* __restrict__ nodes, const int* __restrict__ path, int plen, float leaf_value, int g) {
    const double v = (double)leaf_value;
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
// Network, fp32-exact mode.  Lane g of the group owns output chunks c = g + G*i (i < R/V) of V
// consecutive neurons; each dot product is sequential in k with unfused multiply and add, exactly
// as oracle/cumind_oracle.c::cmo_linear.  x lives in shared memory (one row per tree).
// ------------------------------------------------------------------------------------------------
template <int R> struct VecW { static constexpr int V = R >= 4 ? 4 : R; };

template <int V> CM_DEVICE void load_vec(const float* p, float (&w)[V]);
template <> CM_DEVICE void load_vec<1>(const float* p, float (&w)[1]) { w[0] = p[0]; }
template <> CM_DEVICE void load_vec<2>(const float* p, float (&w)[2]) {
    float2 t = *reinterpret_cast<const float2*>(p); w[0] = t.x; w[1] = t.y;
}
template <> CM_DEVICE void load_vec<4>(const float* p, float (&w)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
}
template <int V> CM_DEVICE void store_vec(float* p, const float (&w)[V]);
template <> CM_DEVICE void store_vec<1>(float* p, const float (&w)[1]) { p[0] = w[0]; }
template <> CM_DEVICE void store_vec<2>(float* p, const float (&w)[2]) { *reinterpret_cast<float2*>(p) = make_float2(w[0], w[1]); }
template <> CM_DEVICE void store_vec<4>(float* p, const float (&w)[4]) { *reinterpret_cast<float4*>(p) = make_float4(w[0], w[1], w[2], w[3]); }

// acc[R] = sum_k x[k] * W[k][own outputs]   (sequential k, unfused)
template <int G, int R>
CM_DEVICE void dot_rows(const float* __restrict__ W /*[H][H]*/, const float* __restrict__ x /*[H]*/, int H, int g,
                        float (&acc)[R]) {
    constexpr int V = VecW<R>::V;
    constexpr int NCH = R / V;
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = 0.0f;
    if constexpr (R == 1) {
        if (g < H) {
            for (int k = 0; k < H; ++k) acc[0] = fadd(acc[0], fmul(x[k], W[(size_t)k * H + g]));
        }
    } else {
#pragma unroll 2
        for (int k4 = 0; k4 < H; k4 += 4) {
            float xv[4];
            load_vec<4>(x + k4, xv);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float* wrow = W + (size_t)(k4 + kk) * H;
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    float w[V];
                    load_vec<V>(wrow + (g + G * i) * V, w);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[i * V + v] = fadd(acc[i * V + v], fmul(xv[kk], w[v]));
                }
            }
        }
    }
}

// x_out = relu(x_in @ K + b) + x_in   (one dynamics layer, network.py:230-233)
template <int G, int R>
CM_DEVICE void dyn_layer(const float* __restrict__ K, const float* __restrict__ b, const float* __restrict__ xin,
                         float* __restrict__ xout, int H, int g) {
    constexpr int V = VecW<R>::V;
    constexpr int NCH = R / V;
    float acc[R];
    dot_rows<G, R>(K, xin, H, g, acc);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j0 = (g + G * i) * V;
        if (R > 1 || j0 < H) {
            float bb[V], xo[V], y[V];
            load_vec<V>(b + j0, bb);
            load_vec<V>(xin + j0, xo);
#pragma unroll
            for (int v = 0; v < V; ++v) y[v] = fadd(frelu(fadd(acc[i * V + v], bb[v])), xo[v]);
            store_vec<V>(xout + j0, y);
        }
    }
}

// x = h + Emb[action]   (network.py:226-227)
template <int G, int R>
CM_DEVICE void embed_add(const float* __restrict__ h /*global*/, const float* __restrict__ emb_row /*smem*/,
                         float* __restrict__ xout, int H, int g) {
    constexpr int V = VecW<R>::V;
    constexpr int NCH = R / V;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j0 = (g + G * i) * V;
        if (R > 1 || j0 < H) {
            float hv[V], ev[V], y[V];
            load_vec<V>(h + j0, hv);
            load_vec<V>(emb_row + j0, ev);
#pragma unroll
            for (int v = 0; v < V; ++v) y[v] = fadd(hv[v], ev[v]);
            store_vec<V>(xout + j0, y);
        }
    }
}

template <int G, int R>
CM_DEVICE void copy_row(const float* __restrict__ src, float* __restrict__ dst, int H, int g) {
    constexpr int V = VecW<R>::V;
    constexpr int NCH = R / V;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int j0 = (g + G * i) * V;
        if (R > 1 || j0 < H) {
            float t[V];
            load_vec<V>(src + j0, t);
            store_vec<V>(dst + j0, t);
        }
    }
}

// Heads: out[o] = x @ Wh[:,o] + bh[o] for o < n_out (columns: policy 0..A-1, value A, reward A+1).
template <int G>
CM_DEVICE void heads(const float* __restrict__ Wh /*[H][HP]*/, const float* __restrict__ bh, const float* __restrict__ x,
                     int H, int HP, int n_out, float* __restrict__ out, int g) {
    for (int o = g; o < n_out; o += G) {
        float acc = 0.0f;
        for (int k = 0; k < H; ++k) acc = fadd(acc, fmul(x[k], Wh[(size_t)k * HP + o]));
        out[o] = fadd(acc, bh[o]);
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
