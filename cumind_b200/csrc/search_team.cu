// search_team.cu — MCTS.search (src/cumind/core/mcts.py:115-155) for B trees in ONE launch,
// warp-specialised: the tree walk and the network run on different warps of the same CTA, and the
// tree statistics live in shared memory while a tree is being searched.
//
// The per-tree work of one simulation is a serial chain (walk d levels of float64 PUCT -> dynamics
// MLP -> heads -> softmax -> backup).  search_fused.cu gives every warp the whole chain of its own
// few trees, so each phase runs at a fraction of a warp's lanes and an SM holds only a handful of
// warps.  Here a CTA is split into TEAMS; a team owns up to TT trees (a batch) and has two kinds of warps:
//
//   tree warps     G = pow2 >= A lanes per tree (32/G trees per warp): select_child / ucb_score
//                  (mcts.py:51-76, 165-171), softmax -> Node.expand (78-89, 189-196), backup (199-203).
//                  Latency-bound.  They work on "hot" records in shared memory that hold what one level
//                  of the walk needs in a form that keeps its dependency chain short (see HotSel): the
//                  value term W/n is refreshed by the backup (off the walk's critical path), c_puct*prior
//                  is fixed at expansion, and RN(1/(1+n)) for the exact table-driven division rides along.
//   network warps  dynamics + prediction (network.py:217-236, 254-266) for all trees of the team as one
//                  register-tiled product: lane = (group of TQ trees, NV neurons), operands double
//                  buffered from shared memory, TQ*NV sequential-k FMA chains per lane.
//
// The two roles hand over through shared memory and two named barriers per simulation (bar.sync id,
// count: no CTA-wide barrier, teams never wait for each other).  The canonical tree arrays in HBM
// (cm::Node records, child_eidx, exp_node, hidden rows) are written once: priors and hidden rows when
// a node is expanded, statistics when the search of the tree ends.  Same arithmetic and operation
// order as search_fused.cu and oracle/cumind_oracle.c (sequential-k FMA chains, float64 statistics).
#include "search_core.cuh"
#include "tc_common.cuh"
#include "kernels.h"

namespace cm {

// bar.sync id, count: all `count` threads (whole warps) of the named barrier; the warp reconverges first
CM_DEVICE void named_bar_sync(int id, int count) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Philox Dirichlet draw kept out of line: it runs once per tree and would otherwise dominate the register
// allocation of the whole kernel (inlined log / pow).
__device__ __noinline__ void dirichlet_row_cold(double* out, int A, double alpha, unsigned long long seed, unsigned long long tree) {
    dirichlet_row(out, A, alpha, seed, tree);
}

// What one level of the walk reads for a child (shared memory, 32 bytes, two LDS.128):
//   q  = value_sum / visit_count as Node.value computes it (mcts.py:41-49; 0 while unvisited)
//   cp = c_puct * prior in float64 (mcts.py:65; root children: the prior after noise mixing)
//   r  = RN(1 / (1 + n)) for the exact division of the exploration term
//   n  = visit_count, ce = expansion index of the child's own children (-1: not expanded)
struct __align__(16) HotSel {
    double q, cp, r;
    int32_t n, ce;
};
static_assert(sizeof(HotSel) == 32, "HotSel must be 32 bytes");

CM_DEVICE void lds_hot(const HotSel* p, int4& a, int4& b) {
    a = *reinterpret_cast<const int4*>(p);
    b = *(reinterpret_cast<const int4*>(p) + 1);
}

// Selection on the hot records, one child per lane (A <= G), SPAN = pow2 >= A lanes in the butterfly.
// The records of the next level are requested before the warp votes on whether anyone still walks.
template <int G, int SPAN>
CM_DEVICE int select_hot(const HotSel* __restrict__ hs, const double* __restrict__ sqrt_tab, int A, int root_visits,
                         int* __restrict__ path, int g, bool active, int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, n_parent = root_visits, plen = 0, action = 0, parent_e = 0;
    int walking = active ? 1 : 0;
    const bool has_child = g < A;
    const HotSel* mine = hs + 1 + (has_child ? g : 0);
    int4 ra, rb;
    lds_hot(mine, ra, rb);
    while (true) {
        const double q = __hiloint2double(ra.y, ra.x), cp = __hiloint2double(ra.w, ra.z), r = __hiloint2double(rb.y, rb.x);
        const int n = rb.z, ce = rb.w;
        const double num = __dmul_rn(cp, sqrt_tab[n_parent]);
        double s = __dadd_rn(q, div_small(num, (double)(1 + n), r));
        s = (n == 0) ? dinf() : s;
        if (s != s) s = (g == 0) ? dinf() : -dinf();
        const bool valid = has_child && walking;
        double best_s = valid ? s : -dinf();
        unsigned best_ae = valid ? (((unsigned)g << 16) | (unsigned)(ce + 1)) : 0xffffffffu;
        int best_n = n;
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
        lds_hot(mine + (walking ? e * A : 0), ra, rb);
        if (!__any_sync(0xffffffffu, walking != 0)) break;
    }
    plen_out = plen;
    parent_e_out = parent_e;
    action_out = action;
    return node;
}

// ---- the fast walk -------------------------------------------------------------------------------
// Same selection with everything that is not arithmetic taken off the per-level dependency chain:
//   * shared memory is addressed through 32-bit window addresses kept in registers (ld.shared / st.shared);
//   * the division is the bare Markstein sequence: the caller guarantees (per tree, see cp_is_safe) that
//     the numerator c_puct*prior*sqrt(N) is 0, NaN or within [2^-900, 2^900], where the sequence is the
//     correctly rounded quotient (0 and NaN come out as 0 and NaN), so no range test and no branch;
//   * scores are compared as order-preserving 64-bit integer keys (-0 folded into +0, NaN and unvisited
//     mapped as in select_hot), one ISETP pair instead of three dependent DSETPs;
//   * sqrt(N_parent) of the next level is requested together with the next records.
CM_DEVICE void lds128_s(uint32_t addr, int4& v) {
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
}
CM_DEVICE double lds64_s(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
CM_DEVICE void sts32_pred(uint32_t addr, int v, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.s32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"(pred) : "memory");
}

CM_DEVICE void sts64_pred(uint32_t addr, double v, int pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" ::"r"(addr), "d"(v), "r"(pred) : "memory");
}
CM_DEVICE void sts128_s(uint32_t addr, int a, int b, int c, int d) {
    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
CM_DEVICE int lds32_s(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// c_puct*prior values for which the bare Markstein division of select_hot_fast is exact
CM_DEVICE bool cp_is_safe(double cp) {
    // 0, NaN, or a biased exponent in [1023-890, 1023+890): integer tests on the high word
    const unsigned hi = (unsigned)__double2hiint(cp) & 0x7fffffffu;
    const unsigned lo = (unsigned)__double2loint(cp);
    const unsigned ex = hi >> 20;
    const bool zero = (hi | lo) == 0u;
    const bool nan = hi > 0x7ff00000u || (hi == 0x7ff00000u && lo != 0u);
    return zero || nan || (ex - 133u) < 1780u;
}

template <int G, int SPAN>
CM_DEVICE int select_hot_fast(uint32_t hs_addr, uint32_t sqrt_addr, int A, int root_visits, uint32_t path_addr, int g, bool active,
                              int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, plen = 0, action = 0, parent_e = 0;
    int walking = active ? 1 : 0;
    const bool has_child = g < A;
    const uint32_t mine = hs_addr + 32u * (uint32_t)(1 + (has_child ? g : 0));
    const uint32_t stride = 32u * (uint32_t)A;
    const unsigned gbits = (unsigned)g << 16;
    const long long PINF = 0x7ff0000000000000LL, NINF = (long long)0xfff0000000000000ULL;
    const long long nan_key_src = (g == 0) ? PINF : NINF;
    int4 ra, rb;
    lds128_s(mine, ra);
    lds128_s(mine + 16u, rb);
    double sq = lds64_s(sqrt_addr + 8u * (uint32_t)root_visits);
    while (true) {
        const double q = __hiloint2double(ra.y, ra.x), cp = __hiloint2double(ra.w, ra.z), r = __hiloint2double(rb.y, rb.x);
        const int n = rb.z, ce = rb.w;
        const double num = __dmul_rn(cp, sq);
        const double d = (double)(1 + n);
        double t = __dmul_rn(num, r);
        t = __fma_rn(__fma_rn(-d, t, num), r, t);
        t = __fma_rn(__fma_rn(-d, t, num), r, t);
        long long b = __double_as_longlong(__dadd_rn(q, t));
        const long long mag = b & 0x7fffffffffffffffLL;
        b = mag > PINF ? nan_key_src : (mag == 0 ? 0LL : b);
        b = (n == 0) ? PINF : b;
        long long key = b ^ ((b >> 63) & 0x7fffffffffffffffLL);
        const bool valid = has_child && walking;
        key = valid ? key : (long long)0x8000000000000000ULL;
        unsigned ae = valid ? (gbits | (unsigned)(ce + 1)) : 0xffffffffu;
        int bn = n;
#pragma unroll
        for (int o = SPAN / 2; o > 0; o >>= 1) {
            const long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            const unsigned oae = __shfl_xor_sync(0xffffffffu, ae, o);
            const int on = __shfl_xor_sync(0xffffffffu, bn, o);
            const bool take = ok > key || (ok == key && oae < ae);
            key = take ? ok : key;
            ae = take ? oae : ae;
            bn = take ? on : bn;
        }
        if constexpr (SPAN < G) {
            ae = __shfl_sync(0xffffffffu, ae, 0, G);
            bn = __shfl_sync(0xffffffffu, bn, 0, G);
        }
        sts32_pred(path_addr + 4u * (uint32_t)plen, node, walking && g == 0);
        const int act = (int)(ae >> 16);
        const int ne = (int)(ae & 0xffffu) - 1;
        plen += walking;
        parent_e = walking ? e : parent_e;
        action = walking ? act : action;
        node = walking ? 1 + e * A + act : node;
        e = walking ? ne : e;
        walking = (walking && ne >= 0) ? 1 : 0;
        const uint32_t nxt = mine + (walking ? (uint32_t)e * stride : 0u);
        lds128_s(nxt, ra);
        lds128_s(nxt + 16u, rb);
        // lanes that do not walk (finished, or slots without a tree aliasing another warp's records, which may
        // be mid-update) must not turn what they loaded into an address
        sq = lds64_s(sqrt_addr + 8u * (uint32_t)(walking ? bn : 0));
        if (!__any_sync(0xffffffffu, walking != 0)) break;
    }
    plen_out = plen;
    parent_e_out = parent_e;
    action_out = action;
    return node;
}

template <int G>
CM_DEVICE int select_hot_fast_any(uint32_t hs_addr, uint32_t sqrt_addr, int A, int root_visits, uint32_t path_addr, int g, bool active,
                                  int& plen_out, int& parent_e_out, int& action_out) {
    if constexpr (G >= 4) {
        if (A <= G / 2)
            return select_hot_fast<G, G / 2>(hs_addr, sqrt_addr, A, root_visits, path_addr, g, active, plen_out, parent_e_out, action_out);
    }
    return select_hot_fast<G, G>(hs_addr, sqrt_addr, A, root_visits, path_addr, g, active, plen_out, parent_e_out, action_out);
}

// A > G (only with G == 32): every lane scores children g, g+32, ... of the current node.
template <int G>
CM_DEVICE int select_hot_loop(const HotSel* __restrict__ hs, const double* __restrict__ sqrt_tab, int A, int root_visits,
                              int* __restrict__ path, int g, bool active, int& plen_out, int& parent_e_out, int& action_out) {
    int node = 0, e = 0, n_parent = root_visits, plen = 0, action = 0, parent_e = 0;
    bool walking = active;
    while (__any_sync(0xffffffffu, walking)) {
        double best_s = -dinf();
        unsigned best_ae = 0xffffffffu;
        int best_n = 0;
        if (walking) {
            const double sq = sqrt_tab[n_parent];
            for (int a = g; a < A; a += G) {
                int4 ra, rb;
                lds_hot(hs + 1 + e * A + a, ra, rb);
                const double q = __hiloint2double(ra.y, ra.x), cp = __hiloint2double(ra.w, ra.z), r = __hiloint2double(rb.y, rb.x);
                const int n = rb.z, ce = rb.w;
                double s = __dadd_rn(q, div_small(__dmul_rn(cp, sq), (double)(1 + n), r));
                s = (n == 0) ? dinf() : s;
                if (s != s) s = (a == 0) ? dinf() : -dinf();
                if (best_ae == 0xffffffffu || s > best_s) { best_s = s; best_ae = ((unsigned)a << 16) | (unsigned)(ce + 1); best_n = n; }
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best_s, o);
            const unsigned oae = __shfl_xor_sync(0xffffffffu, best_ae, o);
            const int on = __shfl_xor_sync(0xffffffffu, best_n, o);
            if (os > best_s || (os == best_s && oae < best_ae)) { best_s = os; best_ae = oae; best_n = on; }
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
CM_DEVICE int select_hot_any(const HotSel* __restrict__ hs, const double* __restrict__ sqrt_tab, int A, int root_visits,
                             int* __restrict__ path, int g, bool active, int& plen_out, int& parent_e_out, int& action_out) {
    if constexpr (G >= 2) {
        if (A <= G / 2 && G / 2 >= 1)
            return select_hot<G, (G / 2 >= 1 ? G / 2 : 1)>(hs, sqrt_tab, A, root_visits, path, g, active, plen_out, parent_e_out, action_out);
    }
    if (A <= G) return select_hot<G, G>(hs, sqrt_tab, A, root_visits, path, g, active, plen_out, parent_e_out, action_out);
    return select_hot_loop<G>(hs, sqrt_tab, A, root_visits, path, g, active, plen_out, parent_e_out, action_out);
}

// Backup on the hot records (mcts.py:199-203): every parent on the path gets n += 1, W += v, the root
// (path[0]) a second time; the value term q = W/n and RN(1/(1+n)) are refreshed here so that the next
// walk does not have to.  U path entries per lane are processed together with no branch in between (entries
// past the end of the path stand in as the root and their stores are predicated off), so their dependency
// chains overlap.  q = W/n is the Markstein sequence of div_small; a value sum outside the range where
// that is exact (never in practice) is redone with the IEEE division after the block.  A zero sum stays on
// the sequence (it yields a zero; the sign of a zero value term never reaches a decision: scores are
// compared with -0 folded into +0, and root values are divided separately).
template <int G, int U>
CM_DEVICE void backup_hot(uint32_t hs_addr, uint32_t hw_addr, uint32_t path_addr, int plen, float leaf_value, uint32_t recip_addr,
                          int g) {
    const double v = (double)leaf_value;
#pragma unroll 1
    for (int i0 = g; i0 < plen; i0 += U * G) {
        int nd[U], n[U], ok[U];
        double w[U];
        int slow = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * G;
            ok[u] = i < plen ? 1 : 0;
            nd[u] = lds32_s(path_addr + 4u * (uint32_t)(ok[u] ? i : 0));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            w[u] = lds64_s(hw_addr + 8u * (uint32_t)nd[u]);
            n[u] = lds32_s(hs_addr + 32u * (uint32_t)nd[u] + 24u);
        }
        // the shared-memory accesses are volatile asm and keep their program order: all loads of all entries
        // first, then the arithmetic, then all stores, or one entry's stores would fence the next entry's loads
        double ww[U], rn[U], r1[U], q[U], d[U], t[U], e[U];
        int nn[U];
        // one operation of every entry at a time: a warp issues in order, so the U dependency chains only
        // overlap if the instruction stream itself interleaves them
#pragma unroll
        for (int u = 0; u < U; ++u) ww[u] = __dadd_rn(w[u], v);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool root = nd[u] == 0;
            const double w2 = __dadd_rn(ww[u], v);
            ww[u] = root ? w2 : ww[u];
            nn[u] = n[u] + (root ? 2 : 1);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            rn[u] = lds64_s(recip_addr + 8u * (uint32_t)nn[u]);
            r1[u] = lds64_s(recip_addr + 8u * (uint32_t)nn[u] + 8u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) d[u] = (double)nn[u];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = __dmul_rn(ww[u], rn[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = __fma_rn(-d[u], t[u], ww[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = __fma_rn(e[u], rn[u], t[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = __fma_rn(-d[u], t[u], ww[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) q[u] = __fma_rn(e[u], rn[u], t[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned hi = (unsigned)__double2hiint(ww[u]) & 0x7fffffffu;
            const bool exact = ((hi >> 20) - 123u) < 1800u || (hi | (unsigned)__double2loint(ww[u])) == 0u;
            slow |= (ok[u] && !exact) ? 1 : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t rec = hs_addr + 32u * (uint32_t)nd[u];
            sts64_pred(hw_addr + 8u * (uint32_t)nd[u], ww[u], ok[u]);
            sts64_pred(rec, q[u], ok[u]);
            sts64_pred(rec + 16u, r1[u], ok[u]);
            sts32_pred(rec + 24u, nn[u], ok[u]);
        }
        if (slow) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) sts64_pred(hs_addr + 32u * (uint32_t)nd[u], __ddiv_rn(ww[u], (double)nn[u]), 1);
            }
        }
    }
}

// softmax over lg[0..A_net) with the pinned exp, straight-line per lane; every lane returns the sum
template <int G>
CM_DEVICE float softmax_team(const float* __restrict__ lg, float* __restrict__ e_buf, int A_net, int g, bool active) {
    float m = lg[0];
    for (int a = 1; a < A_net; ++a) { const float l = lg[a]; m = l > m ? l : m; }
    for (int a = g; a < A_net; a += G) {
        const float e = __double2float_rn(pexp_nb((double)__fsub_rn(lg[a], m)));
        if (active) e_buf[a] = e;
    }
    __syncwarp();
    float s = 0.0f;
    for (int a = 0; a < A_net; ++a) s = fadd(s, e_buf[a]);
    return s;
}

// ---- network warps -------------------------------------------------------------------------------
template <int N>
CM_DEVICE void lds_vec(const float* p, float (&v)[N]) {
    if constexpr (N == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (N == 2) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; }
    else { v[0] = p[0]; }
}
template <int N>
CM_DEVICE void sts_vec(float* p, const float (&v)[N]) {
    if constexpr (N == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (N == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else p[0] = v[0];
}

// One dynamics layer for the team's trees: xout[j][t] = relu(dot_k(xin[k][t], W[k][j]) + b[j]) + xin[j][t].
// lane = (q, np): trees TQ*q.., neurons j0 = blk*NB + NV*np..; HS = compile-time row stride of W (0: runtime Hp);
// HC = compile-time hidden size (0: runtime, guarded loop).  With HC the k loop has no branches: the operands
// of the next 4 k are requested before the FMAs of the current 4 (register double buffering).
// hid != nullptr (last layer): also write the new hidden rows (tree ts at hid + ts*hstride) for ts < n_here,
// and the activations transposed, xt[t][j] (row stride H + 4), which is how the heads read them.
template <int TQ, int NV, int XS, int HS, int HC>
CM_DEVICE void team_layer(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ xin,
                          float* __restrict__ xout, int H_rt, int Hp_rt, int q, int np, int NB, int nblk, int mw, int nmw,
                          float* __restrict__ hid, size_t hstride, int n_here, float* __restrict__ xt) {
    const int Hp = HS > 0 ? HS : Hp_rt;
    const int H = HC > 0 ? HC : H_rt;
    for (int blk = mw; blk < nblk; blk += nmw) {
        const int j0 = blk * NB + NV * np;
        float acc[NV][TQ];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int t = 0; t < TQ; ++t) acc[v][t] = 0.0f;
        const float* x = xin + TQ * q;
        const float* w = W + j0;
        float xb[2][4][TQ], wb[2][4][NV];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
        if constexpr (HC > 0) {
            static_assert(HC % 8 == 0 && HC >= 16, "compile-time hidden size must be a multiple of 8");
#pragma unroll 1
            for (int k0 = 0; k0 < HC - 8; k0 += 8) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + (4 + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(4 + kk) * Hp, wb[1][kk]); }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[0][kk][t], wb[0][kk][v], acc[v][t]);
                x += 8 * XS;
                w += (size_t)8 * Hp;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[1][kk][t], wb[1][kk][v], acc[v][t]);
            }
            // last 8 k: nothing left to prefetch
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + (4 + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(4 + kk) * Hp, wb[1][kk]); }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[0][kk][t], wb[0][kk][v], acc[v][t]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[1][kk][t], wb[1][kk][v], acc[v][t]);
        } else {
#pragma unroll 1
            for (int k0 = 0; k0 < H; k0 += 8) {
                if (k0 + 4 < H) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + (4 + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(4 + kk) * Hp, wb[1][kk]); }
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[0][kk][t], wb[0][kk][v], acc[v][t]);
                x += 8 * XS;
                w += (size_t)8 * Hp;
                if (k0 + 8 < H) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
                }
                if (k0 + 4 < H) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int v = 0; v < NV; ++v)
#pragma unroll
                            for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(xb[1][kk][t], wb[1][kk][v], acc[v][t]);
                }
            }
        }
        if (j0 < H) {  // H % 4 == 0 and NV in {1,2,4} divides it: the lane's NV neurons are all valid or all padding
            float y[NV][TQ];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float bb = b[j0 + v];
                float r[TQ];
                lds_vec<TQ>(xin + (size_t)(j0 + v) * XS + TQ * q, r);
#pragma unroll
                for (int t = 0; t < TQ; ++t) y[v][t] = fadd(frelu(fadd(acc[v][t], bb)), r[t]);
                sts_vec<TQ>(xout + (size_t)(j0 + v) * XS + TQ * q, y[v]);
            }
            if (hid) {
#pragma unroll
                for (int t = 0; t < TQ; ++t) {
                    const int ts = TQ * q + t;
                    float o[NV];
#pragma unroll
                    for (int v = 0; v < NV; ++v) o[v] = y[v][t];
                    sts_vec<NV>(xt + (size_t)ts * (H + 4) + j0, o);
                    if (ts < n_here) {
                        float* dst = hid + (size_t)ts * hstride + j0;
                        if constexpr (NV == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                        else if constexpr (NV == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
                        else dst[0] = o[0];
                    }
                }
            }
        }
    }
}

// heads for the team's trees: lg[t][o] = dot_k(xt[t][k], WhT[o][k]) + bh[o], o < n_out; one sequential-k chain
// per lane, both operands k-contiguous so that one LDS.128 each feeds four steps of the chain.
template <int HC>
CM_DEVICE void team_heads(const float* __restrict__ WhT, const float* __restrict__ bh, const float* __restrict__ xt, int H_rt, int HP,
                          int TT, int n_out, float* __restrict__ lg, int mlane, int nml) {
    const int H = HC > 0 ? HC : H_rt;
    for (int c = mlane; c < n_out * TT; c += nml) {
        const int o = c / TT, t = c - o * TT;
        const float* xp = xt + (size_t)t * (H + 4);
        const float* wp = WhT + (size_t)o * H;
        float acc = 0.0f;
        if constexpr (HC > 0) {
#pragma unroll 8
            for (int k = 0; k < HC; k += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + k);
                const float4 wv = *reinterpret_cast<const float4*>(wp + k);
                acc = ffma(xv.x, wv.x, acc);
                acc = ffma(xv.y, wv.y, acc);
                acc = ffma(xv.z, wv.z, acc);
                acc = ffma(xv.w, wv.w, acc);
            }
        } else {
#pragma unroll 4
            for (int k = 0; k < H; k += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(xp + k);
                const float4 wv = *reinterpret_cast<const float4*>(wp + k);
                acc = ffma(xv.x, wv.x, acc);
                acc = ffma(xv.y, wv.y, acc);
                acc = ffma(xv.z, wv.z, acc);
                acc = ffma(xv.w, wv.w, acc);
            }
        }
        lg[t * HP + o] = fadd(acc, bh[o]);
    }
}

// G lanes per tree; TT tree slots per team; lane tile of the network warps = TQ trees x NV neurons;
// HS = compile-time hidden size (64: H == Hp == 64, loops without branches) or 0 (any H % 4 == 0).
template <int G, int TT, int TQ, int NV, int HS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) search_team_kernel(const TeamParams p) {
    constexpr int TPW = 32 / G;       // trees per tree warp
    constexpr int XS = TT + 4;        // row stride of the activation tiles (floats)
    constexpr int Q = TT / TQ;        // tree groups across a network warp
    constexpr int NPL = 32 / Q;       // lanes along the neuron axis
    constexpr int NB = NV * NPL;      // neurons per warp and block
    static_assert(TT % TQ == 0 && 32 % Q == 0, "bad lane tile");
    extern __shared__ __align__(128) unsigned char smem[];
    float* w_s = reinterpret_cast<float*>(smem);
    double* sq_s = reinterpret_cast<double*>(smem + p.off_sqrt);
    double* rc_s = reinterpret_cast<double*>(smem + p.off_recip);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int H = p.pk.H, L = p.pk.L, HP = p.pk.HP, A = p.tv.A, A_net = p.A_net;
    const int S_max = p.tv.S_max, N = p.tv.N;
    const int team_warps = p.ntw + p.nmw;
    const int team = warp / team_warps, wt = warp - team * team_warps;
    const bool is_tree = wt < p.ntw;

    // ---- stage parameters (TMA bulk) + sqrt / reciprocal tables ------------------------------------
    if (tid == 0) {
        tc_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = (uint32_t)(p.pk.total * sizeof(float));
        tc_mbar_expect_tx(bar, total);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < total; o += CH) {
            const uint32_t n = total - o < CH ? total - o : CH;
            tc_bulk_g2s(smem + o, reinterpret_cast<const unsigned char*>(p.pack) + o, n, bar);
        }
    }
    for (int i = tid; i < p.n_sqrt; i += blockDim.x) {
        sq_s[i] = __dsqrt_rn((double)i);
        rc_s[i] = __ddiv_rn(1.0, (double)i);  // RN(1/i); entry 0 is never used
    }
    tc_mbar_wait(bar, 0);
    __syncthreads();

    const float* emb_s = w_s + p.pk.emb;
    const float* wht_s = w_s + p.pk.heads_wt;
    const float* bh_s = w_s + p.pk.heads_b;

    // team scratch: x0,x1 [H][XS] | xt [TT][H+4] | lg,eb [TT][HP] | sel [4][TT] | path [tpt][S_max+2] | hot W [tpt][N] | hot records [tpt][N]
    unsigned char* my = smem + p.off_teams + (size_t)team * p.team_bytes;
    float* x0 = reinterpret_cast<float*>(my);
    float* x1 = x0 + (size_t)H * XS;
    float* xt = x1 + (size_t)H * XS;
    float* lg_t = xt + (size_t)TT * (H + 4);
    float* eb_t = lg_t + TT * HP;
    int* sel = reinterpret_cast<int*>(eb_t + TT * HP);   // [3][TT] parent_e, action, path length; then [TT] "checked walk" flags
    int* unsafe_t = sel + 3 * TT;
    int* path_t = reinterpret_cast<int*>(my + p.off_path);
    double* hw_t = reinterpret_cast<double*>(my + p.off_hotw);
    HotSel* hs_t = reinterpret_cast<HotSel*>(my + p.off_hot);
    const int bar_team = 1 + team, bar_mlp = 1 + p.n_team + team;
    const int cnt_team = team_warps * 32, cnt_mlp = p.nmw * 32;
    const size_t hstride = (size_t)(S_max + 1) * H;
    const bool clk_on = p.clk != nullptr && team == 0 && lane == 0 && (wt == 0 || wt == p.ntw);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, tmark = 0;
    if (clk_on) tmark = clock64();
    // Phase clocks (diagnostic).  The clock read sits behind a branch on a value that only exists once the
    // phase is over (a load that follows the barrier, a result of the loop): neither nvcc nor ptxas can
    // hoist it across the barrier or the loop it is meant to time.
#define CM_CLK(acc, dep)                                      \
    if (clk_on) {                                             \
        if ((int)(dep) != 0x7fedcba9) {                       \
            const long long now_ = clock64();                 \
            acc += now_ - tmark;                              \
            tmark = now_;                                     \
        }                                                     \
    }

    // Teams of one CTA start `stagger` cycles apart: they all run the same phases of the same length, and
    // started together they would walk their trees at the same time and then fight for the FMA pipe at the
    // same time; offset, one team's network phase overlaps the others' tree walks.
    if (p.stagger > 0 && team > 0) {
        const long long t_go = clock64() + (long long)team * p.stagger;
        while (clock64() < t_go) {}
    }
    const int team_global = blockIdx.x * p.n_team + team, team_stride = gridDim.x * p.n_team;
    for (int batch = team_global; batch < p.n_batches; batch += team_stride) {
        const int t_begin = batch * p.tpt;
        const int n_here = (p.tv.B - t_begin) < p.tpt ? (p.tv.B - t_begin) : p.tpt;

        if (is_tree) {
            // =============================== tree warps =============================================
            const int ts = wt * TPW + lane / G;
            const int g = lane & (G - 1);
            const bool active = ts < n_here;
            const int tsc = active ? ts : 0;  // inactive lanes alias slot 0 (reads only)
            const int t = t_begin + tsc;
            const size_t tt = (size_t)t;
            Node* nodes = p.tv.nodes + tt * N;
            int32_t* eidx = p.tv.child_eidx + tt * N;
            double* rprior = p.tv.root_prior + tt * A;
            int32_t* exp_node = p.tv.exp_node + tt * (S_max + 1);
            float* lg = lg_t + tsc * HP;
            float* eb = eb_t + tsc * HP;
            int* path = path_t + tsc * (S_max + 2);
            HotSel* hs = hs_t + (size_t)tsc * N;
            double* hw = hw_t + (size_t)tsc * N;
            const uint32_t hs_addr = smem_addr(hs), sq_addr = smem_addr(sq_s), path_addr = smem_addr(path);
            const uint32_t hw_addr = smem_addr(hw), rc_addr = smem_addr(rc_s);
            int* unsafe = unsafe_t + tsc;

            // ---- root: softmax of the root logits -> expand -> noise (mcts.py:126-136) -------------
            named_bar_sync(bar_team, cnt_team);  // the previous batch's logits have been consumed
            if (active && p.noise_mode == 2 && g == 0)
                dirichlet_row_cold(rprior, A, p.alpha, p.seed, p.tree_base + (unsigned long long)t);
            named_bar_sync(bar_team, cnt_team);  // root logits published
            if (active && g == 0) *unsafe = 0;
            __syncwarp();
            {
                const float ssum = softmax_team<G>(lg, eb, A_net, g, active);
                if (active) {
                    for (int a = g; a < A; a += G) {
                        const float p32 = prior_div(eb[a], ssum);
                        double pd = (double)p32;
                        if (p.noise_mode == 1)
                            pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(p.noise[tt * A + a], p.frac));
                        else if (p.noise_mode == 2)
                            pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(rprior[a], p.frac));
                        rprior[a] = pd;
                        store_node(nodes + 1 + a, 0.0, 0, p32);
                        HotSel h;
                        h.q = 0.0; h.cp = __dmul_rn(p.c_puct, pd); h.r = 1.0; h.n = 0; h.ce = -1;
                        hs[1 + a] = h;
                        hw[1 + a] = 0.0;
                        if (!cp_is_safe(h.cp)) *unsafe = 1;
                    }
                    if (g == 0) {
                        store_node(nodes, 0.0, 0, 1.0f);
                        HotSel h;
                        h.q = 0.0; h.cp = 0.0; h.r = 1.0; h.n = 0; h.ce = 0;
                        hs[0] = h;
                        hw[0] = 0.0;
                        exp_node[0] = 0;
                    }
                }
            }
            __syncwarp();
            CM_CLK(c3, hs[0].ce)

            // ---- S simulations (mcts.py:139-140, 157-203) -------------------------------------------
            long long depth = 0;
            for (int s = 0; s < p.S; ++s) {
                int plen, parent_e, action;
                // bare-division walk unless one of the warp's trees holds a c_puct*prior outside its exact range
                int leaf;
                if (A <= G && !__any_sync(0xffffffffu, active && *unsafe != 0))
                    leaf = select_hot_fast_any<G>(hs_addr, sq_addr, A, 2 * s, path_addr, g, active, plen, parent_e, action);
                else
                    leaf = select_hot_any<G>(hs, sq_s, A, 2 * s, path, g, active, plen, parent_e, action);
                if (g == 0 && ts < TT) {
                    sel[ts] = active ? parent_e : 0;
                    sel[TT + ts] = active ? action : 0;
                    sel[2 * TT + ts] = active ? plen : 0;
                }
                CM_CLK(c0, plen + leaf)
                named_bar_sync(bar_team, cnt_team);  // (A) selection published
                named_bar_sync(bar_team, cnt_team);  // (B) logits + value published
                CM_CLK(c1, __float_as_int(lg[A_net]))
                const int k = s + 1;  // expansion index of the leaf
                const float ssum = softmax_team<G>(lg, eb, A_net, g, active);
                CM_CLK(c5, __float_as_int(ssum))   // softmax alone; c6: expansion; c2: the backup that follows
                if (active) {
                    const float value = lg[A_net];
                    const int cbase = 1 + k * A;
                    for (int a = g; a < A; a += G) {
                        const float p32 = prior_div(eb[a], ssum);
                        store_node(nodes + cbase + a, 0.0, 0, p32);
                        const double cp = __dmul_rn(p.c_puct, (double)p32);
                        const uint32_t rec = hs_addr + 32u * (uint32_t)(cbase + a);
                        sts128_s(rec, 0, 0, __double2loint(cp), __double2hiint(cp));       // q = 0, cp
                        sts128_s(rec + 16u, 0, 0x3ff00000, 0, -1);                          // r = 1.0, n = 0, ce = -1
                        sts64_pred(hw_addr + 8u * (uint32_t)(cbase + a), 0.0, 1);
                        if (!cp_is_safe(cp)) *unsafe = 1;
                    }
                    if (g == 0) { hs[leaf].ce = k; exp_node[k] = leaf; }
                    depth += plen;
                }
                CM_CLK(c6, hs[1].n)
                named_bar_sync(bar_team, cnt_team);  // (C) the network warps have backed the value up
                __syncwarp();
                CM_CLK(c2, hs[0].n)
            }

            // ---- statistics -> canonical records; visit counts -> probabilities (mcts.py:143-155) ---
            if (active) {
                const int used = 1 + A * (p.S + 1);
#pragma unroll 4
                for (int i = g; i < used; i += G) {
                    nodes[i].value_sum = hw[i];
                    nodes[i].visit = hs[i].n;
                    eidx[i] = hs[i].ce;
                }
                long long total = 0;
                for (int a = 0; a < A; ++a) total += hs[1 + a].n;
                for (int a = g; a < p.A_cfg; a += G) {
                    const int c = a < A ? hs[1 + a].n : 0;
                    p.out_visits[tt * p.A_cfg + a] = c;
                    p.out_probs[tt * p.A_cfg + a] =
                        total == 0 ? __ddiv_rn(1.0, (double)p.A_cfg) : __ddiv_rn((double)c, (double)total);
                }
                if (g == 0) {
                    const int rn = hs[0].n;
                    if (p.out_root_value) p.out_root_value[t] = rn == 0 ? 0.0 : __ddiv_rn(hw[0], (double)rn);
                    p.tv.n_expanded[t] = p.S + 1;
                    p.tv.depth_sum[t] = depth;
                }
            }
            __syncwarp();
            CM_CLK(c3, hs[0].n)
        } else {
            // =============================== network warps ==========================================
            const int mw = wt - p.ntw;
            const int mlane = mw * 32 + lane, nml = p.nmw * 32;
            const int q = lane / NPL, np = lane - q * NPL;
            const int nblk = (H + NB - 1) / NB;
            const int H4 = H >> 2;
            float* hid0 = p.tv.hidden + (size_t)t_begin * hstride;

            // ---- root: x0 = root hidden (also hidden row 0), heads -> logits --------------------------
            for (int i = mlane; i < TT * H4; i += nml) {
                const int ts = i / H4, c = i - ts * H4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ts < n_here) {
                    v = *reinterpret_cast<const float4*>(p.root_hidden + (size_t)(t_begin + ts) * H + 4 * c);
                    *reinterpret_cast<float4*>(hid0 + (size_t)ts * hstride + 4 * c) = v;
                }
                *reinterpret_cast<float4*>(xt + (size_t)ts * (H + 4) + 4 * c) = v;
            }
            named_bar_sync(bar_team, cnt_team);  // xt filled; the tree warps are done with the previous batch's logits
            team_heads<HS>(wht_s, bh_s, xt, H, HP, TT, A_net, lg_t, mlane, nml);
            named_bar_sync(bar_team, cnt_team);  // root logits published

            for (int s = 0; s < p.S; ++s) {
                named_bar_sync(bar_team, cnt_team);  // (A) selection published
                CM_CLK(c3, sel[0])
                // gather: x0[j][ts] = hidden[ts][parent_e][j] + Emb[action][j]   (network.py:229-230).
                // Two items per lane and pass with both global loads in flight before either is used (the
                // parent rows come from L2); with a compile-time H the item -> (tree, chunk) split is shifts.
                {
                    const int H4c = HS > 0 ? HS / 4 : H4;
                    for (int i0 = mlane; i0 < TT * H4c; i0 += 2 * nml) {
                        int ts[2], c[2];
                        bool on[2];
                        float4 hv[2], ev[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = i0 + u * nml;
                            ts[u] = i / H4c;
                            c[u] = i - ts[u] * H4c;
                            on[u] = i < TT * H4c && ts[u] < n_here;
                            hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            ev[u] = hv[u];
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (on[u]) {
                                const int pe = sel[ts[u]], ac = sel[TT + ts[u]];
                                hv[u] = *reinterpret_cast<const float4*>(hid0 + (size_t)ts[u] * hstride + (size_t)pe * H + 4 * c[u]);
                                ev[u] = *reinterpret_cast<const float4*>(emb_s + (size_t)ac * H + 4 * c[u]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (i0 + u * nml < TT * H4c) {
                                float* xd = x0 + (size_t)(4 * c[u]) * XS + ts[u];
                                xd[0] = fadd(hv[u].x, ev[u].x);
                                xd[XS] = fadd(hv[u].y, ev[u].y);
                                xd[2 * XS] = fadd(hv[u].z, ev[u].z);
                                xd[3 * XS] = fadd(hv[u].w, ev[u].w);
                            }
                        }
                    }
                }
                named_bar_sync(bar_mlp, cnt_mlp);
                CM_CLK(c0, __float_as_int(x0[0]))
                float* cur = x0;
                float* nxt = x1;
                for (int l = 0; l < L; ++l) {
                    float* hid = (l == L - 1) ? (hid0 + (size_t)(s + 1) * H) : nullptr;
                    team_layer<TQ, NV, XS, HS, HS>(w_s + p.pk.layer_w(l), w_s + p.pk.layer_b(l), cur, nxt, H, p.pk.Hp, q, np, NB, nblk, mw,
                                                   p.nmw, hid, hstride, n_here, xt);
                    CM_CLK(c4, __float_as_int(nxt[0]))   // the layer itself (network warp 0); c1 then only counts the barriers
                    named_bar_sync(bar_mlp, cnt_mlp);
                    float* tmp = cur; cur = nxt; nxt = tmp;
                }
                CM_CLK(c1, __float_as_int(xt[0]))
                team_heads<HS>(wht_s, bh_s, xt, H, HP, TT, A_net + 1, lg_t, mlane, nml);
                __syncwarp();
                CM_CLK(c2, __float_as_int(lg_t[0]))
                named_bar_sync(bar_team, cnt_team);  // (B) logits + value published
                // Backup (mcts.py:199-203) on the network warps, one path entry per lane, while the tree warps
                // turn the logits into the new node's children: W += v, n += 1 on every parent of the path, the
                // root (path[0]) twice; q = W/n and RN(1/(1+n)) refreshed for the next walk.
                {
                    const int ts = mlane % TT;
                    if (ts < n_here) {
                        const int plen = sel[2 * TT + ts];
                        const double v = (double)lg_t[ts * HP + A_net];
                        const int* path = path_t + ts * (S_max + 2);
                        HotSel* hs = hs_t + (size_t)ts * N;
                        double* hw = hw_t + (size_t)ts * N;
                        for (int i = mlane / TT; i < plen; i += nml / TT) {
                            const int nd = path[i];
                            const bool root = nd == 0;
                            const double w1 = __dadd_rn(hw[nd], v);
                            const double ww = root ? __dadd_rn(w1, v) : w1;
                            const int nn = hs[nd].n + (root ? 2 : 1);
                            hw[nd] = ww;
                            hs[nd].q = div_small(ww, (double)nn, rc_s[nn]);
                            hs[nd].r = rc_s[1 + nn];
                            hs[nd].n = nn;
                        }
                    }
                }
                CM_CLK(c5, hs_t[0].n)
                named_bar_sync(bar_team, cnt_team);  // (C) statistics updated
            }
        }
    }
#undef CM_CLK
    if (clk_on) {
        long long* o = p.clk + (size_t)blockIdx.x * 16 + (is_tree ? 0 : 8);
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3; o[4] = c4; o[5] = c5; o[6] = c6;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch: the instantiations that exist (team_variant_ok) and the dispatch over them
// ------------------------------------------------------------------------------------------------
template <int G, int TT, int TQ, int NV, int HS, int MAXT>
static cudaError_t launch_one(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    auto k = search_team_kernel<G, TT, TQ, NV, HS, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, block, smem, st>>>(p);
    return cudaGetLastError();
}

// lane tiles of the network warps (TT tree slots, TQ trees x NV neurons per lane)
template <int G, int TT, int TQ, int NV>
static cudaError_t launch_t(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    const bool big = block > 512;
    if (p.pk.H == 64)
        return big ? launch_one<G, TT, TQ, NV, 64, 1024>(p, grid, block, smem, st) : launch_one<G, TT, TQ, NV, 64, 512>(p, grid, block, smem, st);
    return big ? launch_one<G, TT, TQ, NV, 0, 1024>(p, grid, block, smem, st) : launch_one<G, TT, TQ, NV, 0, 512>(p, grid, block, smem, st);
}

bool team_variant_ok(int G, int TT) {
    if (TT == 16) return G == 2 || G == 4 || G == 8;
    if (TT == 8) return G == 2 || G == 4 || G == 8 || G == 16 || G == 32;
    return false;
}

// p.tile: 0 = default (16: 4x2, 8: 2x2); tuning alternatives exist for G = 2, Hp = 64, <= 512 threads only
cudaError_t launch_search_team(int G, const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    if (!team_variant_ok(G, p.TT)) return cudaErrorInvalidValue;
    if (p.tile != 0) {
        if (p.TT == 16 && p.tile == 2 && G == 2 && p.pk.H == 64 && block <= 640) return launch_one<2, 16, 2, 2, 64, 640>(p, grid, block, smem, st);
        if (p.TT == 16 && p.tile == 3 && G == 2 && p.pk.H == 64 && block <= 640) return launch_one<2, 16, 4, 1, 64, 640>(p, grid, block, smem, st);
        if (G != 2 || p.pk.H != 64 || block > 512) return cudaErrorInvalidValue;
        if (p.TT == 16 && p.tile == 1) return launch_one<2, 16, 4, 4, 64, 512>(p, grid, block, smem, st);
        if (p.TT == 8 && p.tile == 1) return launch_one<2, 8, 4, 2, 64, 512>(p, grid, block, smem, st);
        if (p.TT == 8 && p.tile == 2) return launch_one<2, 8, 2, 4, 64, 512>(p, grid, block, smem, st);
        return cudaErrorInvalidValue;
    }
    if (p.TT == 16) {
        switch (G) {
            case 2: return launch_t<2, 16, 4, 2>(p, grid, block, smem, st);
            case 4: return launch_t<4, 16, 4, 2>(p, grid, block, smem, st);
            case 8: return launch_t<8, 16, 4, 2>(p, grid, block, smem, st);
        }
    } else {
        switch (G) {
            case 2: return launch_t<2, 8, 2, 2>(p, grid, block, smem, st);
            case 4: return launch_t<4, 8, 2, 2>(p, grid, block, smem, st);
            case 8: return launch_t<8, 8, 2, 2>(p, grid, block, smem, st);
            case 16: return launch_t<16, 8, 2, 2>(p, grid, block, smem, st);
            case 32: return launch_t<32, 8, 2, 2>(p, grid, block, smem, st);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace cm
