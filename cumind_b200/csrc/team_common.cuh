// team_common.cuh — device routines shared by the two team-schedule search kernels (search_team.cu, search_teamx.cu):
// named barriers, the shared-memory node records (HotSel), selection as a pointer chase, backup merged with the
// re-evaluation of select_child, softmax -> priors, and the prediction heads.  Reference anchors are given per routine
// (src/cumind/core/mcts.py, src/cumind/core/network.py).
#pragma once

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

// The statistics of a node while its tree is being searched (shared memory, 32 bytes, two LDS.128):
//   q  = value_sum / visit_count as Node.value computes it (mcts.py:41-49; 0 while unvisited)
//   cp = c_puct * prior in float64 (mcts.py:65; root children: the prior after noise mixing)
//   r  = RN(1 / (1 + n)) for the exact table-driven division of the exploration term
//   n  = visit_count, ce = expansion index of the node's own children (-1: not expanded)
// value_sum itself lives in a parallel array (hw) and the child select_child would pick in nxt (below).
struct __align__(16) HotSel {
    double q, cp, r;
    int32_t n, ce;
};
static_assert(sizeof(HotSel) == 32, "HotSel must be 32 bytes");

// shared memory through 32-bit window addresses kept in registers (ld.shared / st.shared)
CM_DEVICE void lds128_s(uint32_t addr, int4& v) {
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
}
CM_DEVICE double lds64_s(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
CM_DEVICE int lds32_s(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
CM_DEVICE void sts32_s(uint32_t addr, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
CM_DEVICE void sts64_s(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
CM_DEVICE void sts128_s(uint32_t addr, int a, int b, int c, int d) {
    asm volatile("st.shared.v4.s32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- selection as a pointer chase ------------------------------------------------------------------
// Node.select_child (mcts.py:67-76) is a function of the parent's visit count and its children's
// statistics, and all of those only change in a simulation whose path goes through the parent: the backup
// (mcts.py:199-203) touches the parents on the path and nothing else, and a child is on the path only if
// its parent is.  So the child the NEXT visit of a node will select is known as soon as the backup of
// the current simulation is done.  backup_rescore (network warps) stores it in nxt[node] for every parent
// on the path; a freshly expanded node gets nxt = its child 0 (all children unvisited: every score is +inf
// and max() keeps the first, mcts.py:76); an unexpanded node has nxt = -1.  The walk of mcts.py:165-171 is
// then one dependent shared-memory load per level.  Returns the leaf; path[0..plen) = the parents.
CM_DEVICE int select_chase(uint32_t nxt_addr, uint32_t path_addr, int g, bool active, int& plen_out) {
    int node = 0, plen = 0;
    if (active) {
        int nx = lds32_s(nxt_addr);
        while (nx >= 0) {
            if (g == 0) sts32_s(path_addr + 4u * (uint32_t)plen, node);
            ++plen;
            node = nx;
            nx = lds32_s(nxt_addr + 4u * (uint32_t)nx);
        }
    }
    __syncwarp();
    plen_out = plen;
    return node;
}

// Node.ucb_score (mcts.py:51-65) of one child as an order-preserving 64-bit integer key:
//   +inf if n == 0, else q + ((c_puct*prior)*sqrt(N_parent)) / (1 + n) with q = W/n already divided.
// IEEE: the division is __ddiv_rn; otherwise it is the Markstein sequence of div_small, which is the correctly
// rounded quotient when the numerator is 0, NaN or within [2^-900, 2^900] (`exact` reports whether it was).
// -0 is folded into +0 (Python compares them equal), NaN maps to +inf for child 0 and -inf otherwise (max()
// never replaces its first candidate by a NaN and never replaces a NaN first candidate).
// The exploration term ((c_puct*prior)*sqrt(N_parent)) / (1 + n) of Node.ucb_score (mcts.py:65); it does not depend on
// the value being backed up.
template <bool IEEE>
CM_DEVICE double explo_term(double cp, double r, int n, double sq, bool& exact) {
    const double num = __dmul_rn(cp, sq);
    const double d = (double)(1 + n);
    double t;
    if constexpr (IEEE) {
        t = __ddiv_rn(num, d);
        exact = true;
    } else {
        t = __dmul_rn(num, r);
        t = __fma_rn(__fma_rn(-d, t, num), r, t);
        t = __fma_rn(__fma_rn(-d, t, num), r, t);
        const unsigned hi = (unsigned)__double2hiint(num) & 0x7fffffffu;
        const unsigned lo = (unsigned)__double2loint(num);
        exact = ((hi >> 20) - 123u) < 1800u || (hi | lo) == 0u || hi > 0x7ff00000u || (hi == 0x7ff00000u && lo != 0u) || n == 0;
    }
    return t;
}
// q + t as the order-preserving key (see score_key)
CM_DEVICE long long key_of(double q, double t, int n, bool first) {
    const long long PINF = 0x7ff0000000000000LL, NINF = (long long)0xfff0000000000000ULL;
    long long b = __double_as_longlong(__dadd_rn(q, t));
    const long long mag = b & 0x7fffffffffffffffLL;
    b = mag > PINF ? (first ? PINF : NINF) : (mag == 0 ? 0LL : b);
    b = (n == 0) ? PINF : b;
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
template <bool IEEE>
CM_DEVICE long long score_key(double q, double cp, double r, int n, double sq, bool first, bool& exact) {
    return key_of(q, explo_term<IEEE>(cp, r, n, sq, exact), n, first);
}

// Markstein sequence of div_small: a / d from r = RN(1/d); the correctly rounded quotient when markstein_ok(a)
CM_DEVICE double markstein(double a, double d, double r) {
    double q = __dmul_rn(a, r);
    q = __fma_rn(__fma_rn(-d, q, a), r, q);
    return __fma_rn(__fma_rn(-d, q, a), r, q);
}
// biased exponent in [123, 1923): the residuals of the sequence are exact
CM_DEVICE bool markstein_ok(double a) { return ((((unsigned)__double2hiint(a) & 0x7fffffffu) >> 20) - 123u) < 1800u; }

// Lane mapping of backup_rescore, fixed for the whole launch: the path entries of a tree are spread over
// `lpt` lanes of ONE network warp.
struct BackupLanes {
    int ts;    // tree slot of this lane (-1: none)
    int il;    // first path entry of this lane
    int lpt;   // stride between the entries of a lane
};
CM_DEVICE BackupLanes backup_lanes(int TT, int nmw, int mw, int lane) {
    const int tpw = (TT + nmw - 1) / nmw;  // trees per network warp
    BackupLanes b;
    b.lpt = 32 / tpw;
    const int tl = lane / b.lpt;
    b.il = lane - tl * b.lpt;
    b.ts = tl < tpw ? mw * tpw + tl : -1;
    return b;
}

// Backup (mcts.py:199-203) and the selection cache, on the network warps.  The lane of entry i owns parent
// P = path[i]: it adds the leaf value to P (W += v, n += 1; the root, path[0], twice) and refreshes q = W/n and
// RN(1/(1+n)).  To re-evaluate which child P will select next it needs the new statistics of the one child that
// changed, C = path[i+1] (the leaf is not backed up): it recomputes them privately from C's old record, which is
// why every lane reads before any lane writes (one warp: shared-memory accesses complete in program order).
// The other children are not on the path and are read as they are.  Straight-line arithmetic; operands outside
// the exact range of the Markstein sequence (never in practice) are redone with IEEE divisions at the end.
CM_DEVICE void backup_rescore(uint32_t hs_t, uint32_t hw_t, uint32_t nxt_t, uint32_t path_t, const int* __restrict__ sel,
                              const float* __restrict__ lg, uint32_t sq_addr, uint32_t rc_addr, int N, int A, int S_max, int TT,
                              int HP, int A_net, int n_here, const BackupLanes bl) {
    const bool valid = bl.ts >= 0 && bl.ts < n_here;
    const int tsc = valid ? bl.ts : 0;
    const int plen = valid ? sel[2 * TT + tsc] : 0;
    const double v = (double)lg[tsc * HP + A_net];
    const uint32_t path = path_t + 4u * (uint32_t)(tsc * (S_max + 2));
    const uint32_t hs = hs_t + 32u * (uint32_t)(tsc * N);
    const uint32_t hw = hw_t + 8u * (uint32_t)(tsc * N);
    const uint32_t nxt = nxt_t + 4u * (uint32_t)(tsc * N);
    const int maxlen = __reduce_max_sync(0xffffffffu, plen);
    for (int i = bl.il; i - bl.il < maxlen; i += bl.lpt) {
        const bool on = i < plen;
        const bool has_c = i + 1 < plen;
        const int P = lds32_s(path + 4u * (uint32_t)(on ? i : 0));
        const int Cl = lds32_s(path + 4u * (uint32_t)(has_c ? i + 1 : 0));
        // old statistics of P and of the child on the path
        const uint32_t recP = hs + 32u * (uint32_t)P, recC = hs + 32u * (uint32_t)Cl;
        const double wP = lds64_s(hw + 8u * (uint32_t)P);
        const int nP = lds32_s(recP + 24u);
        const int ceP = lds32_s(recP + 28u);
        const double wC = lds64_s(hw + 8u * (uint32_t)Cl);
        const int nC = lds32_s(recC + 24u);
        const bool root = P == 0;
        const int nPn = nP + (root ? 2 : 1);
        const int nCn = nC + 1;
        const double sq = lds64_s(sq_addr + 8u * (uint32_t)nPn);
        const double rP = lds64_s(rc_addr + 8u * (uint32_t)nPn), rP1 = lds64_s(rc_addr + 8u * (uint32_t)nPn + 8u);
        const double rC = lds64_s(rc_addr + 8u * (uint32_t)nCn), rC1 = lds64_s(rc_addr + 8u * (uint32_t)nCn + 8u);
        // the other children's records (unchanged by this backup); C's own record is loaded too and ignored
        const int base = 1 + ceP * A;
        const int C = has_c ? Cl : -1;
        const double wP1 = __dadd_rn(wP, v);
        const double wPn = root ? __dadd_rn(wP1, v) : wP1;
        const double wCn = __dadd_rn(wC, v);
        double qPn = markstein(wPn, (double)nPn, rP);
        double qCn = markstein(wCn, (double)nCn, rC);
        if (!(markstein_ok(wPn) && markstein_ok(wCn))) {
            if (!markstein_ok(wPn)) qPn = (wPn == 0.0) ? wPn : __ddiv_rn(wPn, (double)nPn);
            if (!markstein_ok(wCn)) qCn = (wCn == 0.0) ? wCn : __ddiv_rn(wCn, (double)nCn);
        }
        __syncwarp();  // every lane has read the old records
        if (on) {
            sts64_s(hw + 8u * (uint32_t)P, wPn);
            sts64_s(recP, qPn);
            sts64_s(recP + 16u, rP1);
            sts32_s(recP + 24u, nPn);
            // which child does P select next time?
            long long best = 0;
            int best_c = base;
            bool all_exact = true;
#pragma unroll 2
            for (int a = 0; a < A; ++a) {
                const int c = base + a;
                int4 ra, rb;
                lds128_s(hs + 32u * (uint32_t)c, ra);
                lds128_s(hs + 32u * (uint32_t)c + 16u, rb);
                const bool is_c = c == C;
                const double q = is_c ? qCn : __hiloint2double(ra.y, ra.x);
                const double cp = __hiloint2double(ra.w, ra.z);
                const double r = is_c ? rC1 : __hiloint2double(rb.y, rb.x);
                const int n = is_c ? nCn : rb.z;
                bool exact;
                const long long key = score_key<false>(q, cp, r, n, sq, a == 0, exact);
                all_exact = all_exact && exact;
                if (a == 0 || key > best) { best = key; best_c = c; }
            }
            if (!all_exact) {  // same scan with IEEE divisions
                best_c = base;
                for (int a = 0; a < A; ++a) {
                    const int c = base + a;
                    int4 ra, rb;
                    lds128_s(hs + 32u * (uint32_t)c, ra);
                    lds128_s(hs + 32u * (uint32_t)c + 16u, rb);
                    const bool is_c = c == C;
                    const double q = is_c ? qCn : __hiloint2double(ra.y, ra.x);
                    const double cp = __hiloint2double(ra.w, ra.z);
                    const int n = is_c ? nCn : rb.z;
                    bool exact;
                    const long long key = score_key<true>(q, cp, 0.0, n, sq, a == 0, exact);
                    if (a == 0 || key > best) { best = key; best_c = c; }
                }
            }
            sts32_s(nxt + 4u * (uint32_t)P, best_c);
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

}  // namespace cm
