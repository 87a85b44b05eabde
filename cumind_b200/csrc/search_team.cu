// search_team.cu — MCTS.search (src/cumind/core/mcts.py:115-155) for B trees in ONE launch,
// warp-specialised: the tree walk and the network run on different warps of the same CTA, and the
// tree statistics live in shared memory while a tree is being searched.
//
// The per-tree work of one simulation is a serial chain (walk d levels -> dynamics MLP -> heads ->
// backup with float64 PUCT re-scoring; softmax -> priors trails one simulation behind).  search_fused.cu gives every warp the whole chain of its own
// few trees, so each phase runs at a fraction of a warp's lanes and an SM holds only a handful of
// warps.  Here a CTA is split into TEAMS; a team owns up to TT trees (a batch) and has two kinds of warps:
//
//   tree warps     G = pow2 >= A lanes per tree (32/G trees per warp): the walk of mcts.py:165-171 as a pointer chase
//                  through the selection cache (select_chase), Node.expand (78-89, 195-196) in two parts — the
//                  children are created right after the walk, their priors (softmax, 189-196) one simulation
//                  later, off the critical path — and the visit-count normalisation at the end (143-155).
//   network warps  dynamics + prediction (network.py:217-236, 254-266) for all trees of the team as one
//                  register-tiled product: lane = (group of TQ trees, NV neurons), operands double
//                  buffered from shared memory, TQ*NV sequential-k FMA chains per lane; then the backup
//                  (mcts.py:199-203) merged with the re-evaluation of select_child / ucb_score (51-76) for
//                  every parent on the path (backup_rescore), which is what fills the selection cache.
//                  The "hot" records in shared memory (HotSel) hold what that needs in a form that keeps its
//                  dependency chain short: W/n already divided, c_puct*prior fixed at expansion, RN(1/(1+n))
//                  for the exact table-driven division.
//
// The two roles hand over through shared memory and two named barriers per simulation (bar.sync id,
// count: no CTA-wide barrier, teams never wait for each other).  The canonical tree arrays in HBM
// (cm::Node records, child_eidx, exp_node, hidden rows) are written once: priors and hidden rows when
// a node is expanded, statistics when the search of the tree ends.  Same arithmetic and operation
// order as search_fused.cu and oracle/cumind_oracle.c (sequential-k FMA chains, float64 statistics).
#include "team_common.cuh"

#ifndef CM_TEAM_KB
#define CM_TEAM_KB 4  // k per half buffer of the layer loop's operand double buffering
#endif

namespace cm {

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

// acc[v][t] = fma(x[t], w[v], acc[v][t]) for the lane tile; neuron pairs go through FFMA2 (x as the scalar operand,
// {w[v], w[v+1]} and the accumulators as register pairs): a 3-register FFMA issues every other cycle on sm_100, an
// FFMA2 carries two of them in the same slot.  Every product-sum is still one IEEE fma of the same operands.
template <int TQ, int NV>
CM_DEVICE void fma_tile(const float (&x)[TQ], const float (&w)[NV], float (&acc)[NV][TQ]) {
    if constexpr (NV % 2 == 0) {
#pragma unroll
        for (int v = 0; v < NV; v += 2)
#pragma unroll
            for (int t = 0; t < TQ; ++t) {
                const float2 r = ffma2(x[t], make_float2(w[v], w[v + 1]), make_float2(acc[v][t], acc[v + 1][t]));
                acc[v][t] = r.x;
                acc[v + 1][t] = r.y;
            }
    } else {
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int t = 0; t < TQ; ++t) acc[v][t] = ffma(x[t], w[v], acc[v][t]);
    }
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
        constexpr int KB = HC > 0 ? CM_TEAM_KB : 4;  // k per half buffer (runtime H: 4, any H % 4 == 0): the operands of the next KB k are in flight while KB k of FMAs issue
        float xb[2][KB][TQ], wb[2][KB][NV];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
        if constexpr (HC > 0) {
            static_assert(HC % (2 * KB) == 0 && HC >= 4 * KB, "compile-time hidden size must be a multiple of 8");
#pragma unroll 1
            for (int k0 = 0; k0 < HC - 2 * KB; k0 += 2 * KB) {
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + (KB + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(KB + kk) * Hp, wb[1][kk]); }
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[0][kk], wb[0][kk], acc);
                x += 2 * KB * XS;
                w += (size_t)2 * KB * Hp;
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[1][kk], wb[1][kk], acc);
            }
            // last 2*KB k: nothing left to prefetch
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + (KB + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(KB + kk) * Hp, wb[1][kk]); }
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[0][kk], wb[0][kk], acc);
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[1][kk], wb[1][kk], acc);
        } else {
#pragma unroll 1
            for (int k0 = 0; k0 < H; k0 += 2 * KB) {
                if (k0 + KB < H) {
#pragma unroll
                    for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + (KB + kk) * XS, xb[1][kk]); lds_vec<NV>(w + (size_t)(KB + kk) * Hp, wb[1][kk]); }
                }
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[0][kk], wb[0][kk], acc);
                x += 2 * KB * XS;
                w += (size_t)2 * KB * Hp;
                if (k0 + 2 * KB < H) {
#pragma unroll
                    for (int kk = 0; kk < KB; ++kk) { lds_vec<TQ>(x + kk * XS, xb[0][kk]); lds_vec<NV>(w + (size_t)kk * Hp, wb[0][kk]); }
                }
                if (k0 + KB < H) {
#pragma unroll
                    for (int kk = 0; kk < KB; ++kk) fma_tile<TQ, NV>(xb[1][kk], wb[1][kk], acc);
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

// ---- layer warpgroup (KR > 0, hidden_dim 64) ------------------------------------------------------
// One dynamics layer for NT trees of ONE warp: lane = the neuron pair (2*lane, 2*lane+1), so a warp covers all 64
// neurons of its trees and the layers of a tree never leave the warp.  The activations are row-major per tree
// (row stride RS floats): x[t][k..k+3] is one broadcast LDS.128 for the whole warp and feeds four FFMA2
// (x_k as the scalar operand, the weight pair {W[k][2l], W[k][2l+1]}, the accumulator pair).  The weight pairs of
// k < KR stay in registers for the whole launch (wr), the others come from the k-major copy in shared memory
// (one conflict-free LDS.64 per k).  Same sequential-k FMA chain per (tree, neuron) as team_layer.
// hid != nullptr (last layer): also store the new hidden rows of the first n_valid trees (global, 256 B per row).
template <int NT, int KR, int RS>
CM_DEVICE void wg_layer(const float2 (&wr)[KR > 0 ? KR : 1], const float* __restrict__ Wl, float2 bias, const float* __restrict__ xin,
                        float* __restrict__ xout, float* __restrict__ hid, size_t hstride, int n_valid, int lane) {
    // trees in groups of GT: GT independent FMA chains issue round-robin (an FFMA2 result is needed again GT instructions
    // later), and only GT activation rows are in flight at a time, which is what keeps the live registers within budget
    constexpr int GT = NT < 4 ? NT : 4;
    static_assert(NT % GT == 0, "tree slots per layer warp must be a multiple of the group size");
#pragma unroll
    for (int g0 = 0; g0 < NT; g0 += GT) {
        const float* xg = xin + g0 * RS;
        float2 acc[GT];
#pragma unroll
        for (int t = 0; t < GT; ++t) acc[t] = make_float2(0.0f, 0.0f);
        // the rows are read PD blocks of 4 k ahead of their use: under load a shared-memory access takes several times
        // the issue time of one block's FFMA2s
        constexpr int PD = 2, NB = PD + 1;
        float4 a[NB][GT];
#pragma unroll
        for (int b = 0; b < PD; ++b)
#pragma unroll
            for (int t = 0; t < GT; ++t) a[b][t] = *reinterpret_cast<const float4*>(xg + t * RS + 4 * b);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            if (b + PD < 16) {
#pragma unroll
                for (int t = 0; t < GT; ++t) a[(b + PD) % NB][t] = *reinterpret_cast<const float4*>(xg + t * RS + 4 * (b + PD));
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = 4 * b + kk;
                float2 w;
                if (k < KR) w = wr[k < KR ? k : 0];
                else w = *reinterpret_cast<const float2*>(Wl + k * 64);
#pragma unroll
                for (int t = 0; t < GT; ++t) {
                    const float4 av = a[b % NB][t];
                    const float xv = kk == 0 ? av.x : (kk == 1 ? av.y : (kk == 2 ? av.z : av.w));
                    acc[t] = ffma2(xv, w, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < GT; ++t) {
            const float2 r = *reinterpret_cast<const float2*>(xg + t * RS + 2 * lane);
            float2 y;
            y.x = fadd(frelu(fadd(acc[t].x, bias.x)), r.x);
            y.y = fadd(frelu(fadd(acc[t].y, bias.y)), r.y);
            *reinterpret_cast<float2*>(xout + (g0 + t) * RS + 2 * lane) = y;
            if (hid != nullptr && g0 + t < n_valid) *reinterpret_cast<float2*>(hid + (size_t)(g0 + t) * hstride) = y;
        }
    }
}

// G lanes per tree; TT tree slots per team; lane tile of the network warps = TQ trees x NV neurons;
// HS = compile-time hidden size (64: H == Hp == 64, loops without branches) or 0 (any H % 4 == 0).
// KR > 0: LAYER WARPGROUP mode (hidden_dim 64, two dynamics layers).  Warps 0-3 of the CTA do nothing but the dynamics
// layers, for every team in turn: warps 0,1 hold layer 0 and warps 2,3 layer 1 as 64 weight pairs per lane in registers
// (wg_layer; the register file is re-divided between the roles with setmaxnreg).  The network warps of a team gather
// the parent rows row-major, hand them over (named barriers X0R -> X1R -> X2R), and compute heads + backup as before.
// The layers then read nothing but their activation rows from shared memory (one broadcast LDS.128 per 4 k and tree),
// a third of the wavefronts of the register-tiled product of team_layer, and the two stages pipeline across teams.
// KR = 0: the network warps compute the layers themselves (team_layer).
template <int G, int TT, int TQ, int NV, int HS, int KR, int AUXREG = 104>
CM_DEVICE void search_team_body(const TeamParams& p) {
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
    constexpr int LWW = KR > 0 ? 4 : 0;   // warps of the layer warpgroup
    constexpr int RS = 68;                // row stride of the row-major activation rows (KR > 0)
    const int team_warps = p.ntw + p.nmw;
    const bool is_layerwg = KR > 0 && warp < LWW;
    const int team = is_layerwg ? 0 : (warp - LWW) / team_warps, wt = is_layerwg ? 0 : (warp - LWW) - team * team_warps;
    const bool is_tree = !is_layerwg && wt < p.ntw;

    // ---- stage parameters (TMA bulk) + sqrt / reciprocal tables ------------------------------------
    // layer-warpgroup hand-overs (KR > 0), four mbarriers per team after the staging barrier:
    //   X0R "gathered rows in x0" (one arrival per network warp)  ->  X1R[half] "rows after layer 0 in x1" (one arrival)
    //   ->  X2R "rows after layer 1 in xt" (two arrivals); each completes once per simulation, so its parity is the job count & 1
    uint64_t* hand = bar + 1;
    if (tid == 0) {
        tc_mbar_init(bar, 1);
        if constexpr (KR > 0) {
            for (int tm = 0; tm < p.n_team; ++tm) {
                tc_mbar_init(hand + 4 * tm + 0, (uint32_t)p.nmw);
                tc_mbar_init(hand + 4 * tm + 1, 1);
                tc_mbar_init(hand + 4 * tm + 2, 1);
                tc_mbar_init(hand + 4 * tm + 3, 2);
            }
        }
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
    pdl_wait();  // root_hidden (and the noise) may come from the previous kernel of the stream

    const float* emb_s = w_s + p.pk.emb;
    const float* wht_s = w_s + p.pk.heads_wt;
    const float* bh_s = w_s + p.pk.heads_b;

    // team scratch: x0,x1 [H][XS] | xt [TT][H+4] | lg[2],eb [TT][HP] | sel [4][TT] | path [tpt][S_max+2] | nxt [tpt][N] | hot W [tpt][N] | hot records [tpt][N]
    unsigned char* my = smem + p.off_teams + (size_t)team * p.team_bytes;
    // KR > 0: x0 [TT][RS] (gathered rows, row-major) | x1 [TT][RS] (after layer 0) | xt [TT][RS] (after layer 1) | lg ...
    float* x0 = reinterpret_cast<float*>(my);
    float* x1 = KR > 0 ? x0 + (size_t)TT * RS : x0 + (size_t)H * XS;
    float* xt = KR > 0 ? x1 + (size_t)TT * RS : x1 + (size_t)H * XS;
    float* lg_t = xt + (size_t)TT * (H + 4);              // logits + value of simulation s in buffer s & 1 (the root's in buffer 1)
    float* eb_t = lg_t + 2 * TT * HP;
    int* sel = reinterpret_cast<int*>(eb_t + TT * HP);   // [3][TT] parent_e, action, path length
    int* path_t = reinterpret_cast<int*>(my + p.off_path);
    int* nxt_t = reinterpret_cast<int*>(my + p.off_nxt);
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
    if constexpr (KR > 0) {
        // Register reallocation between the roles (setmaxnreg, whole warpgroups; the pool is what the launch allocated):
        // 16 warps are launched with 128 registers per thread (four warps per SM sub-partition, 16 384 registers each);
        // the layer warps grow to 200 — a full layer of weight pairs is 128 of them — and the twelve other warps shrink
        // to AUXREG = 104, which is what fits beside one layer warp per sub-partition (32 * (200 + 3 * 104) = 16 384).
        // 12 warps (one team) are launched with 168 and the eight other warps shrink to 152 (32 * (200 + 2 * 152)).
        if (is_layerwg) {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 200;" ::: "memory");
            // warps 0,1: layer 0 of tree slots [0, NT) / [NT, 2 NT); warps 2,3: layer 1 of the same halves.  Each warp keeps
            // ITS layer's 64 weight pairs in registers for the whole launch; a team's rows go x0 -> (warps 0,1) -> x1 ->
            // (warps 2,3) -> xt, and while warps 2,3 finish one team warps 0,1 already run the other.
            constexpr int NT = TT / 2;
            const int stage = warp >> 1, half = warp & 1;
            const int ts0 = half * NT;
            float2 wr[64];
            const float* Wl = w_s + p.pk.layer_w(stage) + 2 * lane;
#pragma unroll
            for (int k = 0; k < 64; ++k) wr[k] = *reinterpret_cast<const float2*>(Wl + k * 64);
            const float2 bias = *reinterpret_cast<const float2*>(w_s + p.pk.layer_b(stage) + 2 * lane);
            const int team_stride = gridDim.x * p.n_team;
            const bool wclk = p.clk != nullptr && half == 0 && lane == 0;
            long long lc0 = 0, lc1 = 0;
            uint32_t job = 0;   // jobs done per team so far: the parity of every hand-over barrier
            for (int first = blockIdx.x * p.n_team; first < p.n_batches; first += team_stride) {
                for (int s = 0; s < p.S; ++s, ++job) {
                    for (int tm = 0; tm < p.n_team; ++tm) {
                        const int batch = first + tm;
                        if (batch >= p.n_batches) break;
                        const int t_begin = batch * p.tpt;
                        const int n_here = (p.tv.B - t_begin) < p.tpt ? (p.tv.B - t_begin) : p.tpt;
                        float* tsc = reinterpret_cast<float*>(smem + p.off_teams + (size_t)tm * p.team_bytes);
                        const float* x0t = tsc + ts0 * RS;
                        float* x1t = tsc + (size_t)TT * RS + ts0 * RS;
                        float* x2t = tsc + (size_t)2 * TT * RS + ts0 * RS;
                        uint64_t* hb = hand + 4 * tm;
                        const long long tw0 = wclk ? clock64() : 0;
                        if (stage == 0) {
                            tc_mbar_wait_spin(hb + 0, job & 1u);                    // X0R: gathered rows of team tm
                            const long long tw1 = wclk ? clock64() : 0;
                            wg_layer<NT, 64, RS>(wr, Wl, bias, x0t, x1t, nullptr, 0, 0, lane);
                            __syncwarp();
                            if (lane == 0) tc_mbar_arrive(hb + 1 + half);      // X1R[half]
                            if (wclk) { const long long tw2 = clock64(); lc0 += tw1 - tw0; lc1 += tw2 - tw1; }
                        } else {
                            tc_mbar_wait_spin(hb + 1 + half, job & 1u);
                            const long long tw1 = wclk ? clock64() : 0;
                            float* hid = p.tv.hidden + ((size_t)(t_begin + ts0) * (S_max + 1) + (size_t)(s + 1)) * H + 2 * lane;
                            wg_layer<NT, 64, RS>(wr, Wl, bias, x1t, x2t, hid, (size_t)(S_max + 1) * H, n_here - ts0, lane);
                            __syncwarp();
                            if (lane == 0) tc_mbar_arrive(hb + 3);             // X2R
                            if (wclk) { const long long tw2 = clock64(); lc0 += tw1 - tw0; lc1 += tw2 - tw1; }
                        }
                    }
                }
            }
            // clock slots: layer-0 warp 0 -> {7: waiting for rows, 15: computing}; layer-1 warp 2 -> added to the same slots is
            // avoided: it reports through the tree / network wait clocks instead
            if (wclk && stage == 0) { p.clk[(size_t)blockIdx.x * 16 + 7] = lc0; p.clk[(size_t)blockIdx.x * 16 + 15] = lc1; }
            return;
        }
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUXREG) : "memory");
    }
    // (leaf - 1) / A by multiplication: exact for leaf - 1 < 2^32 / A
    const unsigned magic_A = A > 1 ? (unsigned)((0x100000000ULL + (unsigned)A - 1u) / (unsigned)A) : 0u;
    const int team_global = blockIdx.x * p.n_team + team, team_stride = gridDim.x * p.n_team;
    uint32_t job = 0;   // layer-warpgroup mode: simulations this team has handed over so far (hand-over barrier parity)
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
            float* eb = eb_t + tsc * HP;
            int* path = path_t + tsc * (S_max + 2);
            HotSel* hs = hs_t + (size_t)tsc * N;
            double* hw = hw_t + (size_t)tsc * N;
            int* nxt = nxt_t + (size_t)tsc * N;
            const uint32_t hs_addr = smem_addr(hs), path_addr = smem_addr(path);
            const uint32_t hw_addr = smem_addr(hw), nxt_addr = smem_addr(nxt);

            // softmax of the logits of expansion k (buffer `buf`) -> priors of its children (mcts.py:189-196)
            auto finish_expansion = [&](int k, int buf) {
                const float* lg = lg_t + buf * TT * HP + tsc * HP;
                const float ssum = softmax_team<G>(lg, eb, A_net, g, active);
                if (active) {
                    const int cbase = 1 + k * A;
                    for (int a = g; a < A; a += G) {
                        const float p32 = prior_div(eb[a], ssum);
                        store_node(nodes + cbase + a, 0.0, 0, p32);
                        sts64_s(hs_addr + 32u * (uint32_t)(cbase + a) + 8u, __dmul_rn(p.c_puct, (double)p32));
                    }
                }
                __syncwarp();
            };

            // ---- root: softmax of the root logits -> expand -> noise (mcts.py:126-136) -------------
            named_bar_sync(bar_team, cnt_team);  // the previous batch's logits have been consumed
            if (active && p.noise_mode == 2 && g == 0)
                dirichlet_row_cold(rprior, A, p.alpha, p.seed, p.tree_base + (unsigned long long)t);
            named_bar_sync(bar_team, cnt_team);  // root logits published
            __syncwarp();
            {
                const float* lg = lg_t + TT * HP + tsc * HP;
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
                        nxt[1 + a] = -1;
                    }
                    if (g == 0) {
                        store_node(nodes, 0.0, 0, 1.0f);
                        HotSel h;
                        h.q = 0.0; h.cp = 0.0; h.r = 1.0; h.n = 0; h.ce = 0;
                        hs[0] = h;
                        hw[0] = 0.0;
                        nxt[0] = 1;  // every child unvisited: the first one
                        exp_node[0] = 0;
                    }
                }
            }
            __syncwarp();
            CM_CLK(c3, hs[0].ce)

            // ---- S simulations (mcts.py:139-140, 157-203) -------------------------------------------
            long long depth = 0;
            for (int s = 0; s < p.S; ++s) {
                int plen;
                const int leaf = select_chase(nxt_addr, path_addr, g, active, plen);
                const unsigned lm1 = (unsigned)(leaf - 1);
                const int parent_e = A > 1 ? (int)__umulhi(lm1, magic_A) : (int)lm1;
                const int action = (int)lm1 - parent_e * A;
                if (g == 0 && ts < TT) {
                    sel[ts] = active ? parent_e : 0;
                    sel[TT + ts] = active ? action : 0;
                    sel[2 * TT + ts] = active ? plen : 0;
                }
                CM_CLK(c0, plen + leaf)
                named_bar_sync(bar_team, cnt_team);  // (A) selection published
                CM_CLK(c6, sel[0])
                // Node.expand of the leaf (mcts.py:78-89, 195-196) without the priors: the children exist and are
                // unvisited, so the next walks score them +inf whatever the prior is.  The priors follow one
                // simulation later (finish_expansion), before any of these children can have been visited.
                const int k = s + 1;  // expansion index of the leaf
                if (active) {
                    const int cbase = 1 + k * A;
                    for (int a = g; a < A; a += G) {
                        const uint32_t rec = hs_addr + 32u * (uint32_t)(cbase + a);
                        sts128_s(rec, 0, 0, 0, 0);                    // q = 0, cp = 0 (set by finish_expansion)
                        sts128_s(rec + 16u, 0, 0x3ff00000, 0, -1);    // r = 1.0, n = 0, ce = -1
                        sts64_s(hw_addr + 8u * (uint32_t)(cbase + a), 0.0);
                        sts32_s(nxt_addr + 4u * (uint32_t)(cbase + a), -1);
                    }
                    if (g == 0) {
                        sts32_s(hs_addr + 32u * (uint32_t)leaf + 28u, k);
                        sts32_s(nxt_addr + 4u * (uint32_t)leaf, cbase);
                        exp_node[k] = leaf;
                    }
                    depth += plen;
                }
                __syncwarp();
                CM_CLK(c1, hs[1].n)
                if (s > 0) finish_expansion(s, (s - 1) & 1);  // the leaf of the previous simulation
                CM_CLK(c5, hs[1].n)
                named_bar_sync(bar_team, cnt_team);  // (C) the network warps have backed the value up
                __syncwarp();
                CM_CLK(c2, hs[0].n)
            }
            if (p.S > 0) finish_expansion(p.S, (p.S - 1) & 1);

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
                    if (p.out_actions) {   // first maximum of the visit counts = np.argmax of the probabilities
                        int best = 0, bc = hs[1].n;
                        for (int a = 1; a < A; ++a) { const int c = hs[1 + a].n; if (c > bc) { bc = c; best = a; } }
                        p.out_actions[t] = bc == 0 ? 0 : best;
                    }
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
            const uint32_t hs_a = smem_addr(hs_t), hw_a = smem_addr(hw_t), nxt_a = smem_addr(nxt_t), path_a = smem_addr(path_t);
            const uint32_t sq_a = smem_addr(sq_s), rc_a = smem_addr(rc_s);
            const BackupLanes bl = backup_lanes(p.tpt, p.nmw, mw, lane);   // the path entries of the team's trees over all network lanes

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
            team_heads<HS>(wht_s, bh_s, xt, H, HP, TT, A_net, lg_t + TT * HP, mlane, nml);
            named_bar_sync(bar_team, cnt_team);  // root logits published

            for (int s = 0; s < p.S; ++s) {
                float* lg_s = lg_t + (s & 1) * TT * HP;
                named_bar_sync(bar_team, cnt_team);  // (A) selection published
                CM_CLK(c3, sel[0])
                if constexpr (KR > 0) {
                    // gather, row-major: x0[ts][j] = hidden[ts][parent_e][j] + Emb[action][j]   (network.py:229-230); 16 lanes
                    // per row, every load and store 16 bytes and conflict-free; all loads of a lane in flight together
                    constexpr int PER = (TT * 16 + 127) / 128;   // items per lane with four network warps
                    float4 hv[PER], ev[PER];
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int i = mlane + u * nml;
                        const int ts = i >> 4, c = i & 15;
                        hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        ev[u] = hv[u];
                        if (i < TT * 16 && ts < n_here) {
                            const int pe = sel[ts], ac = sel[TT + ts];
                            hv[u] = *reinterpret_cast<const float4*>(hid0 + (size_t)ts * hstride + (size_t)pe * H + 4 * c);
                            ev[u] = *reinterpret_cast<const float4*>(emb_s + (size_t)ac * H + 4 * c);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PER; ++u) {
                        const int i = mlane + u * nml;
                        if (i < TT * 16)
                            *reinterpret_cast<float4*>(x0 + (size_t)(i >> 4) * RS + 4 * (i & 15)) =
                                make_float4(fadd(hv[u].x, ev[u].x), fadd(hv[u].y, ev[u].y), fadd(hv[u].z, ev[u].z), fadd(hv[u].w, ev[u].w));
                    }
                    for (int i = mlane + PER * nml; i < TT * 16; i += nml) {   // fewer than four network warps
                        const int ts = i >> 4, c = i & 15;
                        float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = h4;
                        if (ts < n_here) {
                            h4 = *reinterpret_cast<const float4*>(hid0 + (size_t)ts * hstride + (size_t)sel[ts] * H + 4 * c);
                            e4 = *reinterpret_cast<const float4*>(emb_s + (size_t)sel[TT + ts] * H + 4 * c);
                        }
                        *reinterpret_cast<float4*>(x0 + (size_t)ts * RS + 4 * c) = make_float4(fadd(h4.x, e4.x), fadd(h4.y, e4.y), fadd(h4.z, e4.z), fadd(h4.w, e4.w));
                    }
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(hand + 4 * team);          // X0R: rows handed to the layer warpgroup
                    CM_CLK(c0, __float_as_int(x0[0]))
                    tc_mbar_wait_spin(hand + 4 * team + 3, job & 1u);             // X2R: new hidden rows in xt
                    ++job;
                    CM_CLK(c4, __float_as_int(xt[0]))
                } else {
                // gather: x0[j][ts] = hidden[ts][parent_e][j] + Emb[action][j]   (network.py:229-230).
                // Two items per lane and pass with both global loads in flight before either is used (the
                // parent rows come from L2); consecutive lanes read consecutive 16-byte chunks of a row (the transposed stores
                // below are 8-way bank conflicted; the conflict-free mapping, trees fastest, un-coalesces the L2 reads and
                // measured slower: 1 320 vs 1 100 cycles).
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
                }
                CM_CLK(c1, __float_as_int(xt[0]))
                team_heads<HS>(wht_s, bh_s, xt, H, HP, TT, A_net + 1, lg_s, mlane, nml);
                named_bar_sync(bar_mlp, cnt_mlp);  // logits + value visible to every network warp
                CM_CLK(c2, __float_as_int(lg_s[0]))
                backup_rescore(hs_a, hw_a, nxt_a, path_a, sel, lg_s, sq_a, rc_a, N, A, S_max, TT, HP, A_net, n_here, bl);
                CM_CLK(c5, hs_t[0].n)
                named_bar_sync(bar_team, cnt_team);  // (C) statistics and selection cache updated; logits of s published
            }
        }
    }
#undef CM_CLK
    if (clk_on) {
        long long* o = p.clk + (size_t)blockIdx.x * 16 + (is_tree ? 0 : 8);
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3; o[4] = c4; o[5] = c5; o[6] = c6;
    }
}

template <int G, int TT, int TQ, int NV, int HS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) search_team_kernel(const TeamParams p) {
    search_team_body<G, TT, TQ, NV, HS, 0>(p);
}
// Layer-warpgroup variants: the register budget is stated directly (ptxas derives it from __launch_bounds__ with the
// thread count rounded up to whole warpgroups, which would give a 14-warp CTA 128 registers instead of 144).
template <int G, int TT, int TQ, int NV, int HS, int NREG, int KR>
__global__ void __maxnreg__(NREG) search_team_lw_kernel(const TeamParams p) {
    search_team_body<G, TT, TQ, NV, HS, KR, (NREG == 128 ? 104 : 152)>(p);
}

// ------------------------------------------------------------------------------------------------
// host-side launch: the instantiations that exist (team_variant_ok) and the dispatch over them
// ------------------------------------------------------------------------------------------------
template <int G, int TT, int TQ, int NV, int HS, int NREG, int KR>
static cudaError_t launch_one_lw(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    auto k = search_team_lw_kernel<G, TT, TQ, NV, HS, NREG, KR>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel_pdl(k, grid, block, smem, st, p.pdl != 0, p);
}

template <int G, int TT, int TQ, int NV, int HS, int MAXT>
static cudaError_t launch_one(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    auto k = search_team_kernel<G, TT, TQ, NV, HS, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel_pdl(k, grid, block, smem, st, p.pdl != 0, p);
}

// lane tiles of the network warps (TT tree slots, TQ trees x NV neurons per lane)
template <int G, int TT, int TQ, int NV>
static cudaError_t launch_t(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    const bool big = block > 512;
    if (p.pk.H == 64) {
        if (block <= 384) return launch_one<G, TT, TQ, NV, 64, 384>(p, grid, block, smem, st);
        return big ? launch_one<G, TT, TQ, NV, 64, 1024>(p, grid, block, smem, st) : launch_one<G, TT, TQ, NV, 64, 512>(p, grid, block, smem, st);
    }
    return big ? launch_one<G, TT, TQ, NV, 0, 1024>(p, grid, block, smem, st) : launch_one<G, TT, TQ, NV, 0, 512>(p, grid, block, smem, st);
}

// layer-warpgroup instantiations: hidden_dim 64, G in {2,4,8}; register-resident k per block size (KR)
bool team_lw_variant_ok(int G, int TT, int H) { return H == 64 && (TT == 8 || TT == 16) && (G == 2 || G == 4 || G == 8); }
// warps per team in layer-warpgroup mode: the CTA must consist of whole warpgroups (4 layer warps + n_team * team warps)
int team_lw_team_warps(int n_team) { return n_team == 1 ? 8 : (n_team == 2 ? 6 : (n_team == 3 ? 4 : 3)); }

template <int G, int TT>
static cudaError_t launch_lw(const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    constexpr int TQ = TT == 16 ? 4 : 2;
    // 16 warps launched with 128 registers per thread, 12 warps with 168 (setmaxnreg re-divides them inside the kernel)
    if (block == 512) return launch_one_lw<G, TT, TQ, 2, 64, 128, 64>(p, grid, block, smem, st);
    if (block == 384) return launch_one_lw<G, TT, TQ, 2, 64, 168, 64>(p, grid, block, smem, st);
    return cudaErrorInvalidValue;
}

bool team_variant_ok(int G, int TT) {
    if (TT == 16) return G == 2 || G == 4 || G == 8;
    if (TT == 8) return G == 2 || G == 4 || G == 8 || G == 16 || G == 32;
    return false;
}

// p.tile: 0 = default (16: 4x2, 8: 2x2); tuning alternatives exist for G = 2, Hp = 64, <= 512 threads only
cudaError_t launch_search_team(int G, const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    if (!team_variant_ok(G, p.TT)) return cudaErrorInvalidValue;
    if (p.lw > 0) {
        if (!team_lw_variant_ok(G, p.TT, p.pk.H)) return cudaErrorInvalidValue;
        if (p.TT == 16) {
            switch (G) {
                case 2: return launch_lw<2, 16>(p, grid, block, smem, st);
                case 4: return launch_lw<4, 16>(p, grid, block, smem, st);
                case 8: return launch_lw<8, 16>(p, grid, block, smem, st);
            }
        } else {
            switch (G) {
                case 2: return launch_lw<2, 8>(p, grid, block, smem, st);
                case 4: return launch_lw<4, 8>(p, grid, block, smem, st);
                case 8: return launch_lw<8, 8>(p, grid, block, smem, st);
            }
        }
        return cudaErrorInvalidValue;
    }
    if (p.tile != 0) {
        if (p.TT == 16 && p.tile == 2 && G == 2 && p.pk.H == 64 && block <= 640) return launch_one<2, 16, 2, 2, 64, 640>(p, grid, block, smem, st);
        if (p.TT == 16 && p.tile == 3 && G == 2 && p.pk.H == 64 && block <= 640) return launch_one<2, 16, 4, 1, 64, 640>(p, grid, block, smem, st);
        if (G != 2 || p.pk.H != 64 || block > 512) return cudaErrorInvalidValue;
        if (p.TT == 16 && p.tile == 1) return launch_one<2, 16, 4, 4, 64, 512>(p, grid, block, smem, st);
        if (p.TT == 8 && p.tile == 1) return block <= 384 ? launch_one<2, 8, 4, 2, 64, 384>(p, grid, block, smem, st) : launch_one<2, 8, 4, 2, 64, 512>(p, grid, block, smem, st);
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
