// search_tc.cu — MCTS.search (src/cumind/core/mcts.py:115-155) for B trees in ONE launch with the per-simulation
// network on the 5th-generation tensor cores (CUMIND_NET_BF16_TC).
//
// A CTA owns 128 trees: thread i = tree i = row i of a 128-row UMMA tile = TMEM lane i.  Every simulation
//   select   each thread walks ITS tree from the root (mcts.py:165-171; Node.ucb_score 51-65 in float64 with the
//            table-driven exact division, first maximum in action order) — the node records are in HBM / L2,
//            128 independent walks per CTA and several CTAs per SM hide the latency;
//   expand   x = hidden[parent] + Emb[action] becomes row i of the bf16 A tile (UMMA K-major core-matrix layout);
//            per dynamics layer H/16 tcgen05.mma (M=128, N=H, K=16; SASS UTCHMMA) issued by one thread accumulate
//            in TMEM, completion through tcgen05.commit -> mbarrier; the epilogue reads the accumulators with
//            tcgen05.ld.32x32b (LDTM), applies bias / ReLU / residual in fp32 registers and writes the next A tile;
//            the heads (policy | value) are one more MMA with N = 16 (network.py:217-236, 254-266);
//   softmax -> children priors (mcts.py:189-196), backup of the parents on the path, the root twice (199-203).
// The bf16 weight tiles are staged once per CTA with TMA bulk copies (UBLKCP).  The tree logic is the reference's,
// in float64, exactly as in the fp32-exact kernels; only the network arithmetic differs (bf16 operands, fp32
// accumulation), so this mode is judged on tolerances (tests/test_gpu_parity.py), not bit for bit.
#include <cuda_bf16.h>

#include <type_traits>

#include "kernels.h"
#include "search_core.cuh"
#include "tc_common.cuh"

namespace cm {

// CACHE: the child the next visit of a node selects (team_common.cuh: select_child only changes for nodes on the path of
// the last backup) and the path are kept in shared memory as 16-bit node indices, [node][tree]: the walk is a pointer
// chase of one shared-memory load per level instead of one L2 round trip per level, and the backup re-scores the
// children of every parent on the path with the record loads of several levels in flight together.
// Without CACHE (trees whose node count does not fit) every walk scores its way down from HBM / L2.
template <int H, bool CACHE>
__global__ void __launch_bounds__(128) search_tc_kernel(const SearchTcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcPackLayout& pk = p.pk;
    unsigned char* w_s = smem;                        // the whole pack (bf16 tiles, fp32 biases, fp32 embedding)
    unsigned char* a_s = smem + pk.total_bytes;       // A tile: 128 x H bf16
    double* sq_s = reinterpret_cast<double*>(a_s + 128 * H * 2);
    double* rc_s = sq_s + p.n_sqrt;
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(rc_s + p.n_sqrt);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    uint16_t* nxt_s = reinterpret_cast<uint16_t*>(bar_w + 4);            // CACHE: [N][128]
    uint16_t* path_s = nxt_s + (size_t)p.tv.N * 128;                      // CACHE: [S_max + 2][128]
    constexpr int NP = 16;
    constexpr uint32_t TMEM_COLS = (H + NP <= 32) ? 32 : (H + NP <= 64) ? 64 : (H + NP <= 128) ? 128 : 256;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = pk.L, A_net = pk.A, A = p.tv.A, N = p.tv.N, S_max = p.tv.S_max;
    // (c - 1) / A by multiplication: exact for c - 1 < 2^32 / A
    const unsigned magic_A = A > 1 ? (unsigned)((0x100000000ULL + (unsigned)A - 1u) / (unsigned)A) : 0u;
    auto block_of = [&](int c) { return A > 1 ? 1 + (int)__umulhi((unsigned)(c - 1), magic_A) * A : c; };   // first child of c's sibling block

    if (tid == 0) {
        tc_mbar_init(bar_w, 1);
        tc_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        tc_mbar_expect_tx(bar_w, (uint32_t)pk.total_bytes);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < (uint32_t)pk.total_bytes; o += CH) {
            const uint32_t n = (uint32_t)pk.total_bytes - o < CH ? (uint32_t)pk.total_bytes - o : CH;
            tc_bulk_g2s(w_s + o, p.pack + o, n, bar_w);
        }
    }
    for (int i = tid; i < p.n_sqrt; i += 128) {
        sq_s[i] = __dsqrt_rn((double)i);
        rc_s[i] = __ddiv_rn(1.0, (double)i);  // RN(1/i); entry 0 is never used
    }
    tc_mbar_wait(bar_w, 0);
    __syncthreads();
    pdl_wait();  // root_hidden (and the noise) may come from the previous kernel of the stream

    const float* bias_s = reinterpret_cast<const float*>(w_s + pk.bias_off);   // [L][H]
    const float* bh_s = reinterpret_cast<const float*>(w_s + pk.bh_off);       // [NP]
    const float* emb_s = reinterpret_cast<const float*>(w_s + pk.emb_off);     // [A][H]
    const uint32_t a_addr = smem_addr(a_s);
    const uint32_t idesc_layer = umma_idesc(128, H);
    const uint32_t idesc_heads = umma_idesc(128, NP);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t mma_phase = 0;

    // row `tid` of the A tile = bf16(x)
    auto write_a_row = [&](const float (&x)[H]) {
#pragma unroll
        for (int kc = 0; kc < H / 8; ++kc) {
            uint32_t w4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(x[kc * 8 + 2 * i], x[kc * 8 + 2 * i + 1]);
                w4[i] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            *reinterpret_cast<uint4*>(a_s + ((size_t)(kc * 16 + (tid >> 3)) * 64 + (tid & 7) * 8) * 2) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    // D[128 x Ncols] (TMEM column offset col0) = A tile x B tile at byte offset b_off of the pack; returns when done
    auto mma_tile = [&](size_t b_off, int Ncols, uint32_t col0, uint32_t idesc) {
        if (tid == 0) {
            tc_fence_after();
            const uint32_t b_addr = smem_addr(w_s + b_off);
            const uint32_t a_lbo = 16 * 128, b_lbo = (uint32_t)(Ncols / 8) * 128;
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                const uint64_t ad = umma_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128);
                const uint64_t bd = umma_desc(b_addr + ks * 2 * b_lbo, b_lbo, 128);
                umma_f16(tmem_base + col0, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
        tc_mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
    };
    // policy logits / value of the rows in the A tile: acc[0..A_net) logits, acc[A_net] value
    auto heads_tile = [&](float (&acc)[16]) {
        mma_tile(pk.heads_off, NP, H, idesc_heads);
        tmem_ld16(lane_taddr + H, acc);
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += bh_s[i];
        tc_fence_before();
        __syncthreads();  // TMEM and the A tile are reused by the next MMA
    };
    // softmax of acc[0..A_net) with the pinned exp; e[i] and the sum in fp32 (children priors = e[i] / s)
    auto softmax14 = [&](const float (&acc)[16], float (&e)[14], float& s) {
        float m = acc[0];
#pragma unroll
        for (int i = 1; i < 14; ++i) if (i < A_net && acc[i] > m) m = acc[i];
        s = 0.f;
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            e[i] = 0.f;
            if (i < A_net) { e[i] = __double2float_rn(pexp((double)__fsub_rn(acc[i], m))); s = fadd(s, e[i]); }
        }
    };

    for (int row0 = blockIdx.x * 128; row0 < p.tv.B; row0 += gridDim.x * 128) {
        const int r = row0 + tid;
        const bool live = r < p.tv.B;
        const size_t rr = live ? (size_t)r : 0;
        Node* nodes = p.tv.nodes + rr * N;
        int32_t* eidx = p.tv.child_eidx + rr * N;
        double* rprior = p.tv.root_prior + rr * A;
        int32_t* exp_node = p.tv.exp_node + rr * (S_max + 1);
        float* hidden = p.tv.hidden + rr * (size_t)(S_max + 1) * H;
        int32_t* path = p.tv.path + rr * (S_max + 2);

        // ---- root: prediction -> softmax -> expand -> noise (mcts.py:126-136) -------------------------------
        // The root logits are computed once per search in pinned-order fp32 (the chain of the fp32-exact kernels and of
        // the per-phase init_root kernel), not on the tensor cores: H * A fused multiply-adds per tree buy a root policy
        // without bf16 rounding, and make this kernel and the per-phase tensor-core path agree bit for bit.
        if (live) {
            const float* xr = p.root_hidden + rr * H;
#pragma unroll
            for (int j = 0; j < H; j += 4) *reinterpret_cast<float4*>(hidden + j) = *reinterpret_cast<const float4*>(xr + j);
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
            const float* Wh = p.pack_f32 + p.pkf.heads_w;
            const float* bh = p.pack_f32 + p.pkf.heads_b;
            const int HP = p.pkf.HP;
            for (int k = 0; k < H; ++k) {
                const float xk = xr[k];
#pragma unroll
                for (int i = 0; i < 14; ++i)
                    if (i < A_net) acc[i] = ffma(xk, Wh[(size_t)k * HP + i], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < 14; ++i)
                if (i < A_net) acc[i] = fadd(acc[i], bh[i]);
            if (p.noise_mode == 2) dirichlet_row(rprior, A, p.alpha, p.seed, p.tree_base + (unsigned long long)r);
            float e[14], s;
            softmax14(acc, e, s);
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                if (i < A) {
                    const float p32 = __fdiv_rn(e[i], s);
                    double pd = (double)p32;
                    if (p.noise_mode == 1) pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(p.noise[rr * A + i], p.frac));
                    else if (p.noise_mode == 2) pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(rprior[i], p.frac));
                    rprior[i] = pd;
                    store_node(nodes + 1 + i, 0.0, 0, p32);
                    eidx[1 + i] = -1;
                }
            }
            store_node(nodes, 0.0, 0, 1.0f);
            eidx[0] = 0;
            exp_node[0] = 0;
            if constexpr (CACHE) {
                nxt_s[tid] = 1;   // every root child unvisited: the first one
                for (int a = 0; a < A; ++a) nxt_s[(1 + a) * 128 + tid] = 0xFFFF;
            }
        }

        // ---- S simulations (mcts.py:139-140, 157-203) ---------------------------------------------------------
        long long depth = 0;
        for (int s = 0; s < p.S; ++s) {
            int node = 0, plen = 0;
            if constexpr (CACHE) {
                // selection: follow the cached choices from the root to the first unexpanded node
                if (live) {
                    int c = nxt_s[tid];
                    while (c != 0xFFFF) {
                        path_s[plen * 128 + tid] = (uint16_t)node;
                        ++plen;
                        node = c;
                        c = nxt_s[node * 128 + tid];
                    }
                    depth += plen;
                }
            } else {
            // selection: walk from the root while the node is expanded.  One dependent memory round trip per level: the
            // records and the expansion indices of all children of a node are requested together, and the parent's visit
            // count for the next level is the chosen child's, already loaded.
            if (live) {
                int e = 0;                                   // the root is expanded
                int n_parent = load_node(nodes).visit;
                while (e >= 0) {
                    const int base = 1 + e * A;
                    const double sq = sq_s[n_parent];
                    int best = 0, best_e = -1, best_n = 0;
                    double best_sc = 0.0;
                    for (int a = 0; a < A; ++a) {
                        const Node nd = load_node(nodes + base + a);
                        const int ce = eidx[base + a];
                        const double pr = node == 0 ? rprior[a] : (double)nd.prior;
                        const double sc = ucb_score_tab(nd.visit, nd.value_sum, pr, sq, p.c_puct, rc_s);
                        if (a == 0 || sc > best_sc) { best = a; best_sc = sc; best_e = ce; best_n = nd.visit; }   // first maximum (Python max)
                    }
                    path[plen++] = node;
                    node = base + best;
                    e = best_e;
                    n_parent = best_n;
                }
                depth += plen;
            }
            }
            const int kexp = s + 1;                            // expansion index of the leaf
            const int pe = live ? (A > 1 ? (int)__umulhi((unsigned)(node - 1), magic_A) : node - 1) : 0;
            const int act = live ? (node - 1) - pe * A : 0;
            // expansion: dynamics + prediction on the tensor cores
            float x[H];
            {
                const float* hsrc = hidden + (size_t)pe * H;
#pragma unroll
                for (int j = 0; j < H; j += 4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (live) v = *reinterpret_cast<const float4*>(hsrc + j);
                    const float4 em = *reinterpret_cast<const float4*>(emb_s + (size_t)act * H + j);
                    x[j] = v.x + em.x; x[j + 1] = v.y + em.y; x[j + 2] = v.z + em.z; x[j + 3] = v.w + em.w;
                }
            }
            for (int l = 0; l < L; ++l) {
                write_a_row(x);
                mma_tile(pk.layer_off + (size_t)l * H * H * 2, H, 0, idesc_layer);
#pragma unroll
                for (int c = 0; c < H / 16; ++c) {
                    float acc[16];
                    tmem_ld16(lane_taddr + c * 16, acc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float t = acc[i] + bias_s[l * H + c * 16 + i];
                        x[c * 16 + i] = (t > 0.f ? t : 0.f) + x[c * 16 + i];
                    }
                }
                tc_fence_before();
                __syncthreads();
            }
            if (live) {
                float* dst = hidden + (size_t)kexp * H;
#pragma unroll
                for (int j = 0; j < H; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
            }
            write_a_row(x);
            float acc[16];
            heads_tile(acc);
            if (live) {
                float val = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) if (i == A_net) val = acc[i];
                float e[14], ssum;
                softmax14(acc, e, ssum);
                const int cbase = 1 + kexp * A;
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    if (i < A) {
                        store_node(nodes + cbase + i, 0.0, 0, __fdiv_rn(e[i], ssum));
                        eidx[cbase + i] = -1;
                    }
                }
                eidx[node] = kexp;
                exp_node[kexp] = node;
                const double v = (double)val;
                if constexpr (CACHE) {
                    // Node.expand: the leaf's next visit takes its first child (all scores +inf, max() keeps the first)
                    nxt_s[node * 128 + tid] = (uint16_t)cbase;
                    for (int a = 0; a < A; ++a) nxt_s[(cbase + a) * 128 + tid] = 0xFFFF;
                    // backup (mcts.py:199-203) from the root down to the leaf's parent, and for every parent P the choice its
                    // next visit will make (Node.select_child, 67-76).  The only records read are the children blocks of
                    // the parents: P's own old record sits in the block of ITS parent (the root's is loaded once), and the
                    // new statistics of the child on the path are its old ones + v, + 1 (what its own turn as a parent
                    // stores).  Block addresses come from shared memory alone, so D blocks are requested together and
                    // a path costs ceil(len / D) memory round trips instead of one per level.
                    auto backup = [&](auto AB_, auto D_) {
                        constexpr int AB = decltype(AB_)::value, D = decltype(D_)::value;
                        int4 rec = *reinterpret_cast<const int4*>(nodes);   // old record of the next parent (the root first)
                        for (int i0 = 0; i0 < plen; i0 += D) {
                            int4 blk[D][AB];
                            int base[D];
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                base[d] = 1;
                                if (i0 + d < plen) {
                                    const int P = path_s[(i0 + d) * 128 + tid];
                                    const int old = nxt_s[P * 128 + tid];
                                    base[d] = block_of(old);
                                }
#pragma unroll
                                for (int a = 0; a < AB; ++a)
                                    if (a < A && i0 + d < plen) blk[d][a] = *reinterpret_cast<const int4*>(nodes + base[d] + a);
                            }
#pragma unroll
                            for (int d = 0; d < D; ++d) {
                                const int i = i0 + d;
                                if (i < plen) {
                                    const int P = path_s[i * 128 + tid];
                                    const bool root = i == 0;
                                    const double w1 = __dadd_rn(__hiloint2double(rec.y, rec.x), v);
                                    const double wP = root ? __dadd_rn(w1, v) : w1;
                                    const int nP = rec.z + (root ? 2 : 1);
                                    store_node(nodes + P, wP, nP, __int_as_float(rec.w));
                                    const int C = i + 1 < plen ? (int)path_s[(i + 1) * 128 + tid] : -1;
                                    const double sq = sq_s[nP];
                                    int best = 0;
                                    double best_sc = 0.0;
#pragma unroll
                                    for (int a = 0; a < AB; ++a) {
                                        if (a < A) {
                                            const int4 r4 = blk[d][a];
                                            int n = r4.z;
                                            double w = __hiloint2double(r4.y, r4.x);
                                            if (base[d] + a == C) { rec = r4; n += 1; w = __dadd_rn(w, v); }
                                            const double pr = root ? rprior[a] : (double)__int_as_float(r4.w);
                                            const double sc = ucb_score_tab(n, w, pr, sq, p.c_puct, rc_s);
                                            if (a == 0 || sc > best_sc) { best = a; best_sc = sc; }
                                        }
                                    }
                                    nxt_s[P * 128 + tid] = (uint16_t)(base[d] + best);
                                }
                            }
                        }
                    };
                    if (A <= 2) backup(std::integral_constant<int, 2>{}, std::integral_constant<int, 4>{});
                    else if (A <= 4) backup(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{});
                    else backup(std::integral_constant<int, 14>{}, std::integral_constant<int, 1>{});
                } else {
                    // backup (mcts.py:199-203): every parent on the path, then the root once more
                    for (int i = plen - 1; i >= 0; --i) {
                        Node* q = nodes + path[i];
                        const Node nd = load_node(q);
                        const bool root = i == 0;
                        const double w1 = __dadd_rn(nd.value_sum, v);
                        store_node(q, root ? __dadd_rn(w1, v) : w1, nd.visit + (root ? 2 : 1), nd.prior);
                    }
                }
            }
        }

        // ---- visit counts -> probabilities (mcts.py:143-155) ---------------------------------------------------
        if (live) {
            long long total = 0;
            for (int a = 0; a < A; ++a) total += load_node(nodes + 1 + a).visit;
            for (int a = 0; a < p.A_cfg; ++a) {
                const int c = a < A ? load_node(nodes + 1 + a).visit : 0;
                p.out_visits[rr * p.A_cfg + a] = c;
                p.out_probs[rr * p.A_cfg + a] = total == 0 ? __ddiv_rn(1.0, (double)p.A_cfg) : __ddiv_rn((double)c, (double)total);
            }
            const Node rt = load_node(nodes);
            if (p.out_root_value) p.out_root_value[r] = rt.visit == 0 ? 0.0 : __ddiv_rn(rt.value_sum, (double)rt.visit);
            if (p.out_actions) {   // first maximum of the visit counts = np.argmax of the probabilities
                int best = 0, bc = load_node(nodes + 1).visit;
                for (int a = 1; a < A; ++a) { const int c = load_node(nodes + 1 + a).visit; if (c > bc) { bc = c; best = a; } }
                p.out_actions[r] = bc == 0 ? 0 : best;
            }
            p.tv.n_expanded[r] = p.S + 1;
            p.tv.depth_sum[r] = depth;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int H, bool CACHE>
static cudaError_t launch_hc(const SearchTcParams& p, int sm_count, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(search_tc_kernel<H, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // CTAs per SM: shared memory and the 512 TMEM columns (each CTA allocates the power of two >= H + 16)
    constexpr int tmem_cols = (H + 16 <= 32) ? 32 : (H + 16 <= 64) ? 64 : (H + 16 <= 128) ? 128 : 256;
    int per_sm = 512 / tmem_cols;
    const int by_smem = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    const int tiles = (p.tv.B + 127) / 128;
    const int cap = sm_count * per_sm;
    const int grid = tiles < cap ? tiles : cap;
    return launch_kernel_pdl(search_tc_kernel<H, CACHE>, grid, 128, smem, st, p.pdl != 0, p);
}

template <int H>
static cudaError_t launch_h(const SearchTcParams& p, int sm_count, cudaStream_t st) {
    const size_t base = p.pk.total_bytes + (size_t)128 * H * 2 + (size_t)2 * p.n_sqrt * sizeof(double) + 64;
    const size_t cache = ((size_t)p.tv.N + (size_t)p.tv.S_max + 2) * 128 * sizeof(uint16_t);
    if (p.tv.N < 0xFFFF && base + cache <= 200 * 1024) return launch_hc<H, true>(p, sm_count, base + cache, st);
    return launch_hc<H, false>(p, sm_count, base, st);
}

cudaError_t launch_search_tc(const SearchTcParams& p, int sm_count, cudaStream_t st) {
    switch (p.pk.H) {
        case 16: return launch_h<16>(p, sm_count, st);
        case 32: return launch_h<32>(p, sm_count, st);
        case 64: return launch_h<64>(p, sm_count, st);
        case 128: return launch_h<128>(p, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cm
