// search_fused.cu — MCTS.search (src/cumind/core/mcts.py:115-155) for B trees in ONE launch.
//
// Persistent CTAs (grid <= #SMs x CTAs/SM).  Each CTA stages the recurrent-path parameters
// (embedding, L dynamics layers, policy/value/reward heads: PackLayout) into shared memory with
// TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier, builds a table of
// correctly rounded sqrt(N) for the PUCT term, then every group of G lanes runs the complete
// search of one tree:  root prediction + softmax + expand (+ Dirichlet mix), then S x
// { select -> dynamics + prediction -> expand -> backup }, then the visit-count normalisation.
// There is no block or grid barrier inside the simulation loop: trees never wait for each other.
// Tree records live in HBM (cm::Node, 16 B) and are served from L1/L2 while a tree is being searched.
#include "search_core.cuh"
#include "kernels.h"

namespace cm {


CM_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

CM_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
CM_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
CM_DEVICE void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`.
CM_DEVICE void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// GS = lanes per tree in the select / expand / backup phases (TW = 32/GS trees per warp);
// R = neurons per lane in the warp-cooperative network tile (32*R >= H).

template <int GS, int R>
__global__ void __launch_bounds__(512, 1) search_fused_kernel(const FusedParams p) {
    constexpr int TW = 32 / GS;
    extern __shared__ __align__(128) unsigned char smem[];
    float* w_s = reinterpret_cast<float*>(smem);
    double* sq_s = reinterpret_cast<double*>(smem + p.off_sqrt);
    double* rc_s = reinterpret_cast<double*>(smem + p.off_recip);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + p.off_bar);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane & (GS - 1);   // lane within the tree's sub-group
    const int tw = lane / GS;        // tree within the warp
    const int TPB = (blockDim.x >> 5) * TW;
    const int H = p.pk.H, L = p.pk.L, HP = p.pk.HP, A = p.tv.A, A_net = p.A_net;

    // ---- stage parameters (TMA bulk) + sqrt table -------------------------------------------
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = (uint32_t)(p.pk.total * sizeof(float));
        mbar_expect_tx(bar, total);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < total; o += CH) {
            const uint32_t n = total - o < CH ? total - o : CH;
            tma_bulk_g2s(smem + o, reinterpret_cast<const unsigned char*>(p.pack) + o, n, bar);
        }
    }
    for (int i = tid; i < p.n_sqrt; i += blockDim.x) {
        sq_s[i] = __dsqrt_rn((double)i);
        rc_s[i] = __ddiv_rn(1.0, (double)i);  // RN(1/i); entry 0 is never used
    }
    mbar_wait(bar, 0);
    __syncthreads();
    pdl_wait();  // root_hidden (and the noise) may come from the previous kernel of the stream

    const float* emb_s = w_s + p.pk.emb;
    const float* wh_s = w_s + p.pk.heads_w;
    const float* bh_s = w_s + p.pk.heads_b;

    // per-warp scratch: x0,x1 [H][TW] | lg,eb [TW][HP] | path [TW][S_max+2]
    unsigned char* my = smem + p.off_slots + (size_t)warp * p.slot_bytes;
    float* x0 = reinterpret_cast<float*>(my);
    float* x1 = x0 + (size_t)H * TW;
    float* lg_w = x1 + (size_t)H * TW;
    float* eb_w = lg_w + TW * HP;
    float* lg = lg_w + tw * HP;
    float* eb = eb_w + tw * HP;
    int* path_s = reinterpret_cast<int*>(eb_w + TW * HP) + tw * (p.tv.S_max + 2);
    const int j0 = lane * R;
    const bool own = j0 < H;
    const size_t hstride = (size_t)(p.tv.S_max + 1) * H;

    for (int t0 = blockIdx.x * TPB; t0 < p.tv.B; t0 += gridDim.x * TPB) {
        const int tbase = t0 + warp * TW;          // first tree of this warp
        const int t = tbase + tw;
        const bool active = t < p.tv.B;
        const size_t tt = active ? (size_t)t : 0;
        Node* nodes = p.tv.nodes + tt * p.tv.N;
        int32_t* eidx = p.tv.child_eidx + tt * p.tv.N;
        double* rprior = p.tv.root_prior + tt * A;
        int32_t* exp_node = p.tv.exp_node + tt * (p.tv.S_max + 1);
        float* hidden = p.tv.hidden + tt * hstride;
        int* path = p.path_in_smem ? path_s : (p.tv.path + tt * (p.tv.S_max + 2));

        // ---- root: prediction -> softmax -> expand -> noise (mcts.py:126-136) ---------------
        if (active) {
            for (int j = g; j < H; j += GS) {
                const float v = p.root_hidden[tt * H + j];
                x0[(size_t)j * TW + tw] = v;
                hidden[j] = v;
            }
        }
        __syncwarp();
        warp_heads<TW>(wh_s, bh_s, x0, H, HP, A_net, lg_w, lane);
        if (active && p.noise_mode == 2 && g == 0)
            dirichlet_row(rprior, A, p.alpha, p.seed, p.tree_base + (unsigned long long)t);
        __syncwarp();
        {
            const float ssum = softmax_prepare<GS>(lg, eb, A_net, g, active);
            if (active) {
                for (int a = g; a < A; a += GS) {
                    const float p32 = prior_div(eb[a], ssum);
                    double pd = (double)p32;
                    if (p.noise_mode == 1)
                        pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(p.noise[tt * A + a], p.frac));
                    else if (p.noise_mode == 2)
                        pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(rprior[a], p.frac));
                    rprior[a] = pd;
                    store_node(nodes + 1 + a, 0.0, 0, p32);
                    eidx[1 + a] = -1;
                }
                if (g == 0) {
                    store_node(nodes, 0.0, 0, 1.0f);
                    eidx[0] = 0;
                    exp_node[0] = 0;
                }
            }
        }
        __syncwarp();

        // ---- S simulations (mcts.py:139-140, 157-203) ---------------------------------------
        long long depth = 0;
        for (int s = 0; s < p.S; ++s) {
            int plen, parent_e, action;
            const int leaf = select_leaf_tab<GS>(nodes, eidx, rprior, sq_s, rc_s, A, p.c_puct, 2 * s, path, g, active, plen,
                                                 parent_e, action);
            __syncwarp();
            const int k = s + 1;  // expansion index of the leaf
            // x0[j][t'] = hidden_{t'}[parent_e][j] + Emb[action_{t'}][j] for the warp's TW trees
            if (own) {
                float hv[TW][R];
#pragma unroll
                for (int u = 0; u < TW; ++u) {
                    const int pe = __shfl_sync(0xffffffffu, parent_e, u * GS);
                    const int ac = __shfl_sync(0xffffffffu, action, u * GS);
                    const size_t tu = (tbase + u < p.tv.B) ? (size_t)(tbase + u) : 0;
                    float ev[R];
                    load_vec<R>(p.tv.hidden + tu * hstride + (size_t)pe * H + j0, hv[u]);
                    load_vec<R>(emb_s + (size_t)ac * H + j0, ev);
#pragma unroll
                    for (int v = 0; v < R; ++v) hv[u][v] = fadd(hv[u][v], ev[v]);
                }
#pragma unroll
                for (int v = 0; v < R; ++v) {
                    float y[TW];
#pragma unroll
                    for (int u = 0; u < TW; ++u) y[u] = hv[u][v];
                    store_vec<TW>(x0 + (size_t)(j0 + v) * TW, y);
                }
            } else {
#pragma unroll
                for (int u = 0; u < TW; ++u) {  // keep the shuffles warp-uniform
                    (void)__shfl_sync(0xffffffffu, parent_e, u * GS);
                    (void)__shfl_sync(0xffffffffu, action, u * GS);
                }
            }
            __syncwarp();
            float* cur = x0;
            float* nxt = x1;
            for (int l = 0; l < L; ++l) {
                warp_layer<TW, R>(w_s + p.pk.layer_w(l), w_s + p.pk.layer_b(l), cur, nxt, H, lane);
                __syncwarp();
                float* tmp = cur; cur = nxt; nxt = tmp;
            }
            if (own) {
                float hv[R][TW];
#pragma unroll
                for (int v = 0; v < R; ++v) load_vec<TW>(cur + (size_t)(j0 + v) * TW, hv[v]);
#pragma unroll
                for (int u = 0; u < TW; ++u) {
                    if (tbase + u < p.tv.B) {
                        float y[R];
#pragma unroll
                        for (int v = 0; v < R; ++v) y[v] = hv[v][u];
                        store_vec<R>(p.tv.hidden + (size_t)(tbase + u) * hstride + (size_t)k * H + j0, y);
                    }
                }
            }
            warp_heads<TW>(wh_s, bh_s, cur, H, HP, A_net + 1, lg_w, lane);
            __syncwarp();
            const float ssum = softmax_prepare<GS>(lg, eb, A_net, g, active);
            if (active) {
                const float value = lg[A_net];
                const int cbase = 1 + k * A;
                for (int a = g; a < A; a += GS) {
                    store_node(nodes + cbase + a, 0.0, 0, prior_div(eb[a], ssum));
                    eidx[cbase + a] = -1;
                }
                if (g == 0) { eidx[leaf] = k; exp_node[k] = leaf; }
                backup_path<GS>(nodes, path, plen, value, g);
                depth += plen;
            }
            __syncwarp();
        }

        // ---- visit counts -> probabilities (mcts.py:143-155) --------------------------------
        if (active) {
            long long total = 0;
            for (int a = 0; a < A; ++a) total += load_node(nodes + 1 + a).visit;
            for (int a = g; a < p.A_cfg; a += GS) {
                const int c = a < A ? load_node(nodes + 1 + a).visit : 0;
                p.out_visits[tt * p.A_cfg + a] = c;
                p.out_probs[tt * p.A_cfg + a] =
                    total == 0 ? __ddiv_rn(1.0, (double)p.A_cfg) : __ddiv_rn((double)c, (double)total);
            }
            if (g == 0) {
                const Node r = load_node(nodes);
                if (p.out_root_value) p.out_root_value[t] = r.visit == 0 ? 0.0 : __ddiv_rn(r.value_sum, (double)r.visit);
                if (p.out_actions) {   // first maximum of the visit counts = np.argmax of the probabilities
                    int best = 0, bc = load_node(nodes + 1).visit;
                    for (int a = 1; a < A; ++a) { const int c = load_node(nodes + 1 + a).visit; if (c > bc) { bc = c; best = a; } }
                    p.out_actions[t] = (total == 0 || bc == 0) ? 0 : best;
                }
                p.tv.n_expanded[t] = p.S + 1;
                p.tv.depth_sum[t] = depth;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch
// ------------------------------------------------------------------------------------------------
template <int GS, int R>
static cudaError_t launch_one(const FusedParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(search_fused_kernel<GS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel_pdl(search_fused_kernel<GS, R>, grid, block, smem, st, p.pdl != 0, p);
}

template <int GS>
static cudaError_t launch_g(int R, const FusedParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    switch (R) {
        case 1: return launch_one<GS, 1>(p, grid, block, smem, st);
        case 2: return launch_one<GS, 2>(p, grid, block, smem, st);
        case 4: return launch_one<GS, 4>(p, grid, block, smem, st);
        case 8: return launch_one<GS, 8>(p, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_search_fused(int GS, int R, const FusedParams& p, int grid, int block, size_t smem, cudaStream_t st) {
    switch (GS) {
        case 4: return launch_g<4>(R, p, grid, block, smem, st);
        case 8: return launch_g<8>(R, p, grid, block, smem, st);
        case 16: return launch_g<16>(R, p, grid, block, smem, st);
        case 32: return launch_g<32>(R, p, grid, block, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cm
