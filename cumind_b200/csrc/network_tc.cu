// network_tc.cu — dynamics + prediction networks on the 5th-generation tensor cores (bf16 mode).
//
//   DynamicsNetwork.__call__    network.py:217-236   x = h + Emb[a];  L x { x = relu(x @ K_l + b_l) + x }
//   PredictionNetwork.__call__  network.py:254-266   logits = x @ Kp + bp, value = x @ Kv + bv (+ reward head)
// for a tile of 128 rows (trees) per CTA:
//   * operands bf16, accumulation fp32 in TMEM, bias / ReLU / residual in fp32 registers;
//   * the weights (pre-tiled on the host into the UMMA K-major core-matrix layout) are staged once per
//     CTA with TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier;
//   * each layer is H/16 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N=H, K=16) issued by one
//     thread, completion through `tcgen05.commit` -> mbarrier; the epilogue reads the accumulators
//     with `tcgen05.ld.32x32b` (thread i of warp w owns TMEM lane 32w+i = row 32w+i of the tile),
//     applies bias/ReLU/residual and writes the next bf16 A tile;
//   * the heads are one more MMA with N = 16 (policy logits | value | reward, zero padded).
// Two front ends share the kernel: the batched `cumind_recurrent_inference` / `cumind_prediction`
// entry points, and the expansion step of the stepwise search (gathers the parent hidden state of
// every tree's selected leaf, then writes children priors / hidden / leaf value) — mcts.py:181-196.
// Judged on rtol 2e-2 against the fp32-exact path (bf16 rounding of activations and weights).
#include <cuda_bf16.h>

#include "kernels.h"
#include "search_core.cuh"
#include "tc_common.cuh"

namespace cm {

struct TcParams {
    const unsigned char* pack;  // TcPackLayout bytes
    TcPackLayout pk;
    int n_rows;
    int with_dynamics;
    // batched front end
    const float* hidden;    // [n][H]
    const int32_t* action;  // [n]
    float* next;            // [n][H] or null
    float* reward;          // [n] or null
    float* logits;          // [n][A] or null
    float* value;           // [n] or null
    // search front end (expand): used when `tree` != 0
    int tree;
    TreeView tv;
    float* leaf_value;
};

template <int H>
__global__ void __launch_bounds__(128, 1) mlp_tc_kernel(const TcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TcPackLayout& pk = p.pk;
    unsigned char* w_s = smem;                                     // the whole pack
    unsigned char* a_s = smem + pk.total_bytes;                     // A tile: 128 x H bf16
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(a_s + 128 * H * 2);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    constexpr int NP = 16;
    constexpr uint32_t TMEM_COLS = (H + NP <= 32) ? 32 : (H + NP <= 64) ? 64 : (H + NP <= 128) ? 128 : 256;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = pk.L, A = pk.A;

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
    tc_mbar_wait(bar_w, 0);

    const float* bias_s = reinterpret_cast<const float*>(w_s + pk.bias_off);   // [L][H]
    const float* bh_s = reinterpret_cast<const float*>(w_s + pk.bh_off);       // [NP]
    const float* emb_s = reinterpret_cast<const float*>(w_s + pk.emb_off);     // [A][H]
    const uint32_t a_addr = smem_addr(a_s);
    const uint32_t idesc_layer = umma_idesc(128, H);
    const uint32_t idesc_heads = umma_idesc(128, NP);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t mma_phase = 0;

    for (int row0 = blockIdx.x * 128; row0 < p.n_rows; row0 += gridDim.x * 128) {
        const int r = row0 + tid;
        const bool live = r < p.n_rows;
        // ---- load the row: x = h (+ Emb[a]) ------------------------------------------------------
        float x[H];
        int leaf = 0, kexp = 0;
        const float* hsrc = nullptr;
        int act = 0;
        if (live) {
            if (p.tree) {
                leaf = p.tv.leaf[r];
                kexp = p.tv.n_expanded[r];
                const int pe = (leaf - 1) / p.tv.A;
                act = (leaf - 1) - pe * p.tv.A;
                hsrc = p.tv.hidden + ((size_t)r * (p.tv.S_max + 1) + pe) * H;
            } else {
                hsrc = p.hidden + (size_t)r * H;
                if (p.with_dynamics) {
                    act = p.action[r];
                    act = act < 0 ? 0 : (act >= A ? A - 1 : act);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) v = *reinterpret_cast<const float4*>(hsrc + j);
            if (p.with_dynamics) {
                const float4 e = *reinterpret_cast<const float4*>(emb_s + (size_t)act * H + j);
                v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
            }
            x[j] = v.x; x[j + 1] = v.y; x[j + 2] = v.z; x[j + 3] = v.w;
        }
        const int n_layers = p.with_dynamics ? L : 0;
        for (int l = 0; l <= n_layers; ++l) {
            // ---- A tile = bf16(x), row `tid` --------------------------------------------------------
#pragma unroll
            for (int kc = 0; kc < H / 8; ++kc) {
                uint32_t w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(x[kc * 8 + 2 * i], x[kc * 8 + 2 * i + 1]);
                    w4[i] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(a_s + ((size_t)(kc * 16 + (tid >> 3)) * 64 + (tid & 7) * 8) * 2) =
                    make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            const bool heads = (l == n_layers);
            if (tid == 0) {
                tc_fence_after();
                const int N = heads ? NP : H;
                const uint32_t b_addr = smem_addr(w_s + (heads ? pk.heads_off : pk.layer_off + (size_t)l * H * H * 2));
                const uint32_t a_lbo = 16 * 128, b_lbo = (uint32_t)(N / 8) * 128;
#pragma unroll
                for (int ks = 0; ks < H / 16; ++ks) {
                    const uint64_t ad = umma_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128);
                    const uint64_t bd = umma_desc(b_addr + ks * 2 * b_lbo, b_lbo, 128);
                    umma_f16(tmem_base + (heads ? H : 0), ad, bd, heads ? idesc_heads : idesc_layer, ks > 0 ? 1u : 0u);
                }
                umma_commit(bar_mma);
            }
            tc_mbar_wait(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
            if (!heads) {
                // ---- epilogue: x = relu(acc + b_l) + x ----------------------------------------------
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
            } else {
                float acc[16];
                tmem_ld16(lane_taddr + H, acc);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] += bh_s[i];
                // value / reward sit at columns A and A+1 (register arrays need compile-time indices)
                float val = 0.f, rew = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) { if (i == A) val = acc[i]; if (i == A + 1) rew = acc[i]; }
                if (live) {
                    if (!p.tree) {
                        if (p.logits) {
#pragma unroll
                            for (int i = 0; i < 14; ++i) if (i < A) p.logits[(size_t)r * A + i] = acc[i];
                        }
                        if (p.value) p.value[r] = val;
                        if (p.reward) p.reward[r] = rew;
                    } else {
                        // softmax + Node.expand + bookkeeping of the stepwise search (mcts.py:189-196)
                        float m = acc[0];
#pragma unroll
                        for (int i = 1; i < 14; ++i) if (i < A && acc[i] > m) m = acc[i];
                        float e[14], s = 0.f;
#pragma unroll
                        for (int i = 0; i < 14; ++i) {
                            e[i] = 0.f;
                            if (i < A) { e[i] = __double2float_rn(pexp((double)__fsub_rn(acc[i], m))); s = fadd(s, e[i]); }
                        }
                        Node* nodes = p.tv.nodes + (size_t)r * p.tv.N;
                        int32_t* eidx = p.tv.child_eidx + (size_t)r * p.tv.N;
                        const int cbase = 1 + kexp * p.tv.A;
#pragma unroll
                        for (int i = 0; i < 14; ++i) {
                            if (i < p.tv.A) {
                                store_node(nodes + cbase + i, 0.0, 0, __fdiv_rn(e[i], s));
                                eidx[cbase + i] = -1;
                            }
                        }
                        eidx[leaf] = kexp;
                        p.tv.exp_node[(size_t)r * (p.tv.S_max + 1) + kexp] = leaf;
                        p.tv.n_expanded[r] = kexp + 1;
                        p.leaf_value[r] = val;
                    }
                }
            }
            // the dynamics output is final after the last layer: store it before the heads
            if (!heads && l + 1 == n_layers && live) {
                float* dst = p.tree ? (p.tv.hidden + ((size_t)r * (p.tv.S_max + 1) + kexp) * H) : (p.next ? p.next + (size_t)r * H : nullptr);
                if (dst) {
#pragma unroll
                    for (int j = 0; j < H; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
                }
            }
            // TMEM and the A tile are reused by the next MMA: order the reads before it
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int H>
static cudaError_t launch_tc_h(const TcParams& p, int sm_count, cudaStream_t st) {
    const size_t smem = p.pk.total_bytes + (size_t)128 * H * 2 + 64;
    cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int tiles = (p.n_rows + 127) / 128;
    int grid = tiles < sm_count ? tiles : sm_count;
    mlp_tc_kernel<H><<<grid, 128, smem, st>>>(p);
    return cudaGetLastError();
}

size_t tc_tile_offset_host(int r, int k, int rows) { return tc_tile_offset(r, k, rows); }

bool tc_supported(int H, int L, int A) { return (H == 16 || H == 32 || H == 64 || H == 128) && A + 2 <= 16 && L >= 1 && L <= 8; }

static cudaError_t launch_tc(const TcParams& p, int sm_count, cudaStream_t st) {
    switch (p.pk.H) {
        case 16: return launch_tc_h<16>(p, sm_count, st);
        case 32: return launch_tc_h<32>(p, sm_count, st);
        case 64: return launch_tc_h<64>(p, sm_count, st);
        case 128: return launch_tc_h<128>(p, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_recurrent_tc(const unsigned char* pack, const TcPackLayout& pk, int sm_count, const float* hidden,
                                const int32_t* action, int B, float* next, float* reward, float* logits, float* value,
                                int with_dynamics, cudaStream_t st) {
    TcParams p{};
    p.pack = pack; p.pk = pk; p.n_rows = B; p.with_dynamics = with_dynamics;
    p.hidden = hidden; p.action = action; p.next = next; p.reward = reward; p.logits = logits; p.value = value;
    p.tree = 0;
    return launch_tc(p, sm_count, st);
}

cudaError_t launch_expand_tc(const unsigned char* pack, const TcPackLayout& pk, int sm_count, const TreeView& tv, float* leaf_value,
                             cudaStream_t st) {
    TcParams p{};
    p.pack = pack; p.pk = pk; p.n_rows = tv.B; p.with_dynamics = 1;
    p.tree = 1; p.tv = tv; p.leaf_value = leaf_value;
    return launch_tc(p, sm_count, st);
}

}  // namespace cm
