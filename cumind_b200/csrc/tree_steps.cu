// tree_steps.cu — the search as one launch per phase (warp per tree, any hidden_dim / action count).
//
// These are the fine-grained C-ABI entry points (cumind_tree_init_root / _select / _expand /
// _backup / _finalize / _dirichlet): the tests cross-check the fused kernel against them, ncu
// attributes time per phase with them, and cumind_search falls back to them when the parameters
// do not fit the fused kernel's shared-memory staging.  Reference anchors are in search_core.cuh.
#include "kernels.h"
#include "search_core.cuh"

namespace cm {

static constexpr int kWarpsPerBlock = 4;

CM_DEVICE int warp_tree_index() { return blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); }

// generic-H layer: lane computes outputs j = lane, lane+32, ... (sequential-k FMA chain)
CM_DEVICE void dyn_layer_any(const float* __restrict__ K, const float* __restrict__ b, const float* __restrict__ xin,
                             float* __restrict__ xout, int H, int lane) {
    for (int j = lane; j < H; j += 32) {
        float acc = 0.0f;
#pragma unroll 8
        for (int k = 0; k < H; ++k) acc = ffma(xin[k], K[(size_t)k * H + j], acc);
        xout[j] = fadd(frelu(fadd(acc, b[j])), xin[j]);
    }
}

// ---- root: prediction -> softmax -> expand -> noise (mcts.py:126-136) ----------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) init_root_kernel(const StepParams p, const float* __restrict__ root_hidden,
                                                                       const double* __restrict__ noise) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.pk.H, HP = p.pk.HP, A = p.tv.A, A_net = p.A_net;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* x = sm + (size_t)w * (H + 2 * HP);
    float* lg = x + H;
    float* eb = lg + HP;
    const int t = warp_tree_index();
    const bool active = t < p.tv.B;
    const size_t tt = active ? (size_t)t : 0;
    Node* nodes = p.tv.nodes + tt * p.tv.N;
    int32_t* eidx = p.tv.child_eidx + tt * p.tv.N;
    double* rprior = p.tv.root_prior + tt * A;
    float* hidden = p.tv.hidden + tt * (size_t)(p.tv.S_max + 1) * H;
    if (active)
        for (int j = lane; j < H; j += 32) { const float v = root_hidden[tt * H + j]; x[j] = v; hidden[j] = v; }
    __syncwarp();
    if (active) heads<32>(p.pack + p.pk.heads_w, p.pack + p.pk.heads_b, x, H, HP, A_net, lg, lane);
    if (active && p.noise_mode == 2 && lane == 0) dirichlet_row(rprior, A, p.alpha, p.seed, p.tree_base + (unsigned long long)t);
    __syncwarp();
    const float ssum = softmax_prepare<32>(lg, eb, A_net, lane, active);
    if (active) {
        for (int a = lane; a < A; a += 32) {
            const float p32 = __fdiv_rn(eb[a], ssum);
            double pd = (double)p32;
            if (p.noise_mode == 1) pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(noise[tt * A + a], p.frac));
            else if (p.noise_mode == 2) pd = __dadd_rn(__dmul_rn(pd, p.one_minus_frac), __dmul_rn(rprior[a], p.frac));
            rprior[a] = pd;
            store_node(nodes + 1 + a, 0.0, 0, p32);
            eidx[1 + a] = -1;
            if (p.tree_mode == CUMIND_TREE_MINMAX_DISCOUNT) p.tv.reward[tt * p.tv.N + 1 + a] = 0.0f;
        }
        if (lane == 0) {
            store_node(nodes, 0.0, 0, 1.0f);
            eidx[0] = 0;
            p.tv.exp_node[tt * (p.tv.S_max + 1)] = 0;
            p.tv.n_expanded[t] = 1;
            p.tv.depth_sum[t] = 0;
            p.tv.path_len[t] = 0;
            p.tv.leaf[t] = 0;
            if (p.tree_mode == CUMIND_TREE_MINMAX_DISCOUNT) {
                p.tv.reward[tt * p.tv.N] = 0.0f;
                p.tv.minmax[2 * tt] = dinf();
                p.tv.minmax[2 * tt + 1] = -dinf();
            }
        }
    }
}

// ---- selection (mcts.py:165-171) ----------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) select_kernel(const StepParams p, int n_sqrt) {
    extern __shared__ __align__(16) double sq_s[];
    for (int i = threadIdx.x; i < n_sqrt; i += blockDim.x) sq_s[i] = __dsqrt_rn((double)i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int t = warp_tree_index();
    const bool active = t < p.tv.B;
    const size_t tt = active ? (size_t)t : 0;
    const Node* nodes = p.tv.nodes + tt * p.tv.N;
    const int32_t* eidx = p.tv.child_eidx + tt * p.tv.N;
    int* path = p.tv.path + tt * (p.tv.S_max + 2);
    const int root_visits = load_node(nodes).visit;
    int plen, parent_e, action;
    const int leaf = select_leaf<32>(nodes, eidx, p.tv.root_prior + tt * p.tv.A, sq_s, p.tv.A, p.c_puct, root_visits, path,
                                     lane, active, plen, parent_e, action);
    if (active && lane == 0) {
        p.tv.path_len[t] = plen;
        p.tv.leaf[t] = leaf;
        p.tv.depth_sum[t] += plen;
    }
}

// ---- expansion + evaluation (mcts.py:181-196) ---------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) expand_kernel(const StepParams p, float* __restrict__ leaf_value) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.pk.H, L = p.pk.L, HP = p.pk.HP, A = p.tv.A, A_net = p.A_net;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* x0 = sm + (size_t)w * (2 * H + 2 * HP);
    float* x1 = x0 + H;
    float* lg = x1 + H;
    float* eb = lg + HP;
    const int t = warp_tree_index();
    const bool active = t < p.tv.B;
    const size_t tt = active ? (size_t)t : 0;
    Node* nodes = p.tv.nodes + tt * p.tv.N;
    int32_t* eidx = p.tv.child_eidx + tt * p.tv.N;
    float* hidden = p.tv.hidden + tt * (size_t)(p.tv.S_max + 1) * H;
    const int leaf = p.tv.leaf[tt];
    const int k = p.tv.n_expanded[tt];
    const int parent_e = (leaf - 1) / A, action = (leaf - 1) % A;
    if (active) {
        const float* h = hidden + (size_t)parent_e * H;
        const float* emb = p.pack + p.pk.emb + (size_t)action * H;
        for (int j = lane; j < H; j += 32) x0[j] = fadd(h[j], emb[j]);
    }
    __syncwarp();
    float* cur = x0;
    float* nxt = x1;
    for (int l = 0; l < L; ++l) {
        if (active) dyn_layer_any(p.pack + p.pk.layer_k(l), p.pack + p.pk.layer_b(l), cur, nxt, H, lane);
        __syncwarp();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (active) {
        for (int j = lane; j < H; j += 32) hidden[(size_t)k * H + j] = cur[j];
        // policy logits, value and (min-max mode) the dynamics reward: head columns 0..A_net-1, A_net, A_net+1
        heads<32>(p.pack + p.pk.heads_w, p.pack + p.pk.heads_b, cur, H, HP, A_net + (p.tree_mode == CUMIND_TREE_MINMAX_DISCOUNT ? 2 : 1), lg, lane);
    }
    __syncwarp();
    const float ssum = softmax_prepare<32>(lg, eb, A_net, lane, active);
    if (active) {
        const int cbase = 1 + k * A;
        for (int a = lane; a < A; a += 32) {
            store_node(nodes + cbase + a, 0.0, 0, __fdiv_rn(eb[a], ssum));
            eidx[cbase + a] = -1;
            if (p.tree_mode == CUMIND_TREE_MINMAX_DISCOUNT) p.tv.reward[tt * p.tv.N + cbase + a] = 0.0f;
        }
        if (lane == 0) {
            eidx[leaf] = k;
            p.tv.exp_node[tt * (p.tv.S_max + 1) + k] = leaf;
            p.tv.n_expanded[t] = k + 1;
            leaf_value[t] = lg[A_net];
            if (p.tree_mode == CUMIND_TREE_MINMAX_DISCOUNT) p.tv.reward[tt * p.tv.N + leaf] = lg[A_net + 1];
        }
    }
}

// ---- CUMIND_TREE_MINMAX_DISCOUNT: the optional textbook rules (oracle/cumind_oracle.c cmo_search_one_mm) --------------
// score = vs + ((c_puct*prior)*sqrt(N_parent))/(1+n), vs = 0 if n == 0 else normalise(reward + discount*(W/n)); IEEE
// float64 operations in that order.  Returned as an order-preserving integer key with the rules of Python's max()
// (-0 == +0; a NaN wins only as the first candidate).
CM_DEVICE long long mm_score_key(const Node nd, double prior, double reward, double sqrtN, double c_puct, double discount, double mn,
                                 double mx, bool first) {
    const long long PINF = 0x7ff0000000000000LL, NINF = (long long)0xfff0000000000000ULL;
    const double explo = __ddiv_rn(__dmul_rn(__dmul_rn(c_puct, prior), sqrtN), (double)(1 + nd.visit));
    double vs = 0.0;
    if (nd.visit != 0) {
        const double x = __dadd_rn(reward, __dmul_rn(discount, __ddiv_rn(nd.value_sum, (double)nd.visit)));
        vs = mx > mn ? __ddiv_rn(__dsub_rn(x, mn), __dsub_rn(mx, mn)) : x;
    }
    long long b = __double_as_longlong(__dadd_rn(vs, explo));
    const long long mag = b & 0x7fffffffffffffffLL;
    b = mag > PINF ? (first ? PINF : NINF) : (mag == 0 ? 0LL : b);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

// selection: from the root down while the node is expanded; lanes share the children of the current node, the first
// maximum in action order wins (warp arg-max on (key, lower action))
__global__ void __launch_bounds__(kWarpsPerBlock * 32) select_mm_kernel(const StepParams p) {
    const int lane = threadIdx.x & 31;
    const int t = warp_tree_index();
    if (t >= p.tv.B) return;
    const size_t tt = (size_t)t;
    const int A = p.tv.A;
    const Node* nodes = p.tv.nodes + tt * p.tv.N;
    const int32_t* eidx = p.tv.child_eidx + tt * p.tv.N;
    const float* reward = p.tv.reward + tt * p.tv.N;
    const double* rprior = p.tv.root_prior + tt * A;
    const double mn = p.tv.minmax[2 * tt], mx = p.tv.minmax[2 * tt + 1];
    int* path = p.tv.path + tt * (p.tv.S_max + 2);
    int node = 0, plen = 0;
    int e = eidx[0];
    while (e >= 0) {
        const int base = 1 + e * A;
        const double sqrtN = __dsqrt_rn((double)load_node(nodes + node).visit);
        long long best = 0;
        int best_a = 0x7fffffff;
        for (int a = lane; a < A; a += 32) {
            const Node nd = load_node(nodes + base + a);
            const double pr = node == 0 ? rprior[a] : (double)nd.prior;
            const long long key = mm_score_key(nd, pr, (double)reward[base + a], sqrtN, p.c_puct, p.discount, mn, mx, a == 0);
            if (best_a == 0x7fffffff || key > best) { best = key; best_a = a; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long ok = __shfl_xor_sync(0xffffffffu, best, off);
            const int oa = __shfl_xor_sync(0xffffffffu, best_a, off);
            if (oa != 0x7fffffff && (best_a == 0x7fffffff || ok > best || (ok == best && oa < best_a))) { best = ok; best_a = oa; }
        }
        if (lane == 0) path[plen] = node;
        ++plen;
        node = base + best_a;
        e = eidx[node];
    }
    if (lane == 0) {
        p.tv.path_len[t] = plen;
        p.tv.leaf[t] = node;
        p.tv.depth_sum[t] += plen;
    }
}

// backup: the leaf first (it is counted), then its ancestors up to the root; G <- reward + discount*G on the way
__global__ void backup_mm_kernel(const StepParams p, const float* __restrict__ leaf_value) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.tv.B) return;
    const size_t tt = (size_t)t;
    Node* nodes = p.tv.nodes + tt * p.tv.N;
    const float* reward = p.tv.reward + tt * p.tv.N;
    const int* path = p.tv.path + tt * (p.tv.S_max + 2);
    const int plen = p.tv.path_len[t], leaf = p.tv.leaf[t];
    double mn = p.tv.minmax[2 * tt], mx = p.tv.minmax[2 * tt + 1];
    double G = (double)leaf_value[t];
    for (int i = plen; i >= 0; --i) {
        const int nd = i == plen ? leaf : path[i];
        Node r = load_node(nodes + nd);
        r.value_sum = __dadd_rn(r.value_sum, G);
        r.visit += 1;
        store_node(nodes + nd, r.value_sum, r.visit, r.prior);
        const double rew = (double)reward[nd];
        const double q = __dadd_rn(rew, __dmul_rn(p.discount, __ddiv_rn(r.value_sum, (double)r.visit)));
        if (q < mn) mn = q;
        if (q > mx) mx = q;
        G = __dadd_rn(rew, __dmul_rn(p.discount, G));
    }
    p.tv.minmax[2 * tt] = mn;
    p.tv.minmax[2 * tt + 1] = mx;
}

// ---- backup (mcts.py:199-203) -------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) backup_kernel(const StepParams p, const float* __restrict__ leaf_value) {
    const int lane = threadIdx.x & 31;
    const int t = warp_tree_index();
    if (t >= p.tv.B) return;
    backup_path<32>(p.tv.nodes + (size_t)t * p.tv.N, p.tv.path + (size_t)t * (p.tv.S_max + 2), p.tv.path_len[t], leaf_value[t], lane);
}

// ---- visit counts -> probabilities (mcts.py:143-155) ---------------------------------------------
__global__ void finalize_kernel(const StepParams p, int32_t* __restrict__ out_visits, double* __restrict__ out_probs,
                                double* __restrict__ out_root_value) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.tv.B) return;
    const Node* nodes = p.tv.nodes + (size_t)t * p.tv.N;
    const int A = p.tv.A;
    long long total = 0;
    for (int a = 0; a < A; ++a) total += load_node(nodes + 1 + a).visit;
    for (int a = 0; a < p.A_cfg; ++a) {
        const int c = a < A ? load_node(nodes + 1 + a).visit : 0;
        if (out_visits) out_visits[(size_t)t * p.A_cfg + a] = c;
        if (out_probs)
            out_probs[(size_t)t * p.A_cfg + a] = total == 0 ? __ddiv_rn(1.0, (double)p.A_cfg) : __ddiv_rn((double)c, (double)total);
    }
    if (out_root_value) {
        const Node r = load_node(nodes);
        out_root_value[t] = r.visit == 0 ? 0.0 : __ddiv_rn(r.value_sum, (double)r.visit);
    }
}

__global__ void dirichlet_kernel(double* __restrict__ noise, int B, int A, double alpha, unsigned long long seed,
                                 unsigned long long tree_base) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    dirichlet_row(noise + (size_t)t * A, A, alpha, seed, tree_base + (unsigned long long)t);
}

// int(np.argmax(action_probs)) (agent.py:86): first maximum
__global__ void argmax_first_kernel(const double* __restrict__ probs, int B, int A, int32_t* __restrict__ actions) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    int best = 0;
    double bv = probs[(size_t)t * A];
    for (int a = 1; a < A; ++a) { const double v = probs[(size_t)t * A + a]; if (v > bv) { bv = v; best = a; } }
    actions[t] = best;
}

// Host-buffer path of cumind_select_actions_host when the caller's buffers are pinned (mapped) host memory:
// the GPU reads the inputs and writes the results over PCIe itself, one kernel each way instead of two
// cudaMemcpyAsync each way (a copy-engine transfer costs ~6 us of set-up for these 16-64 kB buffers).
__global__ void stage_in_kernel(const float4* __restrict__ h_a, float4* __restrict__ d_a, size_t n_a, const float4* __restrict__ h_b,
                                float4* __restrict__ d_b, size_t n_b) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_a) d_a[i] = h_a[i];
    else if (i - n_a < n_b) d_b[i - n_a] = h_b[i - n_a];
}
// int(np.argmax(action_probs)) (agent.py:86) + the copy of actions and probabilities to the host buffers.
// Threads [0, B): one tree each (first maximum, 4-byte coalesced writes); threads [B, B + n_vec): one 16-byte
// vector of the probability matrix each, so the PCIe writes are full 128-byte lines per warp.
__global__ void stage_out_kernel(const double* __restrict__ probs, int B, int A, int32_t* __restrict__ d_actions,
                                 int32_t* __restrict__ h_actions, double* __restrict__ h_probs, size_t n_vec, int vec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)B) {
        const int t = (int)i;
        int best = 0;
        double bv = probs[(size_t)t * A];
        for (int a = 1; a < A; ++a) { const double v = probs[(size_t)t * A + a]; if (v > bv) { bv = v; best = a; } }
        d_actions[t] = best;
        h_actions[t] = best;
    } else if (i - B < n_vec) {
        const size_t j = i - B;
        if (vec) reinterpret_cast<double2*>(h_probs)[j] = reinterpret_cast<const double2*>(probs)[j];
        else h_probs[j] = probs[j];
    }
}

// uint8 observations (image frames) -> float32 / 255, correctly rounded: the same value as observation.astype(float32) / 255
__global__ void u8_to_f32_kernel(const uchar4* __restrict__ in, float4* __restrict__ out, size_t n4, const unsigned char* __restrict__ in1,
                                 float* __restrict__ out1, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uchar4 v = in[i];
        out[i] = make_float4(__fdiv_rn((float)v.x, 255.0f), __fdiv_rn((float)v.y, 255.0f), __fdiv_rn((float)v.z, 255.0f), __fdiv_rn((float)v.w, 255.0f));
    }
    for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out1[i] = __fdiv_rn((float)in1[i], 255.0f);
}
cudaError_t launch_u8_to_f32(const unsigned char* in, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t n4 = (((uintptr_t)in & 3) == 0) ? n / 4 : 0;
    size_t blocks = (n / 4 + 255) / 256 + 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    u8_to_f32_kernel<<<(int)blocks, 256, 0, st>>>((const uchar4*)in, (float4*)out, n4, in, out, n);
    return cudaGetLastError();
}

// Self-test of div_small (the table-driven division of the fused select): random and structured
// numerators against the IEEE division, bitwise.  counts[0] = mismatches, counts[1] = cases.
__global__ void selftest_div_kernel(unsigned long long n_per_thread, unsigned long long seed, int max_d,
                                    unsigned long long* __restrict__ counts) {
    Philox rng;
    rng.init(seed, (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad = 0, total = 0;
    for (unsigned long long i = 0; i < n_per_thread; ++i) {
        const unsigned r0 = rng.next(), r1 = rng.next(), r2 = rng.next(), r3 = rng.next();
        const int d = 1 + (int)(r0 % (unsigned)max_d);
        const double dd = (double)d;
        const double recip = __ddiv_rn(1.0, dd);
        double a;
        const unsigned mode = r3 & 7u;
        const unsigned long long mant = (((unsigned long long)r1 << 32) | r2) & 0x000FFFFFFFFFFFFFULL;
        if (mode < 3) {            // random significand, moderate exponent (the values a search produces)
            const long long ex = 1023 + (long long)((r3 >> 3) % 41) - 20;
            a = __longlong_as_double(((unsigned long long)(r3 >> 31) << 63) | ((unsigned long long)ex << 52) | mant);
        } else if (mode < 5) {     // exact multiples of d and their neighbours (ties / near-ties of the quotient)
            const double q = __longlong_as_double((1023ULL << 52) | mant);
            const double m = __dmul_rn(q, dd);
            const long long step = (long long)((r3 >> 3) % 5) - 2;
            a = __longlong_as_double(__double_as_longlong(m) + step);
        } else if (mode < 7) {     // any exponent, including subnormals, huge values, inf and NaN
            a = __longlong_as_double(((unsigned long long)r1 << 32) | r2);
        } else {                   // small integers and zeros
            a = (double)((int)(r1 % 4001u) - 2000);
            if ((r3 >> 3) & 1) a = -0.0;
        }
        const double want = __ddiv_rn(a, dd);
        const double got = div_small(a, dd, recip);
        const bool same = (__double_as_longlong(want) == __double_as_longlong(got)) || (want != want && got != got);
        if (!same) ++bad;
        ++total;
    }
    atomicAdd(counts, bad);
    atomicAdd(counts + 1, total);
}

cudaError_t launch_selftest_div(unsigned long long n_per_thread, unsigned long long seed, int max_d, unsigned long long* counts,
                                cudaStream_t st) {
    selftest_div_kernel<<<148 * 4, 256, 0, st>>>(n_per_thread, seed, max_d, counts);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
static inline int warp_grid(int B) { return (B + kWarpsPerBlock - 1) / kWarpsPerBlock; }

cudaError_t launch_init_root(const StepParams& p, const float* root_hidden, const double* noise, cudaStream_t st) {
    const size_t smem = (size_t)kWarpsPerBlock * (p.pk.H + 2 * p.pk.HP) * sizeof(float);
    init_root_kernel<<<warp_grid(p.tv.B), kWarpsPerBlock * 32, smem, st>>>(p, root_hidden, noise);
    return cudaGetLastError();
}
cudaError_t launch_select(const StepParams& p, cudaStream_t st) {
    const int n_sqrt = 2 * p.tv.S_max + 2;
    select_kernel<<<warp_grid(p.tv.B), kWarpsPerBlock * 32, n_sqrt * sizeof(double), st>>>(p, n_sqrt);
    return cudaGetLastError();
}
cudaError_t launch_expand(const StepParams& p, float* leaf_value, cudaStream_t st) {
    const size_t smem = (size_t)kWarpsPerBlock * (2 * p.pk.H + 2 * p.pk.HP) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    expand_kernel<<<warp_grid(p.tv.B), kWarpsPerBlock * 32, smem, st>>>(p, leaf_value);
    return cudaGetLastError();
}
cudaError_t launch_backup(const StepParams& p, const float* leaf_value, cudaStream_t st) {
    backup_kernel<<<warp_grid(p.tv.B), kWarpsPerBlock * 32, 0, st>>>(p, leaf_value);
    return cudaGetLastError();
}
cudaError_t launch_select_mm(const StepParams& p, cudaStream_t st) {
    select_mm_kernel<<<warp_grid(p.tv.B), kWarpsPerBlock * 32, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_backup_mm(const StepParams& p, const float* leaf_value, cudaStream_t st) {
    backup_mm_kernel<<<(p.tv.B + 127) / 128, 128, 0, st>>>(p, leaf_value);
    return cudaGetLastError();
}
cudaError_t launch_finalize(const StepParams& p, int32_t* out_visits, double* out_probs, double* out_root_value, cudaStream_t st) {
    finalize_kernel<<<(p.tv.B + 127) / 128, 128, 0, st>>>(p, out_visits, out_probs, out_root_value);
    return cudaGetLastError();
}
cudaError_t launch_dirichlet(double* noise, int B, int A, double alpha, unsigned long long seed, unsigned long long tree_base,
                             cudaStream_t st) {
    dirichlet_kernel<<<(B + 127) / 128, 128, 0, st>>>(noise, B, A, alpha, seed, tree_base);
    return cudaGetLastError();
}
// n_a / n_b in bytes (multiples of 16, 16-byte aligned pointers); either may be 0
cudaError_t launch_stage_in(const void* h_a, void* d_a, size_t n_a, const void* h_b, void* d_b, size_t n_b, cudaStream_t st) {
    const size_t va = n_a / 16, vb = n_b / 16, n = va + vb;
    if (n == 0) return cudaSuccess;
    stage_in_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float4*)h_a, (float4*)d_a, va, (const float4*)h_b, (float4*)d_b, vb);
    return cudaGetLastError();
}
cudaError_t launch_stage_out(const double* probs, int B, int A, int32_t* d_actions, int32_t* h_actions, double* h_probs, cudaStream_t st) {
    const size_t n = (size_t)B * A;
    const int vec = (n % 2 == 0) && ((uintptr_t)probs % 16 == 0) && ((uintptr_t)h_probs % 16 == 0);
    const size_t n_vec = vec ? n / 2 : n, threads = (size_t)B + n_vec;
    stage_out_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(probs, B, A, d_actions, h_actions, h_probs, n_vec, vec);
    return cudaGetLastError();
}
cudaError_t launch_argmax_first(const double* probs, int B, int A, int32_t* actions, cudaStream_t st) {
    argmax_first_kernel<<<(B + 127) / 128, 128, 0, st>>>(probs, B, A, actions);
    return cudaGetLastError();
}

}  // namespace cm
