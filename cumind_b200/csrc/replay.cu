// replay.cu — the sum tree of TreeBuffer (src/cumind/data/memory.py:190-299) in device memory.
//
// The reference keeps `sum_tree` as a flat float64 array of 2*tree_size-1 nodes (root 0, children 2i+1 / 2i+2,
// leaves at tree_size-1 ...) and updates it one priority at a time: `change = p - tree[leaf]; tree[leaf] = p`,
// then `tree[ancestor] += change` for every ancestor up to the root, and — because `_propagate` recurses once
// more from the root, where `(0 - 1) // 2 == -1` — also `tree[-1] += change`, the last leaf (memory.py:212-217,
// 249-253).  The float64 additions of successive updates hit the same ancestors, so the order of the updates is
// part of the result; sumtree_update_kernel applies a batch IN ORDER with exactly those additions.  The ancestors
// of one update are independent of each other, so one warp handles an update with a lane per tree level; the whole
// tree is staged in shared memory when it fits (capacity up to ~14 000), else the warp works in global memory.
//
// Retrieval (memory.py:219-231, 265-273) is independent per sample: one thread per sample walks
// `s <= tree[left] ? left : (s -= tree[left], right)` from the root; `s = a + (b - a) * u` is `random.uniform`.
#include "kernels.h"

namespace cm {

// One update applied by one warp exactly as the reference does it (a lane per ancestor level): the order-preserving path
// for anything the batched scheme below does not cover.
__device__ void sumtree_update_one(double* tree, long long n_nodes, long long leaf, double p, int lane) {
    const double change = p - tree[leaf];   // every lane reads the same word
    __syncwarp();
    if (lane == 0) tree[leaf] = p;
    __syncwarp();
    // an update may name any node (the reference does not check); ancestors are ((leaf + 1) >> j) - 1 for j = 1 .. level(leaf)
    const long long l1 = leaf + 1;
    const int level = 63 - __clzll(l1);
    for (int j0 = 1; j0 <= level; j0 += 32) {
        const int j = j0 + lane;
        if (j <= level) tree[(l1 >> j) - 1] += change;
    }
    __syncwarp();
    if (lane == 0) tree[n_nodes - 1] += change;   // the reference's extra step from the root: sum_tree[-1]
    __syncwarp();
}

// A batch of updates, applied with the float64 additions of the sequential reference IN ITS ORDER PER NODE, in parallel
// over the nodes.  What one update does (memory.py:249-253, 212-217): change = p - tree[leaf]; tree[leaf] = p; every
// ancestor += change; tree[-1] += change.  For a chunk of updates that name only real leaves other than the last node:
//   * change_k depends only on the previous update of the SAME leaf (or the stored value): every thread finds it by a
//     scan of the chunk's leaf indices in shared memory;
//   * the value of a leaf after the chunk is the priority of its LAST update;
//   * an internal node receives the changes of the updates below it, in chunk order.  Nodes are independent of each
//     other, so warp j owns level j and lane (node & 31) owns the node: each lane walks the chunk in order and adds the
//     changes that are its own, accumulating runs on the same node in a register.  Warp 0 lane 0 owns tree[-1].
// The additions that reach one node are the reference's, in the reference's order: the tree is bit-identical.
// A chunk that names an internal node or the last node goes through sumtree_update_one, update by update.
constexpr int kSumtreeChunk = 1024;

template <bool SMEM>
__global__ void __launch_bounds__(1024, 1) sumtree_update_kernel(double* __restrict__ tree_g, long long n_nodes, long long tree_size,
                                                                    const long long* __restrict__ idx, const double* __restrict__ pri,
                                                                    long long n) {
    extern __shared__ double tree_s[];
    __shared__ __align__(16) int leaf_s[kSumtreeChunk];
    __shared__ double chg_s[kSumtreeChunk];
    __shared__ int irregular;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* tree = SMEM ? tree_s : tree_g;
    if (SMEM) {
        for (long long i = tid; i < n_nodes; i += 1024) tree_s[i] = tree_g[i];
    }
    int depth = 0;
    while (((long long)1 << depth) < tree_size) ++depth;
    for (long long c0 = 0; c0 < n; c0 += kSumtreeChunk) {
        const int m = (int)((n - c0 < kSumtreeChunk) ? n - c0 : kSumtreeChunk);
        if (tid == 0) irregular = 0;
        __syncthreads();
        long long my_leaf = -1;
        if (tid < m) {
            my_leaf = idx[c0 + tid];
            leaf_s[tid] = (int)(my_leaf - (tree_size - 1));   // leaf number; the regular range is [0, tree_size - 1)
            if (my_leaf < tree_size - 1 || my_leaf >= n_nodes - 1) irregular = 1;
        } else {
            leaf_s[tid] = -1 - tid;
        }
        __syncthreads();
        if (irregular) {   // block-uniform
            if (warp == 0) {
                for (int k = 0; k < m; ++k) {
                    sumtree_update_one(tree, n_nodes, idx[c0 + k], pri[c0 + k], lane);
                    if (!SMEM) __threadfence_block();
                }
            }
            __syncthreads();
            continue;
        }
        // 1. change of every update; is it the last one of its leaf?  (leaf_s is padded with -1 - j: never equal to a leaf)
        bool last = false;
        double p = 0.0;
        if (tid < m) {
            const int l = leaf_s[tid];
            int prev = -1;
            const int4* l4 = reinterpret_cast<const int4*>(leaf_s);
            const int q = tid >> 2;
            for (int j = 0; j < q; ++j) {   // whole groups before mine
                const int4 v = l4[j];
                prev = v.x == l ? 4 * j : prev;
                prev = v.y == l ? 4 * j + 1 : prev;
                prev = v.z == l ? 4 * j + 2 : prev;
                prev = v.w == l ? 4 * j + 3 : prev;
            }
            last = true;
            for (int j = 4 * q; j < min(4 * q + 4, m); ++j) {
                const bool same = leaf_s[j] == l;
                if (same && j < tid) prev = j;
                if (same && j > tid) last = false;
            }
            bool later = false;
            for (int j = q + 1; j < (m + 3) / 4; ++j) {   // whole groups after mine
                const int4 v = l4[j];
                later |= (v.x == l) | (v.y == l) | (v.z == l) | (v.w == l);
            }
            last = last && !later;
            p = pri[c0 + tid];
            chg_s[tid] = __dsub_rn(p, prev >= 0 ? pri[c0 + prev] : tree[my_leaf]);
        }
        __syncthreads();
        // 2. leaves
        if (last) tree[my_leaf] = p;
        // 3. ancestors: warp j owns level j (warp 0 lane 0: tree[-1])
        const long long base = tree_size;   // leaf number + tree_size = node index + 1
        if (warp == 0) {
            if (lane == 0) {
                double v = tree[n_nodes - 1];
                for (int k = 0; k < m; ++k) v = __dadd_rn(v, chg_s[k]);
                tree[n_nodes - 1] = v;
            }
        } else {
            for (int j = warp; j <= depth; j += 31) {
                const long long level_nodes = tree_size >> j;
                if (level_nodes <= 32) {
                    // few nodes: a lane per node, every lane walks the chunk and takes what is its own (predicated, in order)
                    const long long own = level_nodes - 1 + lane;
                    const bool valid = lane < level_nodes;
                    double v = valid ? tree[own] : 0.0;
                    for (int k = 0; k < m; ++k) {
                        const long long a = ((base + leaf_s[k]) >> j) - 1;
                        const double c = chg_s[k];
                        if (a == own) v = __dadd_rn(v, c);
                    }
                    if (valid) tree[own] = v;
                } else {
                    // many nodes: 32 updates at a time; lanes that name the same node form a group whose first lane adds the
                    // group's changes in lane (= chunk) order, all groups at once
                    for (int k0 = 0; k0 < m; k0 += 32) {
                        const int k = k0 + lane;
                        const bool valid = k < m;
                        const long long a = valid ? ((base + leaf_s[k]) >> j) - 1 : -1 - (long long)lane;
                        const double c = valid ? chg_s[k] : 0.0;
                        const unsigned mask = __match_any_sync(0xffffffffu, a);
                        const bool leader = valid && (__ffs(mask) - 1 == lane);
                        unsigned rem = leader ? mask : 0u;
                        double v = leader ? tree[a] : 0.0;
                        while (__any_sync(0xffffffffu, rem != 0u)) {
                            const int src = rem ? __ffs(rem) - 1 : lane;
                            const double cc = __shfl_sync(0xffffffffu, c, src);
                            if (rem) {
                                v = __dadd_rn(v, cc);
                                rem &= rem - 1;
                            }
                        }
                        if (leader) tree[a] = v;
                        __syncwarp();
                    }
                }
            }
        }
        __syncthreads();
    }
    if (SMEM) {
        for (long long i = tid; i < n_nodes; i += 1024) tree_g[i] = tree_s[i];
    }
}

__global__ void sumtree_sample_kernel(const double* __restrict__ tree, long long n_nodes, long long tree_size,
                                      const double* __restrict__ uniform, long long batch, long long* __restrict__ tree_idx,
                                      long long* __restrict__ data_idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const double segment = __ddiv_rn(tree[0], (double)batch);
    const double a = __dmul_rn(segment, (double)i), b = __dmul_rn(segment, (double)(i + 1));
    double s = __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), uniform[i]));
    long long node = 0;
    while (true) {
        const long long left = 2 * node + 1;
        if (left >= n_nodes) break;
        const double l = tree[left];
        if (s <= l) node = left;
        else { s = __dsub_rn(s, l); node = left + 1; }
    }
    tree_idx[i] = node;
    data_idx[i] = node - tree_size + 1;
}

cudaError_t launch_sumtree_update(double* tree, long long n_nodes, int depth, const long long* idx, const double* pri, long long n,
                                  cudaStream_t st) {
    const size_t bytes = (size_t)n_nodes * sizeof(double);
    const long long tree_size = (n_nodes + 1) / 2;
    (void)depth;
    if (bytes <= 200 * 1024) {
        auto k = sumtree_update_kernel<true>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        k<<<1, 1024, bytes, st>>>(tree, n_nodes, tree_size, idx, pri, n);
    } else {
        sumtree_update_kernel<false><<<1, 1024, 0, st>>>(tree, n_nodes, tree_size, idx, pri, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_sumtree_sample(const double* tree, long long n_nodes, long long tree_size, const double* uniform, long long batch,
                                  long long* tree_idx, long long* data_idx, cudaStream_t st) {
    sumtree_sample_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(tree, n_nodes, tree_size, uniform, batch, tree_idx, data_idx);
    return cudaGetLastError();
}

// ---- trajectory records in device memory ----------------------------------------------------------------------------
// What Trainer._prepare_batch (trainer.py:118-170) reads of a sampled trajectory is fixed when the trajectory is stored:
// the first observation and search policy, the first K actions and rewards (zero padded), the n-step discounted reward
// sum of its first td_steps rewards and, if the trajectory is longer than td_steps, the observation the bootstrap value
// is computed from.  One record holds exactly that (layout in include/cumind_b200.h); a batch is then one gather launch
// over the sampled indices instead of a host loop over Python lists and a host-to-device copy per step.
__global__ void __launch_bounds__(128) replay_gather_kernel(const unsigned char* __restrict__ ring, long long capacity, long long rec_bytes,
                                                            long long obs_size, int A, int K, const long long* __restrict__ index,
                                                            long long head, long long n_valid, float* __restrict__ obs,
                                                            float* __restrict__ boot_obs, float* __restrict__ policy,
                                                            float* __restrict__ rewards, int* __restrict__ actions,
                                                            double* __restrict__ partial, int* __restrict__ has_boot, int* __restrict__ flags) {
    const int b = blockIdx.x, tid = threadIdx.x;
    long long i = index[b];
    const bool oob = i < 0 || i >= n_valid;   // deque[i] of the reference raises IndexError
    if (oob) i = 0;
    const unsigned char* rec = ring + (size_t)((head + i) % capacity) * rec_bytes;
    const int length = ((const int*)rec)[2];
    if (tid == 0) {
        partial[b] = *(const double*)rec;
        has_boot[b] = ((const int*)rec)[3];
        if (oob) atomicAdd(flags + 0, 1);
        if (length == 0) atomicAdd(flags + 1, 1);   // "Encountered empty item in batch."
    }
    const float* w = (const float*)(rec + 16);
    for (long long k = tid; k < obs_size; k += 128) {
        obs[(size_t)b * obs_size + k] = w[k];
        boot_obs[(size_t)b * obs_size + k] = w[obs_size + k];
    }
    w += 2 * obs_size;
    for (int k = tid; k < A; k += 128) policy[(size_t)b * A + k] = w[k];
    w += A;
    for (int k = tid; k < K; k += 128) {
        rewards[(size_t)b * K + k] = w[k];
        actions[(size_t)b * K + k] = ((const int*)(w + K))[k];
    }
}

// values[b] = float32(partial[b] + discount**n * float64(bootstrap value)), the bootstrap only where the trajectory has one
__global__ void replay_values_kernel(const double* __restrict__ partial, const int* __restrict__ has_boot, const float* __restrict__ boot_value,
                                     double discount_pow_n, int B, float* __restrict__ values) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double v = partial[b];
    if (has_boot[b]) v = __dadd_rn(v, __dmul_rn(discount_pow_n, (double)boot_value[b]));
    values[b] = (float)v;
}

cudaError_t launch_replay_gather(const unsigned char* ring, long long capacity, long long rec_bytes, long long obs_size, int A, int K,
                                 const long long* index, long long head, long long n_valid, int B, float* obs, float* boot_obs, float* policy,
                                 float* rewards, int* actions, double* partial, int* has_boot, int* flags, cudaStream_t st) {
    replay_gather_kernel<<<B, 128, 0, st>>>(ring, capacity, rec_bytes, obs_size, A, K, index, head, n_valid, obs, boot_obs, policy, rewards,
                                            actions, partial, has_boot, flags);
    return cudaGetLastError();
}

cudaError_t launch_replay_values(const double* partial, const int* has_boot, const float* boot_value, double discount_pow_n, int B, float* values,
                                 cudaStream_t st) {
    replay_values_kernel<<<(B + 127) / 128, 128, 0, st>>>(partial, has_boot, boot_value, discount_pow_n, B, values);
    return cudaGetLastError();
}

}  // namespace cm
