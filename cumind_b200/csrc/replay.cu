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

template <bool SMEM>
__global__ void __launch_bounds__(32, 1) sumtree_update_kernel(double* __restrict__ tree_g, long long n_nodes, int depth,
                                                                  const long long* __restrict__ idx, const double* __restrict__ pri,
                                                                  long long n) {
    extern __shared__ double tree_s[];
    const int lane = threadIdx.x;
    double* tree = SMEM ? tree_s : tree_g;
    if (SMEM) {
        for (long long i = lane; i < n_nodes; i += 32) tree_s[i] = tree_g[i];
        __syncwarp();
    }
    for (long long k = 0; k < n; ++k) {
        const long long leaf = idx[k];
        const double p = pri[k];
        const double change = p - tree[leaf];   // every lane reads the same word
        __syncwarp();
        if (lane == 0) tree[leaf] = p;
        __syncwarp();
        // the level of `leaf` in the tree: leaves of the complete tree sit at `depth`, but an update may name any node
        // (the reference does not check); ancestors are ((leaf + 1) >> j) - 1 for j = 1 .. level(leaf)
        const long long l1 = leaf + 1;
        const int level = 63 - __clzll(l1);
        for (int j0 = 1; j0 <= level; j0 += 32) {
            const int j = j0 + lane;
            if (j <= level) {
                const long long a = (l1 >> j) - 1;
                tree[a] += change;
            }
        }
        __syncwarp();
        if (lane == 0) tree[n_nodes - 1] += change;   // the reference's extra step from the root: sum_tree[-1]
        __syncwarp();
        if (!SMEM) __threadfence_block();
    }
    if (SMEM) {
        __syncwarp();
        for (long long i = lane; i < n_nodes; i += 32) tree_g[i] = tree_s[i];
    }
    (void)depth;
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
    if (bytes <= 200 * 1024) {
        auto k = sumtree_update_kernel<true>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
        k<<<1, 32, bytes, st>>>(tree, n_nodes, depth, idx, pri, n);
    } else {
        sumtree_update_kernel<false><<<1, 32, 0, st>>>(tree, n_nodes, depth, idx, pri, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_sumtree_sample(const double* tree, long long n_nodes, long long tree_size, const double* uniform, long long batch,
                                  long long* tree_idx, long long* data_idx, cudaStream_t st) {
    sumtree_sample_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(tree, n_nodes, tree_size, uniform, batch, tree_idx, data_idx);
    return cudaGetLastError();
}

}  // namespace cm
