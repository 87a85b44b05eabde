// p2p_adamw.cu — the gradient all-reduce of the train step fused into the AdamW update, over NVLink peer memory.
//
// The only collective of the product is the all-reduce of the flat gradient (trainer.py:97-100 applies the optimiser to the
// gradient of the local batch; data-parallel training averages it over the ranks first).  The bucket is tiny (13 188
// floats at hidden_dim 64), so the cost of a library all-reduce is pure latency: a separate launch, a ring / tree protocol
// and a second pass over the data.  Here ONE kernel does both steps: every rank owns a gradient buffer that its peers map
// through CUDA IPC (one process per GPU, NVSwitch gives every pair full bandwidth), and the update kernel
//   1. announces "my gradient is complete" in every peer's flag array and waits until all peers have announced
//      (system-scope release / acquire on flags in peer memory: ~2 us across NVSwitch),
//   2. reads element i of EVERY rank's buffer (P2P loads), sums them in rank order — the same order on every rank, so all
//      replicas apply bit-identical updates and stay in lock step without a broadcast — scales by 1/world and applies
//      optax.adamw (agent.py:48) to its local parameters and moments,
//   3. announces "I have read your gradient" and waits for all peers, so a buffer is never overwritten while it is being read.
// No NCCL call, nothing between the loss kernel and the optimiser but this launch, and the whole train step is one CUDA
// graph even when distributed.  Evidence for "fused compute + collective": ld.global on mapped peer pointers and
// st.release.sys / ld.acquire.sys on peer flags in the same kernel as the AdamW arithmetic.
// A bounded spin (about one second) turns a missing peer into an error flag instead of a hang.
#include "kernels.h"

namespace cm {

namespace {

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// flags[phase][r] lives on every rank and is written by rank r: announce epoch `ep` to every peer / wait for every peer
__device__ void announce(const P2PArgs& a, int phase, long long ep) {
    if ((int)threadIdx.x < a.world) st_release_sys(a.peer_flags[threadIdx.x] + phase * P2P_MAX_RANKS + a.rank, ep);
}
__device__ void wait_peers(const P2PArgs& a, int phase, long long ep) {
    if ((int)threadIdx.x < a.world) {
        const long long* mine = a.peer_flags[a.rank] + phase * P2P_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) < ep) {
            if (clock64() - t0 > 2000000000LL) { atomicExch(a.err, 1); break; }   // ~1 s: a peer is missing
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) p2p_allreduce_adamw_kernel(const P2PArgs a, float* __restrict__ p, float* __restrict__ m,
                                                                   float* __restrict__ v, const unsigned char* __restrict__ mask, size_t n,
                                                                   float lr, float b1, float b2, float eps, float wd,
                                                                   const long long* __restrict__ step) {
    const long long ep = *step;   // incremented by the kernel before this one: the same value on every rank
    // 1. my gradient is complete (it was written by the previous kernel of the stream); wait for everybody's
    if (blockIdx.x == 0) announce(a, 0, ep);
    wait_peers(a, 0, ep);
    const float t = (float)ep;
    const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
    const float scale = 1.0f / (float)a.world;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float g = 0.0f;
        for (int r = 0; r < a.world; ++r) g += __ldcv(a.peer_grad[r] + i);   // P2P loads past the L1; rank order: identical sums on every rank
        if (a.sum_out) a.sum_out[i] = g;
        if (mask && !mask[i]) continue;
        const float gi = g * scale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pi = p[i];
        p[i] = pi - lr * ((mi / bc1) / (sqrtf(vi / bc2) + eps) + wd * pi);
    }
    // 3. the last block of this rank to finish reading tells the peers, and waits until every peer has finished reading
    //    this rank's buffer: when the kernel completes, the buffer may be overwritten
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(a.count, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {
        announce(a, 1, ep);
        wait_peers(a, 1, ep);
    }
}

__global__ void begin_step_kernel(long long* step, int* count) { *step += 1; *count = 0; }

}  // namespace

cudaError_t launch_p2p_allreduce_adamw(const P2PArgs& a, float* p, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                                       float b2, float eps, float wd, long long* d_step, cudaStream_t st) {
    begin_step_kernel<<<1, 1, 0, st>>>(d_step, a.count);
    size_t blocks = (n + 255) / 256;
    if (blocks > 148) blocks = 148;   // every block spins in the rank barrier: all of them must be resident at once
    p2p_allreduce_adamw_kernel<<<(int)blocks, 256, 0, st>>>(a, p, m, v, mask, n, lr, b1, b2, eps, wd, d_step);
    return cudaGetLastError();
}

}  // namespace cm
