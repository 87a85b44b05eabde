// kernels.h — host-visible launchers of the CUDA kernels (implemented in the .cu files of this dir).
#pragma once

#include "common.cuh"

namespace cm {

struct FusedParams {
    TreeView tv;
    const float* pack;      // device, FusedPackLayout order
    FusedPackLayout pk;
    const float* root_hidden;  // [B][H]
    const double* noise;       // [B][A] or nullptr
    int32_t* out_visits;       // [B][A_cfg]
    double* out_probs;         // [B][A_cfg]
    double* out_root_value;    // [B] or nullptr
    int32_t* out_actions;      // [B] or nullptr: int(np.argmax(probs)) (agent.py:86), first maximum
    int S, A_cfg, A_net;
    double c_puct, frac, one_minus_frac, alpha;
    int noise_mode;
    unsigned long long seed, tree_base;
    int path_in_smem;
    // shared-memory carve-up (bytes)
    unsigned off_sqrt, off_recip, off_slots, slot_bytes, off_bar;
    int n_sqrt;
    int pdl;                   // host only: launch with programmatic stream serialization (root_hidden comes from the previous kernel)
};

#ifdef __CUDACC__
// <<<grid, block, smem, st>>> with an optional programmatic-dependent-launch attribute: the kernel may start while
// the previous kernel of the stream drains and must call pdl_wait() before it reads that kernel's output.
template <class... KArgs, class... Args>
inline cudaError_t launch_kernel_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// search_fused.cu
cudaError_t launch_search_fused(int G, int R, const FusedParams& p, int grid, int block, size_t smem, cudaStream_t st);

// search_team.cu — the same search with warp-specialised teams (tree warps + network warps)
struct TeamParams {
    TreeView tv;
    const float* pack;      // device, TeamPackLayout order
    TeamPackLayout pk;
    const float* root_hidden;  // [B][H]
    const double* noise;       // [B][A] or nullptr
    int32_t* out_visits;       // [B][A_cfg]
    double* out_probs;         // [B][A_cfg]
    double* out_root_value;    // [B] or nullptr
    int32_t* out_actions;      // [B] or nullptr: int(np.argmax(probs)) (agent.py:86), first maximum
    int S, A_cfg, A_net;
    double c_puct, frac, one_minus_frac, alpha;
    int noise_mode;
    unsigned long long seed, tree_base;
    // team geometry: n_team teams per CTA, each ntw tree warps + nmw network warps; TT tree slots per
    // team (compile-time in the kernel: 8 or 16), of which tpt are used per batch
    int n_team, ntw, nmw, TT, tpt, n_batches;
    int tile;                  // lane tile of the network warps (0 = default; see launch_search_team)
    int stagger;               // SM cycles between the starts of the teams of one CTA
    // shared-memory carve-up (bytes); off_path / off_nxt / off_hotw / off_hot are relative to the team's scratch
    unsigned off_sqrt, off_recip, off_teams, team_bytes, off_bar, off_path, off_nxt, off_hotw, off_hot;
    int lw;                    // layer-warpgroup mode (warps 0-3 run the dynamics layers with register-resident weights); 0 = classic
    int n_sqrt;
    long long* clk;            // [grid][16] phase clocks (diagnostic) or nullptr
    int pdl;                   // host only: launch with programmatic stream serialization
};
bool team_variant_ok(int G, int TT);
bool team_lw_variant_ok(int G, int TT, int H);   // a layer-warpgroup instantiation exists
int team_lw_team_warps(int n_team);
cudaError_t launch_search_team(int G, const TeamParams& p, int grid, int block, size_t smem, cudaStream_t st);

// tree_steps.cu — one launch per phase (warp per tree, any H / A)
struct StepParams {
    TreeView tv;
    const float* pack;
    PackLayout pk;
    int S, A_cfg, A_net;
    double c_puct, frac, one_minus_frac, alpha;
    int noise_mode;
    unsigned long long seed, tree_base;
    int tree_mode;     // CUMIND_TREE_*
    double discount;   // CUMIND_TREE_MINMAX_DISCOUNT
};
cudaError_t launch_init_root(const StepParams& p, const float* root_hidden, const double* noise, cudaStream_t st);
cudaError_t launch_select(const StepParams& p, cudaStream_t st);
cudaError_t launch_expand(const StepParams& p, float* leaf_value, cudaStream_t st);
cudaError_t launch_backup(const StepParams& p, const float* leaf_value, cudaStream_t st);
// CUMIND_TREE_MINMAX_DISCOUNT (optional mode, see include/cumind_b200.h): select and backup of that mode
cudaError_t launch_select_mm(const StepParams& p, cudaStream_t st);
cudaError_t launch_backup_mm(const StepParams& p, const float* leaf_value, cudaStream_t st);
cudaError_t launch_finalize(const StepParams& p, int32_t* out_visits, double* out_probs, double* out_root_value, cudaStream_t st);
cudaError_t launch_dirichlet(double* noise, int B, int A, double alpha, unsigned long long seed,
                             unsigned long long tree_base, cudaStream_t st);
cudaError_t launch_selftest_div(unsigned long long n_per_thread, unsigned long long seed, int max_d, unsigned long long* counts,
                                cudaStream_t st);
cudaError_t launch_u8_to_f32(const unsigned char* in, float* out, size_t n, cudaStream_t st);
cudaError_t launch_argmax_first(const double* probs, int B, int A, int32_t* actions, cudaStream_t st);
cudaError_t launch_stage_in(const void* h_a, void* d_a, size_t n_a, const void* h_b, void* d_b, size_t n_b, cudaStream_t st);
cudaError_t launch_stage_out(const double* probs, int B, int A, int32_t* d_actions, int32_t* h_actions, double* h_probs, cudaStream_t st);

// network_fp32.cu — batched fp32-exact network entry points (warp per row)
cudaError_t launch_vector_encoder(const NetView& nv, const float* obs, int B, float* hidden, cudaStream_t st);
cudaError_t launch_recurrent(const NetView& nv, const float* hidden, const int32_t* action, int B, float* next,
                             float* reward, float* logits, float* value, cudaStream_t st);
cudaError_t launch_prediction(const NetView& nv, const float* hidden, int B, float* logits, float* value, cudaStream_t st);
cudaError_t launch_softmax(const float* logits, int B, int A, float* probs, cudaStream_t st);

// network_tc.cu — bf16 tensor-core (tcgen05) dynamics + prediction, 128-row tiles
bool tc_supported(int H, int L, int A);
size_t tc_tile_offset_host(int r, int k, int rows);
cudaError_t launch_recurrent_tc(const unsigned char* pack, const TcPackLayout& pk, int sm_count, const float* hidden,
                                const int32_t* action, int B, float* next, float* reward, float* logits, float* value,
                                int with_dynamics, cudaStream_t st);
cudaError_t launch_expand_tc(const unsigned char* pack, const TcPackLayout& pk, int sm_count, const TreeView& tv, float* leaf_value,
                             cudaStream_t st);

// search_tc.cu — the whole search in one launch with the per-simulation network on tcgen05 (CUMIND_NET_BF16_TC)
struct SearchTcParams {
    TreeView tv;
    const unsigned char* pack;   // TcPackLayout bytes (bf16 weight tiles, fp32 biases and embedding)
    TcPackLayout pk;
    const float* pack_f32;       // PackLayout (fp32): the root prediction is computed in pinned-order fp32
    PackLayout pkf;
    const float* root_hidden;    // [B][H]
    const double* noise;         // [B][A] or nullptr
    int32_t* out_visits;         // [B][A_cfg]
    double* out_probs;           // [B][A_cfg]
    double* out_root_value;      // [B] or nullptr
    int32_t* out_actions;        // [B] or nullptr: first arg-max of the visit counts
    int S, A_cfg;
    double c_puct, frac, one_minus_frac, alpha;
    int noise_mode;
    unsigned long long seed, tree_base;
    int n_sqrt;                  // entries of the sqrt / reciprocal tables (2 S_max + 4)
    int pdl;                     // host only: programmatic dependent launch
};
cudaError_t launch_search_tc(const SearchTcParams& p, int sm_count, cudaStream_t st);

// conv_tc.cu — image encoder as implicit GEMM on tcgen05 (bf16 mode)
bool conv_tc_supported(const cumind_net_desc& d);
ConvTcPackLayout make_conv_tc_pack_layout(const cumind_net_desc& d);
size_t conv_tc_work_bytes(const cumind_net_desc& d, int chunk);
cudaError_t launch_conv_encoder_tc(const NetView& nv, const unsigned char* cpack, const ConvTcPackLayout& cpk, int sm_count,
                                   const float* obs, int B, float* hidden, unsigned char* work, int chunk, cudaStream_t st);

// train.cu — parameter re-layout on the device (gather maps) and fused AdamW over the flat parameter blob
cudaError_t launch_repack_f32(const float* blob, const int* map, float* dst, size_t n, cudaStream_t st);
cudaError_t launch_repack_bf16(const float* blob, const int* map, unsigned char* dst, size_t n, cudaStream_t st);
cudaError_t launch_bn_fold(const float* blob, const BnFold* folds, int n, float* cpack_f32, cudaStream_t st);
cudaError_t launch_adamw_dev_step(float* p, const float* g, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                                  float b2, float eps, float wd, long long* d_step, float grad_scale, cudaStream_t st);
cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                         float b2, float eps, float wd, long long step, float grad_scale, cudaStream_t st);

// train_fused.cu — loss + gradient of the train step in one kernel (vector observations, hidden_dim 32 / 64)
bool train_fused_supported(const cumind_net_desc& d, int K);
size_t train_fused_work_bytes(const cumind_net_desc& d, int B);
cudaError_t launch_train_fwd_bwd(const cumind_net_desc& d, const float* blob, const float* obs, const int32_t* actions, const float* tv,
                                 const float* tp, const float* tr, int B, int K, float* work, float* grad, float* losses, cudaStream_t st);

// p2p_adamw.cu — gradient all-reduce over NVLink peer memory fused into AdamW (one process per GPU, CUDA IPC)
constexpr int P2P_MAX_RANKS = 16;
struct P2PArgs {
    int rank, world;
    const float* peer_grad[P2P_MAX_RANKS];   // every rank's gradient buffer as mapped here (own buffer at [rank])
    long long* peer_flags[P2P_MAX_RANKS];    // every rank's flag array [2][P2P_MAX_RANKS]
    float* sum_out;                          // optional: the summed gradient (tests)
    int* err;                                // set when a peer did not arrive within the spin bound
    int* count;                              // blocks of this rank that finished reading (reset by the pre-kernel)
};
cudaError_t launch_p2p_allreduce_adamw(const P2PArgs& a, float* p, float* m, float* v, const unsigned char* mask, size_t n, float lr, float b1,
                                       float b2, float eps, float wd, long long* d_step, cudaStream_t st);

// replay.cu — TreeBuffer's sum tree on the device (memory.py:190-299)
cudaError_t launch_sumtree_update(double* tree, long long n_nodes, int depth, const long long* idx, const double* pri, long long n,
                                  cudaStream_t st);
cudaError_t launch_replay_gather(const unsigned char* ring, long long capacity, long long rec_bytes, long long obs_size, int A, int K,
                                 const long long* index, long long head, long long n_valid, int B, float* obs, float* boot_obs, float* policy,
                                 float* rewards, int* actions, double* partial, int* has_boot, int* flags, cudaStream_t st);
cudaError_t launch_replay_values(const double* partial, const int* has_boot, const float* boot_value, double discount_pow_n, int B, float* values,
                                 cudaStream_t st);
cudaError_t launch_sumtree_sample(const double* tree, long long n_nodes, long long tree_size, const double* uniform, long long batch,
                                  long long* tree_idx, long long* data_idx, cudaStream_t st);

// conv_fp32.cu — image encoder, fp32-exact (ConvEncoder, network.py:111-147)
size_t conv_encoder_work_bytes(const cumind_net_desc& d, int chunk);
cudaError_t launch_conv_encoder(const NetView& nv, const float* obs, int B, float* hidden, float* work, int chunk,
                                cudaStream_t st);

}  // namespace cm
