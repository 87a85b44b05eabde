/*
 * cumind_b200.h — C ABI of the B200-native CuMind search hot path.
 *
 * The reference (carletonai/CuMind v0.1.7) has no FFI of its own: its boundary is the
 * Python surface Agent.select_action -> MCTS.search -> CuMindNetwork.{initial,recurrent}_inference
 * (src/cumind/agent/agent.py:61-89, src/cumind/core/mcts.py:115-222,
 * src/cumind/core/network.py:288-315).  Each entry point below names the reference
 * function it replaces.  Everything is `extern "C"`, plain pointers and sizes; no torch
 * or jax types.  INTEGRATION.md shows the ctypes and jax.ffi bindings.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; cumind_last_error()
 *     returns a thread-local, NUL-terminated description of the last failure.
 *   - pointers named d_* are DEVICE pointers owned by the caller (torch / XLA buffers);
 *     pointers named h_* are HOST pointers.  Nothing is allocated after *_create.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all
 *     device entry points are asynchronous with respect to the host.
 *   - one host thread per device.
 */
#ifndef CUMIND_B200_H
#define CUMIND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUMIND_ABI_VERSION 3

/* ------------------------------------------------------------------------------------------
 * Network descriptor + flat parameter blob.
 *
 * Mirrors the constructor of CuMindNetwork (network.py:272-286).  The parameter blob is one
 * contiguous fp32 array holding the reference's parameter tree in this
 * order (all kernels row-major exactly as flax stores them, [in,out] / [kh,kw,cin,cout]):
 *
 *   representation_network/encoder
 *     vector obs (obs_ndim==1):  for i in 0..L-1:  layers/i/kernel[in_i,H], layers/i/bias[H]
 *                                (in_0 = obs_shape[0], in_i = H)
 *     image  obs (obs_ndim==3):  initial_conv/kernel[3,3,C,Cc], initial_conv/bias[Cc]
 *                                for i in 0..L-1: conv1/kernel[3,3,Cc,Cc], conv1/bias[Cc],
 *                                   bn1/{scale,bias,mean,var}[Cc], conv2/kernel, conv2/bias,
 *                                   bn2/{scale,bias,mean,var}[Cc]
 *                                final_dense/kernel[Cc,H], final_dense/bias[H]
 *                                (C = obs_shape[2]: the LAST axis is channels, network.py:127)
 *   dynamics_network:   action_embedding/embedding[A,H]; for l in 0..L-1: layers/l/kernel[H,H],
 *                       layers/l/bias[H]; reward_head/kernel[H,1], reward_head/bias[1]
 *   prediction_network: policy_head/kernel[H,A], policy_head/bias[A];
 *                       value_head/kernel[H,1], value_head/bias[1]
 * ------------------------------------------------------------------------------------------ */
typedef struct cumind_net_desc {
    int32_t obs_ndim;          /* 1 (vector) or 3 (image, NHWC)                       */
    int32_t obs_shape[3];      /* (D,0,0) or (height, width, channels)                */
    int32_t hidden_dim;        /* H   (config.hidden_dim)                             */
    int32_t num_blocks;        /* L   (config.num_blocks)                             */
    int32_t action_space_size; /* A   (logits of the policy head / embedding rows)    */
    int32_t conv_channels;     /* Cc  (config.conv_channels; ignored for vector obs)  */
} cumind_net_desc;

/* Arithmetic mode of the network kernels. */
enum {
    CUMIND_NET_FP32_EXACT = 0, /* pinned-order fp32 SIMT: bit-identical to oracle/ (parity mode) */
    CUMIND_NET_BF16_TC    = 1  /* bf16 operands, fp32 accumulate on tcgen05 tensor cores          */
};

/* Tree-policy mode. */
enum {
    CUMIND_TREE_REFERENCE = 0, /* exactly mcts.py: +inf for n==0, no min-max Q, undiscounted
                                  backup of parents only, root backed up twice (mcts.py:51-65,
                                  157-203).  The default, and the only mode parity with the
                                  reference is defined on.                                   */
    CUMIND_TREE_MINMAX_DISCOUNT = 1
                               /* OPTIONAL textbook-MuZero rules the reference does NOT implement (they are what
                                  mcts.py:51-65, 199-203 deviate from): score = normalise(reward + discount*W/n)
                                  (0 while unvisited) + ((c_puct*prior)*sqrt(N_parent))/(1+n) with per-tree min-max
                                  bounds of the backed-up q values; the backup runs from the leaf (counted) to the
                                  root (once) with G <- reward + discount*G.  Defined by oracle/cumind_oracle.c
                                  (cmo_search_mm) and oracle/mcts_minmax.py; float64 statistics, IEEE operations in
                                  the order written there.  Runs on the one-launch-per-phase kernels.            */
};

typedef struct cumind_search_cfg {
    int32_t num_simulations;      /* S        config.num_simulations          (mcts.py:139) */
    int32_t action_space_size;    /* A_cfg    config.action_space_size; children per node =
                                              min(A_cfg, net A) (zip truncation, mcts.py:88) */
    double  c_puct;               /*          config.c_puct                   (mcts.py:169) */
    double  exploration_fraction; /* f        config.exploration_fraction     (mcts.py:222) */
    double  dirichlet_alpha;      /* alpha    config.dirichlet_alpha          (mcts.py:218) */
    int32_t noise_mode;           /* 0 none (add_noise=False); 1 explicit d_noise[B,A_tree] f64;
                                     2 on-device Philox Dirichlet(alpha) from (seed, tree index) */
    int32_t net_mode;             /* CUMIND_NET_*                                            */
    int32_t tree_mode;            /* CUMIND_TREE_*                                           */
    int32_t lanes_per_tree;       /* 0 = auto; else 2,4,8,16,32 (tuning knob, no effect on results) */
    uint64_t seed;                /* Philox key for noise_mode 2                             */
    uint64_t tree_index_base;     /* global index of local tree 0 (multi-GPU sharding: results
                                     do not depend on the GPU count)                         */
    /* launch tuning of cumind_search; none of these changes a result bit */
    int32_t schedule;             /* CUMIND_SCHED_*: 0 = auto                                */
    int32_t teams_per_cta;        /* team schedule: 0 = auto, else 1..7                      */
    int32_t net_warps_per_team;   /* team schedule: 0 = auto (4), else 1..8                  */
    int32_t trees_per_team;       /* team schedule: 0 = auto, else tree slots used per batch */
    double  discount;             /* config.discount (config.py:49); read by CUMIND_TREE_MINMAX_DISCOUNT only */
} cumind_search_cfg;

/* How cumind_search maps the work onto the GPU (all produce identical results):
 *   WARP  search_fused.cu: every group of `lanes_per_tree` lanes runs the whole search of one tree
 *   TEAM  search_team.cu : warp-specialised teams (tree-walk warps + network warps); with
 *         hidden_dim 64 the dynamics layers of all teams run on a dedicated layer warpgroup whose weights are
 *         register-resident
 * CUMIND_NET_BF16_TC adds
 *   TENSOR search_tc.cu : 128 trees per CTA (thread = tree = row of a UMMA tile), dynamics + prediction of every
 *         simulation as tcgen05 MMAs with TMEM accumulators; results are those of the bf16 network, not bit-pinned.
 *         AUTO picks it when the batch is large enough for it to win (more than CUMIND_TENSOR_MIN_TREES_PER_SM trees
 *         per SM: measured per search 0.73 / 0.92 / 1.85 ms against 0.83 / 1.72 / 3.29 ms for the fp32 kernels at 16 384 /
 *         32 768 / 65 536 trees, and 0.73 against 0.53 ms at 8 192); below that AUTO keeps the
 *         fp32-exact fused kernels, which are faster there and exact — cumind_search_plan reports the choice, and
 *         schedule = CUMIND_SCHED_TENSOR forces the tensor-core kernel at any size. */
#define CUMIND_TENSOR_MIN_TREES_PER_SM 80
#define CUMIND_SCHED_AUTO 0
#define CUMIND_SCHED_WARP 1
#define CUMIND_SCHED_TEAM 2
#define CUMIND_SCHED_TEAM_CLASSIC 3 /* TEAM without the layer warpgroup (the network warps compute the layers) */
#define CUMIND_SCHED_TENSOR 4       /* CUMIND_NET_BF16_TC only: search_tc.cu, the per-simulation network on tcgen05 */

typedef struct cumind_net  cumind_net;   /* opaque: packed parameters on one device  */
typedef struct cumind_tree cumind_tree;  /* opaque: view over a caller-owned SoA slab */

const char* cumind_last_error(void);
int         cumind_abi_version(void);

/* Number of fp32 values in the parameter blob for `desc` (0 on invalid desc). */
size_t cumind_param_count(const cumind_net_desc* desc);

/* Replaces CuMindNetwork.__init__ + nnx.update (network.py:272-286, agent.py:109-121):
 * copies the host blob to the current device and builds the kernel-side layouts
 * (bf16 copies for the tensor-core mode).  h_params has cumind_param_count(desc) floats. */
int cumind_net_create(const cumind_net_desc* desc, const float* h_params, size_t n_params,
                      cumind_net** out);
int cumind_net_update_params(cumind_net* net, const float* h_params, size_t n_params, void* stream);
int cumind_net_destroy(cumind_net* net);
/* Image encoders keep their layer activations in a device workspace owned by `net`.  cumind_net_reserve sizes it
 * for batches of up to max_batch observations (both arithmetic modes); call it once after cumind_net_create —
 * it allocates and synchronises the device.  cumind_initial_inference never allocates: it fails with an error
 * that names this function when the reserved workspace is too small.  Vector-observation networks need none. */
int cumind_net_reserve(cumind_net* net, int32_t max_batch);

/* CuMindNetwork.initial_inference (network.py:288-300).
 * d_obs [B, *obs_shape] fp32 -> d_hidden [B,H], d_logits [B,A] (may be NULL), d_value [B] (may be NULL). */
int cumind_initial_inference(const cumind_net* net, const float* d_obs, int32_t B,
                             float* d_hidden, float* d_logits, float* d_value,
                             int32_t net_mode, void* stream);

/* CuMindNetwork.recurrent_inference (network.py:302-315) = DynamicsNetwork.__call__ (217-236)
 * followed by PredictionNetwork.__call__ (254-266).
 * d_hidden [B,H], d_action [B] int32 -> d_next [B,H], d_reward [B], d_logits [B,A], d_value [B]
 * (any output may be NULL). */
int cumind_recurrent_inference(const cumind_net* net, const float* d_hidden, const int32_t* d_action,
                               int32_t B, float* d_next, float* d_reward, float* d_logits,
                               float* d_value, int32_t net_mode, void* stream);

/* PredictionNetwork.__call__ alone (network.py:254-266; called directly at mcts.py:127,190). */
int cumind_prediction(const cumind_net* net, const float* d_hidden, int32_t B,
                      float* d_logits, float* d_value, int32_t net_mode, void* stream);

/* jax.nn.softmax over the last axis with the pinned exp (mcts.py:128,191). d_logits/d_probs [B,A]. */
int cumind_softmax(const float* d_logits, int32_t B, int32_t A, float* d_probs, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tree storage: B independent trees as flat arrays in one caller-owned device slab.
 * Replaces the Python `Node` object graph (mcts.py:18-98).  Node slots per tree:
 *   N = 1 + A_tree*(S_max+1); slot 0 is the root; the e-th expansion (e=0 is the root) owns
 *   child slots 1+e*A_tree .. 1+(e+1)*A_tree-1 and hidden-state row e.
 * ------------------------------------------------------------------------------------------ */
size_t cumind_tree_bytes(int32_t B, int32_t S_max, int32_t A_tree, int32_t H);
int cumind_tree_create(int32_t B, int32_t S_max, int32_t A_tree, int32_t H,
                       void* d_slab, size_t slab_bytes, cumind_tree** out);
int cumind_tree_destroy(cumind_tree* tree);

/* One 16-byte record per node; the A children of a node are contiguous, so one 32-byte sector
 * serves two children and one 128-bit load fetches everything Node.ucb_score reads. */
typedef struct cumind_node {
    double  value_sum;  /* Node.value_sum  (float64, mcts.py:29)                          */
    int32_t visit;      /* Node.visit_count                                               */
    float   prior;      /* Node.prior as produced by the fp32 softmax (root children: the
                           float64 value after noise mixing is in root_prior)             */
} cumind_node;

/* Device pointers of the arrays inside the slab (for parity dumps / zero-copy readers).
 * Slots that no expansion has created yet are uninitialised. */
typedef struct cumind_tree_view {
    int32_t B, S_max, A, H, N;
    cumind_node* nodes;    /* [B,N]        node records                                      */
    int32_t* child_eidx;   /* [B,N]        -1 if not expanded, else expansion index e        */
    double*  root_prior;   /* [B,A]        root children priors after noise mixing (float64) */
    int32_t* exp_node;     /* [B,S_max+1]  slot of the e-th expanded node                    */
    float*   hidden;       /* [B,S_max+1,H] Node.hidden_state of the e-th expanded node      */
    int32_t* path;         /* [B,S_max+2]  scratch: parents on the current path (step API)   */
    int32_t* path_len;     /* [B]                                                            */
    int32_t* leaf;         /* [B]          scratch: slot of the selected leaf                */
    int32_t* n_expanded;   /* [B]          expansions so far (root included)                 */
    int64_t* depth_sum;    /* [B]          sum of path lengths over all simulations          */
    float*   reward;       /* [B,N]        CUMIND_TREE_MINMAX_DISCOUNT: the dynamics reward of each expanded node */
    double*  minmax;       /* [B,2]        CUMIND_TREE_MINMAX_DISCOUNT: bounds {min, max} of the backed-up q values */
} cumind_tree_view;
int cumind_tree_get_view(const cumind_tree* tree, cumind_tree_view* out);

/* MCTS.search (mcts.py:115-155) for B trees in lockstep, all on device, one fused launch:
 * root prediction + softmax + root expand + optional noise + S x (select, dynamics, prediction,
 * expand, backup) + visit-count normalisation.
 *   d_root_hidden [B,H] fp32            (output of initial_inference)
 *   d_noise       [B,A_tree] f64 or NULL (noise_mode 1)
 *   d_out_visits  [B,A_cfg] int32       root.children[a].visit_count
 *   d_out_probs   [B,A_cfg] f64         visits/sum, uniform if sum==0 (mcts.py:143-153)
 *   d_out_root_value [B] f64 or NULL    root.value_sum / root.visit_count
 * The full trees stay in `tree` for inspection. */
int cumind_search(cumind_tree* tree, const cumind_net* net, const float* d_root_hidden,
                  const cumind_search_cfg* cfg, const double* d_noise,
                  int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                  void* stream);

/* Agent.select_action for a batch with every buffer already on the device (agent.py:73-79: the reference encodes
 * and searches in one call too): initial_inference (representation network, network.py:288-300) into the caller's
 * d_hidden_work [B,H], then cumind_search on it, both issued from this one call so no host work separates the
 * launches; for vector observations the search is a programmatic dependent launch whose parameter staging
 * overlaps the encoder.  Other arguments as cumind_search. */
int cumind_search_from_obs(cumind_tree* tree, const cumind_net* net, const float* d_obs,
                           const cumind_search_cfg* cfg, const double* d_noise, float* d_hidden_work,
                           int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                           void* stream);

/* The launch configuration cumind_search would use: out = {schedule (0 = stepwise fallback,
 * CUMIND_SCHED_WARP, CUMIND_SCHED_TEAM), lanes per tree G, WARP: outputs per lane R / TEAM: trees
 * per team batch, grid, block, dynamic shared memory bytes, TEAM: teams per CTA, TEAM: network
 * warps per team + 256 * (register-resident k of the layer warpgroup, 0 = none)}. */
int cumind_search_plan(const cumind_tree* tree, const cumind_net* net, const cumind_search_cfg* cfg,
                       int32_t* out8);

/* Diagnostic: cumind_search on the TEAM schedule with phase clocks.  d_clocks is a zero-initialised
 * device int64[grid*16] (grid from cumind_search_plan): per CTA, SM cycles spent by tree warp 0 of
 * team 0 in {select, waiting for the network, softmax+expand+backup, root+final} and by network
 * warp 0 in {gather, layers, heads, waiting for the tree warps}. */
int cumind_search_phase_clocks(cumind_tree* tree, const cumind_net* net, const float* d_root_hidden,
                               const cumind_search_cfg* cfg, const double* d_noise,
                               int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                               int64_t* d_clocks, void* stream);

/* Same search issued as separate kernels per phase (the fine-grained entry points below);
 * used by the tests to cross-check the fused kernel and by ncu to attribute time per phase. */
int cumind_search_stepwise(cumind_tree* tree, const cumind_net* net, const float* d_root_hidden,
                           const cumind_search_cfg* cfg, const double* d_noise,
                           int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                           void* stream);

/* Agent.select_action for a batch (agent.py:61-89, training=False argmax path; the sampling
 * draw of training=True stays on the host): HOST observations in, HOST results out.
 * h_obs [B,*obs_shape] fp32 (pinned or pageable) -> h_actions [B] int32 (first arg-max of
 * probs), h_probs [B,A_cfg] f64.  Performs H2D, initial_inference, search, D2H and
 * synchronises `stream`.  d_work is a caller-owned device scratch of
 * cumind_select_actions_work_bytes() bytes. */
size_t cumind_select_actions_work_bytes(const cumind_net_desc* desc, int32_t B, int32_t A_cfg);
int cumind_select_actions_host(cumind_tree* tree, const cumind_net* net, const float* h_obs,
                               const cumind_search_cfg* cfg, const double* h_noise,
                               int32_t* h_actions, double* h_probs, void* d_work, size_t work_bytes,
                               void* stream);

/* The same call for uint8 observations (image frames): h_obs_u8 [B,*obs_shape] uint8; the device computes
 * float32(u8) / 255 correctly rounded — what `frames.astype(np.float32) / 255` gives on the host — so the results equal
 * cumind_select_actions_host on those floats, with a quarter of the bytes crossing PCIe.  Batches of 512 or more frames are
 * copied in two slices on an internal second stream (one per device, created on first use) while the encoder already works
 * on the first; the call still returns with every result in place.  One call per device at a time. */
int cumind_select_actions_host_u8(cumind_tree* tree, const cumind_net* net, const uint8_t* h_obs_u8,
                                  const cumind_search_cfg* cfg, const double* h_noise,
                                  int32_t* h_actions, double* h_probs, void* d_work, size_t work_bytes,
                                  void* stream);

/* Fine-grained phases (one launch each).  Reference anchors:
 *   cumind_tree_init_root : mcts.py:126-136  (root prediction softmax, expand, noise)
 *   cumind_select         : mcts.py:165-171  (Node.select_child / ucb_score, 51-76)
 *   cumind_expand         : mcts.py:181-196  (dynamics, prediction, softmax, Node.expand 78-89)
 *   cumind_backup         : mcts.py:199-203  (Node.backup 91-98; root twice)
 *   cumind_finalize       : mcts.py:143-155
 *   cumind_dirichlet      : mcts.py:218      (np.random.dirichlet; Philox4x32-10 here)        */
int cumind_tree_init_root(cumind_tree* tree, const cumind_net* net, const float* d_root_hidden,
                          const cumind_search_cfg* cfg, const double* d_noise, void* stream);
int cumind_select(cumind_tree* tree, const cumind_search_cfg* cfg, void* stream);
int cumind_expand(cumind_tree* tree, const cumind_net* net, const cumind_search_cfg* cfg,
                  float* d_leaf_value, void* stream);
int cumind_backup(cumind_tree* tree, const float* d_leaf_value, void* stream);
int cumind_finalize(cumind_tree* tree, const cumind_search_cfg* cfg, int32_t* d_out_visits,
                    double* d_out_probs, double* d_out_root_value, void* stream);
int cumind_dirichlet(double* d_noise, int32_t B, int32_t A, double alpha, uint64_t seed,
                     uint64_t tree_index_base, void* stream);

/* Train-step support (trainer.py:97-100, agent.py:48).
 * cumind_adamw_step: optax.adamw over a flat fp32 device blob, fused with the 1/world scaling of an
 *   all-reduced gradient (grad_scale); d_mask (uint8, may be NULL) is 0 for leaves that are not
 *   trained (BatchNorm running statistics).  `step` counts from 1.
 * cumind_adamw_step_dev: the same with the step counter in device memory (CUDA-graph replayable).
 * cumind_net_update_params_device: nnx.update(network, params) from a device blob; asynchronous
 *   (one device copy + gather kernels that rebuild the kernel-side layouts), no host round trip. */
int cumind_adamw_step(float* d_params, const float* d_grads, float* d_m, float* d_v, const uint8_t* d_mask,
                      size_t n, float lr, float b1, float b2, float eps, float weight_decay, int64_t step,
                      float grad_scale, void* stream);
int cumind_adamw_step_dev(float* d_params, const float* d_grads, float* d_m, float* d_v, const uint8_t* d_mask,
                          size_t n, float lr, float b1, float b2, float eps, float weight_decay,
                          int64_t* d_step /* device counter, incremented first */, float grad_scale, void* stream);
int cumind_net_update_params_device(cumind_net* net, const float* d_params, size_t n_params, void* stream);

/* Trainer.compute_losses + nnx.value_and_grad (trainer.py:88-95, 172-205) in one kernel for vector-observation networks
 * (hidden_dim 32 or 64, num_blocks <= 3, K = num_unroll_steps in 1..16; cumind_train_supported says whether a shape is
 * covered): K-step unroll, value MSE + policy cross-entropy at all K+1 states against the same targets, reward MSE at the
 * K unrolled states, divided by K+1 / K+1 / K.  d_params / d_grad are flat fp32 blobs in the layout above;
 * d_actions [B,K] int32, d_value_targets [B], d_policy_targets [B,A], d_reward_targets [B,K];
 * d_losses4 = {value_loss, policy_loss, reward_loss, total}.  d_work: cumind_train_work_bytes() bytes of scratch (per-CTA
 * partial gradients, summed in a fixed order: the result is reproducible).  Asynchronous on `stream`, graph-capturable. */
int    cumind_train_supported(const cumind_net_desc* desc, int32_t K);
size_t cumind_train_work_bytes(const cumind_net_desc* desc, int32_t B);
int    cumind_train_loss_grad(const cumind_net_desc* desc, const float* d_params, const float* d_obs, const int32_t* d_actions,
                              const float* d_value_targets, const float* d_policy_targets, const float* d_reward_targets,
                              int32_t B, int32_t K, void* d_work, size_t work_bytes, float* d_grad, float* d_losses4, void* stream);

/* Data-parallel train step: the all-reduce of the flat gradient fused into AdamW over NVLink peer memory (p2p_adamw.cu),
 * one process per GPU.  Every rank: cumind_p2p_create (allocates its gradient buffer and flag array, returns two CUDA IPC
 * handles as 128 bytes) -> exchange the handles (torch.distributed.all_gather_object) -> cumind_p2p_open(all handles in
 * rank order).  Each step: write the local gradient to cumind_p2p_grad_ptr (cumind_train_loss_grad can write there
 * directly), then cumind_p2p_allreduce_adamw on every rank: waits for all ranks' gradients (flags in peer memory), sums
 * them in rank order with P2P loads (every rank computes the identical sum), applies optax.adamw with the 1/world scaling
 * to the local parameters, and returns only when no peer still reads this rank's buffer.  d_step as in
 * cumind_adamw_step_dev (incremented first; doubles as the barrier epoch).  d_sum_out (may be NULL) receives the summed
 * gradient.  Graph-capturable; a peer that never arrives sets cumind_p2p_error instead of hanging (bounded spin). */
typedef struct cumind_p2p cumind_p2p;
int    cumind_p2p_create(size_t n, int32_t rank, int32_t world, cumind_p2p** out, void* h_handles128);
int    cumind_p2p_open(cumind_p2p* p2p, const void* h_all_handles /* world x 128 bytes */);
float* cumind_p2p_grad_ptr(cumind_p2p* p2p);
int    cumind_p2p_allreduce_adamw(cumind_p2p* p2p, float* d_params, float* d_m, float* d_v, const uint8_t* d_mask, size_t n,
                                  float lr, float b1, float b2, float eps, float weight_decay, int64_t* d_step,
                                  float* d_sum_out, void* stream);
int    cumind_p2p_reset(cumind_p2p* p2p);   /* after the step counter was set back: every rank, between two host barriers */
int    cumind_p2p_error(cumind_p2p* p2p);
int    cumind_p2p_destroy(cumind_p2p* p2p);

/* Replay: the sum tree of TreeBuffer (src/cumind/data/memory.py:190-299) in device memory.
 * A flat float64 array of 2*tree_size-1 nodes, tree_size = the power of two >= capacity (memory.py:204-208).
 * cumind_sumtree_update applies `_update(tree_idx[k], priority[k])` (memory.py:249-253, 212-217) for k = 0..n-1 IN
 *   ORDER, with the reference's float64 additions — including its extra `sum_tree[-1] += change` step from the
 *   root — so the array equals the reference's bit for bit.  Indices are TREE indices (leaf = data + tree_size-1).
 * cumind_sumtree_sample is the retrieval of TreeBuffer.sample (memory.py:219-231, 265-273) for a whole batch:
 *   s_i = random.uniform(segment*i, segment*(i+1)) with d_uniform[i] in [0,1) standing for random.random();
 *   writes the tree index and the data index (tree index - tree_size + 1) of every sample.
 * cumind_sumtree_read copies the array (device to device) for inspection. */
typedef struct cumind_sumtree cumind_sumtree;
int cumind_sumtree_create(int64_t capacity, cumind_sumtree** out);
void cumind_sumtree_destroy(cumind_sumtree* tree);
int64_t cumind_sumtree_tree_size(const cumind_sumtree* tree);
int cumind_sumtree_clear(cumind_sumtree* tree, void* stream);
int cumind_sumtree_update(cumind_sumtree* tree, const int64_t* d_tree_idx, const double* d_priority, int64_t n, void* stream);
int cumind_sumtree_sample(const cumind_sumtree* tree, const double* d_uniform, int64_t batch, int64_t* d_tree_idx,
                          int64_t* d_data_idx, void* stream);
int cumind_sumtree_read(const cumind_sumtree* tree, double* d_out, void* stream);

/* ---- trajectory records of the replay buffers in device memory ------------------------------------------------------
 * The reference keeps trajectories (lists of step dicts) in a deque (memory.py:13-85) and Trainer._prepare_batch
 * (trainer.py:118-170) reads, of every sampled trajectory: step 0's observation and policy, the first K =
 * num_unroll_steps actions and rewards (zero padded), the discounted sum of its first td_steps rewards and — if it is
 * longer than td_steps — the observation at td_steps for the bootstrap value.  A record is exactly that, fixed when
 * the trajectory is stored; `capacity` records form a ring (slot = number of earlier adds mod capacity).
 * Record layout (cumind_replay_record_bytes, a multiple of 16):
 *     [0,8)   float64  sum_{i < min(len, td_steps)} reward_i * discount**i     (computed by the caller, in its arithmetic)
 *     [8,12)  int32    trajectory length (0 = empty: flagged by the gather)
 *     [12,16) int32    1 if the trajectory is longer than td_steps (the bootstrap observation is valid)
 *     float32 observation_0[obs_size] | float32 bootstrap observation[obs_size] | float32 policy_0[A]
 *     float32 rewards[K] | int32 actions[K]
 * cumind_replay_write copies n packed records (host or device memory) to slots (slot0 + i) mod capacity.
 * cumind_replay_gather builds a batch: row b comes from slot (head + d_index[b]) mod capacity, i.e. element d_index[b]
 * of the reference's deque when `head` is the slot of its oldest element; an index outside [0, n_valid) counts in
 * d_flags[0] (the reference's IndexError) and reads slot `head`; an empty trajectory counts in d_flags[1] (the
 * reference's RuntimeError).  d_flags is int32[2], accumulated (the caller zeroes it).
 * cumind_replay_values: values[b] = float32(partial[b] + discount_pow_n * float64(boot_value[b])) where has_boot[b],
 * else float32(partial[b]) — the n-step return target of trainer.py:172-205. */
typedef struct cumind_replay cumind_replay;
int cumind_replay_create(int64_t capacity, int64_t obs_size, int32_t n_actions, int32_t unroll_steps, cumind_replay** out);
void cumind_replay_destroy(cumind_replay* ring);
int64_t cumind_replay_record_bytes(const cumind_replay* ring);
int cumind_replay_write(cumind_replay* ring, int64_t slot0, const void* records, int64_t n, void* stream);
int cumind_replay_gather(const cumind_replay* ring, const int64_t* d_index, int64_t head, int64_t n_valid, int32_t B,
                         float* d_obs, float* d_boot_obs, float* d_policy, float* d_rewards, int32_t* d_actions,
                         double* d_partial, int32_t* d_has_boot, int32_t* d_flags, void* stream);
int cumind_replay_values(const double* d_partial, const int32_t* d_has_boot, const float* d_boot_value,
                         double discount_pow_n, int32_t B, float* d_values, void* stream);

/* Diagnostic: checks the table-driven small-divisor division used by the fused select kernel
 * (a/d from RN(1/d) with two FMA residual corrections) against the IEEE division on the device,
 * bitwise, over 148*4*256*n_per_thread random and structured numerators and divisors 1..max_divisor.
 * d_counts2 is a zero-initialised device uint64[2]: {mismatches, cases}. */
int cumind_selftest_division(uint64_t n_per_thread, uint64_t seed, int32_t max_divisor,
                             uint64_t* d_counts2, void* stream);

/* Counters: kernels launched by this library since load (bench.py reports the delta). */
uint64_t cumind_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CUMIND_B200_H */
