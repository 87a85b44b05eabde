// cumind_ffi.cc — XLA FFI handlers (jax.ffi custom calls) over the C ABI of include/cumind_b200.h.
//
// north_star: "Python host code calling hand-written sm_100a CUDA through thin jax.ffi C-ABI custom calls".  Each handler
// is a few lines: unpack XLA buffers, forward to the extern "C" entry point on the stream XLA hands over, map the return
// code.  Build against jaxlib's headers (python -c "import jax; print(jax.ffi.include_dir())") and link libcumind_b200.so:
//     g++ -std=c++17 -shared -fPIC -I$(jax ffi include dir) -I../../include cumind_ffi.cc -L../_lib -lcumind_b200 -o libcumind_ffi.so
// jax / jaxlib are NOT installed in this image (no network), so this file is compile-guarded: without the XLA headers it
// builds to an empty translation unit, and cumind_b200/ffi/jax_ffi.py raises a clear error instead of importing jax.
// Handles (cumind_net*, cumind_tree*) travel as int64 attributes; they are created with the plain host calls
// cumind_net_create / cumind_tree_create (the tree slab can be any XLA device buffer: the call allocates nothing).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define CUMIND_HAVE_XLA_FFI 1
#endif
#endif

#ifdef CUMIND_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "cumind_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error status(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(cumind_last_error()); }

// CuMindNetwork.initial_inference (network.py:288-300): obs [B,...] -> hidden [B,H], logits [B,A], value [B]
static ffi::Error InitialImpl(cudaStream_t stream, int64_t net, int32_t net_mode, ffi::Buffer<ffi::F32> obs,
                              ffi::ResultBuffer<ffi::F32> hidden, ffi::ResultBuffer<ffi::F32> logits, ffi::ResultBuffer<ffi::F32> value) {
    return status(cumind_initial_inference(reinterpret_cast<const cumind_net*>(net), obs.typed_data(), (int32_t)obs.dimensions()[0],
                                           hidden->typed_data(), logits->typed_data(), value->typed_data(), net_mode, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CumindInitialInference, InitialImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("net").Attr<int32_t>("net_mode")
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// CuMindNetwork.recurrent_inference (network.py:302-315)
static ffi::Error RecurrentImpl(cudaStream_t stream, int64_t net, int32_t net_mode, ffi::Buffer<ffi::F32> hidden, ffi::Buffer<ffi::S32> action,
                                ffi::ResultBuffer<ffi::F32> next, ffi::ResultBuffer<ffi::F32> reward, ffi::ResultBuffer<ffi::F32> logits,
                                ffi::ResultBuffer<ffi::F32> value) {
    return status(cumind_recurrent_inference(reinterpret_cast<const cumind_net*>(net), hidden.typed_data(), action.typed_data(),
                                             (int32_t)hidden.dimensions()[0], next->typed_data(), reward->typed_data(), logits->typed_data(),
                                             value->typed_data(), net_mode, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CumindRecurrentInference, RecurrentImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("net").Attr<int32_t>("net_mode")
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// MCTS.search for B trees (mcts.py:115-155): root hidden [B,H] (+ noise [B,A] f64) -> visits [B,A] i32, probs [B,A] f64, root value [B] f64
static ffi::Error SearchImpl(cudaStream_t stream, int64_t tree, int64_t net, int32_t num_simulations, int32_t action_space_size, double c_puct,
                             double exploration_fraction, double dirichlet_alpha, int32_t noise_mode, int32_t net_mode, int32_t tree_mode,
                             double discount, int64_t seed, int64_t tree_index_base, ffi::Buffer<ffi::F32> root_hidden, ffi::Buffer<ffi::F64> noise,
                             ffi::ResultBuffer<ffi::S32> visits, ffi::ResultBuffer<ffi::F64> probs, ffi::ResultBuffer<ffi::F64> root_value) {
    cumind_search_cfg cfg = {};
    cfg.num_simulations = num_simulations;
    cfg.action_space_size = action_space_size;
    cfg.c_puct = c_puct;
    cfg.exploration_fraction = exploration_fraction;
    cfg.dirichlet_alpha = dirichlet_alpha;
    cfg.noise_mode = noise_mode;
    cfg.net_mode = net_mode;
    cfg.tree_mode = tree_mode;
    cfg.discount = discount;
    cfg.seed = (uint64_t)seed;
    cfg.tree_index_base = (uint64_t)tree_index_base;
    return status(cumind_search(reinterpret_cast<cumind_tree*>(tree), reinterpret_cast<const cumind_net*>(net), root_hidden.typed_data(), &cfg,
                                noise_mode == 1 ? noise.typed_data() : nullptr, visits->typed_data(), probs->typed_data(),
                                root_value->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CumindSearch, SearchImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("tree").Attr<int64_t>("net")
                                  .Attr<int32_t>("num_simulations").Attr<int32_t>("action_space_size").Attr<double>("c_puct")
                                  .Attr<double>("exploration_fraction").Attr<double>("dirichlet_alpha").Attr<int32_t>("noise_mode")
                                  .Attr<int32_t>("net_mode").Attr<int32_t>("tree_mode").Attr<double>("discount").Attr<int64_t>("seed")
                                  .Attr<int64_t>("tree_index_base").Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());
#endif  // CUMIND_HAVE_XLA_FFI
