// abi.cu — the C ABI of include/cumind_b200.h: handles, validation, kernel selection.
#include <atomic>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"

using namespace cm;

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct cumind_net {
    cumind_net_desc desc;
    int device;
    int sm_count;
    size_t n_params;
    float* d_blob;
    float* d_pack;
    float* d_fpack;       // FusedPackLayout copy for the fused search kernel
    float* d_team_pack;   // TeamPackLayout copy for the team search kernel
    TeamPackLayout team_pk;
    int* d_team_map;
    unsigned char* d_tcpack;  // TcPackLayout (bf16 tiles) for the tensor-core kernels, or null
    TcPackLayout tpk;
    bool tc_ok;
    unsigned char* d_cpack;   // ConvTcPackLayout (bf16 conv tiles) or null
    ConvTcPackLayout cpk;
    bool conv_tc_ok;
    BlobLayout lo;
    PackLayout pk;
    FusedPackLayout fpk;
    float* d_work;        // activation workspace of the image encoder (lazily sized, grow-only)
    size_t work_bytes;
    int *d_pack_map, *d_fpack_map, *d_tc_half_map, *d_tc_f32_map, *d_c_half_map;
    BnFold* d_folds;
    int n_folds;
};

struct cumind_tree {
    TreeView tv;
    float* leaf_value;  // [B] scratch of the stepwise path
    size_t bytes;
};

static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};

static int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}
static int fail_cuda(const char* what, cudaError_t e) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 2;
}
#define CM_CUDA(call)                                         \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return fail_cuda(#call, _e);   \
    } while (0)

extern "C" const char* cumind_last_error(void) { return g_err.c_str(); }
extern "C" int cumind_abi_version(void) { return CUMIND_ABI_VERSION; }
extern "C" uint64_t cumind_launch_count(void) { return g_launches.load(); }

extern "C" size_t cumind_param_count(const cumind_net_desc* desc) {
    if (!desc_valid(desc)) return 0;
    return make_blob_layout(desc).total;
}

static NetView view_of(const cumind_net* n) {
    NetView v;
    v.desc = n->desc;
    v.blob = n->d_blob;
    v.lo = n->lo;
    v.pack = n->d_pack;
    v.pk = n->pk;
    return v;
}

// ------------------------------------------------------------------------------------------------
// Kernel-side parameter layouts are rebuilt ON THE DEVICE from the canonical blob: at create time the
// host only computes index maps (destination element -> blob index, -1 = zero padding, -2 = leave
// alone) and BatchNorm-fold records; every parameter update is then a blob copy plus a few gather
// kernels on the caller's stream (no host round trip, capturable in a CUDA graph).
// ------------------------------------------------------------------------------------------------
static void build_pack_map(const cumind_net* n, std::vector<int>& m) {
    const PackLayout& pk = n->pk;
    const BlobLayout& lo = n->lo;
    const int H = pk.H, L = pk.L, A = pk.A, HP = pk.HP;
    m.assign(pk.total, -1);
    for (size_t i = 0; i < (size_t)A * H; ++i) m[pk.emb + i] = (int)(lo.emb + i);
    for (size_t i = 0; i < (size_t)L * ((size_t)H * H + H); ++i) m[pk.layers + i] = (int)(lo.dyn + i);
    for (int k = 0; k < H; ++k) {
        for (int a = 0; a < A; ++a) m[pk.heads_w + (size_t)k * HP + a] = (int)(lo.pol_k + (size_t)k * A + a);
        m[pk.heads_w + (size_t)k * HP + A] = (int)(lo.val_k + k);
        m[pk.heads_w + (size_t)k * HP + A + 1] = (int)(lo.rew_k + k);
    }
    for (int a = 0; a < A; ++a) m[pk.heads_b + a] = (int)(lo.pol_b + a);
    m[pk.heads_b + A] = (int)lo.val_b;
    m[pk.heads_b + A + 1] = (int)lo.rew_b;
}

// FusedPackLayout: lane-blocked layer kernels, row-major heads
static void build_fused_map(const cumind_net* n, std::vector<int>& m) {
    const FusedPackLayout& pk = n->fpk;
    const BlobLayout& lo = n->lo;
    const int H = pk.H, L = pk.L, A = pk.A, R = pk.R > 0 ? pk.R : 1;
    m.assign(pk.total, -1);
    for (size_t i = 0; i < (size_t)A * H; ++i) m[pk.emb + i] = (int)(lo.emb + i);
    for (int l = 0; l < L; ++l) {
        const size_t K = lo.dyn + (size_t)l * ((size_t)H * H + H);
        if (pk.R > 0) {
            for (int k = 0; k < H; ++k)
                for (int lane = 0; lane < 32; ++lane)
                    for (int v = 0; v < R; ++v) {
                        const int j = lane * R + v;
                        if (j < H) m[pk.layer_w(l) + FusedPackLayout::w_index(k, lane, v, R)] = (int)(K + (size_t)k * H + j);
                    }
        }
        for (int j = 0; j < H; ++j) m[pk.layer_b(l) + j] = (int)(K + (size_t)H * H + j);
    }
    for (int k = 0; k < H; ++k) {
        for (int a = 0; a < A; ++a) m[pk.heads_w + (size_t)k * pk.HP + a] = (int)(lo.pol_k + (size_t)k * A + a);
        m[pk.heads_w + (size_t)k * pk.HP + A] = (int)(lo.val_k + k);
        m[pk.heads_w + (size_t)k * pk.HP + A + 1] = (int)(lo.rew_k + k);
    }
    for (int a = 0; a < A; ++a) m[pk.heads_b + a] = (int)(lo.pol_b + a);
    m[pk.heads_b + A] = (int)lo.val_b;
    m[pk.heads_b + A + 1] = (int)lo.rew_b;
}

// TeamPackLayout: k-major layer kernels padded to Hp output columns, every weight stored twice when dup
static void build_team_map(const cumind_net* n, std::vector<int>& m) {
    const TeamPackLayout& pk = n->team_pk;
    const BlobLayout& lo = n->lo;
    const int H = pk.H, L = pk.L, A = pk.A, Hp = pk.Hp, rep = pk.dup ? 2 : 1;
    m.assign(pk.total, -1);
    for (size_t i = 0; i < (size_t)A * H; ++i) m[pk.emb + i] = (int)(lo.emb + i);
    for (int l = 0; l < L; ++l) {
        const size_t K = lo.dyn + (size_t)l * ((size_t)H * H + H);
        for (int k = 0; k < H; ++k)
            for (int j = 0; j < H; ++j)
                for (int r = 0; r < rep; ++r) m[pk.layer_w(l) + ((size_t)k * Hp + j) * rep + r] = (int)(K + (size_t)k * H + j);
        for (int j = 0; j < H; ++j) m[pk.layer_b(l) + j] = (int)(K + (size_t)H * H + j);
    }
    for (int k = 0; k < H; ++k) {
        for (int a = 0; a < A; ++a) m[pk.heads_w + (size_t)k * pk.HP + a] = (int)(lo.pol_k + (size_t)k * A + a);
        m[pk.heads_w + (size_t)k * pk.HP + A] = (int)(lo.val_k + k);
        m[pk.heads_w + (size_t)k * pk.HP + A + 1] = (int)(lo.rew_k + k);
    }
    for (int a = 0; a < A; ++a) m[pk.heads_b + a] = (int)(lo.pol_b + a);
    m[pk.heads_b + A] = (int)lo.val_b;
    m[pk.heads_b + A + 1] = (int)lo.rew_b;
    for (int k = 0; k < H; ++k) {
        for (int a = 0; a < A; ++a) m[pk.heads_wt + (size_t)a * H + k] = (int)(lo.pol_k + (size_t)k * A + a);
        m[pk.heads_wt + (size_t)A * H + k] = (int)(lo.val_k + k);
        m[pk.heads_wt + (size_t)(A + 1) * H + k] = (int)(lo.rew_k + k);
    }
}

// TcPackLayout: bf16 tiles over [0, bias_off) (map in 2-byte units), fp32 over [bias_off, total) (4-byte units)
static void build_tc_maps(const cumind_net* n, std::vector<int>& half_map, std::vector<int>& f32_map) {
    const TcPackLayout& pk = n->tpk;
    const BlobLayout& lo = n->lo;
    const int H = pk.H, L = pk.L, A = pk.A, NP = pk.NP;
    half_map.assign(pk.bias_off / 2, -1);
    f32_map.assign((pk.total_bytes - pk.bias_off) / 4, -1);
    auto f32_at = [&](size_t byte_off) -> int& { return f32_map[(byte_off - pk.bias_off) / 4]; };
    for (int l = 0; l < L; ++l) {
        const size_t K = lo.dyn + (size_t)l * ((size_t)H * H + H);
        const size_t tile = pk.layer_off + (size_t)l * H * H * 2;
        for (int k = 0; k < H; ++k)
            for (int j = 0; j < H; ++j) half_map[(tile + tc_tile_offset_host(j, k, H)) / 2] = (int)(K + (size_t)k * H + j);
        for (int j = 0; j < H; ++j) f32_at(pk.bias_off + ((size_t)l * H + j) * 4) = (int)(K + (size_t)H * H + j);
    }
    for (int k = 0; k < H; ++k) {
        for (int a = 0; a < A; ++a) half_map[(pk.heads_off + tc_tile_offset_host(a, k, NP)) / 2] = (int)(lo.pol_k + (size_t)k * A + a);
        half_map[(pk.heads_off + tc_tile_offset_host(A, k, NP)) / 2] = (int)(lo.val_k + k);
        half_map[(pk.heads_off + tc_tile_offset_host(A + 1, k, NP)) / 2] = (int)(lo.rew_k + k);
    }
    for (int a = 0; a < A; ++a) f32_at(pk.bh_off + (size_t)a * 4) = (int)(lo.pol_b + a);
    f32_at(pk.bh_off + (size_t)A * 4) = (int)lo.val_b;
    f32_at(pk.bh_off + (size_t)(A + 1) * 4) = (int)lo.rew_b;
    for (size_t i = 0; i < (size_t)A * H; ++i) f32_at(pk.emb_off + i * 4) = (int)(lo.emb + i);
}

// ConvTcPackLayout: bf16 tiles (map in 2-byte units over the whole pack, -2 over the fp32 gain/shift
// regions) + one BatchNorm-fold record per output channel of every conv layer
static void build_conv_maps(const cumind_net* n, std::vector<int>& half_map, std::vector<BnFold>& folds) {
    const ConvTcPackLayout& pk = n->cpk;
    const int C = n->desc.obs_shape[2], Cc = n->desc.conv_channels, L = n->desc.num_blocks;
    half_map.assign(pk.total_bytes / 2, -1);
    folds.clear();
    size_t p = n->lo.enc;
    auto emit = [&](int i, size_t K, size_t bias, long long bn, int cin) {
        const size_t tile = pk.off[i];
        for (int k = 0; k < 9 * cin; ++k)
            for (int co = 0; co < Cc; ++co) half_map[(tile + tc_tile_offset_host(co, k, Cc)) / 2] = (int)(K + (size_t)k * Cc + co);
        const size_t goff = tile + pk.gain_off[i];
        for (size_t h = goff / 2; h < (goff + (size_t)2 * Cc * 4) / 2; ++h) half_map[h] = -2;
        for (int co = 0; co < Cc; ++co) {
            BnFold f;
            f.gain_dst = (int)((goff + (size_t)co * 4) / 4);
            f.shift_dst = (int)((goff + (size_t)(Cc + co) * 4) / 4);
            f.bias = (int)(bias + co);
            f.scale = bn >= 0 ? (int)(bn + co) : -1;
            f.beta = bn >= 0 ? (int)(bn + Cc + co) : -1;
            f.mean = bn >= 0 ? (int)(bn + 2 * Cc + co) : -1;
            f.var = bn >= 0 ? (int)(bn + 3 * Cc + co) : -1;
            folds.push_back(f);
        }
    };
    emit(0, p, p + (size_t)9 * C * Cc, -1, C);
    p += (size_t)9 * C * Cc + Cc;
    for (int l = 0; l < L; ++l) {
        const size_t k1 = p, b1 = k1 + (size_t)9 * Cc * Cc, bn1 = b1 + Cc, k2 = bn1 + 4 * Cc, b2 = k2 + (size_t)9 * Cc * Cc, bn2 = b2 + Cc;
        emit(1 + 2 * l, k1, b1, (long long)bn1, Cc);
        emit(2 + 2 * l, k2, b2, (long long)bn2, Cc);
        p = bn2 + 4 * Cc;
    }
}

template <class T>
static cudaError_t upload(const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)d, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// rebuild every kernel-side layout from n->d_blob on `st`
static int repack_device(cumind_net* n, cudaStream_t st) {
    CM_CUDA(launch_repack_f32(n->d_blob, n->d_pack_map, n->d_pack, n->pk.total, st));
    CM_CUDA(launch_repack_f32(n->d_blob, n->d_fpack_map, n->d_fpack, n->fpk.total, st));
    CM_CUDA(launch_repack_f32(n->d_blob, n->d_team_map, n->d_team_pack, n->team_pk.total, st));
    g_launches += 3;
    if (n->tc_ok) {
        CM_CUDA(launch_repack_bf16(n->d_blob, n->d_tc_half_map, n->d_tcpack, n->tpk.bias_off / 2, st));
        CM_CUDA(launch_repack_f32(n->d_blob, n->d_tc_f32_map, reinterpret_cast<float*>(n->d_tcpack + n->tpk.bias_off),
                                  (n->tpk.total_bytes - n->tpk.bias_off) / 4, st));
        g_launches += 2;
    }
    if (n->conv_tc_ok) {
        CM_CUDA(launch_repack_bf16(n->d_blob, n->d_c_half_map, n->d_cpack, n->cpk.total_bytes / 2, st));
        CM_CUDA(launch_bn_fold(n->d_blob, n->d_folds, n->n_folds, reinterpret_cast<float*>(n->d_cpack), st));
        g_launches += 2;
    }
    return 0;
}

extern "C" int cumind_net_create(const cumind_net_desc* desc, const float* h_params, size_t n_params, cumind_net** out) {
    if (!out) return fail("cumind_net_create: out is NULL");
    *out = nullptr;
    if (!desc_valid(desc)) {
        if (desc && desc->obs_ndim != 1 && desc->obs_ndim != 3)
            return fail("Unsupported observation shape: observation must be 1-D or 3-D");  // network.py:179-180
        return fail("cumind_net_create: invalid descriptor");
    }
    const BlobLayout lo = make_blob_layout(desc);
    if (n_params != lo.total || !h_params) return fail("cumind_net_create: parameter blob has the wrong size");
    cumind_net* n = new (std::nothrow) cumind_net();
    if (!n) return fail("cumind_net_create: out of host memory");
    n->desc = *desc;
    n->lo = lo;
    n->pk = make_pack_layout(desc->hidden_dim, desc->num_blocks, desc->action_space_size);
    n->fpk = make_fused_pack_layout(desc->hidden_dim, desc->num_blocks, desc->action_space_size);
    {
        n->team_pk = make_team_pack_layout(desc->hidden_dim, desc->num_blocks, desc->action_space_size, 0);
    }
    n->n_params = n_params;
    n->d_team_pack = nullptr;
    n->d_team_map = nullptr;
    n->d_blob = n->d_pack = n->d_fpack = n->d_work = nullptr;
    n->d_tcpack = nullptr;
    n->d_cpack = nullptr;
    n->conv_tc_ok = conv_tc_supported(*desc) && desc->num_blocks <= 8;
    if (n->conv_tc_ok) n->cpk = make_conv_tc_pack_layout(*desc);
    n->tc_ok = tc_supported(desc->hidden_dim, desc->num_blocks, desc->action_space_size);
    n->tpk = make_tc_pack_layout(desc->hidden_dim, desc->num_blocks, desc->action_space_size);
    n->work_bytes = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&n->device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&n->sm_count, cudaDevAttrMultiProcessorCount, n->device)) != cudaSuccess ||
        (e = cudaMalloc(&n->d_blob, sizeof(float) * lo.total)) != cudaSuccess ||
        (e = cudaMalloc(&n->d_pack, sizeof(float) * n->pk.total)) != cudaSuccess ||
        (e = cudaMalloc(&n->d_fpack, sizeof(float) * n->fpk.total)) != cudaSuccess ||
        (e = cudaMalloc(&n->d_team_pack, sizeof(float) * n->team_pk.total)) != cudaSuccess ||
        (n->tc_ok && (e = cudaMalloc(&n->d_tcpack, n->tpk.total_bytes)) != cudaSuccess) ||
        (n->conv_tc_ok && (e = cudaMalloc(&n->d_cpack, n->cpk.total_bytes)) != cudaSuccess)) {
        cudaFree(n->d_blob); cudaFree(n->d_pack); cudaFree(n->d_fpack); cudaFree(n->d_team_pack);
        delete n;
        return fail_cuda("cumind_net_create", e);
    }
    {
        std::vector<int> m;
        n->d_pack_map = n->d_fpack_map = n->d_tc_half_map = n->d_tc_f32_map = n->d_c_half_map = nullptr;
        n->d_folds = nullptr;
        n->n_folds = 0;
        build_pack_map(n, m);
        if ((e = upload(m, &n->d_pack_map)) != cudaSuccess) { cumind_net_destroy(n); return fail_cuda("cumind_net_create", e); }
        build_fused_map(n, m);
        if ((e = upload(m, &n->d_fpack_map)) != cudaSuccess) { cumind_net_destroy(n); return fail_cuda("cumind_net_create", e); }
        build_team_map(n, m);
        if ((e = upload(m, &n->d_team_map)) != cudaSuccess) { cumind_net_destroy(n); return fail_cuda("cumind_net_create", e); }
        if (n->tc_ok) {
            std::vector<int> f;
            build_tc_maps(n, m, f);
            if ((e = upload(m, &n->d_tc_half_map)) != cudaSuccess || (e = upload(f, &n->d_tc_f32_map)) != cudaSuccess) {
                cumind_net_destroy(n);
                return fail_cuda("cumind_net_create", e);
            }
        }
        if (n->conv_tc_ok) {
            std::vector<BnFold> folds;
            build_conv_maps(n, m, folds);
            n->n_folds = (int)folds.size();
            if ((e = upload(m, &n->d_c_half_map)) != cudaSuccess || (e = upload(folds, &n->d_folds)) != cudaSuccess) {
                cumind_net_destroy(n);
                return fail_cuda("cumind_net_create", e);
            }
        }
    }
    *out = n;
    int rc = cumind_net_update_params(n, h_params, n_params, nullptr);
    if (rc != 0) { cumind_net_destroy(n); *out = nullptr; return rc; }
    CM_CUDA(cudaStreamSynchronize(nullptr));
    return 0;
}

extern "C" int cumind_net_update_params(cumind_net* n, const float* h_params, size_t n_params, void* stream) {
    if (!n || !h_params) return fail("cumind_net_update_params: NULL argument");
    if (n_params != n->n_params) return fail("cumind_net_update_params: parameter blob has the wrong size");
    cudaStream_t st = (cudaStream_t)stream;
    CM_CUDA(cudaMemcpyAsync(n->d_blob, h_params, sizeof(float) * n_params, cudaMemcpyHostToDevice, st));
    if (repack_device(n, st)) return 1;
    CM_CUDA(cudaStreamSynchronize(st));  // the caller may reuse h_params right away
    return 0;
}

static int encoder_chunk(bool tc, int B) { return tc ? (B < 1024 ? B : 1024) : (B < 128 ? B : 128); }

extern "C" int cumind_net_reserve(cumind_net* n, int32_t max_batch) {
    if (!n || max_batch <= 0) return fail("cumind_net_reserve: bad argument");
    if (n->desc.obs_ndim != 3) return 0;
    size_t need = conv_encoder_work_bytes(n->desc, encoder_chunk(false, max_batch));
    if (n->conv_tc_ok) {
        const size_t t = conv_tc_work_bytes(n->desc, encoder_chunk(true, max_batch));
        if (t > need) need = t;
    }
    if (need <= n->work_bytes) return 0;
    CM_CUDA(cudaDeviceSynchronize());
    cudaFree(n->d_work);
    n->d_work = nullptr;
    n->work_bytes = 0;
    CM_CUDA(cudaMalloc(&n->d_work, need));
    n->work_bytes = need;
    return 0;
}

extern "C" int cumind_net_destroy(cumind_net* n) {
    if (!n) return 0;
    cudaFree(n->d_blob);
    cudaFree(n->d_pack);
    cudaFree(n->d_fpack);
    cudaFree(n->d_team_pack);
    cudaFree(n->d_team_map);
    cudaFree(n->d_tcpack);
    cudaFree(n->d_cpack);
    cudaFree(n->d_pack_map); cudaFree(n->d_fpack_map); cudaFree(n->d_tc_half_map); cudaFree(n->d_tc_f32_map);
    cudaFree(n->d_c_half_map); cudaFree(n->d_folds);
    cudaFree(n->d_work);
    delete n;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// network entry points
// ------------------------------------------------------------------------------------------------
static int check_mode(const cumind_net* n, int net_mode, bool tc_allowed) {
    if (net_mode == CUMIND_NET_FP32_EXACT) return 0;
    if (net_mode == CUMIND_NET_BF16_TC) {
        if (!tc_allowed) return fail("CUMIND_NET_BF16_TC is not available for this entry point");
        if (!n->tc_ok) return fail("CUMIND_NET_BF16_TC needs hidden_dim in {16,32,64,128} and action_space_size <= 14");
        return 0;
    }
    return fail("unknown net_mode");
}

extern "C" int cumind_prediction(const cumind_net* n, const float* d_hidden, int32_t B, float* d_logits, float* d_value,
                                 int32_t net_mode, void* stream) {
    if (!n || !d_hidden || B < 0) return fail("cumind_prediction: bad argument");
    if (check_mode(n, net_mode, true)) return 1;
    if (B == 0) return 0;
    if (net_mode == CUMIND_NET_BF16_TC)
        CM_CUDA(launch_recurrent_tc(n->d_tcpack, n->tpk, n->sm_count, d_hidden, nullptr, B, nullptr, nullptr, d_logits, d_value, 0,
                                    (cudaStream_t)stream));
    else
        CM_CUDA(launch_prediction(view_of(n), d_hidden, B, d_logits, d_value, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_initial_inference(const cumind_net* cn, const float* d_obs, int32_t B, float* d_hidden, float* d_logits,
                                        float* d_value, int32_t net_mode, void* stream) {
    cumind_net* n = const_cast<cumind_net*>(cn);
    if (!n || !d_obs || !d_hidden || B < 0) return fail("cumind_initial_inference: bad argument");
    if (check_mode(n, net_mode, true)) return 1;
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const NetView nv = view_of(n);
    if (n->desc.obs_ndim == 1) {
        CM_CUDA(launch_vector_encoder(nv, d_obs, B, d_hidden, st));
        g_launches += 1;
    } else {
        if ((n->desc.conv_channels & 3) != 0) return fail("image encoder needs conv_channels % 4 == 0");
        const bool tc = net_mode == CUMIND_NET_BF16_TC;
        if (tc && !n->conv_tc_ok) return fail("CUMIND_NET_BF16_TC image encoder needs conv_channels % 16 == 0 and <= 256");
        // 1024 images per pass: smaller passes that would keep the three activation buffers inside L2 measured slower
        // (0.91 ms at 512, 0.98 at 256 vs 0.85 at 1024: the layers are bound by per-tile latency, not by DRAM)
        const int chunk = encoder_chunk(tc, B);
        const size_t need = tc ? conv_tc_work_bytes(n->desc, chunk) : conv_encoder_work_bytes(n->desc, chunk);
        if (need > n->work_bytes)
            return fail("cumind_initial_inference: image-encoder workspace too small for this batch; call cumind_net_reserve(net, max_batch) first");
        if (tc)
            CM_CUDA(launch_conv_encoder_tc(nv, n->d_cpack, n->cpk, n->sm_count, d_obs, B, d_hidden, (unsigned char*)n->d_work, chunk, st));
        else
            CM_CUDA(launch_conv_encoder(nv, d_obs, B, d_hidden, n->d_work, chunk, st));
        g_launches += (unsigned long long)((B + chunk - 1) / chunk) * (2 + 2 * n->desc.num_blocks);
    }
    if (d_logits || d_value) {
        if (net_mode == CUMIND_NET_BF16_TC)
            CM_CUDA(launch_recurrent_tc(n->d_tcpack, n->tpk, n->sm_count, d_hidden, nullptr, B, nullptr, nullptr, d_logits, d_value, 0, st));
        else
            CM_CUDA(launch_prediction(nv, d_hidden, B, d_logits, d_value, st));
        g_launches += 1;
    }
    return 0;
}

extern "C" int cumind_recurrent_inference(const cumind_net* n, const float* d_hidden, const int32_t* d_action, int32_t B,
                                          float* d_next, float* d_reward, float* d_logits, float* d_value, int32_t net_mode,
                                          void* stream) {
    if (!n || !d_hidden || !d_action || B < 0) return fail("cumind_recurrent_inference: bad argument");
    if (check_mode(n, net_mode, true)) return 1;
    if (B == 0) return 0;
    if (net_mode == CUMIND_NET_BF16_TC)
        CM_CUDA(launch_recurrent_tc(n->d_tcpack, n->tpk, n->sm_count, d_hidden, d_action, B, d_next, d_reward, d_logits, d_value, 1,
                                    (cudaStream_t)stream));
    else
        CM_CUDA(launch_recurrent(view_of(n), d_hidden, d_action, B, d_next, d_reward, d_logits, d_value, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_softmax(const float* d_logits, int32_t B, int32_t A, float* d_probs, void* stream) {
    if (!d_logits || !d_probs || B < 0 || A <= 0) return fail("cumind_softmax: bad argument");
    if (B == 0) return 0;
    CM_CUDA(launch_softmax(d_logits, B, A, d_probs, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// tree slab
// ------------------------------------------------------------------------------------------------
struct SlabOffsets {
    size_t nodes, eidx, rprior, exp_node, hidden, path, path_len, leaf, n_exp, depth, leaf_value, reward, minmax, total;
};
static SlabOffsets slab_offsets(int64_t B, int64_t S, int64_t A, int64_t H) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t N = 1 + A * (S + 1);
    SlabOffsets o;
    size_t p = 0;
    o.nodes = p;      p = al(p + B * N * sizeof(Node));
    o.eidx = p;       p = al(p + B * N * sizeof(int32_t));
    o.rprior = p;     p = al(p + B * A * sizeof(double));
    o.exp_node = p;   p = al(p + B * (S + 1) * sizeof(int32_t));
    o.hidden = p;     p = al(p + B * (S + 1) * H * sizeof(float));
    o.path = p;       p = al(p + B * (S + 2) * sizeof(int32_t));
    o.path_len = p;   p = al(p + B * sizeof(int32_t));
    o.leaf = p;       p = al(p + B * sizeof(int32_t));
    o.n_exp = p;      p = al(p + B * sizeof(int32_t));
    o.depth = p;      p = al(p + B * sizeof(long long));
    o.leaf_value = p; p = al(p + B * sizeof(float));
    o.reward = p;     p = al(p + B * N * sizeof(float));
    o.minmax = p;     p = al(p + B * 2 * sizeof(double));
    o.total = p;
    return o;
}

extern "C" size_t cumind_tree_bytes(int32_t B, int32_t S_max, int32_t A_tree, int32_t H) {
    if (B <= 0 || S_max <= 0 || A_tree <= 0 || H <= 0) return 0;
    return slab_offsets(B, S_max, A_tree, H).total;
}

extern "C" int cumind_tree_create(int32_t B, int32_t S_max, int32_t A_tree, int32_t H, void* d_slab, size_t slab_bytes,
                                  cumind_tree** out) {
    if (!out) return fail("cumind_tree_create: out is NULL");
    *out = nullptr;
    if (B <= 0 || S_max <= 0 || A_tree <= 0 || H <= 0) return fail("cumind_tree_create: sizes must be positive");
    const SlabOffsets o = slab_offsets(B, S_max, A_tree, H);
    if (!d_slab || slab_bytes < o.total) return fail("cumind_tree_create: slab too small (see cumind_tree_bytes)");
    if (((uintptr_t)d_slab & 255) != 0) return fail("cumind_tree_create: slab must be 256-byte aligned");
    if ((int64_t)1 + (int64_t)A_tree * (S_max + 1) > 0x7fffffff) return fail("cumind_tree_create: tree too large");
    cumind_tree* t = new (std::nothrow) cumind_tree();
    if (!t) return fail("cumind_tree_create: out of host memory");
    unsigned char* b = (unsigned char*)d_slab;
    t->tv.B = B; t->tv.S_max = S_max; t->tv.A = A_tree; t->tv.H = H; t->tv.N = 1 + A_tree * (S_max + 1);
    t->tv.nodes = (Node*)(b + o.nodes);
    t->tv.child_eidx = (int32_t*)(b + o.eidx);
    t->tv.root_prior = (double*)(b + o.rprior);
    t->tv.exp_node = (int32_t*)(b + o.exp_node);
    t->tv.hidden = (float*)(b + o.hidden);
    t->tv.path = (int32_t*)(b + o.path);
    t->tv.path_len = (int32_t*)(b + o.path_len);
    t->tv.leaf = (int32_t*)(b + o.leaf);
    t->tv.n_expanded = (int32_t*)(b + o.n_exp);
    t->tv.depth_sum = (long long*)(b + o.depth);
    t->leaf_value = (float*)(b + o.leaf_value);
    t->tv.reward = (float*)(b + o.reward);
    t->tv.minmax = (double*)(b + o.minmax);
    t->bytes = o.total;
    *out = t;
    return 0;
}

extern "C" int cumind_tree_destroy(cumind_tree* t) {
    delete t;
    return 0;
}

extern "C" int cumind_tree_get_view(const cumind_tree* t, cumind_tree_view* out) {
    if (!t || !out) return fail("cumind_tree_get_view: NULL argument");
    out->B = t->tv.B; out->S_max = t->tv.S_max; out->A = t->tv.A; out->H = t->tv.H; out->N = t->tv.N;
    out->nodes = (cumind_node*)t->tv.nodes;
    out->child_eidx = t->tv.child_eidx;
    out->root_prior = t->tv.root_prior;
    out->exp_node = t->tv.exp_node;
    out->hidden = t->tv.hidden;
    out->path = t->tv.path;
    out->path_len = t->tv.path_len;
    out->leaf = t->tv.leaf;
    out->n_expanded = t->tv.n_expanded;
    out->depth_sum = (int64_t*)t->tv.depth_sum;
    out->reward = t->tv.reward;
    out->minmax = t->tv.minmax;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------------
static int check_search(const char* fn, const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c, const double* d_noise) {
    if (!t || !n || !c) return fail(std::string(fn) + ": NULL argument");
    if (c->num_simulations <= 0) return fail("num_simulations must be positive");      // config.py:160-162
    if (c->action_space_size <= 0) return fail("action_space_size must be positive");  // config.py:143-145
    if (c->num_simulations > t->tv.S_max) return fail(std::string(fn) + ": num_simulations exceeds the tree's S_max");
    const int A_tree = c->action_space_size < n->desc.action_space_size ? c->action_space_size : n->desc.action_space_size;
    if (A_tree != t->tv.A) return fail(std::string(fn) + ": tree was created for a different action count");
    if (n->desc.hidden_dim != t->tv.H) return fail(std::string(fn) + ": tree was created for a different hidden_dim");
    if (c->noise_mode < 0 || c->noise_mode > 2) return fail("noise_mode must be 0, 1 or 2");
    if (c->noise_mode == 1 && !d_noise) return fail("noise_mode 1 needs d_noise");
    if (c->tree_mode != CUMIND_TREE_REFERENCE && c->tree_mode != CUMIND_TREE_MINMAX_DISCOUNT) return fail("unknown tree_mode");
    if (c->tree_mode == CUMIND_TREE_MINMAX_DISCOUNT && c->net_mode != CUMIND_NET_FP32_EXACT)
        return fail("CUMIND_TREE_MINMAX_DISCOUNT runs in CUMIND_NET_FP32_EXACT mode only");
    if (check_mode(n, c->net_mode, true)) return 1;
    return 0;
}

static StepParams step_params(const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c) {
    StepParams p;
    p.tv = t->tv;
    p.pack = n->d_pack;
    p.pk = n->pk;
    p.S = c->num_simulations;
    p.A_cfg = c->action_space_size;
    p.A_net = n->desc.action_space_size;
    p.c_puct = c->c_puct;
    p.frac = c->exploration_fraction;
    p.one_minus_frac = 1.0 - c->exploration_fraction;  // (1 - f) in float64, mcts.py:222
    p.alpha = c->dirichlet_alpha;
    p.noise_mode = c->noise_mode;
    p.seed = c->seed;
    p.tree_base = c->tree_index_base;
    p.tree_mode = c->tree_mode;
    p.discount = c->discount;
    return p;
}

struct FusedPlan { int G, R, grid, block; size_t smem; FusedParams p; };

static int plan_fused(const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c, FusedPlan& pl, std::string& why) {
    const int H = n->desc.hidden_dim, B = t->tv.B;
    const int sm = n->sm_count;
    pl.R = n->fpk.R;
    if (pl.R <= 0) { why = "hidden_dim has no lane mapping"; return 1; }
    if (t->tv.A > 65535 || t->tv.S_max > 65533) { why = "tree too large for the packed select payload"; return 1; }
    // lanes per tree: measured optimum on B200 at H=64 (profiles/README.md, sweep over B x G):
    // <= 16 trees per SM -> 16 lanes, <= 64 -> 8, more -> 4 (enough warps to hide the per-tree
    // dependency chain without paying the per-warp select overhead for too few trees).
    int G;
    if (c->lanes_per_tree > 0) {
        G = 4;
        while (G * 2 <= c->lanes_per_tree && G < 32) G *= 2;
    } else {
        const int tps = (B + sm - 1) / sm;
        G = tps <= 16 ? 16 : (tps <= 64 ? 8 : 4);
    }
    pl.G = G;
    const int TW = 32 / G;
    const FusedPackLayout& pk = n->fpk;
    const size_t pack_bytes = pk.total * sizeof(float);
    const int n_sqrt = 2 * t->tv.S_max + 4;
    const size_t off_sqrt = (pack_bytes + 15) & ~(size_t)15;
    const size_t off_recip = (off_sqrt + (size_t)n_sqrt * sizeof(double) + 15) & ~(size_t)15;
    const size_t off_slots = (off_recip + (size_t)n_sqrt * sizeof(double) + 15) & ~(size_t)15;
    const size_t warp_base = ((size_t)(2 * H * TW + 2 * TW * pk.HP) * sizeof(float) + 15) & ~(size_t)15;
    const size_t path_bytes = ((size_t)TW * (t->tv.S_max + 2) * sizeof(int) + 15) & ~(size_t)15;
    const size_t kMax = 227 * 1024;
    // warps per block: one CTA per SM when the whole batch fits in <= 16 warps, else 16 warps per CTA
    const int warps_total = (B + TW - 1) / TW;
    int wps = (warps_total + sm - 1) / sm;
    int nw = wps <= 16 ? wps : 16;
    if (nw < 1) nw = 1;
    int path_in_smem = 1;
    size_t slot = warp_base + path_bytes;
    auto total = [&](int nw_, size_t slot_) { return off_slots + (size_t)nw_ * slot_ + 16; };
    if (total(nw, slot) > kMax) { path_in_smem = 0; slot = warp_base; }
    while (nw > 1 && total(nw, slot) > kMax) nw -= 1;
    if (total(nw, slot) > kMax) { why = "recurrent-path parameters do not fit in shared memory"; return 1; }
    pl.block = nw * 32;
    pl.smem = total(nw, slot);
    const int blocks = (warps_total + nw - 1) / nw;
    int per_sm = (int)(kMax / pl.smem);
    if (per_sm < 1) per_sm = 1;
    const int by_threads = 2048 / pl.block;
    if (per_sm > by_threads) per_sm = by_threads < 1 ? 1 : by_threads;
    if (per_sm * pl.block > 512) per_sm = 512 / pl.block > 0 ? 512 / pl.block : 1;  // register budget of __launch_bounds__(512,1)
    const int cap = sm * per_sm;
    pl.grid = blocks < cap ? blocks : cap;
    FusedParams& p = pl.p;
    p.tv = t->tv;
    p.pack = n->d_fpack;
    p.pk = pk;
    p.S = c->num_simulations;
    p.A_cfg = c->action_space_size;
    p.A_net = n->desc.action_space_size;
    p.c_puct = c->c_puct;
    p.frac = c->exploration_fraction;
    p.one_minus_frac = 1.0 - c->exploration_fraction;
    p.alpha = c->dirichlet_alpha;
    p.noise_mode = c->noise_mode;
    p.seed = c->seed;
    p.tree_base = c->tree_index_base;
    p.path_in_smem = path_in_smem;
    p.off_sqrt = (unsigned)off_sqrt;
    p.off_recip = (unsigned)off_recip;
    p.off_slots = (unsigned)off_slots;
    p.slot_bytes = (unsigned)slot;
    p.off_bar = (unsigned)((off_slots + (size_t)nw * slot + 7) & ~(size_t)7);
    p.n_sqrt = n_sqrt;
    p.pdl = 0;
    return 0;
}

struct TeamPlan { int G, grid, block; size_t smem; TeamParams p; };

// Team schedule (search_team.cu): G = pow2 >= A lanes per tree, TT tree slots per team (8 or 16), ntw tree
// warps + nmw network warps per team, n_team teams per CTA, one CTA per SM.  The tree statistics of the
// trees in flight live in shared memory (44 bytes per node), which bounds the trees per team.
static int plan_team(const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c, TeamPlan& pl, std::string& why) {
    const int H = n->desc.hidden_dim, B = t->tv.B, A = t->tv.A, S_max = t->tv.S_max, N = t->tv.N;
    const int sm = n->sm_count;
    if (H % 4 != 0) { why = "hidden_dim is not a multiple of 4"; return 1; }
    if (A > 65535 || S_max > 65533) { why = "tree too large for the packed select payload"; return 1; }
    int G = 2;
    while (G < A && G < 32) G *= 2;
    if (c->lanes_per_tree > 0) {  // explicit lanes per tree: round down to a power of two in [2, 32]
        G = 2;
        while (G * 2 <= c->lanes_per_tree && G < 32) G *= 2;
    }
    const int TPW = 32 / G;
    const int tps = (B + sm - 1) / sm;  // trees per SM
    int TT = (G <= 8 && tps > 8) ? 16 : 8;
    if (c->trees_per_team > 0) TT = (c->trees_per_team > 8 && G <= 8) ? 16 : 8;
    // layer-warpgroup mode prefers up to four small teams (8 tree slots each) over two large ones: the chain of one
    // simulation is mostly latency, so more teams in flight overlap more of it and a team's layers are two short jobs
    const bool lw_wanted = H == 64 && n->desc.num_blocks == 2 && c->schedule != CUMIND_SCHED_TEAM_CLASSIC && G <= 8;
    // (also beyond one round per SM: shared memory holds ~28 trees per CTA whatever the slot count, and with 16 slots per team the
    // layer warps would compute 16 rows for 7 trees — measured 1.51 ms instead of 0.52 ms at 8192 trees)
    if (lw_wanted && c->trees_per_team <= 0 && c->teams_per_cta <= 0 && G <= 4) TT = 8;
    if (!team_variant_ok(G, TT)) { why = "no team kernel for this lane mapping"; return 1; }
    const int ntw = (TT + TPW - 1) / TPW;
    int nmw = c->net_warps_per_team > 0 ? c->net_warps_per_team : 4;
    if (nmw > 8) nmw = 8;
    int team_warps = ntw + nmw;
    const int XS = TT + 4;
    // layer-warpgroup mode (search_team.cu, KR > 0): hidden_dim 64, two dynamics layers; warps 0-3 of the CTA run the layers
    // of all teams, the CTA consists of whole warpgroups and a team's warp count follows from the number of teams
    bool lw = team_lw_variant_ok(G, TT, H) && n->desc.num_blocks == 2 && ntw <= 2 && c->schedule != CUMIND_SCHED_TEAM_CLASSIC;
    if (const char* e = getenv("CUMIND_TEAM_LW")) lw = lw && atoi(e) != 0;   // tuning only
    const int lw_warps = lw ? 4 : 0;
    const int RS = 68;
    const TeamPackLayout& pk = n->team_pk;
    const size_t pack_bytes = pk.total * sizeof(float);
    const int n_sqrt = 2 * S_max + 4;
    auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t off_sqrt = al16(pack_bytes);
    const size_t off_recip = al16(off_sqrt + (size_t)n_sqrt * sizeof(double));
    const size_t off_teams = (off_recip + (size_t)n_sqrt * sizeof(double) + 127) & ~(size_t)127;
    const size_t team_base = lw ? al16((size_t)(3 * TT * RS + 3 * TT * pk.HP + 4 * TT) * 4)
                                : al16((size_t)(2 * H * XS + TT * (H + 4) + 3 * TT * pk.HP + 4 * TT) * 4);
    const size_t per_tree = (size_t)(S_max + 2) * sizeof(int) + (size_t)N * 44;  // path + selection cache + hot W + hot records
    const size_t kMax = 227 * 1024;
    if (off_teams + team_base + per_tree + 256 > kMax) { why = "one tree does not fit the team kernel's shared memory"; return 1; }
    int n_team = c->teams_per_cta > 0 ? c->teams_per_cta : (tps + TT - 1) / TT;
    if (n_team < 1) n_team = 1;
    if (lw && c->teams_per_cta <= 0 && TT == 8) n_team = (tps + 7) / 8;
    if (n_team > (lw ? 4 : 7)) n_team = lw ? 4 : 7;
    if (!lw) {
        // the automatic choice keeps the CTA at <= 512 threads (128 registers per thread)
        const int max_threads = c->teams_per_cta > 0 ? 1024 : 512;
        while (n_team > 1 && n_team * team_warps * 32 > max_threads) --n_team;
        if (n_team * team_warps * 32 > 1024) { why = "team does not fit a CTA"; return 1; }
    }
    // trees per batch: spread the SM's trees evenly over its teams, within TT and within shared memory
    int tpt = 0;
    for (;; --n_team) {
        const size_t budget = (kMax - 64 - off_teams) / n_team;
        const long long cap = budget > team_base + 512 ? (long long)((budget - team_base - 512) / per_tree) : 0;
        tpt = (tps + n_team - 1) / n_team;
        if (tpt > TT) tpt = TT;
        if (c->trees_per_team > 0 && c->trees_per_team < tpt) tpt = c->trees_per_team;
        if (cap >= 1) {
            if (tpt > cap) tpt = (int)cap;
            break;
        }
        if (n_team == 1) { why = "one tree does not fit the team kernel's shared memory"; return 1; }
    }
    if (tpt < 1) tpt = 1;
    if (lw) {
        team_warps = team_lw_team_warps(n_team);
        nmw = team_warps - ntw;
    }
    const size_t off_path = team_base;
    const size_t off_nxt = al16(off_path + (size_t)tpt * (S_max + 2) * sizeof(int));
    const size_t off_hotw = al16(off_nxt + (size_t)tpt * N * sizeof(int));
    const size_t off_hot = al16(off_hotw + (size_t)tpt * N * sizeof(double));
    const size_t team_bytes = (off_hot + (size_t)tpt * N * 32 + 127) & ~(size_t)127;
    const size_t total = off_teams + (size_t)n_team * team_bytes + 16 + (lw ? 32 * (size_t)n_team : 0);
    if (total > kMax) { why = "internal: team scratch exceeds shared memory"; return 1; }
    const int n_batches = (B + tpt - 1) / tpt;
    int grid = (n_batches + n_team - 1) / n_team;
    if (grid > sm) grid = sm;
    pl.G = G;
    pl.grid = grid;
    pl.block = (lw_warps + n_team * team_warps) * 32;
    pl.smem = total;
    TeamParams& p = pl.p;
    p.tv = t->tv;
    p.pack = n->d_team_pack;
    p.pk = pk;
    p.S = c->num_simulations;
    p.A_cfg = c->action_space_size;
    p.A_net = n->desc.action_space_size;
    p.c_puct = c->c_puct;
    p.frac = c->exploration_fraction;
    p.one_minus_frac = 1.0 - c->exploration_fraction;
    p.alpha = c->dirichlet_alpha;
    p.noise_mode = c->noise_mode;
    p.seed = c->seed;
    p.tree_base = c->tree_index_base;
    p.n_team = n_team; p.ntw = ntw; p.nmw = nmw; p.TT = TT; p.tpt = tpt; p.n_batches = n_batches;
    p.off_sqrt = (unsigned)off_sqrt;
    p.off_recip = (unsigned)off_recip;
    p.off_teams = (unsigned)off_teams;
    p.team_bytes = (unsigned)team_bytes;
    p.off_path = (unsigned)off_path;
    p.off_nxt = (unsigned)off_nxt;
    p.off_hotw = (unsigned)off_hotw;
    p.off_hot = (unsigned)off_hot;
    p.off_bar = (unsigned)((off_teams + (size_t)n_team * team_bytes + 7) & ~(size_t)7);
    p.n_sqrt = n_sqrt;
    p.clk = nullptr;
    p.pdl = 0;
    p.lw = lw ? 64 : 0;
    p.tile = 0;
    p.stagger = 0;
    if (const char* e = getenv("CUMIND_TEAM_TILE")) p.tile = atoi(e);        // tuning only
    if (const char* e = getenv("CUMIND_TEAM_STAGGER")) p.stagger = atoi(e);  // tuning only
    return 0;
}

// CUMIND_SCHED_AUTO: the team schedule wins while every tree of an SM is in flight at once (its trees live in
// shared memory, so it holds a few dozen per SM; measured 0.33 vs 0.39 ms at 28 trees per SM); beyond one
// round per CTA the warp schedule, which keeps its trees in HBM/L2 and runs 100+ per SM, is faster.
// With the layer warpgroup two nearly full rounds still win: 8192 trees (BASELINE configs[3] on 8 GPUs) take 0.518 ms per
// step in two rounds of 4736 against 0.545 ms on the warp schedule, while 6144 trees take the same 0.508 ms against 0.472 ms.
static bool team_is_preferred(const TeamPlan& tp, int sm_count) {
    const long long in_flight = (long long)tp.p.n_team * tp.p.tpt * (tp.grid < sm_count ? tp.grid : sm_count);
    if (in_flight >= tp.p.tv.B) return true;
    return tp.p.lw > 0 && tp.p.tv.B <= 2 * in_flight && 100LL * tp.p.tv.B >= 77LL * 2 * in_flight;
}

extern "C" int cumind_search_stepwise(cumind_tree* t, const cumind_net* n, const float* d_root_hidden, const cumind_search_cfg* c,
                                      const double* d_noise, int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                                      void* stream) {
    if (check_search("cumind_search_stepwise", t, n, c, d_noise)) return 1;
    if (!d_root_hidden) return fail("cumind_search_stepwise: d_root_hidden is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const StepParams p = step_params(t, n, c);
    CM_CUDA(launch_init_root(p, d_root_hidden, d_noise, st));
    const bool mm = c->tree_mode == CUMIND_TREE_MINMAX_DISCOUNT;
    for (int s = 0; s < c->num_simulations; ++s) {
        if (mm) CM_CUDA(launch_select_mm(p, st));
        else CM_CUDA(launch_select(p, st));
        if (c->net_mode == CUMIND_NET_BF16_TC)
            CM_CUDA(launch_expand_tc(n->d_tcpack, n->tpk, n->sm_count, t->tv, t->leaf_value, st));
        else
            CM_CUDA(launch_expand(p, t->leaf_value, st));
        if (mm) CM_CUDA(launch_backup_mm(p, t->leaf_value, st));
        else CM_CUDA(launch_backup(p, t->leaf_value, st));
    }
    CM_CUDA(launch_finalize(p, d_out_visits, d_out_probs, d_out_root_value, st));
    g_launches += 2 + 3ull * c->num_simulations;
    return 0;
}

// CUMIND_NET_BF16_TC: the tensor-core search kernel when asked for, or under AUTO when the batch is large enough for it
// to beat the fp32 kernels (include/cumind_b200.h, CUMIND_SCHED_TENSOR)
static bool use_tensor_search(const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c) {
    if (c->net_mode != CUMIND_NET_BF16_TC || !n->tc_ok || c->tree_mode != CUMIND_TREE_REFERENCE) return false;
    if (c->schedule == CUMIND_SCHED_TENSOR) return true;
    if (c->schedule != CUMIND_SCHED_AUTO) return false;
    return (long long)t->tv.B > (long long)CUMIND_TENSOR_MIN_TREES_PER_SM * n->sm_count;
}

// pdl: d_root_hidden is written by the kernel launched just before on `stream`, which calls pdl_launch_dependents():
// the search kernel is launched with programmatic stream serialization so that its prologue overlaps that kernel.
// out_actions (may be NULL): first arg-max of the root visit counts, written by the fused kernels themselves
static int search_impl(cumind_tree* t, const cumind_net* n, const float* d_root_hidden, const cumind_search_cfg* c,
                       const double* d_noise, int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                       void* stream, bool pdl, int32_t* out_actions = nullptr, bool* wrote_actions = nullptr) {
    {
        static const bool no_pdl = getenv("CUMIND_NO_PDL") != nullptr;   // tuning only
        if (no_pdl) pdl = false;
    }
    if (wrote_actions) *wrote_actions = false;
    if (check_search("cumind_search", t, n, c, d_noise)) return 1;
    if (!d_root_hidden || !d_out_visits || !d_out_probs) return fail("cumind_search: NULL device pointer");
    FusedPlan pl;
    std::string why;
    // CUMIND_NET_FP32_EXACT: the fused kernels below (pinned-order fp32 SIMT network).  CUMIND_NET_BF16_TC: the fused
    // tensor-core kernel (search_tc.cu: tcgen05 MMAs per simulation, 128 trees per CTA).
    if (c->schedule < 0 || c->schedule > CUMIND_SCHED_TENSOR) return fail("unknown schedule");
    if (c->schedule == CUMIND_SCHED_TENSOR && c->net_mode != CUMIND_NET_BF16_TC) return fail("CUMIND_SCHED_TENSOR needs CUMIND_NET_BF16_TC");
    if (c->tree_mode == CUMIND_TREE_MINMAX_DISCOUNT)   // the optional mode exists on the one-launch-per-phase kernels only
        return cumind_search_stepwise(t, n, d_root_hidden, c, d_noise, d_out_visits, d_out_probs, d_out_root_value, stream);
    if (use_tensor_search(t, n, c)) {
        SearchTcParams q{};
        q.tv = t->tv;
        q.pack = n->d_tcpack;
        q.pk = n->tpk;
        q.pack_f32 = n->d_pack;
        q.pkf = n->pk;
        q.root_hidden = d_root_hidden;
        q.noise = d_noise;
        q.out_visits = d_out_visits;
        q.out_probs = d_out_probs;
        q.out_root_value = d_out_root_value;
        q.out_actions = out_actions;
        if (wrote_actions) *wrote_actions = out_actions != nullptr;
        q.S = c->num_simulations;
        q.A_cfg = c->action_space_size;
        q.c_puct = c->c_puct;
        q.frac = c->exploration_fraction;
        q.one_minus_frac = 1.0 - c->exploration_fraction;
        q.alpha = c->dirichlet_alpha;
        q.noise_mode = c->noise_mode;
        q.seed = c->seed;
        q.tree_base = c->tree_index_base;
        q.n_sqrt = 2 * t->tv.S_max + 4;
        q.pdl = 0;   // as for the warp schedule below: measured 0.08 ms slower per step at 32 768 trees with the dependent launch
        CM_CUDA(launch_search_tc(q, n->sm_count, (cudaStream_t)stream));
        g_launches += 1;
        return 0;
    }
    if (c->schedule != CUMIND_SCHED_WARP) {
        TeamPlan tp;
        if (plan_team(t, n, c, tp, why) == 0 && (c->schedule == CUMIND_SCHED_TEAM || c->schedule == CUMIND_SCHED_TEAM_CLASSIC || team_is_preferred(tp, n->sm_count))) {
            tp.p.root_hidden = d_root_hidden;
            tp.p.noise = d_noise;
            tp.p.out_visits = d_out_visits;
            tp.p.out_probs = d_out_probs;
            tp.p.out_root_value = d_out_root_value;
            tp.p.out_actions = out_actions;
            if (wrote_actions) *wrote_actions = out_actions != nullptr;
            tp.p.pdl = pdl ? 1 : 0;
            CM_CUDA(launch_search_team(tp.G, tp.p, tp.grid, tp.block, tp.smem, (cudaStream_t)stream));
            g_launches += 1;
            return 0;
        }
    }
    if (plan_fused(t, n, c, pl, why) != 0)
        return cumind_search_stepwise(t, n, d_root_hidden, c, d_noise, d_out_visits, d_out_probs, d_out_root_value, stream);
    pl.p.root_hidden = d_root_hidden;
    pl.p.noise = d_noise;
    pl.p.out_visits = d_out_visits;
    pl.p.out_probs = d_out_probs;
    pl.p.out_root_value = d_out_root_value;
    pl.p.out_actions = out_actions;
    if (wrote_actions) *wrote_actions = out_actions != nullptr;
    pl.p.pdl = 0;   // measured at 8 192 trees: 0.642 ms per step with the dependent launch, 0.548 without (the many-CTA grid gains nothing from an early start)
    CM_CUDA(launch_search_fused(pl.G, pl.R, pl.p, pl.grid, pl.block, pl.smem, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_search(cumind_tree* t, const cumind_net* n, const float* d_root_hidden, const cumind_search_cfg* c,
                             const double* d_noise, int32_t* d_out_visits, double* d_out_probs, double* d_out_root_value,
                             void* stream) {
    return search_impl(t, n, d_root_hidden, c, d_noise, d_out_visits, d_out_probs, d_out_root_value, stream, false);
}

// Agent.select_action for a batch with every buffer on the device (agent.py:73-79: encode -> search is one call in
// the reference too): initial_inference (representation network only) and the search are issued back to back from
// here, so no host work sits between the two launches, and for vector observations the search kernel is a
// programmatic dependent launch whose parameter staging overlaps the encoder.
extern "C" int cumind_search_from_obs(cumind_tree* t, const cumind_net* n, const float* d_obs, const cumind_search_cfg* c,
                                      const double* d_noise, float* d_hidden_work, int32_t* d_out_visits, double* d_out_probs,
                                      double* d_out_root_value, void* stream) {
    if (!t || !n || !c || !d_obs || !d_hidden_work) return fail("cumind_search_from_obs: NULL argument");
    int rc = cumind_initial_inference(n, d_obs, t->tv.B, d_hidden_work, nullptr, nullptr, c->net_mode, stream);
    if (rc) return rc;
    return search_impl(t, n, d_hidden_work, c, d_noise, d_out_visits, d_out_probs, d_out_root_value, stream, n->desc.obs_ndim == 1);
}

// introspection for tests / bench: the launch configuration cumind_search would use
extern "C" int cumind_search_plan(const cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c, int32_t* out /*[8]*/) {
    if (!t || !n || !c || !out) return fail("cumind_search_plan: NULL argument");
    std::string why;
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (use_tensor_search(t, n, c)) {
        out[0] = 4; out[1] = 1; out[3] = (t->tv.B + 127) / 128; out[4] = 128;   // fused tensor-core kernel: one thread per tree
        return 0;
    }
    if (c->schedule != CUMIND_SCHED_WARP) {
        TeamPlan tp;
        if (plan_team(t, n, c, tp, why) == 0 && (c->schedule == CUMIND_SCHED_TEAM || c->schedule == CUMIND_SCHED_TEAM_CLASSIC || team_is_preferred(tp, n->sm_count))) {
            out[0] = CUMIND_SCHED_TEAM; out[1] = tp.G; out[2] = tp.p.tpt; out[3] = tp.grid; out[4] = tp.block;
            out[5] = (int32_t)tp.smem; out[6] = tp.p.n_team; out[7] = tp.p.nmw + 256 * tp.p.lw;
            return 0;
        }
    }
    FusedPlan pl;
    if (plan_fused(t, n, c, pl, why) != 0) return 0;
    out[0] = CUMIND_SCHED_WARP; out[1] = pl.G; out[2] = pl.R; out[3] = pl.grid; out[4] = pl.block; out[5] = (int32_t)pl.smem;
    return 0;
}

extern "C" int cumind_search_phase_clocks(cumind_tree* t, const cumind_net* n, const float* d_root_hidden, const cumind_search_cfg* c,
                                          const double* d_noise, int32_t* d_out_visits, double* d_out_probs,
                                          double* d_out_root_value, int64_t* d_clocks, void* stream) {
    if (check_search("cumind_search_phase_clocks", t, n, c, d_noise)) return 1;
    if (!d_root_hidden || !d_out_visits || !d_out_probs || !d_clocks) return fail("cumind_search_phase_clocks: NULL device pointer");
    TeamPlan tp;
    std::string why;
    if (plan_team(t, n, c, tp, why) != 0) return fail("cumind_search_phase_clocks: " + why);
    tp.p.root_hidden = d_root_hidden;
    tp.p.noise = d_noise;
    tp.p.out_visits = d_out_visits;
    tp.p.out_probs = d_out_probs;
    tp.p.out_root_value = d_out_root_value;
    tp.p.out_actions = nullptr;
    tp.p.clk = (long long*)d_clocks;
    CM_CUDA(launch_search_team(tp.G, tp.p, tp.grid, tp.block, tp.smem, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// ---- fine-grained phases ------------------------------------------------------------------------
static cumind_search_cfg cfg_for_tree(const cumind_tree* t, const cumind_search_cfg* c) { (void)t; return *c; }

extern "C" int cumind_tree_init_root(cumind_tree* t, const cumind_net* n, const float* d_root_hidden, const cumind_search_cfg* c,
                                     const double* d_noise, void* stream) {
    if (check_search("cumind_tree_init_root", t, n, c, d_noise)) return 1;
    if (!d_root_hidden) return fail("cumind_tree_init_root: d_root_hidden is NULL");
    CM_CUDA(launch_init_root(step_params(t, n, c), d_root_hidden, d_noise, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_select(cumind_tree* t, const cumind_search_cfg* c, void* stream) {
    if (!t || !c) return fail("cumind_select: NULL argument");
    StepParams p{};
    p.tv = t->tv;
    p.c_puct = c->c_puct;
    CM_CUDA(launch_select(p, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_expand(cumind_tree* t, const cumind_net* n, const cumind_search_cfg* c, float* d_leaf_value, void* stream) {
    if (!t || !n || !c || !d_leaf_value) return fail("cumind_expand: NULL argument");
    if (n->desc.hidden_dim != t->tv.H) return fail("cumind_expand: tree was created for a different hidden_dim");
    const cumind_search_cfg cc = cfg_for_tree(t, c);
    CM_CUDA(launch_expand(step_params(t, n, &cc), d_leaf_value, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_backup(cumind_tree* t, const float* d_leaf_value, void* stream) {
    if (!t || !d_leaf_value) return fail("cumind_backup: NULL argument");
    StepParams p{};
    p.tv = t->tv;
    CM_CUDA(launch_backup(p, d_leaf_value, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_finalize(cumind_tree* t, const cumind_search_cfg* c, int32_t* d_out_visits, double* d_out_probs,
                               double* d_out_root_value, void* stream) {
    if (!t || !c) return fail("cumind_finalize: NULL argument");
    StepParams p{};
    p.tv = t->tv;
    p.A_cfg = c->action_space_size;
    CM_CUDA(launch_finalize(p, d_out_visits, d_out_probs, d_out_root_value, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

extern "C" int cumind_dirichlet(double* d_noise, int32_t B, int32_t A, double alpha, uint64_t seed, uint64_t tree_index_base,
                                void* stream) {
    if (!d_noise || B <= 0 || A <= 0 || !(alpha > 0.0)) return fail("cumind_dirichlet: bad argument");
    CM_CUDA(launch_dirichlet(d_noise, B, A, alpha, seed, tree_index_base, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// optax.adamw step over a flat device blob (agent.py:48, trainer.py:97-100); see train.cu.
extern "C" int cumind_adamw_step(float* d_params, const float* d_grads, float* d_m, float* d_v, const uint8_t* d_mask, size_t n,
                                 float lr, float b1, float b2, float eps, float weight_decay, int64_t step, float grad_scale,
                                 void* stream) {
    if (!d_params || !d_grads || !d_m || !d_v || step < 1) return fail("cumind_adamw_step: bad argument");
    CM_CUDA(launch_adamw(d_params, d_grads, d_m, d_v, d_mask, n, lr, b1, b2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// Trainer.compute_losses + nnx.value_and_grad (trainer.py:88-95, 172-205) for vector-observation networks: see train_fused.cu.
extern "C" int cumind_train_supported(const cumind_net_desc* desc, int32_t K) {
    return desc_valid(desc) && train_fused_supported(*desc, K) ? 1 : 0;
}
extern "C" size_t cumind_train_work_bytes(const cumind_net_desc* desc, int32_t B) {
    if (!desc_valid(desc) || B <= 0) return 0;
    return train_fused_work_bytes(*desc, B);
}
extern "C" int cumind_train_loss_grad(const cumind_net_desc* desc, const float* d_params, const float* d_obs, const int32_t* d_actions,
                                      const float* d_value_targets, const float* d_policy_targets, const float* d_reward_targets, int32_t B,
                                      int32_t K, void* d_work, size_t work_bytes, float* d_grad, float* d_losses4, void* stream) {
    if (!desc_valid(desc) || !d_params || !d_obs || !d_actions || !d_value_targets || !d_policy_targets || !d_reward_targets || !d_work ||
        !d_grad || !d_losses4 || B <= 0)
        return fail("cumind_train_loss_grad: bad argument");
    if (!train_fused_supported(*desc, K)) return fail("cumind_train_loss_grad: network shape not supported (see cumind_train_supported)");
    if (work_bytes < train_fused_work_bytes(*desc, B)) return fail("cumind_train_loss_grad: work buffer too small");
    CM_CUDA(launch_train_fwd_bwd(*desc, d_params, d_obs, d_actions, d_value_targets, d_policy_targets, d_reward_targets, B, K, (float*)d_work,
                                 d_grad, d_losses4, (cudaStream_t)stream));
    g_launches += 2;
    return 0;
}

// Same step with the counter on the device (d_step is incremented first): the launch arguments do not
// change between steps, so the whole train step can be replayed from a CUDA graph.
extern "C" int cumind_adamw_step_dev(float* d_params, const float* d_grads, float* d_m, float* d_v, const uint8_t* d_mask, size_t n,
                                     float lr, float b1, float b2, float eps, float weight_decay, int64_t* d_step, float grad_scale,
                                     void* stream) {
    if (!d_params || !d_grads || !d_m || !d_v || !d_step) return fail("cumind_adamw_step_dev: bad argument");
    CM_CUDA(launch_adamw_dev_step(d_params, d_grads, d_m, d_v, d_mask, n, lr, b1, b2, eps, weight_decay, (long long*)d_step, grad_scale,
                                  (cudaStream_t)stream));
    g_launches += 2;
    return 0;
}

// nnx.update(network, params) from a DEVICE blob (after an optimiser step on the GPU): one device copy
// and the gather kernels on `stream`; asynchronous, no host round trip.
extern "C" int cumind_net_update_params_device(cumind_net* n, const float* d_params, size_t n_params, void* stream) {
    if (!n || !d_params) return fail("cumind_net_update_params_device: NULL argument");
    if (n_params != n->n_params) return fail("cumind_net_update_params_device: parameter blob has the wrong size");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_params != n->d_blob)
        CM_CUDA(cudaMemcpyAsync(n->d_blob, d_params, sizeof(float) * n_params, cudaMemcpyDeviceToDevice, st));
    return repack_device(n, st);
}

// ------------------------------------------------------------------------------------------------
// gradient all-reduce over peer memory fused into AdamW (p2p_adamw.cu)
// ------------------------------------------------------------------------------------------------
struct cumind_p2p {
    int rank, world;
    size_t n;
    float* d_grad;
    long long* d_flags;
    int* d_err;       // [0] error flag, [1] block counter
    void* peer_grad[P2P_MAX_RANKS];
    void* peer_flags[P2P_MAX_RANKS];
    bool opened;
};

extern "C" int cumind_p2p_create(size_t n, int32_t rank, int32_t world, cumind_p2p** out, void* h_handles128) {
    if (!out || !h_handles128 || n == 0 || world < 1 || world > P2P_MAX_RANKS || rank < 0 || rank >= world)
        return fail("cumind_p2p_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two IPC handles are exchanged as 128 bytes");
    cumind_p2p* q = new cumind_p2p();
    q->rank = rank; q->world = world; q->n = n; q->opened = false;
    q->d_grad = nullptr; q->d_flags = nullptr; q->d_err = nullptr;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) q->peer_grad[r] = q->peer_flags[r] = nullptr;
    cudaError_t e;
    cudaIpcMemHandle_t hg, hf;
    if ((e = cudaMalloc(&q->d_grad, n * sizeof(float))) != cudaSuccess || (e = cudaMemset(q->d_grad, 0, n * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&q->d_flags, 2 * P2P_MAX_RANKS * sizeof(long long))) != cudaSuccess ||
        (e = cudaMemset(q->d_flags, 0, 2 * P2P_MAX_RANKS * sizeof(long long))) != cudaSuccess ||
        (e = cudaMalloc(&q->d_err, 2 * sizeof(int))) != cudaSuccess || (e = cudaMemset(q->d_err, 0, 2 * sizeof(int))) != cudaSuccess ||
        (e = cudaIpcGetMemHandle(&hg, q->d_grad)) != cudaSuccess || (e = cudaIpcGetMemHandle(&hf, q->d_flags)) != cudaSuccess ||
        (e = cudaDeviceSynchronize()) != cudaSuccess) {
        cudaFree(q->d_grad); cudaFree(q->d_flags); cudaFree(q->d_err);
        delete q;
        return fail_cuda("cumind_p2p_create", e);
    }
    memcpy(h_handles128, &hg, 64);
    memcpy((char*)h_handles128 + 64, &hf, 64);
    q->peer_grad[rank] = q->d_grad;
    q->peer_flags[rank] = q->d_flags;
    *out = q;
    return 0;
}

// h_all_handles: world x 128 bytes, the output of every rank's cumind_p2p_create in rank order
extern "C" int cumind_p2p_open(cumind_p2p* q, const void* h_all_handles) {
    if (!q || !h_all_handles) return fail("cumind_p2p_open: NULL argument");
    for (int r = 0; r < q->world; ++r) {
        if (r == q->rank) continue;
        cudaIpcMemHandle_t hg, hf;
        memcpy(&hg, (const char*)h_all_handles + (size_t)r * 128, 64);
        memcpy(&hf, (const char*)h_all_handles + (size_t)r * 128 + 64, 64);
        CM_CUDA(cudaIpcOpenMemHandle(&q->peer_grad[r], hg, cudaIpcMemLazyEnablePeerAccess));
        CM_CUDA(cudaIpcOpenMemHandle(&q->peer_flags[r], hf, cudaIpcMemLazyEnablePeerAccess));
    }
    q->opened = true;
    return 0;
}

extern "C" float* cumind_p2p_grad_ptr(cumind_p2p* q) { return q ? q->d_grad : nullptr; }

extern "C" int cumind_p2p_allreduce_adamw(cumind_p2p* q, float* d_params, float* d_m, float* d_v, const uint8_t* d_mask, size_t n, float lr,
                                          float b1, float b2, float eps, float weight_decay, int64_t* d_step, float* d_sum_out, void* stream) {
    if (!q || !d_params || !d_m || !d_v || !d_step || n != q->n) return fail("cumind_p2p_allreduce_adamw: bad argument");
    if (!q->opened && q->world > 1) return fail("cumind_p2p_allreduce_adamw: peers not opened (cumind_p2p_open)");
    P2PArgs a;
    a.rank = q->rank; a.world = q->world;
    for (int r = 0; r < P2P_MAX_RANKS; ++r) { a.peer_grad[r] = (const float*)q->peer_grad[r]; a.peer_flags[r] = (long long*)q->peer_flags[r]; }
    a.sum_out = d_sum_out;
    a.err = q->d_err;
    a.count = q->d_err + 1;
    CM_CUDA(launch_p2p_allreduce_adamw(a, d_params, d_m, d_v, d_mask, n, lr, b1, b2, eps, weight_decay, (long long*)d_step, (cudaStream_t)stream));
    g_launches += 2;
    return 0;
}

// forget the barrier epochs (after the step counter was set back, e.g. a checkpoint was loaded): call on every rank between two
// host-side barriers
extern "C" int cumind_p2p_reset(cumind_p2p* q) {
    if (!q) return fail("cumind_p2p_reset: NULL argument");
    CM_CUDA(cudaDeviceSynchronize());
    CM_CUDA(cudaMemset(q->d_flags, 0, 2 * P2P_MAX_RANKS * sizeof(long long)));
    CM_CUDA(cudaMemset(q->d_err, 0, 2 * sizeof(int)));
    CM_CUDA(cudaDeviceSynchronize());
    return 0;
}

// 1 when a rank barrier timed out since creation (synchronises the device)
extern "C" int cumind_p2p_error(cumind_p2p* q) {
    if (!q) return -1;
    int e = 0;
    if (cudaMemcpy(&e, q->d_err, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return e;
}

extern "C" int cumind_p2p_destroy(cumind_p2p* q) {
    if (!q) return 0;
    for (int r = 0; r < q->world; ++r) {
        if (r == q->rank) continue;
        if (q->peer_grad[r]) cudaIpcCloseMemHandle(q->peer_grad[r]);
        if (q->peer_flags[r]) cudaIpcCloseMemHandle(q->peer_flags[r]);
    }
    cudaFree(q->d_grad); cudaFree(q->d_flags); cudaFree(q->d_err);
    delete q;
    return 0;
}

extern "C" int cumind_selftest_division(uint64_t n_per_thread, uint64_t seed, int32_t max_divisor, uint64_t* d_counts2,
                                        void* stream) {
    if (!d_counts2 || max_divisor <= 0) return fail("cumind_selftest_division: bad argument");
    CM_CUDA(launch_selftest_div(n_per_thread, seed, max_divisor, (unsigned long long*)d_counts2, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// TreeBuffer's sum tree in device memory (memory.py:190-299)
// ------------------------------------------------------------------------------------------------
struct cumind_sumtree {
    double* d_tree;
    long long capacity, tree_size, n_nodes;
    int depth;
};

extern "C" int cumind_sumtree_create(int64_t capacity, cumind_sumtree** out) {
    if (!out || capacity < 1 || capacity > ((int64_t)1 << 40)) return fail("cumind_sumtree_create: bad argument");
    cumind_sumtree* t = new cumind_sumtree();
    t->capacity = capacity;
    t->tree_size = 1;
    t->depth = 0;
    while (t->tree_size < capacity) { t->tree_size *= 2; ++t->depth; }   // memory.py:204-206
    t->n_nodes = 2 * t->tree_size - 1;
    cudaError_t e = cudaMalloc(&t->d_tree, sizeof(double) * (size_t)t->n_nodes);
    if (e == cudaSuccess) e = cudaMemset(t->d_tree, 0, sizeof(double) * (size_t)t->n_nodes);
    if (e != cudaSuccess) { cudaFree(t->d_tree); delete t; return fail_cuda("cumind_sumtree_create", e); }
    *out = t;
    return 0;
}
extern "C" void cumind_sumtree_destroy(cumind_sumtree* t) {
    if (!t) return;
    cudaFree(t->d_tree);
    delete t;
}
extern "C" int64_t cumind_sumtree_tree_size(const cumind_sumtree* t) { return t ? t->tree_size : 0; }
extern "C" int cumind_sumtree_clear(cumind_sumtree* t, void* stream) {
    if (!t) return fail("cumind_sumtree_clear: NULL argument");
    CM_CUDA(cudaMemsetAsync(t->d_tree, 0, sizeof(double) * (size_t)t->n_nodes, (cudaStream_t)stream));
    return 0;
}
extern "C" int cumind_sumtree_update(cumind_sumtree* t, const int64_t* d_tree_idx, const double* d_priority, int64_t n, void* stream) {
    if (!t || !d_tree_idx || !d_priority || n < 0) return fail("cumind_sumtree_update: bad argument");
    if (n == 0) return 0;
    CM_CUDA(launch_sumtree_update(t->d_tree, t->n_nodes, t->depth, (const long long*)d_tree_idx, d_priority, n, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}
extern "C" int cumind_sumtree_sample(const cumind_sumtree* t, const double* d_uniform, int64_t batch, int64_t* d_tree_idx,
                                     int64_t* d_data_idx, void* stream) {
    if (!t || !d_uniform || !d_tree_idx || !d_data_idx || batch < 1) return fail("cumind_sumtree_sample: bad argument");
    CM_CUDA(launch_sumtree_sample(t->d_tree, t->n_nodes, t->tree_size, d_uniform, batch, (long long*)d_tree_idx, (long long*)d_data_idx,
                                  (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}
extern "C" int cumind_sumtree_read(const cumind_sumtree* t, double* d_out, void* stream) {
    if (!t || !d_out) return fail("cumind_sumtree_read: NULL argument");
    CM_CUDA(cudaMemcpyAsync(d_out, t->d_tree, sizeof(double) * (size_t)t->n_nodes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Trajectory records of the replay buffers in device memory (memory.py:13-85 holds them in a deque; trainer.py:118-170 reads them)
// ------------------------------------------------------------------------------------------------
struct cumind_replay {
    unsigned char* d_ring;
    long long capacity, obs_size, rec_bytes;
    int A, K;
};
static long long replay_record_bytes(long long obs_size, int A, int K) {
    const long long b = 16 + 4 * (2 * obs_size + A + 2 * (long long)K);
    return (b + 15) / 16 * 16;
}
extern "C" int cumind_replay_create(int64_t capacity, int64_t obs_size, int32_t n_actions, int32_t unroll_steps, cumind_replay** out) {
    if (!out || capacity < 1 || obs_size < 1 || n_actions < 1 || unroll_steps < 1) return fail("cumind_replay_create: bad argument");
    cumind_replay* r = new cumind_replay();
    r->capacity = capacity;
    r->obs_size = obs_size;
    r->A = n_actions;
    r->K = unroll_steps;
    r->rec_bytes = replay_record_bytes(obs_size, n_actions, unroll_steps);
    cudaError_t e = cudaMalloc(&r->d_ring, (size_t)(r->rec_bytes * capacity));
    if (e == cudaSuccess) e = cudaMemset(r->d_ring, 0, (size_t)(r->rec_bytes * capacity));
    if (e != cudaSuccess) { delete r; return fail_cuda("cumind_replay_create", e); }
    *out = r;
    return 0;
}
extern "C" void cumind_replay_destroy(cumind_replay* r) {
    if (!r) return;
    cudaFree(r->d_ring);
    delete r;
}
extern "C" int64_t cumind_replay_record_bytes(const cumind_replay* r) { return r ? r->rec_bytes : 0; }
extern "C" int cumind_replay_write(cumind_replay* r, int64_t slot0, const void* records, int64_t n, void* stream) {
    if (!r || !records || n < 0 || n > r->capacity || slot0 < 0 || slot0 >= r->capacity) return fail("cumind_replay_write: bad argument");
    const long long first = std::min<long long>(n, r->capacity - slot0);
    cudaStream_t st = (cudaStream_t)stream;
    if (first > 0)
        CM_CUDA(cudaMemcpyAsync(r->d_ring + (size_t)(slot0 * r->rec_bytes), records, (size_t)(first * r->rec_bytes), cudaMemcpyDefault, st));
    if (n > first)   // the ring wraps
        CM_CUDA(cudaMemcpyAsync(r->d_ring, (const unsigned char*)records + (size_t)(first * r->rec_bytes), (size_t)((n - first) * r->rec_bytes),
                                cudaMemcpyDefault, st));
    return 0;
}
extern "C" int cumind_replay_gather(const cumind_replay* r, const int64_t* d_index, int64_t head, int64_t n_valid, int32_t B, float* d_obs,
                                    float* d_boot_obs, float* d_policy, float* d_rewards, int32_t* d_actions, double* d_partial,
                                    int32_t* d_has_boot, int32_t* d_flags, void* stream) {
    if (!r || !d_index || !d_obs || !d_boot_obs || !d_policy || !d_rewards || !d_actions || !d_partial || !d_has_boot || !d_flags || B < 1 ||
        head < 0 || head >= r->capacity || n_valid < 0 || n_valid > r->capacity)
        return fail("cumind_replay_gather: bad argument");
    CM_CUDA(launch_replay_gather(r->d_ring, r->capacity, r->rec_bytes, r->obs_size, r->A, r->K, (const long long*)d_index, head, n_valid, B, d_obs,
                                 d_boot_obs, d_policy, d_rewards, d_actions, d_partial, d_has_boot, d_flags, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}
extern "C" int cumind_replay_values(const double* d_partial, const int32_t* d_has_boot, const float* d_boot_value, double discount_pow_n,
                                    int32_t B, float* d_values, void* stream) {
    if (!d_partial || !d_has_boot || !d_boot_value || !d_values || B < 1) return fail("cumind_replay_values: bad argument");
    CM_CUDA(launch_replay_values(d_partial, d_has_boot, d_boot_value, discount_pow_n, B, d_values, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Agent.select_action for a batch, host buffers in / out (agent.py:61-89)
// ------------------------------------------------------------------------------------------------
// second stream + events of the sliced host-to-device copy (one set per device, created on first use, never destroyed)
struct CopyLane { cudaStream_t stream; cudaEvent_t ready; cudaEvent_t landed[4]; };
static CopyLane* copy_lane(int device) {
    static std::mutex mu;
    static CopyLane* lanes[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!lanes[device]) {
        CopyLane* l = new CopyLane();
        bool ok = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ready, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < 4; ++i) ok = cudaEventCreateWithFlags(&l->landed[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { delete l; return nullptr; }
        lanes[device] = l;
    }
    return lanes[device];
}

struct WorkOffsets { size_t obs, hidden, noise, visits, probs, actions, obs_u8, total; };
static WorkOffsets work_offsets(const cumind_net_desc* d, int64_t B, int64_t A_cfg) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t osz = d->obs_ndim == 1 ? (size_t)d->obs_shape[0] : (size_t)d->obs_shape[0] * d->obs_shape[1] * d->obs_shape[2];
    WorkOffsets o;
    size_t p = 0;
    o.obs = p;     p = al(p + B * osz * sizeof(float));
    o.hidden = p;  p = al(p + B * d->hidden_dim * sizeof(float));
    o.noise = p;   p = al(p + B * d->action_space_size * sizeof(double));
    o.visits = p;  p = al(p + B * A_cfg * sizeof(int32_t));
    o.probs = p;   p = al(p + B * A_cfg * sizeof(double));
    o.actions = p; p = al(p + B * sizeof(int32_t));
    o.obs_u8 = p;  p = al(p + B * osz);
    o.total = p;
    return o;
}

extern "C" size_t cumind_select_actions_work_bytes(const cumind_net_desc* desc, int32_t B, int32_t A_cfg) {
    if (!desc_valid(desc) || B <= 0 || A_cfg <= 0) return 0;
    return work_offsets(desc, B, A_cfg).total;
}

// device-visible alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, nullptr for pageable memory
static void* mapped_ptr(const void* h) {
    if (!h) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}

static int select_actions_host_impl(cumind_tree* t, const cumind_net* n, const float* h_obs, const uint8_t* h_obs_u8, const cumind_search_cfg* c,
                                    const double* h_noise, int32_t* h_actions, double* h_probs, void* d_work, size_t work_bytes, void* stream);

extern "C" int cumind_select_actions_host(cumind_tree* t, const cumind_net* n, const float* h_obs, const cumind_search_cfg* c,
                                          const double* h_noise, int32_t* h_actions, double* h_probs, void* d_work,
                                          size_t work_bytes, void* stream) {
    if (!h_obs) return fail("cumind_select_actions_host: NULL argument");
    return select_actions_host_impl(t, n, h_obs, nullptr, c, h_noise, h_actions, h_probs, d_work, work_bytes, stream);
}

// The same call for uint8 observations (image frames as the emulator produces them): a quarter of the bytes over PCIe; the
// device computes float32(u8) / 255 correctly rounded, i.e. exactly what the caller's `frames.astype(np.float32) / 255` gives.
extern "C" int cumind_select_actions_host_u8(cumind_tree* t, const cumind_net* n, const uint8_t* h_obs_u8, const cumind_search_cfg* c,
                                             const double* h_noise, int32_t* h_actions, double* h_probs, void* d_work,
                                             size_t work_bytes, void* stream) {
    if (!h_obs_u8) return fail("cumind_select_actions_host_u8: NULL argument");
    return select_actions_host_impl(t, n, nullptr, h_obs_u8, c, h_noise, h_actions, h_probs, d_work, work_bytes, stream);
}

static int select_actions_host_impl(cumind_tree* t, const cumind_net* n, const float* h_obs, const uint8_t* h_obs_u8, const cumind_search_cfg* c,
                                    const double* h_noise, int32_t* h_actions, double* h_probs, void* d_work, size_t work_bytes, void* stream) {
    if (!t || !n || !c || !h_actions || !h_probs || !d_work) return fail("cumind_select_actions_host: NULL argument");
    const int B = t->tv.B, A_cfg = c->action_space_size;
    const WorkOffsets o = work_offsets(&n->desc, B, A_cfg);
    if (work_bytes < o.total) return fail("cumind_select_actions_host: work buffer too small");
    if (c->noise_mode == 1 && !h_noise) return fail("noise_mode 1 needs h_noise");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = (unsigned char*)d_work;
    float* d_obs = (float*)(w + o.obs);
    float* d_hidden = (float*)(w + o.hidden);
    double* d_noise = (double*)(w + o.noise);
    int32_t* d_visits = (int32_t*)(w + o.visits);
    double* d_probs = (double*)(w + o.probs);
    int32_t* d_actions = (int32_t*)(w + o.actions);
    const size_t osz = n->desc.obs_ndim == 1 ? (size_t)n->desc.obs_shape[0]
                                             : (size_t)n->desc.obs_shape[0] * n->desc.obs_shape[1] * n->desc.obs_shape[2];
    // Pinned (mapped) host buffers up to a few MB: the GPU moves the data itself, one kernel each way.  Pageable
    // or large buffers go through cudaMemcpyAsync (copy engines).  Either way the host<->device transfers are part of this call.
    const size_t obs_bytes = (size_t)B * osz * sizeof(float), noise_bytes = (size_t)B * t->tv.A * sizeof(double);
    const bool with_noise = c->noise_mode == 1;
    void *m_obs = mapped_ptr(h_obs), *m_noise = with_noise ? mapped_ptr(h_noise) : nullptr;
    void *m_actions = mapped_ptr(h_actions), *m_probs = mapped_ptr(h_probs);
    if (h_obs_u8) {
        // uint8 frames: the copy engine brings them over in slices on a second stream while the encoder already works on the
        // slices that have landed (the frames of a batch are 20+ MB: 0.4 ms of PCIe time against 0.46 ms of encoder)
        unsigned char* d_u8 = w + o.obs_u8;
        CopyLane* lane = copy_lane(n->device);
        if (!lane) return fail("cumind_select_actions_host_u8: cannot create the copy stream");
        int n_slices = B >= 4 * 128 ? 2 : 1;   // measured at B = 1024: 1.11 ms in one piece, 1.00 in two, 1.07 / 1.01 in three / four
        if (const char* e = getenv("CUMIND_U8_SLICES")) { const int v = atoi(e); if (v >= 1 && v <= 4) n_slices = v; }   // tuning only
        if (with_noise) CM_CUDA(cudaMemcpyAsync(d_noise, h_noise, noise_bytes, cudaMemcpyHostToDevice, st));
        CM_CUDA(cudaEventRecord(lane->ready, st));              // the work buffers are free once the stream's earlier work is done
        CM_CUDA(cudaStreamWaitEvent(lane->stream, lane->ready, 0));
        for (int k = 0; k < n_slices; ++k) {
            const int b0 = (int)((long long)B * k / n_slices), b1 = (int)((long long)B * (k + 1) / n_slices);
            CM_CUDA(cudaMemcpyAsync(d_u8 + (size_t)b0 * osz, h_obs_u8 + (size_t)b0 * osz, (size_t)(b1 - b0) * osz, cudaMemcpyHostToDevice, lane->stream));
            CM_CUDA(cudaEventRecord(lane->landed[k], lane->stream));
        }
        for (int k = 0; k < n_slices; ++k) {
            const int b0 = (int)((long long)B * k / n_slices), b1 = (int)((long long)B * (k + 1) / n_slices);
            CM_CUDA(cudaStreamWaitEvent(st, lane->landed[k], 0));
            CM_CUDA(launch_u8_to_f32(d_u8 + (size_t)b0 * osz, d_obs + (size_t)b0 * osz, (size_t)(b1 - b0) * osz, st));
            g_launches += 1;
            int rc = cumind_initial_inference(n, d_obs + (size_t)b0 * osz, b1 - b0, d_hidden + (size_t)b0 * n->desc.hidden_dim, nullptr, nullptr,
                                              c->net_mode, stream);
            if (rc) return rc;
        }
        int rc = search_impl(t, n, d_hidden, c, with_noise ? d_noise : nullptr, d_visits, d_probs, nullptr, stream, false);
        if (rc) return rc;
        CM_CUDA(launch_argmax_first(d_probs, B, A_cfg, d_actions, st));
        g_launches += 1;
        CM_CUDA(cudaMemcpyAsync(h_actions, d_actions, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaMemcpyAsync(h_probs, d_probs, (size_t)B * A_cfg * sizeof(double), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    const bool zero_copy = m_obs && m_actions && m_probs && (!with_noise || m_noise) && obs_bytes <= ((size_t)4 << 20) && obs_bytes % 16 == 0 &&
                           noise_bytes % 16 == 0 && ((uintptr_t)m_obs | (uintptr_t)m_noise) % 16 == 0;
    // Small vector observations: no staging kernels at all — the encoder reads the observations and the search kernel the
    // noise straight from the mapped host buffers, and the search kernel writes actions and probabilities straight back
    // (posted PCIe writes); two launches instead of four.  Larger inputs are staged by one kernel each way.
    const bool direct = zero_copy && n->desc.obs_ndim == 1 && obs_bytes <= ((size_t)256 << 10) && c->tree_mode == CUMIND_TREE_REFERENCE;
    if (direct) {
        int rc = cumind_initial_inference(n, (const float*)m_obs, B, d_hidden, nullptr, nullptr, c->net_mode, stream);
        if (rc) return rc;
        bool wrote = false;
        rc = search_impl(t, n, d_hidden, c, with_noise ? (const double*)m_noise : nullptr, d_visits, (double*)m_probs, nullptr, stream, true,
                         (int32_t*)m_actions, &wrote);
        if (rc) return rc;
        if (!wrote) {   // per-phase fallback kernels: no action output of their own
            CM_CUDA(launch_argmax_first((const double*)m_probs, B, A_cfg, (int32_t*)m_actions, st));
            g_launches += 1;
        }
        CM_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    if (zero_copy) {
        CM_CUDA(launch_stage_in(m_obs, d_obs, obs_bytes, m_noise, d_noise, with_noise ? noise_bytes : 0, st));
        g_launches += 1;
    } else {
        CM_CUDA(cudaMemcpyAsync(d_obs, h_obs, obs_bytes, cudaMemcpyHostToDevice, st));
        if (with_noise) CM_CUDA(cudaMemcpyAsync(d_noise, h_noise, noise_bytes, cudaMemcpyHostToDevice, st));
    }
    int rc = cumind_initial_inference(n, d_obs, B, d_hidden, nullptr, nullptr, c->net_mode, stream);
    if (rc) return rc;
    rc = search_impl(t, n, d_hidden, c, with_noise ? d_noise : nullptr, d_visits, d_probs, nullptr, stream, n->desc.obs_ndim == 1);
    if (rc) return rc;
    if (zero_copy) {
        CM_CUDA(launch_stage_out(d_probs, B, A_cfg, d_actions, (int32_t*)m_actions, (double*)m_probs, st));
        g_launches += 1;
    } else {
        CM_CUDA(launch_argmax_first(d_probs, B, A_cfg, d_actions, st));
        g_launches += 1;
        CM_CUDA(cudaMemcpyAsync(h_actions, d_actions, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaMemcpyAsync(h_probs, d_probs, (size_t)B * A_cfg * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CM_CUDA(cudaStreamSynchronize(st));
    return 0;
}
