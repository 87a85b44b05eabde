// conv_tc.cu — image encoder on the tensor cores (bf16 mode): every 3x3 convolution of ConvEncoder /
// ResidualBlock (network.py:131-147, 29-45) as an implicit GEMM on tcgen05.
//
//   GEMM view:  D[pixels, Cout] = A[pixels, 9*Cin] . B[9*Cin, Cout]
//   M tile = 128 output pixels (flattened over the image chunk) = the 128 TMEM lanes; thread r builds row
//   r of the im2col operand straight into the UMMA K-major core-matrix layout in shared memory (taps in
//   (ky,kx,ci) order, zero rows for padding taps); the bf16 weights (pre-tiled on the host) are staged
//   once per CTA by TMA bulk copies; K is walked in chunks that accumulate in TMEM;
//   the epilogue (thread = pixel) applies  y = acc*gain + shift  (conv bias and the inference
//   BatchNorm folded per channel), the residual, ReLU, and writes bf16 NHWC.
// Activations between layers are bf16; accumulation and the epilogue are fp32.  rtol 2e-2 vs fp32-exact.
#include <cuda.h>   // CUtensorMap (the driver entry point is fetched at run time: no link dependency on libcuda)

#include "kernels.h"
#include "search_core.cuh"
#include "tc_common.cuh"

namespace cm {

struct ConvTcArgs {
    const float* in_f32;            // first layer: fp32 NHWC observations, or null
    const __nv_bfloat16* in_bf16;   // other layers: bf16 NHWC activations
    __nv_bfloat16* out;             // [n][Ho][Wo][Cout] bf16
    const __nv_bfloat16* res;       // residual (same shape as out) or null
    const unsigned char* wpack;     // bf16 weight tile [Kpad/8][Cout/8][8][8] | gain[Cout] | shift[Cout] (fp32)
    int w_bytes;                    // bytes of the whole pack section of this layer
    int gain_off;                   // byte offset of gain inside the section (shift follows)
    int n_img, Hi, Wi, Cin, Ho, Wo, Cout, stride, relu;
    int K, Kpad, KCH;               // 9*Cin, rounded up to 16, chunk size (multiple of 16)
    int chunked;                    // bf16 activations (out, res, and in of the TMA kernel) in the row-chunked layout below
};

// Row-chunked activation layout [img][y][c/8][x][c%8] (bf16): the 8-channel chunks of one image row are contiguous along x, so
// the halo of a tile is 18 x C/8 rows of 10 x 16 bytes — long rows for the TMA unit — and a warp's stores of one chunk are
// 128 contiguous bytes.  Index of element (pixel row `prow` = img * H + y, x, channel c) with NJ = C / 8 chunks per pixel:
CM_HOST_DEVICE size_t chunked_index(size_t prow, int x, int c, int NJ, int W) {
    return ((prow * NJ + (size_t)(c >> 3)) * W + (size_t)x) * 8 + (size_t)(c & 7);
}

// CIN32: Cin == 32 and the whole K = 288 fits one chunk: all 36 16-byte loads of a row are issued
// before the first dependent shared-memory store (a warp issues in order, so a store waiting for its
// load would otherwise serialise the global-memory latency tap by tap).
template <bool CIN32>
__global__ void __launch_bounds__(128) conv_tc_kernel(const ConvTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* w_s = smem;
    unsigned char* a_s = smem + ((a.w_bytes + 127) & ~127);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(a_s + (size_t)128 * a.KCH * 2);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.Cout) tmem_cols <<= 1;

    if (tid == 0) {
        tc_mbar_init(bar_w, 1);
        tc_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        tc_mbar_expect_tx(bar_w, (uint32_t)a.w_bytes);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < (uint32_t)a.w_bytes; o += CH) {
            const uint32_t n = (uint32_t)a.w_bytes - o < CH ? (uint32_t)a.w_bytes - o : CH;
            tc_bulk_g2s(w_s + o, a.wpack + o, n, bar_w);
        }
    }
    tc_mbar_wait(bar_w, 0);
    const float* gain_s = reinterpret_cast<const float*>(w_s + a.gain_off);
    const float* shift_s = gain_s + a.Cout;
    const uint32_t a_addr = smem_addr(a_s), w_addr = smem_addr(w_s);
    const uint32_t idesc = umma_idesc(128, a.Cout);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t b_lbo = (uint32_t)(a.Cout / 8) * 128;
    uint32_t mma_phase = 0;

    const long long n_pix = (long long)a.n_img * a.Ho * a.Wo;
    const int hw = a.Ho * a.Wo;
    for (long long row0 = (long long)blockIdx.x * 128; row0 < n_pix; row0 += (long long)gridDim.x * 128) {
        const long long pix = row0 + tid;
        const bool live = pix < n_pix;
        int img = 0, oy = 0, ox = 0;
        if (live) { img = (int)(pix / hw); const int rem = (int)(pix - (long long)img * hw); oy = rem / a.Wo; ox = rem - oy * a.Wo; }
        unsigned char* my_row = a_s + ((size_t)(tid >> 3) * 64 + (tid & 7) * 8) * 2;  // + kc*16*128 per 8-wide K chunk
        bool first_mma = true;
        for (int k0 = 0; k0 < a.Kpad; k0 += a.KCH) {
            const int kn = a.Kpad - k0 < a.KCH ? a.Kpad - k0 : a.KCH;  // multiple of 16
            // ---- im2col gather of this row, K range [k0, k0+kn): (tap, ci) advance incrementally, loads are
            // issued in batches of 8 independent 16-byte requests before the shared-memory stores -------------
            const int nkc = kn / 8;
            int tap = k0 / a.Cin, ci = k0 - tap * a.Cin;
            int ky = tap / 3, kx = tap - ky * 3;
            const int iy_base = oy * a.stride - 1, ix_base = ox * a.stride - 1;
            if constexpr (CIN32) {
                // K chunk = 3 whole taps of 32 channels (KCH = 96): 12 independent 16-byte loads, then 12 stores
                const __nv_bfloat16* in = a.in_bf16 + (size_t)img * a.Hi * a.Wi * 32;
                uint4 v[12];
#pragma unroll
                for (int t3 = 0; t3 < 3; ++t3) {
                    const int iy = iy_base + ky, ix = ix_base + kx + t3;   // a chunk is one kernel row: ky fixed, kx = 0..2
                    const bool ok = live && iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
                    const uint4* src = reinterpret_cast<const uint4*>(in + ((size_t)(ok ? iy : 0) * a.Wi + (ok ? ix : 0)) * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[t3 * 4 + j] = ok ? src[j] : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int kc = 0; kc < 12; ++kc) *reinterpret_cast<uint4*>(my_row + (size_t)kc * 2048) = v[kc];
            } else if (a.in_bf16 && (a.Cin & 31) == 0 && ci == 0) {
                // whole taps, Cin a multiple of 32: one bounds test and one base pointer per tap, then
                // Cin/8 independent 16-byte loads in groups of 4
                const __nv_bfloat16* in = a.in_bf16 + (size_t)img * a.Hi * a.Wi * a.Cin;
                const int per_tap = a.Cin >> 3;
                int kc = 0;
                for (; tap < 9 && kc + per_tap <= nkc; ++tap) {
                    const int iy = iy_base + ky, ix = ix_base + kx;
                    const bool ok = live && iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
                    const uint4* src = reinterpret_cast<const uint4*>(in + ((size_t)(ok ? iy : 0) * a.Wi + (ok ? ix : 0)) * a.Cin);
                    for (int j = 0; j < per_tap; j += 4) {
                        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
                        if (ok) { v0 = src[j]; v1 = src[j + 1]; v2 = src[j + 2]; v3 = src[j + 3]; }
                        unsigned char* dst = my_row + (size_t)(kc + j) * 2048;
                        *reinterpret_cast<uint4*>(dst) = v0;
                        *reinterpret_cast<uint4*>(dst + 2048) = v1;
                        *reinterpret_cast<uint4*>(dst + 4096) = v2;
                        *reinterpret_cast<uint4*>(dst + 6144) = v3;
                    }
                    kc += per_tap;
                    if (++kx == 3) { kx = 0; ++ky; }
                }
                for (; kc < nkc; ++kc) *reinterpret_cast<uint4*>(my_row + (size_t)kc * 2048) = make_uint4(0, 0, 0, 0);
            } else if (a.in_bf16 && (a.Cin & 7) == 0) {
                const __nv_bfloat16* in = a.in_bf16 + (size_t)img * a.Hi * a.Wi * a.Cin;
                for (int kc0 = 0; kc0 < nkc; kc0 += 8) {
                    uint4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        v[u] = make_uint4(0, 0, 0, 0);
                        if (kc0 + u < nkc) {
                            const int iy = iy_base + ky, ix = ix_base + kx;
                            if (live && tap < 9 && iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi)
                                v[u] = *reinterpret_cast<const uint4*>(in + ((size_t)iy * a.Wi + ix) * a.Cin + ci);
                            ci += 8;
                            if (ci == a.Cin) { ci = 0; ++tap; if (++kx == 3) { kx = 0; ++ky; } }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (kc0 + u < nkc) *reinterpret_cast<uint4*>(my_row + (size_t)(kc0 + u) * 2048) = v[u];
                }
            } else {
                const float* in = a.in_f32 + (size_t)img * a.Hi * a.Wi * a.Cin;
                for (int kc = 0; kc < nkc; ++kc) {
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int iy = iy_base + ky, ix = ix_base + kx;
                        float v = 0.f;
                        if (live && tap < 9 && iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) v = in[((size_t)iy * a.Wi + ix) * a.Cin + ci];
                        f[i] = v;
                        if (++ci == a.Cin) { ci = 0; ++tap; if (++kx == 3) { kx = 0; ++ky; } }
                    }
                    uint32_t w4[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                        w4[i] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(my_row + (size_t)kc * 2048) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                for (int ks = 0; ks < kn / 16; ++ks) {
                    const uint64_t ad = umma_desc(a_addr + (uint32_t)ks * 2 * 2048, 2048, 128);
                    const uint64_t bd = umma_desc(w_addr + (uint32_t)(k0 / 8 + 2 * ks) * b_lbo, b_lbo, 128);
                    umma_f16(tmem_base, ad, bd, idesc, first_mma ? 0u : 1u);
                    first_mma = false;
                }
                umma_commit(bar_mma);
            }
            tc_mbar_wait(bar_mma, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
        }
        // ---- epilogue: y = acc*gain + shift (+res) -> relu -> bf16 NHWC --------------------------------
        for (int c0 = 0; c0 < a.Cout; c0 += 16) {
            float acc[16];
            tmem_ld16(lane_taddr + c0, acc);
            if (live) {
                const size_t o = (size_t)pix * a.Cout + c0;
                const size_t oc0 = chunked_index((size_t)img * a.Ho + oy, ox, c0, a.Cout >> 3, a.Wo);       // chunk c0/8 (and c0/8 + 1 one row of chunks later)
                const size_t oc1 = oc0 + (size_t)a.Wo * 8;
                float r[16];
                if (a.res) {
                    const uint4 r0 = *reinterpret_cast<const uint4*>(a.res + (a.chunked ? oc0 : o)), r1 = *reinterpret_cast<const uint4*>(a.res + (a.chunked ? oc1 : o + 8));
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i]));
                        r[2 * i] = f2.x; r[2 * i + 1] = f2.y;
                    }
                }
                uint32_t ow[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float y0 = acc[i] * gain_s[c0 + i] + shift_s[c0 + i];
                    float y1 = acc[i + 1] * gain_s[c0 + i + 1] + shift_s[c0 + i + 1];
                    if (a.res) { y0 += r[i]; y1 += r[i + 1]; }
                    if (a.relu) { y0 = y0 > 0.f ? y0 : 0.f; y1 = y1 > 0.f ? y1 : 0.f; }
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                    ow[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(a.out + (a.chunked ? oc0 : o)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                *reinterpret_cast<uint4*>(a.out + (a.chunked ? oc1 : o + 8)) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Stride-1 layers (the residual blocks): NO im2col copy.  The M tile is an 8 (x) by 16 (y) block of
// output pixels; the CTA stages the 18 x 10 input halo once, laid out [yy][k-chunk][xx][8 ch] so that
// the 8 pixels of a UMMA core matrix are 8 consecutive x of one image row.  A tap (ky,kx) is then just
// a shifted view of the same tile: descriptor start = base + ((ky*NJ + 2ks)*10 + kx)*16 B with
// SBO = one halo row (NJ*160 B) and LBO = one k-chunk (160 B), so the 9*Cin/16 MMAs of a tile read
// shared memory that was written once (6.4x less gather traffic than materialising im2col rows).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_tc_halo_kernel(const ConvTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int NJ = a.Cin >> 3;                       // 16-byte k-chunks per pixel
    const int halo_bytes = 18 * NJ * 10 * 16;
    unsigned char* w_s = smem;
    unsigned char* h_s = smem + ((a.w_bytes + 127) & ~127);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(h_s + ((halo_bytes + 127) & ~127));
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.Cout) tmem_cols <<= 1;

    if (tid == 0) {
        tc_mbar_init(bar_w, 1);
        tc_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        tc_mbar_expect_tx(bar_w, (uint32_t)a.w_bytes);
        const uint32_t CH = 32768;
        for (uint32_t o = 0; o < (uint32_t)a.w_bytes; o += CH) {
            const uint32_t n = (uint32_t)a.w_bytes - o < CH ? (uint32_t)a.w_bytes - o : CH;
            tc_bulk_g2s(w_s + o, a.wpack + o, n, bar_w);
        }
    }
    tc_mbar_wait(bar_w, 0);
    const float* gain_s = reinterpret_cast<const float*>(w_s + a.gain_off);
    const float* shift_s = gain_s + a.Cout;
    const uint32_t h_addr = smem_addr(h_s), w_addr = smem_addr(w_s);
    const uint32_t idesc = umma_idesc(128, a.Cout);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t b_lbo = (uint32_t)(a.Cout / 8) * 128;
    const uint32_t sbo = (uint32_t)NJ * 160;
    uint32_t mma_phase = 0;

    const int TX = (a.Wo + 7) / 8, TY = (a.Ho + 15) / 16;
    const int tiles_per_img = TX * TY;
    const long long n_tiles = (long long)a.n_img * tiles_per_img;
    const int n_chunks = 18 * 10 * NJ;
    const int y = tid >> 3, x = tid & 7;
    // this thread's halo chunks are tile-invariant: (yy, xx), destination, source offset relative to the tile origin
    // (Cin <= 32: at most six chunks per thread, kept in registers; the index divisions leave the tile loop)
    const bool small = n_chunks <= 6 * 128;
    int c_dst[6], c_rel[6], c_yx[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
        const int idx = u * 128 + tid;
        c_dst[u] = -1; c_rel[u] = 0; c_yx[u] = 0;
        if (idx < n_chunks) {
            const int pixi = idx / NJ, j = idx - pixi * NJ;
            const int yy = pixi / 10, xx = pixi - yy * 10;
            c_dst[u] = ((yy * NJ + j) * 10 + xx) * 16;
            c_rel[u] = ((yy - 1) * a.Wi + (xx - 1)) * a.Cin + j * 8;
            c_yx[u] = (yy << 8) | xx;
        }
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int img = (int)(tile / tiles_per_img);
        const int trem = (int)(tile - (long long)img * tiles_per_img);
        const int ty = trem / TX, tx = trem - ty * TX;
        const int oy0 = ty * 16, ox0 = tx * 8;
        const __nv_bfloat16* in = a.in_bf16 + (size_t)img * a.Hi * a.Wi * a.Cin;
        // ---- stage the halo: chunk idx = (yy*10 + xx)*NJ + j (a pixel's channels are contiguous in global) ----
        if (small) {
            const __nv_bfloat16* org = in + ((size_t)oy0 * a.Wi + ox0) * a.Cin;
            uint4 v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                v[u] = make_uint4(0, 0, 0, 0);
                if (c_dst[u] >= 0) {
                    const int iy = oy0 + (c_yx[u] >> 8) - 1, ix = ox0 + (c_yx[u] & 255) - 1;
                    if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) v[u] = *reinterpret_cast<const uint4*>(org + c_rel[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 6; ++u)
                if (c_dst[u] >= 0) *reinterpret_cast<uint4*>(h_s + c_dst[u]) = v[u];
        } else {
            for (int idx = tid; idx < n_chunks; idx += 128) {
                const int pixi = idx / NJ, j = idx - pixi * NJ;
                const int yy = pixi / 10, xx = pixi - yy * 10;
                const int iy = oy0 + yy - 1, ix = ox0 + xx - 1;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi)
                    v = *reinterpret_cast<const uint4*>(in + ((size_t)iy * a.Wi + ix) * a.Cin + j * 8);
                *reinterpret_cast<uint4*>(h_s + ((yy * NJ + j) * 10 + xx) * 16) = v;
            }
        }
        // prefetch the residual while the MMAs run
        const int oy = oy0 + y, ox = ox0 + x;
        const bool live = oy < a.Ho && ox < a.Wo;
        const size_t opix = ((size_t)img * a.Ho + oy) * a.Wo + ox;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            bool first = true;
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3, kx = tap - ky * 3;
                for (int ks = 0; ks < NJ / 2; ++ks) {
                    const uint64_t ad = umma_desc(h_addr + (uint32_t)(((ky * NJ + 2 * ks) * 10 + kx) * 16), 160, sbo);
                    const uint64_t bd = umma_desc(w_addr + (uint32_t)(tap * NJ + 2 * ks) * b_lbo, b_lbo, 128);
                    umma_f16(tmem_base, ad, bd, idesc, first ? 0u : 1u);
                    first = false;
                }
            }
            umma_commit(bar_mma);
        }
        tc_mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        for (int c0 = 0; c0 < a.Cout; c0 += 16) {
            float acc[16];
            tmem_ld16(lane_taddr + c0, acc);
            if (live) {
                const size_t o = opix * a.Cout + c0;
                float r[16];
                if (a.res) {
                    const uint4 r0 = *reinterpret_cast<const uint4*>(a.res + o), r1 = *reinterpret_cast<const uint4*>(a.res + o + 8);
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i]));
                        r[2 * i] = f2.x; r[2 * i + 1] = f2.y;
                    }
                }
                uint32_t ow[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float y0 = acc[i] * gain_s[c0 + i] + shift_s[c0 + i];
                    float y1 = acc[i + 1] * gain_s[c0 + i + 1] + shift_s[c0 + i + 1];
                    if (a.res) { y0 += r[i]; y1 += r[i + 1]; }
                    if (a.relu) { y0 = y0 > 0.f ? y0 : 0.f; y1 = y1 > 0.f ? y1 : 0.f; }
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                    ow[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(a.out + o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                *reinterpret_cast<uint4*>(a.out + o + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// The same stride-1 layer as a warp-specialised pipeline (the product path when the tensor map can be built):
//   warp 4 (one elected lane)  PRODUCER + MMA ISSUER.  The 18 x 10 halo of a tile is ONE TMA tensor load
//       (cp.async.bulk.tensor.4d, SASS UTMALDG): the activations are kept in the row-chunked layout [img][y][c/8][x][c%8]
//       (chunked_index) and described to the TMA unit as the 4-D tensor (x*8 + c%8, c/8, y, img), so a box of
//       (80, C/8, 18, 1) — 18 * C/8 rows of 160 contiguous bytes — lands in shared memory exactly in the
//       [yy][k-chunk][xx][8 ch] layout the UMMA descriptors below walk, and everything outside the image is zero-filled by
//       the hardware (no bounds tests, no address arithmetic in the SM).  NSTAGE halo buffers cycle through full / empty
//       mbarriers; the 9 * C/16 tcgen05.mma of a tile go to one of TWO TMEM accumulators; tcgen05.commit releases the
//       halo buffer and publishes the accumulator.
//   warps 0-3                  EPILOGUE.  Thread = output pixel = TMEM lane: prefetch the residual, wait for the accumulator,
//       tcgen05.ld, hand the accumulator back, then gain / shift / residual / ReLU and the bf16 NHWC store — while the
//       MMAs of the next tile and the loads of the one after run.
// ------------------------------------------------------------------------------------------------
constexpr int kConvStages = 3;

CM_DEVICE void tma_load_4d(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_addr(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(bar))
                 : "memory");
}

// KS = Cin / 16 (k-steps per tap): the 9 * KS descriptor pairs of a tile are built ONCE, before the tile loop
template <int KS>
__global__ void __launch_bounds__(160) conv_tc_halo_tma_kernel(const ConvTcArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NJ = 2 * KS;
    const int halo_bytes = 18 * NJ * 10 * 16;
    const int halo_stride = (halo_bytes + 127) & ~127;
    unsigned char* w_s = smem;
    unsigned char* h_s = smem + ((a.w_bytes + 127) & ~127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(h_s + (size_t)kConvStages * halo_stride);
    uint64_t* bar_w = bars;                      // weights staged
    uint64_t* full = bars + 1;                   // [kConvStages] halo landed
    uint64_t* empty = full + kConvStages;        // [kConvStages] halo consumed by the MMAs
    uint64_t* acc_full = empty + kConvStages;    // [2] accumulator complete
    uint64_t* acc_free = acc_full + 2;           // [2] accumulator read by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t acc_cols = 32;
    while ((int)acc_cols < a.Cout) acc_cols <<= 1;
    const uint32_t tmem_cols = 2 * acc_cols;

    if (tid == 0) {
        tc_mbar_init(bar_w, 1);
        for (int i = 0; i < kConvStages; ++i) { tc_mbar_init(full + i, 1); tc_mbar_init(empty + i, 1); }
        for (int i = 0; i < 2; ++i) { tc_mbar_init(acc_full + i, 1); tc_mbar_init(acc_free + i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int TX = (a.Wo + 7) / 8, TY = (a.Ho + 15) / 16;
    const int tiles_per_img = TX * TY;
    const long long n_tiles = (long long)a.n_img * tiles_per_img;

    if (warp == 4) {
        if ((tid & 31) == 0) {
            // ---- producer + MMA issuer ----------------------------------------------------------------------
            tc_mbar_expect_tx(bar_w, (uint32_t)a.w_bytes);
            const uint32_t CH = 32768;
            for (uint32_t o = 0; o < (uint32_t)a.w_bytes; o += CH) {
                const uint32_t n = (uint32_t)a.w_bytes - o < CH ? (uint32_t)a.w_bytes - o : CH;
                tc_bulk_g2s(w_s + o, a.wpack + o, n, bar_w);
            }
            auto issue_load = [&](long long tile, int k) {
                const int s = k % kConvStages;
                if (k >= kConvStages) tc_mbar_wait_spin(empty + s, (uint32_t)((k / kConvStages - 1) & 1));
                const int img = (int)(tile / tiles_per_img);
                const int trem = (int)(tile - (long long)img * tiles_per_img);
                const int ty = trem / TX, tx = trem - ty * TX;
                tc_mbar_expect_tx(full + s, (uint32_t)halo_bytes);
                tma_load_4d(h_s + (size_t)s * halo_stride, &tmap, (tx * 8 - 1) * 8, 0, ty * 16 - 1, img, full + s);
            };
            int k_load = 0;
            long long t_load = blockIdx.x;
            for (; k_load < kConvStages - 1 && t_load < n_tiles; ++k_load, t_load += gridDim.x) issue_load(t_load, k_load);
            tc_mbar_wait(bar_w, 0);
            const uint32_t w_addr = smem_addr(w_s);
            const uint32_t idesc = umma_idesc(128, a.Cout);
            const uint32_t b_lbo = (uint32_t)(a.Cout / 8) * 128;
            const uint32_t sbo = (uint32_t)NJ * 160;
            // A tap (ky, kx), k-step ks of a tile is the halo buffer's descriptor plus a constant: the start address sits in the
            // low 14 bits of the descriptor in 16-byte units and shared-memory addresses stay below 2^18, so the sum never
            // carries into the other fields.  One thread issues all MMAs; built per tile, the descriptors cost ~140 cycles each.
            uint64_t a_off[9 * KS], b_desc[9 * KS];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const int ky = tap / 3, kx = tap - ky * 3;
                    a_off[tap * KS + ks] = (uint64_t)((uint32_t)(((ky * NJ + 2 * ks) * 10 + kx) * 16) >> 4);
                    b_desc[tap * KS + ks] = umma_desc(w_addr + (uint32_t)(tap * NJ + 2 * ks) * b_lbo, b_lbo, 128);
                }
            }
            uint64_t a_stage[kConvStages];
#pragma unroll
            for (int i = 0; i < kConvStages; ++i) a_stage[i] = umma_desc(smem_addr(h_s + (size_t)i * halo_stride), 160, sbo);
            int it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                if (t_load < n_tiles) { issue_load(t_load, k_load); ++k_load; t_load += gridDim.x; }
                const int s = it % kConvStages, b = it & 1;
                tc_mbar_wait_spin(full + s, (uint32_t)((it / kConvStages) & 1));
                if (it >= 2) tc_mbar_wait_spin(acc_free + b, (uint32_t)((it / 2 - 1) & 1));
                tc_fence_after();
                uint64_t ad0 = a_stage[0];
#pragma unroll
                for (int i = 1; i < kConvStages; ++i) ad0 = s == i ? a_stage[i] : ad0;
                const uint32_t tacc = tmem_base + (uint32_t)b * acc_cols;
#pragma unroll
                for (int i = 0; i < 9 * KS; ++i) umma_f16(tacc, ad0 + a_off[i], b_desc[i], idesc, i > 0 ? 1u : 0u);
                umma_commit(empty + s);      // the halo buffer may be refilled
                umma_commit(acc_full + b);   // the accumulator is complete
            }
        }
    } else {
        // ---- epilogue warps: thread = pixel (y, x) of the 16 x 8 tile = TMEM lane ---------------------------------
        tc_mbar_wait(bar_w, 0);   // gain / shift arrive with the weights
        const float* gain_s = reinterpret_cast<const float*>(w_s + a.gain_off);
        const float* shift_s = gain_s + a.Cout;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int y = tid >> 3, x = tid & 7;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const int img = (int)(tile / tiles_per_img);
            const int trem = (int)(tile - (long long)img * tiles_per_img);
            const int ty = trem / TX, tx = trem - ty * TX;
            const int oy = ty * 16 + y, ox = tx * 8 + x;
            const bool live = oy < a.Ho && ox < a.Wo;
            tc_mbar_wait_spin(acc_full + b, (uint32_t)((it / 2) & 1));
            tc_fence_after();
            for (int c0 = 0; c0 < a.Cout; c0 += 32) {
                // residual of these 32 channels requested before the accumulator is read
                uint4 rr[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                // chunk q of these 32 channels: row-chunked layout, one row of chunks = Wo * 8 elements apart
                const size_t o = chunked_index((size_t)img * a.Ho + oy, ox, c0, a.Cout >> 3, a.Wo);
                const size_t cs = (size_t)a.Wo * 8;
                if (a.res && live) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) rr[q] = *reinterpret_cast<const uint4*>(a.res + o + q * cs);
                }
                float acc[32];
                {
                    float lo[16], hi[16];
                    tmem_ld16(lane_taddr + (uint32_t)b * acc_cols + c0, lo);
                    tmem_ld16(lane_taddr + (uint32_t)b * acc_cols + c0 + 16, hi);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { acc[i] = lo[i]; acc[16 + i] = hi[i]; }
                }
                if (c0 + 32 >= a.Cout) {   // last read of this accumulator: hand it back to the MMA issuer
                    tc_fence_before();
                    tc_mbar_arrive(acc_free + b);
                }
                if (live) {
                    uint32_t ow[16];
                    const uint32_t rw[16] = {rr[0].x, rr[0].y, rr[0].z, rr[0].w, rr[1].x, rr[1].y, rr[1].z, rr[1].w,
                                             rr[2].x, rr[2].y, rr[2].z, rr[2].w, rr[3].x, rr[3].y, rr[3].z, rr[3].w};
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 g2 = *reinterpret_cast<const float2*>(gain_s + c0 + i), s2 = *reinterpret_cast<const float2*>(shift_s + c0 + i);
                        float y0 = acc[i] * g2.x + s2.x;
                        float y1 = acc[i + 1] * g2.y + s2.y;
                        if (a.res) {
                            const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[i / 2]));
                            y0 += f2.x; y1 += f2.y;
                        }
                        if (a.relu) { y0 = y0 > 0.f ? y0 : 0.f; y1 = y1 > 0.f ? y1 : 0.f; }
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                        ow[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(a.out + o + q * cs) = make_uint4(ow[4 * q], ow[4 * q + 1], ow[4 * q + 2], ow[4 * q + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// The first layer (network.py:137: 3x3, stride 2, pad 1, fp32 observations with a handful of channels): K = 9 * C is one
// or two k-steps, so the layer is bound by getting the observations in.  Tile = 16 x 8 output pixels; its 33 x 17 input
// patch is read ONCE with coalesced loads (an input row segment of 17 pixels is contiguous in NHWC) into shared memory,
// every thread then assembles its im2col row (taps in (ky, kx, ci) order) from there, converts to bf16 and writes it in the
// UMMA K-major core-matrix layout.  (The generic kernel reads each input element up to 2.25 times, four bytes at a time.)
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) conv_tc_first_kernel(const ConvTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int PW = 17 * C;                        // floats per patch row
    constexpr int NU = (PW + 31) / 32;                // patch columns per lane
    constexpr int KR = 9 * C, KPAD = (KR + 15) & ~15; // im2col row: for ky: 3*C contiguous floats of patch row 2y + ky
    unsigned char* w_s = smem;
    unsigned char* a_s = smem + ((a.w_bytes + 127) & ~127);
    float* patch = reinterpret_cast<float*>(a_s + (size_t)128 * a.Kpad * 2);   // [33][17 * C]
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(patch) + (((size_t)33 * PW * 4 + 127) & ~(size_t)127));
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < a.Cout) tmem_cols <<= 1;
    if (tid == 0) {
        tc_mbar_init(bar_w, 1);
        tc_mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        tc_mbar_expect_tx(bar_w, (uint32_t)a.w_bytes);
        tc_bulk_g2s(w_s, a.wpack, (uint32_t)a.w_bytes, bar_w);
    }
    tc_mbar_wait(bar_w, 0);
    const float* gain_s = reinterpret_cast<const float*>(w_s + a.gain_off);
    const float* shift_s = gain_s + a.Cout;
    const uint32_t a_addr = smem_addr(a_s), w_addr = smem_addr(w_s);
    const uint32_t idesc = umma_idesc(128, a.Cout);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t b_lbo = (uint32_t)(a.Cout / 8) * 128;
    uint32_t mma_phase = 0;
    const int TX = (a.Wo + 7) / 8, TY = (a.Ho + 15) / 16;
    const int tiles_per_img = TX * TY;
    const int n_tiles = a.n_img * tiles_per_img;
    const int y = tid >> 3, x = tid & 7;
    unsigned char* my_row = a_s + ((size_t)(tid >> 3) * 64 + (tid & 7) * 8) * 2;
    // patch load: warp w takes rows w, w+4, ...; a lane's columns cc = lane, lane + 32, ... (and their pixel offsets) are fixed
    const int lane = tid & 31;
    int px[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) px[u] = (lane + 32 * u) / C;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int img = tile / tiles_per_img, trem = tile - img * tiles_per_img;
        const int ty = trem / TX, tx = trem - ty * TX;
        const int oy0 = ty * 16, ox0 = tx * 8;
        const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
        const float* in = a.in_f32 + (size_t)img * a.Hi * a.Wi * C;
        // ---- the input patch, zero outside the image ----------------------------------------------------------------
        // all loads of a lane are issued before its first store: the patch costs one memory round trip, not nine
        float v[9][NU];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int r = warp + 4 * q;
            const int iy = iy0 + r;
            const bool row_ok = r < 33 && iy >= 0 && iy < a.Hi;
            const float* src = in + ((long long)iy * a.Wi + ix0) * C;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int cc = lane + 32 * u;
                const int ix = ix0 + px[u];
                v[q][u] = (row_ok && cc < PW && ix >= 0 && ix < a.Wi) ? src[cc] : 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int r = warp + 4 * q;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int cc = lane + 32 * u;
                if (r < 33 && cc < PW) patch[r * PW + cc] = v[q][u];
            }
        }
        __syncthreads();
        // ---- im2col row of pixel (y, x), zero padded to Kpad ------------------------------------------------------------
        {
            const float* p0 = patch + (2 * y) * PW + (2 * x) * C;
            float f[KPAD];
#pragma unroll
            for (int k = 0; k < KPAD; ++k) f[k] = 0.f;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int j = 0; j < 3 * C; ++j) f[ky * 3 * C + j] = p0[ky * PW + j];   // taps (ky, kx = 0..2) x channels are contiguous
#pragma unroll
            for (int kc = 0; kc < KPAD / 8; ++kc) {
                uint32_t w4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[kc * 8 + 2 * i], f[kc * 8 + 2 * i + 1]);
                    w4[i] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(my_row + (size_t)kc * 2048) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < a.Kpad / 16; ++ks) {
                const uint64_t ad = umma_desc(a_addr + (uint32_t)ks * 2 * 2048, 2048, 128);
                const uint64_t bd = umma_desc(w_addr + (uint32_t)(2 * ks) * b_lbo, b_lbo, 128);
                umma_f16(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar_mma);
        }
        tc_mbar_wait_spin(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        const int oy = oy0 + y, ox = ox0 + x;
        const bool live = oy < a.Ho && ox < a.Wo;
        for (int c0 = 0; c0 < a.Cout; c0 += 16) {
            float acc[16];
            tmem_ld16(lane_taddr + c0, acc);
            if (live) {
                uint32_t ow[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 g2 = *reinterpret_cast<const float2*>(gain_s + c0 + i), s2 = *reinterpret_cast<const float2*>(shift_s + c0 + i);
                    float y0 = acc[i] * g2.x + s2.x;
                    float y1 = acc[i + 1] * g2.y + s2.y;
                    if (a.relu) { y0 = y0 > 0.f ? y0 : 0.f; y1 = y1 > 0.f ? y1 : 0.f; }
                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                    ow[i / 2] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                const size_t prow = (size_t)img * a.Ho + oy;
                const size_t o0 = a.chunked ? chunked_index(prow, ox, c0, a.Cout >> 3, a.Wo) : (prow * a.Wo + ox) * a.Cout + c0;
                const size_t o1 = a.chunked ? o0 + (size_t)a.Wo * 8 : o0 + 8;
                *reinterpret_cast<uint4*>(a.out + o0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                *reinterpret_cast<uint4*>(a.out + o1) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// mean over pixels (fp32 accumulation, tree reduction over 8 pixel slices) then final_dense in fp32.
// NHWC: thread = (slice, channel).  Row-chunked layout: warp = (channel chunk, slice), its lanes walk the pixels of the
// chunk with 16-byte loads (an image row of one chunk is Wo * 16 contiguous bytes), 8 channel sums per lane, shuffle tree.
__global__ void __launch_bounds__(256) pool_dense_bf16_kernel(const __nv_bfloat16* __restrict__ x, int npix, int Cc,
                                                             const float* __restrict__ K, const float* __restrict__ b, int H,
                                                             float* __restrict__ hidden, int chunked, int Wo) {
    extern __shared__ __align__(16) float sm[];  // [8][Cc] partial sums, then pooled [Cc]
    const int img = blockIdx.x;
    const __nv_bfloat16* xi = x + (size_t)img * npix * Cc;
    const int slices = 8;
    if (chunked) {
        const int NJ = Cc >> 3, Ho = npix / Wo;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int idx = threadIdx.x; idx < slices * Cc; idx += blockDim.x) sm[idx] = 0.0f;
        __syncthreads();
        for (int job = warp; job < NJ * slices; job += nwarps) {     // (chunk j, slice sl): rows y = sl, sl + slices, ...
            const int j = job % NJ, sl = job / NJ;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            // the (row, pixel) pairs of this slice as one index space, eight 16-byte loads of a lane in flight before the first
            // is used (a row of 42 pixels leaves a third of the lanes idle on its second pass otherwise)
            const int rows = (Ho - sl + slices - 1) / slices, items = rows * Wo;
            for (int i0 = lane; i0 < items; i0 += 32 * 8) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + 32 * u;
                    v[u] = make_uint4(0, 0, 0, 0);
                    if (i < items) {
                        const int r = i / Wo, xx = i - r * Wo;
                        v[u] = *reinterpret_cast<const uint4*>(xi + (((size_t)(sl + r * slices) * NJ + j) * Wo + xx) * 8);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[q]));
                        acc[2 * q] += f2.x; acc[2 * q + 1] += f2.y;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float v = acc[q];
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) sm[sl * Cc + j * 8 + q] = v;
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < slices * Cc; idx += blockDim.x) {
            const int sl = idx / Cc, c = idx - sl * Cc;
            float s = 0.f;
            for (int i = sl; i < npix; i += slices) s += __bfloat162float(xi[(size_t)i * Cc + c]);
            sm[idx] = s;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Cc; c += blockDim.x) {
        float s = 0.f;
        for (int sl = 0; sl < slices; ++sl) s += sm[sl * Cc + c];
        sm[slices * Cc + c] = s / (float)npix;
    }
    __syncthreads();
    const float* pooled = sm + slices * Cc;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < Cc; ++k) acc = fmaf(pooled[k], K[(size_t)k * H + j], acc);
        hidden[(size_t)img * H + j] = acc + b[j];
    }
}

static inline int out_dim2(int n) { return (n - 1) / 2 + 1; }

bool conv_tc_supported(const cumind_net_desc& d) {
    return d.obs_ndim == 3 && d.conv_channels % 16 == 0 && d.conv_channels <= 256 && d.obs_shape[2] <= 512;
}

ConvTcPackLayout make_conv_tc_pack_layout(const cumind_net_desc& d) {
    ConvTcPackLayout p{};
    const int C = d.obs_shape[2], Cc = d.conv_channels, L = d.num_blocks;
    p.n_layers = 1 + 2 * L;
    size_t o = 0;
    for (int i = 0; i < p.n_layers && i < ConvTcPackLayout::kMaxLayers; ++i) {
        const int cin = i == 0 ? C : Cc;
        const int K = 9 * cin, Kpad = (K + 15) & ~15;
        p.K[i] = K; p.Kpad[i] = Kpad; p.cin[i] = cin;
        p.off[i] = o;
        const size_t wbytes = ((size_t)Kpad * Cc * 2 + 127) & ~(size_t)127;
        p.gain_off[i] = wbytes;
        p.bytes[i] = (wbytes + (size_t)2 * Cc * 4 + 127) & ~(size_t)127;
        o += p.bytes[i];
    }
    p.total_bytes = o;
    return p;
}

size_t conv_tc_work_bytes(const cumind_net_desc& d, int chunk) {
    const size_t npix = (size_t)out_dim2(d.obs_shape[0]) * out_dim2(d.obs_shape[1]);
    return 3 * (((size_t)chunk * npix * d.conv_channels * 2 + 255) & ~(size_t)255);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// row-chunked bf16 activations [n][Hi][C/8][Wi][8] as the 4-D tensor (x*8 + c%8, c/8, y, img): a box (80, C/8, 18, 1) is the halo of a tile
static bool make_halo_tensor_map(const ConvTcArgs& a, CUtensorMap* out) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t NJ = (cuuint64_t)a.Cin / 8, W = (cuuint64_t)a.Wi;
    const cuuint64_t dims[4] = {W * 8, NJ, (cuuint64_t)a.Hi, (cuuint64_t)a.n_img};
    const cuuint64_t strides[3] = {W * 16, NJ * W * 16, (cuuint64_t)a.Hi * NJ * W * 16};   // bytes, dims 1..3
    const cuuint32_t box[4] = {80, (cuuint32_t)NJ, 18, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(a.in_bf16), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// the whole trunk runs on the TMA kernel (and keeps its activations row-chunked) when every residual layer qualifies
bool conv_tc_chunked_ok(int Cc, int Ho, int Wo) {
    (void)Ho; (void)Wo;
    return (Cc == 32 || Cc == 64 || Cc == 96 || Cc == 128) && encode_tiled_fn() != nullptr && !getenv("CUMIND_CONV_NO_TMA");
}

static cudaError_t run_conv_tc_halo(const ConvTcArgs& a, int sm_count, cudaStream_t st) {
    const int NJ = a.Cin >> 3;
    CUtensorMap tmap;
    if (a.chunked) {
        if (!make_halo_tensor_map(a, &tmap)) return cudaErrorInvalidValue;
        const size_t halo_stride = ((size_t)(18 * NJ * 10 * 16) + 127) & ~(size_t)127;
        const size_t smem_t = ((a.w_bytes + 127) & ~127) + kConvStages * halo_stride + 16 * 8 + 64;
        if (smem_t <= 227 * 1024) {
            void (*kern)(const ConvTcArgs, const CUtensorMap) = nullptr;
            switch (a.Cin / 16) {
                case 2: kern = conv_tc_halo_tma_kernel<2>; break;
                case 4: kern = conv_tc_halo_tma_kernel<4>; break;
                case 6: kern = conv_tc_halo_tma_kernel<6>; break;
                case 8: kern = conv_tc_halo_tma_kernel<8>; break;
                default: return cudaErrorInvalidValue;
            }
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
            if (e != cudaSuccess) return e;
            const long long tiles = (long long)a.n_img * ((a.Wo + 7) / 8) * ((a.Ho + 15) / 16);
            int acc_cols = 32;
            while (acc_cols < a.Cout) acc_cols <<= 1;
            int per_sm = (int)((227 * 1024) / (smem_t + 1024));
            if (per_sm > 512 / (2 * acc_cols)) per_sm = 512 / (2 * acc_cols);
            if (per_sm > 4) per_sm = 4;
            if (per_sm < 1) per_sm = 1;
            if (const char* ev = getenv("CUMIND_CONV_CTAS")) per_sm = atoi(ev) > 0 ? atoi(ev) : per_sm;   // tuning only
            const long long cap = (long long)sm_count * per_sm;
            kern<<<(int)(tiles < cap ? tiles : cap), 160, smem_t, st>>>(a, tmap);
            return cudaGetLastError();
        }
        return cudaErrorInvalidValue;
    }
    const size_t smem = ((a.w_bytes + 127) & ~127) + ((size_t)(18 * NJ * 10 * 16 + 127) & ~(size_t)127) + 64;
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long tiles = (long long)a.n_img * ((a.Wo + 7) / 8) * ((a.Ho + 15) / 16);
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sm_count * per_sm;
    conv_tc_halo_kernel<<<(int)(tiles < cap ? tiles : cap), 128, smem, st>>>(a);
    return cudaGetLastError();
}

static cudaError_t run_conv_tc(ConvTcArgs a, int sm_count, cudaStream_t st) {
    if (a.in_bf16 && a.stride == 1 && (a.Cin & 15) == 0 && a.Cin <= 128 && a.Hi == a.Ho && a.Wi == a.Wo)
        return run_conv_tc_halo(a, sm_count, st);
    if (a.in_f32 && a.stride == 2 && (a.Cin == 1 || a.Cin == 3 || a.Cin == 4) && a.Kpad == ((9 * a.Cin + 15) & ~15) && a.res == nullptr &&
        a.w_bytes <= 32768 && !getenv("CUMIND_CONV_NO_FIRST")) {
        void (*kern)(const ConvTcArgs) = a.Cin == 1 ? conv_tc_first_kernel<1> : (a.Cin == 3 ? conv_tc_first_kernel<3> : conv_tc_first_kernel<4>);
        const size_t smem_f = ((a.w_bytes + 127) & ~127) + (size_t)128 * a.Kpad * 2 + (((size_t)33 * 17 * a.Cin * 4 + 127) & ~(size_t)127) + 64 +
                              (size_t)a.Kpad * 4;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
        if (e != cudaSuccess) return e;
        const long long tiles = (long long)a.n_img * ((a.Wo + 7) / 8) * ((a.Ho + 15) / 16);
        int per_sm = (int)((227 * 1024) / (smem_f + 1024));
        if (per_sm > 8) per_sm = 8;
        int cols = 32;
        while (cols < a.Cout) cols <<= 1;
        if (per_sm > 512 / cols) per_sm = 512 / cols;
        if (per_sm < 1) per_sm = 1;
        const long long cap = (long long)sm_count * per_sm;
        kern<<<(int)(tiles < cap ? tiles : cap), 128, smem_f, st>>>(a);
        return cudaGetLastError();
    }
    const bool cin32 = a.in_bf16 && a.Cin == 32 && a.Kpad == 288;
    a.KCH = cin32 ? 96 : (a.Kpad < 384 ? a.Kpad : 384);
    const size_t smem = ((a.w_bytes + 127) & ~127) + (size_t)128 * a.KCH * 2 + 64;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long n_pix = (long long)a.n_img * a.Ho * a.Wo;
    const long long tiles = (n_pix + 127) / 128;
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    const long long cap = (long long)sm_count * per_sm;
    const int grid = (int)(tiles < cap ? tiles : cap);
    if (cin32)
        conv_tc_kernel<true><<<grid, 128, smem, st>>>(a);
    else
        conv_tc_kernel<false><<<grid, 128, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_conv_encoder_tc(const NetView& nv, const unsigned char* cpack, const ConvTcPackLayout& cpk, int sm_count,
                                   const float* obs, int B, float* hidden, unsigned char* work, int chunk, cudaStream_t st) {
    const cumind_net_desc& d = nv.desc;
    const int Hi = d.obs_shape[0], Wi = d.obs_shape[1], C = d.obs_shape[2], Cc = d.conv_channels, L = d.num_blocks, H = d.hidden_dim;
    const int Ho = out_dim2(Hi), Wo = out_dim2(Wi);
    const size_t npix = (size_t)Ho * Wo;
    const size_t buf = ((size_t)chunk * npix * Cc * 2 + 255) & ~(size_t)255;
    __nv_bfloat16* x = reinterpret_cast<__nv_bfloat16*>(work);
    __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(work + buf);
    __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(work + 2 * buf);
    // final_dense parameters live in the canonical fp32 blob after the conv / bn parameters
    const float* fd = nv.blob + nv.lo.enc + ((size_t)9 * C * Cc + Cc) + (size_t)L * (2 * ((size_t)9 * Cc * Cc + Cc) + 8 * Cc);
    const int chunked = conv_tc_chunked_ok(Cc, Ho, Wo) ? 1 : 0;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int n = B - b0 < chunk ? B - b0 : chunk;
        auto layer = [&](int i, const float* in_f32, const __nv_bfloat16* in_bf16, __nv_bfloat16* out, const __nv_bfloat16* res,
                         int hi, int wi, int cin, int stride) {
            ConvTcArgs a{};
            a.in_f32 = in_f32; a.in_bf16 = in_bf16; a.out = out; a.res = res;
            a.wpack = cpack + cpk.off[i]; a.w_bytes = (int)cpk.bytes[i]; a.gain_off = (int)cpk.gain_off[i];
            a.n_img = n; a.Hi = hi; a.Wi = wi; a.Cin = cin; a.Ho = Ho; a.Wo = Wo; a.Cout = Cc; a.stride = stride; a.relu = 1;
            a.K = cpk.K[i]; a.Kpad = cpk.Kpad[i];
            a.chunked = chunked;
            return run_conv_tc(a, sm_count, st);
        };
        cudaError_t e = layer(0, obs + (size_t)b0 * Hi * Wi * C, nullptr, x, nullptr, Hi, Wi, C, 2);
        if (e != cudaSuccess) return e;
        for (int l = 0; l < L; ++l) {
            if ((e = layer(1 + 2 * l, nullptr, x, t, nullptr, Ho, Wo, Cc, 1)) != cudaSuccess) return e;
            if ((e = layer(2 + 2 * l, nullptr, t, u, x, Ho, Wo, Cc, 1)) != cudaSuccess) return e;
            __nv_bfloat16* tmp = x; x = u; u = tmp;
        }
        pool_dense_bf16_kernel<<<n, 256, (size_t)9 * Cc * sizeof(float), st>>>(x, (int)npix, Cc, fd, fd + (size_t)Cc * H, H,
                                                                            hidden + (size_t)b0 * H, chunked, Wo);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace cm
