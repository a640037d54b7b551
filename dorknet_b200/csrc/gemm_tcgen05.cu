// gemm_tcgen05.cu -- tensor-core GEMMs for pointwise conv / dense on sm_100a:
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared-memory stages -> tcgen05.mma kind::tf32 with the
// accumulator in TMEM -> tcgen05.ld epilogue with coalesced NCHW stores.
//
// One persistent, warp-specialised kernel (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one
// elected lane), warps 2-5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31).  Three mbarrier rings:
// full/empty per shared-memory stage (TMA <-> MMA) and full/empty per TMEM accumulator (MMA <-> epilogue; two
// accumulators, so the epilogue of tile t overlaps the loads and MMAs of tile t+1).
//
// The NCHW tensors are used in place -- no NHWC copies as in the reference (pointwise_convolution.py:46-55):
//   forward  Y[n][F,HW]  = W[F,C] . X[n][C,HW]    M = pixels (A = X[n], MN-major), N = F (B = W, K-major),  K = C
//   dgrad    dX[n][C,HW] = W^T . dY[n][F,HW]      M = pixels (A = dY[n], MN-major), N = C (B = W, MN-major), K = F
//   wgrad    dW[F,C]     = sum_n dY[n] . X[n]^T   M = F (A = dY[n], K-major), N = C (B = X[n], K-major), K = (n, hw)
// Putting the pixels on M makes TMEM lane i <-> pixel i, so for every output channel a warp stores 32 consecutive
// floats (128 B): the epilogue is coalesced straight from registers.  fp32 MN-major operands use the
// 128B-swizzle-with-32B-atom layout (the only MN-major layout the tensor core accepts for 32-bit types).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace dk { extern int g_bn_fused_enabled; extern int g_bn_split_ctas_per_sm; }

namespace dk {

using namespace tc;

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;                 // MMA M
constexpr int TC_BK = 32;                  // floats of K per stage (= one 128-byte swizzle span)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_BUDGET = 200 * 1024;
constexpr int TC_XR_MAX = 8;                 // slots of the epilogue-operand ring
constexpr uint32_t TC_X_SLOT = 128 * 32 * 4;  // 16 KB: four [32 channels][32 pixels] boxes

struct TcParams {
    int mode;       // 0: tile = (batch, m-block, n-block), K loop over k-blocks.  1: tile = (m-block, n-block, split), K loop over (batch, k-block) items
    int a_mn, b_mn;  // 1 = operand is MN-major in memory (the GEMM's M / N index is the contiguous one)
    int b_batched;   // mode 0: does B have a batch coordinate (0 for weights)
    int M, N, K;     // per-batch GEMM extents (mode 1: K = per-image reduction length)
    int batches;
    int bn;          // MMA N (multiple of 32, <= 256)
    int m_blocks, n_blocks, k_blocks;
    int splits, items_per_split, total_items;  // mode 1
    int kpi;         // mode 1, both operands through TMA: 32-wide k-blocks per pipeline item (k_blocks then counts ITEMS per image)
    int num_tiles;
    int stages;
    int epi;         // 0: out[(b*N + n)*ldo + m] (+bias[n]);  1: out[(split*M + m)*N + n];  2: zero-stuffed strided scatter
    float *out;
    const float *bias;
    int ldo;
    int epi_ow, epi_s;  // epi 2: output-pixel row length and the stride of the zero-stuffed scatter
    const float *epi_w;  // epi 1: optional "+ epi_l2 * epi_w[m*N + n]" (weight decay folded into a dense wgrad)
    float epi_l2;
    // epi 0: optional "+ epi_cb[n] * epi_x[(b*N + n)*ldo + m] + epi_cd[n]" -- the backward of a BatchNorm folded into this
    // dgrad GEMM (bn_fold.cu): epi_x is the BatchNorm's input, laid out like the output
    const float *epi_x, *epi_cb, *epi_cd;
    // epi 1, split-K (mode 1) with every tile of the grid resident at once: the CTAs of an output tile add their partial
    // sums themselves once all of them have arrived (each reduces 1/splits of the tile) -- no reduce kernel behind the GEMM
    unsigned int *sk_counter;  // [2 * m_blocks * n_blocks], zero between launches; nullptr = partials only
    float *sk_out;
    const float *sk_w;
    float sk_l2;
    int xring;       // > 0: epi_x tiles are staged by TMA into a ring of `xring` 16 KB slots (128 pixels x 32 channels) instead of
                     // being loaded by the epilogue warps themselves
    uint32_t tmem_cols, acc_stride;
    uint32_t a_tx;   // bytes one stage's A tile receives from TMA (a K-major A with fewer than 128 rows loads only those)
    // MN-major shared-memory descriptor fields (bytes) -- runtime so a bring-up probe can sweep them
    uint32_t mn_layout, mn_lbo, mn_sbo, mn_kstep;
};

// ---- software-gather operand loaders ---------------------------------------------------------------------
// Operands TMA cannot describe (stride-2 subsampling, 7x7 planes whose row pitch is not a multiple of 16 bytes,
// im2col patches) are written into the SAME swizzled stage layout by four loader warps (warps 6-9), which then
// fence.proxy.async and arrive on the stage's full barrier next to the TMA transaction count.
//   MN-major functor:  Row row(int b, int m) const;   float load(const Row &, int k) const;
//   K-major functor:   KCol kcol(int b, int k) const; float load(int r, const KCol &) const;
struct NoGather {
    static constexpr bool kGather = false;
    static constexpr bool kMN = false;
};

// pixels of a [B, K, SH, SW] tensor subsampled by s: element (b, m = oh*OW + ow, k) = src[b][k][oh*s][ow*s]
struct PixelGatherMN {
    static constexpr bool kGather = true;
    static constexpr bool kMN = true;
    const float *src;
    int K, SH, SW, OW, M, s;
    struct Row { long long off; bool ok; };
    __device__ __forceinline__ Row row(int b, int m) const {
        Row r;
        r.ok = m < M;
        const int oh = m / OW, ow = m - oh * OW;
        r.off = ((long long)b * K * SH + (long long)oh * s) * SW + (long long)ow * s;
        return r;
    }
    __device__ __forceinline__ const float *ptr(const Row &r, int k) const {
        return (r.ok && k < K) ? src + r.off + (long long)k * SH * SW : nullptr;
    }
};
// same tensor as a K-major operand: rows r = channel, k = pixel index inside image b
struct PixelGatherKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    const float *src;
    int R, SH, SW, OW, P, s;  // R channels, P = OH*OW pixels per image
    struct KCol { long long off; bool ok; };
    __device__ __forceinline__ KCol kcol(int b, int k) const {
        KCol c;
        c.ok = k < P;
        const int oh = k / OW, ow = k - oh * OW;
        c.off = ((long long)b * R * SH + (long long)oh * s) * SW + (long long)ow * s;
        return c;
    }
    __device__ __forceinline__ const float *ptr(int r, const KCol &c) const {
        return (c.ok && r < R) ? src + c.off + (long long)r * SH * SW : nullptr;
    }
};

struct ConvGeom {
    int C, H, W, F, kh, kw, s, p, OH, OW;
};
// im2col patches as the MN-major A of the forward GEMM: (b, m = oh*OW+ow, k = (c*kh+i)*kw+j) (im2col.pyx:33-34)
struct ConvPatchMN {
    static constexpr bool kGather = true;
    static constexpr bool kMN = true;
    const float *x;
    ConvGeom g;
    struct Row { long long base; int ih0, iw0; bool ok; };
    __device__ __forceinline__ Row row(int b, int m) const {
        Row r;
        r.ok = m < g.OH * g.OW;
        const int oh = m / g.OW, ow = m - oh * g.OW;
        r.base = (long long)b * g.C * g.H * g.W;
        r.ih0 = oh * g.s - g.p;
        r.iw0 = ow * g.s - g.p;
        return r;
    }
    __device__ __forceinline__ const float *ptr(const Row &r, int k) const {
        const int kk = g.kh * g.kw;
        const int c = k / kk, t = k - c * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ih = r.ih0 + i, iw = r.iw0 + j;
        if (!r.ok || c >= g.C || ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return nullptr;
        return x + r.base + ((long long)c * g.H + ih) * g.W + iw;
    }
};
// the same patches as the K-major B of the wgrad GEMM: rows r = (c,i,j), k = output pixel of image b
struct ConvPatchKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    const float *x;
    ConvGeom g;
    struct KCol { long long base; int ih0, iw0; bool ok; };
    __device__ __forceinline__ KCol kcol(int b, int k) const {
        KCol c;
        c.ok = k < g.OH * g.OW;
        const int oh = k / g.OW, ow = k - oh * g.OW;
        c.base = (long long)b * g.C * g.H * g.W;
        c.ih0 = oh * g.s - g.p;
        c.iw0 = ow * g.s - g.p;
        return c;
    }
    __device__ __forceinline__ const float *ptr(int r, const KCol &kc) const {
        const int kk = g.kh * g.kw;
        const int c = r / kk, t = r - c * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ih = kc.ih0 + i, iw = kc.iw0 + j;
        if (!kc.ok || c >= g.C || ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return nullptr;
        return x + kc.base + ((long long)c * g.H + ih) * g.W + iw;
    }
};
// a dense row-major matrix [R][K] as a K-major operand (filters W[F][C*kh*kw] whose pitch TMA cannot take)
struct MatrixKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    const float *w;
    int R, K;
    struct KCol { int k; bool ok; };
    __device__ __forceinline__ KCol kcol(int, int k) const { return KCol{k, k < K}; }
    __device__ __forceinline__ const float *ptr(int r, const KCol &c) const {
        return (c.ok && r < R) ? w + (long long)r * K + c.k : nullptr;
    }
};
// dgrad of a general convolution, gather form of col2im (im2col.pyx:209-234): A(b, m = input pixel, k = (f,i,j))
struct ConvDgradMN {
    static constexpr bool kGather = true;
    static constexpr bool kMN = true;
    const float *dy;
    ConvGeom g;
    struct Row { long long base; int hp, wp; bool ok; };
    __device__ __forceinline__ Row row(int b, int m) const {
        Row r;
        r.ok = m < g.H * g.W;
        const int h = m / g.W, w = m - h * g.W;
        r.base = (long long)b * g.F * g.OH * g.OW;
        r.hp = h + g.p;
        r.wp = w + g.p;
        return r;
    }
    __device__ __forceinline__ const float *ptr(const Row &r, int k) const {
        const int kk = g.kh * g.kw;
        const int f = k / kk, t = k - f * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ti = r.hp - i, tj = r.wp - j;
        if (!r.ok || f >= g.F || ti < 0 || tj < 0 || (ti % g.s) != 0 || (tj % g.s) != 0) return nullptr;
        const int oh = ti / g.s, ow = tj / g.s;
        if (oh >= g.OH || ow >= g.OW) return nullptr;
        return dy + r.base + ((long long)f * g.OH + oh) * g.OW + ow;
    }
};
// ... and its B(n = c, k = (f,i,j)) = W[f][c][i][j] as a K-major operand
struct ConvDgradWKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    const float *w;
    ConvGeom g;
    struct KCol { long long off; bool ok; };
    __device__ __forceinline__ KCol kcol(int, int k) const {
        const int kk = g.kh * g.kw;
        const int f = k / kk, t = k - f * kk;
        return KCol{(long long)f * g.C * kk + t, f < g.F};
    }
    __device__ __forceinline__ const float *ptr(int r, const KCol &c) const {
        return (c.ok && r < g.C) ? w + c.off + (long long)r * g.kh * g.kw : nullptr;
    }
};

// W[F][C] read as the K-major operand B(n = c, k = f) of the pointwise dgrad when C is not TMA-friendly
struct MatrixTransposedKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    const float *w;
    int R, K;  // R = C (rows of the operand), K = F; element (r, k) = w[k*R + r]
    struct KCol { long long off; bool ok; };
    __device__ __forceinline__ KCol kcol(int, int k) const { return KCol{(long long)k * R, k < K}; }
    __device__ __forceinline__ const float *ptr(int r, const KCol &c) const { return (c.ok && r < R) ? w + c.off + r : nullptr; }
};

// Rows of a dense [B][R][P] tensor (P % 4 == 0, 16-byte aligned) as a K-major operand, copied by the loader warps with
// 16-byte cp.async -- not because TMA could not describe it, but to take this operand OFF the TMA unit: with 128-byte box
// rows the TMA stream of a CTA tops out near 25 GB/s per SM, and a wgrad whose two operands both stream from HBM is bound
// by it.  One operand through TMA, the other through the LSU path lets the two engines share the load.
struct RowsVecKM {
    static constexpr bool kGather = true;
    static constexpr bool kMN = false;
    static constexpr bool kVec16 = true;
    const float *src;
    int R, P;
    struct KCol { long long off; bool ok; };
    __device__ __forceinline__ KCol kcol(int b, int k) const { return KCol{(long long)b * R * P + k, k < P}; }
    __device__ __forceinline__ const float *ptr(int r, const KCol &c) const { return (c.ok && r < R) ? src + c.off + (long long)r * P : nullptr; }
};
template <class G, class = void> struct is_vec16 { static constexpr bool value = false; };
template <class G> struct is_vec16<G, decltype((void)G::kVec16)> { static constexpr bool value = G::kVec16; };

constexpr int TC_GATHER_THREADS = TC_THREADS + 128;

template <class AG, class BG>
// (min 2 CTAs per SM for the all-TMA variant: tc_launch runs forward / dgrad two per SM, which needs <= 168 registers -- a
// resident-weights experiment pushed it to 172 and silently cost those GEMMs 25 -> 17.6 us's worth of overlap)
__global__ void __launch_bounds__((AG::kGather || BG::kGather) ? TC_GATHER_THREADS : TC_THREADS, (AG::kGather || BG::kGather) ? 1 : 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmX, const TcParams p, const AG ag, const BG bg) {
    constexpr bool kAnyGather = AG::kGather || BG::kGather;
    constexpr bool kAnyTma = !AG::kGather || !BG::kGather;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: stages first (1024-byte aligned), then barriers
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.bn * TC_BK * 4;
    const uint32_t kpi = (uint32_t)p.kpi;  // stage layout: kpi A sub-tiles, then kpi B sub-tiles
    const uint32_t stage_bytes = kpi * (TC_A_BYTES + b_bytes);
    const uint32_t xs_base = smem_base + (uint32_t)p.stages * stage_bytes;  // epilogue-operand ring (p.xring slots)
    const uint32_t bar_base = xs_base + (uint32_t)p.xring * TC_X_SLOT;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 4);
    auto xfull_bar = [&](int i) { return bar_base + 8u * (2 * TC_MAX_STAGES + 6 + i); };
    auto xempty_bar = [&](int i) { return bar_base + 8u * (2 * TC_MAX_STAGES + 6 + TC_XR_MAX + i); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        if (!AG::kGather) tma_prefetch_desc(&tmA);
        if (!BG::kGather) tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), (kAnyTma ? 1u : 0u) + (kAnyGather ? 4u : 0u));
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        for (int i = 0; i < p.xring; ++i) {
            mbar_init(xfull_bar(i), 1);
            mbar_init(xempty_bar(i), 4);
        }
        if (p.xring) tma_prefetch_desc(&tmX);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // number of K iterations of a tile (same in every role)
    auto tile_iters = [&](int tile) -> int {
        if (p.mode == 0) return p.k_blocks;
        const int split = tile / (p.m_blocks * p.n_blocks);
        const int beg = split * p.items_per_split;
        int end = beg + p.items_per_split;
        if (end > p.total_items) end = p.total_items;
        return end - beg;
    };

    if (warp == 0) {
        // ================================ TMA producer ================================
        // (whole warp, uniform control flow; one elected lane issues -- see the MMA issuer below)
        if (kAnyTma) {
            int s = 0, xs = 0;
            uint32_t ph = 0, xph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int b = 0, m0, n0, item0 = 0;
                if (p.mode == 0) {
                    const int per_b = p.m_blocks * p.n_blocks;
                    b = tile / per_b;
                    const int r = tile - b * per_b;
                    m0 = (r / p.n_blocks) * TC_BM;
                    n0 = (r % p.n_blocks) * p.bn;
                } else {
                    const int mn = p.m_blocks * p.n_blocks;
                    const int split = tile / mn;
                    const int r = tile - split * mn;
                    m0 = (r / p.n_blocks) * TC_BM;
                    n0 = (r % p.n_blocks) * p.bn;
                    item0 = split * p.items_per_split;
                }
                const int iters = tile_iters(tile);
                for (int it = 0; it <= iters; ++it) {
                    if (it == iters) {
                        // the tile's operands are on their way: now the second operand of its epilogue, one slot per
                        // 32-channel chunk (pixels / channels past the tensor are zero-filled)
                        for (int c = 0; c < p.bn && p.xring > 0; c += 32) {
                            mbar_wait(xempty_bar(xs), xph ^ 1u);
                            const uint32_t xb = xfull_bar(xs), dst = xs_base + (uint32_t)xs * TC_X_SLOT;
                            if (elect_one()) {
                                mbar_expect_tx(xb, TC_X_SLOT);
#pragma unroll
                                for (int j = 0; j < TC_BM / 32; ++j) tma_load_3d(dst + j * 4096u, &tmX, xb, m0 + 32 * j, n0 + c, b);
                            }
                            if (++xs == p.xring) { xs = 0; xph ^= 1u; }
                        }
                        break;
                    }
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sA = smem_base + (uint32_t)s * stage_bytes, sB = sA + kpi * TC_A_BYTES;
                    const uint32_t fb = full_bar(s);
                    const bool issuer = elect_one();
                    if (issuer) mbar_expect_tx(fb, kpi * ((AG::kGather ? 0u : p.a_tx) + (BG::kGather ? 0u : b_bytes)));
                    int kb = it, bb = b;
                    if (p.mode == 1) {
                        const int kk = item0 + it;
                        bb = kk / p.k_blocks;
                        kb = kk - bb * p.k_blocks;
                    }
                    const int k0 = kb * TC_BK * (int)kpi;
                    if (kpi > 1) {
                        // wide items (wgrad, K-major operands): kpi consecutive 128-byte k-blocks of every row arrive
                        // together -- 512 B of each (n, f) plane row per item instead of 128 B, which is what the DRAM pages
                        // want; k-blocks past the plane are zero-filled by the TMA unit
                        if (issuer) {
                            for (uint32_t j = 0; j < kpi; ++j) {
                                tma_load_3d(sA + j * TC_A_BYTES, &tmA, fb, k0 + (int)j * TC_BK, m0, bb);
                                tma_load_3d(sB + j * b_bytes, &tmB, fb, k0 + (int)j * TC_BK, n0, bb);
                            }
                        }
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                        continue;
                    }
                    if (!AG::kGather && issuer) {
                        if (p.a_mn) {
#pragma unroll
                            for (int j = 0; j < TC_BM / 32; ++j) tma_load_3d(sA + j * 4096u, &tmA, fb, m0 + 32 * j, k0, bb);
                        } else {
                            tma_load_3d(sA, &tmA, fb, k0, m0, bb);
                        }
                    }
                    if (!BG::kGather && issuer) {
                        const int bbB = (p.mode == 1 || p.b_batched) ? bb : 0;
                        if (p.b_mn) {
                            for (int j = 0; j < p.bn / 32; ++j) tma_load_3d(sB + j * 4096u, &tmB, fb, n0 + 32 * j, k0, bbB);
                        } else {
                            tma_load_3d(sB, &tmB, fb, k0, n0, bbB);
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // The whole warp walks the loop (uniform control flow: barrier addresses and descriptors stay in uniform registers) and
        // one elected lane issues; the descriptors' upper words are kernel constants, the lower words advance by a fixed step
        // per K = 8 slice.  (Inside an `if (lane == 0)` region every tcgen05.mma was wrapped in an ELECT / R2UR.BROADCAST /
        // BRA.U.ANY waterfall after ~20 scalar instructions of descriptor building: ~100-200 cycles of issue latency per MMA.)
        {
            const uint32_t idesc = idesc_tf32(TC_BM, p.bn, p.a_mn, p.b_mn);
            const uint32_t a_hi = p.a_mn ? smem_desc_hi(p.mn_sbo, p.mn_layout) : smem_desc_hi(1024u, LAYOUT_SW128);
            const uint32_t b_hi = p.b_mn ? smem_desc_hi(p.mn_sbo, p.mn_layout) : smem_desc_hi(1024u, LAYOUT_SW128);
            const uint32_t a_lbo = p.a_mn ? p.mn_lbo : 16u, b_lbo = p.b_mn ? p.mn_lbo : 16u;
            const uint32_t a_step = (p.a_mn ? p.mn_kstep : 32u) >> 4, b_step = (p.b_mn ? p.mn_kstep : 32u) >> 4;
            int s = 0;
            uint32_t ph = 0;
            int local = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
                const int acc = local & 1;
                const uint32_t aph = (uint32_t)(local >> 1) & 1u;
                const int iters = tile_iters(tile);
                mbar_wait(tempty_bar(acc), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * p.acc_stride;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sA0 = smem_base + (uint32_t)s * stage_bytes, sB0 = sA0 + kpi * TC_A_BYTES;
                    int nks = TC_BK / 8;
                    if (p.mode == 0) {
                        const int rem = p.K - it * TC_BK;
                        if (rem < TC_BK) nks = (rem + 7) / 8;
                    }
#pragma unroll 1
                    for (uint32_t j = 0; j < kpi; ++j) {
                        const uint32_t a_lo = smem_desc_lo(sA0 + j * TC_A_BYTES, a_lbo), b_lo = smem_desc_lo(sB0 + j * b_bytes, b_lbo);
                        const uint32_t acc0 = (it > 0 || j > 0) ? 1u : 0u;
                        if (nks == 4) {
                            if (elect_one()) mma_tf32_k4(d_tmem, a_lo, a_hi, b_lo, b_hi, a_step, b_step, idesc, acc0);
                        } else {
#pragma unroll 1
                            for (int ks = 0; ks < nks; ++ks)
                                if (elect_one())
                                    mma_tf32_lohi(d_tmem, a_lo + (uint32_t)ks * a_step, a_hi, b_lo + (uint32_t)ks * b_step, b_hi, idesc,
                                                  (acc0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    if (elect_one()) mma_commit(empty_bar(s));  // frees the stage once these MMAs have read it
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) mma_commit(tfull_bar(acc));  // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ================================ gather loaders (warps 6..9) ===================
        if constexpr (kAnyGather) {
            // The copies are asynchronous (cp.async): a warp issues the whole stage, commits the group and moves on
            // to the next stage; a stage is published (fence.proxy.async + arrive on its full barrier) once its
            // group has landed, `depth` stages later -- so up to depth+1 stages of gathers are in flight per warp.
            const int lw = warp - 6;
            const int depth = p.stages > 3 ? 3 : p.stages - 1;
            const float *safe = p.out;  // any mapped address: never dereferenced (src-size 0)
            int s = 0, sa = 0, inflight = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int b = 0, m0, n0, item0 = 0;
                if (p.mode == 0) {
                    const int per_b = p.m_blocks * p.n_blocks;
                    b = tile / per_b;
                    const int r = tile - b * per_b;
                    m0 = (r / p.n_blocks) * TC_BM;
                    n0 = (r % p.n_blocks) * p.bn;
                } else {
                    const int mn = p.m_blocks * p.n_blocks;
                    const int split = tile / mn;
                    const int r = tile - split * mn;
                    m0 = (r / p.n_blocks) * TC_BM;
                    n0 = (r % p.n_blocks) * p.bn;
                    item0 = split * p.items_per_split;
                }
                const int iters = tile_iters(tile);
                for (int it = 0; it < iters; ++it) {
                    int kb = it, bb = b;
                    if (p.mode == 1) {
                        const int kk = item0 + it;
                        bb = kk / p.k_blocks;
                        kb = kk - bb * p.k_blocks;
                    }
                    const int k0 = kb * TC_BK;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sA = smem_base + (uint32_t)s * stage_bytes, sB = sA + TC_A_BYTES;  // (kpi == 1 with loaders)
                    // Index arithmetic and copies are issued in explicit batches (pointer arrays + fully unrolled
                    // inner loops): the compiler's own unrolling of these loops is not stable across builds, and a
                    // rolled loop serialises address computation behind every single cp.async.
                    if constexpr (AG::kGather) {
                        if constexpr (AG::kMN) {
                            // lanes along m (coalesced pixels), this warp owns k = lw*8 .. lw*8+7
                            typename AG::Row rows[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) rows[j] = ag.row(bb, m0 + 32 * j + lane);
#pragma unroll
                            for (int kk0 = 0; kk0 < 8; kk0 += 2) {
                                const float *pp[2][4];
#pragma unroll
                                for (int u = 0; u < 2; ++u)
#pragma unroll
                                    for (int j = 0; j < 4; ++j) pp[u][j] = ag.ptr(rows[j], k0 + lw * 8 + kk0 + u);
#pragma unroll
                                for (int u = 0; u < 2; ++u)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        cp_async_f32(sA + mn_tile_off(32 * j + lane, lw * 8 + kk0 + u), pp[u][j], safe);
                            }
                        } else {
                            // lanes along k, this warp owns rows lw, lw+4, ...
                            const auto kc = ag.kcol(bb, k0 + lane);
#pragma unroll
                            for (int r0 = 0; r0 < TC_BM; r0 += 32) {
                                const float *pp[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) pp[u] = ag.ptr(m0 + r0 + lw + 4 * u, kc);
#pragma unroll
                                for (int u = 0; u < 8; ++u) cp_async_f32(sA + km_tile_off(r0 + lw + 4 * u, lane), pp[u], safe);
                            }
                        }
                    }
                    if constexpr (is_vec16<BG>::value) {
                        // 16-byte chunks: lane & 7 = chunk of the 128-byte row, (lw, lane >> 3) = row within 16
                        const int bbB = (p.mode == 1 || p.b_batched) ? bb : 0;
                        const int j = lane & 7, rb = lw * 4 + (lane >> 3);
                        const auto kc = bg.kcol(bbB, k0 + 4 * j);
                        for (int r = rb; r < p.bn; r += 16)
                            cp_async_16(sB + (uint32_t)r * 128u + ((uint32_t)(j ^ (r & 7)) << 4), bg.ptr(n0 + r, kc), safe);
                    } else if constexpr (BG::kGather) {
                        static_assert(!BG::kMN, "gathered B operands are K-major");
                        const int bbB = (p.mode == 1 || p.b_batched) ? bb : 0;
                        const auto kc = bg.kcol(bbB, k0 + lane);
                        for (int r0 = 0; r0 < p.bn; r0 += 32) {  // bn is a multiple of 32
                            const float *pp[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) pp[u] = bg.ptr(n0 + r0 + lw + 4 * u, kc);
#pragma unroll
                            for (int u = 0; u < 8; ++u) cp_async_f32(sB + km_tile_off(r0 + lw + 4 * u, lane), pp[u], safe);
                        }
                    }
                    cp_async_commit();
                    ++inflight;
                    if (inflight > depth) {
                        cp_async_wait_pending(depth);
                        fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(sa));
                        if (++sa == p.stages) sa = 0;
                        --inflight;
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
            cp_async_wait_pending(0);
            fence_proxy_async();
            __syncwarp();
            while (inflight > 0) {
                if (lane == 0) mbar_arrive(full_bar(sa));
                if (++sa == p.stages) sa = 0;
                --inflight;
            }
        }
    } else {
        // ================================ epilogue (warps 2..5) ========================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int local = 0, xs = 0;
        uint32_t xph = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
            const int acc = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            int b = 0, m0, n0, split = 0;
            if (p.mode == 0) {
                const int per_b = p.m_blocks * p.n_blocks;
                b = tile / per_b;
                const int r = tile - b * per_b;
                m0 = (r / p.n_blocks) * TC_BM;
                n0 = (r % p.n_blocks) * p.bn;
            } else {
                const int mn = p.m_blocks * p.n_blocks;
                split = tile / mn;
                const int r = tile - split * mn;
                m0 = (r / p.n_blocks) * TC_BM;
                n0 = (r % p.n_blocks) * p.bn;
            }
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const int m = m0 + 32 * q + lane;
            const bool m_ok = m < p.M;
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * p.acc_stride;
            for (int c = 0; c < p.bn; c += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + (uint32_t)c, v);
                const int nb = n0 + c;
                if (p.xring > 0) {
                    // folded BatchNorm backward, second operand staged by TMA: slot = [4 pixel groups][32 channels][32 pixels]
                    mbar_wait(xfull_bar(xs), xph);
                    const uint32_t xa = xs_base + (uint32_t)xs * TC_X_SLOT + (uint32_t)q * 4096u + (uint32_t)lane * 4u;
                    tmem_ld_wait();
                    float *o = p.out + ((long long)b * p.N + nb) * p.ldo + m;
                    if (nb + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float r = fmaf(__ldg(p.epi_cb + nb + j), ld_shared_f32(xa + j * 128u), __uint_as_float(v[j])) +
                                            __ldg(p.epi_cd + nb + j);
                            if (m_ok) o[(long long)j * p.ldo] = r;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (nb + j < p.N) {
                                const float r = fmaf(__ldg(p.epi_cb + nb + j), ld_shared_f32(xa + j * 128u), __uint_as_float(v[j])) +
                                                __ldg(p.epi_cd + nb + j);
                                if (m_ok) o[(long long)j * p.ldo] = r;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(xempty_bar(xs));
                    if (++xs == p.xring) { xs = 0; xph ^= 1u; }
                    continue;
                }
                if (p.epi == 0 && p.epi_x != nullptr) {
                    // folded BatchNorm backward: the second operand of the epilogue is read like the output is written
                    // (lane <-> pixel, 128 B per channel and warp); the loads are in flight while the accumulator arrives
                    const long long off = ((long long)b * p.N + nb) * p.ldo + m;
                    float xv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) xv[j] = (m_ok && nb + j < p.N) ? __ldg(p.epi_x + off + (long long)j * p.ldo) : 0.0f;
                    tmem_ld_wait();
                    if (m_ok) {
                        float *o = p.out + off;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.N)
                                o[(long long)j * p.ldo] = fmaf(__ldg(p.epi_cb + nb + j), xv[j], __uint_as_float(v[j])) + __ldg(p.epi_cd + nb + j);
                    }
                    continue;
                }
                tmem_ld_wait();
                if (p.epi == 0) {
                    // lane <-> pixel: for every output channel the warp stores 32 consecutive floats (one 128 B line)
                    float *o = p.out + ((long long)b * p.N + nb) * p.ldo + m;
                    if (m_ok) {
                        if (nb + 32 <= p.N && p.bias == nullptr) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[(long long)j * p.ldo] = __uint_as_float(v[j]);
                        } else if (nb + 32 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[(long long)j * p.ldo] = __uint_as_float(v[j]) + __ldg(p.bias + nb + j);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (nb + j < p.N) {
                                    float r = __uint_as_float(v[j]);
                                    if (p.bias) r += __ldg(p.bias + nb + j);
                                    o[(long long)j * p.ldo] = r;
                                }
                            }
                        }
                    }
                } else if (p.epi == 2) {
                    // pointwise dgrad with stride s: zero-stuffed dX[b][n][oh*s + di][ow*s + dj] (pointwise_convolution.py:68-72)
                    if (m_ok) {
                        const int oh = m / p.epi_ow, ow = m - oh * p.epi_ow;
                        const int st2 = p.epi_s;
                        const long long plane = (long long)p.epi_ow * st2 * ((long long)(p.M / p.epi_ow) * st2);
                        float *o = p.out + ((long long)b * p.N + nb) * plane + ((long long)oh * st2) * (p.epi_ow * st2) + ow * st2;
                        if (st2 == 2 && nb + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.out) & 7u) == 0) {
                            // the common case: each lane owns a 2x2 cell -> two 8-byte stores, 256 B contiguous per warp and row
                            const int pitch = p.epi_ow * 2;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float *oj = o + (long long)j * plane;
                                *reinterpret_cast<float2 *>(oj) = make_float2(__uint_as_float(v[j]), 0.0f);
                                *reinterpret_cast<float2 *>(oj + pitch) = make_float2(0.0f, 0.0f);
                            }
                            continue;
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (nb + j < p.N) {
                                float *oj = o + (long long)j * plane;
                                for (int di = 0; di < st2; ++di)
                                    for (int dj = 0; dj < st2; ++dj)
                                        oj[(long long)di * (p.epi_ow * st2) + dj] = (di | dj) ? 0.0f : __uint_as_float(v[j]);
                            }
                        }
                    }
                } else {
                    if (m_ok) {
                        const long long row = ((long long)split * p.M + m) * p.N + nb;
                        float *o = p.out + row;
                        if (p.bias || p.epi_w) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (nb + j < p.N) {
                                    float r = __uint_as_float(v[j]);
                                    if (p.bias) r += __ldg(p.bias + nb + j);
                                    if (p.epi_w) r = fmaf(p.epi_l2, __ldg(p.epi_w + row + j), r);
                                    o[j] = r;
                                }
                            }
                        } else if (nb + 32 <= p.N && (p.N & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(o + j) =
                                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.N) o[j] = __uint_as_float(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (p.sk_counter != nullptr) {
                // ---- cooperative split-K reduction (one tile per CTA, all CTAs resident: host guarantees) ----
                const int et = (int)threadIdx.x - 64;  // 0..127 over the four epilogue warps
                const int mn = p.m_blocks * p.n_blocks;
                unsigned int *cnt = p.sk_counter + 2 * (tile - split * mn);
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) {
                    atomicAdd(cnt, 1u);
                    const long long t0 = clock64();
                    unsigned int seen;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
                        if (clock64() - t0 > 8000000000LL) {
                            printf("dorknet_b200: split-K arrival wait timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    } while (seen < (unsigned int)p.splits);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int mrows = p.M - m0 < TC_BM ? p.M - m0 : TC_BM;
                const int cols = p.N - n0 < p.bn ? p.N - n0 : p.bn;
                const long long plane = (long long)p.M * p.N;
                if ((p.N & 3) == 0 && (cols & 3) == 0) {
                    // 16-byte groups; tz threads share a group and split the partials between them (fixed shuffle tree:
                    // deterministic), eight loads in flight per thread
                    const int groups = (mrows * cols) >> 2;
                    const int per = (groups + p.splits - 1) / p.splits;
                    int tz = 1;
                    while (tz < 32 && per * tz * 2 <= 128) tz *= 2;
                    const int zl = et & (tz - 1);
                    const int g_hi = (split + 1) * per < groups ? (split + 1) * per : groups;
                    for (int g0 = split * per; g0 < g_hi; g0 += 128 / tz) {
                        const int g = g0 + et / tz;
                        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                        long long off = 0;
                        if (g < g_hi) {
                            const int e = g << 2;
                            const int rm = e / cols, cn = e - rm * cols;
                            off = (long long)(m0 + rm) * p.N + n0 + cn;
                            int z = zl;
                            for (; z + 7 * tz < p.splits; z += 8 * tz) {
                                float4 v[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4 *>(p.out + (z + u * tz) * plane + off));
#pragma unroll
                                for (int u = 0; u < 8; ++u) { sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w; }
                            }
                            for (; z < p.splits; z += tz) {
                                const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.out + z * plane + off));
                                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
                            }
                        }
                        for (int o = tz >> 1; o > 0; o >>= 1) {
                            sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
                            sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
                            sum.z += __shfl_xor_sync(0xffffffffu, sum.z, o);
                            sum.w += __shfl_xor_sync(0xffffffffu, sum.w, o);
                        }
                        if (g < g_hi && zl == 0) {
                            if (p.sk_l2 != 0.0f) {
                                const float4 wv = __ldg(reinterpret_cast<const float4 *>(p.sk_w + off));
                                sum.x = fmaf(p.sk_l2, wv.x, sum.x); sum.y = fmaf(p.sk_l2, wv.y, sum.y);
                                sum.z = fmaf(p.sk_l2, wv.z, sum.z); sum.w = fmaf(p.sk_l2, wv.w, sum.w);
                            }
                            *reinterpret_cast<float4 *>(p.sk_out + off) = sum;
                        }
                    }
                } else {
                    const int total = mrows * cols;
                    const int per = (total + p.splits - 1) / p.splits;
                    int tz = 1;
                    while (tz < 32 && per * tz * 2 <= 128) tz *= 2;
                    const int zl = et & (tz - 1);
                    const int e_hi = (split + 1) * per < total ? (split + 1) * per : total;
                    for (int e0 = split * per; e0 < e_hi; e0 += 128 / tz) {
                        const int e = e0 + et / tz;
                        float sum = 0.0f;
                        long long off = 0;
                        if (e < e_hi) {
                            const int rm = e / cols, cn = e - rm * cols;
                            off = (long long)(m0 + rm) * p.N + n0 + cn;
                            for (int z = zl; z < p.splits; z += tz) sum += __ldcg(p.out + z * plane + off);
                        }
                        for (int o = tz >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                        if (e < e_hi && zl == 0) p.sk_out[off] = sum + (p.sk_l2 != 0.0f ? p.sk_l2 * __ldg(p.sk_w + off) : 0.0f);
                    }
                }
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (et == 0) {
                    const unsigned int prev = atomicAdd(cnt + 1, 1u);
                    if (prev == (unsigned int)p.splits - 1u) {  // everybody has read the partials: re-arm for the next launch
                        cnt[0] = 0u;
                        cnt[1] = 0u;
                        __threadfence();
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static bool g_tc_ready = false;
static int g_mn_layout = LAYOUT_SW128_BASE32B, g_mn_lbo = 4096, g_mn_sbo = 512, g_mn_kstep = 1024;
static int g_mn_swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
static int g_l2_promo = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
static int g_smem_budget = TC_SMEM_BUDGET;  // bytes of operand stages per CTA
static int g_hybrid_wgrad = 0;               // 1: pointwise wgrad takes X through 16-byte cp.async loaders, dY through TMA
static int g_wide_items = 0;                 // wgrad k-blocks per item: 0 = automatic (pw_wgrad), 1 / 2 / 4 = forced
static int g_two_per_sm = 1;                 // see tc_launch
static int g_ctas_per_sm = 1;                // persistent CTAs per SM the grids / split plans are sized for
static unsigned int *g_sk_counters = nullptr;  // arrival / departure counters of the in-kernel split-K reduction (zero at rest)
constexpr int TC_SK_MAX_TILES = 512;
// 1: the CTAs of a wgrad tile add their split-K partials themselves.  Parity-green but measured SLOWER than the separate
// reduce kernel on every pointwise shape (profiles/r02w_pw_wgrad_fused_reduce.log: +0 .. +5 us of 16 .. 29 us; 128 reducing
// threads per SM against a reduce kernel's thousands, behind the same all-CTAs-arrived barrier): off, kept behind knob 23
static int g_fused_reduce = 0;
static int g_wgrad_bn_cap = 0;               // pointwise wgrad: largest MMA N (0 = automatic, see pw_wgrad)
static int g_epi_ring = 1;                   // 0: the affine dgrad epilogue loads its second operand itself (no TMA ring)
static int g_tc_disable_mask = 0;  // bit0 fwd, bit1 dgrad, bit2 wgrad (bring-up / tests)

int init_gemm_tcgen05() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        g_tc_ready = false;  // dispatcher falls back to the SIMT kernels
        return DK_OK;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (g_sk_counters == nullptr) {
        if (cudaMalloc(&g_sk_counters, 2 * TC_SK_MAX_TILES * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(g_sk_counters, 0, 2 * TC_SK_MAX_TILES * sizeof(unsigned int)) != cudaSuccess) {
            cudaGetLastError();
            g_sk_counters = nullptr;  // (the separate reduce kernel is used instead)
        }
    }
    if (const char *m = getenv("DK_TC_DISABLE_MASK")) g_tc_disable_mask = atoi(m);  // diagnostics
    if (const char *m = getenv("DK_HYBRID_WGRAD")) g_hybrid_wgrad = atoi(m);
    g_tc_ready = true;
    return DK_OK;
}

struct MapKey {
    const void *ptr;
    uint64_t d0, d1, d2;
    uint32_t b0, b1, sw;
    bool operator==(const MapKey &o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && b0 == o.b0 && b1 == o.b1 && sw == o.sw;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        size_t h = reinterpret_cast<size_t>(k.ptr);
        auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
        mix(k.d0); mix(k.d1); mix(k.d2); mix(k.b0); mix(k.b1); mix(k.sw);
        return h;
    }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

// 3-D fp32 tensor map: dims (d0 fastest, d1, d2) with dense strides, box (b0, b1, 1).
static int make_map(CUtensorMap *out, const float *ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                    CUtensorMapSwizzle sw) {
    MapKey key{ptr, d0, d1, d2, b0, b1, (uint32_t)sw};
    {
        std::lock_guard<std::mutex> g(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) {
            *out = it->second;
            return DK_OK;
        }
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, (CUtensorMapL2promotion)g_l2_promo,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for dims (%llu,%llu,%llu) box (%u,%u)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
        return DK_ERR_CUDA;
    }
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps[key] = *out;
    return DK_OK;
}

static bool tma_ok(const void *p, int64_t inner) { return aligned16(p) && (inner % 4) == 0; }

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

static void fill_common(TcParams &p) {
    p.bn = p.N >= 256 ? 256 : round_up(p.N, 32);
    p.m_blocks = (int)ceil_div(p.M, TC_BM);
    p.n_blocks = (int)ceil_div(p.N, p.bn);
    p.k_blocks = (int)ceil_div(p.K, TC_BK);
    p.acc_stride = p.bn <= 32 ? 32 : p.bn <= 64 ? 64 : p.bn <= 128 ? 128 : 256;
    p.tmem_cols = 2 * p.acc_stride;
    const int stage_bytes = TC_A_BYTES + p.bn * TC_BK * 4;
    int st = g_smem_budget / stage_bytes;
    if (st > TC_MAX_STAGES) st = TC_MAX_STAGES;
    p.stages = st;
    p.kpi = 1;
    p.a_tx = TC_A_BYTES;
    p.mn_layout = (uint32_t)g_mn_layout;
    p.mn_lbo = (uint32_t)g_mn_lbo;
    p.mn_sbo = (uint32_t)g_mn_sbo;
    p.mn_kstep = (uint32_t)g_mn_kstep;
}

// two_per_sm: forward / dgrad GEMMs with at least two tiles per SM run TWO persistent CTAs per SM on half the stage budget
// each (the epilogue of one overlaps the loads of the other, and 448 tiles no longer cost 4 rounds of 148): measured at
// batch 64, 28x28x128 fwd 15.2 -> 12.4 us, dgrad 15.6 -> 12.9 us, 56x56x64 19.9 -> 17.6 us; wgrad and the gather variants
// lose (more split-K partials / loader warps competing), so they keep one CTA per SM
template <class AG, class BG>
static int tc_launch(const CUtensorMap &ta, const CUtensorMap &tb, TcParams p, const AG &ag, const BG &bg,
                     cudaStream_t st, bool two_per_sm = false, const CUtensorMap *tx = nullptr) {
    static bool attr_set = false;
    if (!attr_set) {
        DK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<AG, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TC_SMEM_BUDGET + 4096));
        attr_set = true;
    }
    const int stage_bytes = p.kpi * (TC_A_BYTES + p.bn * TC_BK * 4);
    int per_sm = g_ctas_per_sm;
    if (tx == nullptr) p.xring = 0;
    const int xbytes = p.xring * (int)TC_X_SLOT;
    if (two_per_sm && g_two_per_sm && per_sm == 1 && p.num_tiles >= 2 * sm_count() && p.tmem_cols <= 256) {
        // (with an epilogue-operand ring the two CTAs share the whole 227 KB: 110 KB each)
        const int st2 = ((xbytes ? 110 : 98) * 1024 - xbytes) / stage_bytes;
        if (st2 >= 3) {
            per_sm = 2;
            if (p.stages > st2) p.stages = st2;
        }
    }
    if (per_sm == 1 && xbytes) {
        const int st1 = (g_smem_budget - xbytes) / stage_bytes;
        if (st1 < 2) return DK_ERR_UNSUPPORTED;
        if (p.stages > st1) p.stages = st1;
    }
    const size_t smem = (size_t)p.stages * stage_bytes + xbytes + 1024 /*align slack*/ + 8 * (2 * TC_MAX_STAGES + 8 + 2 * TC_XR_MAX);
    const int cap = sm_count() * per_sm;
    const int grid = p.num_tiles < cap ? p.num_tiles : cap;
    const int threads = (AG::kGather || BG::kGather) ? TC_GATHER_THREADS : TC_THREADS;
    tc_gemm_kernel<AG, BG><<<grid, threads, smem, st>>>(ta, tb, tx ? *tx : ta, p, ag, bg);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// split plan for wgrad-type GEMMs (output [M][N], reduction over `total_items` k-blocks): enough CTAs to fill the
// machine, but the partial sums must stay small next to the inputs
static void split_plan(int M, int N, int64_t total_items, int64_t in_bytes, int *splits, int *per, int bn_used = 0) {
    const int bn = bn_used > 0 ? bn_used : N >= 256 ? 256 : round_up(N, 32);
    const int tiles = (int)(ceil_div(M, TC_BM) * ceil_div(N, bn));
    int64_t s = (int64_t)sm_count() * g_ctas_per_sm / tiles;
    const int64_t out_bytes = (int64_t)M * N * 4;
    // partial sums are written once and read once: keep them under ~1/4 of the input bytes, but never refuse the
    // first 32 MB of them (tiny layers would otherwise run on a handful of SMs)
    int64_t cap = in_bytes / (4 * out_bytes);
    const int64_t cap_abs = (32ll << 20) / out_bytes;
    if (cap < cap_abs) cap = cap_abs;
    if (s > cap) s = cap;
    if (s > total_items) s = total_items;
    if (s < 1) s = 1;
    *per = (int)ceil_div(total_items, s);
    *splits = (int)ceil_div(total_items, *per);
}

static int g_short_a = 0;  // 1: wgrad TMA box of the dY operand covers only its F < 128 real rows (measured: no gain, 31.0 us either way at 64x64x56x56)
static int g_repack_mask = 3;  // bit0: misaligned stride-1 planes, bit1: stride-s planes of <= TC_REPACK_MAX_P pixels
constexpr int64_t TC_REPACK_MAX_P = 1024;
static int64_t repack_pitch(int64_t P) { return (P + 3) / 4 * 4; }
static size_t repack_bytes(int64_t rows, int64_t P) { return (size_t)(rows * repack_pitch(P)) * sizeof(float) + 256; }

static ConvGeom mk_geom(int C, int H, int W, int F, int kh, int kw, int s, int p) {
    ConvGeom g{C, H, W, F, kh, kw, s, p, (H + 2 * p - kh) / s + 1, (W + 2 * p - kw) / s + 1};
    return g;
}

size_t tc_conv_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    const ConvGeom g = mk_geom(C, H, W, F, kh, kw, s, p);
    const int64_t P = (int64_t)g.OH * g.OW;
    const int Kf = C * kh * kw;
    int splits = 1, per;
    for (int kpi = 1; kpi <= 4; kpi *= 2) {  // (pw_wgrad may group 2 or 4 k-blocks per item: take the largest plan)
        int sp;
        split_plan(F, Kf, (int64_t)N * ceil_div(ceil_div(P, TC_BK), kpi), (int64_t)N * P * (F + Kf) * 4, &sp, &per);
        if (sp > splits) splits = sp;
    }
    size_t repack = 0;  // padded copies of dY and X (pointwise only; fwd / dgrad need one of the two)
    if (kh == 1 && kw == 1 && p == 0 && ((P % 4) != 0 || s > 1) && P <= 4 * TC_REPACK_MAX_P)
        repack = repack_bytes((int64_t)N * F, P) + repack_bytes((int64_t)N * C, P) + 512;
    return (size_t)splits * F * Kf * sizeof(float) + repack;
}

static bool dims_ok(int64_t a, int64_t b) { return a > 0 && b > 0 && a < (1 << 30) && b < (1 << 30); }

// ---- plane repacking ------------------------------------------------------------------------------------------------
// TMA needs 16-byte row pitches.  Planes whose pixel count is not a multiple of 4 (7x7 = 49 floats = 196 B) and
// stride-s subsampled planes used to go through the loader warps' 4-byte gathers, which are bound by the issue rate of
// four warps (ncu: 33 instructions per cp.async, stages published one item at a time).  Small planes are instead
// copied once into the workspace with the pitch rounded up to 4 floats (pad = 0, so padded k contribute nothing to a
// reduction and padded m are masked by the epilogue) and every operand goes through TMA.

__global__ void __launch_bounds__(256)
repack_planes_kernel(const float *__restrict__ src, float *__restrict__ dst, long long rows, int plane, int SW, int OW, int s,
                     int P, int Pp) {
    const int q4 = Pp >> 2;
    const long long total = rows * q4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / q4;
        const int j0 = (int)(i - row * q4) * 4;
        const float *sp = src + row * plane;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            float t = 0.0f;
            if (j < P) {
                if (s == 1) t = __ldg(sp + j);
                else {
                    const int oh = j / OW, ow = j - oh * OW;
                    t = __ldg(sp + (long long)oh * s * SW + ow * s);
                }
            }
            v[e] = t;
        }
        *reinterpret_cast<float4 *>(dst + row * Pp + j0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// measured on B200 (tests/pw_sweep.py, batch 64): 7x7 planes fwd/dgrad 26 -> 18 us, wgrad 71 -> 33 us; stride-2 wgrad
// 28x28 -> 14x14 26 -> 20 us, 14x14 -> 7x7 40 -> 26 us; a stride-2 FORWARD whose output planes are TMA-friendly is
// faster through the gather loaders (the repack would re-read what the gather reads once), so it keeps them
static bool repack_wanted(const float *ptr, int64_t P, int s, bool reduction_operand) {
    if (s == 1) return (g_repack_mask & 1) && !tma_ok(ptr, P) && P <= 4 * TC_REPACK_MAX_P;
    if (!(g_repack_mask & 2) || P > TC_REPACK_MAX_P) return false;
    return reduction_operand || (P % 4) != 0;
}
// carve `bytes` (256-byte aligned) off the front of the workspace; nullptr if it does not fit
static float *ws_carve(void *&ws, size_t &ws_bytes, size_t bytes) {
    if (ws == nullptr) return nullptr;
    uintptr_t a = (reinterpret_cast<uintptr_t>(ws) + 255u) & ~(uintptr_t)255u;
    const size_t skip = a - reinterpret_cast<uintptr_t>(ws);
    if (ws_bytes < skip + bytes) return nullptr;
    ws = reinterpret_cast<void *>(a + bytes);
    ws_bytes -= skip + bytes;
    return reinterpret_cast<float *>(a);
}
// [rows][SH][SW] sampled at stride s -> dst[rows][Pp]
static int repack_launch(const float *src, float *dst, int64_t rows, int SH, int SW, int OW, int s, int64_t P, cudaStream_t st) {
    const int Pp = (int)repack_pitch(P);
    const long long total = rows * (Pp / 4);
    const int grid = (int)(ceil_div(total, 256) < 8 * sm_count() ? ceil_div(total, 256) : 8 * sm_count());
    repack_planes_kernel<<<grid, 256, 0, st>>>(src, dst, rows, SH * SW, SW, OW, s, (int)P, Pp);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// ---- pointwise (1x1, pad 0, stride s) ------------------------------------------------------------------------
// x_pitch > 0: x is a dense [N][C][x_pitch] tensor (x_pitch % 4 == 0, rows zero-padded past H*W; s must be 1)
static int pw_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int s,
                  void *ws, size_t ws_bytes, cudaStream_t st, int64_t x_pitch = 0) {
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const int64_t P = (int64_t)OH * OW;
    if (!tma_ok(w, C) || !dims_ok(P, (int64_t)H * W)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 0; q.a_mn = 1; q.b_mn = 0; q.b_batched = 0;
    q.M = (int)P; q.N = F; q.K = C; q.batches = N;
    fill_common(q);
    q.num_tiles = N * q.m_blocks * q.n_blocks;
    q.epi = 0; q.out = y; q.bias = bias; q.ldo = (int)P;
    CUtensorMap ta = {}, tb;
    int rc = make_map(&tb, w, C, F, 1, TC_BK, q.bn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (x_pitch > 0) {
        rc = make_map(&ta, x, x_pitch, C, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st, true);
    }
    if (s == 1 && tma_ok(x, P)) {
        rc = make_map(&ta, x, P, C, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st, true);
    }
    if (repack_wanted(x, P, s, false)) {
        if (float *xp = ws_carve(ws, ws_bytes, repack_bytes((int64_t)N * C, P))) {
            rc = repack_launch(x, xp, (int64_t)N * C, H, W, OW, s, P, st);
            if (rc) return rc;
            rc = make_map(&ta, xp, repack_pitch(P), C, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
            if (rc) return rc;
            return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
        }
    }
    return tc_launch(ta, tb, q, PixelGatherMN{x, C, H, W, OW, (int)P, s}, NoGather{}, st);
}

static int pw_dgrad(const float *dy, const float *w, float *dx, int N, int C, int OH, int OW, int F, int s, void *ws,
                    size_t ws_bytes, cudaStream_t st, const float *epi_x = nullptr, const float *epi_cb = nullptr,
                    const float *epi_cd = nullptr, int64_t dy_pitch = 0) {
    const int64_t P = (int64_t)OH * OW;
    if (!tma_ok(w, C) || !dims_ok(P, P * s * s)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 0; q.a_mn = 1; q.b_mn = 1; q.b_batched = 0;
    q.M = (int)P; q.N = C; q.K = F; q.batches = N;
    fill_common(q);
    q.num_tiles = N * q.m_blocks * q.n_blocks;
    q.out = dx; q.bias = nullptr; q.ldo = (int)P;
    q.epi = s == 1 ? 0 : 2; q.epi_ow = OW; q.epi_s = s;
    if (epi_x != nullptr) {
        if (s != 1) return DK_ERR_UNSUPPORTED;
        q.epi_x = epi_x; q.epi_cb = epi_cb; q.epi_cd = epi_cd;
    }
    CUtensorMap ta = {}, tb;
    int rc = make_map(&tb, w, C, F, 1, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);  // B(k=f, n=c) = W[f][c]: n contiguous
    if (rc) return rc;
    if (dy_pitch > 0) {  // dY already re-pitched by the caller (dk_pw_pack): [N][F][dy_pitch]
        rc = make_map(&ta, dy, dy_pitch, F, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
    }
    if (tma_ok(dy, P)) {
        rc = make_map(&ta, dy, P, F, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        if (epi_x != nullptr && g_epi_ring && tma_ok(epi_x, P)) {
            // the epilogue's second operand through TMA: two tiles' worth of 32-channel chunks, at most TC_XR_MAX / what fits
            CUtensorMap tx;
            rc = make_map(&tx, epi_x, P, C, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
            const int chunks = q.bn / 32;
            q.xring = chunks >= 4 ? 4 : 2 * chunks > TC_XR_MAX ? TC_XR_MAX : 2 * chunks;
            if (chunks == 2 && q.num_tiles >= 2 * sm_count()) q.xring = 2;  // (two CTAs per SM: one tile's worth each)
            rc = tc_launch(ta, tb, q, NoGather{}, NoGather{}, st, true, &tx);
            if (rc != DK_ERR_UNSUPPORTED) return rc;
            q.xring = 0;
        }
        return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st, s == 1);
    }
    if (repack_wanted(dy, P, 1, false)) {
        if (float *gp = ws_carve(ws, ws_bytes, repack_bytes((int64_t)N * F, P))) {
            rc = repack_launch(dy, gp, (int64_t)N * F, OH, OW, OW, 1, P, st);
            if (rc) return rc;
            rc = make_map(&ta, gp, repack_pitch(P), F, N, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
            if (rc) return rc;
            return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
        }
    }
    return tc_launch(ta, tb, q, PixelGatherMN{dy, F, OH, OW, OW, (int)P, 1}, NoGather{}, st);
}

static int pw_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                    int s, void *ws, size_t ws_bytes, cudaStream_t st, int64_t x_pitch = 0, int64_t dy_pitch = 0) {
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const int64_t P = (int64_t)OH * OW;
    if (!dims_ok(P, (int64_t)H * W)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 1; q.a_mn = 0; q.b_mn = 0; q.b_batched = 1;
    q.M = F; q.N = C; q.K = (int)P; q.batches = N;
    fill_common(q);
    {
        // Narrower accumulators for wgrads whose output is large next to their inputs (7x7 / 14x14 planes, 256 .. 512
        // channels): every CTA writes a 128 x bn partial tile, so with ~148 CTAs the split-K partials are 148*128*bn*4 bytes
        // whatever the split -- 18.9 MB at bn = 256 against 12.8 MB of operands for 512 -> 512 at 7x7.
        // Measured (tests/pw_sweep.py, batch 64, us at bn 256 / 128 / 64): 256->256 @14x14 28.5 / 19.7 / 24.4,
        // 512->512 @7x7 33.1 / 25.4 / 32.1, 128->256 @14x14 19.3 / 19.2 / 16.0, 256->512 @7x7 24.1 / 23.5 / 21.4; planes of
        // 28x28 and more are indifferent or lose.
        int cap = g_wgrad_bn_cap;
        if (cap == 0) cap = P > 196 ? 256 : C >= 256 ? 128 : C >= 128 ? 64 : 256;
        if (cap != 64 && cap != 128) cap = 256;
        if (q.bn > cap) {
            q.bn = cap;
            q.n_blocks = (int)ceil_div(q.N, q.bn);
            q.acc_stride = q.bn <= 32 ? 32 : q.bn <= 64 ? 64 : q.bn <= 128 ? 128 : 256;
            q.tmem_cols = 2 * q.acc_stride;
            int st = g_smem_budget / (TC_A_BYTES + q.bn * TC_BK * 4);
            q.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
        }
    }
    bool a_tma = dy_pitch > 0 || tma_ok(dy, P), b_tma = x_pitch > 0 || ((s == 1) && tma_ok(x, P));
    const bool hybrid = g_hybrid_wgrad && a_tma && b_tma && x_pitch == 0;
    if (a_tma && b_tma && !hybrid) {
        // wide pipeline items for long planes: 4 (or 2) consecutive k-blocks per item, if at least two stages still fit and
        // the rounding of the plane to whole items wastes little
        const int kb0 = q.k_blocks, one = TC_A_BYTES + q.bn * TC_BK * 4;
        int kpi = 1;
        for (int cand = 4; cand >= 2; cand >>= 1) {
            const int items = (kb0 + cand - 1) / cand;
            if (P >= 512 && g_smem_budget / (cand * one) >= 2 && items * cand * 16 <= kb0 * 17) { kpi = cand; break; }
        }
        if (g_wide_items == 1 || g_wide_items == 2 || g_wide_items == 4) kpi = g_wide_items;
        if (g_smem_budget / (kpi * one) < 1) kpi = 1;
        if (kpi > 1) {
            q.kpi = kpi;
            q.k_blocks = (kb0 + kpi - 1) / kpi;
            int st = g_smem_budget / (kpi * one);
            q.stages = st > TC_MAX_STAGES ? TC_MAX_STAGES : st;
        }
    }
    q.total_items = N * q.k_blocks;
    split_plan(F, C, q.total_items, (int64_t)N * P * (F + C) * 4, &q.splits, &q.items_per_split, q.bn);
    q.num_tiles = q.m_blocks * q.n_blocks * q.splits;
    const size_t need = (size_t)q.splits * F * C * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("pointwise wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return DK_ERR_WORKSPACE;
    }
    q.epi = 1; q.out = reinterpret_cast<float *>(ws); q.bias = nullptr; q.ldo = 0;
    CUtensorMap ta = {}, tb = {};
    int rc = DK_OK;
    int64_t pa = dy_pitch > 0 ? dy_pitch : P, pb = x_pitch > 0 ? x_pitch : P;  // row pitches of the operands the maps describe
    {
        void *rest = reinterpret_cast<char *>(ws) + need;
        size_t rest_bytes = ws_bytes - need;
        if (!a_tma && repack_wanted(dy, P, 1, true)) {
            if (float *gp = ws_carve(rest, rest_bytes, repack_bytes((int64_t)N * F, P))) {
                rc = repack_launch(dy, gp, (int64_t)N * F, OH, OW, OW, 1, P, st);
                if (rc) return rc;
                dy = gp; pa = repack_pitch(P); a_tma = true;
            }
        }
        if (a_tma && !b_tma && repack_wanted(x, P, s, true)) {
            if (float *xp = ws_carve(rest, rest_bytes, repack_bytes((int64_t)N * C, P))) {
                rc = repack_launch(x, xp, (int64_t)N * C, H, W, OW, s, P, st);
                if (rc) return rc;
                x = xp; pb = repack_pitch(P); b_tma = true;
            }
        }
    }
    // F < 128: the box stops at the real rows (the rest of the 128-row MMA operand is stale shared memory whose
    // accumulator rows the epilogue never reads) -- half the TMA row requests per stage at F = 64
    const int a_rows = (g_short_a && F < TC_BM) ? round_up(F, 8) : TC_BM;
    if (a_tma) {
        rc = make_map(&ta, dy, pa, F, N, TC_BK, a_rows, CU_TENSOR_MAP_SWIZZLE_128B);
        q.a_tx = (uint32_t)a_rows * TC_BK * 4;
    }
    if (rc) return rc;
    if (b_tma) rc = make_map(&tb, x, pb, C, N, TC_BK, q.bn, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const PixelGatherKM ga{dy, F, OH, OW, OW, (int)P, 1}, gb{x, C, H, W, OW, (int)P, s};
    // every (tile, split) CTA resident at once (one CTA per SM, grid == num_tiles <= SM count): they add the partial sums
    // themselves (sk_counter), otherwise the reduce kernel follows
    const bool fused = g_fused_reduce && g_sk_counters != nullptr && g_ctas_per_sm == 1 && q.splits > 1 &&
                       q.num_tiles <= sm_count() && q.m_blocks * q.n_blocks <= TC_SK_MAX_TILES;
    if (fused) {
        q.sk_counter = g_sk_counters;
        q.sk_out = dw;
        q.sk_w = w;
        q.sk_l2 = l2;
    }
    if (hybrid) rc = tc_launch(ta, tb, q, NoGather{}, RowsVecKM{x, C, (int)P}, st);
    else if (a_tma && b_tma) rc = tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
    else if (a_tma) rc = tc_launch(ta, tb, q, NoGather{}, gb, st);
    else rc = tc_launch(ta, tb, q, ga, gb, st);
    if (rc) return rc;
    if (fused) return DK_OK;
    splitk_reduce_launch(q.out, w, dw, l2, (int64_t)F * C, q.splits, st);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// ---- general convolution as an implicit GEMM (patches gathered by the loader warps) -------------------------------
static int cv_fwd(const float *x, const float *w, const float *bias, float *y, int N, const ConvGeom &g, cudaStream_t st) {
    const int64_t P = (int64_t)g.OH * g.OW;
    const int Kf = g.C * g.kh * g.kw;
    if (!dims_ok(P, (int64_t)g.H * g.W)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 0; q.a_mn = 1; q.b_mn = 0; q.b_batched = 0;
    q.M = (int)P; q.N = g.F; q.K = Kf; q.batches = N;
    fill_common(q);
    q.num_tiles = N * q.m_blocks * q.n_blocks;
    q.epi = 0; q.out = y; q.bias = bias; q.ldo = (int)P;
    CUtensorMap ta = {}, tb = {};
    const ConvPatchMN ga{x, g};
    if (tma_ok(w, Kf)) {
        int rc = make_map(&tb, w, Kf, g.F, 1, TC_BK, q.bn, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        return tc_launch(ta, tb, q, ga, NoGather{}, st);
    }
    return tc_launch(ta, tb, q, ga, MatrixKM{w, g.F, Kf}, st);
}

static int cv_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, const ConvGeom &g, void *ws,
                    size_t ws_bytes, cudaStream_t st) {
    const int64_t P = (int64_t)g.OH * g.OW;
    const int Kf = g.C * g.kh * g.kw;
    if (!dims_ok(P, (int64_t)g.H * g.W)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 1; q.a_mn = 0; q.b_mn = 0; q.b_batched = 1;
    q.M = g.F; q.N = Kf; q.K = (int)P; q.batches = N;
    fill_common(q);
    q.total_items = N * q.k_blocks;
    split_plan(g.F, Kf, q.total_items, (int64_t)N * P * (g.F + Kf) * 4, &q.splits, &q.items_per_split);
    q.num_tiles = q.m_blocks * q.n_blocks * q.splits;
    const size_t need = (size_t)q.splits * g.F * Kf * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("conv wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return DK_ERR_WORKSPACE;
    }
    q.epi = 1; q.out = reinterpret_cast<float *>(ws); q.bias = nullptr; q.ldo = 0;
    CUtensorMap ta = {}, tb = {};
    const ConvPatchKM gb{x, g};
    int rc;
    if (tma_ok(dy, P)) {
        const int a_rows = (g_short_a && g.F < TC_BM) ? round_up(g.F, 8) : TC_BM;
        rc = make_map(&ta, dy, P, g.F, N, TC_BK, a_rows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        q.a_tx = (uint32_t)a_rows * TC_BK * 4;
        rc = tc_launch(ta, tb, q, NoGather{}, gb, st);
    } else {
        rc = tc_launch(ta, tb, q, PixelGatherKM{dy, g.F, g.OH, g.OW, g.OW, (int)P, 1}, gb, st);
    }
    if (rc) return rc;
    splitk_reduce_launch(q.out, w, dw, l2, (int64_t)g.F * Kf, q.splits, st);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

static int cv_dgrad(const float *dy, const float *w, float *dx, int N, const ConvGeom &g, cudaStream_t st) {
    const int64_t HW = (int64_t)g.H * g.W;
    if (!dims_ok(HW, (int64_t)g.OH * g.OW)) return DK_ERR_UNSUPPORTED;
    TcParams q = {};
    q.mode = 0; q.a_mn = 1; q.b_mn = 0; q.b_batched = 0;
    q.M = (int)HW; q.N = g.C; q.K = g.F * g.kh * g.kw; q.batches = N;
    fill_common(q);
    q.num_tiles = N * q.m_blocks * q.n_blocks;
    q.epi = 0; q.out = dx; q.bias = nullptr; q.ldo = (int)HW;
    CUtensorMap ta = {}, tb = {};
    return tc_launch(ta, tb, q, ConvDgradMN{dy, g}, ConvDgradWKM{w, g}, st);
}

// ---- general convolution through materialised patches ------------------------------------------------------------------
// Shapes the aligned-box kernels of conv_tma.cu / conv_rows.cu cannot take (stride > 1 with K > 128, row pitches that are not
// a multiple of 16 bytes: MNIST's 4x4 stride-2 and 14x14 layers) used to run on the gather variants above, which are bound by
// the issue rate of four loader warps (dgrad of 32 -> 64 4x4 s2 at 28x28, batch 64: 1.27 ms for 9.6 MB).  The reference's own
// decomposition (im2col.pyx:16-36 + matmul + row2im, convolution.py:58-126) maps onto the all-TMA pointwise GEMMs instead: the
// patches are written ONCE, transposed -- Pt[n][k = (c,i,j)][p = (oh,ow)], pitch rounded up to 4 floats -- so that a
// convolution over C channels IS the pointwise convolution of a K = C*kh*kw channel tensor:
//   forward  Y[n]  = W[F,K] . Pt[n]                 (pw_fwd)
//   wgrad    dW    = sum_n dY[n] . Pt[n]^T          (pw_wgrad)
//   dgrad    dPt[n] = W^T . dY[n], then dX = gather-form col2im of dPt (every dX element adds its <= ceil(k/s)^2 taps: no atomics)
// The patch tensor lives in the caller's workspace; for these layers it is a few tens of MB and stays in the 126 MB L2.
constexpr size_t TC_MAT_MAX_BYTES = (size_t)768 << 20;
int g_conv_mat_enabled = 1;

// one warp per (n, k) row of Pt: the (c, i, j) split is done once per row, a lane writes 4 consecutive pixels per pass
// (one 16-byte store; consecutive lanes -> consecutive 16-byte chunks), all index arithmetic in 32 bits
template <int S>
__global__ void __launch_bounds__(256)
im2col_t_kernel(const float *__restrict__ x, float *__restrict__ pt, ConvGeom g, int P, int Pp, int rows) {
    const int s = S > 0 ? S : g.s;
    const int kk = g.kh * g.kw, K = g.C * kk, q4 = Pp >> 2;
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
        const int n = row / K, k = row - n * K;
        const int c = k / kk, t = k - c * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const float *xc = x + ((long long)n * g.C + c) * g.H * g.W;
        float *dst = pt + (long long)row * Pp;
        for (int c4 = lane; c4 < q4; c4 += 32) {
            const int p0 = c4 * 4;
            int oh = p0 / g.OW, ow = p0 - oh * g.OW;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ih = oh * s - g.p + i, iw = ow * s - g.p + j;
                v[e] = (p0 + e < P && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) ? __ldg(xc + ih * g.W + iw) : 0.0f;
                if (++ow == g.OW) { ow = 0; ++oh; }
            }
            *reinterpret_cast<float4 *>(dst + p0) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// dX[n][c][h][w] = sum over the taps (i, j) that reach it of dPt[n][(c,i,j)][oh][ow], oh*s + i - p = h, ow*s + j - p = w
// (im2col.pyx:209-234 scatter-adds; this is the same sum written as a gather, so it needs no atomics and is deterministic).
// A thread owns one dX element; consecutive threads = consecutive w, so for a fixed tap the warp reads dPt with stride 1/s.
// KH, KW, S > 0: compile-time filter geometry (the tap loops unroll and the divisions by S become shifts).
template <int KH, int KW, int S>
__global__ void __launch_bounds__(256)
col2im_t_kernel(const float *__restrict__ dpt, float *__restrict__ dx, ConvGeom g, int P, long long planes) {
    const int kh = KH > 0 ? KH : g.kh, kw = KW > 0 ? KW : g.kw, s = S > 0 ? S : g.s;
    const int HW = g.H * g.W;
    const int per_plane = (HW + 255) / 256;
    for (long long blk = blockIdx.x; blk < planes * per_plane; blk += gridDim.x) {
        const long long plane = blk / per_plane;  // (n, c)
        const int e = (int)(blk - plane * per_plane) * 256 + threadIdx.x;
        if (e >= HW) continue;
        const int h = e / g.W, w = e - h * g.W;
        const float *base = dpt + plane * (long long)(kh * kw) * P;
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < kh; ++i) {
            const int th = h + g.p - i;
            const int oh = th / s;
            if (th < 0 || oh * s != th || oh >= g.OH) continue;
#pragma unroll
            for (int j = 0; j < kw; ++j) {
                const int tw = w + g.p - j;
                const int ow = tw / s;
                if (tw < 0 || ow * s != tw || ow >= g.OW) continue;
                acc += __ldg(base + (i * kw + j) * P + oh * g.OW + ow);
            }
        }
        dx[plane * HW + e] = acc;
    }
}

template <int KH, int KW, int S>
static void col2im_launch(const float *dpt, float *dx, const ConvGeom &g, int P, long long planes, cudaStream_t st) {
    const long long blocks = planes * ((g.H * g.W + 255) / 256);
    const int grid = (int)(blocks < (long long)sm_count() * 16 ? blocks : (long long)sm_count() * 16);
    col2im_t_kernel<KH, KW, S><<<grid, 256, 0, st>>>(dpt, dx, g, P, planes);
}

static size_t mat_patch_bytes(int N, const ConvGeom &g) {
    const int64_t P = (int64_t)g.OH * g.OW;
    return (size_t)N * g.C * g.kh * g.kw * (size_t)repack_pitch(P) * sizeof(float) + 512;
}
static bool mat_ok(const float *w, int N, const ConvGeom &g) {
    const int Kf = g.C * g.kh * g.kw;
    if (!g_conv_mat_enabled || g.kh * g.kw <= 1 || g.OH < 1 || g.OW < 1) return false;
    if (!tma_ok(w, Kf) || (int64_t)N * Kf >= ((int64_t)1 << 30)) return false;
    return mat_patch_bytes(N, g) <= TC_MAT_MAX_BYTES;
}
size_t tc_conv_mat_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    const ConvGeom g = mk_geom(C, H, W, F, kh, kw, s, p);
    if (g.OH < 1 || g.OW < 1 || kh * kw <= 1 || !g_conv_mat_enabled) return 0;
    const size_t pb = mat_patch_bytes(N, g);
    if (pb > TC_MAT_MAX_BYTES || (C * kh * kw) % 4 != 0) return 0;
    // patches (or dPt) + what the pointwise GEMM over K = C*kh*kw channels wants (split-K partials, repacked dY)
    return pb + tc_conv_ws_bytes(N, C * kh * kw, g.OH, g.OW, F, 1, 1, 1, 0) + 1024;
}

static int mat_im2col(const float *x, float *pt, int N, const ConvGeom &g, cudaStream_t st) {
    const int P = g.OH * g.OW, Pp = (int)repack_pitch(P);
    const int rows = N * g.C * g.kh * g.kw;  // < 2^30 (mat_ok)
    const int want = (rows + 7) / 8, cap = sm_count() * 16;
    const int grid = want < cap ? want : cap;
    if (g.s == 1) im2col_t_kernel<1><<<grid, 256, 0, st>>>(x, pt, g, P, Pp, rows);
    else if (g.s == 2) im2col_t_kernel<2><<<grid, 256, 0, st>>>(x, pt, g, P, Pp, rows);
    else im2col_t_kernel<0><<<grid, 256, 0, st>>>(x, pt, g, P, Pp, rows);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

static int cvm_fwd(const float *x, const float *w, const float *bias, float *y, int N, const ConvGeom &g, void *ws,
                   size_t ws_bytes, cudaStream_t st) {
    if (!mat_ok(w, N, g)) return DK_ERR_UNSUPPORTED;
    float *pt = ws_carve(ws, ws_bytes, mat_patch_bytes(N, g));
    if (pt == nullptr) return DK_ERR_UNSUPPORTED;
    int rc = mat_im2col(x, pt, N, g, st);
    if (rc) return rc;
    return pw_fwd(pt, w, bias, y, N, g.C * g.kh * g.kw, g.OH, g.OW, g.F, 1, ws, ws_bytes, st, repack_pitch((int64_t)g.OH * g.OW));
}

static int cvm_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, const ConvGeom &g, void *ws,
                     size_t ws_bytes, cudaStream_t st) {
    if (!mat_ok(w, N, g)) return DK_ERR_UNSUPPORTED;
    {   // dY must reach the GEMM through TMA (directly or repacked): the gather loaders assume unpadded plane pitches
        const int64_t P = (int64_t)g.OH * g.OW;
        if (!tma_ok(dy, P) && !repack_wanted(dy, P, 1, true)) return DK_ERR_UNSUPPORTED;
    }
    float *pt = ws_carve(ws, ws_bytes, mat_patch_bytes(N, g));
    if (pt == nullptr) return DK_ERR_UNSUPPORTED;
    const int Kf = g.C * g.kh * g.kw;
    // (the rest of the workspace must hold the split-K partials; pw_wgrad reports DK_ERR_WORKSPACE otherwise)
    if (ws_bytes < tc_conv_ws_bytes(N, Kf, g.OH, g.OW, g.F, 1, 1, 1, 0)) return DK_ERR_UNSUPPORTED;
    int rc = mat_im2col(x, pt, N, g, st);
    if (rc) return rc;
    return pw_wgrad(dy, pt, w, dw, l2, N, Kf, g.OH, g.OW, g.F, 1, ws, ws_bytes, st, repack_pitch((int64_t)g.OH * g.OW));
}

static int cvm_dgrad(const float *dy, const float *w, float *dx, int N, const ConvGeom &g, void *ws, size_t ws_bytes,
                     cudaStream_t st) {
    if (!mat_ok(w, N, g)) return DK_ERR_UNSUPPORTED;
    float *dpt = ws_carve(ws, ws_bytes, mat_patch_bytes(N, g));
    if (dpt == nullptr) return DK_ERR_UNSUPPORTED;
    const int Kf = g.C * g.kh * g.kw, P = g.OH * g.OW;
    int rc = pw_dgrad(dy, w, dpt, N, Kf, g.OH, g.OW, g.F, 1, ws, ws_bytes, st);  // dPt[n][k][p], pitch P
    if (rc) return rc;
    const long long planes = (long long)N * g.C;
    if (g.kh == 4 && g.kw == 4 && g.s == 2) col2im_launch<4, 4, 2>(dpt, dx, g, P, planes, st);
    else if (g.kh == 3 && g.kw == 3 && g.s == 1) col2im_launch<3, 3, 1>(dpt, dx, g, P, planes, st);
    else if (g.kh == 3 && g.kw == 3 && g.s == 2) col2im_launch<3, 3, 2>(dpt, dx, g, P, planes, st);
    else if (g.kh == 5 && g.kw == 5 && g.s == 2) col2im_launch<5, 5, 2>(dpt, dx, g, P, planes, st);
    else col2im_launch<0, 0, 0>(dpt, dx, g, P, planes, st);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int tc_conv_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 1)) return DK_ERR_UNSUPPORTED;
    if (kh == 1 && kw == 1 && p == 0) return pw_fwd(x, w, bias, y, N, C, H, W, F, s, ws, ws_bytes, st);
    if (g_tc_disable_mask & 8) return DK_ERR_UNSUPPORTED;
    {
        const int rc = cvm_fwd(x, w, bias, y, N, mk_geom(C, H, W, F, kh, kw, s, p), ws, ws_bytes, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    return cv_fwd(x, w, bias, y, N, mk_geom(C, H, W, F, kh, kw, s, p), st);
}

int tc_conv_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw, int s,
                  int p, int OH, int OW, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 2)) return DK_ERR_UNSUPPORTED;
    if (kh == 1 && kw == 1 && p == 0) {
        if (H != OH * s || W != OW * s) return DK_ERR_UNSUPPORTED;
        return pw_dgrad(dy, w, dx, N, C, OH, OW, F, s, ws, ws_bytes, st);
    }
    if (g_tc_disable_mask & 8) return DK_ERR_UNSUPPORTED;
    {
        const int rc = cvm_dgrad(dy, w, dx, N, mk_geom(C, H, W, F, kh, kw, s, p), ws, ws_bytes, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    return cv_dgrad(dy, w, dx, N, mk_geom(C, H, W, F, kh, kw, s, p), st);
}

// ---- operands re-pitched ONCE by the caller (dk_pw_pack) ---------------------------------------------------------------
// Planes whose pitch is not a multiple of 16 bytes (7x7) and small strided planes are copied into a padded [rows][Pp] layout
// before they can go through TMA.  Inside the three entry points that copy happens per call: the same x twice per step
// (forward, wgrad), the same dY twice (dgrad, wgrad).  A caller that keeps the packed copies passes them to the *_packed
// variants instead: half the re-pitch launches and traffic of those layers.
size_t tc_pw_pack_bytes(int N, int C, int H, int W, int s) {
    if (!g_tc_ready || N <= 0 || C <= 0 || s < 1) return 0;
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const int64_t P = (int64_t)OH * OW;
    const bool want = s == 1 ? ((g_repack_mask & 1) && (P % 4) != 0 && P <= 4 * TC_REPACK_MAX_P)
                             : ((g_repack_mask & 2) && P <= TC_REPACK_MAX_P);
    return want ? repack_bytes((int64_t)N * C, P) : 0;
}
int tc_pw_pack(const float *x, float *packed, int N, int C, int H, int W, int s, cudaStream_t st) {
    if (tc_pw_pack_bytes(N, C, H, W, s) == 0 || !aligned16(packed)) return DK_ERR_UNSUPPORTED;
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    return repack_launch(x, packed, (int64_t)N * C, H, W, OW, s, (int64_t)OH * OW, st);
}
int tc_pw_fwd_packed(const float *xp, const float *w, const float *bias, float *y, int N, int C, int OH, int OW, int F, void *ws,
                     size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 1)) return DK_ERR_UNSUPPORTED;
    return pw_fwd(xp, w, bias, y, N, C, OH, OW, F, 1, ws, ws_bytes, st, repack_pitch((int64_t)OH * OW));
}
int tc_pw_dgrad_packed(const float *dyp, const float *w, float *dx, int N, int C, int OH, int OW, int F, int s, void *ws,
                       size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 2)) return DK_ERR_UNSUPPORTED;
    return pw_dgrad(dyp, w, dx, N, C, OH, OW, F, s, ws, ws_bytes, st, nullptr, nullptr, nullptr, repack_pitch((int64_t)OH * OW));
}
// dy_packed / x_packed: which of the two operands are packed copies ([N][F][Pp] / [N][C][Pp], stride already applied)
int tc_pw_wgrad_packed(const float *dy, int dy_packed, const float *x, int x_packed, const float *w, float *dw, float l2, int N,
                       int C, int H, int W, int F, int s, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 4)) return DK_ERR_UNSUPPORTED;
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const int64_t Pp = repack_pitch((int64_t)OH * OW);
    if (x_packed) return pw_wgrad(dy, x, w, dw, l2, N, C, OH, OW, F, 1, ws, ws_bytes, st, Pp, dy_packed ? Pp : 0);
    return pw_wgrad(dy, x, w, dw, l2, N, C, H, W, F, s, ws, ws_bytes, st, 0, dy_packed ? Pp : 0);
}

// pointwise dgrad (stride 1) whose epilogue adds cb[c]*x + cd[c]: the backward of the BatchNorm folded into this layer
int tc_pw_dgrad_affine(const float *dy, const float *w, const float *x, const float *cb, const float *cd, float *dx, int N,
                       int C, int OH, int OW, int F, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 2)) return DK_ERR_UNSUPPORTED;
    return pw_dgrad(dy, w, dx, N, C, OH, OW, F, 1, ws, ws_bytes, st, x, cb, cd);
}

int tc_conv_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                  int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!g_tc_ready || (g_tc_disable_mask & 4)) return DK_ERR_UNSUPPORTED;
    if (kh == 1 && kw == 1 && p == 0) return pw_wgrad(dy, x, w, dw, l2, N, C, H, W, F, s, ws, ws_bytes, st);
    if (g_tc_disable_mask & 8) return DK_ERR_UNSUPPORTED;
    {
        const int rc = cvm_wgrad(dy, x, w, dw, l2, N, mk_geom(C, H, W, F, kh, kw, s, p), ws, ws_bytes, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    return cv_wgrad(dy, x, w, dw, l2, N, mk_geom(C, H, W, F, kh, kw, s, p), ws, ws_bytes, st);
}

// ---- DenseLayer (dense_layer.py:46-67): three small GEMMs, all operands through TMA ----------------------------------
// out_dim % 4 != 0 (MNIST: 10 classes): W[in][out] and dY[B][out] have row pitches TMA cannot take; they are copied into
// the workspace with the pitch rounded up to 4 floats (zero padded: padded k add nothing, padded n are masked by the
// epilogue) -- two tiny copies instead of leaving the tensor cores.
static bool dense_ok(const float *a, const float *b, const float *c, int B, int in_dim, int out_dim) {
    return g_tc_ready && !(g_tc_disable_mask & 16) && aligned16(a) && aligned16(b) && aligned16(c) && in_dim % 4 == 0 && B > 0 &&
           in_dim < (1 << 24) && out_dim < (1 << 24);
}
size_t tc_dense_ws_bytes(int B, int in_dim, int out_dim) {
    if (out_dim % 4 == 0) return 0;
    return repack_bytes(in_dim, out_dim) + repack_bytes(B, out_dim) + 1024;
}

int tc_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim, void *ws,
                 size_t ws_bytes, cudaStream_t st) {
    if (!dense_ok(x, w, x, B, in_dim, out_dim)) return DK_ERR_UNSUPPORTED;
    int64_t wp = out_dim;  // row pitch of the W operand
    if (out_dim % 4 != 0) {
        float *wpad = ws_carve(ws, ws_bytes, repack_bytes(in_dim, out_dim));
        if (wpad == nullptr) return DK_ERR_UNSUPPORTED;
        int rc = repack_launch(w, wpad, in_dim, 1, out_dim, out_dim, 1, out_dim, st);
        if (rc) return rc;
        w = wpad; wp = repack_pitch(out_dim);
    }
    TcParams q = {};
    q.mode = 0; q.a_mn = 0; q.b_mn = 1; q.b_batched = 0;
    q.M = B; q.N = out_dim; q.K = in_dim; q.batches = 1;
    fill_common(q);
    q.num_tiles = q.m_blocks * q.n_blocks;
    q.epi = 1; q.out = y; q.bias = bias;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, x, in_dim, B, 1, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B);          // A(m=b, k): k contiguous
    if (rc) return rc;
    rc = make_map(&tb, w, wp, in_dim, 1, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);         // B(k, n) = W[k][n]: n contiguous
    if (rc) return rc;
    return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
}

int tc_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float l2, int B, int in_dim,
                 int out_dim, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!dense_ok(x, w, dx, B, in_dim, out_dim) || !aligned16(dw)) return DK_ERR_UNSUPPORTED;
    const float *w_l2 = w;  // the weight-decay term reads the real (unpadded) weights
    int64_t op = out_dim;   // row pitch of the W and dY operands
    if (out_dim % 4 != 0) {
        float *wpad = ws_carve(ws, ws_bytes, repack_bytes(in_dim, out_dim));
        float *gpad = ws_carve(ws, ws_bytes, repack_bytes(B, out_dim));
        if (wpad == nullptr || gpad == nullptr) return DK_ERR_UNSUPPORTED;
        int rc = repack_launch(w, wpad, in_dim, 1, out_dim, out_dim, 1, out_dim, st);
        if (rc) return rc;
        rc = repack_launch(dy, gpad, B, 1, out_dim, out_dim, 1, out_dim, st);
        if (rc) return rc;
        w = wpad; dy = gpad; op = repack_pitch(out_dim);
    } else if (!aligned16(dy)) {
        return DK_ERR_UNSUPPORTED;
    }
    {   // dx[B, in] = dy[B, out] @ W^T : A(m=b, k=o) = dy, B(n=i, k=o) = W[i][o] (both K-major)
        TcParams q = {};
        q.mode = 0; q.a_mn = 0; q.b_mn = 0; q.b_batched = 0;
        q.M = B; q.N = in_dim; q.K = (int)op; q.batches = 1;
        fill_common(q);
        q.num_tiles = q.m_blocks * q.n_blocks;
        q.epi = 1; q.out = dx;
        CUtensorMap ta, tb;
        int rc = make_map(&ta, dy, op, B, 1, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = make_map(&tb, w, op, in_dim, 1, TC_BK, q.bn, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
        if (rc) return rc;
    }
    {   // dw[in, out] = x^T @ dy (+ l2*w) : A(m=i, k=b) = x[b][i], B(k=b, n=o) = dy[b][o] (both MN-major)
        TcParams q = {};
        q.mode = 0; q.a_mn = 1; q.b_mn = 1; q.b_batched = 0;
        q.M = in_dim; q.N = out_dim; q.K = B; q.batches = 1;
        fill_common(q);
        q.num_tiles = q.m_blocks * q.n_blocks;
        q.epi = 1; q.out = dw;
        if (l2 != 0.0f) { q.epi_w = w_l2; q.epi_l2 = l2; }
        CUtensorMap ta, tb;
        int rc = make_map(&ta, x, in_dim, B, 1, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        rc = make_map(&tb, dy, op, B, 1, 32, TC_BK, (CUtensorMapSwizzle)g_mn_swizzle);
        if (rc) return rc;
        return tc_launch(ta, tb, q, NoGather{}, NoGather{}, st);
    }
}

}  // namespace dk

extern "C" {
/* Bring-up / test knobs for the tensor-core path (not part of the reference-facing ABI):
 * key 0: disable mask (bit0 fwd, bit1 dgrad, bit2 wgrad, bit3 non-pointwise convs); 1: MN-major layout type;
 * 2: LBO bytes; 3: SBO bytes; 4: bytes per 8-deep K step; 5: TMA swizzle enum for MN-major operands. */
int dk_tc_debug_set(int key, int value) {
    switch (key) {
        case 0: dk::g_tc_disable_mask = value; break;
        case 1: dk::g_mn_layout = value; break;
        case 2: dk::g_mn_lbo = value; break;
        case 3: dk::g_mn_sbo = value; break;
        case 4: dk::g_mn_kstep = value; break;
        case 5: dk::g_mn_swizzle = value; break;
        case 6: dk::g_l2_promo = value; break;  // CUtensorMapL2promotion: 0 none, 1 64B, 2 128B, 3 256B
        case 9: dk::g_bn_fused_enabled = value; break;  // 0: BatchNorm through the split kernels of batchnorm.cu only
        case 12: dk::g_smem_budget = value * 1024; break;  // operand stage bytes per CTA (KB)
        case 13:  // persistent CTAs per SM (2: the stage budget is halved so that two CTAs fit)
            dk::g_ctas_per_sm = value;
            dk::g_smem_budget = value >= 2 ? 98 * 1024 : dk::TC_SMEM_BUDGET;
            break;
        case 16: dk::g_hybrid_wgrad = value; break;  // pointwise wgrad: X through the LSU path (16-byte cp.async), dY through TMA
        case 15: dk::g_wide_items = value; break;  // wgrad k-blocks per pipeline item (0 automatic)
        case 14: dk::g_two_per_sm = value; break;  // 0: forward / dgrad GEMMs never run two CTAs per SM
        case 11: dk::g_short_a = value; break;  // 0: wgrad dY boxes always 128 rows
        case 10: dk::g_repack_mask = value; break;  // bit0: pad misaligned planes for TMA, bit1: repack small strided planes
        case 19: dk::g_ct_wgrad2 = value; break;  // 0: conv_tma wgrad through column-shifted global copies only
        case 18: dk::g_ct_kc16 = value; break;  // conv_tma forward / dgrad: 0 = 32-channel stages only
        case 17: dk::g_conv_tma_enabled = value; break;  // 0: stride-1 k x k convolutions skip conv_tma.cu (gather variants instead)
        case 26: dk::g_cw2_rows = value; break;  // conv_tma wgrad: X rows per step (0 auto, 1, 2)
        case 24: dk::g_bn_split_ctas_per_sm = value; break;  // CTAs per SM the split BatchNorm kernels are planned for
        case 23: dk::g_fused_reduce = value; break;  // 0: pointwise wgrad partials go through the separate reduce kernel
        case 22: dk::g_wgrad_bn_cap = value; break;  // 0 automatic, else 64 / 128 / 256
        case 21: dk::g_epi_ring = value; break;  // 0: dk_pwconv_dgrad_affine loads the BatchNorm input from the epilogue warps
        case 20: dk::g_conv_mat_enabled = value; break;  // 0: no materialised-patch path (general convolutions fall to the gather variants)
        case 8: dk::g_conv_rows_enabled = value; break;  // 0: small-K convolutions use the gather loaders, not conv_rows.cu
        default: dk::set_error("dk_tc_debug_set: unknown key %d", key); return DK_ERR_INVALID;
    }
    {
        std::lock_guard<std::mutex> g(dk::g_maps_mu);
        dk::g_maps.clear();
    }
    return DK_OK;
}
}
