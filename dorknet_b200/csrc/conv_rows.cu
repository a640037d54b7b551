// conv_rows.cu -- small-K convolutions (C*kh*kw <= 128: the first layer of every network here) as a row-staged
// implicit GEMM on tcgen05, forward and wgrad.
//
// The reference builds the whole im2col matrix in memory (layers/im2col.pyx:16-36: 25x the input for conv0) and runs
// a GEMM on it (layers/convolution.py:75-100).  Here the patch matrix only ever exists one tile at a time in shared
// memory: the loader warps copy the kh input rows (all C channels) an output row needs into shared memory with
// coalesced cp.async -- 225-float rows are not 16-byte aligned, so TMA cannot describe them -- and expand them there
// into the swizzled operand layout tcgen05.mma reads (same index map as im2col_cy: k = (c*kh + i)*kw + j, value
// X[n, c, oh*s + i - p, ow*s + j - p], zero outside the image).  The whole reduction (Kf <= 128) fits one tile:
//
//   forward  Y[n, :, oh, ow0:ow0+128]  = P[128 pixels, Kf] . W^T     A = P (MN-major, pixels contiguous), B = W (K-major,
//                                                                    resident in shared memory for the whole kernel)
//   wgrad    dW[F, Kf] = sum_(n,oh) dY[n, :, oh, :] . P[OW, Kf]      A = dY row (K-major, TMA), B = P^T (K-major, built),
//                                                                    ONE TMEM accumulator per CTA over all its rows,
//                                                                    then a deterministic cross-CTA reduction (+ l2*W)
//
// Roles (448 threads): warp 0 = TMA producer (wgrad), warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM lane quadrants),
// warps 6-13 = loaders (stage input rows of tile t+1 while expanding tile t).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int CR_LOADER_WARPS = 8;
constexpr int CR_LOADER_THREADS = 32 * CR_LOADER_WARPS;
constexpr int CR_THREADS = 192 + CR_LOADER_THREADS;
constexpr int CR_SMEM_MAX = 224 * 1024;
constexpr int CR_MAX_A_STAGES = 4;

struct ConvRowsParams {
    const float *x;
    const float *w;
    const float *bias;
    float *out;  // forward: y;  wgrad: partial sums [grid][F][Kf]
    int N, C, H, W, F, kh, kw, s, p, OH, OW;
    int Kf, KP, KC;  // filter taps per output; padded (8 forward / 16 wgrad); 32-wide K chunks (forward)
    int bn;          // forward: MMA N (F rounded up to 32)
    int iwt;         // staged input columns per tile
    int phw, irow;   // staged rows are split by column phase (col % s): phw columns per phase, irow = s*phw floats per row
    int in_floats;   // C*kh*irow rounded up to 4 (span staging: C*slot + over-read slack)
    int slot;        // span staging: floats per channel slot (4 front pad + kh*W + alignment slop)
    int in_stages;   // staged input tiles in flight + 1 (the global-load latency is spread over in_stages-1 tiles)
    int tiles_per_row, num_tiles;
    uint32_t tmem_cols, acc_stride;
    int PC, BMR, rows_total, a_stages;  // wgrad: 32-pixel chunks per output row, dY rows per chunk, (n, oh) rows, dY ring depth
};

__device__ __forceinline__ void cr_loader_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(CR_LOADER_THREADS) : "memory"); }

// input rows ih0 .. ih0+kh-1 of every channel of image n, columns iw0 .. iw0+iwt-1 -> dst[C*kh][irow]; zeros outside.
// A staged row is stored split by column phase -- column col lives at (col % s)*phw + col/s -- so that the stride-s
// walk of the expansion (column m*s + j for pixel m) reads consecutive shared-memory words (no bank conflicts).
__device__ __forceinline__ void cr_stage_input(const ConvRowsParams &p, uint32_t dst, int n, int ih0, int iw0, int lw,
                                               int lane) {
    const int rows = p.C * p.kh;
    for (int r = lw; r < rows; r += CR_LOADER_WARPS) {
        const int c = r / p.kh, i = r - c * p.kh;
        const int ih = ih0 + i;
        const bool rok = ih >= 0 && ih < p.H;
        const float *src = p.x + (((long long)n * p.C + c) * p.H + (rok ? ih : 0)) * p.W;
        const uint32_t d = dst + (uint32_t)(r * p.irow) * 4u;
        for (int col = lane; col < p.iwt; col += 32) {
            const int iw = iw0 + col;
            const bool ok = rok && iw >= 0 && iw < p.W;
            int pos;
            if (p.s == 1) pos = col;
            else if (p.s == 2) pos = (col & 1) * p.phw + (col >> 1);
            else pos = (col % p.s) * p.phw + col / p.s;
            cp_async_f32(d + (uint32_t)pos * 4u, ok ? src + iw : nullptr, p.x);
        }
    }
}

// ktab[k] = offset (floats) inside a staged tile of the first pixel's input for filter tap k = (c*kh + i)*kw + j
__device__ __forceinline__ void cr_fill_ktab(const ConvRowsParams &p, int *ktab) {
    const int kk = p.kh * p.kw;
    for (int k = threadIdx.x; k < p.Kf; k += blockDim.x) {
        const int c = k / kk, t = k - c * kk;
        const int i = t / p.kw, j = t - i * p.kw;
        ktab[k] = (c * p.kh + i) * p.irow + (j % p.s) * p.phw + j / p.s;
    }
}


// ---- span staging (fast path) ---------------------------------------------------------------------------------------
// When an output row spans the whole image width, the kh input rows a tile needs are ONE contiguous span per channel
// (kh*W floats).  A single lane per channel pulls it in with cp.async.bulk (16-byte aligned superset of the span: `mis`
// = first wanted float's index mod 4) -- no per-element address arithmetic, no registers, completion on an mbarrier.
// The expansion then gives every thread 4 consecutive output pixels: it reads its 3*S+KW input columns with aligned
// 16-byte shared loads, realigns them in registers (the misalignment is warp-uniform) and writes each filter tap of
// the 4 pixels as ONE 16-byte store into the swizzled operand tile (4 consecutive pixels are contiguous in both the
// MN-major A of the forward and the K-major B of the wgrad).  Per (c, i) row and warp: NL loads + KW stores instead of
// 4*KW loads + 4*KW stores with their index arithmetic.
__device__ __forceinline__ void cr_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// one warp: rows ih0 .. ih0+kh-1 (clipped to the image) of every channel of image n -> stage slots
__device__ __forceinline__ void cr_span_issue(const ConvRowsParams &p, uint32_t dst_stage, uint32_t bar, int n, int ih0, int lane) {
    const int lo = ih0 > 0 ? ih0 : 0;
    int hi = ih0 + p.kh - 1;
    if (hi > p.H - 1) hi = p.H - 1;
    uint32_t bytes = 0;
    const float *src = p.x;
    if (lane < p.C && hi >= lo) {
        const long long g0 = (((long long)n * p.C + lane) * p.H + lo) * p.W;
        const long long g1 = g0 + (long long)(hi - lo + 1) * p.W;
        const long long al = g0 & ~3ll;
        bytes = (uint32_t)(((g1 - al) + 3ll) & ~3ll) * 4u;
        src = p.x + al;
    }
    uint32_t total = bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (bytes) cr_bulk_g2s(dst_stage + (uint32_t)(lane * p.slot + 4) * 4u, src, bytes, bar);
}

// expand row (c, i) of a staged tile: taps k = kbase .. kbase+KW-1 of pixels 4*lane .. 4*lane+3
template <int S, int KW, bool WGRAD>
__device__ __forceinline__ void cr_expand_row(const float *stage, int row_off, bool rok, int W, int pad, int kbase, int lane,
                                              uint32_t tile_base, uint32_t chunk_bytes) {
    constexpr int WIN = 3 * S + KW;
    constexpr int NL = (WIN + 3 + 3) / 4;
    float w[WIN];
    if (rok) {
        const int a = row_off + S * 4 * lane;
        const int R = a & 3;  // warp-uniform: S*4*lane is a multiple of 4
        const float4 *src = reinterpret_cast<const float4 *>(stage + (a - R));
        float r[NL * 4];
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            const float4 v = src[u];
            r[4 * u] = v.x; r[4 * u + 1] = v.y; r[4 * u + 2] = v.z; r[4 * u + 3] = v.w;
        }
        if (R == 0) {
#pragma unroll
            for (int t = 0; t < WIN; ++t) w[t] = r[t];
        } else if (R == 1) {
#pragma unroll
            for (int t = 0; t < WIN; ++t) w[t] = r[t + 1];
        } else if (R == 2) {
#pragma unroll
            for (int t = 0; t < WIN; ++t) w[t] = r[t + 2];
        } else {
#pragma unroll
            for (int t = 0; t < WIN; ++t) w[t] = r[t + 3];
        }
        // zero padding columns: window element t is input column S*4*lane - pad + t
        const int lo_bad = pad - S * 4 * lane, hi_ok = W + pad - S * 4 * lane;
        if (lo_bad > 0 || hi_ok < WIN) {
#pragma unroll
            for (int t = 0; t < WIN; ++t)
                if (t < lo_bad || t >= hi_ok) w[t] = 0.0f;
        }
    } else {
#pragma unroll
        for (int t = 0; t < WIN; ++t) w[t] = 0.0f;
    }
    const uint32_t l = (uint32_t)lane;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const uint32_t k = (uint32_t)(kbase + j);
        uint32_t dst;
        if (WGRAD) dst = tile_base + (l >> 3) * chunk_bytes + k * 128u + (((l & 7u) ^ (k & 7u)) << 4);  // km_tile_off(k, 4*lane % 32)
        else dst = tile_base + (k >> 5) * 16384u + (l >> 3) * 4096u + (k & 31u) * 128u + ((((l & 7u) >> 1) ^ (k & 3u)) << 5) +
                   ((l & 1u) << 4);  // mn_tile_off(4*lane, k)
        st_shared_v4(dst, w[j], w[S + j], w[2 * S + j], w[3 * S + j]);
    }
}

// all loader warps: expand staged tile (n, oh) into an operand tile
template <int S, int KW, bool WGRAD>
__device__ __forceinline__ void cr_expand_tile(const ConvRowsParams &p, const float *stage, int n, int ih0, int lw, int lane,
                                               uint32_t tile_base, uint32_t chunk_bytes) {
    if (4 * lane >= p.OW) return;
    const int lo = ih0 > 0 ? ih0 : 0;
    const int rows = p.C * p.kh;
    for (int r = lw; r < rows; r += CR_LOADER_WARPS) {
        const int c = r / p.kh, i = r - c * p.kh;
        const int ih = ih0 + i;
        const bool rok = ih >= 0 && ih < p.H;
        const int mis = (int)(((((long long)n * p.C + c) * p.H + lo) * p.W) & 3ll);
        const int row_off = c * p.slot + 4 + mis + (ih - lo) * p.W - p.p;
        cr_expand_row<S, KW, WGRAD>(stage, row_off, rok, p.W, p.p, r * KW, lane, tile_base, chunk_bytes);
    }
}


// The same expansion with everything that does not depend on the tile hoisted out of the tile loop: a loader warp always
// expands the same (c, i) rows (r = lw, lw + 8), so their destination addresses, tap indices and edge masks are computed
// once per kernel; per tile only the row validity, the span misalignment and the source offset change.  (ncu on the first
// span-staging version: ~480 instructions per loader warp and tile, most of them this index arithmetic.)
constexpr int CR_ROWS_PER_WARP = 2;  // C*kh <= 16 (conv0: 15, MobileNet conv0: 9, MNIST conv0: 3); else cr_expand_tile

template <int KW>
struct CrRowPlan {
    int nrows;                              // rows this warp expands (0 .. CR_ROWS_PER_WARP)
    int rowconst[CR_ROWS_PER_WARP];         // c*slot + 4 + i*W - pad
    int chw[CR_ROWS_PER_WARP];              // c*H*W (low bits only matter)
    int irow[CR_ROWS_PER_WARP];             // i
    uint32_t dst[CR_ROWS_PER_WARP][KW];     // byte offset of tap j of this lane's 4 pixels inside an operand tile
    int lo_bad, hi_ok;                      // window elements t < lo_bad or t >= hi_ok are padding columns
    bool active;                            // this lane's 4 pixels exist (4*lane < OW)
};

template <int S, int KW, bool WGRAD>
__device__ __forceinline__ void cr_plan_rows(const ConvRowsParams &p, int lw, int lane, uint32_t chunk_bytes, CrRowPlan<KW> &pl) {
    const int rows = p.C * p.kh;
    pl.nrows = 0;
    const uint32_t l = (uint32_t)lane;
#pragma unroll
    for (int u = 0; u < CR_ROWS_PER_WARP; ++u) {
        const int r = lw + u * CR_LOADER_WARPS;
        pl.rowconst[u] = 0; pl.chw[u] = 0; pl.irow[u] = 0;
        if (r < rows) {
            const int c = r / p.kh, i = r - c * p.kh;
            pl.nrows = u + 1;
            pl.rowconst[u] = c * p.slot + 4 + i * p.W - p.p;
            pl.chw[u] = c * p.H * p.W;
            pl.irow[u] = i;
        }
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const uint32_t k = (uint32_t)(r * KW + j);
            if (WGRAD) pl.dst[u][j] = (l >> 3) * chunk_bytes + k * 128u + (((l & 7u) ^ (k & 7u)) << 4);
            else pl.dst[u][j] = (k >> 5) * 16384u + (l >> 3) * 4096u + (k & 31u) * 128u + ((((l & 7u) >> 1) ^ (k & 3u)) << 5) + ((l & 1u) << 4);
        }
    }
    pl.lo_bad = p.p - S * 4 * lane;
    pl.hi_ok = p.W + p.p - S * 4 * lane;
    pl.active = 4 * lane < p.OW;
}

template <int S, int KW>
__device__ __forceinline__ void cr_expand_planned(const ConvRowsParams &p, const CrRowPlan<KW> &pl, const float *stage, int n,
                                                  int ih0, int lane, uint32_t tile_base) {
    if (!pl.active) return;
    constexpr int WIN = 3 * S + KW;
    constexpr int NL = (WIN + 3 + 3) / 4;
    const int lo = ih0 > 0 ? ih0 : 0;
    const uint32_t nbase = ((uint32_t)n * (uint32_t)(p.C * p.H) + (uint32_t)lo) * (uint32_t)p.W;  // (only the low two bits are used)
    const int dlo = (ih0 - lo) * p.W;
#pragma unroll
    for (int u = 0; u < CR_ROWS_PER_WARP; ++u) {
        if (u < pl.nrows) {
            const int ih = ih0 + pl.irow[u];
            const bool rok = ih >= 0 && ih < p.H;
            float w[WIN];
            if (rok) {
                const int mis = (int)((nbase + (uint32_t)pl.chw[u]) & 3u);
                const int a = pl.rowconst[u] + mis + dlo + S * 4 * lane;
                const int R = a & 3;
                const float4 *src = reinterpret_cast<const float4 *>(stage + (a - R));
                float r[NL * 4];
#pragma unroll
                for (int q = 0; q < NL; ++q) {
                    const float4 v = src[q];
                    r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
                }
                if (R == 0) {
#pragma unroll
                    for (int t = 0; t < WIN; ++t) w[t] = r[t];
                } else if (R == 1) {
#pragma unroll
                    for (int t = 0; t < WIN; ++t) w[t] = r[t + 1];
                } else if (R == 2) {
#pragma unroll
                    for (int t = 0; t < WIN; ++t) w[t] = r[t + 2];
                } else {
#pragma unroll
                    for (int t = 0; t < WIN; ++t) w[t] = r[t + 3];
                }
                if (pl.lo_bad > 0 || pl.hi_ok < WIN) {
#pragma unroll
                    for (int t = 0; t < WIN; ++t)
                        if (t < pl.lo_bad || t >= pl.hi_ok) w[t] = 0.0f;
                }
            } else {
#pragma unroll
                for (int t = 0; t < WIN; ++t) w[t] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < KW; ++j) st_shared_v4(tile_base + pl.dst[u][j], w[j], w[S + j], w[2 * S + j], w[3 * S + j]);
        }
    }
}

// =====================================================================================================================
// forward
// =====================================================================================================================
// S = 0: generic staging (4-byte cp.async, any geometry); S > 0: span staging for stride S, KW filter columns
template <int S, int KW>
__global__ void __launch_bounds__(CR_THREADS, 1) conv_rows_fwd_kernel(const ConvRowsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t a_bytes = (uint32_t)p.KC * 16384u;
    const uint32_t b_bytes = (uint32_t)p.KC * (uint32_t)p.bn * 128u;
    const uint32_t in_bytes = (uint32_t)p.in_floats * 4u;
    const uint32_t sA = base, sB = sA + 2u * a_bytes, sIn = sB + b_bytes, bar = sIn + (uint32_t)p.in_stages * in_bytes;
    auto a_full = [&](int a) { return bar + 8u * a; };
    auto a_empty = [&](int a) { return bar + 8u * (2 + a); };
    auto t_full = [&](int a) { return bar + 8u * (4 + a); };
    auto t_empty = [&](int a) { return bar + 8u * (6 + a); };
    const uint32_t tmem_slot = bar + 64u;
    auto in_full = [&](int st) { return bar + 80u + 8u * st; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int ktab[128];
    if (S == 0) cr_fill_ktab(p, ktab);

    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; ++a) {
            mbar_init(a_full(a), CR_LOADER_WARPS);
            mbar_init(a_empty(a), 1);
            mbar_init(t_full(a), 1);
            mbar_init(t_empty(a), 4);
        }
        for (int st = 0; st < p.in_stages; ++st) mbar_init(in_full(st), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    {   // zero both A buffers once (the K rows past Kf stay zero for the whole kernel) and park the filters in B
        float4 *za = reinterpret_cast<float4 *>(gbase);
        const int n4 = (int)(2u * a_bytes / 16u);
        for (int i = threadIdx.x; i < n4; i += CR_THREADS) za[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int kw32 = p.KC * 32, total = p.bn * kw32;
        for (int idx = threadIdx.x; idx < total; idx += CR_THREADS) {
            const int f = idx / kw32, k = idx - f * kw32;
            const float v = (f < p.F && k < p.Kf) ? __ldg(p.w + (long long)f * p.Kf + k) : 0.0f;
            st_shared_f32(sB + (uint32_t)(k >> 5) * (uint32_t)p.bn * 128u + km_tile_off(f, k & 31), v);
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 1) {
        // ================================ MMA issuer ==================================
        // whole warp, uniform control flow, one elected lane issues (no per-MMA ELECT / R2UR waterfall); descriptors as
        // constant upper and incremented lower words
        {
            const uint32_t idesc = idesc_tf32(128, p.bn, 1, 0);
            const uint32_t a_hi = smem_desc_hi(512u, LAYOUT_SW128_BASE32B), b_hi = smem_desc_hi(1024u, LAYOUT_SW128);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(t_empty(a), ph ^ 1u);
                mbar_wait(a_full(a), ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)a * p.acc_stride;
                const uint32_t sAa = sA + (uint32_t)a * a_bytes;
                uint32_t accum = 0;
#pragma unroll 1
                for (int kc = 0; kc < p.KC; ++kc) {
                    const int rem = p.KP - kc * 32;
                    const int nks = rem >= 32 ? 4 : rem / 8;
                    const uint32_t a_lo = smem_desc_lo(sAa + (uint32_t)kc * 16384u, 4096u);
                    const uint32_t b_lo = smem_desc_lo(sB + (uint32_t)kc * (uint32_t)p.bn * 128u, 16u);
                    if (nks == 4) {
                        if (elect_one()) mma_tf32_k4(d_tmem, a_lo, a_hi, b_lo, b_hi, 64u, 2u, idesc, accum);
                    } else {
#pragma unroll 1
                        for (int ks = 0; ks < nks; ++ks)
                            if (elect_one())
                                mma_tf32_lohi(d_tmem, a_lo + 64u * (uint32_t)ks, a_hi, b_lo + 2u * (uint32_t)ks, b_hi, idesc, (accum || ks > 0) ? 1u : 0u);
                    }
                    accum = 1;
                }
                if (elect_one()) {
                    mma_commit(a_empty(a));
                    mma_commit(t_full(a));
                }
            }
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ================================ loaders =====================================
        const int lw = warp - 6;
        const float *in_f = reinterpret_cast<const float *>(gbase + (sIn - base));
        auto tile_coords = [&](int tile, int &n, int &oh, int &ow0) {
            const int owb = tile % p.tiles_per_row;
            const int r = tile / p.tiles_per_row;
            oh = r % p.OH;
            n = r / p.OH;
            ow0 = owb * 128;
        };
        int n, oh, ow0;
        const int D = p.in_stages - 1;  // prefetch distance in tiles
        if constexpr (S > 0) {
            // span staging: tiles_per_row == 1, so tile = (n, oh)
            if (lw == 0) {
                for (int d = 0; d < D; ++d) {
                    const long long t = (long long)blockIdx.x + (long long)d * gridDim.x;
                    if (t < p.num_tiles) {
                        tile_coords((int)t, n, oh, ow0);
                        cr_span_issue(p, sIn + (uint32_t)d * in_bytes, in_full(d), n, oh * p.s - p.p, lane);
                    }
                }
            }
            const bool planned = p.C * p.kh <= CR_ROWS_PER_WARP * CR_LOADER_WARPS;
            CrRowPlan<KW> pl;
            cr_plan_rows<S, KW, false>(p, lw, lane, 0u, pl);
            // (n, oh) of this CTA's tiles advance by gridDim.x rows per iteration: no division in the loop
            int tn = (int)blockIdx.x / p.OH, toh = (int)blockIdx.x - tn * p.OH;
            const int dn = (int)gridDim.x / p.OH, doh = (int)gridDim.x - dn * p.OH;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const int stg = it % p.in_stages;
                mbar_wait(in_full(stg), (uint32_t)(it / p.in_stages) & 1u);
                cr_loader_barrier();  // every loader is done expanding tile it-1, whose buffer is refilled now
                const long long next = (long long)tile + (long long)D * gridDim.x;
                if (lw == 0 && next < p.num_tiles) {
                    const int sn = (it + D) % p.in_stages;
                    tile_coords((int)next, n, oh, ow0);
                    cr_span_issue(p, sIn + (uint32_t)sn * in_bytes, in_full(sn), n, oh * p.s - p.p, lane);
                }
                const int a = it & 1;
                mbar_wait(a_empty(a), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                if (planned)
                    cr_expand_planned<S, KW>(p, pl, in_f + (size_t)stg * p.in_floats, tn, toh * p.s - p.p, lane,
                                             sA + (uint32_t)a * a_bytes);
                else
                    cr_expand_tile<S, KW, false>(p, in_f + (size_t)stg * p.in_floats, tn, toh * p.s - p.p, lw, lane,
                                                 sA + (uint32_t)a * a_bytes, 0u);
                tn += dn;
                toh += doh;
                if (toh >= p.OH) { toh -= p.OH; ++tn; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full(a));
            }
        } else {
        for (int d = 0; d < D; ++d) {
            const long long t = (long long)blockIdx.x + (long long)d * gridDim.x;
            if (t < p.num_tiles) {
                tile_coords((int)t, n, oh, ow0);
                cr_stage_input(p, sIn + (uint32_t)d * in_bytes, n, oh * p.s - p.p, ow0 * p.s - p.p, lw, lane);
            }
            cp_async_commit();
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            cp_async_wait_pending(D - 1);
            cr_loader_barrier();  // tile `it` is staged; every loader is done expanding tile it-1 (whose buffer is reused now)
            const long long next = (long long)tile + (long long)D * gridDim.x;
            if (next < p.num_tiles) {
                tile_coords((int)next, n, oh, ow0);
                cr_stage_input(p, sIn + (uint32_t)((it + D) % p.in_stages) * in_bytes, n, oh * p.s - p.p, ow0 * p.s - p.p, lw, lane);
            }
            cp_async_commit();
            const int a = it & 1;
            mbar_wait(a_empty(a), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            const float *in = in_f + (size_t)(it % p.in_stages) * p.in_floats;
            const uint32_t sAa = sA + (uint32_t)a * a_bytes;
            const uint32_t lane_off = (((uint32_t)lane & 7u) << 2);
#pragma unroll 2
            for (int k = lw; k < p.Kf; k += CR_LOADER_WARPS) {
                const float *src = in + ktab[k] + lane;
                const uint32_t dst = sAa + (uint32_t)(k >> 5) * 16384u + (uint32_t)(k & 31) * 128u +
                                     ((((uint32_t)lane >> 3) ^ ((uint32_t)k & 3u)) << 5) + lane_off;  // mn_tile_off(lane, k & 31)
                const float v0 = src[0], v1 = src[32], v2 = src[64], v3 = src[96];
                st_shared_f32(dst, v0);
                st_shared_f32(dst + 4096u, v1);
                st_shared_f32(dst + 8192u, v2);
                st_shared_f32(dst + 12288u, v3);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(a));
        }
        cp_async_wait_pending(0);
        }
    } else if (warp >= 2) {
        // ================================ epilogue ====================================
        const int q = warp & 3;
        const long long plane = (long long)p.OH * p.OW;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const int owb = tile % p.tiles_per_row;
            const int r = tile / p.tiles_per_row;
            const int oh = r % p.OH, n = r / p.OH;
            const int ow = owb * 128 + 32 * q + lane;
            mbar_wait(t_full(a), (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)a * p.acc_stride;
            const bool ok = ow < p.OW;
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t_row + (uint32_t)c0, v);
                tmem_ld_wait();
                if (ok) {
                    float *o = p.out + ((long long)n * p.F + c0) * plane + (long long)oh * p.OW + ow;
                    if (c0 + 32 <= p.F && p.bias == nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[(long long)j * plane] = __uint_as_float(v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (c0 + j < p.F) {
                                float rv = __uint_as_float(v[j]);
                                if (p.bias) rv += __ldg(p.bias + c0 + j);
                                o[(long long)j * plane] = rv;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty(a));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =====================================================================================================================
// wgrad
// =====================================================================================================================
template <int S, int KW>
__global__ void __launch_bounds__(CR_THREADS, 1)
conv_rows_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const ConvRowsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw);
    const uint32_t a_chunk = (uint32_t)p.BMR * 128u, a_bytes = (uint32_t)p.PC * a_chunk;
    const uint32_t b_chunk = (uint32_t)p.KP * 128u, b_bytes = (uint32_t)p.PC * b_chunk;
    const uint32_t in_bytes = (uint32_t)p.in_floats * 4u;
    // the dY ring comes first: with F < 128 the M=128 MMA reads 128-BMR rows past a chunk, which must stay inside
    // the allocation (those accumulator rows are never read back)
    const uint32_t sA = base, sB = sA + (uint32_t)p.a_stages * a_bytes, sIn = sB + 2u * b_bytes, bar = sIn + (uint32_t)p.in_stages * in_bytes;
    auto a_full = [&](int a) { return bar + 8u * a; };
    auto a_empty = [&](int a) { return bar + 8u * (CR_MAX_A_STAGES + a); };
    auto b_full = [&](int a) { return bar + 8u * (2 * CR_MAX_A_STAGES + a); };
    auto b_empty = [&](int a) { return bar + 8u * (2 * CR_MAX_A_STAGES + 2 + a); };
    const uint32_t t_full = bar + 8u * (2 * CR_MAX_A_STAGES + 4);
    const uint32_t tmem_slot = t_full + 8u;
    auto in_full = [&](int st) { return t_full + 16u + 8u * st; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ int ktab[128];
    if (S == 0) cr_fill_ktab(p, ktab);

    // this CTA's contiguous range of (n, oh) rows
    const int r_beg = (int)(((long long)p.rows_total * blockIdx.x) / gridDim.x);
    const int r_end = (int)(((long long)p.rows_total * (blockIdx.x + 1)) / gridDim.x);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDy);
        for (int a = 0; a < p.a_stages; ++a) {
            mbar_init(a_full(a), 1);
            mbar_init(a_empty(a), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(b_full(a), CR_LOADER_WARPS);
            mbar_init(b_empty(a), 1);
        }
        mbar_init(t_full, 1);
        for (int st = 0; st < p.in_stages; ++st) mbar_init(in_full(st), 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    {   // zero the patch tiles once: rows Kf..KP-1 stay zero; and clear the dY ring so that no NaN bit pattern of a
        // previous kernel sits in rows the TMA never writes
        float4 *z = reinterpret_cast<float4 *>(gbase);
        const int n4 = (int)(((uint32_t)p.a_stages * a_bytes + 2u * b_bytes) / 16u);
        for (int i = threadIdx.x; i < n4; i += CR_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ================================ TMA producer: dY rows ======================
        if (lane == 0) {
            int it = 0;
            for (int r = r_beg; r < r_end; ++r, ++it) {
                const int sa = it % p.a_stages;
                const uint32_t ph = (uint32_t)(it / p.a_stages) & 1u;
                const int n = r / p.OH, oh = r - n * p.OH;
                mbar_wait(a_empty(sa), ph ^ 1u);
                mbar_expect_tx(a_full(sa), a_bytes);
                for (int pc = 0; pc < p.PC; ++pc)
                    tma_load_3d(sA + (uint32_t)sa * a_bytes + (uint32_t)pc * a_chunk, &tmDy, a_full(sa), pc * 32, oh, n * p.F);
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // (whole warp, uniform control flow, one elected lane issues: see conv_rows_fwd_kernel)
        {
            const uint32_t idesc = idesc_tf32(128, p.KP, 0, 0);
            const uint32_t d_hi = smem_desc_hi(1024u, LAYOUT_SW128);
            uint32_t accum = 0;
            int it = 0, sa = 0;
            uint32_t aph = 0;
            for (int r = r_beg; r < r_end; ++r, ++it) {
                const int sb = it & 1;
                mbar_wait(a_full(sa), aph);
                mbar_wait(b_full(sb), (uint32_t)(it >> 1) & 1u);
                tc_fence_after();
#pragma unroll 1
                for (int pc = 0; pc < p.PC; ++pc) {
                    const int rem = p.OW - pc * 32;
                    const int nks = rem >= 32 ? 4 : (rem + 7) / 8;
                    const uint32_t a_lo = smem_desc_lo(sA + (uint32_t)sa * a_bytes + (uint32_t)pc * a_chunk, 16u);
                    const uint32_t b_lo = smem_desc_lo(sB + (uint32_t)sb * b_bytes + (uint32_t)pc * b_chunk, 16u);
                    if (nks == 4) {
                        if (elect_one()) mma_tf32_k4(tmem_base, a_lo, d_hi, b_lo, d_hi, 2u, 2u, idesc, accum);
                    } else {
#pragma unroll 1
                        for (int ks = 0; ks < nks; ++ks)
                            if (elect_one())
                                mma_tf32_lohi(tmem_base, a_lo + 2u * (uint32_t)ks, d_hi, b_lo + 2u * (uint32_t)ks, d_hi, idesc, (accum || ks > 0) ? 1u : 0u);
                    }
                    accum = 1;
                }
                if (elect_one()) {
                    mma_commit(a_empty(sa));
                    mma_commit(b_empty(sb));
                }
                if (++sa == p.a_stages) { sa = 0; aph ^= 1u; }
            }
            if (elect_one()) mma_commit(t_full);
            __syncwarp();
        }
    } else if (warp >= 6) {
        // ================================ loaders =====================================
        const int lw = warp - 6;
        const float *in_f = reinterpret_cast<const float *>(gbase + (sIn - base));
        const int D = p.in_stages - 1;
        if constexpr (S > 0) {
            if (lw == 0) {
                for (int d = 0; d < D; ++d) {
                    if (r_beg + d < r_end) {
                        const int n = (r_beg + d) / p.OH, oh = (r_beg + d) - n * p.OH;
                        cr_span_issue(p, sIn + (uint32_t)d * in_bytes, in_full(d), n, oh * p.s - p.p, lane);
                    }
                }
            }
            const bool planned = p.C * p.kh <= CR_ROWS_PER_WARP * CR_LOADER_WARPS;
            CrRowPlan<KW> pl;
            cr_plan_rows<S, KW, true>(p, lw, lane, b_chunk, pl);
            int tn = r_beg / p.OH, toh = r_beg - tn * p.OH;
            int it = 0;
            for (int r = r_beg; r < r_end; ++r, ++it) {
                const int stg = it % p.in_stages;
                mbar_wait(in_full(stg), (uint32_t)(it / p.in_stages) & 1u);
                cr_loader_barrier();
                if (lw == 0 && r + D < r_end) {
                    const int sn = (it + D) % p.in_stages;
                    const int n = (r + D) / p.OH, oh = (r + D) - n * p.OH;
                    cr_span_issue(p, sIn + (uint32_t)sn * in_bytes, in_full(sn), n, oh * p.s - p.p, lane);
                }
                const int sb = it & 1;
                mbar_wait(b_empty(sb), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                if (planned)
                    cr_expand_planned<S, KW>(p, pl, in_f + (size_t)stg * p.in_floats, tn, toh * p.s - p.p, lane,
                                             sB + (uint32_t)sb * b_bytes);
                else
                    cr_expand_tile<S, KW, true>(p, in_f + (size_t)stg * p.in_floats, tn, toh * p.s - p.p, lw, lane,
                                                sB + (uint32_t)sb * b_bytes, b_chunk);
                if (++toh >= p.OH) { toh = 0; ++tn; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_full(sb));
            }
        } else {
        for (int d = 0; d < D; ++d) {
            if (r_beg + d < r_end) {
                const int n = (r_beg + d) / p.OH, oh = (r_beg + d) - n * p.OH;
                cr_stage_input(p, sIn + (uint32_t)d * in_bytes, n, oh * p.s - p.p, -p.p, lw, lane);
            }
            cp_async_commit();
        }
        int it = 0;
        for (int r = r_beg; r < r_end; ++r, ++it) {
            cp_async_wait_pending(D - 1);
            cr_loader_barrier();
            if (r + D < r_end) {
                const int n = (r + D) / p.OH, oh = (r + D) - n * p.OH;
                cr_stage_input(p, sIn + (uint32_t)((it + D) % p.in_stages) * in_bytes, n, oh * p.s - p.p, -p.p, lw, lane);
            }
            cp_async_commit();
            const int sb = it & 1;
            mbar_wait(b_empty(sb), ((uint32_t)(it >> 1) & 1u) ^ 1u);
            const float *in = in_f + (size_t)(it % p.in_stages) * p.in_floats;
            const uint32_t sBb = sB + (uint32_t)sb * b_bytes;
            const uint32_t lane_off = (((uint32_t)lane & 3u) << 2);
#pragma unroll 2
            for (int k = lw; k < p.Kf; k += CR_LOADER_WARPS) {
                const float *src = in + ktab[k] + lane;
                const uint32_t dst = sBb + (uint32_t)k * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)k & 7u)) << 4) + lane_off;  // km_tile_off(k, lane)
#pragma unroll 4
                for (int pc = 0; pc < p.PC; ++pc) {
                    const float v = pc * 32 + lane < p.OW ? src[pc * 32] : 0.0f;
                    st_shared_f32(dst + (uint32_t)pc * b_chunk, v);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_full(sb));
        }
        cp_async_wait_pending(0);
        }
    } else if (warp >= 2) {
        // ================================ epilogue: partial dW of this CTA ============
        const int q = warp & 3;
        const int f = 32 * q + lane;
        mbar_wait(t_full, 0u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16);
        float *o = p.out + ((long long)blockIdx.x * p.F + f) * p.Kf;
        for (int c0 = 0; c0 < p.KP; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(t_row + (uint32_t)c0, v);
            tmem_ld_wait();
            if (f < p.F) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < p.Kf) o[c0 + j] = __uint_as_float(v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn3)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 g_encode3 = nullptr;
static bool g_cr_ready = false;
int g_conv_rows_enabled = 1;  // 0: off; 1: span staging where the geometry allows; 2: generic staging only

int init_conv_rows() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return DK_OK;
    }
    g_encode3 = reinterpret_cast<EncodeTiledFn3>(fn);
#define CR_SET_ATTR(S, KW)                                                                                                       \
    DK_CUDA(cudaFuncSetAttribute(conv_rows_fwd_kernel<S, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_MAX));       \
    DK_CUDA(cudaFuncSetAttribute(conv_rows_wgrad_kernel<S, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM_MAX));
    CR_SET_ATTR(0, 0) CR_SET_ATTR(1, 3) CR_SET_ATTR(1, 5) CR_SET_ATTR(2, 3) CR_SET_ATTR(2, 5)
#undef CR_SET_ATTR
    g_cr_ready = true;
    return DK_OK;
}

static int cr_round_up(int v, int m) { return (v + m - 1) / m * m; }

// span staging (cr_span_issue / cr_expand_tile): one output row per tile, bulk-copyable input, instantiated (stride, kw)
static bool cr_span_ok(const ConvRowsParams &q) {
    const bool inst = (q.s == 1 || q.s == 2) && (q.kw == 3 || q.kw == 5);
    return g_conv_rows_enabled == 1 && inst && q.OW <= 128 && q.p <= 4 && q.C <= 32 && aligned16(q.x) &&
           ((long long)q.N * q.C * q.H * q.W) % 4 == 0;
}
static void cr_span_layout(ConvRowsParams &q) {
    q.slot = 4 + cr_round_up(q.kh * q.W + 3, 4) + 4;
    q.in_floats = q.C * q.slot + 128 * q.s + 64;  // + what pixels past OW over-read (never used)
}
#define CR_DISPATCH(KERNEL, q, ...)                                                \
    do {                                                                           \
        if (!span) KERNEL<0, 0> __VA_ARGS__;                                       \
        else if (q.s == 1 && q.kw == 3) KERNEL<1, 3> __VA_ARGS__;                  \
        else if (q.s == 1 && q.kw == 5) KERNEL<1, 5> __VA_ARGS__;                  \
        else if (q.s == 2 && q.kw == 3) KERNEL<2, 3> __VA_ARGS__;                  \
        else KERNEL<2, 5> __VA_ARGS__;                                             \
    } while (0)
static uint32_t cr_tmem_cols(int n) { return n <= 32 ? 32u : n <= 64 ? 64u : n <= 128 ? 128u : n <= 256 ? 256u : 512u; }

static bool cr_common(ConvRowsParams &q, int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    q.N = N; q.C = C; q.H = H; q.W = W; q.F = F; q.kh = kh; q.kw = kw; q.s = s; q.p = p;
    q.OH = (H + 2 * p - kh) / s + 1;
    q.OW = (W + 2 * p - kw) / s + 1;
    q.Kf = C * kh * kw;
    return g_cr_ready && g_conv_rows_enabled && q.Kf <= 128 && q.OH >= 1 && q.OW >= 1 && !(kh == 1 && kw == 1);
}

int conv_rows_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                  int kw, int s, int p, cudaStream_t st) {
    ConvRowsParams q = {};
    q.x = x; q.w = w; q.bias = bias; q.out = y;
    if (!cr_common(q, N, C, H, W, F, kh, kw, s, p) || F > 256) return DK_ERR_UNSUPPORTED;
    q.KP = cr_round_up(q.Kf, 8);
    q.KC = (q.KP + 31) / 32;
    q.bn = cr_round_up(F, 32);
    q.iwt = 127 * s + kw;
    q.phw = (q.iwt + s - 1) / s;
    q.irow = s * q.phw;
    q.in_floats = cr_round_up(C * kh * q.irow, 4);
    const bool span = cr_span_ok(q);
    if (span) cr_span_layout(q);
    q.tiles_per_row = (q.OW + 127) / 128;
    const long long tiles = (long long)N * q.OH * q.tiles_per_row;
    if (tiles >= (1ll << 31)) return DK_ERR_UNSUPPORTED;
    q.num_tiles = (int)tiles;
    q.acc_stride = cr_tmem_cols(q.bn);
    q.tmem_cols = 2 * q.acc_stride;
    const size_t fixed = 1024 + 2 * (size_t)q.KC * 16384 + (size_t)q.KC * q.bn * 128 + 128;
    q.in_stages = 4;
    while (q.in_stages > 2 && fixed + (size_t)q.in_stages * q.in_floats * 4 > (size_t)CR_SMEM_MAX) --q.in_stages;
    const size_t smem = fixed + (size_t)q.in_stages * q.in_floats * 4;
    if (smem > (size_t)CR_SMEM_MAX) return DK_ERR_UNSUPPORTED;
    const int grid = q.num_tiles < sm_count() ? q.num_tiles : sm_count();
    CR_DISPATCH(conv_rows_fwd_kernel, q, <<<grid, CR_THREADS, smem, st>>>(q));
    DK_LAUNCH_CHECK();
    return DK_OK;
}

static int cr_wgrad_plan(ConvRowsParams &q, size_t *smem, int *grid, bool *span_out) {
    if (q.F > 128 || q.OW > 128 || (q.OW % 4) != 0) return DK_ERR_UNSUPPORTED;
    q.KP = cr_round_up(q.Kf, 16);
    q.PC = (q.OW + 31) / 32;
    q.BMR = cr_round_up(q.F, 8);
    q.iwt = (q.OW - 1) * q.s + q.kw;
    q.phw = (q.iwt + q.s - 1) / q.s;
    q.irow = q.s * q.phw;
    q.in_floats = cr_round_up(q.C * q.kh * q.irow, 4);
    *span_out = cr_span_ok(q);
    if (*span_out) cr_span_layout(q);
    q.rows_total = q.N * q.OH;
    q.tmem_cols = cr_tmem_cols(q.KP);
    const size_t a_bytes = (size_t)q.PC * q.BMR * 128, b_bytes = (size_t)q.PC * q.KP * 128;
    // two dY stages first (the HBM stream), then up to four input stages, then more dY stages
    q.in_stages = 4;
    size_t fixed = 1024 + 2 * b_bytes + 256;
    while (q.in_stages > 2 && fixed + 2 * a_bytes + (size_t)q.in_stages * q.in_floats * 4 > (size_t)CR_SMEM_MAX) --q.in_stages;
    fixed += (size_t)q.in_stages * q.in_floats * 4;
    int stages = CR_MAX_A_STAGES;
    while (stages > 1 && fixed + stages * a_bytes > (size_t)CR_SMEM_MAX) --stages;
    // the M=128 MMA over-reads (128 - BMR) rows past the last dY chunk: they must fall inside the patch tiles
    if (fixed + stages * a_bytes > (size_t)CR_SMEM_MAX || (size_t)(128 - q.BMR) * 128 > 2 * b_bytes) return DK_ERR_UNSUPPORTED;
    q.a_stages = stages;
    *smem = fixed + stages * a_bytes;
    *grid = q.rows_total < sm_count() ? q.rows_total : sm_count();
    return DK_OK;
}

size_t conv_rows_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    (void)N; (void)H; (void)W; (void)s; (void)p;
    if (C * kh * kw > 128 || F > 128) return 0;
    return (size_t)sm_count() * F * C * kh * kw * sizeof(float);
}

int conv_rows_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                    int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    ConvRowsParams q = {};
    q.x = x;
    if (!cr_common(q, N, C, H, W, F, kh, kw, s, p) || !aligned16(dy)) return DK_ERR_UNSUPPORTED;
    size_t smem;
    int grid;
    bool span;
    if (cr_wgrad_plan(q, &smem, &grid, &span) != DK_OK) return DK_ERR_UNSUPPORTED;
    const size_t need = (size_t)grid * F * q.Kf * sizeof(float);
    if (ws == nullptr || ws_bytes < need) return DK_ERR_UNSUPPORTED;
    q.x = x; q.w = w; q.bias = nullptr; q.out = reinterpret_cast<float *>(ws);
    CUtensorMap tm;
    {   // dY as (OW, OH, N*F): box (32 pixels, 1 row, BMR filters) -> K-major chunk [BMR][32] with the 128B swizzle;
        // pixels past OW are zero-filled by the TMA unit
        cuuint64_t dims[3] = {(cuuint64_t)q.OW, (cuuint64_t)q.OH, (cuuint64_t)N * F};
        cuuint64_t strides[2] = {(cuuint64_t)q.OW * 4, (cuuint64_t)q.OH * q.OW * 4};
        cuuint32_t box[3] = {32, 1, (cuuint32_t)q.BMR};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = g_encode3(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(dy), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("conv wgrad: cuTensorMapEncodeTiled(dY) failed (%d)", (int)r);
            return DK_ERR_CUDA;
        }
    }
    CR_DISPATCH(conv_rows_wgrad_kernel, q, <<<grid, CR_THREADS, smem, st>>>(tm, q));
    DK_LAUNCH_CHECK();
    splitk_reduce_launch(q.out, w, dw, l2, (int64_t)F * q.Kf, grid, st);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
