// conv_tma.cu -- stride-1 k x k convolutions (ConvLayer: layers/convolution.py:58-126, layers/im2col.pyx:16-36,209-234) as
// TMA-fed tcgen05 implicit GEMMs directly on the NCHW tensors: the kernels for the one tensor-bound shape of the
// baseline (cfg2: 3x3, 64 -> 64 at 56x56) and the MNIST stack's 3x3 layers.
//
// The patch matrix of the reference never exists.  A 4-D tensor map (W, H, C, N) with a (32, 1, rows, 1) box delivers
// 32 consecutive pixels of ONE image row for 32 (or up to 128) channels as a ready-made swizzled operand tile, and a
// filter tap (i, j) is nothing but a coordinate shift (w + j - p, h + i - p): TMA's out-of-bounds zero fill IS the
// padding -- no index arithmetic, no gather warps, no padded copies.
//
//   forward / dgrad (conv_s1_kernel): tile = 4 output rows x 32 columns of one image (M = 128 pixels: TMEM lane =
//       pixel, so the epilogue stores 128-byte runs of NCHW), N = output channels (<= 256), K loop over (channel block
//       of 32, tap column j): ONE pipeline stage = the kh + 3 input-row boxes that serve all kh taps of that column
//       (tap i = boxes i..i+3 of the stage: an MN-major operand is a sequence of 4 KB 32-pixel blocks, so the A
//       descriptor simply starts i blocks later) -- input traffic from L2 is (kh+3)/4 * kw instead of kh * kw times
//       the tensor.  The permuted filters stay RESIDENT in shared memory for the whole persistent CTA when they fit
//       (147 KB for 64 x 64 x 3 x 3), else they stream next to the input boxes.
//       dgrad is the same kernel on dY with the flipped / transposed filters and padding k - 1 - p.
//   wgrad (conv_s1_wgrad_kernel): dW[f][c][i][j] = sum over pixels of dY[f][px] * X[c][px shifted by the tap]: M = F,
//       N = C, K = pixels (both operands K-major: 32 pixels of an image row are contiguous).  A CTA owns one tap column
//       j and a range of (image, 32-column strip) pairs and walks down the rows: each step loads ONE dY row box and
//       ONE new X row box; the X boxes of the last kh steps are the kh row taps (accumulator i in TMEM columns
//       i*stride), so every byte fetched feeds kh MMAs.  Partials per (split, j) go to the workspace, the existing
//       deterministic split-K reduce adds them (+ l2 * W).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int CT_THREADS = 192;       // wgrad: producer, MMA issuer, 4 epilogue warps
constexpr int CS_THREADS = 320;       // forward / dgrad: producer, MMA issuer, 8 epilogue warps
constexpr int CT_BOX = 4096;           // one (32 pixels x 32 channels) fp32 box
constexpr int CT_MAX_STAGES = 8;
constexpr int CT_SMEM_MAX = 227 * 1024 - 2048;
constexpr uint32_t CT_MN_LBO = 4096, CT_MN_SBO = 512, CT_MN_KSTEP = 1024;

int g_ct_wgrad2 = 1;         // 0: wgrad always through the column-shifted global copies
int g_ct_kc16 = 0;           // 0: stages always hold 32 channels
int g_conv_tma_enabled = 1;  // dk_tc_debug_set(17, 0) switches these kernels off (the gather variants take over)

struct CsParams {
    int N, Cin, H, W, Nout, OH, OW, kh, kw, ph, pw;
    int bnF;            // output channels rounded up to 32: one tap column's accumulator width
    int cblocks;        // input-channel blocks of 32
    int nb, nr;         // a tile = nr output rows x nb 32-column blocks (nb * nr = 4: M = 128 pixels)
    int rgroups, num_tiles;
    int stages, nbox;   // nbox = (nr + kh - 1) * nb input boxes per stage (one stage = kc channels)
    int kc;             // channels per stage: 32, or 16 (twice as many, half-size stages: finer overlap of loads and MMAs)
    uint32_t box_bytes; // kc * 128
    int w_resident;
    int jchunk;         // tap columns per MMA (N = jchunk * bnF <= 256)
    int nacc;           // accumulator sets in TMEM (2 = the epilogue overlaps the next tile)
    uint32_t b_bytes;   // one filter tile: bnF x 32 floats
    uint32_t stage_bytes, wres_bytes, xch_bytes;
    uint32_t tmem_cols, acc_stride;
    float *out;
    const float *bias;
};

// Forward-form kernel.  Only ALIGNED boxes can be fetched (a TMA box whose innermost coordinate is not a multiple of
// 16 bytes traps on this hardware: profiles/r01r_tma_alignment_probe.log), so the +-1 pixel shifts of the filter's
// columns cannot be loads.  They are not needed: with Z_j[f][r][c] = sum_{i, ch} W[f][ch][i][j] * X[ch][r + i - p][c]
// (vertical taps only -- whole-row shifts, always aligned) the convolution is Y[f][r][c] = sum_j Z_j[f][r][c + j - p], a
// shift of the RESULT by j - p columns.  The kw partial results Z_j are kw accumulators side by side in TMEM (one
// MMA of N = kw * F columns per (channel block, row tap) computes all of them from ONE read of the input boxes), and
// TMEM lane = pixel, so the epilogue adds them up with warp shuffles (+ a shared-memory hand-over of the edge lanes
// between the 32-column blocks of a row).  Columns outside the image are zero in X, hence in Z: padding again costs
// nothing.
__global__ void __launch_bounds__(CS_THREADS, 1)
conv_s1_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const CsParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t wres = smem_base;                       // resident filter tiles [cblock][i][j][bnF x 128 B]
    const uint32_t stage0 = smem_base + p.wres_bytes;
    const uint32_t xch_base = stage0 + (uint32_t)p.stages * p.stage_bytes;   // edge-lane hand-over (epilogue)
    const uint32_t bar_base = xch_base + p.xch_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (CT_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * CT_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * CT_MAX_STAGES + 2 + a); };
    const uint32_t wfull_bar = bar_base + 8u * (2 * CT_MAX_STAGES + 4);
    const uint32_t tmem_slot = bar_base + 8u * (2 * CT_MAX_STAGES + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        mbar_init(wfull_bar, 1);
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

    const int taps = p.kh * p.kw;
    const int rows_in = p.nr + p.kh - 1;

    if (warp == 0) {
        // ================================ TMA producer ================================
        // lane 0 owns the barriers; the boxes of a stage are issued by as many lanes as there are boxes (a single thread
        // issuing them one after the other was the bottleneck of the first version)
        if (p.w_resident && lane == 0) {
            mbar_expect_tx(wfull_bar, (uint32_t)(taps * p.cblocks) * p.b_bytes);
            for (int cb = 0; cb < p.cblocks; ++cb)
                for (int t = 0; t < taps; ++t)
                    tma_load_3d(wres + (uint32_t)(cb * taps + t) * p.b_bytes, &tmW, wfull_bar, cb * 32, 0, t);
        }
        int s = 0;
        uint32_t ph = 0;
        const uint32_t tx = (uint32_t)p.nbox * p.box_bytes + (p.w_resident ? 0u : (uint32_t)taps * p.b_bytes);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n = tile / p.rgroups, r0 = (tile - n * p.rgroups) * p.nr;
            for (int ch = 0; ch < p.Cin; ch += p.kc) {  // one stage per kc channels
                const uint32_t sA = stage0 + (uint32_t)s * p.stage_bytes, fb = full_bar(s);
                if (lane == 0) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(fb, tx);
                }
                __syncwarp();
                for (int b = lane; b < p.nbox; b += 32) {
                    const int rr = b / p.nb, cb = b - rr * p.nb;
                    tma_load_4d(sA + (uint32_t)b * p.box_bytes, &tmX, fb, 32 * cb, r0 + rr - p.ph, ch, n);
                }
                if (!p.w_resident) {  // (kc == 32 here)
                    const uint32_t sB = sA + (uint32_t)p.nbox * p.box_bytes;
                    for (int t = lane; t < taps; t += 32)
                        tma_load_3d(sB + (uint32_t)t * p.b_bytes, &tmW, fb, ch, 0, t);
                }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // whole warp, uniform control flow, one elected lane issues; descriptors as constant upper / incremented lower words
        // (an `if (lane == 0)` region costs an ELECT / R2UR.BROADCAST waterfall per tcgen05.mma: ~100 cycles of issue latency each)
        {
            if (p.w_resident) {
                mbar_wait(wfull_bar, 0);
                tc_fence_after();
            }
            const uint32_t a_hi = smem_desc_hi(CT_MN_SBO, LAYOUT_SW128_BASE32B), b_hi = smem_desc_hi(1024u, LAYOUT_SW128);
            int s = 0, local = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
                const int acc = p.nacc == 2 ? (local & 1) : 0;
                const uint32_t aph = p.nacc == 2 ? (((uint32_t)(local >> 1)) & 1u) : ((uint32_t)local & 1u);
                mbar_wait(tempty_bar(acc), aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * p.acc_stride;
                for (int ch = 0; ch < p.Cin; ch += p.kc) {
                    int nks = p.kc / 8;
                    const int rem = p.Cin - ch;
                    if (rem < p.kc) nks = (rem + 7) / 8;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sA = stage0 + (uint32_t)s * p.stage_bytes;
                    // the filter tiles keep 32-channel rows (128 B, K-major): a 16-channel stage starts 64 B into them
                    const uint32_t wt = (p.w_resident ? wres + (uint32_t)((ch >> 5) * taps) * p.b_bytes
                                                      : sA + (uint32_t)p.nbox * p.box_bytes) + (uint32_t)(ch & 31) * 4u;
#pragma unroll 1
                    for (int i = 0; i < p.kh; ++i) {
                        const uint32_t a_lo = smem_desc_lo(sA + (uint32_t)(i * p.nb) * p.box_bytes, p.box_bytes);
#pragma unroll 1
                        for (int j0 = 0; j0 < p.kw; j0 += p.jchunk) {
                            const int nj = p.kw - j0 < p.jchunk ? p.kw - j0 : p.jchunk;
                            const uint32_t idesc = idesc_tf32(128, nj * p.bnF, 1, 0);
                            const uint32_t b_lo = smem_desc_lo(wt + (uint32_t)(i * p.kw + j0) * p.b_bytes, 16u);
                            const uint32_t dd = d_tmem + (uint32_t)(j0 * p.bnF);
                            const uint32_t acc0 = (ch > 0 || i > 0) ? 1u : 0u;
                            if (nks == 4) {
                                if (elect_one()) mma_tf32_k4(dd, a_lo, a_hi, b_lo, b_hi, CT_MN_KSTEP >> 4, 2u, idesc, acc0);
                            } else {
#pragma unroll 1
                                for (int ks = 0; ks < nks; ++ks)
                                    if (elect_one())
                                        mma_tf32_lohi(dd, a_lo + (uint32_t)ks * (CT_MN_KSTEP >> 4), a_hi, b_lo + 2u * (uint32_t)ks, b_hi, idesc,
                                                      (acc0 || ks > 0) ? 1u : 0u);
                            }
                        }
                    }
                    if (elect_one()) mma_commit(empty_bar(s));
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) mma_commit(tfull_bar(acc));
            }
            __syncwarp();
        }
    } else {
        // ================================ epilogue (warps 2..9) ========================
        // Eight warps: warp w may only touch TMEM lanes 32*(w%4).. (its quadrant = one 32-pixel block of the tile), so two
        // warps share a quadrant and split the output channels in chunks of 32 (half 0: channels 0-31, 64-95, ...).
        const int q = warp & 3;                      // block (row q / nb, column block q % nb) of the tile
        const int half = (warp - 2) >> 2;
        const int trow = q / p.nb, cblk = q - trow * p.nb;
        const int col = 32 * cblk + lane;
        // edge-lane hand-over between the column blocks of a row: [half][quadrant][tap column][edge lane 0..3][32 channels]
        float *xch = reinterpret_cast<float *>(smem_raw + (xch_base - smem_u32(smem_raw))) + half * (4 * p.kw * 4 * 32);
        const long long plane = (long long)p.OH * p.OW;
        int local = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local) {
            const int acc = p.nacc == 2 ? (local & 1) : 0;
            const uint32_t aph = p.nacc == 2 ? (((uint32_t)(local >> 1)) & 1u) : ((uint32_t)local & 1u);
            const int n = tile / p.rgroups, row = (tile - n * p.rgroups) * p.nr + trow;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const bool ok = row < p.OH && col < p.OW;
            float *o = p.out + (long long)n * p.Nout * plane + (long long)row * p.OW + col;
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)acc * p.acc_stride;
            for (int c = 32 * half; c < p.bnF; c += 64) {
                float y[32];
                {   // the unshifted tap column
                    uint32_t v[32];
                    tmem_ld32(t_row + (uint32_t)(p.pw * p.bnF + c), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) y[jj] = __uint_as_float(v[jj]);
                }
                for (int j = 0; j < p.kw; ++j) {
                    const int d = j - p.pw;  // Y[col] += Z_j[col + d]
                    if (d == 0) continue;
                    uint32_t v[32];
                    tmem_ld32(t_row + (uint32_t)(j * p.bnF + c), v);
                    tmem_ld_wait();
                    if (p.nb > 1) {
                        // d > 0: the block to the LEFT misses my lanes 0 .. d-1;  d < 0: the block to the RIGHT my lanes 32+d .. 31
                        const int e = d > 0 ? lane : lane - (32 + d);
                        if (e >= 0 && e < (d > 0 ? d : -d)) {
                            float4 *dst = reinterpret_cast<float4 *>(xch + ((q * p.kw + j) * 4 + e) * 32);
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4)
                                dst[jj >> 2] = make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]),
                                                           __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
                        }
                    }
                    const int src = lane + d;
                    const float m = (src >= 0 && src < 32) ? 1.0f : 0.0f;  // lanes whose source column is in another block
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        y[jj] = fmaf(__shfl_sync(0xffffffffu, __uint_as_float(v[jj]), src & 31), m, y[jj]);
                }
                if (p.nb > 1) {
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                    for (int j = 0; j < p.kw; ++j) {
                        const int d = j - p.pw, src = lane + d;
                        if (d == 0) continue;
                        // the neighbour published exactly the lanes this block misses: index = lane (d < 0) or src - 32 (d > 0)
                        const bool take = src < 0 ? cblk > 0 : (src >= 32 && cblk + 1 < p.nb);
                        if (take) {
                            const int nq = src < 0 ? q - 1 : q + 1;
                            const float4 *nsrc = reinterpret_cast<const float4 *>(xch + ((nq * p.kw + j) * 4 + (src < 0 ? lane : src - 32)) * 32);
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4) {
                                const float4 t = nsrc[jj >> 2];
                                y[jj] += t.x; y[jj + 1] += t.y; y[jj + 2] += t.z; y[jj + 3] += t.w;
                            }
                        }
                    }
                }
                if (ok) {
                    float *pp = o + (long long)c * plane;
                    if (c + 32 <= p.Nout && p.bias == nullptr) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj, pp += plane) *pp = y[jj];
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj, pp += plane)
                            if (c + jj < p.Nout) *pp = y[jj] + (p.bias ? __ldg(p.bias + c + jj) : 0.0f);
                    }
                }
                if (p.nb > 1) asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");  // the hand-over buffer is free again
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---- wgrad ------------------------------------------------------------------------------------------------------------
struct CwParams {
    int N, C, H, W, F, OH, OW, kh, kw, p;
    int bnC;            // MMA N = round_up(C, 32)
    int a_rows;         // rows of the dY box (F rounded up to 8, <= 128)
    int csegs, strips;  // 32-column strips per image row, N * csegs strips in total
    int splits, strips_per_split;
    int stages;
    uint32_t a_bytes, b_bytes, stage_bytes;
    uint32_t tmem_cols, acc_stride;
    float *partial;     // [splits][F][C][kh][kw]
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_s1_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ CUtensorMap tmXS, const CwParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * p.stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (CT_MAX_STAGES + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * CT_MAX_STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * CT_MAX_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmXS);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
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

    const int j = blockIdx.x % p.kw, split = blockIdx.x / p.kw;
    const int sb = split * p.strips_per_split;
    int se = sb + p.strips_per_split;
    if (se > p.strips) se = p.strips;
    const int steps = p.OH + p.kh - 1;  // per strip: kh - 1 rows of run-in, then one output row per step

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int strip = sb; strip < se; ++strip) {
                const int n = strip / p.csegs, c0 = (strip - n * p.csegs) * 32;
                for (int t = 0; t < steps; ++t) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sA = smem_base + (uint32_t)s * p.stage_bytes, sB = sA + p.a_bytes, fb = full_bar(s);
                    const int oh = t - (p.kh - 1);
                    mbar_expect_tx(fb, p.b_bytes + (oh >= 0 ? (uint32_t)p.a_rows * 128u : 0u));
                    // X row t - p (zero rows above / below) of the copy shifted by j - p columns: aligned boxes only (a box
                    // starting at an odd column traps); the unshifted tap column reads X itself
                    if (j == p.p) tma_load_4d(sB, &tmX, fb, c0, t - p.p, 0, n);
                    else tma_load_4d(sB, &tmXS, fb, c0, t - p.p, 0, (j < p.p ? j : j - 1) * p.N + n);
                    if (oh >= 0) tma_load_4d(sA, &tmDY, fb, c0, oh, 0, n);
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // (whole warp, uniform control flow, one elected lane issues: see conv_s1_kernel)
        {
            const uint32_t idesc = idesc_tf32(128, p.bnC, 0, 0);
            const uint32_t d_hi = smem_desc_hi(1024u, LAYOUT_SW128);
            int s = 0;
            uint32_t ph = 0;
            bool first = true;
            for (int strip = sb; strip < se; ++strip) {
                for (int t = 0; t < steps; ++t) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    if (t >= p.kh - 1) {
                        const uint32_t sA = smem_base + (uint32_t)s * p.stage_bytes;
#pragma unroll 1
                        for (int i = 0; i < p.kh; ++i) {
                            // X row oh + i - p was loaded kh - 1 - i steps ago
                            int sx = s - (p.kh - 1 - i);
                            if (sx < 0) sx += p.stages;
                            const uint32_t sB = smem_base + (uint32_t)sx * p.stage_bytes + p.a_bytes;
                            const uint32_t d_tmem = tmem_base + (uint32_t)i * p.acc_stride;
                            if (elect_one())
                                mma_tf32_k4(d_tmem, smem_desc_lo(sA, 16u), d_hi, smem_desc_lo(sB, 16u), d_hi, 2u, 2u, idesc, first ? 0u : 1u);
                        }
                        first = false;
                        // the oldest X row of this step is not needed again
                        int so = s - (p.kh - 1);
                        if (so < 0) so += p.stages;
                        if (elect_one()) mma_commit(empty_bar(so));
                        if (t == steps - 1) {  // end of the strip: the remaining run-out stages
                            for (int d = p.kh - 2; d >= 0; --d) {
                                int sr = s - d;
                                if (sr < 0) sr += p.stages;
                                if (elect_one()) mma_commit(empty_bar(sr));
                            }
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
            if (elect_one()) mma_commit(tfull_bar);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        if (sb < se) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
        const int f = 32 * q + lane;
        const int taps = p.kh * p.kw;
        float *o = p.partial + ((long long)split * p.F + f) * p.C * taps + j;
        for (int i = 0; i < p.kh; ++i) {
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)i * p.acc_stride;
            for (int c = 0; c < p.bnC; c += 32) {
                uint32_t v[32];
                if (sb < se) {
                    tmem_ld32(t_row + (uint32_t)c, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) v[jj] = 0u;
                }
                if (f < p.F) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        if (c + jj < p.C) o[(long long)(c + jj) * taps + i * p.kw] = __uint_as_float(v[jj]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---- wgrad, F <= 64: all kh * kw taps from ONE dY row box and ONE X row box per step ------------------------------------
// For X row r the kh row taps pair it with the dY rows r + p - i.  Two consecutive dY rows of 64 filters stacked are a full
// M = 128 operand: [dY(oh - 1); dY(oh)] against X row r yields tap i + 1 in TMEM lanes 0-63 and tap i in lanes 64-127, so
// ceil(kh / 2) MMAs per step cover every row tap.  The kw column taps are kw copies of the X row shifted by j - p pixels,
// stacked along N ([j][c] rows, N = kw * C): shifting cannot be a load (unaligned boxes trap), so four warps build the
// shifted K-major tiles in shared memory from one aligned (40 pixel x C channel) raw box -- 16-byte loads, register
// renaming, 16-byte swizzled stores.  A step therefore fetches 8 KB of dY + 10 KB of X for kh * kw taps (the column-shifted
// global copies of the fallback kernel below cost 3 x the traffic plus a pre-pass).  dY rows live in a ring (consecutive
// loads in consecutive 8 KB slots; slot 0 is mirrored behind the last slot so that a pair never wraps).
struct Cw2Params {
    int N, C, H, W, F, OH, OW, kh, kw, p;
    int bnC, a_rows, csegs, chunks, rc;   // rc = X rows per work unit, chunks = row chunks per strip
    int units, units_per_cta;
    int ring;                            // dY ring slots (+1 mirror)
    int nb;                              // shifted-tile buffers (2 .. 4): how far the shifter warps may run ahead of the MMAs
    int raw_stages;                      // raw X row boxes in flight
    int npairs;
    int dbg;                             // timing experiments only (results are garbage): 1 no shifting, 2 no MMAs, 4 no dY loads, 8 no X loads
    uint32_t tmem_cols, acc_stride;      // acc_stride = kw * bnC columns per row-tap pair
    float *partial;                      // [CTAs][tap][C][F]
};

// 16 pixels of one channel row, shifted by D columns, as four 16-byte chunks of a swizzled K-major tile row
template <int D>
__device__ __forceinline__ void cw2_store_shifted(const float (&v)[24], uint8_t *row, uint32_t chunk0, uint32_t sw) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        *reinterpret_cast<float4 *>(row + (((chunk0 + kk) ^ sw) << 4)) =
            make_float4(v[4 + 4 * kk + D], v[5 + 4 * kk + D], v[6 + 4 * kk + D], v[7 + 4 * kk + D]);
}

// dbg & 256: time spent in a wait (clock64 either side); otherwise just the statement
#define CW2_T(stmt, acc)                       \
    do {                                       \
        if (p.dbg & 256) {                     \
            const long long t0_ = clock64();   \
            stmt;                              \
            acc += clock64() - t0_;            \
        } else {                               \
            stmt;                              \
        }                                      \
    } while (0)
constexpr int CW2_RAW_W = 40;  // raw X box: 4 halo pixels left, 32, 4 right
constexpr uint32_t CW2_DY_SLOT = 8192;
constexpr int CW2_RAW_STAGES = 6;   // most raw X row boxes in flight (p.raw_stages)
constexpr int CW2_NB_MAX = 4;       // most shifted-tile buffers (p.nb)

__global__ void __launch_bounds__(CT_THREADS, 1)
conv_s1_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmXR, const Cw2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring_base = smem_base;                                     // (ring + 1) x 8 KB
    const uint32_t b_tile = (uint32_t)(p.kw * p.bnC) * 128u;                  // kw shifted tiles of bnC rows
    const uint32_t b_base = ring_base + (uint32_t)(p.ring + 1) * CW2_DY_SLOT;  // p.nb x b_tile
    const uint32_t raw_bytes = (uint32_t)p.bnC * CW2_RAW_W * 4u;
    const uint32_t raw_base = b_base + (uint32_t)p.nb * b_tile;               // p.raw_stages x raw box
    const uint32_t bar_base = (raw_base + (uint32_t)p.raw_stages * raw_bytes + 15u) & ~15u;
    auto dyfull = [&](int s) { return bar_base + 8u * s; };
    auto dyempty = [&](int s) { return bar_base + 8u * (16 + s); };
    auto rawfull = [&](int s) { return bar_base + 8u * (32 + s); };
    auto rawempty = [&](int s) { return bar_base + 8u * (40 + s); };
    auto bfull = [&](int s) { return bar_base + 8u * (48 + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (52 + s); };
    const uint32_t tfull_bar = bar_base + 8u * 56;
    const uint32_t tmem_slot = bar_base + 8u * 57;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmXR);
        for (int s = 0; s < p.ring; ++s) {
            mbar_init(dyfull(s), 1);
            mbar_init(dyempty(s), 1);
        }
        for (int s = 0; s < p.raw_stages; ++s) {
            mbar_init(rawfull(s), 1);
            mbar_init(rawempty(s), 4);
        }
        for (int s = 0; s < p.nb; ++s) {
            mbar_init(bfull(s), 4);
            mbar_init(bempty(s), 1);
        }
        mbar_init(tfull_bar, 1);
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

    const int u_lo = blockIdx.x * p.units_per_cta;
    int u_hi = u_lo + p.units_per_cta;
    if (u_hi > p.units) u_hi = p.units;
    if (p.dbg & 64) u_hi = u_lo;
    const int run_in = p.kh - 1;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            long long tw0 = 0, tw1 = 0;
            const long long tstart = clock64();
            int L = 0;        // dY loads so far
            int slot = 0, rs = 0;         // dY ring slot, raw X stage
            uint32_t dph = 1u, rph = 1u;  // phases of the "empty" waits (first lap passes)
            for (int u = u_lo; u < u_hi; ++u) {
                const int strip = u / p.chunks, chunk = u - strip * p.chunks;
                const int n = strip / p.csegs, c0 = (strip - n * p.csegs) * 32;
                const int r_lo = chunk * p.rc;
                int r_hi = r_lo + p.rc;
                if (r_hi > p.H) r_hi = p.H;
                const int oh0 = r_lo + p.p - run_in;  // first dY row of the unit (rows outside the tensor arrive as zeros)
                const int nload = (r_hi - r_lo) + run_in;
                for (int t = 0; t < nload; ++t, ++L) {
                    CW2_T(mbar_wait(dyempty(slot), dph), tw0);
                    const uint32_t fb = dyfull(slot), bytes = (uint32_t)p.a_rows * 128u;
                    if (p.dbg & 4) {
                        mbar_arrive(fb);
                    } else {
                        mbar_expect_tx(fb, slot == 0 ? 2u * bytes : bytes);
                        tma_load_4d(ring_base + (uint32_t)slot * CW2_DY_SLOT, &tmDY, fb, c0, oh0 + t, 0, n);
                        if (slot == 0) tma_load_4d(ring_base + (uint32_t)p.ring * CW2_DY_SLOT, &tmDY, fb, c0, oh0 + t, 0, n);
                    }
                    if (t >= run_in) {  // the X row of this step
                        CW2_T(mbar_wait(rawempty(rs), rph), tw1);
                        if (p.dbg & 8) {
                            mbar_arrive(rawfull(rs));
                        } else {
                            mbar_expect_tx(rawfull(rs), raw_bytes);
                            tma_load_4d(raw_base + (uint32_t)rs * raw_bytes, &tmXR, rawfull(rs), c0 - 4, r_lo + (t - run_in), 0, n);
                        }
                        if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
                    }
                    if (++slot == p.ring) { slot = 0; dph ^= 1u; }
                }
            }
            if ((p.dbg & 256) && (blockIdx.x == 0 || blockIdx.x == 73))
                printf("cw2 cta %d producer: total %lld  wait dyempty %lld  wait rawempty %lld  (loads %d)\n", blockIdx.x,
                       clock64() - tstart, tw0, tw1, L);
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // The whole warp walks the loop (uniform control flow, every lane polls the barriers) and one elected lane issues: the
        // issue stream IS the step-time floor, so ring slots and phases are carried incrementally (no divisions) and the
        // descriptors are split into constant upper and incremented lower words (mma_tf32_k4).
        {
            const uint32_t idesc = idesc_tf32(128, (p.dbg & 512) ? 16 : p.kw * p.bnC, 0, 0);
            const uint32_t d_hi = smem_desc_hi(1024u, LAYOUT_SW128);
            int slot = 0, bs = 0, xs = 0;      // dY ring slot of the current load, shifted-tile buffer of the current step
            uint32_t dph = 0, bph = 0;         // their phases
            long long tw0 = 0, tw1 = 0;
            const long long tstart = clock64();
            uint32_t started = 0;  // accumulators written?
            for (int u = u_lo; u < u_hi; ++u) {
                const int chunk = u % p.chunks;
                const int r_lo = chunk * p.rc;
                int r_hi = r_lo + p.rc;
                if (r_hi > p.H) r_hi = p.H;
                const int nload = (r_hi - r_lo) + run_in;
                for (int t = 0; t < nload; ++t) {
                    // every dY load is waited for exactly once, in order (the run-in rows carry no step of their own)
                    CW2_T(mbar_wait(dyfull(slot), dph), tw0);
                    if (t >= run_in) {
                        CW2_T(mbar_wait(bfull(bs), bph), tw1);
                        if (!(p.dbg & 2048)) tc_fence_after();
                        const uint32_t b_lo = smem_desc_lo(b_base + (uint32_t)bs * b_tile, 16u);
                        int sold = slot - run_in;
                        if (sold < 0) sold += p.ring;
                        // row tap i pairs this X row with the dY row loaded i loads ago
#pragma unroll 1
                        for (int k = 0; k < ((p.dbg & 2) ? 0 : p.npairs); ++k) {
                            const int i_lo = 2 * k;                      // tap in lanes 64-127 (or 0-63 when it has no partner)
                            int stop = slot - (i_lo + 1 < p.kh ? i_lo + 1 : i_lo);
                            if (stop < 0) stop += p.ring;
                            const uint32_t a_lo = smem_desc_lo(ring_base + (uint32_t)stop * CW2_DY_SLOT, 16u);
                            if (elect_one())
                                mma_tf32_k4(tmem_base + (uint32_t)k * p.acc_stride, a_lo, d_hi, b_lo, d_hi, 2u, 2u, idesc, started);
                        }
                        started = 1u;
                        if (elect_one()) {
                            if (p.dbg & 8192) {  // (only without MMAs) plain arrives instead of commits
                                mbar_arrive(bempty(bs));
                                mbar_arrive(dyempty(sold));
                            } else {
                                mma_commit(bempty(bs));
                                mma_commit(dyempty(sold));  // the oldest dY row of this step is done
                            }
                        }
                        if (t == nload - 1)
                            for (int d = run_in - 1; d >= 0; --d) {
                                int sd = slot - d;
                                if (sd < 0) sd += p.ring;
                                if (elect_one()) mma_commit(dyempty(sd));
                            }
                        ++xs;
                        if (++bs == p.nb) { bs = 0; bph ^= 1u; }
                    }
                    if (++slot == p.ring) { slot = 0; dph ^= 1u; }
                }
            }
            if (elect_one()) mma_commit(tfull_bar);
            if ((p.dbg & 256) && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 73)) {
                const long long tissue = clock64() - tstart;
                mbar_wait(tfull_bar, 0);
                printf("cw2 cta %d mma: issue loop %lld (until done %lld)  wait dyfull %lld  wait bfull %lld  (steps %d)\n", blockIdx.x,
                       tissue, clock64() - tstart, tw0, tw1, xs);
            }
            __syncwarp();
        }
    } else {
        // ================================ shifter (warps 2..5), then epilogue ==========
        const int t128 = threadIdx.x - 64;          // 0..127
        const int c = t128 >> 1, hh = t128 & 1;      // channel row, 16-pixel half
        int total_steps = 0;
        for (int u = u_lo; u < u_hi; ++u) {
            const int r_lo = (u % p.chunks) * p.rc;
            int r_hi = r_lo + p.rc;
            if (r_hi > p.H) r_hi = p.H;
            total_steps += r_hi - r_lo;
        }
        long long tw0 = 0, tw1 = 0;
        const long long tstart = clock64();
        int rs = 0, bs = 0;
        uint32_t rph = 0, bph = 1u;
        for (int xs = 0; xs < total_steps; ++xs) {
            CW2_T(mbar_wait(rawfull(rs), rph), tw0);
            CW2_T(mbar_wait(bempty(bs), bph), tw1);
            if (c < p.bnC && !(p.dbg & 1)) {
                // raw[c][16*hh .. 16*hh + 23] covers the 16 pixels of this half shifted by -4 .. +4
                const float4 *src = reinterpret_cast<const float4 *>(smem_raw + (raw_base - smem_u32(smem_raw)) + (uint32_t)rs * raw_bytes +
                                                                     (uint32_t)c * (CW2_RAW_W * 4) + (uint32_t)hh * 64u);
                float v[24];
#pragma unroll
                for (int k4 = 0; k4 < 6; ++k4) {
                    const float4 t4 = src[k4];
                    v[4 * k4] = t4.x; v[4 * k4 + 1] = t4.y; v[4 * k4 + 2] = t4.z; v[4 * k4 + 3] = t4.w;
                }
                uint8_t *bt = smem_raw + (b_base - smem_u32(smem_raw)) + (uint32_t)bs * b_tile + (uint32_t)c * 128u;
                for (int j = 0; j < p.kw; ++j) {
                    uint8_t *row = bt + (uint32_t)(j * p.bnC) * 128u;
                    const uint32_t sw = (uint32_t)(c & 7), ch0 = 4u * (uint32_t)hh;
                    switch (j - p.p) {  // tile j holds X[col + d]: raw index 4 + col + d (uniform branch)
                        case -4: cw2_store_shifted<-4>(v, row, ch0, sw); break;
                        case -3: cw2_store_shifted<-3>(v, row, ch0, sw); break;
                        case -2: cw2_store_shifted<-2>(v, row, ch0, sw); break;
                        case -1: cw2_store_shifted<-1>(v, row, ch0, sw); break;
                        case 0: cw2_store_shifted<0>(v, row, ch0, sw); break;
                        case 1: cw2_store_shifted<1>(v, row, ch0, sw); break;
                        case 2: cw2_store_shifted<2>(v, row, ch0, sw); break;
                        case 3: cw2_store_shifted<3>(v, row, ch0, sw); break;
                        default: cw2_store_shifted<4>(v, row, ch0, sw); break;
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bfull(bs));
                mbar_arrive(rawempty(rs));
            }
            if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
            if (++bs == p.nb) { bs = 0; bph ^= 1u; }
        }
        // epilogue: accumulator k, lanes 0-63 = tap 2k+1 (or the unpaired tap), lanes 64-127 = tap 2k
        const int q = warp & 3;
        const long long tloop = clock64() - tstart;
        if (u_lo < u_hi) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
        const long long ttail = clock64() - tstart;
        const int f = 32 * (q & 1) + lane;
        const int taps = p.kh * p.kw;
        // partial sums as [CTA][tap][c][f]: for every (tap, c) a warp stores 32 consecutive floats.  (The first version wrote
        // the final [f][c][tap] order directly: 4-byte stores 36 bytes apart, 49 K of them per CTA -- ncu's stall samples put
        // a fifth of the kernel into this epilogue.  cw2_reduce_kernel does the transposition while it adds the partials.)
        float *o = p.partial + (long long)blockIdx.x * taps * p.C * p.F + f;
        for (int k = 0; k < p.npairs; ++k) {
            const bool paired = 2 * k + 1 < p.kh;
            const int i = paired ? (q < 2 ? 2 * k + 1 : 2 * k) : 2 * k;
            const bool live = paired || q < 2;
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)k * p.acc_stride;
            for (int j = 0; j < p.kw; ++j) {
                for (int cc = 0; cc < p.bnC; cc += 32) {
                    uint32_t v[32];
                    if (u_lo < u_hi) {
                        tmem_ld32(t_row + (uint32_t)(j * p.bnC + cc), v);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) v[jj] = 0u;
                    }
                    if (live && f < p.F && !(p.dbg & 16)) {
                        float *ot = o + (long long)((i * p.kw + j) * p.C + cc) * p.F;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj)
                            if (cc + jj < p.C) ot[(long long)jj * p.F] = __uint_as_float(v[jj]);
                    }
                }
            }
        }
        if ((p.dbg & 256) && warp == 2 && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 73))
            printf("cw2 cta %d shifter: loop %lld  accumulators ready %lld  stores issued %lld  wait rawfull %lld  wait bempty %lld\n",
                   blockIdx.x, tloop, ttail, clock64() - tstart, tw0, tw1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dW[f][c][tap] = sum_z partial[z][tap][c][f] + l2 * W[f][c][tap]: reads coalesced along f, the Z partials spread over 32
// thread rows and combined in a fixed order (deterministic), one scattered write per output element
__global__ void __launch_bounds__(1024)
cw2_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ dw, float l2, int F, int C,
                  int taps, int Z) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long mn = (long long)taps * C * F;
    const long long i = (long long)blockIdx.x * 32 + tx;
    float s = 0.0f;
    if (i < mn) {
        int z = ty;
        for (; z + 96 < Z; z += 128) {
            const float a = partial[(long long)z * mn + i], b = partial[(long long)(z + 32) * mn + i];
            const float c2 = partial[(long long)(z + 64) * mn + i], d = partial[(long long)(z + 96) * mn + i];
            s += (a + b) + (c2 + d);
        }
        for (; z < Z; z += 32) s += partial[(long long)z * mn + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < mn) {
        float t = red[0][tx];
#pragma unroll
        for (int y = 1; y < 32; ++y) t += red[y][tx];
        const int f = (int)(i % F);
        const long long r = i / F;
        const int c = (int)(r % C), tap = (int)(r / C);
        const long long o = ((long long)f * C + c) * taps + tap;
        dw[o] = t + (l2 != 0.0f ? l2 * w[o] : 0.0f);
    }
}

// filters -> [tap][out channel][in channel padded to 32]:  mode 0 (forward) Wp[(i,j)][f][c] = W[f][c][i][j];
// mode 1 (dgrad) Wp[(i',j')][c][f] = W[f][c][kh-1-i'][kw-1-j']
__global__ void conv_tma_permute_kernel(const float *__restrict__ w, float *__restrict__ wp, int F, int C, int kh, int kw,
                                        int Nout, int Cp, int mode) {
    const int taps = kh * kw;
    const long long total = (long long)taps * Nout * Cp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % Cp);
        const long long r = idx / Cp;
        const int no = (int)(r % Nout), t = (int)(r / Nout);
        const int i = t / kw, jj = t - i * kw;
        float v = 0.0f;
        if (mode == 0) {
            if (ci < C) v = __ldg(w + (((long long)no * C + ci) * kh + i) * kw + jj);
        } else {
            if (ci < F) v = __ldg(w + (((long long)ci * C + no) * kh + (kh - 1 - i)) * kw + (kw - 1 - jj));
        }
        wp[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_ct_encode = nullptr;
static bool g_ct_ready = false;

int init_conv_tma() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return DK_OK;
    }
    g_ct_encode = reinterpret_cast<EncodeTiledFn>(fn);
    if (cudaFuncSetAttribute(conv_s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess ||
        cudaFuncSetAttribute(conv_s1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess ||
        cudaFuncSetAttribute(conv_s1_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess) {
        cudaGetLastError();
        return DK_OK;
    }
    if (const char *m = getenv("DK_CONV_TMA")) g_conv_tma_enabled = atoi(m);
    g_ct_ready = true;
    return DK_OK;
}

// fp32 tensor map of rank 3 or 4 with dense strides
static int ct_map(CUtensorMap *out, const float *ptr, int rank, const uint64_t *dims, const uint32_t *box, CUtensorMapSwizzle sw) {
    cuuint64_t d[4], strides[3];
    cuuint32_t b[4], es[4] = {1, 1, 1, 1};
    uint64_t pitch = 4;
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        pitch *= dims[i];
        if (i < rank - 1) strides[i] = pitch;
    }
    CUresult r = g_ct_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float *>(ptr), d, strides, b, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_tma: cuTensorMapEncodeTiled failed (%d), rank %d dims (%llu,%llu,%llu) box (%u,%u,%u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2]);
        return DK_ERR_CUDA;
    }
    return DK_OK;
}

static int ct_round_up(int v, int m) { return (v + m - 1) / m * m; }
static uint32_t ct_acc_stride(int bn) { return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u; }
static uint32_t ct_pow2_cols(uint32_t c) { uint32_t v = 32; while (v < c) v <<= 1; return v; }

// shapes the forward-form kernel takes: input [N, Cin, H, W] -> output [N, Nout, OH, OW], stride 1, pads (ph, pw)
static bool ct_fwd_ok(const float *x, int Cin, int H, int W, int Nout, int OW, int kh, int kw, int pw) {
    const int wmax = W > OW ? W : OW;
    return g_ct_ready && g_conv_tma_enabled && aligned16(x) && (W % 4) == 0 && wmax <= 128 && kh <= 5 && kw <= 7 && kh * kw > 1 &&
           Cin >= 8 && W >= 8 && pw <= 4 && kw - 1 - pw <= 4 && ct_round_up(Nout, 32) <= 256 && kw * ct_round_up(Nout, 32) <= 512;
}

size_t conv_tma_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    if (s != 1 || kh * kw <= 1) return 0;
    const int taps = kh * kw;
    const size_t big = (size_t)(F > C ? F : C);
    const size_t perm = (size_t)taps * big * (size_t)ct_round_up((int)big, 32) * 4 + 1024;
    // wgrad: the kw - 1 column-shifted copies of X, then the split-K partials (splits * kw <= SM count)
    const size_t shifted = (size_t)(kw - 1) * N * C * H * W * 4 + 1024;
    const size_t partial = (size_t)sm_count() * F * C * taps * 4 + 1024;
    return (perm > shifted + partial ? perm : shifted + partial) + 1024;
}

static int ct_run_fwd(const float *x, const float *w, const float *bias, float *y, int N, int Cin, int H, int W, int Nout,
                      int OH, int OW, int kh, int kw, int ph, int pw, int F, int C, int mode, void *ws, size_t ws_bytes,
                      cudaStream_t st) {
    CsParams q = {};
    q.N = N; q.Cin = Cin; q.H = H; q.W = W; q.Nout = Nout; q.OH = OH; q.OW = OW; q.kh = kh; q.kw = kw; q.ph = ph; q.pw = pw;
    q.bnF = ct_round_up(Nout, 32);
    const int Cp = ct_round_up(Cin, 32);
    q.cblocks = Cp / 32;
    const int wmax = W > OW ? W : OW;
    q.nb = wmax <= 32 ? 1 : wmax <= 64 ? 2 : 4;
    q.nr = 4 / q.nb;
    q.rgroups = (OH + q.nr - 1) / q.nr;
    q.num_tiles = N * q.rgroups;
    q.nbox = (q.nr + kh - 1) * q.nb;
    q.b_bytes = (uint32_t)q.bnF * 128u;
    q.jchunk = 256 / q.bnF;
    if (q.jchunk > kw) q.jchunk = kw;
    const uint32_t acc_cols = (uint32_t)(kw * q.bnF);
    q.nacc = 2 * acc_cols <= 512 ? 2 : 1;
    q.acc_stride = acc_cols;
    q.tmem_cols = ct_pow2_cols((uint32_t)q.nacc * acc_cols);
    q.xch_bytes = (uint32_t)kw * 4096u;
    const int taps = kh * kw;
    const size_t perm_bytes = (size_t)taps * Nout * Cp * sizeof(float);
    if (ws == nullptr || ws_bytes < perm_bytes + 256) return DK_ERR_UNSUPPORTED;
    float *wp = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ws) + 255u) & ~(uintptr_t)255u);
    // resident filters if they leave room for at least 2 stages of input boxes
    const uint32_t wres = (uint32_t)(taps * q.cblocks) * q.b_bytes;
    uint32_t a_stage = (uint32_t)q.nbox * CT_BOX;
    const int avail = CT_SMEM_MAX - 1024 - 256 - (int)q.xch_bytes;
    q.kc = 32;
    q.box_bytes = CT_BOX;
    if ((int64_t)wres + 2 * (int64_t)a_stage <= avail) {
        q.w_resident = 1;
        q.wres_bytes = wres;
        // little room next to the resident filters: half-size stages (16 channels) keep more loads in flight behind the
        // stage the tensor core is reading
        if ((avail - (int64_t)wres) / a_stage < 4 && g_ct_kc16 && Cin % 16 == 0) {
            q.kc = 16;
            q.box_bytes = CT_BOX / 2;
            a_stage /= 2;
        }
        q.stage_bytes = a_stage;
    } else {
        q.w_resident = 0;
        q.wres_bytes = 0;
        q.stage_bytes = a_stage + (uint32_t)taps * q.b_bytes;
    }
    int stg = (avail - (int)q.wres_bytes) / (int)q.stage_bytes;
    if (stg < 2) return DK_ERR_UNSUPPORTED;
    q.stages = stg > CT_MAX_STAGES ? CT_MAX_STAGES : stg;
    q.out = y;
    q.bias = bias;
    {
        const long long total = (long long)taps * Nout * Cp;
        const int grid = (int)(ceil_div(total, 256) < 4 * sm_count() ? ceil_div(total, 256) : 4 * sm_count());
        conv_tma_permute_kernel<<<grid, 256, 0, st>>>(w, wp, F, C, kh, kw, Nout, Cp, mode);
        DK_LAUNCH_CHECK();
    }
    CUtensorMap tx, tw;
    const uint64_t dx[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)Cin, (uint64_t)N};
    const uint32_t bx[4] = {32, 1, (uint32_t)q.kc, 1};
    int rc = ct_map(&tx, x, 4, dx, bx, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    const uint64_t dw[3] = {(uint64_t)Cp, (uint64_t)Nout, (uint64_t)taps};
    const uint32_t bw[3] = {32, (uint32_t)q.bnF, 1};
    rc = ct_map(&tw, wp, 3, dw, bw, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = (size_t)q.wres_bytes + (size_t)q.stages * q.stage_bytes + q.xch_bytes + 1024 + 8 * (2 * CT_MAX_STAGES + 8);
    const int grid = q.num_tiles < sm_count() ? q.num_tiles : sm_count();
    conv_s1_kernel<<<grid, CS_THREADS, smem, st>>>(tx, tw, q);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int conv_tma_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                 int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (s != 1) return DK_ERR_UNSUPPORTED;
    const int OH = H + 2 * p - kh + 1, OW = W + 2 * p - kw + 1;
    if (OH < 1 || OW < 1 || !ct_fwd_ok(x, C, H, W, F, OW, kh, kw, p)) return DK_ERR_UNSUPPORTED;
    return ct_run_fwd(x, w, bias, y, N, C, H, W, F, OH, OW, kh, kw, p, p, F, C, 0, ws, ws_bytes, st);
}

int conv_tma_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw, int s,
                   int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (s != 1) return DK_ERR_UNSUPPORTED;
    const int OH = H + 2 * p - kh + 1, OW = W + 2 * p - kw + 1;
    if (OH < 1 || OW < 1 || kh - 1 - p < 0 || kw - 1 - p < 0) return DK_ERR_UNSUPPORTED;
    if (!ct_fwd_ok(dy, F, OH, OW, C, W, kh, kw, kw - 1 - p)) return DK_ERR_UNSUPPORTED;
    // dX = dY (*) flipped filters, padding k - 1 - p: output H x W again
    return ct_run_fwd(dy, w, nullptr, dx, N, F, OH, OW, C, H, W, kh, kw, kh - 1 - p, kw - 1 - p, F, C, 1, ws, ws_bytes, st);
}

// xs[jj][n][c][h][w] = x[n][c][h][w + d(jj)] (0 outside the row), d running over the kw - 1 non-zero values of j - p:
// the only unaligned access of the wgrad path, done once per call by plain loads
__global__ void __launch_bounds__(256)
conv_tma_shift_kernel(const float *__restrict__ x, float *__restrict__ xs, long long rows, int W, int kw, int p) {
    const int w4 = W >> 2;
    const long long per = rows * w4;
    const long long total = per * (kw - 1);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int jj = (int)(idx / per);
        const long long r = idx - (long long)jj * per;
        const long long row = r / w4;
        const int c0 = (int)(r - row * w4) * 4;
        const int j = jj < p ? jj : jj + 1;
        const int d = j - p;
        const float *src = x + row * W;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int cc = c0 + e + d;
            v[e] = (cc >= 0 && cc < W) ? __ldg(src + cc) : 0.0f;
        }
        *reinterpret_cast<float4 *>(xs + ((long long)jj * rows + row) * W + c0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// shifted-tile buffers of conv_s1_wgrad2_kernel (dk_tc_debug_set key 26).  Measured at cfg2 (3x3 64 -> 64 @56x56, batch 128):
// 188.2 / 188.1 / 188.1 us with 2 / 3 / 4 buffers -- letting the shifter run further ahead changes nothing, so the exposed
// hand-over latency the stall samples suggested is NOT what holds the kernel at 25 % tensor pipe; 2 keeps all 6 raw stages.
int g_cw2_nb = 2;
int g_cw2_dbg = 0;  // dk_tc_debug_set key 27: Cw2Params::dbg
int conv_tma_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                   int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (s != 1 || !g_ct_ready || !g_conv_tma_enabled) return DK_ERR_UNSUPPORTED;
    const int OH = H + 2 * p - kh + 1, OW = W + 2 * p - kw + 1;
    if (OH < 1 || OW < 1 || kh * kw <= 1 || kh > 5 || kw > 7 || p >= kw) return DK_ERR_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(dy) || (W % 4) != 0 || (OW % 4) != 0 || F > 128 || F < 8 || C < 8 || OW < 8) return DK_ERR_UNSUPPORTED;
    if ((int64_t)N * kw >= (1 << 30)) return DK_ERR_UNSUPPORTED;
    if (g_ct_wgrad2 && F <= 64 && ct_round_up(C, 32) <= 64 && kw * ct_round_up(C, 32) <= 256 && p <= 4 && kw - 1 - p <= 4 &&
        ((kh + 1) / 2) * kw * ct_round_up(C, 32) <= 512) {
        Cw2Params q = {};
        q.N = N; q.C = C; q.H = H; q.W = W; q.F = F; q.OH = OH; q.OW = OW; q.kh = kh; q.kw = kw; q.p = p;
        q.bnC = ct_round_up(C, 32);
        q.a_rows = ct_round_up(F, 8);
        q.csegs = (OW + 31) / 32;
        const int strips = N * q.csegs;
        // work unit = rc X rows of one strip; pick the chunking that fills the SMs best (each unit re-reads kh - 1 dY rows)
        int best_rc = H;
        double best_cost = 1e30;
        for (int div = 1; div <= 8; ++div) {
            const int rc = (H + div - 1) / div;
            if (rc < 4 && div > 1) break;
            const int chunks = (H + rc - 1) / rc;
            const long long units = (long long)strips * chunks;
            const double cost = (double)ceil_div(units, sm_count()) * (rc + kh - 1);
            if (cost < best_cost - 1e-9) { best_cost = cost; best_rc = rc; }
        }
        q.rc = best_rc;
        q.chunks = (H + q.rc - 1) / q.rc;
        q.units = strips * q.chunks;
        q.units_per_cta = (int)ceil_div(q.units, sm_count());
        const int ctas = (int)ceil_div(q.units, q.units_per_cta);
        q.npairs = (kh + 1) / 2;
        q.acc_stride = (uint32_t)(kw * q.bnC);
        q.tmem_cols = ct_pow2_cols((uint32_t)q.npairs * q.acc_stride);
        q.ring = kh + 5 > 15 ? 15 : kh + 5;
        // Shifted-tile buffers (g_cw2_nb): more than two let the shifter warps run further ahead of the MMAs, at the price of
        // raw-box stages (measured: no effect, see g_cw2_nb).
        q.nb = g_cw2_nb < 2 ? 2 : g_cw2_nb > CW2_NB_MAX ? CW2_NB_MAX : g_cw2_nb;
        q.raw_stages = CW2_RAW_STAGES;
        q.dbg = g_cw2_dbg;
        while (q.nb > 2 || q.raw_stages > 3) {
            const size_t need = (size_t)(q.ring + 1) * CW2_DY_SLOT + (size_t)q.nb * (kw * q.bnC) * 128 +
                                (size_t)q.raw_stages * q.bnC * CW2_RAW_W * 4 + 1024 + 16 + 8 * 64;
            if (need <= (size_t)CT_SMEM_MAX) break;
            if (q.raw_stages > 3) --q.raw_stages;
            else --q.nb;
        }
        const int taps = kh * kw;
        const size_t need = (size_t)ctas * F * C * taps * sizeof(float);
        if (ws == nullptr || ws_bytes < need + 1024) return DK_ERR_UNSUPPORTED;
        q.partial = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ws) + 255u) & ~(uintptr_t)255u);
        CUtensorMap ta, tr;
        const uint64_t da[4] = {(uint64_t)OW, (uint64_t)OH, (uint64_t)F, (uint64_t)N};
        const uint32_t ba[4] = {32, 1, (uint32_t)q.a_rows, 1};
        int rc2 = ct_map(&ta, dy, 4, da, ba, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc2) return rc2;
        const uint64_t dbr[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
        const uint32_t bbr[4] = {CW2_RAW_W, 1, (uint32_t)q.bnC, 1};
        rc2 = ct_map(&tr, x, 4, dbr, bbr, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc2) return rc2;
        const size_t smem = (size_t)(q.ring + 1) * CW2_DY_SLOT + (size_t)q.nb * (kw * q.bnC) * 128 + (size_t)q.raw_stages * q.bnC * CW2_RAW_W * 4 + 1024 + 16 + 8 * 64;
        if (smem > (size_t)CT_SMEM_MAX) return DK_ERR_UNSUPPORTED;
        conv_s1_wgrad2_kernel<<<ctas, CT_THREADS, smem, st>>>(ta, tr, q);
        DK_LAUNCH_CHECK();
        if (!(q.dbg & 32))
            cw2_reduce_kernel<<<(unsigned)ceil_div((int64_t)F * C * taps, 32), 1024, 0, st>>>(q.partial, w, dw, l2, F, C, taps, ctas);
        DK_LAUNCH_CHECK();
        return DK_OK;
    }
    CwParams q = {};
    q.N = N; q.C = C; q.H = H; q.W = W; q.F = F; q.OH = OH; q.OW = OW; q.kh = kh; q.kw = kw; q.p = p;
    q.bnC = ct_round_up(C, 32);
    if (q.bnC > 256) return DK_ERR_UNSUPPORTED;
    q.acc_stride = ct_acc_stride(q.bnC);
    if ((uint32_t)kh * q.acc_stride > 512) return DK_ERR_UNSUPPORTED;
    q.tmem_cols = ct_pow2_cols((uint32_t)kh * q.acc_stride);
    q.a_rows = ct_round_up(F, 8);
    q.a_bytes = 128u * 128u;  // the MMA reads 128 rows: the slot is always 16 KB, the box fills the first a_rows
    q.b_bytes = (uint32_t)q.bnC * 128u;
    q.stage_bytes = q.a_bytes + q.b_bytes;
    q.csegs = (OW + 31) / 32;
    q.strips = N * q.csegs;
    int splits = sm_count() / kw;
    if (splits < 1) splits = 1;
    if (splits > q.strips) splits = q.strips;
    q.strips_per_split = (int)ceil_div(q.strips, splits);
    q.splits = (int)ceil_div(q.strips, q.strips_per_split);
    int stg = (CT_SMEM_MAX - 1024 - 256) / (int)q.stage_bytes;
    if (stg < kh + 2) return DK_ERR_UNSUPPORTED;
    q.stages = stg > CT_MAX_STAGES ? CT_MAX_STAGES : stg;
    const int taps = kh * kw;
    const size_t shifted = (size_t)(kw - 1) * N * C * H * W * sizeof(float);
    const size_t need = (size_t)q.splits * F * C * taps * sizeof(float);
    if (ws == nullptr || ws_bytes < shifted + need + 1024) return DK_ERR_UNSUPPORTED;
    float *xs = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ws) + 255u) & ~(uintptr_t)255u);
    q.partial = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(xs) + shifted + 255u) & ~(uintptr_t)255u);
    if (kw > 1) {
        const long long rows = (long long)N * C * H;
        const long long total = rows * (W / 4) * (kw - 1);
        conv_tma_shift_kernel<<<stream_grid(total, 256), 256, 0, st>>>(x, xs, rows, W, kw, p);
        DK_LAUNCH_CHECK();
    }
    CUtensorMap ta, tb, tbs;
    const uint64_t da[4] = {(uint64_t)OW, (uint64_t)OH, (uint64_t)F, (uint64_t)N};
    const uint32_t ba[4] = {32, 1, (uint32_t)q.a_rows, 1};
    int rc = ct_map(&ta, dy, 4, da, ba, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const uint64_t db[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N};
    const uint32_t bb[4] = {32, 1, (uint32_t)q.bnC, 1};
    rc = ct_map(&tb, x, 4, db, bb, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    tbs = tb;
    if (kw > 1) {
        const uint64_t dbs[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)C, (uint64_t)N * (uint64_t)(kw - 1)};
        rc = ct_map(&tbs, xs, 4, dbs, bb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = (size_t)q.stages * q.stage_bytes + 1024 + 8 * (2 * CT_MAX_STAGES + 8);
    conv_s1_wgrad_kernel<<<q.splits * kw, CT_THREADS, smem, st>>>(ta, tb, tbs, q);
    DK_LAUNCH_CHECK();
    splitk_reduce_launch(q.partial, w, dw, l2, (int64_t)F * C * taps, q.splits, st);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
