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
        // Whole warp, uniform control flow; ONE elected lane issues every box of a stage from uniform registers (a few cycles
        // per box).  History: a single lane inside `if (lane == 0)` was the bottleneck of the first version, one box per lane
        // (divergent coordinates) the second -- both pay an ELECT / R2UR.BROADCAST waterfall of ~100 cycles per box, which for
        // 12-21 boxes per stage is more than the stage's MMAs take.
        if (p.w_resident && elect_one()) {
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
                mbar_wait(empty_bar(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(fb, tx);
                    uint32_t dst = sA;
                    for (int rr = 0; rr * p.nb < p.nbox; ++rr)
                        for (int cb = 0; cb < p.nb; ++cb, dst += p.box_bytes)
                            tma_load_4d(dst, &tmX, fb, 32 * cb, r0 + rr - p.ph, ch, n);
                    if (!p.w_resident) {  // (kc == 32 here)
                        const uint32_t sB = sA + (uint32_t)p.nbox * p.box_bytes;
                        for (int t = 0; t < taps; ++t) tma_load_3d(sB + (uint32_t)t * p.b_bytes, &tmW, fb, ch, 0, t);
                    }
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
        // (whole warp, uniform control flow, one elected lane issues)
        {
            int s = 0;
            uint32_t ph = 0;
            for (int strip = sb; strip < se; ++strip) {
                const int n = strip / p.csegs, c0 = (strip - n * p.csegs) * 32;
                for (int t = 0; t < steps; ++t) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sA = smem_base + (uint32_t)s * p.stage_bytes, sB = sA + p.a_bytes, fb = full_bar(s);
                    const int oh = t - (p.kh - 1);
                    if (elect_one()) {
                        mbar_expect_tx(fb, p.b_bytes + (oh >= 0 ? (uint32_t)p.a_rows * 128u : 0u));
                        // X row t - p (zero rows above / below) of the copy shifted by j - p columns: aligned boxes only (a box
                        // starting at an odd column traps); the unshifted tap column reads X itself
                        if (j == p.p) tma_load_4d(sB, &tmX, fb, c0, t - p.p, 0, n);
                        else tma_load_4d(sB, &tmXS, fb, c0, t - p.p, 0, (j < p.p ? j : j - 1) * p.N + n);
                        if (oh >= 0) tma_load_4d(sA, &tmDY, fb, c0, oh, 0, n);
                    }
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
    int ring;                            // dY ring slots (+1 mirror), a multiple of the rows per step
    int nb;                              // shifted-tile buffers (of R rows each)
    int raw_stages;                      // raw X stages (of R row boxes each) in flight
    int npairs;
    uint32_t tmem_cols, acc_stride;      // acc_stride = kw * bnC columns per row-tap pair
    float *partial;                      // [CTAs][tap][F][C]
};

// 16 pixels of one channel row, shifted by D columns, as four 16-byte chunks of a swizzled K-major tile row
template <int D>
__device__ __forceinline__ void cw2_store_shifted(const float (&v)[24], uint8_t *row, uint32_t chunk0, uint32_t sw) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
        *reinterpret_cast<float4 *>(row + (((chunk0 + kk) ^ sw) << 4)) =
            make_float4(v[4 + 4 * kk + D], v[5 + 4 * kk + D], v[6 + 4 * kk + D], v[7 + 4 * kk + D]);
}

constexpr int CW2_RAW_W = 40;  // raw X box: 4 halo pixels left, 32, 4 right
constexpr uint32_t CW2_DY_SLOT = 8192;

// R = X rows per pipeline step.  Every role of this kernel is a single warp (or four) walking a loop of mbarrier waits,
// address arithmetic and issues, i.e. bound by instruction LATENCY: ~600 cycles of such overhead per step were measured in
// the MMA warp with every load, shift and MMA removed, against 768 cycles of tensor work (8 MMAs of 128 x 192 x 8) in a 3 x 3
// one-row step -- and tcgen05.mma itself occupies its issuing warp for ~70 cycles.  Two rows per step (R = 2: one barrier
// round trip, 16 MMAs) halve that overhead per MMA; data stays row-granular (8 KB dY slots, one raw box and kw shifted
// tiles per row), only the hand-shakes cover R rows.  R = 2 needs kh == 3 (the run-in of kh - 1 dY rows is then exactly one group).
// KH = kh as a compile-time constant (0: read it from the parameters): unrolls the row-tap pairs of the issue loop.
template <int R, int KH>
__global__ void __launch_bounds__(CT_THREADS, 1)
conv_s1_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmXR, const Cw2Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring_base = smem_base;                                     // (ring + 1) x 8 KB
    const uint32_t b_row = (uint32_t)(p.kw * p.bnC) * 128u;                   // kw shifted tiles of bnC rows: one X row
    const uint32_t b_buf = (uint32_t)R * b_row;
    const uint32_t b_base = ring_base + (uint32_t)(p.ring + 1) * CW2_DY_SLOT;  // p.nb buffers of R rows
    const uint32_t raw_bytes = (uint32_t)p.bnC * CW2_RAW_W * 4u;
    const uint32_t raw_base = b_base + (uint32_t)p.nb * b_buf;                // p.raw_stages stages of R raw boxes
    const uint32_t bar_base = (raw_base + (uint32_t)(p.raw_stages * R) * raw_bytes + 15u) & ~15u;
    auto dyfull = [&](int g) { return bar_base + 8u * g; };          // per GROUP of R dY rows
    auto dyempty = [&](int g) { return bar_base + 8u * (16 + g); };
    auto rawfull = [&](int s) { return bar_base + 8u * (32 + s); };
    auto rawempty = [&](int s) { return bar_base + 8u * (40 + s); };
    auto bfull = [&](int s) { return bar_base + 8u * (48 + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (52 + s); };
    const uint32_t tfull_bar = bar_base + 8u * 56;
    const uint32_t tmem_slot = bar_base + 8u * 57;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = p.ring / R;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmXR);
        for (int g = 0; g < groups; ++g) {
            mbar_init(dyfull(g), 1);
            mbar_init(dyempty(g), 1);
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
    const int kh = KH ? KH : p.kh, npairs = (kh + 1) / 2;
    const int run_in = kh - 1;     // dY rows loaded ahead of the first step of a unit (a multiple of R)
    const int lag = run_in / R;    // a dY group is last read `lag` steps after the one that waited for it

    // Roles keep their control flow warp-uniform and elect one lane around the issues (see elect_one()); ring positions and
    // phases are carried incrementally -- the loops below ARE the step time, every division or branch in them counts.
    if (warp == 0) {
        // ================================ TMA producer ================================
        int gs = 0, rs = 0;           // dY group slot, raw X stage
        uint32_t dph = 1u, rph = 1u;  // phases of the "empty" waits (first lap passes)
        const uint32_t dy_bytes = (uint32_t)p.a_rows * 128u;
        for (int u = u_lo; u < u_hi; ++u) {
            // unit order (image, row chunk, column segment): the segments of a row band follow each other on the same CTA, so the
            // 64-byte DRAM granules and the halo columns they share are still in L2 when the second one asks for them
            const int cseg = u % p.csegs, t = u / p.csegs;
            const int chunk = t % p.chunks, n = t / p.chunks, c0 = cseg * 32;
            const int r_lo = chunk * p.rc;
            int r_hi = r_lo + p.rc;
            if (r_hi > p.H) r_hi = p.H;
            const int oh0 = r_lo + p.p - run_in;  // first dY row of the unit (rows outside the tensor arrive as zeros)
            const int nload = (r_hi - r_lo) + run_in;
            for (int t0 = 0; t0 < nload; t0 += R) {
                const int nv = nload - t0 < R ? nload - t0 : R;
                mbar_wait(dyempty(gs), dph);
                if (elect_one()) {
                    const uint32_t fb = dyfull(gs);
                    mbar_expect_tx(fb, (uint32_t)(nv + (gs == 0 ? 1 : 0)) * dy_bytes);
#pragma unroll
                    for (int rr = 0; rr < R; ++rr)
                        if (rr < nv) tma_load_4d(ring_base + (uint32_t)(gs * R + rr) * CW2_DY_SLOT, &tmDY, fb, c0, oh0 + t0 + rr, 0, n);
                    // slot 0 is mirrored behind the last slot: a pair of consecutive rows never wraps
                    if (gs == 0) tma_load_4d(ring_base + (uint32_t)p.ring * CW2_DY_SLOT, &tmDY, fb, c0, oh0 + t0, 0, n);
                }
                if (t0 >= run_in) {  // the X rows of this step
                    mbar_wait(rawempty(rs), rph);
                    if (elect_one()) {
                        mbar_expect_tx(rawfull(rs), (uint32_t)nv * raw_bytes);
#pragma unroll
                        for (int rr = 0; rr < R; ++rr)
                            if (rr < nv)
                                tma_load_4d(raw_base + (uint32_t)(rs * R + rr) * raw_bytes, &tmXR, rawfull(rs), c0 - 4,
                                            r_lo + (t0 - run_in) + rr, 0, n);
                    }
                    if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
                }
                if (++gs == groups) { gs = 0; dph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // tcgen05.mma blocks its issuing warp until the tensor pipe accepts it and the pipe queues next to nothing, so every
        // cycle this warp spends between two steps is a cycle the pipe idles (measured: 769 cycles per 8 MMAs of N = 192
        // back to back, ~1200 with this loop's first version).  Hence ONE barrier per step here -- the shifter warps wait for
        // the dY groups and their "tiles ready" arrival covers both operands -- and one elected block with all the issues.
        const uint32_t idesc = idesc_tf32(128, p.kw * p.bnC, 0, 0);
        const uint32_t d_hi = smem_desc_hi(1024u, LAYOUT_SW128);
        int gs = 0, bs = 0;
        uint32_t bph = 0, started = 0;
        for (int u = u_lo; u < u_hi; ++u) {
            const int r_lo = ((u / p.csegs) % p.chunks) * p.rc;
            int r_hi = r_lo + p.rc;
            if (r_hi > p.H) r_hi = p.H;
            const int nrows = r_hi - r_lo;
            gs += lag;  // the run-in groups
            if (gs >= groups) gs -= groups;
            for (int x0 = 0; x0 < nrows; x0 += R) {
                const int nv = nrows - x0 < R ? nrows - x0 : R;
                int gold = gs - lag;  // the oldest dY group of this step
                if (gold < 0) gold += groups;
                const uint32_t b_lo = smem_desc_lo(b_base + (uint32_t)bs * b_buf, 16u);
                const uint32_t a_slot = smem_desc_lo(ring_base + (uint32_t)(gs * R) * CW2_DY_SLOT, 16u);
                const uint32_t a_wrap = (uint32_t)p.ring * (CW2_DY_SLOT >> 4);
                const uint32_t a_min = smem_desc_lo(ring_base, 16u);
                mbar_wait(bfull(bs), bph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        if (rr < nv) {
                            // row tap i pairs this X row with the dY row loaded i loads ago: accumulator k holds taps 2k + 1
                            // (TMEM lanes 0-63) and 2k (lanes 64-127), or the unpaired last tap in lanes 0-63
#pragma unroll
                            for (int k = 0; k < (KH ? (KH + 1) / 2 : 3); ++k) {
                                if (k < npairs) {
                                    const int back = (2 * k + 1 < kh ? 2 * k + 1 : 2 * k) - rr;  // slots behind the group's first
                                    uint32_t a_lo = a_slot - (uint32_t)back * (CW2_DY_SLOT >> 4);
                                    if ((int)(a_lo - a_min) < 0) a_lo += a_wrap;
                                    mma_tf32_k4(tmem_base + (uint32_t)k * p.acc_stride, a_lo, d_hi, b_lo + (uint32_t)rr * (b_row >> 4), d_hi, 2u,
                                                2u, idesc, started | (uint32_t)rr);
                                }
                            }
                        }
                    }
                    mma_commit(bempty(bs));
                    mma_commit(dyempty(gold));
                    if (x0 + R >= nrows)  // end of the unit: the groups a next step would still have read
                        for (int d = lag - 1; d >= 0; --d) {
                            int gd = gs - d;
                            if (gd < 0) gd += groups;
                            mma_commit(dyempty(gd));
                        }
                }
                started = 1u;
                if (++bs == p.nb) { bs = 0; bph ^= 1u; }
                if (++gs == groups) gs = 0;
            }
        }
        if (elect_one()) mma_commit(tfull_bar);
        __syncwarp();
    } else {
        // ================================ shifter (warps 2..5), then epilogue ==========
        const int t128 = threadIdx.x - 64;          // 0..127
        const int c = t128 >> 1, hh = t128 & 1;      // channel row, 16-pixel half
        int rs = 0, bs = 0, gs = 0;
        uint32_t rph = 0, bph = 1u, dph = 0;
        for (int u = u_lo; u < u_hi; ++u) {
            const int r_lo = ((u / p.csegs) % p.chunks) * p.rc;
            int r_hi = r_lo + p.rc;
            if (r_hi > p.H) r_hi = p.H;
            const int nrows = r_hi - r_lo;
            // every dY group is waited for exactly once, in order, HERE: the "tiles ready" arrival below then tells the MMA warp
            // that both operands of the step have landed (TMA loads may complete out of order; the run-in groups carry no step)
            for (int d = 0; d < lag; ++d) {
                mbar_wait(dyfull(gs), dph);
                if (++gs == groups) { gs = 0; dph ^= 1u; }
            }
            for (int x0 = 0; x0 < nrows; x0 += R) {
                const int nv = nrows - x0 < R ? nrows - x0 : R;
                mbar_wait(rawfull(rs), rph);
                mbar_wait(bempty(bs), bph);
                if (c < p.bnC) {
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) {
                        if (rr < nv) {
                            // raw[c][16*hh .. 16*hh + 23] covers the 16 pixels of this half shifted by -4 .. +4
                            const float4 *src = reinterpret_cast<const float4 *>(smem_raw + (raw_base - smem_u32(smem_raw)) +
                                                                                 (uint32_t)(rs * R + rr) * raw_bytes +
                                                                                 (uint32_t)c * (CW2_RAW_W * 4) + (uint32_t)hh * 64u);
                            // The 160-byte row pitch puts channels c and c + 4 of a quarter-warp (threads c = 4 consecutive channels x
                            // two halves) on the same banks: odd channel pairs read their six chunks rotated by one and
                            // un-rotate in registers (selects) -- conflict-free instead of two wavefronts per load.
                            const bool rot = (c >> 1) & 1;
                            float4 w4[6];
#pragma unroll
                            for (int k4 = 0; k4 < 6; ++k4) w4[k4] = src[rot ? (k4 == 5 ? 0 : k4 + 1) : k4];
                            float v[24];
#pragma unroll
                            for (int k4 = 0; k4 < 6; ++k4) {
                                const float4 a4 = w4[k4], b4 = w4[k4 == 0 ? 5 : k4 - 1];
                                const float4 t4 = rot ? b4 : a4;
                                v[4 * k4] = t4.x; v[4 * k4 + 1] = t4.y; v[4 * k4 + 2] = t4.z; v[4 * k4 + 3] = t4.w;
                            }
                            uint8_t *bt = smem_raw + (b_base - smem_u32(smem_raw)) + (uint32_t)bs * b_buf + (uint32_t)rr * b_row +
                                          (uint32_t)c * 128u;
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
                    }
                }
                fence_proxy_async();
                mbar_wait(dyfull(gs), dph);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bfull(bs));
                    mbar_arrive(rawempty(rs));
                }
                if (++rs == p.raw_stages) { rs = 0; rph ^= 1u; }
                if (++bs == p.nb) { bs = 0; bph ^= 1u; }
                if (++gs == groups) { gs = 0; dph ^= 1u; }
            }
        }
        // epilogue: accumulator k, lanes 0-63 = tap 2k+1 (or the unpaired tap), lanes 64-127 = tap 2k
        const int q = warp & 3;
        if (u_lo < u_hi) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
        const int f = 32 * (q & 1) + lane;
        const int taps = p.kh * p.kw;
        // partial sums as [CTA][tap][f][c]: a lane owns filter f and stores its 32 channels of a tap as eight 16-byte vectors.
        // (The first version wrote the final [f][c][tap] order directly: 4-byte stores 36 bytes apart, 49 K of them per CTA -- a
        // fifth of the kernel, and 173 MB of DRAM read-for-ownership; the second, [tap][c][f] with coalesced 4-byte stores, was
        // still 1152 store instructions per warp.  cw2_reduce_kernel transposes while it adds the partials.)
        float *o = p.partial + ((long long)blockIdx.x * taps * p.F + f) * p.C;
        const bool vec4 = (p.C & 3) == 0;
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
                    if (live && f < p.F) {
                        float *ot = o + (long long)(i * p.kw + j) * p.F * p.C + cc;
                        if (vec4) {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (cc + 4 * e < p.C)
                                    *reinterpret_cast<float4 *>(ot + 4 * e) = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                                                                          __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj)
                                if (cc + jj < p.C) ot[jj] = __uint_as_float(v[jj]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// dW[f][c][tap] = sum_z partial[z][tap][f][c] + l2 * W[f][c][tap]: V consecutive channels per thread (16-byte loads when C % 4
// == 0), the Z partials spread over 32 thread rows and combined in a fixed order (deterministic), scattered 4-byte writes of
// the (small) result
template <int V>
__global__ void __launch_bounds__(1024)
cw2_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ dw, float l2, int F, int C,
                  int taps, int Z) {
    __shared__ float red[32][32 * V + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long mn = (long long)taps * C * F;
    const long long i = ((long long)blockIdx.x * 32 + tx) * V;
    float s[V];
#pragma unroll
    for (int e = 0; e < V; ++e) s[e] = 0.0f;
    if (i < mn) {
        for (int z = ty; z < Z; z += 32) {
            const float *q = partial + (long long)z * mn + i;
            if (V == 4) {
                const float4 t = *reinterpret_cast<const float4 *>(q);
                s[0] += t.x; s[V > 1 ? 1 : 0] += t.y; s[V > 2 ? 2 : 0] += t.z; s[V > 3 ? 3 : 0] += t.w;
            } else {
                s[0] += q[0];
            }
        }
    }
#pragma unroll
    for (int e = 0; e < V; ++e) red[ty][tx * V + e] = s[e];
    __syncthreads();
    // 32 * V results per block: thread t < 32 * V adds the 32 rows of column t
    if (threadIdx.x < 32 * V) {
        const long long ii = (long long)blockIdx.x * 32 * V + threadIdx.x;
        if (ii < mn) {
            float t = red[0][threadIdx.x];
#pragma unroll
            for (int y = 1; y < 32; ++y) t += red[y][threadIdx.x];
            const int c = (int)(ii % C);
            const long long r = ii / C;
            const int f = (int)(r % F), tap = (int)(r / F);
            const long long o = ((long long)f * C + c) * taps + tap;
            dw[o] = t + (l2 != 0.0f ? l2 * w[o] : 0.0f);
        }
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
        cudaFuncSetAttribute(conv_s1_wgrad2_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess ||
        cudaFuncSetAttribute(conv_s1_wgrad2_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess ||
        cudaFuncSetAttribute(conv_s1_wgrad2_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM_MAX + 2048) != cudaSuccess) {
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

// X rows per step of conv_s1_wgrad2_kernel (dk_tc_debug_set key 26): 0 = two when kh == 3 and the buffers fit, 1 = always one
// (measured at cfg2: 87.6 us with two, 90.6 us with one).
// (More shifted-tile buffers -- 2 / 3 / 4, measured 188.2 / 188.1 / 188.1 us on the first version -- never helped: the step
// time is the issue warps' instruction latency, not a hand-over the shifter could hide by running ahead.)
int g_cw2_rows = 0;
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
        q.nb = 2;
        const size_t b_row = (size_t)(kw * q.bnC) * 128, raw_box = (size_t)q.bnC * CW2_RAW_W * 4, fixed = 1024 + 16 + 8 * 64;
        auto smem_need = [&](int rows) {
            return (size_t)(q.ring + 1) * CW2_DY_SLOT + (size_t)q.nb * rows * b_row + (size_t)q.raw_stages * rows * raw_box + fixed;
        };
        // two rows per step (kh == 3: the run-in is one group): 4 groups of dY rows, the raw stages that still fit (>= 2)
        int rows = 1;
        if (kh == 3 && g_cw2_rows != 1) {
            q.ring = 8;
            q.raw_stages = 3;
            if (smem_need(2) > (size_t)CT_SMEM_MAX) q.raw_stages = 2;
            if (smem_need(2) <= (size_t)CT_SMEM_MAX) rows = 2;
        }
        if (rows == 1) {
            q.ring = kh + 5 > 15 ? 15 : kh + 5;
            q.raw_stages = 6;
            while (q.raw_stages > 3 && smem_need(1) > (size_t)CT_SMEM_MAX) --q.raw_stages;
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
        const size_t smem = smem_need(rows);
        if (smem > (size_t)CT_SMEM_MAX) return DK_ERR_UNSUPPORTED;
        if (rows == 2) conv_s1_wgrad2_kernel<2, 3><<<ctas, CT_THREADS, smem, st>>>(ta, tr, q);
        else if (kh == 3) conv_s1_wgrad2_kernel<1, 3><<<ctas, CT_THREADS, smem, st>>>(ta, tr, q);
        else conv_s1_wgrad2_kernel<1, 0><<<ctas, CT_THREADS, smem, st>>>(ta, tr, q);
        DK_LAUNCH_CHECK();
        if (C % 4 == 0) cw2_reduce_kernel<4><<<(unsigned)ceil_div((int64_t)F * C * taps, 128), 1024, 0, st>>>(q.partial, w, dw, l2, F, C, taps, ctas);
        else cw2_reduce_kernel<1><<<(unsigned)ceil_div((int64_t)F * C * taps, 32), 1024, 0, st>>>(q.partial, w, dw, l2, F, C, taps, ctas);
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
