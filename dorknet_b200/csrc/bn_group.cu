// bn_group.cu -- BatchNorm forward / backward for SMALL planes (7x7, 14x14: the last two stages of ResNet-18-depsep) as one
// kernel each, with a channel GROUP resident in shared memory.
//
// At 7x7 a plane is 196 bytes: not a 16-byte multiple, so the cluster kernels of bn_fused.cu fall back to plain
// per-element loads and are latency-bound (12.6 us for a 6.4 MB tensor).  But G = 4 consecutive channels of one image ARE
// one contiguous, 16-byte aligned run of G*HW floats (784 B at 7x7; 14x14 has that with G = 1), so a CTA that owns G
// channels of ALL N images pulls its whole working set in with N bulk-async copies -- a [N][G*HW] tile, 50 KB at batch 64
// -- and both passes run out of shared memory: 2n (forward) / 3n (backward) of traffic, no cluster, no second kernel.
// A thread owns one float4 COLUMN of the tile and a range of rows: consecutive threads read consecutive 16-byte words
// (conflict-free), the per-column partial sums are reduced per channel in a fixed order (deterministic), and the second
// pass writes 16-byte words straight to global memory (G*HW*4-byte runs per image, coalesced).
// Semantics are those of bn_fused.cu / batchnorm.cu (batch_norm.py:54-174): shifted single-pass sums, biased variance.
#include <cuda.h>
#include <stdlib.h>

#include "bn.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int BG_THREADS = 256;
constexpr int BG_MAX_G = 4;
constexpr int BG_SMEM_MAX = 200 * 1024;

struct BgGeom {
    int N, C, HW, G;
    int L4;    // float4 columns of a tile row (G*HW/4)
    int RG;    // row groups (threads = RG * L4 active)
    int rows;  // rows per group
};

__device__ __forceinline__ void bg_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// rows n = 0..N-1 of NT tensors: the G channels c0.. of image n -> tile[t][n][G*HW]
template <int NT>
__device__ __forceinline__ void bg_load(const float *const (&src)[NT], float *const (&dst)[NT], const BgGeom &g, int c0,
                                        uint32_t bar) {
    const int L = 4 * g.L4;
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(NT * g.N * L) * 4u);
    __syncthreads();
    for (int n = threadIdx.x; n < g.N; n += BG_THREADS) {
        const long long goff = ((long long)n * g.C + c0) * g.HW;
#pragma unroll
        for (int t = 0; t < NT; ++t) bg_bulk_g2s(smem_u32(dst[t] + (size_t)n * L), src[t] + goff, (uint32_t)L * 4u, bar);
    }
    mbar_wait(bar, 0u);
}

// per-channel totals of two per-thread float4 partial sums (component e of column q belongs to channel (4q+e)/HW):
// part[2][RG][L] -> colsum over the row groups -> per-channel sums in tot[2][G]; fixed order
__device__ __forceinline__ void bg_channel_sums(const float (&pa)[4], const float (&pb)[4], bool active, int q, int rg,
                                                const BgGeom &g, float *part, float *tot) {
    const int L = 4 * g.L4;
    if (active) {
        *reinterpret_cast<float4 *>(part + (size_t)rg * L + 4 * q) = make_float4(pa[0], pa[1], pa[2], pa[3]);
        *reinterpret_cast<float4 *>(part + (size_t)(g.RG + rg) * L + 4 * q) = make_float4(pb[0], pb[1], pb[2], pb[3]);
    }
    __syncthreads();
    // column sums over the row groups, into row 0 of each half
    for (int f = threadIdx.x; f < 2 * L; f += BG_THREADS) {
        const int half = f >= L ? 1 : 0, col = f - half * L;
        float *p = part + (size_t)half * g.RG * L + col;
        float s = p[0];
        for (int r = 1; r < g.RG; ++r) s += p[(size_t)r * L];
        p[0] = s;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int w = warp; w < 2 * g.G; w += BG_THREADS / 32) {
        const int half = w >= g.G ? 1 : 0, ch = w - half * g.G;
        const float *p = part + (size_t)half * g.RG * L + ch * g.HW;
        float s = 0.0f;
        for (int i = lane; i < g.HW; i += 32) s += p[i];
        s = warp_sum(s);
        if (lane == 0) tot[half * BG_MAX_G + ch] = s;
    }
    __syncthreads();
}

template <bool RELU>
__global__ void __launch_bounds__(BG_THREADS)
bn_group_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ add, const BgGeom g,
                    const BnFinalize fin) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ float tot[2 * BG_MAX_G];
    __shared__ float chs[BG_MAX_G], chh[BG_MAX_G];
    const int L = 4 * g.L4;
    float *tile = smem, *part = smem + (size_t)g.N * L;
    const int c0 = blockIdx.x * g.G;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    BnChannelParams cp = {1.0f, 0.0f, 0.0f, 0.0f};
    if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < g.G) cp = bn_load_channel_params(fin, c0 + (threadIdx.x >> 5));
    __syncthreads();
    {
        const float *const src[1] = {x};
        float *const dst[1] = {tile};
        bg_load<1>(src, dst, g, c0, bar);
    }
    const int q = threadIdx.x % g.L4, rg = threadIdx.x / g.L4;
    const bool active = rg < g.RG;
    int ce[4];
    float shift[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        ce[e] = (4 * q + e) / g.HW;
        shift[e] = tile[ce[e] * g.HW];  // the channel's first value (image 0): one shift per channel, as in bn_fused.cu
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const int n0 = rg * g.rows, n1 = (n0 + g.rows < g.N) ? n0 + g.rows : g.N;
    if (active) {
        for (int n = n0; n < n1; ++n) {
            const float4 v = *reinterpret_cast<const float4 *>(tile + (size_t)n * L + 4 * q);
            const float d0 = v.x - shift[0], d1 = v.y - shift[1], d2 = v.z - shift[2], d3 = v.w - shift[3];
            s[0] += d0; s[1] += d1; s[2] += d2; s[3] += d3;
            ss[0] = fmaf(d0, d0, ss[0]); ss[1] = fmaf(d1, d1, ss[1]); ss[2] = fmaf(d2, d2, ss[2]); ss[3] = fmaf(d3, d3, ss[3]);
        }
    }
    bg_channel_sums(s, ss, active, q, rg, g, part, tot);
    if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < g.G) {
        const int ch = threadIdx.x >> 5;
        const float inv_n = 1.0f / (float)(g.N * g.HW);
        const float s1 = tot[ch], s2 = tot[BG_MAX_G + ch];
        const float mean = tile[ch * g.HW] + s1 * inv_n;
        const float var = fmaxf(s2 - s1 * s1 * inv_n, 0.0f) * inv_n;  // biased (batch_norm_stats_cy.pyx:44)
        float sc, sh;
        bn_finalize_channel(fin, cp, c0 + ch, mean, var, true, &sc, &sh);
        chs[ch] = sc;
        chh[ch] = sh;
    }
    __syncthreads();
    if (y != nullptr && active) {
        float sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            sc[e] = chs[ce[e]];
            sh[e] = chh[ce[e]];
        }
        for (int nb = n0; nb < n1; nb += 4) {
            // the skip values of four rows are requested before anything is stored (see bn_fused.cu)
            float4 a[4];
            if (add != nullptr) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (nb + u < n1) a[u] = ld_stream4(add + ((long long)(nb + u) * g.C + c0) * g.HW + 4 * q);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = nb + u;
                if (n < n1) {
                    float4 v = *reinterpret_cast<const float4 *>(tile + (size_t)n * L + 4 * q);
                    v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]); v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
                    if (add != nullptr) {
                        v.x += a[u].x; v.y += a[u].y; v.z += a[u].z; v.w += a[u].w;
                    }
                    if (RELU) {
                        v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                        v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
                    }
                    st_stream4(y + ((long long)n * g.C + c0) * g.HW + 4 * q, v);
                }
            }
        }
    }
}

template <bool RELU>
__global__ void __launch_bounds__(BG_THREADS)
bn_group_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dx, const BgGeom g,
                    const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                    const float *__restrict__ save_scale, const float *__restrict__ save_shift, float *__restrict__ dgamma,
                    float *__restrict__ dbeta, const float *__restrict__ join_out, float *__restrict__ join_g) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ float tot[2 * BG_MAX_G];
    const int L = 4 * g.L4;
    float *tg = smem, *tx = smem + (size_t)g.N * L, *part = smem + (size_t)2 * g.N * L;
    const int c0 = blockIdx.x * g.G;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    {
        const float *const src[2] = {dy, x};
        float *const dst[2] = {tg, tx};
        bg_load<2>(src, dst, g, c0, bar);
    }
    const int q = threadIdx.x % g.L4, rg = threadIdx.x / g.L4;
    const bool active = rg < g.RG;
    float mean[4], invstd[4], sc[4], sh[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = c0 + (4 * q + e) / g.HW;
        mean[e] = save_mean[c];
        invstd[e] = save_invstd[c];
        sc[e] = save_scale[c];
        sh[e] = RELU ? save_shift[c] : 0.0f;
    }
    // pass 1: sum(g), sum(g * x_hat) per column (g = dY, masked by the recomputed ReLU when fused)
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    const int n0 = rg * g.rows, n1 = (n0 + g.rows < g.N) ? n0 + g.rows : g.N;
    if (active && join_out != nullptr) {
        // ResidualBlock join folded in (RELU = false): g = dY * (out > 0), `out` streamed from global memory four rows
        // ahead; g replaces dY in the tile (pass 2) and is written out once for the skip path
        for (int nb = n0; nb < n1; nb += 4) {
            float4 m[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (nb + u < n1) m[u] = ld_stream4(join_out + ((long long)(nb + u) * g.C + c0) * g.HW + 4 * q);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int n = nb + u;
                if (n < n1) {
                    float4 gq = *reinterpret_cast<const float4 *>(tg + (size_t)n * L + 4 * q);
                    const float4 t = *reinterpret_cast<const float4 *>(tx + (size_t)n * L + 4 * q);
                    gq.x = m[u].x > 0.f ? gq.x : 0.f; gq.y = m[u].y > 0.f ? gq.y : 0.f;
                    gq.z = m[u].z > 0.f ? gq.z : 0.f; gq.w = m[u].w > 0.f ? gq.w : 0.f;
                    *reinterpret_cast<float4 *>(tg + (size_t)n * L + 4 * q) = gq;
                    st_stream4(join_g + ((long long)n * g.C + c0) * g.HW + 4 * q, gq);
                    const float gv[4] = {gq.x, gq.y, gq.z, gq.w};
                    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        a[e] += gv[e];
                        b[e] = fmaf(gv[e], (tv[e] - mean[e]) * invstd[e], b[e]);
                    }
                }
            }
        }
    } else if (active) {
        for (int n = n0; n < n1; ++n) {
            const float4 gq = *reinterpret_cast<const float4 *>(tg + (size_t)n * L + 4 * q);
            const float4 t = *reinterpret_cast<const float4 *>(tx + (size_t)n * L + 4 * q);
            float gv[4] = {gq.x, gq.y, gq.z, gq.w};
            const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (RELU) gv[e] = fmaf(tv[e], sc[e], sh[e]) > 0.f ? gv[e] : 0.f;
                a[e] += gv[e];
                b[e] = fmaf(gv[e], (tv[e] - mean[e]) * invstd[e], b[e]);
            }
        }
    }
    bg_channel_sums(a, b, active, q, rg, g, part, tot);
    if ((int)threadIdx.x < g.G) {
        dbeta[c0 + threadIdx.x] = tot[threadIdx.x];                 // batch_norm.py:171
        dgamma[c0 + threadIdx.x] = tot[BG_MAX_G + threadIdx.x];     // batch_norm.py:164
    }
    // pass 2: dx = scale * (g - mean(g) - x_hat * mean(g * x_hat))   (batch_norm.py:127-156)
    if (active) {
        const float inv_n = 1.0f / (float)(g.N * g.HW);
        float k1[4], k2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ch = (4 * q + e) / g.HW;
            k1[e] = tot[ch] * inv_n;
            k2[e] = tot[BG_MAX_G + ch] * inv_n;
        }
        for (int n = n0; n < n1; ++n) {
            const float4 gq = *reinterpret_cast<const float4 *>(tg + (size_t)n * L + 4 * q);
            const float4 t = *reinterpret_cast<const float4 *>(tx + (size_t)n * L + 4 * q);
            float gv[4] = {gq.x, gq.y, gq.z, gq.w};
            const float tv[4] = {t.x, t.y, t.z, t.w};
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (RELU) gv[e] = fmaf(tv[e], sc[e], sh[e]) > 0.f ? gv[e] : 0.f;
                r[e] = sc[e] * (gv[e] - k1[e] - ((tv[e] - mean[e]) * invstd[e]) * k2[e]);
            }
            st_stream4(dx + ((long long)n * g.C + c0) * g.HW + 4 * q, make_float4(r[0], r[1], r[2], r[3]));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
static bool g_bg_ready = false;

int bn_group_init() {
    DK_CUDA(cudaFuncSetAttribute(bn_group_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_group_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_group_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_group_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM_MAX));
    g_bg_ready = true;
    return DK_OK;
}

// channel groups whose per-image run is a 16-byte multiple, small enough that all N images of `ntensors` tensors fit;
// only worth it when the grid still covers most of the machine
static bool bg_plan(int N, int C, int HW, int ntensors, BgGeom *g, size_t *smem) {
    if (!g_bg_ready || g_bn_fused_enabled != 1) return false;
    const int G = (HW % 4 == 0) ? 1 : (HW % 2 == 0) ? 2 : 4;
    if (C % G != 0 || N < 1) return false;
    const int L = G * HW, L4 = L / 4;
    if (L4 < 1 || L4 > BG_THREADS) return false;
    int RG = BG_THREADS / L4;
    if (RG > N) RG = N;
    const size_t bytes = ((size_t)ntensors * N * L + (size_t)2 * RG * L) * 4;
    if (bytes > (size_t)110 * 1024) return false;  // at least two CTAs per SM (228 KB, 1 KB reserved per CTA)
    if (C / G < sm_count() / 2) return false;
    g->N = N; g->C = C; g->HW = HW; g->G = G; g->L4 = L4; g->RG = RG; g->rows = (N + RG - 1) / RG;
    *smem = bytes;
    return true;
}

int bn_group_fwd(const float *x, float *y, const BnFinalize &fin, int relu, int N, int C, int HW, cudaStream_t st,
                 const float *add) {
    BgGeom g;
    size_t smem;
    if (fin.mode == 0 || !aligned16(x) || (y != nullptr && !aligned16(y)) || (add != nullptr && (y == nullptr || !aligned16(add))) ||
        !bg_plan(N, C, HW, 1, &g, &smem))
        return DK_ERR_UNSUPPORTED;
    if (relu) bn_group_fwd_kernel<true><<<C / g.G, BG_THREADS, smem, st>>>(x, y, add, g, fin);
    else bn_group_fwd_kernel<false><<<C / g.G, BG_THREADS, smem, st>>>(x, y, add, g, fin);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int bn_group_bwd(const float *dy, const float *x, const float *save_mean, const float *save_invstd, const float *save_scale,
                 const float *save_shift, float *dx, float *dgamma, float *dbeta, int relu, int N, int C, int HW,
                 cudaStream_t st, const float *join_out, float *join_g) {
    BgGeom g;
    size_t smem;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx) || !bg_plan(N, C, HW, 2, &g, &smem)) return DK_ERR_UNSUPPORTED;
    if (join_out != nullptr && (relu || join_g == nullptr || !aligned16(join_out) || !aligned16(join_g))) return DK_ERR_UNSUPPORTED;
    if (relu)
        bn_group_bwd_kernel<true><<<C / g.G, BG_THREADS, smem, st>>>(dy, x, dx, g, save_mean, save_invstd, save_scale, save_shift,
                                                                     dgamma, dbeta, join_out, join_g);
    else
        bn_group_bwd_kernel<false><<<C / g.G, BG_THREADS, smem, st>>>(dy, x, dx, g, save_mean, save_invstd, save_scale,
                                                                      save_shift, dgamma, dbeta, join_out, join_g);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
