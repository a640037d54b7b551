// bn_fused.cu -- BatchNorm forward (statistics + normalise [+ReLU]) and backward (reductions + dX) as ONE kernel each,
// with the channel resident in shared memory between the two passes.
//
// The reference makes ~8 passes over the activation per BatchNorm direction (batch_norm.py:66-74, 118-156); the split
// kernels of batchnorm.cu make the algorithmic minimum of an unfused implementation, 3n forward (read for the
// statistics, read + write for the apply) and 5n backward.  Here a thread-block cluster owns a channel: each of its S
// CTAs pulls its slice of the channel's N*HW values (and of dY, backward) into shared memory with bulk-async copies
// (cp.async.bulk, one per image plane, completion on an mbarrier), reduces it, exchanges the partial sums with the
// other CTAs through distributed shared memory, and then runs the second pass straight out of shared memory:
// 2n forward, 3n backward from HBM.  Everything is fixed-order -> deterministic.
#include <cuda.h>
#include <stdlib.h>

#include "bn.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int BF_THREADS = 256;
constexpr int BF_SMEM_MAX = 200 * 1024;  // dynamic bytes per CTA the slices may use
constexpr int BF_MAX_CLUSTER = 16;          // > 8 is the non-portable cluster size (opt-in per kernel)
constexpr int BF_SMEM_TARGET = 64 * 1024;   // preferred slice bytes per CTA: several CTAs per SM overlap their load / compute / store phases

struct BfGeom {
    int N, C, HW, S;
    int per;   // slice length in elements (multiple of 4 when bulk)
    int bulk;  // planes are 16-byte multiples and aligned: cp.async.bulk; else plain loads
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float ld_dsmem(const float *local, unsigned rank) {
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// slice [v0, v1) of channel c (virtual index v = n*HW + off) of NT tensors -> shared memory, contiguous in v
template <int NT>
__device__ __forceinline__ void bf_load_slice(const float *const (&src)[NT], float *const (&dst)[NT], const BfGeom &g, int c,
                                              int v0, int v1, uint32_t bar) {
    const int len = v1 - v0;
    if (g.bulk) {
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(NT * len) * 4u);
            __syncwarp();
            const int n_first = v0 / g.HW, n_last = (v1 - 1) / g.HW;
            for (int n = n_first + (int)threadIdx.x; n <= n_last; n += 32) {
                const int a = n * g.HW > v0 ? n * g.HW : v0;
                const int b = (n + 1) * g.HW < v1 ? (n + 1) * g.HW : v1;
                const long long goff = ((long long)n * g.C + c) * g.HW + (a - n * g.HW);
#pragma unroll
                for (int t = 0; t < NT; ++t) bulk_g2s(smem_u32(dst[t] + (a - v0)), src[t] + goff, (uint32_t)(b - a) * 4u, bar);
            }
        }
        mbar_wait(bar, 0u);
    } else {
        // plain loads, four elements per thread in flight per tensor before anything is stored
        for (int i0 = threadIdx.x; i0 < len; i0 += 4 * BF_THREADS) {
            float tmp[NT][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * BF_THREADS;
                if (i < len) {
                    const int v = v0 + i;
                    const int n = v / g.HW, off = v - n * g.HW;
                    const long long goff = ((long long)n * g.C + c) * g.HW + off;
#pragma unroll
                    for (int t = 0; t < NT; ++t) tmp[t][u] = __ldg(src[t] + goff);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * BF_THREADS;
                if (i < len) {
#pragma unroll
                    for (int t = 0; t < NT; ++t) dst[t][i] = tmp[t][u];
                }
            }
        }
        __syncthreads();
    }
}

// block-wide sum of two values; result valid in every thread (fixed order)
__device__ __forceinline__ void bf_block_sum2(float &a, float &b, float *red /* >= 2*8+2 floats */) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) {
        red[wid] = a;
        red[8 + wid] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ta = red[0], tb = red[8];
#pragma unroll
        for (int w = 1; w < BF_THREADS / 32; ++w) {
            ta += red[w];
            tb += red[8 + w];
        }
        red[16] = ta;
        red[17] = tb;
    }
    __syncthreads();
    a = red[16];
    b = red[17];
}

// ---- forward ------------------------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(BF_THREADS)
bn_fused_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ add, const BfGeom g,
                    const BnFinalize fin) {
    extern __shared__ __align__(16) float slice[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ float red[18];
    __shared__ float xch[2];  // this CTA's shifted (sum, sum of squares), read by the whole cluster
    __shared__ float ss[2];
    const int c = blockIdx.y;
    const unsigned rank = blockIdx.x;  // cluster = the S CTAs of one channel
    const int total = g.N * g.HW;
    const int v0 = (int)rank * g.per;
    const int v1 = v0 + g.per < total ? v0 + g.per : total;
    const int len = v1 - v0;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    BnChannelParams cp = {1.0f, 0.0f, 0.0f, 0.0f};
    if (threadIdx.x == 0) cp = bn_load_channel_params(fin, c);  // in flight while the slice lands and is reduced
    // One shift for the whole channel (its first value), so that the CTAs' shifted sums simply add up.  The earlier
    // version merged per-slice (n, mean, M2) with Chan's formula: S - 1 dependent divisions on the critical path
    // between the two passes, ~0.5 us each on one warp while the rest of the CTA waits (measured with %globaltimer
    // stamps: 8 us of a 12 us channel at S = 16).
    const float shift = __ldg(x + (long long)c * g.HW);
    {
        const float *const src[1] = {x};
        float *const dst[1] = {slice};
        bf_load_slice<1>(src, dst, g, c, v0, v1, bar);
    }
    // pass 1: shifted sums
    float sum = 0.0f, sq = 0.0f;
    const int len4 = len & ~3;
    for (int i = 4 * threadIdx.x; i < len4; i += 4 * BF_THREADS) {
        const float4 t = *reinterpret_cast<const float4 *>(slice + i);
        const float d0 = t.x - shift, d1 = t.y - shift, d2 = t.z - shift, d3 = t.w - shift;
        sum += (d0 + d1) + (d2 + d3);
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    if ((int)threadIdx.x < len - len4) {
        const float d = slice[len4 + threadIdx.x] - shift;
        sum += d;
        sq += d * d;
    }
    bf_block_sum2(sum, sq, red);
    if (threadIdx.x == 0) {
        xch[0] = sum;
        xch[1] = sq;
    }
    __syncwarp();
    cluster_arrive();
    cluster_wait();
    if (threadIdx.x < 32) {
        // every CTA adds the S partials with the same butterfly (additions commute: bit-identical statistics in every
        // lane and CTA); lane r fetches rank r's partial through distributed shared memory
        const unsigned lane = threadIdx.x;
        float s1 = 0.0f, s2 = 0.0f;
        if (lane < (unsigned)g.S) {
            s1 = ld_dsmem(xch + 0, lane);
            s2 = ld_dsmem(xch + 1, lane);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            const float inv_n = 1.0f / (float)total;
            const float mean = shift + s1 * inv_n;
            const float var = fmaxf(s2 - s1 * s1 * inv_n, 0.0f) * inv_n;  // biased (batch_norm_stats_cy.pyx:44)
            float sc, sh;
            bn_finalize_channel(fin, cp, c, mean, var, rank == 0, &sc, &sh);
            ss[0] = sc;
            ss[1] = sh;
        }
    }
    __syncwarp();
    cluster_arrive();  // peers may exit only after everybody has read their xch (waited for at the end)
    __syncthreads();
    if (y != nullptr) {
        // pass 2 out of shared memory: y = x*scale + shift (+ReLU)
        const float sc = ss[0], sh = ss[1];
        if (add != nullptr) {
            // residual join (host side: only with bulk planes).  The skip values come straight from global memory:
            // four independent 16-byte loads per thread are issued before anything is stored (the streaming stores
            // order later loads behind them, which left ONE load in flight per thread: 13 us for a 51 MB tensor)
            for (int i0 = 4 * threadIdx.x; i0 < len4; i0 += 16 * BF_THREADS) {
                float4 a[4];
                long long goff[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 4 * BF_THREADS;
                    if (i < len4) {
                        const int v = v0 + i;
                        const int n = v / g.HW, off = v - n * g.HW;
                        goff[u] = ((long long)n * g.C + c) * g.HW + off;
                        a[u] = ld_stream4(add + goff[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 4 * BF_THREADS;
                    if (i < len4) {
                        float4 t = *reinterpret_cast<const float4 *>(slice + i);
                        t.x = fmaf(t.x, sc, sh) + a[u].x; t.y = fmaf(t.y, sc, sh) + a[u].y;
                        t.z = fmaf(t.z, sc, sh) + a[u].z; t.w = fmaf(t.w, sc, sh) + a[u].w;
                        if (RELU) {
                            t.x = t.x > 0.f ? t.x : 0.f; t.y = t.y > 0.f ? t.y : 0.f;
                            t.z = t.z > 0.f ? t.z : 0.f; t.w = t.w > 0.f ? t.w : 0.f;
                        }
                        st_stream4(y + goff[u], t);
                    }
                }
            }
        } else
        for (int i = 4 * threadIdx.x; i < len4; i += 4 * BF_THREADS) {
            float4 t = *reinterpret_cast<const float4 *>(slice + i);
            t.x = fmaf(t.x, sc, sh); t.y = fmaf(t.y, sc, sh); t.z = fmaf(t.z, sc, sh); t.w = fmaf(t.w, sc, sh);
            const int v = v0 + i;
            const int n = v / g.HW, off = v - n * g.HW;
            float *o = y + ((long long)n * g.C + c) * g.HW + off;
            if (RELU) {
                t.x = t.x > 0.f ? t.x : 0.f; t.y = t.y > 0.f ? t.y : 0.f;
                t.z = t.z > 0.f ? t.z : 0.f; t.w = t.w > 0.f ? t.w : 0.f;
            }
            if (g.bulk) {
                st_stream4(o, t);  // HW % 4 == 0: the four values share a plane and the address is 16-byte aligned
            } else {
                const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int ve = v + e;
                    const int ne = ve / g.HW, oe = ve - ne * g.HW;
                    y[((long long)ne * g.C + c) * g.HW + oe] = tv[e];
                }
            }
        }
        if ((int)threadIdx.x < len - len4) {
            const int v = v0 + len4 + threadIdx.x;
            float t = fmaf(slice[len4 + threadIdx.x], sc, sh);
            const int n = v / g.HW, off = v - n * g.HW;
            const long long goff = ((long long)n * g.C + c) * g.HW + off;
            if (add != nullptr) t += __ldg(add + goff);
            if (RELU) t = t > 0.f ? t : 0.f;
            y[goff] = t;
        }
    }
    cluster_wait();
}

// ---- backward -----------------------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(BF_THREADS)
bn_fused_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dx, const BfGeom g,
                    const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                    const float *__restrict__ save_scale, const float *__restrict__ save_shift, float *__restrict__ dgamma,
                    float *__restrict__ dbeta, const float *__restrict__ join_out, float *__restrict__ join_g) {
    extern __shared__ __align__(16) float slice[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ float red[18];
    __shared__ float xch[2];
    __shared__ float kk[2];
    const int c = blockIdx.y;
    const unsigned rank = blockIdx.x;
    const int total = g.N * g.HW;
    const int v0 = (int)rank * g.per;
    const int v1 = v0 + g.per < total ? v0 + g.per : total;
    const int len = v1 - v0;
    float *sg = slice, *sx = slice + ((g.per + 3) & ~3);  // 16-byte aligned for the float4 passes
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    {
        const float *const src[2] = {dy, x};
        float *const dst[2] = {sg, sx};
        bf_load_slice<2>(src, dst, g, c, v0, v1, bar);
    }
    const float mean = save_mean[c], invstd = save_invstd[c], sc = save_scale[c];
    const float sh = RELU ? save_shift[c] : 0.0f;
    // pass 1: sum(g), sum(g * x_hat); with a fused ReLU the masked gradient is written back for pass 2
    float a = 0.0f, b = 0.0f;
    const int len4 = len & ~3;
    if (join_out != nullptr) {
        // ResidualBlock join folded in (host side: bulk planes only, RELU = false): the incoming gradient is first masked by
        // the block's ReLU, g = dY * (out > 0) with `out` the block's output streamed from global memory (four loads in
        // flight per thread), kept in the shared-memory slice for pass 2 and written out once for the skip path
        for (int i0 = 4 * threadIdx.x; i0 < len4; i0 += 16 * BF_THREADS) {
            float4 m[4];
            long long goff[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 4 * BF_THREADS;
                if (i < len4) {
                    const int v = v0 + i;
                    const int n = v / g.HW, off = v - n * g.HW;
                    goff[u] = ((long long)n * g.C + c) * g.HW + off;
                    m[u] = ld_stream4(join_out + goff[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 4 * BF_THREADS;
                if (i < len4) {
                    float4 gq = *reinterpret_cast<const float4 *>(sg + i);
                    const float4 t = *reinterpret_cast<const float4 *>(sx + i);
                    gq.x = m[u].x > 0.f ? gq.x : 0.f; gq.y = m[u].y > 0.f ? gq.y : 0.f;
                    gq.z = m[u].z > 0.f ? gq.z : 0.f; gq.w = m[u].w > 0.f ? gq.w : 0.f;
                    *reinterpret_cast<float4 *>(sg + i) = gq;
                    st_stream4(join_g + goff[u], gq);
                    a += (gq.x + gq.y) + (gq.z + gq.w);
                    b += (gq.x * ((t.x - mean) * invstd) + gq.y * ((t.y - mean) * invstd)) +
                         (gq.z * ((t.z - mean) * invstd) + gq.w * ((t.w - mean) * invstd));
                }
            }
        }
        if ((int)threadIdx.x < len - len4) {
            const int i = len4 + threadIdx.x;
            const int v = v0 + i;
            const int n = v / g.HW, off = v - n * g.HW;
            const long long goff = ((long long)n * g.C + c) * g.HW + off;
            const float gv = __ldg(join_out + goff) > 0.f ? sg[i] : 0.f;
            sg[i] = gv;
            join_g[goff] = gv;
            a += gv;
            b += gv * ((sx[i] - mean) * invstd);
        }
    } else {
    for (int i = 4 * threadIdx.x; i < len4; i += 4 * BF_THREADS) {
        float4 gq = *reinterpret_cast<const float4 *>(sg + i);
        const float4 t = *reinterpret_cast<const float4 *>(sx + i);
        if (RELU) {
            gq.x = fmaf(t.x, sc, sh) > 0.f ? gq.x : 0.f; gq.y = fmaf(t.y, sc, sh) > 0.f ? gq.y : 0.f;
            gq.z = fmaf(t.z, sc, sh) > 0.f ? gq.z : 0.f; gq.w = fmaf(t.w, sc, sh) > 0.f ? gq.w : 0.f;
            *reinterpret_cast<float4 *>(sg + i) = gq;
        }
        a += (gq.x + gq.y) + (gq.z + gq.w);
        b += (gq.x * ((t.x - mean) * invstd) + gq.y * ((t.y - mean) * invstd)) +
             (gq.z * ((t.z - mean) * invstd) + gq.w * ((t.w - mean) * invstd));
    }
    if ((int)threadIdx.x < len - len4) {
        const int i = len4 + threadIdx.x;
        float gv = sg[i];
        const float t = sx[i];
        if (RELU) {
            gv = fmaf(t, sc, sh) > 0.f ? gv : 0.f;
            sg[i] = gv;
        }
        a += gv;
        b += gv * ((t - mean) * invstd);
    }
    }
    bf_block_sum2(a, b, red);
    if (threadIdx.x == 0) {
        xch[0] = a;
        xch[1] = b;
    }
    __syncwarp();
    cluster_arrive();
    cluster_wait();
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        float pa = 0.0f, pb = 0.0f;
        if (lane < (unsigned)g.S) {
            pa = ld_dsmem(xch + 0, lane);
            pb = ld_dsmem(xch + 1, lane);
        }
        float ta = pa, tb = pb;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // fixed butterfly; additions commute -> same bits in every lane and CTA
            ta += __shfl_xor_sync(0xffffffffu, ta, o);
            tb += __shfl_xor_sync(0xffffffffu, tb, o);
        }
        if (lane == 0) {
            if (rank == 0) {
                dbeta[c] = ta;   // batch_norm.py:171
                dgamma[c] = tb;  // batch_norm.py:164
            }
            const float inv_n = 1.0f / (float)total;
            kk[0] = ta * inv_n;
            kk[1] = tb * inv_n;
        }
    }
    __syncwarp();
    cluster_arrive();
    __syncthreads();
    // pass 2 out of shared memory: dx = scale * (g - mean(g) - x_hat * mean(g * x_hat))   (batch_norm.py:127-156)
    const float k1 = kk[0], k2 = kk[1];
    for (int i = 4 * threadIdx.x; i < len4; i += 4 * BF_THREADS) {
        const float4 gq = *reinterpret_cast<const float4 *>(sg + i);
        const float4 t = *reinterpret_cast<const float4 *>(sx + i);
        float4 r;
        r.x = sc * (gq.x - k1 - ((t.x - mean) * invstd) * k2);
        r.y = sc * (gq.y - k1 - ((t.y - mean) * invstd) * k2);
        r.z = sc * (gq.z - k1 - ((t.z - mean) * invstd) * k2);
        r.w = sc * (gq.w - k1 - ((t.w - mean) * invstd) * k2);
        const int v = v0 + i;
        const int n = v / g.HW, off = v - n * g.HW;
        float *o = dx + ((long long)n * g.C + c) * g.HW + off;
        if (g.bulk) {
            st_stream4(o, r);
        } else {
            const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ve = v + e;
                const int ne = ve / g.HW, oe = ve - ne * g.HW;
                dx[((long long)ne * g.C + c) * g.HW + oe] = rv[e];
            }
        }
    }
    if ((int)threadIdx.x < len - len4) {
        const int i = len4 + threadIdx.x;
        const int v = v0 + i;
        const int n = v / g.HW, off = v - n * g.HW;
        dx[((long long)n * g.C + c) * g.HW + off] = sc * (sg[i] - k1 - ((sx[i] - mean) * invstd) * k2);
    }
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------ host
static bool g_bf_ready = false;
static int g_bf_max_cluster = 8;
int g_bn_fused_enabled = 1;

int bn_fused_init() {
    DK_CUDA(cudaFuncSetAttribute(bn_fused_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_fused_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(bn_fused_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BF_SMEM_MAX));
    // clusters of 16 CTAs are opt-in; keep 8 if the device refuses
    bool np_ok = true;
    np_ok &= cudaFuncSetAttribute(bn_fused_fwd_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    np_ok &= cudaFuncSetAttribute(bn_fused_fwd_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    np_ok &= cudaFuncSetAttribute(bn_fused_bwd_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    np_ok &= cudaFuncSetAttribute(bn_fused_bwd_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (np_ok) {
        // make sure a 16-CTA cluster with the largest slices we would ask for can be co-scheduled
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16, 1, 1);
        cfg.blockDim = dim3(BF_THREADS, 1, 1);
        cfg.dynamicSmemBytes = 100 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 16;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, bn_fused_bwd_kernel<false>, &cfg) == cudaSuccess && nclusters > 0)
            g_bf_max_cluster = 16;
    }
    cudaGetLastError();
    if (const char *e = getenv("DK_BN_MAX_CLUSTER")) g_bf_max_cluster = atoi(e) >= 16 ? 16 : atoi(e) >= 8 ? 8 : atoi(e) >= 4 ? 4 : atoi(e) >= 2 ? 2 : 1;
    g_bf_ready = true;
    return bn_group_init();
}

// cluster size and slice length: enough CTAs to fill the machine twice, slices of at most BF_SMEM_TARGET bytes when a
// cluster of <= 16 CTAs allows it (else up to BF_SMEM_MAX), no empty trailing CTA
static bool bf_plan(int N, int C, int HW, int ntensors, bool bulk, BfGeom *g) {
    if (!g_bf_ready || !g_bn_fused_enabled || C > 65535) return false;
    const int64_t total = (int64_t)N * HW;
    const int64_t cap = BF_SMEM_MAX / (4 * ntensors) - 4, want = BF_SMEM_TARGET / (4 * ntensors);
    const int64_t unit = bulk ? 4 : 1;
    auto per_of = [&](int s) { return ceil_div(ceil_div(total, s), unit) * unit; };
    int s = 1;
    while (s < BF_MAX_CLUSTER && ((int64_t)C * s < 2 * sm_count() || per_of(s) > want)) s *= 2;
    if (s > g_bf_max_cluster) s = g_bf_max_cluster;
    while (s > 1 && (int64_t)(s - 1) * per_of(s) >= total) s /= 2;
    const int64_t per = per_of(s);
    if (per > cap || (int64_t)(s - 1) * per >= total) return false;
    g->N = N; g->C = C; g->HW = HW; g->S = s; g->per = (int)per; g->bulk = bulk ? 1 : 0;
    return true;
}

template <class... Exp, class... Act>
static int bf_launch(void (*kernel)(Exp...), const BfGeom &g, size_t smem, cudaStream_t st, Act &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)g.S, (unsigned)g.C, 1);
    cfg.blockDim = dim3(BF_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DK_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...));
    count_launch();
    return DK_OK;
}

int bn_fused_fwd(const float *x, float *y, const BnFinalize &fin, int relu, int N, int C, int HW, cudaStream_t st,
                 const float *add) {
    {
        const int rc = bn_group_fwd(x, y, fin, relu, N, C, HW, st, add);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    const bool bulk = (HW % 4 == 0) && aligned16(x) && (y == nullptr || aligned16(y));
    BfGeom g;
    if (add != nullptr && (!bulk || y == nullptr || !aligned16(add))) return DK_ERR_UNSUPPORTED;
    if (!bf_plan(N, C, HW, 1, bulk, &g)) return DK_ERR_UNSUPPORTED;
    const size_t smem = (size_t)g.per * 4;
    if (relu) return bf_launch(bn_fused_fwd_kernel<true>, g, smem, st, x, y, add, g, fin);
    return bf_launch(bn_fused_fwd_kernel<false>, g, smem, st, x, y, add, g, fin);
}

int bn_fused_bwd(const float *dy, const float *x, const float *save_mean, const float *save_invstd, const float *save_scale,
                 const float *save_shift, float *dx, float *dgamma, float *dbeta, int relu, int N, int C, int HW,
                 cudaStream_t st, const float *join_out, float *join_g) {
    if (join_out != nullptr && (relu || join_g == nullptr)) return DK_ERR_UNSUPPORTED;
    {
        const int rc = bn_group_bwd(dy, x, save_mean, save_invstd, save_scale, save_shift, dx, dgamma, dbeta, relu, N, C, HW, st,
                                    join_out, join_g);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    const bool bulk = (HW % 4 == 0) && aligned16(x) && aligned16(dy) && aligned16(dx);
    if (join_out != nullptr && (!bulk || !aligned16(join_out) || !aligned16(join_g))) return DK_ERR_UNSUPPORTED;
    BfGeom g;
    if (!bf_plan(N, C, HW, 2, bulk, &g)) return DK_ERR_UNSUPPORTED;
    const size_t smem = (size_t)((g.per + 3) & ~3) * 8;
    if (relu)
        return bf_launch(bn_fused_bwd_kernel<true>, g, smem, st, dy, x, dx, g, save_mean, save_invstd, save_scale, save_shift,
                         dgamma, dbeta, join_out, join_g);
    return bf_launch(bn_fused_bwd_kernel<false>, g, smem, st, dy, x, dx, g, save_mean, save_invstd, save_scale, save_shift,
                     dgamma, dbeta, join_out, join_g);
}

}  // namespace dk
