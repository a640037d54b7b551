// depthwise_group.cu -- 3x3 / stride 1 / pad 1 depthwise convolution forward and backward for SMALL square planes (7x7,
// 14x14: the last two stages of ResNet-18-depsep), a channel GROUP per CTA with all N images resident in shared memory.
//
// Same idea as bn_group.cu: a 7x7 plane (196 B) is not bulk-copyable, but G = 4 consecutive channels of one image are
// one 16-byte aligned run of G*H*W floats (14x14: G = 1), so the CTA's whole working set -- a [N][G*H*W] tile per
// tensor, 50 KB at batch 64 -- arrives with N cp.async.bulk copies per tensor and every window read comes from shared
// memory.  A thread works on ONE channel (its 9 taps and its 10 dW/db sums stay in registers) and takes output rows
// r = (n, h) of that channel round-robin, so consecutive lanes read consecutive plane rows (stride PW words:
// conflict-free at 7, 2-way at 14).  dW / db are reduced over the channel's threads in a fixed order inside the CTA: no
// partial buffer, no reduce kernel, deterministic.  The register-window kernels these replace were latency-bound at
// these sizes (7x7: fwd 16 us, bwd 25.7 us for 6.4 MB tensors).
// Semantics: depthwise_convolution.py:72-83,186-196 / im2col.pyx:109-178 (cross-correlation, zero padding).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int DG_THREADS = 256;
constexpr int DG_MAX_G = 4;

struct DgGeom {
    int N, C, H, G;
    int L;    // floats per tile row (G*H*PW)
    int TPC;  // threads per channel (DG_THREADS / G)
};

__device__ __forceinline__ void dg_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NT>
__device__ __forceinline__ void dg_load(const float *const (&src)[NT], float *const (&dst)[NT], const DgGeom &g, int c0,
                                        uint32_t bar) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, (uint32_t)(NT * g.N * g.L) * 4u);
    __syncthreads();
    for (int n = threadIdx.x; n < g.N; n += DG_THREADS) {
        const long long goff = ((long long)n * g.C + c0) * (g.L / g.G);
#pragma unroll
        for (int t = 0; t < NT; ++t) dg_bulk_g2s(smem_u32(dst[t] + (size_t)n * g.L), src[t] + goff, (uint32_t)g.L * 4u, bar);
    }
    mbar_wait(bar, 0u);
}

// rows h-1, h, h+1 of a plane (zero outside), with one zero column on each side
template <int PW>
__device__ __forceinline__ void dg_rows3(const float *plane, int h, int H, float (&r)[3][PW + 2]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hh = h + i - 1;
        const bool ok = hh >= 0 && hh < H;
        const float *p = plane + hh * PW;
        r[i][0] = 0.0f;
        r[i][PW + 1] = 0.0f;
#pragma unroll
        for (int w = 0; w < PW; ++w) r[i][w + 1] = ok ? p[w] : 0.0f;
    }
}

// tile (results written back in place) -> global, 16-byte words, coalesced; optional "+ add" (residual join gradient)
__device__ __forceinline__ void dg_store_tile(const float *tile, float *out, const float *add, const DgGeom &g, int c0) {
    const int L4 = g.L / 4, HWG = g.L;
    for (int idx = threadIdx.x; idx < g.N * L4; idx += DG_THREADS) {
        const int n = idx / L4, q = idx - n * L4;
        float4 v = *reinterpret_cast<const float4 *>(tile + (size_t)n * HWG + 4 * q);
        const long long goff = ((long long)n * g.C + c0) * (g.L / g.G) + 4 * q;
        if (add != nullptr) {
            const float4 a = ld_stream4(add + goff);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
        }
        st_stream4(out + goff, v);
    }
}

// ITEMS = output rows per thread (compile time: the results wait in registers until every thread has read its windows,
// then overwrite the tile, which leaves as coalesced 16-byte stores)
template <int PW, int ITEMS>
__global__ void __launch_bounds__(DG_THREADS, 2)
dw_group_fwd_kernel(const float *__restrict__ x, const float *__restrict__ wt, const float *__restrict__ bias,
                    float *__restrict__ y, const DgGeom g) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    float *tile = smem;
    const int c0 = blockIdx.x * g.G;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    const int ch = threadIdx.x / g.TPC, u = threadIdx.x - ch * g.TPC;
    const int c = c0 + ch;
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = __ldg(wt + (long long)c * 9 + t);
    const float b = bias ? __ldg(bias + c) : 0.0f;
    __syncthreads();
    {
        const float *const src[1] = {x};
        float *const dst[1] = {tile};
        dg_load<1>(src, dst, g, c0, bar);
    }
    const int HW = g.H * PW;
    const int items = g.N * g.H;
    float out[ITEMS][PW];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int r = u + it * g.TPC;
        if (r < items) {
            const int n = r / g.H, h = r - n * g.H;
            float in[3][PW + 2];
            dg_rows3<PW>(tile + (size_t)n * g.L + ch * HW, h, g.H, in);
#pragma unroll
            for (int w = 0; w < PW; ++w) {
                float acc = b;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc = fmaf(in[i][w + j], k[i * 3 + j], acc);
                out[it][w] = acc;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int r = u + it * g.TPC;
        if (r < items) {
            const int n = r / g.H, h = r - n * g.H;
            float *o = tile + (size_t)n * g.L + ch * HW + h * PW;
#pragma unroll
            for (int w = 0; w < PW; ++w) o[w] = out[it][w];
        }
    }
    __syncthreads();
    dg_store_tile(tile, y, nullptr, g, c0);
}

template <int PW, int ITEMS>
__global__ void __launch_bounds__(DG_THREADS, 2)
dw_group_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ wt,
                    float *__restrict__ dx, const float *__restrict__ dx_add, float *__restrict__ dw,
                    float *__restrict__ dbias, float l2, const DgGeom g) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ float red[DG_THREADS / 32][10];
    float *tg = smem, *tx = smem + (size_t)g.N * g.L;
    const int c0 = blockIdx.x * g.G;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    const int ch = threadIdx.x / g.TPC, u = threadIdx.x - ch * g.TPC;
    const int c = c0 + ch;
    float k[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) k[t] = __ldg(wt + (long long)c * 9 + t);
    __syncthreads();
    {
        const float *const src[2] = {dy, x};
        float *const dst[2] = {tg, tx};
        dg_load<2>(src, dst, g, c0, bar);
    }
    const int HW = g.H * PW;
    const int items = g.N * g.H;
    float sw[9], sb = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) sw[t] = 0.0f;
    float out[ITEMS][PW];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int r = u + it * g.TPC;
        if (r < items) {
            const int n = r / g.H, h = r - n * g.H;
            float gy[3][PW + 2];
            dg_rows3<PW>(tg + (size_t)n * g.L + ch * HW, h, g.H, gy);
            // dX[h][w] = sum_ij dY[h+1-i][w+1-j] * K[i][j]   (im2col.pyx:159-178, stride 1, pad 1)
#pragma unroll
            for (int w = 0; w < PW; ++w) {
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc = fmaf(gy[2 - i][w + 2 - j], k[i * 3 + j], acc);
                out[it][w] = acc;
            }
            // dW[i][j] += sum_w dY[h][w] * X[h+i-1][w+j-1];  db += sum_w dY[h][w]
            float xi[3][PW + 2];
            dg_rows3<PW>(tx + (size_t)n * g.L + ch * HW, h, g.H, xi);
#pragma unroll
            for (int w = 0; w < PW; ++w) {
                const float gv = gy[1][w + 1];
                sb += gv;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) sw[i * 3 + j] = fmaf(gv, xi[i][w + j], sw[i * 3 + j]);
            }
        }
    }
    // fixed-order reduction over the channel's threads: warp shuffles, then the channel's warps in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < 9; ++t) sw[t] = warp_sum(sw[t]);
    sb = warp_sum(sb);
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < 9; ++t) red[warp][t] = sw[t];
        red[warp][9] = sb;
    }
    __syncthreads();  // also: every thread is done reading the dY tile, which now takes dX
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int r = u + it * g.TPC;
        if (r < items) {
            const int n = r / g.H, h = r - n * g.H;
            float *o = tg + (size_t)n * g.L + ch * HW + h * PW;
#pragma unroll
            for (int w = 0; w < PW; ++w) o[w] = out[it][w];
        }
    }
    const int wpc = g.TPC / 32;  // warps per channel
    if (u < 10) {
        float s = 0.0f;
        for (int wi = 0; wi < wpc; ++wi) s += red[ch * wpc + wi][u];
        if (u < 9) dw[(long long)c * 9 + u] = l2 != 0.0f ? fmaf(l2, __ldg(wt + (long long)c * 9 + u), s) : s;
        else if (dbias != nullptr) dbias[c] = s;
    }
    __syncthreads();
    dg_store_tile(tg, dx, dx_add, g, c0);
}

// ------------------------------------------------------------------------------------------------ host
static bool g_dg_ready = false;
int g_dw_group_enabled = 1;

int init_dw_group() {
    const int smem = 200 * 1024;
    DK_CUDA(cudaFuncSetAttribute(dw_group_fwd_kernel<7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DK_CUDA(cudaFuncSetAttribute(dw_group_fwd_kernel<14, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DK_CUDA(cudaFuncSetAttribute(dw_group_bwd_kernel<7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DK_CUDA(cudaFuncSetAttribute(dw_group_bwd_kernel<14, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_dg_ready = true;
    return DK_OK;
}

static bool dg_plan(DgGeom &g, size_t *smem, int N, int C, int H, int W, int kh, int kw, int s, int p, int ntensors) {
    if (!g_dg_ready || !g_dw_group_enabled || kh != 3 || kw != 3 || s != 1 || p != 1 || H != W || (W != 7 && W != 14)) return false;
    const int HW = H * W;
    const int G = (HW % 4 == 0) ? 1 : (HW % 2 == 0) ? 2 : 4;
    if (C % G != 0 || C / G < sm_count() / 2) return false;
    const size_t bytes = (size_t)ntensors * N * G * HW * 4;
    if (bytes > (size_t)110 * 1024) return false;  // two CTAs per SM
    g.N = N; g.C = C; g.H = H; g.G = G; g.L = G * HW; g.TPC = DG_THREADS / G;
    const int items_per_thread = (N * H + g.TPC - 1) / g.TPC;
    if (items_per_thread > (W == 7 ? 8 : 4)) return false;  // the kernels' compile-time ITEMS
    *smem = bytes;
    return true;
}

int dw_group_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                 int s, int p, cudaStream_t st) {
    DgGeom g;
    size_t smem;
    if (!aligned16(x) || !aligned16(y) || !dg_plan(g, &smem, N, C, H, W, kh, kw, s, p, 1)) return DK_ERR_UNSUPPORTED;
    if (W == 7) dw_group_fwd_kernel<7, 8><<<C / g.G, DG_THREADS, smem, st>>>(x, w, bias, y, g);
    else dw_group_fwd_kernel<14, 4><<<C / g.G, DG_THREADS, smem, st>>>(x, w, bias, y, g);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dw_group_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                 float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st) {
    DgGeom g;
    size_t smem;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(dx) || (dx_add != nullptr && !aligned16(dx_add)) ||
        !dg_plan(g, &smem, N, C, H, W, kh, kw, s, p, 2))
        return DK_ERR_UNSUPPORTED;
    if (W == 7) dw_group_bwd_kernel<7, 8><<<C / g.G, DG_THREADS, smem, st>>>(dy, x, w, dx, dx_add, dw, dbias, l2, g);
    else dw_group_bwd_kernel<14, 4><<<C / g.G, DG_THREADS, smem, st>>>(dy, x, w, dx, dx_add, dw, dbias, l2, g);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
