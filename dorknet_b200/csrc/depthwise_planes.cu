// depthwise_planes.cu -- depthwise 3x3 / pad 1 / stride 1 or 2 with the image planes resident in shared memory.
//
// A CTA takes a group of G consecutive (n, c) planes -- contiguous in NCHW memory -- and pulls them into shared memory
// with ONE bulk-async copy per tensor (cp.async.bulk + mbarrier: no registers, no per-thread address arithmetic, the
// whole group in flight at once; several CTAs per SM overlap their load / compute / store phases).  All window reads
// then come from shared memory, so the 3x3 reuse never touches L1/L2 again and the zero padding is index arithmetic
// (the reference pads a copy of the input: depthwise_convolution.py:72-83).
//
// Forward (im2col.pyx:109-139): y = sum_{i,j} xpad[.., oh*s+i, ow*s+j] * w[c,i,j] (+ b).
// Backward (im2col.pyx:143-178, one pass over dY and X like the reference's fused routine): dX (+ the residual
// gradient dx_add), and per-plane partial sums of dW / db which dw_planes_reduce_kernel adds over the batch in a fixed
// order (the reference sums its per-image dW: depthwise_convolution.py:193) and tops up with l2*W.
//
// Work split inside a CTA: 8 warps; with G >= 8 planes a warp owns whole planes (g = warp, warp+8, ...), with fewer
// planes 8/G warps share one.  A lane's work item is VEC adjacent pixels of one row.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dk {

using namespace tc;

constexpr int DP_THREADS = 256;
constexpr int DP_WARPS = DP_THREADS / 32;
constexpr int DP_SMEM_MAX = 100 * 1024;
constexpr int DP_TARGET_BYTES = 12 * 1024;  // preferred bytes per tensor and CTA

struct DpGeom {
    int C, H, W, OH, OW, s;
    int G;             // planes per CTA (power of two <= 8, or a multiple of 8)
    long long planes;  // N*C
    int bulk;
};

__device__ __forceinline__ void dp_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// cnt consecutive planes of `ps` floats starting at plane p0 of NT tensors -> shared memory
template <int NT>
__device__ __forceinline__ void dp_load(const float *const (&src)[NT], float *const (&dst)[NT], const int (&ps)[NT],
                                        long long p0, int cnt, int bulk, uint32_t bar) {
    if (bulk) {
        if (threadIdx.x == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int t = 0; t < NT; ++t) total += (uint32_t)(cnt * ps[t]) * 4u;
            mbar_expect_tx(bar, total);
#pragma unroll
            for (int t = 0; t < NT; ++t)
                dp_bulk_g2s(smem_u32(dst[t]), src[t] + p0 * ps[t], (uint32_t)(cnt * ps[t]) * 4u, bar);
        }
        mbar_wait(bar, 0u);
    } else {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const float *s = src[t] + p0 * ps[t];
            const int n = cnt * ps[t];
            for (int i = threadIdx.x; i < n; i += DP_THREADS) dst[t][i] = __ldg(s + i);
        }
        __syncthreads();
    }
}

// r[0..VEC+1] = row[w0-1 .. w0+VEC] with zeros outside [0, W); row == nullptr: a padding row
template <int VEC>
__device__ __forceinline__ void dp_row_s1(const float *row, int w0, int W, float (&r)[VEC + 2]) {
    if (row == nullptr) {
#pragma unroll
        for (int i = 0; i < VEC + 2; ++i) r[i] = 0.0f;
        return;
    }
    r[0] = w0 > 0 ? row[w0 - 1] : 0.0f;
    if (VEC == 4) {
        const float4 q = *reinterpret_cast<const float4 *>(row + w0);
        r[1] = q.x; r[2] = q.y; r[VEC > 2 ? 3 : 0] = q.z; r[VEC > 3 ? 4 : 0] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) r[1 + i] = row[w0 + i];
    }
    r[VEC + 1] = w0 + VEC < W ? row[w0 + VEC] : 0.0f;
}

// r[0..2*VEC] = row[2*o0-1 .. 2*o0+2*VEC-1] (the columns VEC stride-2 outputs starting at o0 read), zeros outside
template <int VEC>
__device__ __forceinline__ void dp_row_s2(const float *row, int o0, int W, float (&r)[2 * VEC + 1]) {
    if (row == nullptr) {
#pragma unroll
        for (int i = 0; i < 2 * VEC + 1; ++i) r[i] = 0.0f;
        return;
    }
    const int c0 = 2 * o0;
    r[0] = c0 > 0 ? row[c0 - 1] : 0.0f;
    if (VEC == 4 && c0 + 8 <= W) {
        const float4 a = *reinterpret_cast<const float4 *>(row + c0), b = *reinterpret_cast<const float4 *>(row + c0 + 4);
        r[1] = a.x; r[2] = a.y; r[VEC > 1 ? 3 : 0] = a.z; r[VEC > 1 ? 4 : 0] = a.w;
        r[VEC > 2 ? 5 : 0] = b.x; r[VEC > 2 ? 6 : 0] = b.y; r[VEC > 3 ? 7 : 0] = b.z; r[VEC > 3 ? 8 : 0] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 2 * VEC; ++i) r[1 + i] = c0 + i < W ? row[c0 + i] : 0.0f;
    }
}

template <int VEC>
__device__ __forceinline__ void dp_store(float *p, const float (&v)[VEC]) {
    if (VEC == 4) st_stream4(p, make_float4(v[0], v[1], v[VEC > 2 ? 2 : 0], v[VEC > 3 ? 3 : 0]));
    else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = v[i];
    }
}

// ---------------------------------------------------------------------------------------------- forward
template <int S, int VEC>
__global__ void __launch_bounds__(DP_THREADS)
dw_planes_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                     float *__restrict__ y, const DpGeom g) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar_mem;
    const long long p0 = (long long)blockIdx.x * g.G;
    const int cnt = (int)(g.planes - p0 < g.G ? g.planes - p0 : g.G);
    const int HW = g.H * g.W, OHW = g.OH * g.OW;
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    {
        const float *const src[1] = {x};
        float *const dst[1] = {sm};
        const int ps[1] = {HW};
        dp_load<1>(src, dst, ps, p0, cnt, g.bulk, bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpp = g.G >= DP_WARPS ? 1 : DP_WARPS / g.G;  // warps per plane
    const int SP = g.OW / VEC, items = g.OH * SP;
    for (int gl = warp / wpp; gl < cnt; gl += DP_WARPS / wpp) {
        const int c = (int)((p0 + gl) % g.C);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
        const float bv = bias ? __ldg(bias + c) : 0.0f;
        const float *xp = sm + (size_t)gl * HW;
        float *yp = y + (p0 + gl) * OHW;
        for (int t = (warp % wpp) * 32 + lane; t < items; t += wpp * 32) {
            const int oh = t / SP, o0 = (t - oh * SP) * VEC;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] = bv;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ih = oh * S - 1 + i;
                const float *row = (ih >= 0 && ih < g.H) ? xp + ih * g.W : nullptr;
                if (S == 1) {
                    float r[VEC + 2];
                    dp_row_s1<VEC>(row, o0, g.W, r);
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
#pragma unroll
                        for (int j = 0; j < 3; ++j) o[v] = fmaf(r[v + j], k[i * 3 + j], o[v]);
                } else {
                    float r[2 * VEC + 1];
                    dp_row_s2<VEC>(row, o0, g.W, r);
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
#pragma unroll
                        for (int j = 0; j < 3; ++j) o[v] = fmaf(r[2 * v + j], k[i * 3 + j], o[v]);
                }
            }
            dp_store<VEC>(yp + oh * g.OW + o0, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward
// partial[plane][10]: dW (9) and db of one plane
template <int S, int VEC>
__global__ void __launch_bounds__(DP_THREADS)
dw_planes_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ w,
                     float *__restrict__ dx, float *__restrict__ partial, const float *__restrict__ dx_add,
                     const DpGeom g) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar_mem;
    const long long p0 = (long long)blockIdx.x * g.G;
    const int cnt = (int)(g.planes - p0 < g.G ? g.planes - p0 : g.G);
    const int HW = g.H * g.W, OHW = g.OH * g.OW;
    float *sg = sm;                                          // dY planes [G][OH*OW]
    float *sx = sm + (((size_t)g.G * OHW + 3) & ~(size_t)3);  // X planes  [G][H*W]
    float *psum = sx + (((size_t)g.G * HW + 3) & ~(size_t)3); // [G][wpp][10]
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    {
        const float *const src[2] = {dy, x};
        float *const dst[2] = {sg, sx};
        const int ps[2] = {OHW, HW};
        dp_load<2>(src, dst, ps, p0, cnt, g.bulk, bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpp = g.G >= DP_WARPS ? 1 : DP_WARPS / g.G;
    const int sub = warp % wpp;
    for (int gl = warp / wpp; gl < cnt; gl += DP_WARPS / wpp) {
        const int c = (int)((p0 + gl) % g.C);
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
        const float *gp = sg + (size_t)gl * OHW;
        const float *xp = sx + (size_t)gl * HW;
        float *dxp = dx + (p0 + gl) * HW;
        const float *addp = dx_add ? dx_add + (p0 + gl) * HW : nullptr;
        float acc[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = 0.0f;
        if (S == 1) {
            // dY and X share the geometry: one item yields VEC values of dX and feeds the dW sums
            const int SP = g.W / VEC, items = g.H * SP;
            for (int t = sub * 32 + lane; t < items; t += wpp * 32) {
                const int h = t / SP, w0 = (t - h * SP) * VEC;
                float rg[3][VEC + 2], rx[3][VEC + 2];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int hh = h - 1 + a;
                    const bool ok = hh >= 0 && hh < g.H;
                    dp_row_s1<VEC>(ok ? gp + hh * g.W : nullptr, w0, g.W, rg[a]);
                    dp_row_s1<VEC>(ok ? xp + hh * g.W : nullptr, w0, g.W, rx[a]);
                }
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    // dX[h][w] = sum_{a,b} dY[h+a-1][w+b-1] * k[2-a][2-b]
                    float s = 0.0f;
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) s = fmaf(rg[a][v + b], k[(2 - a) * 3 + (2 - b)], s);
                    o[v] = s;
                    const float gv = rg[1][v + 1];
                    acc[9] += gv;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fmaf(gv, rx[i][v + j], acc[i * 3 + j]);
                }
                if (addp) {
                    if (VEC == 4) {
                        const float4 q = ld_stream4(addp + h * g.W + w0);
                        o[0] += q.x; o[VEC > 1 ? 1 : 0] += q.y; o[VEC > 2 ? 2 : 0] += q.z; o[VEC > 3 ? 3 : 0] += q.w;
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) o[v] += __ldg(addp + h * g.W + w0 + v);
                    }
                }
                dp_store<VEC>(dxp + h * g.W + w0, o);
            }
        } else {
            // (A) dW / db over the dY pixels
            {
                const int items = g.OH * g.OW;
                for (int t = sub * 32 + lane; t < items; t += wpp * 32) {
                    const int oh = t / g.OW, ow = t - oh * g.OW;
                    const float gv = gp[t];
                    acc[9] += gv;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int ih = 2 * oh - 1 + i;
                        float r[3];
                        dp_row_s2<1>((ih >= 0 && ih < g.H) ? xp + ih * g.W : nullptr, ow, g.W, r);
#pragma unroll
                        for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fmaf(gv, r[j], acc[i * 3 + j]);
                    }
                }
            }
            // (B) dX[h][w] = sum over the taps (i, j) whose output pixel ((h+1-i)/2, (w+1-j)/2) is integral and exists
            {
                const int SP = g.W / VEC, items = g.H * SP;
                for (int t = sub * 32 + lane; t < items; t += wpp * 32) {
                    const int h = t / SP, w0 = (t - h * SP) * VEC;
                    float o[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[v] = 0.0f;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int th = h + 1 - i;
                        if (th < 0 || (th & 1) || (th >> 1) >= g.OH) continue;
                        const float *grow = gp + (th >> 1) * g.OW;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const int tw = w0 + v + 1 - j;
                                if (tw >= 0 && !(tw & 1) && (tw >> 1) < g.OW) o[v] = fmaf(grow[tw >> 1], k[i * 3 + j], o[v]);
                            }
                        }
                    }
                    if (addp) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) o[v] += __ldg(addp + h * g.W + w0 + v);
                    }
                    dp_store<VEC>(dxp + h * g.W + w0, o);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) acc[i] = warp_sum(acc[i]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 10; ++i) psum[((size_t)gl * wpp + sub) * 10 + i] = acc[i];
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 10; t += DP_THREADS) {
        const int gl = t / 10, i = t - gl * 10;
        float s = 0.0f;
        for (int q = 0; q < wpp; ++q) s += psum[((size_t)gl * wpp + q) * 10 + i];
        partial[(p0 + gl) * 10 + i] = s;
    }
}

// dW[c][t] = sum_n partial[n*C + c][t] (+ l2*W), db[c] = sum_n partial[n*C + c][9]; fixed order
__global__ void __launch_bounds__(128)
dw_planes_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ dw,
                        float *__restrict__ dbias, float l2, int N, int C) {
    __shared__ float red[4][10];
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = 0.0f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *src = partial + ((long long)n * C + c) * 10;
#pragma unroll
        for (int t = 0; t < 10; ++t) acc[t] += src[t];
    }
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = warp_sum(acc[t]);
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < 10; ++t) red[wid][t] = acc[t];
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        const int t = threadIdx.x;
        const float s = (red[0][t] + red[1][t]) + (red[2][t] + red[3][t]);
        if (t < 9) dw[c * 9 + t] = s + (l2 != 0.0f ? l2 * w[c * 9 + t] : 0.0f);
        else if (dbias) dbias[c] = s;
    }
}


// ============================================================================================== per-channel backward
// dw_chan_bwd_kernel: stride 1.  A CTA owns ONE channel c and a range of images; it streams the (dY, X) planes of that
// channel through a ring of shared-memory stages filled by cp.async.bulk (P images per stage, `nst` stages in flight:
// the loads need no registers, so ~100 KB per CTA are on their way while the warps compute), and keeps the 9 dW sums
// and db in registers across ALL its planes -- one block reduction at the very end instead of one per plane, and no
// [N*C][10] partial buffer.  The R CTAs of a channel form a thread-block cluster; rank 0 adds their partials in rank
// order through distributed shared memory (deterministic, no atomics, no second kernel).
constexpr int DC_THREADS = 224;  // 7 compute warps: 14 bands x 14 strips of a 56x56 plane (and 4 x 7x7 of 28x28) fill them exactly once
constexpr int DC_WARPS = DC_THREADS / 32;
constexpr int DC_BLOCK = DC_THREADS + 32;  // + the copy warp
constexpr int DC_MAX_STAGES = 8;
constexpr int DC_SMEM_2CTA = 100 * 1024;  // ring bytes that still let two CTAs share an SM
constexpr int DC_SMEM_MAX = 200 * 1024;

struct DcGeom {
    int N, C, H, W;
    int R;    // CTAs per channel (cluster size along y)
    int per;  // images per CTA (multiple of P)
    int P;    // images per stage
    int nst;  // stages
};

__device__ __forceinline__ void dc_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void dc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ float dc_ld_dsmem(const float *local, unsigned rank) {
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

constexpr int DC_RB = 4;  // rows per band

// r[0..VEC+1] = columns w0-1 .. w0+VEC of one row of a plane in shared memory.  `row` points at column w0 of the row, or
// into a row of zeros kept behind the ring (rows outside the image, idle lanes), so the loads are unconditional.  The VEC
// inner values are one vector load; the halo columns are the neighbouring lanes' edge values (lane l-1 holds the strip to
// the left whenever sp > 0), except at the two ends of a warp, where that one lane reads shared memory (fl / fr), and at
// the image border (neither shuffle nor load: zero).  Branch-free: the first version of this kernel spent more issue slots
// on BSSY/BSYNC/ISETP and index divisions than on its FMAs (ncu: 29 M warp instructions, 9 M of them FFMA).
template <int VEC>
__device__ __forceinline__ void dc_load_row(const float *row, bool sl, bool sr, bool fl, bool fr, float (&r)[VEC + 2]) {
    float q[VEC];
    if (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(row);
        q[0] = t.x; q[VEC > 1 ? 1 : 0] = t.y; q[VEC > 2 ? 2 : 0] = t.z; q[VEC > 3 ? 3 : 0] = t.w;
    } else if (VEC == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(row);
        q[0] = t.x; q[VEC > 1 ? 1 : 0] = t.y;
    } else {
        q[0] = row[0];
    }
    float lf = 0.0f, rt = 0.0f;
    if (fl) lf = row[-1];
    if (fr) rt = row[VEC];
    const float left = __shfl_up_sync(0xffffffffu, q[VEC - 1], 1);
    const float right = __shfl_down_sync(0xffffffffu, q[0], 1);
    r[0] = sl ? left : lf;   // sl: the left neighbour's value is the halo; else lf (0 at the image border)
#pragma unroll
    for (int e = 0; e < VEC; ++e) r[1 + e] = q[e];
    r[VEC + 1] = sr ? right : rt;
}

template <int VEC>
__global__ void __launch_bounds__(DC_BLOCK, 2)
dw_chan_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ w,
                   float *__restrict__ dx, const float *__restrict__ dx_add, float *__restrict__ dw,
                   float *__restrict__ dbias, float l2, const DcGeom g) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bars[DC_MAX_STAGES], ebars[DC_MAX_STAGES];  // stage filled / stage consumed
    __shared__ float red[DC_WARPS][10];
    __shared__ float xch[10];
    const int c = blockIdx.x, rank = blockIdx.y;
    const int H = g.H, W = g.W, HW = H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = rank * g.per;
    const int n1 = n0 + g.per < g.N ? n0 + g.per : g.N;
    const int groups = n1 > n0 ? (n1 - n0 + g.P - 1) / g.P : 0;
    const int stage_floats = 2 * g.P * HW;
    if (threadIdx.x == 0) {
        for (int s = 0; s < g.nst; ++s) {
            mbar_init(smem_u32(&bars[s]), 1);
            mbar_init(smem_u32(&ebars[s]), DC_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();
    // The copy warp (warp DC_WARPS) refills a stage as soon as the seven compute warps have each released it; the compute
    // warps never wait for one another.  (The first version ended every stage with __syncthreads() and let warp 0 issue the
    // refill: ncu showed 1.07 warps stalled at that barrier per issued instruction -- a third of all warp time -- because the
    // 196 items of a plane leave one warp of seven half empty and everybody waited for the slowest.)
    // fill stage (gi % nst) with the planes of images na .. na+cnt-1 (one bulk copy per plane and tensor)
    auto issue = [&](int gi) {
        const int slot = gi % g.nst;
        const uint32_t bar = smem_u32(&bars[slot]);
        const int na = n0 + gi * g.P;
        const int cnt = n1 - na < g.P ? n1 - na : g.P;
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)(2 * cnt * HW) * 4u);
        __syncwarp();
        float *sg = sm + (size_t)slot * stage_floats, *sx = sg + (size_t)g.P * HW;
        for (int j = lane; j < cnt; j += 32) {
            const long long off = ((long long)(na + j) * g.C + c) * HW;
            dp_bulk_g2s(smem_u32(sg + (size_t)j * HW), dy + off, (uint32_t)HW * 4u, bar);
            dp_bulk_g2s(smem_u32(sx + (size_t)j * HW), x + off, (uint32_t)HW * 4u, bar);
        }
    };
    if (warp == DC_WARPS)
        for (int gi = 0; gi < g.nst && gi < groups; ++gi) issue(gi);
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0f;
    // An item is a band of DC_RB rows x a strip of VEC columns.  The thread walks down the band with a 3-row register
    // window of dY and X, so every staged row is read from shared memory 1.5 times instead of 3, and the two halo
    // columns come from the neighbouring lanes by shuffle (a strip-of-4 access pattern makes scalar shared-memory
    // loads 4-way bank conflicted: they were the bottleneck of the first version of this kernel).
    const int SP = W / VEC, NB = (H + DC_RB - 1) / DC_RB;
    // a row of zeros behind the ring: what out-of-image rows and idle lanes read
    float *zrow = sm + (size_t)g.nst * stage_floats + 4;
    for (int i = threadIdx.x; i < W + 8; i += DC_BLOCK) zrow[i - 4] = 0.0f;
    __syncthreads();
    if (warp == DC_WARPS) {
        for (int gi = g.nst; gi < groups; ++gi) {
            mbar_wait(smem_u32(&ebars[gi % g.nst]), (uint32_t)(gi / g.nst - 1) & 1u);
            issue(gi);
        }
    }
    // item t = (plane p, band, strip sp) for t = tid, and how it moves when t advances by the CTA size
    int sp_0, band_0, p_0;
    {
        const int pb = (int)threadIdx.x / SP;
        sp_0 = (int)threadIdx.x - pb * SP;
        p_0 = pb / NB;
        band_0 = pb - p_0 * NB;
    }
    const int d_sp = DC_THREADS % SP, d_pb = DC_THREADS / SP;
    const int d_band = d_pb % NB, d_p = d_pb / NB;
    for (int gi = 0; gi < (warp < DC_WARPS ? groups : 0); ++gi) {
        const int slot = gi % g.nst;
        mbar_wait(smem_u32(&bars[slot]), (uint32_t)(gi / g.nst) & 1u);
        const int na = n0 + gi * g.P;
        const int cnt = n1 - na < g.P ? n1 - na : g.P;
        const float *sg = sm + (size_t)slot * stage_floats, *sx = sg + (size_t)g.P * HW;
        const int items = cnt * NB * SP;
        int sp = sp_0, band = band_0, p = p_0;
        for (int base = 0; base < items; base += DC_THREADS) {  // CTA-uniform trip count: the shuffles need whole warps
            const bool valid = p < cnt;
            const int w0 = sp * VEC, h0 = band * DC_RB;
            // halo sources: neighbour lane (sl/sr), own shared-memory read at a warp end (fl/fr), or the zero border
            const bool sl = sp != 0 && lane != 0, sr = sp != SP - 1 && lane != 31;
            const bool fl = sp != 0 && lane == 0, fr = sp != SP - 1 && lane == 31;
            const int pv = valid ? p : 0;
            const float *gr = sg + (size_t)pv * HW + (h0 - 1) * W + w0;  // row h0-1 of dY (never read when outside)
            const float *xr = sx + (size_t)pv * HW + (h0 - 1) * W + w0;
            const float *zr = zrow + w0;
            float rg[3][VEC + 2], rx[3][VEC + 2];
            {
                const bool ok0 = valid && h0 > 0;  // row h0-1; row h0 exists whenever the item does
                dc_load_row<VEC>(ok0 ? gr : zr, sl, sr, fl, fr, rg[0]);
                dc_load_row<VEC>(ok0 ? xr : zr, sl, sr, fl, fr, rx[0]);
                dc_load_row<VEC>(valid ? gr + W : zr, sl, sr, fl, fr, rg[1]);
                dc_load_row<VEC>(valid ? xr + W : zr, sl, sr, fl, fr, rx[1]);
            }
            float *orow = dx + ((long long)(na + pv) * g.C + c) * HW + h0 * W + w0;
            const float *arow = dx_add ? dx_add + ((long long)(na + pv) * g.C + c) * HW + h0 * W + w0 : nullptr;
            // the residual gradient folded into dX comes straight from global memory (it is read once): all DC_RB rows of
            // the item are requested here, ahead of the item's window arithmetic.  (Measured: no change -- 56.7 us either way
            // at 56x56x64; the dx_add variant is 68 us against 38 us without it because it moves 16n instead of 12n bytes
            // at the same ~3-4 TB/s this issue-bound kernel reaches, not because it waits on these loads.)
            float4 addv[DC_RB];
            if (VEC == 4) {
#pragma unroll
                for (int rr = 0; rr < DC_RB; ++rr)
                    addv[rr] = (arow != nullptr && valid && h0 + rr < H) ? ld_stream4(arow + rr * W) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int rr = 0; rr < DC_RB; ++rr) {
                const int h = h0 + rr;
                const bool okn = valid && h + 1 < H;
                dc_load_row<VEC>(okn ? gr + (rr + 2) * W : zr, sl, sr, fl, fr, rg[2]);
                dc_load_row<VEC>(okn ? xr + (rr + 2) * W : zr, sl, sr, fl, fr, rx[2]);
                float o[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    // dX[h][w] = sum_{a,b} dY[h+a-1][w+b-1] * k[2-a][2-b]
                    float sacc = 0.0f;
#pragma unroll
                    for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
                        for (int b2 = 0; b2 < 3; ++b2) sacc = fmaf(rg[a2][v + b2], k[(2 - a2) * 3 + (2 - b2)], sacc);
                    o[v] = sacc;
                    const float gv = rg[1][v + 1];  // zero outside the image / for idle lanes
                    acc[9] += gv;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fmaf(gv, rx[i][v + j], acc[i * 3 + j]);
                }
                if (valid && h < H) {
                    if (arow) {
                        if (VEC == 4) {
                            const float4 q = addv[rr];
                            o[0] += q.x; o[VEC > 1 ? 1 : 0] += q.y; o[VEC > 2 ? 2 : 0] += q.z; o[VEC > 3 ? 3 : 0] += q.w;
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) o[v] += __ldg(arow + rr * W + v);
                        }
                    }
                    if (VEC == 2) *reinterpret_cast<float2 *>(orow + rr * W) = make_float2(o[0], o[VEC > 1 ? 1 : 0]);
                    else dp_store<VEC>(orow + rr * W, o);
                }
#pragma unroll
                for (int e = 0; e < VEC + 2; ++e) {
                    rg[0][e] = rg[1][e]; rg[1][e] = rg[2][e];
                    rx[0][e] = rx[1][e]; rx[1][e] = rx[2][e];
                }
            }
            // next item of this thread: t += DC_THREADS
            sp += d_sp;
            int carry = 0;
            if (sp >= SP) { sp -= SP; carry = 1; }
            band += d_band + carry;
            carry = 0;
            if (band >= NB) { band -= NB; carry = 1; }
            p += d_p + carry;
        }
        __syncwarp();  // this warp is done with the stage
        if (lane == 0) mbar_arrive(smem_u32(&ebars[slot]));
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0 && warp < DC_WARPS) {
#pragma unroll
        for (int i = 0; i < 10; ++i) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < DC_WARPS; ++q) t += red[q][threadIdx.x];
        xch[threadIdx.x] = t;
    }
    if (g.R > 1) {
        __syncthreads();
        dc_cluster_arrive();
        dc_cluster_wait();
    } else {
        __syncthreads();
    }
    if (rank == 0 && threadIdx.x < 10) {
        const int i = threadIdx.x;
        float t = xch[i];
        for (int q = 1; q < g.R; ++q) t += dc_ld_dsmem(xch + i, (unsigned)q);  // rank order
        if (i < 9) dw[c * 9 + i] = t + (l2 != 0.0f ? l2 * __ldg(w + c * 9 + i) : 0.0f);
        else if (dbias) dbias[c] = t;
    }
    if (g.R > 1) {
        // nobody leaves while rank 0 may still read its shared memory
        dc_cluster_arrive();
        dc_cluster_wait();
    }
}

// ---------------------------------------------------------------------------------------------------- host
static bool g_dc_ready = false;
int g_dw_chan_enabled = 1;

template <int VEC>
static int dc_set_attrs() {
    DK_CUDA(cudaFuncSetAttribute(dw_chan_bwd_kernel<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM_MAX));
    return DK_OK;
}

int g_dw_planes_enabled = 1;
static bool g_dp_ready = false;

template <int S, int VEC>
static int dp_set_attrs() {
    DK_CUDA(cudaFuncSetAttribute(dw_planes_fwd_kernel<S, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
    DK_CUDA(cudaFuncSetAttribute(dw_planes_bwd_kernel<S, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX + 4096));
    return DK_OK;
}

int init_dw_planes() {
    int rc;
    if ((rc = dp_set_attrs<1, 4>())) return rc;
    if ((rc = dp_set_attrs<1, 1>())) return rc;
    if ((rc = dp_set_attrs<2, 4>())) return rc;
    if ((rc = dp_set_attrs<2, 1>())) return rc;
    g_dp_ready = true;
    if ((rc = dc_set_attrs<4>())) return rc;
    if ((rc = dc_set_attrs<2>())) return rc;
    if ((rc = dc_set_attrs<1>())) return rc;
    g_dc_ready = true;
    return DK_OK;
}

static bool dp_plan(DpGeom &g, int N, int C, int H, int W, int kh, int kw, int s, int p, int per_plane_floats, bool al16) {
    if (!g_dp_ready || !g_dw_planes_enabled || kh != 3 || kw != 3 || p != 1 || (s != 1 && s != 2) || H < 1 || W < 1) return false;
    g.C = C; g.H = H; g.W = W; g.s = s;
    g.OH = (H + 2 - 3) / s + 1;
    g.OW = (W + 2 - 3) / s + 1;
    g.planes = (long long)N * C;
    const long long bytes_per_plane = (long long)per_plane_floats * 4;
    if (bytes_per_plane > DP_SMEM_MAX) return false;
    int G = 1;
    while (G < 64 && (long long)(2 * G) * bytes_per_plane <= DP_TARGET_BYTES) G *= 2;
    if ((long long)G > g.planes) {
        G = 1;
        while ((long long)(2 * G) <= g.planes && G < 64) G *= 2;
    }
    g.G = G;
    // one bulk copy per tensor needs 16-byte multiples at every group boundary
    const int HW = H * W, OHW = g.OH * g.OW;
    g.bulk = al16 && ((long long)G * HW) % 4 == 0 && ((long long)G * OHW) % 4 == 0 && (g.planes * HW) % 4 == 0 &&
             (g.planes * OHW) % 4 == 0;
    return true;
}


// ---- per-channel backward: plan + launch

static bool dc_plan(DcGeom &g, int N, int C, int H, int W, size_t *smem) {
    const long long HW = (long long)H * W;
    const long long pair = 2 * HW * 4;  // bytes of one image's (dY, X) planes
    if (HW % 4 != 0 || 2 * pair > DC_SMEM_MAX || C > 65535) return false;
    int P = 1;
    const int SP = W / (W % 4 == 0 ? 4 : W % 2 == 0 ? 2 : 1);
    while (P < 16 && P < N && (long long)(2 * P) * pair <= 32 * 1024 && (long long)P * ceil_div(H, DC_RB) * SP < DC_THREADS) P *= 2;
    const long long stage = P * pair;
    const long long budget = 2 * stage <= DC_SMEM_2CTA ? DC_SMEM_2CTA : DC_SMEM_MAX;
    int nst = (int)(budget / stage);
    if (nst > DC_MAX_STAGES) nst = DC_MAX_STAGES;
    if (nst < 2) return false;
    const int ctas_target = sm_count() * (budget == DC_SMEM_2CTA ? 2 : 1);
    int R = ctas_target / C;
    if (R < 1) R = 1;
    if (R > 8) R = 8;
    const int groups_total = (int)ceil_div(N, P);
    if (R > groups_total) R = groups_total;
    int per = (int)ceil_div(groups_total, R) * P;
    R = (int)ceil_div(N, per);
    const int groups = per / P;
    if (nst > groups) nst = groups < 2 ? 2 : groups;
    g.N = N; g.C = C; g.H = H; g.W = W; g.R = R; g.per = per; g.P = P; g.nst = nst;
    *smem = (size_t)nst * stage + (size_t)(W + 8) * 4;
    return true;
}

int dw_chan_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st) {
    if (!g_dc_ready || !g_dw_chan_enabled || kh != 3 || kw != 3 || s != 1 || p != 1) return DK_ERR_UNSUPPORTED;
    if (!(aligned16(dy) && aligned16(x) && aligned16(dx) && (dx_add == nullptr || aligned16(dx_add)))) return DK_ERR_UNSUPPORTED;
    DcGeom g;
    size_t smem = 0;
    if (!dc_plan(g, N, C, H, W, &smem)) return DK_ERR_UNSUPPORTED;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C, (unsigned)g.R, 1);
    cfg.blockDim = dim3(DC_BLOCK, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)g.R;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (W % 4 == 0) DK_CUDA(cudaLaunchKernelEx(&cfg, dw_chan_bwd_kernel<4>, dy, x, w, dx, dx_add, dw, dbias, l2, g));
    else if (W % 2 == 0) DK_CUDA(cudaLaunchKernelEx(&cfg, dw_chan_bwd_kernel<2>, dy, x, w, dx, dx_add, dw, dbias, l2, g));
    else DK_CUDA(cudaLaunchKernelEx(&cfg, dw_chan_bwd_kernel<1>, dy, x, w, dx, dx_add, dw, dbias, l2, g));
    count_launch();
    return DK_OK;
}

size_t dw_planes_ws_bytes(int N, int C) { return (size_t)N * C * 10 * sizeof(float); }

int dw_planes_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                  int s, int p, cudaStream_t st) {
    DpGeom g;
    if (!dp_plan(g, N, C, H, W, kh, kw, s, p, H * W, aligned16(x) && aligned16(y))) return DK_ERR_UNSUPPORTED;
    const size_t smem = (size_t)g.G * H * W * 4;
    if (smem > (size_t)DP_SMEM_MAX) return DK_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(g.planes, g.G);
    const bool v4 = (W % 4 == 0) && (g.OW % 4 == 0) && aligned16(y) && (s == 1 || W % 8 == 0);
    if (s == 1) {
        if (v4) dw_planes_fwd_kernel<1, 4><<<grid, DP_THREADS, smem, st>>>(x, w, bias, y, g);
        else dw_planes_fwd_kernel<1, 1><<<grid, DP_THREADS, smem, st>>>(x, w, bias, y, g);
    } else {
        if (v4) dw_planes_fwd_kernel<2, 4><<<grid, DP_THREADS, smem, st>>>(x, w, bias, y, g);
        else dw_planes_fwd_kernel<2, 1><<<grid, DP_THREADS, smem, st>>>(x, w, bias, y, g);
    }
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dw_planes_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                  float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, void *ws, size_t ws_bytes,
                  cudaStream_t st) {
    DpGeom g;
    const int OH = (H - 1) / s + 1, OW = (W - 1) / s + 1;
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx) && (dx_add == nullptr || aligned16(dx_add));
    if (!dp_plan(g, N, C, H, W, kh, kw, s, p, H * W + OH * OW, al)) return DK_ERR_UNSUPPORTED;
    if (ws == nullptr || ws_bytes < dw_planes_ws_bytes(N, C)) return DK_ERR_UNSUPPORTED;
    const int wpp = g.G >= DP_WARPS ? 1 : DP_WARPS / g.G;
    const size_t smem = ((((size_t)g.G * g.OH * g.OW + 3) & ~(size_t)3) + (((size_t)g.G * H * W + 3) & ~(size_t)3) +
                         (size_t)g.G * wpp * 10) * 4;
    if (smem > (size_t)DP_SMEM_MAX + 4096) return DK_ERR_UNSUPPORTED;
    float *partial = reinterpret_cast<float *>(ws);
    const unsigned grid = (unsigned)ceil_div(g.planes, g.G);
    const bool v4 = (W % 4 == 0) && al;
    if (s == 1) {
        if (v4) dw_planes_bwd_kernel<1, 4><<<grid, DP_THREADS, smem, st>>>(dy, x, w, dx, partial, dx_add, g);
        else dw_planes_bwd_kernel<1, 1><<<grid, DP_THREADS, smem, st>>>(dy, x, w, dx, partial, dx_add, g);
    } else {
        if (v4) dw_planes_bwd_kernel<2, 4><<<grid, DP_THREADS, smem, st>>>(dy, x, w, dx, partial, dx_add, g);
        else dw_planes_bwd_kernel<2, 1><<<grid, DP_THREADS, smem, st>>>(dy, x, w, dx, partial, dx_add, g);
    }
    DK_LAUNCH_CHECK();
    dw_planes_reduce_kernel<<<C, 128, 0, st>>>(partial, w, dw, dbias, l2, N, C);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
