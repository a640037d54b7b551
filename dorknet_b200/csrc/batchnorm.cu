// batchnorm.cu -- BatchNorm statistics, apply and backward for NCHW float32 (HBM-bound).
//
// Statistics: each channel's N*HW elements are cut into S splits (one CTA each, S chosen so that
// C*S is a whole number of waves over the SMs).  A thread accumulates shifted sums (shift = first
// value it sees, which removes the catastrophic cancellation of sum(x^2)-sum(x)^2), converts to
// (count, mean, M2) and the partials are merged pairwise with Chan's formula: warp shuffles, then
// shared memory, then -- by the last CTA of the channel to arrive, in split order, so the result is
// deterministic -- across CTAs.  That last CTA also finalises the channel (std, scale/shift,
// running statistics), so statistics + finalise are ONE launch with no memset: the arrival
// counters live in a caller-provided, zero-initialised workspace and are reset on exit.
#include "bn.cuh"

namespace dk {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_SPLITS = 64;
constexpr int BN_WS_FLOATS_PER_SPLIT = 4;  // stats use 3 (n, mean, M2); backward uses 2

__device__ __forceinline__ Moments warp_merge(Moments m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Moments other;
        other.n = __shfl_xor_sync(0xffffffffu, m.n, o);
        other.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
        other.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
        // keep the merge order identical on both lanes of a pair (lower lane first) so every lane
        // ends with bit-identical values
        const bool low = ((threadIdx.x & o) == 0);
        m = low ? merge(m, other) : merge(other, m);
    }
    return m;
}

// x viewed as [N, C, HW]; channel c's virtual index v in [0, N*HW) maps to x[(v/HW)*C*HW + c*HW + v%HW].
template <bool VEC>
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float *__restrict__ x, int N, int C, int HW, int S, int64_t per_split,
                float *__restrict__ ws_part, unsigned int *__restrict__ ws_count, BnFinalize fin) {
    const int c = blockIdx.x, s = blockIdx.y;
    const int64_t total = (int64_t)N * HW;
    const int64_t v0 = (int64_t)s * per_split;
    const int64_t v1 = v0 + per_split < total ? v0 + per_split : total;
    const int64_t cstride = (int64_t)C * HW;

    float shift = 0.0f, sum = 0.0f, sq = 0.0f, cnt = 0.0f;
    bool have = false;
    if (VEC) {
        // four independent 128-bit loads in flight per thread before any arithmetic
        constexpr int U = 4;
        int64_t v = v0 + 4 * (int64_t)threadIdx.x;
        for (; v + (U - 1) * 4 * BN_THREADS < v1; v += U * 4 * BN_THREADS) {
            float4 t[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t vv = v + (int64_t)u * 4 * BN_THREADS;
                const uint32_t n = (uint32_t)vv / (uint32_t)HW, off = (uint32_t)vv - n * (uint32_t)HW;  // N*HW < 2^31 (host check)  // HW % 4 == 0: a float4 never straddles planes
                t[u] = ld_stream4(x + n * cstride + (int64_t)c * HW + off);
            }
            if (!have) { shift = t[0].x; have = true; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float d0 = t[u].x - shift, d1 = t[u].y - shift, d2 = t[u].z - shift, d3 = t[u].w - shift;
                sum += (d0 + d1) + (d2 + d3);
                sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
            cnt += 4.0f * U;
        }
        for (; v < v1; v += 4 * BN_THREADS) {
            const uint32_t n = (uint32_t)v / (uint32_t)HW, off = (uint32_t)v - n * (uint32_t)HW;
            const float4 t = ld_stream4(x + n * cstride + (int64_t)c * HW + off);
            if (!have) { shift = t.x; have = true; }
            const float d0 = t.x - shift, d1 = t.y - shift, d2 = t.z - shift, d3 = t.w - shift;
            sum += (d0 + d1) + (d2 + d3);
            sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            cnt += 4.0f;
        }
    } else {
        for (int64_t v = v0 + threadIdx.x; v < v1; v += BN_THREADS) {
            const uint32_t n = (uint32_t)v / (uint32_t)HW, off = (uint32_t)v - n * (uint32_t)HW;
            const float t = x[n * cstride + (int64_t)c * HW + off];
            if (!have) { shift = t; have = true; }
            const float d = t - shift;
            sum += d;
            sq += d * d;
            cnt += 1.0f;
        }
    }
    Moments m;
    m.n = cnt;
    m.mean = cnt > 0.0f ? shift + sum / cnt : 0.0f;
    m.m2 = cnt > 0.0f ? fmaxf(sq - sum * sum / cnt, 0.0f) : 0.0f;
    m = warp_merge(m);

    __shared__ Moments sm[BN_THREADS / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        Moments tot = sm[0];
        for (int w = 1; w < BN_THREADS / 32; ++w) tot = merge(tot, sm[w]);
        float *p = ws_part + ((size_t)c * BN_MAX_SPLITS + s) * BN_WS_FLOATS_PER_SPLIT;
        p[0] = tot.n;
        p[1] = tot.mean;
        p[2] = tot.m2;
        __threadfence();
        const unsigned int prev = atomicAdd(&ws_count[c], 1u);
        is_last = (prev == (unsigned int)(S - 1));
        if (is_last) {
            __threadfence();
            const volatile float *vp = ws_part + (size_t)c * BN_MAX_SPLITS * BN_WS_FLOATS_PER_SPLIT;
            Moments all;
            all.n = vp[0]; all.mean = vp[1]; all.m2 = vp[2];
            for (int k = 1; k < S; ++k) {
                Moments o;
                o.n = vp[k * BN_WS_FLOATS_PER_SPLIT + 0];
                o.mean = vp[k * BN_WS_FLOATS_PER_SPLIT + 1];
                o.m2 = vp[k * BN_WS_FLOATS_PER_SPLIT + 2];
                all = merge(all, o);
            }
            const float mean = all.mean;
            const float var = all.n > 0.0f ? all.m2 / all.n : 0.0f;  // biased (batch_norm_stats_cy.pyx:44)
            if (fin.mode == 0) {
                fin.mean_out[c] = mean;
                fin.var_out[c] = var;
            } else {
                const float std = sqrtf(var + fin.eps);  // batch_norm.py:69
                const float invstd = 1.0f / std;
                const float scale = fin.gamma[c] * invstd;
                fin.save_mean[c] = mean;
                fin.save_invstd[c] = invstd;
                fin.save_scale[c] = scale;
                fin.save_shift[c] = fin.beta[c] - mean * scale;
                if (fin.running_mean) {  // batch_norm.py:76-89 (tracks std, not var)
                    if (fin.first_batch) {
                        fin.running_mean[c] = mean;
                        fin.running_std[c] = std;
                    } else {
                        const float mo = fin.momentum;
                        fin.running_mean[c] = mo * fin.running_mean[c] + (1.0f - mo) * mean;
                        fin.running_std[c] = mo * fin.running_std[c] + (1.0f - mo) * std;
                    }
                }
            }
            ws_count[c] = 0;  // leave the workspace clean for the next launch
        }
    }
}

// y = x*scale[c] + shift[c] (+ReLU)
template <bool VEC, bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ scale,
                const float *__restrict__ shift, int64_t total, int C, int HW) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t nvec = total >> 2;
        const int hw4 = HW >> 2;
        for (int64_t i = tid; i < nvec; i += nthreads) {
            const int c = (int)((i / hw4) % C);
            const float sc = __ldg(scale + c), sh = __ldg(shift + c);
            float4 v = ld_stream4(x + 4 * i);
            v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
            if (RELU) {
                v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
            }
            st_stream4(y + 4 * i, v);
        }
    } else {
        for (int64_t i = tid; i < total; i += nthreads) {
            const int c = (int)((i / HW) % C);
            float v = fmaf(x[i], __ldg(scale + c), __ldg(shift + c));
            if (RELU) v = v > 0.f ? v : 0.f;
            y[i] = v;
        }
    }
}

// y[n,c,oh,ow] = relu?(x[n,c,oh*s,ow*s]*scale[c] + shift[c]): the normalisation pass restricted to the pixels a following
// stride-s pointwise convolution reads (X[:, :, ::s, ::s], pointwise_convolution.py:48), written compactly.  In
// ResNet-18-depsep conv0_bn -> ReLU -> pw0 (stride 2) uses one pixel in four: the full-size apply wrote 205 MB per
// step that nobody read.  VEC: stride 2, W % 8 == 0: a thread turns 8 consecutive inputs into 4 consecutive outputs.
template <bool VEC, bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_strided_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ scale,
                        const float *__restrict__ shift, int64_t total_out, int C, int H, int W, int OH, int OW, int s) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int ow4 = OW >> 2;
        const int64_t nvec = total_out >> 2;
        constexpr int U = 2;  // 2 x 32 bytes in flight per thread before anything is stored
        for (int64_t i0 = tid; i0 < nvec; i0 += U * nthreads) {
            float4 a[U], b[U];
            float sc[U], sh[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * nthreads;
                if (i < nvec) {
                    const int q = (int)(i % ow4);
                    const int64_t r = i / ow4;  // (n*C + c)*OH + oh
                    const int oh = (int)(r % OH);
                    const int64_t plane = r / OH;
                    const int c = (int)(plane % C);
                    sc[u] = __ldg(scale + c);
                    sh[u] = __ldg(shift + c);
                    const float *src = x + (plane * H + (int64_t)oh * 2) * W + 8 * q;
                    a[u] = ld_stream4(src);
                    b[u] = ld_stream4(src + 4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + u * nthreads;
                if (i < nvec) {
                    float4 v = make_float4(fmaf(a[u].x, sc[u], sh[u]), fmaf(a[u].z, sc[u], sh[u]), fmaf(b[u].x, sc[u], sh[u]),
                                           fmaf(b[u].z, sc[u], sh[u]));
                    if (RELU) {
                        v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                        v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
                    }
                    st_stream4(y + 4 * i, v);
                }
            }
        }
    } else {
        for (int64_t i = tid; i < total_out; i += nthreads) {
            const int ow = (int)(i % OW);
            const int64_t r = i / OW;
            const int oh = (int)(r % OH);
            const int64_t plane = r / OH;
            const int c = (int)(plane % C);
            float v = fmaf(x[(plane * H + (int64_t)oh * s) * W + (int64_t)ow * s], __ldg(scale + c), __ldg(shift + c));
            if (RELU) v = v > 0.f ? v : 0.f;
            y[i] = v;
        }
    }
}

// test mode: y = gamma*((x - rm)/rs) + beta, evaluated as the reference does (batch_norm.py:112-115)
template <bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_infer_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ gamma,
                const float *__restrict__ beta, const float *__restrict__ rm, const float *__restrict__ rs,
                int64_t total, int C, int HW) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int c = (int)((i / HW) % C);
        float v = gamma[c] * ((x[i] - rm[c]) / rs[c]) + beta[c];
        if (RELU) v = v > 0.f ? v : 0.f;
        y[i] = v;
    }
}

// backward pass 1: per channel sum(g), sum(g*x_hat); last CTA writes dgamma/dbeta and k1,k2.
template <bool VEC, bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(const float *__restrict__ dy, const float *__restrict__ x, int N, int C, int HW, int S,
                     int64_t per_split, const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                     const float *__restrict__ save_scale, const float *__restrict__ save_shift,
                     float *__restrict__ ws_part, unsigned int *__restrict__ ws_count, float *__restrict__ coef,
                     float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x, s = blockIdx.y;
    const int64_t total = (int64_t)N * HW;
    const int64_t v0 = (int64_t)s * per_split;
    const int64_t v1 = v0 + per_split < total ? v0 + per_split : total;
    const int64_t cstride = (int64_t)C * HW;
    const float mean = save_mean[c], invstd = save_invstd[c];
    const float sc = RELU ? save_scale[c] : 0.f, sh = RELU ? save_shift[c] : 0.f;
    float sg = 0.0f, sgx = 0.0f;
    if (VEC) {
        constexpr int U = 2;  // 2 x (dY, X) 128-bit loads in flight per thread
        int64_t v = v0 + 4 * (int64_t)threadIdx.x;
        for (; v + (U - 1) * 4 * BN_THREADS < v1; v += U * 4 * BN_THREADS) {
            float4 gq[U], tq[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t vv = v + (int64_t)u * 4 * BN_THREADS;
                const uint32_t n = (uint32_t)vv / (uint32_t)HW, off = (uint32_t)vv - n * (uint32_t)HW;  // N*HW < 2^31 (host check)
                const int64_t idx = n * cstride + (int64_t)c * HW + off;
                gq[u] = ld_stream4(dy + idx);
                tq[u] = ld_stream4(x + idx);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float4 g = gq[u];
                const float4 t = tq[u];
                if (RELU) {
                    g.x = fmaf(t.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(t.y, sc, sh) > 0.f ? g.y : 0.f;
                    g.z = fmaf(t.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(t.w, sc, sh) > 0.f ? g.w : 0.f;
                }
                sg += (g.x + g.y) + (g.z + g.w);
                sgx += (g.x * ((t.x - mean) * invstd) + g.y * ((t.y - mean) * invstd)) +
                       (g.z * ((t.z - mean) * invstd) + g.w * ((t.w - mean) * invstd));
            }
        }
        for (; v < v1; v += 4 * BN_THREADS) {
            const uint32_t n = (uint32_t)v / (uint32_t)HW, off = (uint32_t)v - n * (uint32_t)HW;
            const int64_t idx = n * cstride + (int64_t)c * HW + off;
            float4 g = ld_stream4(dy + idx);
            const float4 t = ld_stream4(x + idx);
            if (RELU) {
                g.x = fmaf(t.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(t.y, sc, sh) > 0.f ? g.y : 0.f;
                g.z = fmaf(t.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(t.w, sc, sh) > 0.f ? g.w : 0.f;
            }
            sg += (g.x + g.y) + (g.z + g.w);
            sgx += (g.x * ((t.x - mean) * invstd) + g.y * ((t.y - mean) * invstd)) +
                   (g.z * ((t.z - mean) * invstd) + g.w * ((t.w - mean) * invstd));
        }
    } else {
        for (int64_t v = v0 + threadIdx.x; v < v1; v += BN_THREADS) {
            const uint32_t n = (uint32_t)v / (uint32_t)HW, off = (uint32_t)v - n * (uint32_t)HW;
            const int64_t idx = n * cstride + (int64_t)c * HW + off;
            float g = dy[idx];
            const float t = x[idx];
            if (RELU) g = fmaf(t, sc, sh) > 0.f ? g : 0.f;
            sg += g;
            sgx += g * ((t - mean) * invstd);
        }
    }
    __shared__ float red[33];
    __shared__ bool is_last;
    sg = block_sum(sg, red);
    sgx = block_sum(sgx, red);
    if (threadIdx.x == 0) {
        float *p = ws_part + ((size_t)c * BN_MAX_SPLITS + s) * BN_WS_FLOATS_PER_SPLIT;
        p[0] = sg;
        p[1] = sgx;
        __threadfence();
        const unsigned int prev = atomicAdd(&ws_count[c], 1u);
        is_last = (prev == (unsigned int)(S - 1));
        if (is_last) {
            __threadfence();
            const volatile float *vp = ws_part + (size_t)c * BN_MAX_SPLITS * BN_WS_FLOATS_PER_SPLIT;
            float a = 0.0f, b = 0.0f;
            for (int k = 0; k < S; ++k) {
                a += vp[k * BN_WS_FLOATS_PER_SPLIT + 0];
                b += vp[k * BN_WS_FLOATS_PER_SPLIT + 1];
            }
            dbeta[c] = a;
            dgamma[c] = b;
            const float inv_n = 1.0f / (float)total;
            coef[2 * c + 0] = a * inv_n;  // mean(dy)
            coef[2 * c + 1] = b * inv_n;  // mean(dy * x_hat)
            ws_count[c] = 0;
        }
    }
}

// backward pass 2: dx = scale * (g - k1 - x_hat*k2)
template <bool VEC, bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_dx_kernel(const float *__restrict__ dy, const float *__restrict__ x, float *__restrict__ dx,
                 const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                 const float *__restrict__ save_scale, const float *__restrict__ save_shift,
                 const float *__restrict__ coef, int64_t total, int C, int HW) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t nvec = total >> 2;
        const int hw4 = HW >> 2;
        for (int64_t i = tid; i < nvec; i += nthreads) {
            const int c = (int)((i / hw4) % C);
            const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c), sc = __ldg(save_scale + c);
            const float k1 = __ldg(coef + 2 * c), k2 = __ldg(coef + 2 * c + 1);
            float4 g = ld_stream4(dy + 4 * i);
            const float4 t = ld_stream4(x + 4 * i);
            if (RELU) {
                const float sh = __ldg(save_shift + c);
                g.x = fmaf(t.x, sc, sh) > 0.f ? g.x : 0.f; g.y = fmaf(t.y, sc, sh) > 0.f ? g.y : 0.f;
                g.z = fmaf(t.z, sc, sh) > 0.f ? g.z : 0.f; g.w = fmaf(t.w, sc, sh) > 0.f ? g.w : 0.f;
            }
            float4 r;
            r.x = sc * (g.x - k1 - ((t.x - mean) * invstd) * k2);
            r.y = sc * (g.y - k1 - ((t.y - mean) * invstd) * k2);
            r.z = sc * (g.z - k1 - ((t.z - mean) * invstd) * k2);
            r.w = sc * (g.w - k1 - ((t.w - mean) * invstd) * k2);
            st_stream4(dx + 4 * i, r);
        }
    } else {
        for (int64_t i = tid; i < total; i += nthreads) {
            const int c = (int)((i / HW) % C);
            const float mean = save_mean[c], invstd = save_invstd[c], sc = save_scale[c];
            float g = dy[i];
            const float t = x[i];
            if (RELU) g = fmaf(t, sc, save_shift[c]) > 0.f ? g : 0.f;
            dx[i] = sc * (g - coef[2 * c] - ((t - mean) * invstd) * coef[2 * c + 1]);
        }
    }
}

// ---- backward with a stride-2 COMPACT upstream gradient ------------------------------------------------------------------
// The gradient a stride-2 pointwise convolution sends back is zero except at (even row, even column) pixels
// (pointwise_convolution.py:68-72 zero-stuffs it to full size).  Taking it in its compact [N, C, H/2, W/2] form, pass 1
// only visits those pixels (a quarter of dY, half of X's sectors) and pass 2 reads a quarter of the dY bytes; nobody
// writes or reads the 3/4 zeros (154 MB each way for conv0_bn at batch 64).  Requires W % 8 == 0, H % 2 == 0.
template <bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_sub_kernel(const float *__restrict__ dys, const float *__restrict__ x, int N, int C, int H, int W, int S,
                         int64_t per_split, const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                         const float *__restrict__ save_scale, const float *__restrict__ save_shift,
                         float *__restrict__ ws_part, unsigned int *__restrict__ ws_count, float *__restrict__ coef,
                         float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x, s = blockIdx.y;
    const int OH = H >> 1, OW = W >> 1, ow4 = OW >> 2;
    const int64_t total4 = (int64_t)N * OH * ow4;  // float4 groups of the compact gradient of this channel
    const int64_t v0 = (int64_t)s * per_split;
    const int64_t v1 = v0 + per_split < total4 ? v0 + per_split : total4;
    const float mean = save_mean[c], invstd = save_invstd[c];
    const float sc = RELU ? save_scale[c] : 0.f, sh = RELU ? save_shift[c] : 0.f;
    float sg = 0.0f, sgx = 0.0f;
    constexpr int U = 2;
    for (int64_t v = v0 + threadIdx.x; v < v1; v += U * BN_THREADS) {
        float4 gq[U], ta[U], tb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t vv = v + (int64_t)u * BN_THREADS;
            if (vv < v1) {
                const int q = (int)(vv % ow4);
                const int64_t r = vv / ow4;
                const int oh = (int)(r % OH);
                const int64_t n = r / OH;
                gq[u] = ld_stream4(dys + ((n * C + c) * OH + oh) * OW + 4 * q);
                const float *xp = x + ((n * C + c) * H + 2 * oh) * (int64_t)W + 8 * q;
                ta[u] = ld_stream4(xp);
                tb[u] = ld_stream4(xp + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v + (int64_t)u * BN_THREADS < v1) {
                float g[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
                const float t[4] = {ta[u].x, ta[u].z, tb[u].x, tb[u].z};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (RELU) g[e] = fmaf(t[e], sc, sh) > 0.f ? g[e] : 0.f;
                    sg += g[e];
                    sgx = fmaf(g[e], (t[e] - mean) * invstd, sgx);
                }
            }
        }
    }
    __shared__ float red[33];
    sg = block_sum(sg, red);
    sgx = block_sum(sgx, red);
    if (threadIdx.x == 0) {
        float *p = ws_part + ((size_t)c * BN_MAX_SPLITS + s) * BN_WS_FLOATS_PER_SPLIT;
        p[0] = sg;
        p[1] = sgx;
        __threadfence();
        const unsigned int prev = atomicAdd(&ws_count[c], 1u);
        if (prev == (unsigned int)(S - 1)) {
            __threadfence();
            const volatile float *vp = ws_part + (size_t)c * BN_MAX_SPLITS * BN_WS_FLOATS_PER_SPLIT;
            float a = 0.0f, b = 0.0f;
            for (int k = 0; k < S; ++k) {
                a += vp[k * BN_WS_FLOATS_PER_SPLIT + 0];
                b += vp[k * BN_WS_FLOATS_PER_SPLIT + 1];
            }
            dbeta[c] = a;
            dgamma[c] = b;
            const float inv_n = 1.0f / (float)((int64_t)N * H * W);  // means over ALL pixels (the others contribute 0)
            coef[2 * c + 0] = a * inv_n;
            coef[2 * c + 1] = b * inv_n;
            ws_count[c] = 0;
        }
    }
}

template <bool RELU>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_dx_sub_kernel(const float *__restrict__ dys, const float *__restrict__ x, float *__restrict__ dx,
                     const float *__restrict__ save_mean, const float *__restrict__ save_invstd,
                     const float *__restrict__ save_scale, const float *__restrict__ save_shift,
                     const float *__restrict__ coef, int64_t total, int C, int H, int W) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w4 = W >> 2, OW = W >> 1, OH = H >> 1;
    const int64_t nvec = total >> 2;
    for (int64_t i = tid; i < nvec; i += nthreads) {
        const int q = (int)(i % w4);
        const int64_t r = i / w4;  // (n*C + c)*H + h
        const int h = (int)(r % H);
        const int64_t plane = r / H;
        const int c = (int)(plane % C);
        const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c), sc = __ldg(save_scale + c);
        const float k1 = __ldg(coef + 2 * c), k2 = __ldg(coef + 2 * c + 1);
        const float4 t = ld_stream4(x + 4 * i);
        float g0 = 0.0f, g2 = 0.0f;
        if ((h & 1) == 0) {
            const float2 gg = __ldg(reinterpret_cast<const float2 *>(dys + (plane * OH + (h >> 1)) * OW + 2 * q));
            g0 = gg.x;
            g2 = gg.y;
            if (RELU) {
                const float sh = __ldg(save_shift + c);
                g0 = fmaf(t.x, sc, sh) > 0.f ? g0 : 0.f;
                g2 = fmaf(t.z, sc, sh) > 0.f ? g2 : 0.f;
            }
        }
        float4 o;
        o.x = sc * (g0 - k1 - ((t.x - mean) * invstd) * k2);
        o.y = sc * (0.f - k1 - ((t.y - mean) * invstd) * k2);
        o.z = sc * (g2 - k1 - ((t.z - mean) * invstd) * k2);
        o.w = sc * (0.f - k1 - ((t.w - mean) * invstd) * k2);
        st_stream4(dx + 4 * i, o);
    }
}

// ---- host side ------------------------------------------------------------------------------------
struct BnWs {
    float *part;          // [C][BN_MAX_SPLITS][4]
    float *coef;          // [C][2]
    unsigned int *count;  // [C]
};

static size_t bn_ws_bytes(int C) {
    return (size_t)C * (BN_MAX_SPLITS * BN_WS_FLOATS_PER_SPLIT + 2 + 1) * sizeof(float);
}

static BnWs bn_ws_carve(void *ws, int C) {
    BnWs w;
    w.part = reinterpret_cast<float *>(ws);
    w.coef = w.part + (size_t)C * BN_MAX_SPLITS * BN_WS_FLOATS_PER_SPLIT;
    w.count = reinterpret_cast<unsigned int *>(w.coef + (size_t)C * 2);
    return w;
}

// choose the number of splits per channel: ~4 CTAs per SM in total, each split a multiple of 4 elements
int g_bn_split_ctas_per_sm = 0;  // split-kernel CTAs per SM the plan aims for (0 = default; dk_tc_debug_set key 24)
static void bn_plan(int N, int C, int HW, int *S, int64_t *per_split, int ctas_per_sm = 8) {
    const int64_t total = (int64_t)N * HW;
    if (g_bn_split_ctas_per_sm > 0) ctas_per_sm = g_bn_split_ctas_per_sm;
    int64_t want = ceil_div((int64_t)sm_count() * ctas_per_sm, C);  // resident CTAs of 256 threads per SM
    const int64_t max_by_work = ceil_div(total, 16 * BN_THREADS);  // at least four float4 per thread
    if (want > max_by_work) want = max_by_work;
    if (want > BN_MAX_SPLITS) want = BN_MAX_SPLITS;
    if (want < 1) want = 1;
    int64_t per = ceil_div(total, want);
    per = ceil_div(per, 4) * 4;
    *S = (int)ceil_div(total, per);
    *per_split = per;
}

static int bn_check(const char *who, int N, int C, int HW, const void *ws, size_t ws_bytes) {
    DK_REQUIRE(N > 0 && C > 0 && HW > 0, "%s: bad shape N=%d C=%d HW=%d", who, N, C, HW);
    DK_REQUIRE((int64_t)N * C * HW < ((int64_t)1 << 40) && (int64_t)N * HW < ((int64_t)1 << 31), "%s: tensor too large", who);
    if (ws_bytes < bn_ws_bytes(C) || ws == nullptr) {
        set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, bn_ws_bytes(C));
        return DK_ERR_WORKSPACE;
    }
    return DK_OK;
}

static int launch_stats(const float *x, int N, int C, int HW, void *ws, const BnFinalize &fin, cudaStream_t st) {
    int S;
    int64_t per;
    // statistics pass: 4 CTAs per SM (tests/bn_stats_sweep.py, batch 64, us at 2 / 4 / 8 / 16 per SM: conv0_bn 77 / 49.8 / 57.7 /
    // 52.3; 56x56x64 27.1 / 21.3 / 24.8 / 31.6; 28x28x128 15.9 / 14.8 / 17.1 / 20.5)
    bn_plan(N, C, HW, &S, &per, 4);
    BnWs w = bn_ws_carve(ws, C);
    dim3 grid(C, S);
    const bool vec = (HW % 4 == 0) && aligned16(x);
    if (vec) bn_stats_kernel<true><<<grid, BN_THREADS, 0, st>>>(x, N, C, HW, S, per, w.part, w.count, fin);
    else bn_stats_kernel<false><<<grid, BN_THREADS, 0, st>>>(x, N, C, HW, S, per, w.part, w.count, fin);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

static int launch_apply(const float *x, float *y, const float *scale, const float *shift, int relu,
                        int N, int C, int HW, cudaStream_t st) {
    const int64_t total = (int64_t)N * C * HW;
    const bool vec = (HW % 4 == 0) && aligned16(x) && aligned16(y);
    const int grid = stream_grid(vec ? total / 4 : total, BN_THREADS * 2);
    if (vec) {
        if (relu) bn_apply_kernel<true, true><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, HW);
        else bn_apply_kernel<true, false><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, HW);
    } else {
        if (relu) bn_apply_kernel<false, true><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, HW);
        else bn_apply_kernel<false, false><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, HW);
    }
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk

using namespace dk;

extern "C" {

size_t dk_bn_ws_bytes(int C) { return C > 0 ? bn_ws_bytes(C) : 0; }

int dk_bn_stats(const float *x, float *mean, float *var, int N, int C, int HW, void *ws, size_t ws_bytes,
                dk_stream_t stream) {
    int rc = bn_check("dk_bn_stats", N, C, HW, ws, ws_bytes);
    if (rc) return rc;
    DK_REQUIRE(x && mean && var, "dk_bn_stats: NULL pointer");
    BnFinalize fin = {};
    fin.mode = 0;
    fin.mean_out = mean;
    fin.var_out = var;
    return launch_stats(x, N, C, HW, ws, fin, as_stream(stream));
}

static int bn_fwd_train_impl(const float *x, const float *add, float *y, const float *gamma, const float *beta,
                             float *running_mean, float *running_std, int first_batch, float momentum, float eps,
                             float *save_mean, float *save_invstd, float *save_scale, float *save_shift, int fuse_relu, int N,
                             int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = bn_check("dk_bn_fwd_train", N, C, HW, ws, ws_bytes);
    if (rc) return rc;
    DK_REQUIRE(x && gamma && beta && save_mean && save_invstd && save_scale && save_shift,
               "dk_bn_fwd_train: NULL pointer");
    DK_REQUIRE((running_mean == nullptr) == (running_std == nullptr), "dk_bn_fwd_train: running stats must come in pairs");
    BnFinalize fin = {};
    fin.mode = 1;
    fin.gamma = gamma;
    fin.beta = beta;
    fin.running_mean = running_mean;
    fin.running_std = running_std;
    fin.first_batch = first_batch;
    fin.momentum = momentum;
    fin.eps = eps;
    fin.save_mean = save_mean;
    fin.save_invstd = save_invstd;
    fin.save_scale = save_scale;
    fin.save_shift = save_shift;
    if (y != nullptr) {  // statistics + normalisation in one cluster kernel when the channel slices fit shared memory
        rc = bn_fused_fwd(x, y, fin, fuse_relu, N, C, HW, as_stream(stream), add);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    rc = launch_stats(x, N, C, HW, ws, fin, as_stream(stream));
    if (rc || y == nullptr) return rc;
    if (add == nullptr) return launch_apply(x, y, save_scale, save_shift, fuse_relu, N, C, HW, as_stream(stream));
    // residual join on the split path: plain apply, then y = relu(y + add) in place (never the case in the networks here)
    rc = launch_apply(x, y, save_scale, save_shift, 0, N, C, HW, as_stream(stream));
    if (rc) return rc;
    if (fuse_relu) return dk_add_relu_fwd(y, add, y, (int64_t)N * C * HW, stream);
    return dk_add(y, add, y, (int64_t)N * C * HW, stream);
}

int dk_bn_fwd_train(const float *x, float *y, const float *gamma, const float *beta, float *running_mean,
                    float *running_std, int first_batch, float momentum, float eps, float *save_mean,
                    float *save_invstd, float *save_scale, float *save_shift, int fuse_relu, int N, int C, int HW,
                    void *ws, size_t ws_bytes, dk_stream_t stream) {
    return bn_fwd_train_impl(x, nullptr, y, gamma, beta, running_mean, running_std, first_batch, momentum, eps, save_mean,
                             save_invstd, save_scale, save_shift, fuse_relu, N, C, HW, ws, ws_bytes, stream);
}

int dk_bn_fwd_train_add(const float *x, const float *add, float *y, const float *gamma, const float *beta,
                        float *running_mean, float *running_std, int first_batch, float momentum, float eps,
                        float *save_mean, float *save_invstd, float *save_scale, float *save_shift, int fuse_relu, int N,
                        int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream) {
    DK_REQUIRE(add != nullptr && y != nullptr, "dk_bn_fwd_train_add: NULL pointer");
    return bn_fwd_train_impl(x, add, y, gamma, beta, running_mean, running_std, first_batch, momentum, eps, save_mean,
                             save_invstd, save_scale, save_shift, fuse_relu, N, C, HW, ws, ws_bytes, stream);
}

int dk_bn_apply(const float *x, float *y, const float *scale, const float *shift, int fuse_relu, int N, int C,
                int HW, dk_stream_t stream) {
    DK_REQUIRE(N > 0 && C > 0 && HW > 0 && x && y && scale && shift, "dk_bn_apply: bad arguments");
    return launch_apply(x, y, scale, shift, fuse_relu, N, C, HW, as_stream(stream));
}

int dk_bn_apply_strided(const float *x, float *y, const float *scale, const float *shift, int fuse_relu, int N, int C,
                        int H, int W, int stride, dk_stream_t stream) {
    DK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && stride >= 1 && x && y && scale && shift, "dk_bn_apply_strided: bad arguments");
    const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
    const int64_t total = (int64_t)N * C * OH * OW;
    const bool vec = stride == 2 && (W % 8 == 0) && aligned16(x) && aligned16(y);
    const int grid = stream_grid(vec ? total / 4 : total, BN_THREADS * 2);
    cudaStream_t st = as_stream(stream);
    if (vec) {
        if (fuse_relu) bn_apply_strided_kernel<true, true><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, H, W, OH, OW, stride);
        else bn_apply_strided_kernel<true, false><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, H, W, OH, OW, stride);
    } else {
        if (fuse_relu) bn_apply_strided_kernel<false, true><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, H, W, OH, OW, stride);
        else bn_apply_strided_kernel<false, false><<<grid, BN_THREADS, 0, st>>>(x, y, scale, shift, total, C, H, W, OH, OW, stride);
    }
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_bn_fwd_infer(const float *x, float *y, const float *gamma, const float *beta, const float *running_mean,
                    const float *running_std, int fuse_relu, int N, int C, int HW, dk_stream_t stream) {
    DK_REQUIRE(N > 0 && C > 0 && HW > 0 && x && y && gamma && beta && running_mean && running_std,
               "dk_bn_fwd_infer: bad arguments");
    const int64_t total = (int64_t)N * C * HW;
    const int grid = stream_grid(total, BN_THREADS * 4);
    if (fuse_relu) bn_infer_kernel<true><<<grid, BN_THREADS, 0, as_stream(stream)>>>(x, y, gamma, beta, running_mean, running_std, total, C, HW);
    else bn_infer_kernel<false><<<grid, BN_THREADS, 0, as_stream(stream)>>>(x, y, gamma, beta, running_mean, running_std, total, C, HW);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_bn_bwd(const float *dy, const float *x, const float *gamma, const float *save_mean, const float *save_invstd,
              const float *save_scale, const float *save_shift, float *dx, float *dgamma, float *dbeta,
              int fuse_relu, int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = bn_check("dk_bn_bwd", N, C, HW, ws, ws_bytes);
    if (rc) return rc;
    DK_REQUIRE(dy && x && save_mean && save_invstd && save_scale && save_shift && dx && dgamma && dbeta,
               "dk_bn_bwd: NULL pointer");
    (void)gamma;  // gamma*invstd is save_scale
    cudaStream_t st = as_stream(stream);
    rc = bn_fused_bwd(dy, x, save_mean, save_invstd, save_scale, save_shift, dx, dgamma, dbeta, fuse_relu, N, C, HW, st);
    if (rc != DK_ERR_UNSUPPORTED) return rc;
    int S;
    int64_t per;
    bn_plan(N, C, HW, &S, &per);
    BnWs w = bn_ws_carve(ws, C);
    const bool vec = (HW % 4 == 0) && aligned16(x) && aligned16(dy) && aligned16(dx);
    dim3 grid(C, S);
    const int64_t total = (int64_t)N * C * HW;
    const int grid2 = stream_grid(vec ? total / 4 : total, BN_THREADS * 2);
#define DK_BN_BWD(V, R)                                                                                          \
    do {                                                                                                         \
        bn_bwd_reduce_kernel<V, R><<<grid, BN_THREADS, 0, st>>>(dy, x, N, C, HW, S, per, save_mean, save_invstd,   \
                                                                save_scale, save_shift, w.part, w.count, w.coef,  \
                                                                dgamma, dbeta);                                   \
        DK_LAUNCH_CHECK();                                                                                       \
        bn_bwd_dx_kernel<V, R><<<grid2, BN_THREADS, 0, st>>>(dy, x, dx, save_mean, save_invstd, save_scale,        \
                                                             save_shift, w.coef, total, C, HW);                   \
        DK_LAUNCH_CHECK();                                                                                       \
    } while (0)
    if (vec) {
        if (fuse_relu) DK_BN_BWD(true, true); else DK_BN_BWD(true, false);
    } else {
        if (fuse_relu) DK_BN_BWD(false, true); else DK_BN_BWD(false, false);
    }
#undef DK_BN_BWD
    return DK_OK;
}

int dk_bn_bwd_join(const float *dout, const float *out, const float *x, const float *gamma, const float *save_mean,
                   const float *save_invstd, const float *save_scale, const float *save_shift, float *dx, float *d_join,
                   float *dgamma, float *dbeta, int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = bn_check("dk_bn_bwd_join", N, C, HW, ws, ws_bytes);
    if (rc) return rc;
    DK_REQUIRE(dout && out && x && save_mean && save_invstd && save_scale && save_shift && dx && d_join && dgamma && dbeta,
               "dk_bn_bwd_join: NULL pointer");
    rc = bn_fused_bwd(dout, x, save_mean, save_invstd, save_scale, save_shift, dx, dgamma, dbeta, 0, N, C, HW, as_stream(stream),
                      out, d_join);
    if (rc != DK_ERR_UNSUPPORTED) return rc;
    // shapes the resident kernels do not take: the two passes the fusion replaces
    rc = dk_relu_bwd(dout, out, d_join, (int64_t)N * C * HW, stream);
    if (rc) return rc;
    return dk_bn_bwd(d_join, x, gamma, save_mean, save_invstd, save_scale, save_shift, dx, dgamma, dbeta, 0, N, C, HW, ws, ws_bytes,
                     stream);
}

int dk_bn_bwd_strided(const float *dy_sub, const float *x, const float *gamma, const float *save_mean,
                      const float *save_invstd, const float *save_scale, const float *save_shift, float *dx, float *dgamma,
                      float *dbeta, int fuse_relu, int N, int C, int H, int W, int stride, void *ws, size_t ws_bytes,
                      dk_stream_t stream) {
    int rc = bn_check("dk_bn_bwd_strided", N, C, H * W, ws, ws_bytes);
    if (rc) return rc;
    DK_REQUIRE(dy_sub && x && save_mean && save_invstd && save_scale && save_shift && dx && dgamma && dbeta,
               "dk_bn_bwd_strided: NULL pointer");
    DK_REQUIRE(stride == 2 && W % 8 == 0 && H % 2 == 0 && aligned16(dy_sub) && aligned16(x) && aligned16(dx),
               "dk_bn_bwd_strided: needs stride 2, W %% 8 == 0, H %% 2 == 0 and 16-byte aligned tensors (got stride %d, %dx%d)",
               stride, H, W);
    (void)gamma;
    cudaStream_t st = as_stream(stream);
    BnWs w = bn_ws_carve(ws, C);
    const int64_t total4 = (int64_t)N * (H / 2) * (W / 8);
    int64_t want = ceil_div((int64_t)sm_count() * 8, C);
    const int64_t max_by_work = ceil_div(total4, 4 * BN_THREADS);
    if (want > max_by_work) want = max_by_work;
    if (want > BN_MAX_SPLITS) want = BN_MAX_SPLITS;
    if (want < 1) want = 1;
    const int64_t per = ceil_div(total4, want);
    const int S = (int)ceil_div(total4, per);
    dim3 grid(C, S);
    const int64_t total = (int64_t)N * C * H * W;
    const int grid2 = stream_grid(total / 4, BN_THREADS * 2);
    if (fuse_relu) {
        bn_bwd_reduce_sub_kernel<true><<<grid, BN_THREADS, 0, st>>>(dy_sub, x, N, C, H, W, S, per, save_mean, save_invstd, save_scale,
                                                                    save_shift, w.part, w.count, w.coef, dgamma, dbeta);
        DK_LAUNCH_CHECK();
        bn_bwd_dx_sub_kernel<true><<<grid2, BN_THREADS, 0, st>>>(dy_sub, x, dx, save_mean, save_invstd, save_scale, save_shift,
                                                                 w.coef, total, C, H, W);
    } else {
        bn_bwd_reduce_sub_kernel<false><<<grid, BN_THREADS, 0, st>>>(dy_sub, x, N, C, H, W, S, per, save_mean, save_invstd, save_scale,
                                                                     save_shift, w.part, w.count, w.coef, dgamma, dbeta);
        DK_LAUNCH_CHECK();
        bn_bwd_dx_sub_kernel<false><<<grid2, BN_THREADS, 0, st>>>(dy_sub, x, dx, save_mean, save_invstd, save_scale, save_shift,
                                                                  w.coef, total, C, H, W);
    }
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// grads["bias"] = sum over (N, HW) -- same channel reduction, sum only (convolution.py:91-92)
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float *__restrict__ dy, float *__restrict__ dbias, int N, int F, int HW) {
    __shared__ float red[33];
    const int f = blockIdx.x;
    const int64_t total = (int64_t)N * HW;
    float s = 0.0f;
    for (int64_t v = threadIdx.x; v < total; v += blockDim.x) {
        const uint32_t n = (uint32_t)v / (uint32_t)HW, off = (uint32_t)v - n * (uint32_t)HW;
        s += dy[(n * F + f) * (int64_t)HW + off];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) dbias[f] = s;
}

int dk_bias_grad(const float *dy, float *dbias, int N, int F, int HW, void *ws, size_t ws_bytes, dk_stream_t stream) {
    (void)ws; (void)ws_bytes;
    DK_REQUIRE(N > 0 && F > 0 && HW > 0 && dy && dbias, "dk_bias_grad: bad arguments");
    bias_grad_kernel<<<F, 256, 0, as_stream(stream)>>>(dy, dbias, N, F, HW);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // extern "C"
