// elementwise.cu -- streaming kernels: ReLU, residual join, mixup, pooling, loss, l2, optimisers.
// All are HBM-bound: 128-bit coalesced accesses, grid-stride over a whole number of waves.
#include "common.cuh"

namespace dk {

constexpr int EW_THREADS = 256;
constexpr int EW_UNROLL = 4;  // float4 per thread per iteration -> 64 B in flight per operand

// Generic 1/2-input, 1/2-output streaming map.  Op::apply works on scalars.
template <class Op, int NIN, int NOUT>
__global__ void __launch_bounds__(EW_THREADS)
ew_map_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ o0,
              float *__restrict__ o1, int64_t n, int vec_ok, Op op) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    int64_t done = 0;
    if (vec_ok) {
        const int64_t nvec = n >> 2;
        const int64_t step = nthreads * EW_UNROLL;
        for (int64_t base = tid; base < nvec; base += step) {
            float4 va[EW_UNROLL], vb[EW_UNROLL];
#pragma unroll
            for (int u = 0; u < EW_UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < nvec) {
                    va[u] = ld_stream4(a + 4 * i);
                    if (NIN > 1) vb[u] = ld_stream4(b + 4 * i);
                }
            }
#pragma unroll
            for (int u = 0; u < EW_UNROLL; ++u) {
                const int64_t i = base + (int64_t)u * nthreads;
                if (i < nvec) {
                    float4 r0, r1;
                    op.apply(va[u].x, NIN > 1 ? vb[u].x : 0.f, r0.x, r1.x);
                    op.apply(va[u].y, NIN > 1 ? vb[u].y : 0.f, r0.y, r1.y);
                    op.apply(va[u].z, NIN > 1 ? vb[u].z : 0.f, r0.z, r1.z);
                    op.apply(va[u].w, NIN > 1 ? vb[u].w : 0.f, r0.w, r1.w);
                    st_stream4(o0 + 4 * i, r0);
                    if (NOUT > 1) st_stream4(o1 + 4 * i, r1);
                }
            }
        }
        done = nvec << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        float r0, r1;
        op.apply(a[i], NIN > 1 ? b[i] : 0.f, r0, r1);
        o0[i] = r0;
        if (NOUT > 1) o1[i] = r1;
    }
}

template <class Op, int NIN, int NOUT>
static int launch_map(const float *a, const float *b, float *o0, float *o1, int64_t n, Op op, cudaStream_t st) {
    if (n <= 0) return DK_OK;
    const int vec_ok = aligned16(a) && (NIN < 2 || aligned16(b)) && aligned16(o0) && (NOUT < 2 || aligned16(o1));
    const int grid = stream_grid(n, EW_THREADS * 4 * EW_UNROLL);
    ew_map_kernel<Op, NIN, NOUT><<<grid, EW_THREADS, 0, st>>>(a, b, o0, o1, n, vec_ok, op);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

struct ReluFwdOp {
    __device__ __forceinline__ void apply(float x, float, float &y, float &m) const {
        const bool pos = x > 0.0f;  // NaN -> 0, as relu_cy.pyx:33
        y = pos ? x : 0.0f;
        m = pos ? 1.0f : 0.0f;
    }
};
struct ReluBwdOp {  // a = dy, b = y (or mask)
    __device__ __forceinline__ void apply(float dy, float y, float &dx, float &) const { dx = y > 0.0f ? dy : 0.0f; }
};
struct AddReluOp {
    __device__ __forceinline__ void apply(float a, float b, float &y, float &) const {
        const float s = a + b;
        y = s > 0.0f ? s : 0.0f;
    }
};
struct AddOp {
    __device__ __forceinline__ void apply(float a, float b, float &y, float &) const { y = a + b; }
};
struct MixupOp {
    float lam, oml;
    __device__ __forceinline__ void apply(float xa, float xb, float &y, float &) const { y = lam * xb + oml * xa; }
};

// ---- device-side input pipeline -----------------------------------------------------------------------
// The reference's loader turns every decoded uint8 HWC image into fp32 CHW minus 128 on the host
// (data_loading/image_preprocessor.py:36-37: astype(float32).transpose(2,0,1); im -= 128) and blends mixup pairs there
// too (image_data_loader.py:100-110).  Here the batch crosses PCIe as uint8 NHWC (a quarter of the bytes) and ONE kernel
// does transpose + conversion + offset (+ mixup): out[n,c,h,w] = (1-lam)*(xa[n,h,w,c] - sub) + lam*(xb[n,h,w,c] - sub).
// A thread owns one pixel: consecutive threads read consecutive C-byte groups and write consecutive floats of each plane.
__global__ void __launch_bounds__(256)
input_u8_nhwc_kernel(const unsigned char *__restrict__ xa, const unsigned char *__restrict__ xb, float *__restrict__ out,
                     float lam, float sub, long long pixels, int C, long long HW) {
    const float oml = 1.0f - lam;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / HW, hw = i - n * HW;
        const unsigned char *pa = xa + i * C;
        float *o = out + n * C * HW + hw;
        if (xb != nullptr) {
            const unsigned char *pb = xb + i * C;
            for (int c = 0; c < C; ++c) o[c * HW] = oml * ((float)pa[c] - sub) + lam * ((float)pb[c] - sub);
        } else {
            for (int c = 0; c < C; ++c) o[c * HW] = (float)pa[c] - sub;
        }
    }
}

// ---- global average pooling: one warp per (n, c) plane -----------------------------------------
__global__ void gap_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int planes, int HW, float inv) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < planes; p += gridDim.x * warps_per_block) {
        const float *src = x + (size_t)p * HW;
        float s = 0.0f;
        for (int i = lane; i < HW; i += 32) s += src[i];
        s = warp_sum(s);
        if (lane == 0) y[p] = s * inv;
    }
}
__global__ void gap_bwd_kernel(const float *__restrict__ dy, float *__restrict__ dx, int64_t total, int HW, float inv) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads)
        dx[i] = inv * dy[i / HW];
}

// ---- max pooling (s x s, stride s): one thread per output ---------------------------------------
template <bool TRAIN>
__global__ void maxpool_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int32_t *__restrict__ mask,
                                   int64_t total_out, int H, int W, int s) {
    const int OH = H / s, OW = W / s;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total_out; o += nthreads) {
        const int q = (int)(o % OW);
        const int p = (int)((o / OW) % OH);
        const int64_t plane = o / ((int64_t)OW * OH);
        const float *src = x + (plane * H + (int64_t)p * s) * W + (int64_t)q * s;
        float best = src[0];
        int r = 0, t = 0;
        for (int m = 0; m < s; ++m)
            for (int n2 = 0; n2 < s; ++n2) {
                const float v = src[(int64_t)m * W + n2];
                if (v > best) { best = v; r = m; t = n2; }  // strict '>': first maximum wins
            }
        y[o] = best;
        if (TRAIN) {
            int32_t *mdst = mask + (plane * H + (int64_t)p * s) * W + (int64_t)q * s;
            for (int m = 0; m < s; ++m)
                for (int n2 = 0; n2 < s; ++n2) mdst[(int64_t)m * W + n2] = (m == r && n2 == t) ? 1 : 0;
        }
    }
}
__global__ void maxpool_bwd_kernel(const int32_t *__restrict__ mask, const float *__restrict__ dy,
                                   float *__restrict__ dx, int64_t total_in, int H, int W, int s) {
    const int OH = H / s, OW = W / s;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_in; i += nthreads) {
        const int l = (int)(i % W);
        const int k = (int)((i / W) % H);
        const int64_t plane = i / ((int64_t)W * H);
        dx[i] = (mask[i] == 1) ? dy[(plane * OH + k / s) * OW + l / s] : 0.0f;
    }
}

// ---- softmax + cross entropy: one warp per row, single CTA (B x K is tiny) ----------------------
__global__ void __launch_bounds__(1024)
softmax_xent_fwd_kernel(const float *__restrict__ logits, const float *__restrict__ y, float *__restrict__ probs,
                        float *__restrict__ loss, int B, int K) {
    __shared__ float red[33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float acc = 0.0f;  // lane 0 of each warp accumulates -log(sum_j p_j y_j) over its rows
    for (int b = wid; b < B; b += nw) {
        const float *row = logits + (size_t)b * K;
        float s = 0.0f;
        for (int j = lane; j < K; j += 32) s += expf(row[j]);  // no max subtraction (losses.py:15-16)
        s = warp_sum(s);
        const float inv = 1.0f / s;
        float dot = 0.0f;
        for (int j = lane; j < K; j += 32) {
            const float p = inv * expf(row[j]);
            probs[(size_t)b * K + j] = p;
            if (y) dot += p * y[(size_t)b * K + j];
        }
        if (y) {
            dot = warp_sum(dot);
            if (lane == 0) acc += -logf(dot);
        }
    }
    if (y) {
        const float tot = block_sum(lane == 0 ? acc : 0.0f, red);
        if (threadIdx.x == 0) loss[0] = tot * (1.0f / (float)B);
    }
}
struct XentBwdOp {
    float invB;
    __device__ __forceinline__ void apply(float p, float y, float &dx, float &) const { dx = invB * (p - y); }
};

// ---- l2.forward: scale * sum(w^2), single CTA, fixed order ---------------------------------------
__global__ void __launch_bounds__(1024) sumsq_kernel(const float *__restrict__ w, float *__restrict__ out, float scale, int64_t n) {
    __shared__ float red[33];
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += w[i] * w[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = scale * s;
}

// all l2 terms of a step in one launch: block b reduces tensor b (fixed order) into its own result slot
__global__ void __launch_bounds__(1024) sumsq_multi_kernel(const dk_sumsq_task *__restrict__ tasks) {
    __shared__ float red[33];
    const dk_sumsq_task t = tasks[blockIdx.x];
    float s = 0.0f;
    int64_t done = 0;
    if (aligned16(t.w)) {  // 4 x 128-bit loads in flight per thread
        const int64_t nvec = t.n >> 2;
        const float4 *w4 = reinterpret_cast<const float4 *>(t.w);
        int64_t i = threadIdx.x;
        for (; i + 3 * 1024 < nvec; i += 4 * 1024) {
            const float4 a = w4[i], b = w4[i + 1024], c = w4[i + 2048], d = w4[i + 3072];
            s += ((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w)) + ((b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w)) +
                 ((c.x * c.x + c.y * c.y) + (c.z * c.z + c.w * c.w)) + ((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w));
        }
        for (; i < nvec; i += 1024) {
            const float4 a = w4[i];
            s += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
        }
        done = nvec << 2;
    }
    for (int64_t i = done + threadIdx.x; i < t.n; i += blockDim.x) s += t.w[i] * t.w[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) t.out[0] = t.scale * s;
}

// ---- fused multi-tensor optimisers -----------------------------------------------------------------
// grid = (chunks of the largest tensor, num_tensors); blocks past a tensor's end exit at once.
constexpr int OPT_THREADS = 256;
constexpr int OPT_CHUNK = OPT_THREADS * 8;

enum { OPT_SGD = 0, OPT_SGDM = 1, OPT_RMSPROP = 2 };

template <int KIND>
__global__ void __launch_bounds__(OPT_THREADS)
opt_multi_kernel(const dk_opt_tensor *__restrict__ table, float lr, float hp, float grad_scale,
                 const float *__restrict__ hyper) {
    if (hyper) {  // hyper-parameters live in device memory: a captured CUDA graph sees later changes
        lr = hyper[0];
        hp = hyper[1];
        grad_scale = hyper[2];
    }
    const dk_opt_tensor t = table[blockIdx.y];
    const int64_t start = (int64_t)blockIdx.x * OPT_CHUNK;
    if (start >= t.n) return;
    const int64_t end = start + OPT_CHUNK < t.n ? start + OPT_CHUNK : t.n;
    for (int64_t i = start + threadIdx.x; i < end; i += OPT_THREADS) {
        const float g = t.grad[i] * grad_scale;
        float w = t.param[i];
        if (KIND == OPT_SGD) {
            w += -lr * g;
        } else if (KIND == OPT_SGDM) {
            const float v = -lr * g + hp * t.state[i];
            w += v;
            t.state[i] = v;
        } else {
            const float c = hp * t.state[i] + (1.0f - hp) * (g * g);
            t.state[i] = c;
            w += -lr * g / sqrtf(c + 1e-5f);
        }
        t.param[i] = w;
    }
}

template <int KIND>
static int launch_opt(const dk_opt_tensor *table, int num_tensors, int64_t max_n, float lr, float hp,
                      float grad_scale, const float *hyper, cudaStream_t st) {
    if (num_tensors <= 0 || max_n <= 0) return DK_OK;
    DK_REQUIRE(table != nullptr, "optimiser: NULL tensor table");
    DK_REQUIRE(num_tensors <= 65535, "optimiser: too many tensors (%d)", num_tensors);
    dim3 grid((unsigned)ceil_div(max_n, OPT_CHUNK), (unsigned)num_tensors);
    opt_multi_kernel<KIND><<<grid, OPT_THREADS, 0, st>>>(table, lr, hp, grad_scale, hyper);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_relu_fwd(const float *x, float *y, float *mask, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && (n == 0 || (x && y)), "dk_relu_fwd: bad arguments");
    if (mask) return launch_map<ReluFwdOp, 1, 2>(x, nullptr, y, mask, n, ReluFwdOp{}, as_stream(stream));
    return launch_map<ReluFwdOp, 1, 1>(x, nullptr, y, nullptr, n, ReluFwdOp{}, as_stream(stream));
}

int dk_relu_bwd(const float *dy, const float *y, float *dx, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && (n == 0 || (dy && y && dx)), "dk_relu_bwd: bad arguments");
    return launch_map<ReluBwdOp, 2, 1>(dy, y, dx, nullptr, n, ReluBwdOp{}, as_stream(stream));
}

int dk_add_relu_fwd(const float *a, const float *b, float *y, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && (n == 0 || (a && b && y)), "dk_add_relu_fwd: bad arguments");
    return launch_map<AddReluOp, 2, 1>(a, b, y, nullptr, n, AddReluOp{}, as_stream(stream));
}

int dk_add(const float *a, const float *b, float *out, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && (n == 0 || (a && b && out)), "dk_add: bad arguments");
    return launch_map<AddOp, 2, 1>(a, b, out, nullptr, n, AddOp{}, as_stream(stream));
}

int dk_mixup(const float *xa, const float *xb, float *out, float lam, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && (n == 0 || (xa && xb && out)), "dk_mixup: bad arguments");
    return launch_map<MixupOp, 2, 1>(xa, xb, out, nullptr, n, MixupOp{lam, 1.0f - lam}, as_stream(stream));
}

int dk_input_u8_nhwc(const unsigned char *xa, const unsigned char *xb, float *out, float lam, float sub, int N, int C, int H,
                     int W, dk_stream_t stream) {
    DK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && xa && out, "dk_input_u8_nhwc: bad arguments");
    const long long pixels = (long long)N * H * W;
    input_u8_nhwc_kernel<<<stream_grid(pixels, 256), 256, 0, as_stream(stream)>>>(xa, xb, out, xb ? lam : 0.0f, sub, pixels, C,
                                                                                 (long long)H * W);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_gap_fwd(const float *x, float *y, int N, int C, int HW, dk_stream_t stream) {
    DK_REQUIRE(N >= 0 && C > 0 && HW > 0, "dk_gap_fwd: bad shape");
    const int planes = N * C;
    if (planes == 0) return DK_OK;
    const int grid = stream_grid(planes, 8);
    gap_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, planes, HW, 1.0f / (float)HW);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_gap_bwd(const float *dy, float *dx, int N, int C, int HW, dk_stream_t stream) {
    DK_REQUIRE(N >= 0 && C > 0 && HW > 0, "dk_gap_bwd: bad shape");
    const int64_t total = (int64_t)N * C * HW;
    if (total == 0) return DK_OK;
    gap_bwd_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(dy, dx, total, HW, 1.0f / (float)HW);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

static int maxpool_check(int N, int C, int H, int W, int s, const char *who) {
    DK_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && s > 0, "%s: bad shape", who);
    // the reference reads out of bounds otherwise (layers/pooling_cy.pyx:54-58)
    DK_REQUIRE(H % s == 0 && W % s == 0, "%s: H=%d and W=%d must be divisible by the pooling stride %d", who, H, W, s);
    return DK_OK;
}

int dk_maxpool_fwd(const float *x, float *y, int N, int C, int H, int W, int s, dk_stream_t stream) {
    int rc = maxpool_check(N, C, H, W, s, "dk_maxpool_fwd");
    if (rc) return rc;
    const int64_t total = (int64_t)N * C * (H / s) * (W / s);
    if (total == 0) return DK_OK;
    maxpool_fwd_kernel<false><<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, nullptr, total, H, W, s);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_maxpool_fwd_train(const float *x, float *y, int32_t *mask, int N, int C, int H, int W, int s, dk_stream_t stream) {
    int rc = maxpool_check(N, C, H, W, s, "dk_maxpool_fwd_train");
    if (rc) return rc;
    DK_REQUIRE(mask != nullptr, "dk_maxpool_fwd_train: NULL mask");
    const int64_t total = (int64_t)N * C * (H / s) * (W / s);
    if (total == 0) return DK_OK;
    maxpool_fwd_kernel<true><<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(x, y, mask, total, H, W, s);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_maxpool_bwd(const int32_t *mask, const float *dy, float *dx, int N, int C, int H, int W, int s, dk_stream_t stream) {
    int rc = maxpool_check(N, C, H, W, s, "dk_maxpool_bwd");
    if (rc) return rc;
    const int64_t total = (int64_t)N * C * H * W;
    if (total == 0) return DK_OK;
    maxpool_bwd_kernel<<<stream_grid(total, 256), 256, 0, as_stream(stream)>>>(mask, dy, dx, total, H, W, s);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_softmax_xent_fwd(const float *logits, const float *y_one_hot, float *probs, float *loss, int B, int K,
                        dk_stream_t stream) {
    DK_REQUIRE(B > 0 && K > 0 && logits && probs, "dk_softmax_xent_fwd: bad arguments");
    DK_REQUIRE(!y_one_hot || loss, "dk_softmax_xent_fwd: labels given but loss is NULL");
    int threads = B * 32;
    threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
    softmax_xent_fwd_kernel<<<1, threads, 0, as_stream(stream)>>>(logits, y_one_hot, probs, loss, B, K);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_softmax_xent_bwd(const float *probs, const float *y_one_hot, float *dx, int B, int K, dk_stream_t stream) {
    DK_REQUIRE(B > 0 && K > 0 && probs && y_one_hot && dx, "dk_softmax_xent_bwd: bad arguments");
    return launch_map<XentBwdOp, 2, 1>(probs, y_one_hot, dx, nullptr, (int64_t)B * K, XentBwdOp{1.0f / (float)B},
                                      as_stream(stream));
}

int dk_sumsq(const float *w, float *out, float scale, int64_t n, dk_stream_t stream) {
    DK_REQUIRE(n >= 0 && out && (n == 0 || w), "dk_sumsq: bad arguments");
    sumsq_kernel<<<1, 1024, 0, as_stream(stream)>>>(w, out, scale, n);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_sumsq_multi(const dk_sumsq_task *tasks, int num_tasks, dk_stream_t stream) {
    if (num_tasks <= 0) return DK_OK;
    DK_REQUIRE(tasks != nullptr && num_tasks <= 65535, "dk_sumsq_multi: bad arguments");
    sumsq_multi_kernel<<<num_tasks, 1024, 0, as_stream(stream)>>>(tasks);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_opt_sgd_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n, float lr, float grad_scale,
                     const float *hyper, dk_stream_t stream) {
    return launch_opt<OPT_SGD>(table, num_tensors, max_n, lr, 0.0f, grad_scale, hyper, as_stream(stream));
}
int dk_opt_sgdm_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n, float lr, float momentum,
                      float grad_scale, const float *hyper, dk_stream_t stream) {
    return launch_opt<OPT_SGDM>(table, num_tensors, max_n, lr, momentum, grad_scale, hyper, as_stream(stream));
}
int dk_opt_rmsprop_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n, float lr, float decay,
                         float grad_scale, const float *hyper, dk_stream_t stream) {
    return launch_opt<OPT_RMSPROP>(table, num_tensors, max_n, lr, decay, grad_scale, hyper, as_stream(stream));
}

}  // extern "C"
