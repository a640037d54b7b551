// dp_p2p.cu -- data-parallel gradient exchange fused into the optimiser kernel, over NVLink peer memory.
//
// The reference has no multi-GPU path (SURVEY §8e defines it: G replicas, gradients averaged, identical update).  With
// NCCL the 6 MB all-reduce of ResNet-18-depsep cost 0.3 ms of a 3.5 ms step on 2 x B200 (its kernels hold SMs while they
// wait for the peer, and the 148-CTA persistent kernels next to them lose a wave) -- for 40 us of wire time.  Here every
// rank's flat gradient buffer is cudaMalloc'ed, exported with cudaIpc and mapped by all peers, and the exchange is two
// kernels of this library, with no reduced-gradient buffer and no NCCL launch on the step path:
//   1. reduce-scatter, in place: rank r sums slice r of the flat buffer over all ranks (its own from HBM, the others over
//      NVLink with 16-byte loads, in rank order) and writes the sum back into ITS OWN slice r;
//   2. all-gather fused into the optimiser: the update kernel reads gradient element i from its owner's buffer (one load,
//      16 bytes wide) and applies SGD / SGDMomentum / RMSProp -- every replica reads the same bits.
// Per rank that is 2 (G-1)/G x the gradient bytes over NVLink instead of the (G-1) x of the first version (every rank
// summing every element of every peer with 4-byte loads: +98 us per step at 8 GPUs).  Three flag handshakes per step
// (system-scope release / acquire on peer-mapped words) order it: "my gradients are complete" before anybody reads them,
// "my slice is reduced" before anybody gathers it, "I have read everything" before the next backward overwrites them.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dk {

constexpr int P2P_MAX = 8;

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p has reached epoch e (wrap-safe); a dead peer must surface as a trapped kernel, never as a hung GPU
__device__ __forceinline__ void p2p_spin(const unsigned int *p, unsigned int e) {
    if ((int)(ld_acquire_sys(p) - e) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - e) < 0) {
        if (clock64() - t0 > 60000000000LL) {  // ~30 s
            printf("dorknet_b200: peer flag wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}

// which = 0: epoch += 1, then ready[p][rank] = epoch on every rank p ("my gradients of this step are complete")
// which = 1: done[p][rank] = epoch on every rank p ("I have read everybody's gradients of this step")
// which = 2: reduced[p][rank] = epoch ("slice `rank` of my buffer holds the sum over all ranks")
__global__ void p2p_signal_kernel(const dk_p2p_ctx *__restrict__ ctx, int which) {
    const int lane = threadIdx.x;
    unsigned int e = 0;
    if (lane == 0) {
        e = *ctx->epoch + (which == 0 ? 1u : 0u);
        if (which == 0) *ctx->epoch = e;
    }
    e = __shfl_sync(0xffffffffu, e, 0);
    __threadfence_system();
    if (lane < ctx->world)
        st_release_sys((which == 0 ? ctx->ready[lane] : which == 1 ? ctx->done[lane] : ctx->reduced[lane]) + ctx->rank, e);
}

__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float *p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v;
    asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// step 1: my slice of the flat gradient buffer, summed over all ranks in rank order, written back in place.  Two 16-byte
// loads per peer in flight per thread.
constexpr int RS_THREADS = 256;
__global__ void __launch_bounds__(RS_THREADS)
p2p_reduce_slice_kernel(const dk_p2p_ctx *__restrict__ ctx) {
    const int world = ctx->world, rank = ctx->rank;
    if ((int)threadIdx.x < world) p2p_spin(ctx->ready[rank] + threadIdx.x, *ctx->epoch);
    __syncthreads();
    long long delta[P2P_MAX];
#pragma unroll
    for (int p = 0; p < P2P_MAX; ++p) delta[p] = p < world ? ctx->grad_delta[p] : 0;
    const long long lo = (long long)rank * ctx->slice;
    long long hi = lo + ctx->slice;
    if (hi > ctx->nfloats) hi = ctx->nfloats;  // (nfloats is a multiple of 4: tensors are 128-byte aligned in the buffer)
    float *base = ctx->grad_base;
    const long long stride = (long long)gridDim.x * RS_THREADS * 4;
    for (long long i = lo + ((long long)blockIdx.x * RS_THREADS + threadIdx.x) * 4; i < hi; i += 2 * stride) {
        const bool two = i + stride < hi;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
        for (int p = 0; p < P2P_MAX; ++p) {
            if (p < world) {
                const float *src = reinterpret_cast<const float *>(reinterpret_cast<const char *>(base + i) + delta[p]);
                const float4 u = ld_relaxed_sys_v4(src);
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (two) w = ld_relaxed_sys_v4(src + stride);
                a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
                b.x += w.x; b.y += w.y; b.z += w.z; b.w += w.w;
            }
        }
        *reinterpret_cast<float4 *>(base + i) = a;
        if (two) *reinterpret_cast<float4 *>(base + i + stride) = b;
    }
}

// start of a step: nobody may still be reading the gradients the coming backward overwrites
__global__ void p2p_wait_done_kernel(const dk_p2p_ctx *__restrict__ ctx) {
    const int lane = threadIdx.x;
    if (lane < ctx->world) p2p_spin(ctx->done[ctx->rank] + lane, *ctx->epoch);
}

constexpr int OPTP_THREADS = 256;
constexpr int OPTP_CHUNK = OPTP_THREADS * 8;

__device__ __forceinline__ float opt_apply(int kind, float w, float g, float *state, float lr, float hp) {
    if (kind == 0) return w - lr * g;
    if (kind == 1) {
        const float v = -lr * g + hp * *state;
        *state = v;
        return w + v;
    }
    const float c = hp * *state + (1.0f - hp) * (g * g);
    *state = c;
    return w - lr * g / sqrtf(c + 1e-5f);
}

// step 2: the optimiser update; gradient element i comes from the rank that owns (reduced) its slice
template <int KIND>
__global__ void __launch_bounds__(OPTP_THREADS)
opt_multi_p2p_kernel(const dk_opt_tensor *__restrict__ table, const float *__restrict__ hyper,
                     const dk_p2p_ctx *__restrict__ ctx) {
    const float lr = hyper[0], hp = hyper[1], grad_scale = hyper[2];
    const dk_opt_tensor t = table[blockIdx.y];
    const int64_t start = (int64_t)blockIdx.x * OPTP_CHUNK;
    if (start >= t.n) return;
    const int world = ctx->world;
    if ((int)threadIdx.x < world) p2p_spin(ctx->reduced[ctx->rank] + threadIdx.x, *ctx->epoch);
    __syncthreads();
    const int64_t end = start + OPTP_CHUNK < t.n ? start + OPTP_CHUNK : t.n;
    const long long flat0 = t.grad - ctx->grad_base;  // this tensor's offset in the flat buffer (a multiple of 32 floats)
    const long long slice = ctx->slice;
    const bool vec = (flat0 & 3) == 0 && (reinterpret_cast<uintptr_t>(t.param) & 15u) == 0 &&
                     (KIND == 0 || (reinterpret_cast<uintptr_t>(t.state) & 15u) == 0);
    if (vec) {
        // 4 elements per thread and pass: they share an owner (slices and tensor offsets are multiples of 4 floats)
        for (int64_t i = start + 4 * (int64_t)threadIdx.x; i < end; i += 4 * OPTP_THREADS) {
            const int owner = (int)((flat0 + i) / slice);
            const float *gp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(t.grad + i) + ctx->grad_delta[owner]);
            if (i + 4 <= end) {
                const float4 g = ld_relaxed_sys_v4(gp);
                float4 w = *reinterpret_cast<const float4 *>(t.param + i);
                float4 st = KIND == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4 *>(t.state + i);
                w.x = opt_apply(KIND, w.x, g.x * grad_scale, &st.x, lr, hp);
                w.y = opt_apply(KIND, w.y, g.y * grad_scale, &st.y, lr, hp);
                w.z = opt_apply(KIND, w.z, g.z * grad_scale, &st.z, lr, hp);
                w.w = opt_apply(KIND, w.w, g.w * grad_scale, &st.w, lr, hp);
                *reinterpret_cast<float4 *>(t.param + i) = w;
                if (KIND != 0) *reinterpret_cast<float4 *>(t.state + i) = st;
            } else {
                for (int64_t k = i; k < end; ++k) {
                    float st = KIND == 0 ? 0.0f : t.state[k];
                    t.param[k] = opt_apply(KIND, t.param[k], ld_relaxed_sys(gp + (k - i)) * grad_scale, &st, lr, hp);
                    if (KIND != 0) t.state[k] = st;
                }
            }
        }
        return;
    }
    for (int64_t i = start + threadIdx.x; i < end; i += OPTP_THREADS) {
        const int owner = (int)((flat0 + i) / slice);
        const float *gp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(t.grad + i) + ctx->grad_delta[owner]);
        float st = KIND == 0 ? 0.0f : t.state[i];
        t.param[i] = opt_apply(KIND, t.param[i], ld_relaxed_sys(gp) * grad_scale, &st, lr, hp);
        if (KIND != 0) t.state[i] = st;
    }
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_p2p_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
    DK_REQUIRE(bytes > 0 && ptr && handle64, "dk_p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    DK_CUDA(cudaMalloc(&p, bytes));
    DK_CUDA(cudaMemset(p, 0, bytes));
    DK_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    DK_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *ptr = p;
    return DK_OK;
}

int dk_p2p_open(const unsigned char *handle64, void **ptr) {
    DK_REQUIRE(handle64 && ptr, "dk_p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    DK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return DK_OK;
}

int dk_p2p_close(void *ptr) {
    if (ptr) DK_CUDA(cudaIpcCloseMemHandle(ptr));
    return DK_OK;
}

int dk_p2p_free(void *ptr) {
    if (ptr) DK_CUDA(cudaFree(ptr));
    return DK_OK;
}

int dk_p2p_wait_done(const dk_p2p_ctx *ctx, dk_stream_t stream) {
    DK_REQUIRE(ctx != nullptr, "dk_p2p_wait_done: NULL context");
    p2p_wait_done_kernel<<<1, 32, 0, as_stream(stream)>>>(ctx);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_opt_multi_p2p(int kind, const dk_opt_tensor *table, int num_tensors, int64_t max_n, const float *hyper,
                     const dk_p2p_ctx *ctx, dk_stream_t stream) {
    DK_REQUIRE(kind >= 0 && kind <= 2 && hyper && ctx, "dk_opt_multi_p2p: bad arguments");
    if (num_tensors <= 0 || max_n <= 0) return DK_OK;
    DK_REQUIRE(table != nullptr && num_tensors <= 65535, "dk_opt_multi_p2p: bad tensor table");
    cudaStream_t st = as_stream(stream);
    p2p_signal_kernel<<<1, 32, 0, st>>>(ctx, 0);
    DK_LAUNCH_CHECK();
    // (the slice length lives in device memory: size the grid for the largest slice a rank can own, world >= 2)
    p2p_reduce_slice_kernel<<<sm_count(), RS_THREADS, 0, st>>>(ctx);
    DK_LAUNCH_CHECK();
    p2p_signal_kernel<<<1, 32, 0, st>>>(ctx, 2);
    DK_LAUNCH_CHECK();
    dim3 grid((unsigned)ceil_div(max_n, OPTP_CHUNK), (unsigned)num_tensors);
    if (kind == 0) opt_multi_p2p_kernel<0><<<grid, OPTP_THREADS, 0, st>>>(table, hyper, ctx);
    else if (kind == 1) opt_multi_p2p_kernel<1><<<grid, OPTP_THREADS, 0, st>>>(table, hyper, ctx);
    else opt_multi_p2p_kernel<2><<<grid, OPTP_THREADS, 0, st>>>(table, hyper, ctx);
    DK_LAUNCH_CHECK();
    p2p_signal_kernel<<<1, 32, 0, st>>>(ctx, 1);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // extern "C"
