// dp_p2p.cu -- data-parallel gradient exchange fused into the optimiser kernel, over NVLink peer memory.
//
// The reference has no multi-GPU path (SURVEY §8e defines it: G replicas, gradients averaged, identical update).  With
// NCCL the 6 MB all-reduce of ResNet-18-depsep cost 0.3 ms of a 3.5 ms step on 2 x B200 (its kernels hold SMs while they
// wait for the peer, and the 148-CTA persistent kernels next to them lose a wave) -- for 40 us of wire time.  Here every
// rank's flat gradient buffer is cudaMalloc'ed, exported with cudaIpc and mapped by all peers; the optimiser kernel reads
// gradient element i from EVERY rank's buffer (its own from HBM, the others over NVLink), adds them in rank order (same
// bits on every replica) and applies the update: collective + update are ONE kernel, there is no reduced-gradient
// buffer and no NCCL launch on the step path.  Two flag handshakes per step (system-scope release / acquire on
// peer-mapped words) order it: "my gradients are complete" before anybody reads them, "I have read yours" before
// anybody overwrites them in the next backward.
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
__global__ void p2p_signal_kernel(const dk_p2p_ctx *__restrict__ ctx, int which) {
    const int lane = threadIdx.x;
    unsigned int e = 0;
    if (lane == 0) {
        e = *ctx->epoch + (which == 0 ? 1u : 0u);
        if (which == 0) *ctx->epoch = e;
    }
    e = __shfl_sync(0xffffffffu, e, 0);
    __threadfence_system();
    if (lane < ctx->world) st_release_sys((which == 0 ? ctx->ready[lane] : ctx->done[lane]) + ctx->rank, e);
}

// start of a step: nobody may still be reading the gradients the coming backward overwrites
__global__ void p2p_wait_done_kernel(const dk_p2p_ctx *__restrict__ ctx) {
    const int lane = threadIdx.x;
    if (lane < ctx->world) p2p_spin(ctx->done[ctx->rank] + lane, *ctx->epoch);
}

constexpr int OPTP_THREADS = 256;
constexpr int OPTP_CHUNK = OPTP_THREADS * 8;

template <int KIND>
__global__ void __launch_bounds__(OPTP_THREADS)
opt_multi_p2p_kernel(const dk_opt_tensor *__restrict__ table, const float *__restrict__ hyper,
                     const dk_p2p_ctx *__restrict__ ctx) {
    const float lr = hyper[0], hp = hyper[1], grad_scale = hyper[2];
    const dk_opt_tensor t = table[blockIdx.y];
    const int64_t start = (int64_t)blockIdx.x * OPTP_CHUNK;
    if (start >= t.n) return;
    const int world = ctx->world;
    if ((int)threadIdx.x < world) p2p_spin(ctx->ready[ctx->rank] + threadIdx.x, *ctx->epoch);
    __syncthreads();
    long long delta[P2P_MAX];
#pragma unroll
    for (int p = 0; p < P2P_MAX; ++p) delta[p] = p < world ? ctx->grad_delta[p] : 0;
    const int64_t end = start + OPTP_CHUNK < t.n ? start + OPTP_CHUNK : t.n;
    for (int64_t i = start + threadIdx.x; i < end; i += OPTP_THREADS) {
        float g = 0.0f;
#pragma unroll
        for (int p = 0; p < P2P_MAX; ++p) {  // rank order: the same sum, bit for bit, on every replica
            if (p < world) {
                const float *gp = reinterpret_cast<const float *>(reinterpret_cast<const char *>(t.grad + i) + delta[p]);
                float v;
                asm volatile("ld.global.relaxed.sys.f32 %0, [%1];" : "=f"(v) : "l"(gp) : "memory");
                g += v;
            }
        }
        g *= grad_scale;
        float w = t.param[i];
        if (KIND == 0) {
            w += -lr * g;
        } else if (KIND == 1) {
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
