// bulk_bw_probe.cu -- how fast can a persistent CTA stream HBM through shared memory with cp.async.bulk?
//   mode 0: bulk G2S only (data discarded)            -> read bandwidth of the bulk-copy path
//   mode 1: bulk G2S, then LDS.128 + st.global.v4     -> read + write, stores from registers
//   mode 2: bulk G2S, then bulk S2G                   -> read + write, both through the async proxy
//   mode 3: plain ld.global.v4 / st.global.v4 copy    -> reference
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_bw_probe tools/bulk_bw_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256)
probe_kernel(const float *__restrict__ src, float *__restrict__ dst, long long chunks, int chunk_floats, int nst, int split) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bars[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t cb = (uint32_t)chunk_floats * 4u;
    auto issue = [&](long long ch, int slot) {
        const uint32_t bar = smem_u32(&bars[slot]);
        mbar_expect_tx(bar, cb);
        const uint32_t piece = cb / split;
        for (int q = 0; q < split; ++q)
            bulk_g2s(smem_u32(sm + (size_t)slot * chunk_floats) + q * piece, (const char *)(src + ch * chunk_floats) + q * piece, piece, bar);
    };
    long long mine = 0;
    for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) ++mine;
    if (threadIdx.x == 0)
        for (int i = 0; i < nst && i < mine; ++i) issue(blockIdx.x + (long long)i * gridDim.x, i);
    float sink = 0.f;
    for (long long i = 0; i < mine; ++i) {
        const int slot = (int)(i % nst);
        const long long ch = blockIdx.x + i * gridDim.x;
        mbar_wait(smem_u32(&bars[slot]), (uint32_t)(i / nst) & 1u);
        const float *s = sm + (size_t)slot * chunk_floats;
        if (MODE == 0) {
            sink += s[threadIdx.x];
        } else if (MODE == 1) {
            float *o = dst + ch * chunk_floats;
            for (int j = 4 * threadIdx.x; j < chunk_floats; j += 4 * 256) {
                const float4 v = *reinterpret_cast<const float4 *>(s + j);
                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + j), "f"(v.x * 2.f), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
        } else if (MODE == 2) {
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_s2g(dst + ch * chunk_floats, smem_u32(s), cb);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && i + nst < mine) issue(blockIdx.x + (i + nst) * gridDim.x, slot);
    }
    if (MODE == 2 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (sink == 123.456f) dst[0] = sink;
}

__global__ void __launch_bounds__(256) plain_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n4) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256 * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long j = i + (long long)u * gridDim.x * 256;
            if (j < n4) v[u] = src[j];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long j = i + (long long)u * gridDim.x * 256;
            if (j < n4) dst[j] = v[u];
        }
    }
}

template <int MODE>
static float run(const float *src, float *dst, long long total_floats, int chunk_floats, int nst, int ctas_per_sm, int split) {
    const long long chunks = total_floats / chunk_floats;
    const size_t smem = (size_t)nst * chunk_floats * 4;
    cudaFuncSetAttribute(probe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) probe_kernel<MODE><<<grid, 256, smem>>>(src, dst, chunks, chunk_floats, nst, split);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) probe_kernel<MODE><<<grid, 256, smem>>>(src, dst, chunks, chunk_floats, nst, split);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return ms / reps;
}

int main() {
    const long long total = 128ll << 20;  // 128 Mi floats = 512 MB per tensor (>> L2)
    float *src, *dst;
    cudaMalloc(&src, total * 4);
    cudaMalloc(&dst, total * 4);
    cudaMemset(src, 1, total * 4);
    cudaMemset(dst, 0, total * 4);
    {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        plain_copy<<<148 * 8, 256>>>((const float4 *)src, (float4 *)dst, total / 4);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) plain_copy<<<148 * 8, 256>>>((const float4 *)src, (float4 *)dst, total / 4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("plain copy: %.1f GB/s (read+write)\n", 2.0 * total * 4 / (ms / 10 * 1e-3) / 1e9);
    }
    const int cfgs[][4] = {  // chunk KB, stages, CTAs/SM, split
        {12, 4, 2, 1}, {12, 8, 2, 1}, {24, 4, 2, 1}, {48, 2, 2, 1}, {48, 4, 1, 1}, {12, 4, 4, 1}, {6, 8, 4, 1}, {24, 8, 1, 1},
        {48, 2, 2, 4}, {48, 4, 1, 4}, {96, 2, 1, 8}, {24, 2, 4, 2},
    };
    for (auto &c : cfgs) {
        const int cf = c[0] * 256;
        const float r0 = run<0>(src, dst, total, cf, c[1], c[2], c[3]);
        const float r1 = run<1>(src, dst, total, cf, c[1], c[2], c[3]);
        const float r2 = run<2>(src, dst, total, cf, c[1], c[2], c[3]);
        printf("chunk %3d KB stages %d ctas/SM %d split %d: bulk read %.0f GB/s | read + st.global %.0f GB/s | read + bulk store %.0f GB/s\n",
               c[0], c[1], c[2], c[3], total * 4 / (r0 * 1e-3) / 1e9, 2.0 * total * 4 / (r1 * 1e-3) / 1e9,
               2.0 * total * 4 / (r2 * 1e-3) / 1e9);
    }
    return 0;
}
