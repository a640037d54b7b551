// mma_rate.cu -- what a tcgen05.mma kind::tf32 (M = 128, K = 8) costs on B200, one CTA per SM, fixed (uninitialised)
// shared-memory operands.  A measurement tool for DESIGN.md, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I dorknet_b200/csrc -o tools/micro/mma_rate.bin tools/micro/mma_rate.cu
// part 1: back-to-back MMAs as a function of N (issue-to-issue and issue-to-completion)
// part 2: N = 192, steps of 8 MMAs with the per-step work of a pipelined kernel added piece by piece:
//   bit 0  tcgen05.commit onto an mbarrier after every step
//   bit 1  wait (mbarrier.try_wait) for the commit of two steps ago before issuing a step
//   bit 2  tcgen05.fence::after_thread_sync after the wait
//   bit 3  three more warps spin on an mbarrier that never completes (pollers next to the tensor pipe's operand reads)
//   bit 4  a second warp relays: it waits for the commit and arrives on the barrier the issuer waits for (one more hop)
#include <cstdio>
#include <cstdlib>
#include "tc_ptx.cuh"
using namespace dk::tc;

__global__ void __launch_bounds__(192, 1) rate_kernel(int N, int count, int mode, long long *out, int acc1 = 256) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bar = base + 200 * 1024;  // [0] final, [1] never, [2..3] commit ring, [4..5] relay ring
    const uint32_t slot = bar + 64;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bar + 8 * i, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    volatile int *stop = reinterpret_cast<volatile int *>(smem + (bar + 128 - smem_u32(smem)));
    if (threadIdx.x == 0) *stop = 0;
    __syncthreads();
    const int steps = count / 8;
    if (warp == 1) {
        const uint32_t idesc = idesc_tf32(128, N, 0, 0);
        const uint32_t hi = smem_desc_hi(1024u, LAYOUT_SW128);
        const uint32_t a_lo = smem_desc_lo(base, 16u), b_lo = smem_desc_lo(base + 64 * 1024, 16u);
        const uint32_t wbase = (mode & 16) ? bar + 32 : bar + 16;
        const long long t0 = clock64();
        for (int s = 0; s < steps; ++s) {
            if ((mode & 2) && s >= 2) mbar_wait(wbase + 8 * (s & 1), (uint32_t)((s - 2) >> 1) & 1u);
            if (mode & 4) tc_fence_after();
            if (elect_one()) {
                mma_tf32_k4(tmem_base, a_lo, hi, b_lo, hi, 2u, 2u, idesc, 1u);
                mma_tf32_k4(tmem_base + (uint32_t)acc1, a_lo + 512, hi, b_lo, hi, 2u, 2u, idesc, 1u);
            }
            if ((mode & 1) && elect_one()) mma_commit(bar + 16 + 8 * (s & 1));
        }
        const long long t1 = clock64();
        if (elect_one()) mma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        *stop = 1;
        if (threadIdx.x == 32 && blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    } else if (warp == 2 && (mode & 16)) {
        for (int s = 0; s < steps; ++s) {
            mbar_wait(bar + 16 + 8 * (s & 1), (uint32_t)(s >> 1) & 1u);
            __syncwarp();
            if (threadIdx.x == 64) mbar_arrive(bar + 32 + 8 * (s & 1));
        }
    } else if (warp >= 3 && (mode & 8)) {
        while (!*stop) {
            if (mbar_try_wait(bar + 8, 0)) break;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    const int count = 4096;
    auto run = [&](int N, int mode, int acc1 = 256) {
        for (int rep = 0; rep < 2; ++rep) {
            rate_kernel<<<148, 192, 206 * 1024>>>(N, count, mode, out, acc1);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
                exit(1);
            }
        }
    };
    for (int N : {16, 32, 64, 96, 128, 160, 192, 224, 256}) {
        run(N, 0);
        printf("N %3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA  (math floor at 2048 FMA/clk: %.0f)\n", N, (double)out[0] / count,
               (double)out[1] / count, 128.0 * N * 8 / 2048);
    }
    for (int mode : {0, 1, 3, 7, 8, 9, 11, 15, 19, 23, 31}) {
        run(192, mode);
        printf("N 192 mode %2d: issue %.1f cycles/MMA, complete %.1f cycles/MMA = %.0f cycles per step of 8\n", mode,
               (double)out[0] / count, (double)out[1] / count, (double)out[1] / count * 8);
    }
    for (int acc1 : {256, 192, 200, 224, 128, 64, 0}) {
        run(192, 0, acc1);
        printf("N 192, second accumulator at column %3d: complete %.1f cycles/MMA\n", acc1, (double)out[1] / count);
    }
    for (int acc1 : {128, 96, 64}) {
        run(64, 0, acc1);
        printf("N  64, second accumulator at column %3d: complete %.1f cycles/MMA\n", acc1, (double)out[1] / count);
    }
    return 0;
}
