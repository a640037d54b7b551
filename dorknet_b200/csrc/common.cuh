// common.cuh -- shared helpers for libdorknet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dorknet_b200.h"

namespace dk {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const char *fmt, ...);
int sm_count();
int gemm_backend();
void count_launch();  // every kernel launch site goes through DK_LAUNCH_CHECK, which counts it

#define DK_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            dk::set_error(__VA_ARGS__);       \
            return DK_ERR_INVALID;            \
        }                                     \
    } while (0)

#define DK_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DK_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define DK_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        dk::count_launch();                                                                    \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            dk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return DK_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

static inline cudaStream_t as_stream(dk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for a grid-stride streaming kernel: enough CTAs to cover the work, capped at a whole
// number of waves over the 148 SMs.
static inline int stream_grid(int64_t work_items, int per_block, int ctas_per_sm = 8) {
    int64_t need = ceil_div(work_items, per_block);
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static __host__ __device__ inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__

// streaming (read-once) 128-bit load / store: keep L1 for data with reuse
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); result valid in every thread.
// `red` must hold >= 33 floats of shared memory.  Fixed order -> deterministic.
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        float t = lane < nw ? red[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

#endif  // __CUDACC__

}  // namespace dk
