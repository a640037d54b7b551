// Stand-alone TMA behaviour probe (bring-up tool): one box load per case, reports whether the mbarrier completed and a
// checksum of what landed.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tm, int rank, int c0, int c1, int c2, int c3, uint32_t bytes,
                      float *out, int *status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t dst = (smem_u32(smem) + 1023u) & ~1023u, b = smem_u32(&bar);
    float *s = reinterpret_cast<float *>(smem + (dst - smem_u32(smem)));
    for (int i = threadIdx.x; i < (int)(bytes / 4); i += blockDim.x) s[i] = -777.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
        if (rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dst), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(dst), "l"(&tm), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        long long t0 = clock64();
        uint32_t ok = 0;
        while (!ok && clock64() - t0 < 200000000LL) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(b) : "memory");
        }
        *status = (int)ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(bytes / 4); i += blockDim.x) out[i] = s[i];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Enc enc = (Enc)fn;
    const int W = 64, H = 8, C = 32, N = 2;
    std::vector<float> h((size_t)N * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *out;
    int *st;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&out, 65536);
    cudaMalloc(&st, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    struct Case { const char *name; int rank; cuuint64_t dims[4]; cuuint64_t str[3]; cuuint32_t box[4]; int c[4]; int sw; };
#define D3 {W, H, (cuuint64_t)N * C, 1}, {W * 4, H * W * 4, 0}
    Case all_cases[] = {
        {"atom32B box(32,1,32) c0=0 in-bounds", 3, D3, {32, 1, 32, 1}, {0, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c0=8 aligned in-bounds", 3, D3, {32, 1, 32, 1}, {8, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c0=2 unaligned in-bounds", 3, D3, {32, 1, 32, 1}, {2, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c0=4 16B-aligned in-bounds", 3, D3, {32, 1, 32, 1}, {4, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c0=40 aligned partial OOB", 3, D3, {32, 1, 32, 1}, {40, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c0=-8 aligned partial OOB", 3, D3, {32, 1, 32, 1}, {-8, 0, 0, 0}, 4},
        {"atom32B box(32,1,32) c1=-1 fully OOB", 3, D3, {32, 1, 32, 1}, {0, -1, 0, 0}, 4},
        {"sw128 box(32,1,32) c0=2 unaligned in-bounds", 3, D3, {32, 1, 32, 1}, {2, 0, 0, 0}, 3},
        {"sw128 box(32,1,32) c0=40 partial OOB", 3, D3, {32, 1, 32, 1}, {40, 0, 0, 0}, 3},
        {"sw128 box(32,1,32) c0=-1 unaligned partial OOB", 3, D3, {32, 1, 32, 1}, {-1, 0, 0, 0}, 3},
        {"atom32B 4d (W,H,C,N) box(32,1,32,1) in-bounds", 4, {W, H, C, N}, {W * 4, H * W * 4, (cuuint64_t)C * H * W * 4}, {32, 1, 32, 1}, {0, 0, 0, 0}, 4},
        {"noswizzle box(40,3,8) c0=-1 c1=-1 halo tile", 3, D3, {40, 3, 8, 1}, {-1, -1, 0, 0}, 0},
        {"noswizzle box(36,3,8) c0=3 unaligned in-bounds", 3, D3, {36, 3, 8, 1}, {3, 1, 0, 0}, 0},
    };
    const int ncases = (int)(sizeof(all_cases) / sizeof(all_cases[0]));
    if (argc < 2) { printf("%d\n", ncases); return 0; }
    const int which = atoi(argv[1]);
    if (which < 0 || which >= ncases) return 1;
    Case cases[1] = {all_cases[which]};
    for (auto &cs : cases) {
        CUtensorMap tm;
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cs.rank, d, cs.dims, cs.str, cs.box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         (CUtensorMapSwizzle)cs.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        uint32_t bytes = 4;
        for (int i = 0; i < cs.rank; ++i) bytes *= cs.box[i];
        int hst = -1;
        cudaMemset(st, 0xff, 4);
        if (r == CUDA_SUCCESS) {
            probe<<<1, 128, 66000>>>(tm, cs.rank, cs.c[0], cs.c[1], cs.c[2], cs.c[3], bytes, out, st);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&hst, st, 4, cudaMemcpyDeviceToHost);
            std::vector<float> o(bytes / 4);
            cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
            int untouched = 0, zeros = 0;
            for (float v : o) { untouched += (v == -777.0f); zeros += (v == 0.0f); }
            printf("%-52s encode ok, barrier %s, err=%s, untouched %d / %u, zeros %d, first %.0f %.0f %.0f %.0f\n", cs.name,
                   hst == 1 ? "completed" : "TIMEOUT", cudaGetErrorString(e), untouched, bytes / 4, zeros, o[0], o[1], o[8], o[32]);
        } else {
            printf("%-52s encode FAILED (%d)\n", cs.name, (int)r);
        }
    }
    return 0;
}
