// pdl_gap.cu -- per-node cost of a CUDA graph of dependent kernels with and without programmatic dependent launch
// (griddepcontrol.launch_dependents at the top of every kernel, griddepcontrol.wait before its first memory access).
// A measurement tool for DESIGN.md, not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/pdl_gap.bin tools/micro/pdl_gap.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) step_kernel(float *x, long long n, int pdl) {
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] += 1.0f;
}

static float run(float *x, long long n, int nodes, int pdl, int grid) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < nodes; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, step_kernel, x, n, pdl);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 20; ++i) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) printf("error\n");
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    cudaStreamDestroy(st);
    return ms * 1000.0f / 20;
}

int main() {
    float *x;
    const long long big = 64ll << 20;  // 256 MB
    cudaMalloc(&x, big * 4);
    cudaMemset(x, 0, big * 4);
    const long long sizes[] = {1024, 1ll << 20, 8ll << 20, 32ll << 20};
    for (long long n : sizes)
        for (int grid : {148, 1184}) {
            const float a = run(x, n, 217, 0, grid), b = run(x, n, 217, 1, grid);
            printf("n = %9lld floats, grid %4d: 217 nodes  plain %.1f us (%.2f / node)   PDL %.1f us (%.2f / node)   saved %.2f us / node\n", n,
                   grid, a, a / 217, b, b / 217, (a - b) / 217);
        }
    // correctness of the chain under PDL: every element was incremented the same number of times
    cudaMemset(x, 0, big * 4);
    run(x, 1 << 20, 50, 1, 148);
    float h[4];
    cudaMemcpy(h, x, 16, cudaMemcpyDeviceToHost);
    printf("check: x[0] = %.0f (expect %d)\n", h[0], 50 * 23);
    return 0;
}
