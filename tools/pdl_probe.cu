// pdl_probe.cu -- what does a kernel boundary cost inside a CUDA graph on B200, and how much of it does programmatic
// dependent launch (griddepcontrol) remove?  Chain of NK dependent kernels (each reads what the previous one wrote),
// captured into a graph with and without cudaLaunchAttributeProgrammaticStreamSerialization.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_probe tools/pdl_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <bool PDL, bool EARLY>
__global__ void __launch_bounds__(256) step_kernel(const float *__restrict__ in, float *__restrict__ out, int n) {
    if (PDL && EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) out[i] = in[i] * 1.0001f + 1.0f;
    if (PDL && !EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <bool PDL, bool EARLY>
static float run(float *a, float *b, int n, int grid, int nk) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int k = 0; k < nk; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(256);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = PDL ? 1 : 0;
        const float *in = (k & 1) ? b : a;
        float *out = (k & 1) ? a : b;
        cudaLaunchKernelEx(&cfg, step_kernel<PDL, EARLY>, in, out, n);
    }
    cudaStreamEndCapture(st, &g);
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return -1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    e = cudaGetLastError();
    if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
    return ms * 1e3f / (10 * nk);
}

int main() {
    const int nmax = 64 << 20;
    float *a, *b;
    cudaMalloc(&a, nmax * 4);
    cudaMalloc(&b, nmax * 4);
    cudaMemset(a, 0, nmax * 4);
    cudaMemset(b, 0, nmax * 4);
    const int sizes[] = {1 << 10, 1 << 18, 1 << 20, 1 << 22, 1 << 24};  // floats per kernel: 4 KB .. 64 MB
    for (int n : sizes) {
        const int grid = n / 256 < 148 * 8 ? (n / 256 > 0 ? n / 256 : 1) : 148 * 8;
        const float t0 = run<false, false>(a, b, n, grid, 200);
        const float t1 = run<true, false>(a, b, n, grid, 200);
        const float t2 = run<true, true>(a, b, n, grid, 200);
        printf("n = %8d floats, grid %4d: plain %.2f us/kernel | PDL (trigger at end) %.2f | PDL (trigger at start) %.2f\n", n, grid, t0, t1, t2);
    }
    return 0;
}
