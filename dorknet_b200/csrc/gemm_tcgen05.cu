// gemm_tcgen05.cu -- tcgen05 / TMEM / TMA GEMM kernels (placeholder until the first GPU bring-up).
#include "common.cuh"
#include "gemm.cuh"

namespace dk {

int init_gemm_tcgen05() { return DK_OK; }

int tc_conv_fwd(const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int,
                void *, size_t, cudaStream_t) { return DK_ERR_UNSUPPORTED; }
int tc_conv_dgrad(const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int,
                  void *, size_t, cudaStream_t) { return DK_ERR_UNSUPPORTED; }
int tc_conv_wgrad(const float *, const float *, const float *, float *, float, int, int, int, int, int, int, int, int,
                  int, void *, size_t, cudaStream_t) { return DK_ERR_UNSUPPORTED; }
int tc_dense_fwd(const float *, const float *, const float *, float *, int, int, int, void *, size_t, cudaStream_t) {
    return DK_ERR_UNSUPPORTED;
}
int tc_dense_bwd(const float *, const float *, const float *, float *, float *, float, int, int, int, void *, size_t,
                 cudaStream_t) { return DK_ERR_UNSUPPORTED; }
size_t tc_conv_ws_bytes(int, int, int, int, int, int, int, int, int) { return 0; }
size_t tc_dense_ws_bytes(int, int, int) { return 0; }

}  // namespace dk
