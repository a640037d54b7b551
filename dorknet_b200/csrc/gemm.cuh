// gemm.cuh -- internal interface between the C-ABI dispatch (conv_api.cu) and the two GEMM backends.
#pragma once
#include "common.cuh"

namespace dk {

// ---- SIMT implicit GEMM (gemm_simt.cu) ------------------------------------------------------------
int simt_conv_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                  int kh, int kw, int s, int p, cudaStream_t st);
int simt_conv_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int s, int p, int OH, int OW, cudaStream_t st);
int simt_conv_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W,
                    int F, int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
int simt_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim,
                   cudaStream_t st);
int simt_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float l2, int B, int in_dim,
                   int out_dim, void *ws, size_t ws_bytes, cudaStream_t st);
size_t simt_wgrad_ws_bytes(int64_t M, int Nn, int64_t K);
size_t simt_dense_ws_bytes(int B, int in_dim, int out_dim);
int simt_im2col(const float *x, float *P, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st);

// ---- tcgen05 / TMEM / TMA GEMM (gemm_tcgen05.cu) --------------------------------------------------
// Each returns DK_ERR_UNSUPPORTED (without setting an error) when the shape is outside what the
// tensor-core kernels cover; the dispatcher then uses the SIMT kernel for that call.
int tc_conv_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                  int s, int p, int OH, int OW, void *ws, size_t ws_bytes, cudaStream_t st);
int tc_conv_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W,
                  int F, int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
int tc_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim,
                 void *ws, size_t ws_bytes, cudaStream_t st);
int tc_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float l2, int B, int in_dim,
                 int out_dim, void *ws, size_t ws_bytes, cudaStream_t st);
size_t tc_conv_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p);
int tc_pw_dgrad_affine(const float *dy, const float *w, const float *x, const float *cb, const float *cd, float *dx, int N,
                       int C, int OH, int OW, int F, void *ws, size_t ws_bytes, cudaStream_t st);
size_t tc_pw_pack_bytes(int N, int C, int H, int W, int s);
int tc_pw_pack(const float *x, float *packed, int N, int C, int H, int W, int s, cudaStream_t st);
int tc_pw_fwd_packed(const float *xp, const float *w, const float *bias, float *y, int N, int C, int OH, int OW, int F, void *ws,
                     size_t ws_bytes, cudaStream_t st);
int tc_pw_dgrad_packed(const float *dyp, const float *w, float *dx, int N, int C, int OH, int OW, int F, int s, void *ws,
                       size_t ws_bytes, cudaStream_t st);
int tc_pw_wgrad_packed(const float *dy, int dy_packed, const float *x, int x_packed, const float *w, float *dw, float l2, int N,
                       int C, int H, int W, int F, int s, void *ws, size_t ws_bytes, cudaStream_t st);
int affine_add_launch(float *out, const float *x, const float *cb, const float *cd, int N, int C, int P, cudaStream_t st);  // bn_fold.cu
size_t tc_dense_ws_bytes(int B, int in_dim, int out_dim);
size_t tc_conv_mat_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p);

// ---- small-K (C*kh*kw <= 128) row-staged implicit-GEMM convolution: forward + wgrad (conv_rows.cu) -------------------
int conv_rows_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                  int kw, int s, int p, cudaStream_t st);
int conv_rows_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                    int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
size_t conv_rows_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p);
extern int g_conv_rows_enabled;

// ---- stride-1 k x k convolutions on 4-D tensor maps (conv_tma.cu): forward, dgrad, wgrad -------------------------------
int conv_tma_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                 int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
int conv_tma_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw, int s,
                   int p, void *ws, size_t ws_bytes, cudaStream_t st);
int conv_tma_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W, int F,
                   int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st);
size_t conv_tma_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int s, int p);
extern int g_conv_tma_enabled, g_ct_kc16, g_ct_wgrad2, g_cw2_rows;
void splitk_reduce_launch(const float *partial, const float *w, float *out, float l2, int64_t mn, int Z, cudaStream_t st);

}  // namespace dk
