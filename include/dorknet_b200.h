/*
 * dorknet_b200.h -- C ABI of libdorknet_b200.so: hand-written sm_100a CUDA kernels for the
 * CNN training hot path of WJGiles/Dorknet (the path BASELINE.json:north_star names).
 *
 * This is the drop-in boundary.  One family of entry points per reference call site; the
 * reference file:line each one replaces is cited on the declaration (paths relative to the
 * reference tree).  The reference reaches its native code through Cython modules (CPU) or
 * cupy.RawKernel / cupy.dot (GPU); a maintainer binds these functions instead with the
 * ctypes stubs shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous float32 (int32 where stated), NCHW;
 *   - the library owns no tensors: outputs, caches and scratch ("ws") are caller-allocated.
 *     `ws` regions must be zero-filled once when allocated; kernels leave the counter words
 *     they use zeroed on exit, so a workspace can be reused launch after launch (and inside
 *     CUDA graphs) without memsets.  dk_*_ws_bytes() returns the size a call needs;
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*) and returns
 *     0 on success or a DK_ERR_* code; dk_last_error() gives the thread-local message;
 *   - conv output size rule (layers/im2col.pyx:18-21): OH = (H + 2*pad - kh) / stride + 1
 *     (floor).  Backward results always have the full input shape (im2col.pyx:212-213).
 */
#ifndef DORKNET_B200_H
#define DORKNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *dk_stream_t; /* cudaStream_t */

#define DK_OK 0
#define DK_ERR_INVALID 1   /* bad argument (raised as ValueError by the Python mirror) */
#define DK_ERR_CUDA 2      /* CUDA runtime / launch error */
#define DK_ERR_WORKSPACE 3 /* workspace too small */
#define DK_ERR_UNSUPPORTED 4

/* ---- library ---------------------------------------------------------------------------- */
int dk_version(void);
const char *dk_last_error(void);
/* Bind to `device`, query SM count, opt kernels in to large dynamic shared memory. */
int dk_init(int device);
int dk_destroy(void);
int dk_sm_count(void);
/* Number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches). */
unsigned long long dk_kernel_launches(void);
/* How many conv / pointwise / dense GEMM calls went to the tcgen05 and the SIMT backend since load. */
void dk_gemm_call_counts(unsigned long long *tc, unsigned long long *simt);
/* Diagnostics for the tensor-core path (bring-up and tests; not a reference call site): key 0 = disable mask
 * (bit0 fwd, bit1 dgrad, bit2 wgrad -> those calls use the SIMT kernels); keys 1-5 override the MN-major
 * shared-memory descriptor fields / TMA swizzle mode. */
int dk_tc_debug_set(int key, int value);
/* Depthwise: 1 (default) = register-window kernels for 3x3 / stride 1 / pad 1, 0 = shared-memory tile kernels only. */
int dk_dw_debug_set(int enable_rows);
/* GEMM backend for conv / pointwise / dense: 0 = tcgen05+TMEM+TMA (product path, default),
 * 1 = plain SIMT implicit GEMM (GPU-side cross-check used by tests only). */
int dk_set_gemm_backend(int backend);
int dk_get_gemm_backend(void);

/* ---- ReLu: layers/activations.py:14-47, layers/relu_cy.pyx:11-107 ------------------------- */
/* y = x > 0 ? x : 0; mask (nullable) = float 0/1 (the reference's positive_locs). */
int dk_relu_fwd(const float *x, float *y, float *mask, int64_t n, dk_stream_t stream);
/* dx = dy * (y > 0): `y` is the forward OUTPUT (or the 0/1 mask -- same test). */
int dk_relu_bwd(const float *dy, const float *y, float *dx, int64_t n, dk_stream_t stream);

/* ---- ResidualBlock join: layers/residual_block.py:75,93-95 ------------------------------- */
/* y = relu(a + b)  (post_skip_activation.forward(X_tmp + skippee)) */
int dk_add_relu_fwd(const float *a, const float *b, float *y, int64_t n, dk_stream_t stream);
/* out = a + b  (dx + joined_dx) */
int dk_add(const float *a, const float *b, float *out, int64_t n, dk_stream_t stream);

/* ---- BatchNormLayer: layers/batch_norm.py:54-174, layers/batch_norm_stats_cy.pyx:17-46 ---- */
/* X is [N, C, HW] (HW = H*W for 4-D inputs, 1 for 2-D inputs). */
size_t dk_bn_ws_bytes(int C);
/* mean[c], biased var[c] over (N, HW). */
int dk_bn_stats(const float *x, float *mean, float *var, int N, int C, int HW,
                void *ws, size_t ws_bytes, dk_stream_t stream);
/* Training forward.  Writes y = gamma*(x-mean)/std + beta (optionally followed by ReLU when
 * fuse_relu != 0), save_mean[c], save_invstd[c] = 1/sqrt(var+eps), save_scale[c] = gamma*invstd,
 * save_shift[c] = beta - mean*scale, and updates running_mean / running_STD in place:
 * r = momentum*r + (1-momentum)*batch, or r = batch when first_batch != 0
 * (batch_norm.py:76-89; note the reference tracks std = sqrt(var+eps), not var).
 * y may be NULL: statistics only (the apply is then fused into the consumer). */
int dk_bn_fwd_train(const float *x, float *y, const float *gamma, const float *beta,
                    float *running_mean, float *running_std, int first_batch, float momentum, float eps,
                    float *save_mean, float *save_invstd, float *save_scale, float *save_shift,
                    int fuse_relu, int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream);
/* The same with the residual join of a ResidualBlock folded into the normalisation pass:
 * y = relu?(batchnorm(x) + add)  (residual_block.py:75: post_skip_activation(X_tmp + skippee), X_tmp the output of the
 * branch's last BatchNormLayer).  Statistics and saved values are those of dk_bn_fwd_train; the backward of the join
 * stays with the caller (dk_relu_bwd on y, then dk_bn_bwd with fuse_relu = 0). */
int dk_bn_fwd_train_add(const float *x, const float *add, float *y, const float *gamma, const float *beta,
                        float *running_mean, float *running_std, int first_batch, float momentum, float eps,
                        float *save_mean, float *save_invstd, float *save_scale, float *save_shift,
                        int fuse_relu, int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream);
/* Test mode: y = gamma*(x - running_mean)/running_std + beta (batch_norm.py:112-115). */
int dk_bn_fwd_infer(const float *x, float *y, const float *gamma, const float *beta,
                    const float *running_mean, const float *running_std, int fuse_relu,
                    int N, int C, int HW, dk_stream_t stream);
/* y = x*scale[c] + shift[c] (+ReLU): materialises a deferred BN apply. */
int dk_bn_apply(const float *x, float *y, const float *scale, const float *shift, int fuse_relu,
                int N, int C, int HW, dk_stream_t stream);
/* The same restricted to the pixels a following stride-s pointwise convolution reads, written compactly:
 * y[n,c,oh,ow] = x[n,c,oh*s,ow*s]*scale[c] + shift[c] (+ReLU), y of shape [N, C, (H-1)/s+1, (W-1)/s+1]
 * (= BN output [:, :, ::s, ::s], pointwise_convolution.py:48). */
int dk_bn_apply_strided(const float *x, float *y, const float *scale, const float *shift, int fuse_relu,
                        int N, int C, int H, int W, int stride, dk_stream_t stream);
/* Backward (batch_norm.py:118-174): dgamma = sum(dy*x_hat), dbeta = sum(dy),
 * dx = gamma*invstd*(dy - mean(dy) - x_hat*mean(dy*x_hat)).  With fuse_relu != 0 the incoming
 * dy is first masked by (x*scale+shift > 0), i.e. the backward of a fused BN+ReLU. */
int dk_bn_bwd(const float *dy, const float *x, const float *gamma,
              const float *save_mean, const float *save_invstd, const float *save_scale, const float *save_shift,
              float *dx, float *dgamma, float *dbeta, int fuse_relu,
              int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream);
/* Backward of a ResidualBlock join folded into the backward of the branch's last BatchNorm
 * (residual_block.py:86-88: joined_dx = post_skip_activation.backward(upstream_dx); layer_list[-1].backward(joined_dx)):
 * d_join = dout * (out > 0) with `out` the block's output (activations.py:44-47), written once for the skip path, and
 * (dx, dgamma, dbeta) = BatchNorm backward of d_join -- one pass over (dout, out, x) instead of two kernels. */
int dk_bn_bwd_join(const float *dout, const float *out, const float *x, const float *gamma,
                   const float *save_mean, const float *save_invstd, const float *save_scale,
                   const float *save_shift, float *dx, float *d_join, float *dgamma, float *dbeta,
                   int N, int C, int HW, void *ws, size_t ws_bytes, dk_stream_t stream);
/* The same for the upstream gradient of a stride-2 pointwise convolution taken in its COMPACT form
 * dy_sub[N, C, H/2, W/2] (= the non-zero entries of the zero-stuffed gradient, pointwise_convolution.py:68-72:
 * dy[n,c,2i,2j] = dy_sub[n,c,i,j], 0 elsewhere).  Results are those of dk_bn_bwd on the zero-stuffed tensor.
 * Requires stride == 2, W % 8 == 0, H % 2 == 0. */
int dk_bn_bwd_strided(const float *dy_sub, const float *x, const float *gamma,
                      const float *save_mean, const float *save_invstd, const float *save_scale,
                      const float *save_shift, float *dx, float *dgamma, float *dbeta, int fuse_relu,
                      int N, int C, int H, int W, int stride, void *ws, size_t ws_bytes, dk_stream_t stream);

/* ---- DepthwiseConvLayer: layers/depthwise_convolution.py:72-83,186-196, im2col.pyx:109-178 - */
size_t dk_dwconv_ws_bytes(int N, int C, int H, int W, int kh, int kw, int stride, int pad);
/* y[n,c,oh,ow] = sum_{i,j} xpad[n,c,oh*s+i,ow*s+j] * w[c,i,j] (+ bias[c]).  When in_scale /
 * in_shift are non-NULL the input is read as relu?(x*in_scale[c] + in_shift[c]) (a deferred
 * BatchNorm(+ReLU) of the producer applied on load; zero padding is applied AFTER it). */
int dk_dwconv_fwd(const float *x, const float *w, const float *bias, float *y,
                  const float *in_scale, const float *in_shift, int in_relu,
                  int N, int C, int H, int W, int kh, int kw, int stride, int pad, dk_stream_t stream);
/* dx [N,C,H,W]; dw [C,kh,kw] (+ l2*w); dbias [C] nullable.  dx_add (nullable) is added into dx
 * (the residual join dx + joined_dx fused into the branch's first backward). */
int dk_dwconv_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias,
                  const float *in_scale, const float *in_shift, int in_relu, const float *dx_add,
                  float l2, int N, int C, int H, int W, int kh, int kw, int stride, int pad,
                  void *ws, size_t ws_bytes, dk_stream_t stream);

/* ---- ConvLayer: layers/convolution.py:58-126, layers/im2col.pyx:16-36,209-234 ------------- */
/* Implicit GEMM; no patch matrix is materialised.  w is [F, C, kh, kw]. */
size_t dk_conv2d_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad);
int dk_conv2d_fwd(const float *x, const float *w, const float *bias, float *y,
                  int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad,
                  void *ws, size_t ws_bytes, dk_stream_t stream);
/* dx [N, C, H, W] = crop(col2im(dY_rows @ Wflat)). */
int dk_conv2d_dgrad(const float *dy, const float *w, float *dx,
                    int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad,
                    void *ws, size_t ws_bytes, dk_stream_t stream);
/* dw [F,C,kh,kw] = dY_rows^T @ patches (+ l2*w); dbias [F] nullable = sum(dy,(0,2,3)). */
int dk_conv2d_wgrad(const float *dy, const float *x, const float *w, float *dw, float *dbias, float l2,
                    int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad,
                    void *ws, size_t ws_bytes, dk_stream_t stream);
/* Debug / parity: materialise the reference's patch matrix P[N*OH*OW, C*kh*kw] (bit-exact index
 * map of im2col_cy, layers/im2col.pyx:33-34). */
int dk_im2col_materialise(const float *x, float *patches, int N, int C, int H, int W,
                          int kh, int kw, int stride, int pad, dk_stream_t stream);

/* ---- PointwiseConvLayer: layers/pointwise_convolution.py:46-75 ---------------------------- */
/* x [N,C,H,W] subsampled [::stride, ::stride]; w [F,C]; y [N,F,OH,OW], OH = ceil(H/stride). */
size_t dk_pwconv_ws_bytes(int N, int C, int H, int W, int F, int stride);
int dk_pwconv_fwd(const float *x, const float *w, const float *bias, float *y,
                  int N, int C, int H, int W, int F, int stride,
                  void *ws, size_t ws_bytes, dk_stream_t stream);
/* dx is [N, C, OH*stride, OW*stride], zero-stuffed (pointwise_convolution.py:68-72). */
int dk_pwconv_dgrad(const float *dy, const float *w, float *dx,
                    int N, int C, int OH, int OW, int F, int stride,
                    void *ws, size_t ws_bytes, dk_stream_t stream);
int dk_pwconv_wgrad(const float *dy, const float *x, const float *w, float *dw, float *dbias, float l2,
                    int N, int C, int H, int W, int F, int stride,
                    void *ws, size_t ws_bytes, dk_stream_t stream);

/* Operands re-pitched once.  Planes whose pitch is not a multiple of 16 bytes (7x7) and small strided planes cannot go
 * through TMA as they lie; dk_pwconv_fwd / dgrad / wgrad copy them into a padded layout in their workspace on every call --
 * the same x twice per training step (forward, wgrad), the same dY twice (dgrad, wgrad).  A caller that keeps the copies
 * avoids half of that: dk_pw_pack_bytes (0 = this shape needs no packing; otherwise the buffer size) and dk_pw_pack produce
 * [N][C][round_up(OH*OW, 4)] with the stride already applied and zero padding; the *_packed entry points take such buffers
 * (wgrad: flags say which of dy / x is a packed copy; dbias needs the unpacked dY).  Tensor-core backend only: they fail
 * (no fallback -- the unpacked tensor is not at hand) when C % 4 != 0 or the weights are not 16-byte aligned. */
size_t dk_pw_pack_bytes(int N, int C, int H, int W, int stride);
int dk_pw_pack(const float *x, float *packed, int N, int C, int H, int W, int stride, dk_stream_t stream);
int dk_pwconv_fwd_packed(const float *x_packed, const float *w, const float *bias, float *y, int N, int C, int OH, int OW, int F,
                         void *ws, size_t ws_bytes, dk_stream_t stream);
int dk_pwconv_dgrad_packed(const float *dy_packed, const float *w, float *dx, int N, int C, int OH, int OW, int F, int stride,
                           void *ws, size_t ws_bytes, dk_stream_t stream);
int dk_pwconv_wgrad_packed(const float *dy, int dy_is_packed, const float *x, int x_is_packed, const float *w, float *dw,
                           float *dbias, float l2, int N, int C, int H, int W, int F, int stride, void *ws, size_t ws_bytes,
                           dk_stream_t stream);

/* ---- BatchNorm (no ReLU) -> PointwiseConvLayer, folded (training): layers/batch_norm.py:54-174 applied to the input of
 * layers/pointwise_convolution.py:46-75.  Between the two layers there is no non-linearity, so with the batch statistics
 * known (dk_bn_fwd_train with y = NULL) the normalisation is absorbed by the GEMM and its backward by the GEMM's epilogue;
 * the normalised activation is never written or read (csrc/bn_fold.cu has the algebra).
 *   dk_bn_fold_fwd: w_out[f,c] = w[f,c]*scale[c]; b_out[f] = (bias[f]) + sum_c w[f,c]*shift[c]; then
 *                   dk_pwconv_fwd(x_raw, w_out, b_out) == pointwise(batchnorm(x_raw)).  With mean (and trunc_resid, the
 *                   per-channel mean of x - tf32_truncate(x) left by dk_dwconv_fwd_bn) non-NULL, w_out is stored
 *                   TF32-truncated and b_out compensates what the tensor core's operand truncation does to the output
 *                   mean of a GEMM over un-centred activations (csrc/bn_fold.cu).
 *   dk_dwconv_fwd_bn: the depthwise forward of dk_dwconv_fwd AND the training statistics of the BatchNorm that follows it
 *                   (dk_bn_fwd_train with y = NULL: save_*, running_*), accumulated while the outputs are still in
 *                   registers -- no pass over the activation.  dk_dwconv_fwd_bn_ws_bytes returns 0 for shapes it does not
 *                   cover (then call the two functions separately).
 *   dk_bn_fold_bwd: from g_raw = dk_pwconv_wgrad(dy, x_raw) (l2 = 0) and s_col[f] = sum_{n,p} dy (dk_bias_grad; NULL when the
 *                   caller knows it is zero, e.g. dy is a BatchNorm's input gradient): dw (+ l2*w), the BatchNorm's dgamma /
 *                   dbeta [C], and the coefficients cb, cd [C] of its input gradient.  count = N*H*W of the BatchNorm.
 *   dk_pwconv_dgrad_affine: dx[n,c,p] = sum_f w[f,c]*dy[n,f,p] + cb[c]*x[n,c,p] + cd[c] (stride 1): called with the FOLDED
 *                   weights and the BatchNorm's input x it yields the BatchNorm's input gradient directly. */
int dk_bn_fold_fwd(const float *w, const float *bias, const float *scale, const float *shift, const float *mean,
                   const float *trunc_resid, float *w_out, float *b_out, int F, int C, dk_stream_t stream);
size_t dk_dwconv_fwd_bn_ws_bytes(int N, int C, int H, int W, int kh, int kw, int stride, int pad);
int dk_dwconv_fwd_bn(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                     int stride, int pad, const float *gamma, const float *beta, float *running_mean, float *running_std,
                     int first_batch, float momentum, float eps, float *save_mean, float *save_invstd, float *save_scale,
                     float *save_shift, float *trunc_resid, void *ws, size_t ws_bytes, dk_stream_t stream);
int dk_bn_fold_bwd(const float *g_raw, const float *s_col, const float *w, const float *gamma, const float *save_mean,
                   const float *save_invstd, const float *save_scale, const float *save_shift, float l2, int64_t count,
                   float *dw, float *dgamma, float *dbeta, float *cb, float *cd, int F, int C, dk_stream_t stream);
int dk_pwconv_dgrad_affine(const float *dy, const float *w, const float *x, const float *cb, const float *cd, float *dx,
                           int N, int C, int OH, int OW, int F, void *ws, size_t ws_bytes, dk_stream_t stream);

/* ---- DenseLayer: layers/dense_layer.py:46-67 ------------------------------------------------ */
/* y[B,out] = x[B,in] @ w[in,out] (+ bias). */
size_t dk_dense_ws_bytes(int B, int in_dim, int out_dim);
int dk_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim,
                 void *ws, size_t ws_bytes, dk_stream_t stream);
/* dx = dy @ w^T; dw = x^T @ dy (+ l2*w); dbias (nullable) = sum(dy, 0). */
int dk_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, float l2,
                 int B, int in_dim, int out_dim, void *ws, size_t ws_bytes, dk_stream_t stream);

/* per-channel sum over (N, HW) of dy[N,F,HW]: grads["bias"] (convolution.py:91-92). */
int dk_bias_grad(const float *dy, float *dbias, int N, int F, int HW,
                 void *ws, size_t ws_bytes, dk_stream_t stream);

/* ---- pooling: layers/pooling.py:23-77, layers/pooling_cy.pyx:10-88 ------------------------ */
int dk_gap_fwd(const float *x, float *y, int N, int C, int HW, dk_stream_t stream);
int dk_gap_bwd(const float *dy, float *dx, int N, int C, int HW, dk_stream_t stream);
/* s x s window, stride s, H and W divisible by s; strict '>' so the first maximum in the
 * row-major window scan wins; mask is int32 one-hot in INPUT geometry (bit-exact). */
int dk_maxpool_fwd(const float *x, float *y, int N, int C, int H, int W, int s, dk_stream_t stream);
int dk_maxpool_fwd_train(const float *x, float *y, int32_t *mask, int N, int C, int H, int W, int s,
                         dk_stream_t stream);
int dk_maxpool_bwd(const int32_t *mask, const float *dy, float *dx, int N, int C, int H, int W, int s,
                   dk_stream_t stream);

/* ---- loss + regulariser: layers/losses.py:13-34, regularisers/l2.py:12-17 ----------------- */
/* probs = exp(x)/sum(exp(x)) (NO max subtraction, as the reference); when y_one_hot != NULL also
 * loss[0] = (1/B) * sum_b -log(sum_j probs[b,j]*y[b,j]). */
int dk_softmax_xent_fwd(const float *logits, const float *y_one_hot, float *probs, float *loss,
                        int B, int K, dk_stream_t stream);
/* dx = (probs - y)/B */
int dk_softmax_xent_bwd(const float *probs, const float *y_one_hot, float *dx, int B, int K, dk_stream_t stream);
/* out[0] = scale * sum(w^2)   (l2.forward with scale = 0.5*strength) */
int dk_sumsq(const float *w, float *out, float scale, int64_t n, dk_stream_t stream);
/* The same for many tensors in ONE launch (`tasks` is a DEVICE array): every layer's regulariser_forward()
 * (layers/layer.py:42-46) of a step, flushed by the loss layer. */
typedef struct {
    const float *w;
    float *out;
    int64_t n;
    float scale;
    int32_t pad_;
} dk_sumsq_task;
int dk_sumsq_multi(const dk_sumsq_task *tasks, int num_tasks, dk_stream_t stream);

/* ---- optimisers: optimisers/SGD.py:20-24, SGDMomentum.py:31-39, RMSProp.py:28-36 ----------- */
/* One fused multi-tensor launch.  `table` is a DEVICE array of num_tensors descriptors;
 * grad_scale multiplies every gradient first (1/world_size after a data-parallel all-reduce).
 * `hyper` (nullable) is a DEVICE array {lr, momentum-or-decay, grad_scale}: when given it overrides the scalar
 * arguments, so a training step captured in a CUDA graph follows later set_learning_rate() calls. */
typedef struct {
    float *param;
    const float *grad;
    float *state; /* momentum buffer (SGDMomentum) / squared-grad cache (RMSProp); unused for SGD */
    int64_t n;
} dk_opt_tensor;
int dk_opt_sgd_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n,
                     float lr, float grad_scale, const float *hyper, dk_stream_t stream);
/* v = -lr*g + momentum*v ; w += v */
int dk_opt_sgdm_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n,
                      float lr, float momentum, float grad_scale, const float *hyper, dk_stream_t stream);
/* c = decay*c + (1-decay)*g^2 ; w -= lr*g/sqrt(c + 1e-5) */
int dk_opt_rmsprop_multi(const dk_opt_tensor *table, int num_tensors, int64_t max_n,
                         float lr, float decay, float grad_scale, const float *hyper, dk_stream_t stream);

/* ---- data parallel over NVLink peer memory: gradient exchange fused into the optimiser kernel ----
 * Each rank allocates its flat gradient buffer and a flag block with dk_p2p_alloc, exchanges the 64-byte handles
 * (any host channel: torch.distributed here), maps the peers' buffers with dk_p2p_open and fills a dk_p2p_ctx in
 * DEVICE memory.  Per step: dk_p2p_wait_done before the first kernel that writes gradients, dk_opt_multi_p2p in place of
 * dk_opt_*_multi: "gradients ready" handshake, ONE kernel that sums element i of every rank's gradient buffer in rank
 * order (peers read over NVLink), scales by hyper[2] and applies optimiser `kind` (0 SGD, 1 SGDMomentum, 2 RMSProp;
 * hyper = {lr, momentum|decay, grad_scale} in device memory), then the "done reading" handshake. */
#define DK_P2P_MAX_RANKS 8
typedef struct {
    int world, rank;
    long long grad_delta[DK_P2P_MAX_RANKS]; /* byte offset from MY gradient buffer to rank p's mapping of its own */
    unsigned int *ready[DK_P2P_MAX_RANKS];  /* rank p's flag block (uint32[8], indexed by writer), as mapped HERE */
    unsigned int *done[DK_P2P_MAX_RANKS];
    unsigned int *epoch;                    /* local step counter (device memory) */
    unsigned int *reduced[DK_P2P_MAX_RANKS]; /* "slice p of rank p's buffer holds the sum" flags, same indexing */
    float *grad_base;                       /* MY flat gradient buffer */
    long long nfloats;                      /* its length (a multiple of 4) */
    long long slice;                        /* floats per rank slice (a multiple of 4; world*slice >= nfloats) */
} dk_p2p_ctx;
int dk_p2p_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int dk_p2p_open(const unsigned char *handle64, void **ptr);
int dk_p2p_close(void *ptr);
int dk_p2p_free(void *ptr);
int dk_p2p_wait_done(const dk_p2p_ctx *ctx, dk_stream_t stream);
int dk_opt_multi_p2p(int kind, const dk_opt_tensor *table, int num_tensors, int64_t max_n, const float *hyper,
                     const dk_p2p_ctx *ctx, dk_stream_t stream);

/* ---- input pipeline (next row): data_loading/image_data_loader.py:100-112 ---- */
/* out = lam*xb + (1-lam)*xa  (mixup of two batches / label sets) */
int dk_mixup(const float *xa, const float *xb, float *out, float lam, int64_t n, dk_stream_t stream);
/* uint8 NHWC batch(es) as decoded -> fp32 NCHW network input, one kernel (image_preprocessor.py:36-37 + mixup):
 * out[n,c,h,w] = (1-lam)*(xa[n,h,w,c] - sub) + lam*(xb[n,h,w,c] - sub); xb may be NULL (no mixup, lam ignored). */
int dk_input_u8_nhwc(const unsigned char *xa, const unsigned char *xb, float *out, float lam, float sub,
                     int N, int C, int H, int W, dk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DORKNET_B200_H */
