// conv_api.cu -- C-ABI entry points for ConvLayer / PointwiseConvLayer / DenseLayer.
// Dispatch: tcgen05+TMEM+TMA kernels (backend 0, default) with the SIMT implicit GEMM for shapes
// they do not cover; backend 1 forces SIMT (tests use it as the on-GPU cross-check).
#include "common.cuh"
#include "gemm.cuh"

namespace dk {

static unsigned long long g_tc_calls = 0, g_simt_calls = 0;

static int conv_check(const char *who, int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    DK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && F > 0 && kh > 0 && kw > 0 && s > 0 && p >= 0,
               "%s: bad shape N=%d C=%d H=%d W=%d F=%d k=%dx%d s=%d p=%d", who, N, C, H, W, F, kh, kw, s, p);
    DK_REQUIRE(H + 2 * p >= kh && W + 2 * p >= kw, "%s: filter larger than the padded input", who);
    DK_REQUIRE((int64_t)N * C * H * W < ((int64_t)1 << 31) * 8, "%s: tensor too large", who);
    return DK_OK;
}

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

static size_t conv_ws(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    const int OH = (H + 2 * p - kh) / s + 1, OW = (W + 2 * p - kw) / s + 1;
    const size_t simt = simt_wgrad_ws_bytes(F, C * kh * kw, (int64_t)N * OH * OW);
    const size_t tc = max_sz(max_sz(tc_conv_ws_bytes(N, C, H, W, F, kh, kw, s, p), conv_rows_ws_bytes(N, C, H, W, F, kh, kw, s, p)),
                             conv_tma_ws_bytes(N, C, H, W, F, kh, kw, s, p));
    return max_sz(max_sz(simt, tc), tc_conv_mat_ws_bytes(N, C, H, W, F, kh, kw, s, p)) + 256;
}

}  // namespace dk

using namespace dk;

#define DK_TRY_TC(call)                          \
    if (gemm_backend() == 0) {                   \
        const int _rc = (call);                  \
        if (_rc != DK_ERR_UNSUPPORTED) {         \
            if (_rc == DK_OK) ++g_tc_calls;      \
            return _rc;                          \
        }                                        \
    }

extern "C" {

/* test hook: how many conv/pointwise/dense GEMM calls went to each backend since load */
void dk_gemm_call_counts(unsigned long long *tc, unsigned long long *simt) {
    if (tc) *tc = g_tc_calls;
    if (simt) *simt = g_simt_calls;
}

size_t dk_conv2d_ws_bytes(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0) return 0;
    return conv_ws(N, C, H, W, F, kh, kw, stride, pad);
}

int dk_conv2d_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                  int kh, int kw, int stride, int pad, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = conv_check("dk_conv2d_fwd", N, C, H, W, F, kh, kw, stride, pad);
    if (rc) return rc;
    DK_REQUIRE(x && w && y, "dk_conv2d_fwd: NULL pointer");
    DK_TRY_TC(conv_rows_fwd(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad, as_stream(stream)));
    DK_TRY_TC(conv_tma_fwd(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    DK_TRY_TC(tc_conv_fwd(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_conv_fwd(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad, as_stream(stream));
}

int dk_conv2d_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = conv_check("dk_conv2d_dgrad", N, C, H, W, F, kh, kw, stride, pad);
    if (rc) return rc;
    DK_REQUIRE(dy && w && dx, "dk_conv2d_dgrad: NULL pointer");
    const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
    DK_TRY_TC(conv_tma_dgrad(dy, w, dx, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    DK_TRY_TC(tc_conv_dgrad(dy, w, dx, N, C, H, W, F, kh, kw, stride, pad, OH, OW, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_conv_dgrad(dy, w, dx, N, C, H, W, F, kh, kw, stride, pad, OH, OW, as_stream(stream));
}

int dk_conv2d_wgrad(const float *dy, const float *x, const float *w, float *dw, float *dbias, float l2, int N, int C,
                    int H, int W, int F, int kh, int kw, int stride, int pad, void *ws, size_t ws_bytes,
                    dk_stream_t stream) {
    int rc = conv_check("dk_conv2d_wgrad", N, C, H, W, F, kh, kw, stride, pad);
    if (rc) return rc;
    DK_REQUIRE(dy && x && w && dw, "dk_conv2d_wgrad: NULL pointer");
    const int OH = (H + 2 * pad - kh) / stride + 1, OW = (W + 2 * pad - kw) / stride + 1;
    if (dbias) {
        rc = dk_bias_grad(dy, dbias, N, F, OH * OW, nullptr, 0, stream);
        if (rc) return rc;
    }
    DK_TRY_TC(conv_rows_wgrad(dy, x, w, dw, l2, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    DK_TRY_TC(conv_tma_wgrad(dy, x, w, dw, l2, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    DK_TRY_TC(tc_conv_wgrad(dy, x, w, dw, l2, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_conv_wgrad(dy, x, w, dw, l2, N, C, H, W, F, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream));
}

int dk_im2col_materialise(const float *x, float *patches, int N, int C, int H, int W, int kh, int kw, int stride,
                          int pad, dk_stream_t stream) {
    int rc = conv_check("dk_im2col_materialise", N, C, H, W, 1, kh, kw, stride, pad);
    if (rc) return rc;
    DK_REQUIRE(x && patches, "dk_im2col_materialise: NULL pointer");
    return simt_im2col(x, patches, N, C, H, W, kh, kw, stride, pad, as_stream(stream));
}

/* Pointwise = 1x1 convolution with pad 0 and stride s: OH = (H-1)/s + 1 = ceil(H/s), which is exactly
 * X[:, :, ::s, ::s] (pointwise_convolution.py:48-49). */
size_t dk_pwconv_ws_bytes(int N, int C, int H, int W, int F, int stride) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || F <= 0 || stride <= 0) return 0;
    return conv_ws(N, C, H, W, F, 1, 1, stride, 0);
}

int dk_pwconv_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                  int stride, void *ws, size_t ws_bytes, dk_stream_t stream) {
    return dk_conv2d_fwd(x, w, bias, y, N, C, H, W, F, 1, 1, stride, 0, ws, ws_bytes, stream);
}

int dk_pwconv_dgrad(const float *dy, const float *w, float *dx, int N, int C, int OH, int OW, int F, int stride,
                    void *ws, size_t ws_bytes, dk_stream_t stream) {
    /* dx is [N, C, OH*s, OW*s] zero-stuffed, NOT the forward input's H x W (pointwise_convolution.py:68-72) */
    const int H = OH * stride, W = OW * stride;
    int rc = conv_check("dk_pwconv_dgrad", N, C, H, W, F, 1, 1, stride, 0);
    if (rc) return rc;
    DK_REQUIRE(dy && w && dx, "dk_pwconv_dgrad: NULL pointer");
    DK_TRY_TC(tc_conv_dgrad(dy, w, dx, N, C, H, W, F, 1, 1, stride, 0, OH, OW, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_conv_dgrad(dy, w, dx, N, C, H, W, F, 1, 1, stride, 0, OH, OW, as_stream(stream));
}

int dk_pwconv_dgrad_affine(const float *dy, const float *w, const float *x, const float *cb, const float *cd, float *dx,
                           int N, int C, int OH, int OW, int F, void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = conv_check("dk_pwconv_dgrad_affine", N, C, OH, OW, F, 1, 1, 1, 0);
    if (rc) return rc;
    DK_REQUIRE(dy && w && x && cb && cd && dx, "dk_pwconv_dgrad_affine: NULL pointer");
    DK_TRY_TC(tc_pw_dgrad_affine(dy, w, x, cb, cd, dx, N, C, OH, OW, F, ws, ws_bytes, as_stream(stream)));
    /* fallback: the plain dgrad, then the affine term as its own pass */
    rc = dk_pwconv_dgrad(dy, w, dx, N, C, OH, OW, F, 1, ws, ws_bytes, stream);
    if (rc) return rc;
    return affine_add_launch(dx, x, cb, cd, N, C, OH * OW, as_stream(stream));
}

int dk_pwconv_wgrad(const float *dy, const float *x, const float *w, float *dw, float *dbias, float l2, int N, int C,
                    int H, int W, int F, int stride, void *ws, size_t ws_bytes, dk_stream_t stream) {
    return dk_conv2d_wgrad(dy, x, w, dw, dbias, l2, N, C, H, W, F, 1, 1, stride, 0, ws, ws_bytes, stream);
}

/* ---- pointwise operands re-pitched once (gemm_tcgen05.cu: tc_pw_pack*) -------------------------------------------------- */
size_t dk_pw_pack_bytes(int N, int C, int H, int W, int stride) {
    return gemm_backend() == 0 ? tc_pw_pack_bytes(N, C, H, W, stride) : 0;
}

int dk_pw_pack(const float *x, float *packed, int N, int C, int H, int W, int stride, dk_stream_t stream) {
    DK_REQUIRE(x && packed && N > 0 && C > 0 && H > 0 && W > 0 && stride > 0, "dk_pw_pack: bad arguments");
    const int rc = tc_pw_pack(x, packed, N, C, H, W, stride, as_stream(stream));
    DK_REQUIRE(rc != DK_ERR_UNSUPPORTED, "dk_pw_pack: this shape needs no packing (dk_pw_pack_bytes returns 0) or `packed` is misaligned");
    return rc;
}

int dk_pwconv_fwd_packed(const float *x_packed, const float *w, const float *bias, float *y, int N, int C, int OH, int OW, int F,
                         void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = conv_check("dk_pwconv_fwd_packed", N, C, OH, OW, F, 1, 1, 1, 0);
    if (rc) return rc;
    DK_REQUIRE(x_packed && w && y, "dk_pwconv_fwd_packed: NULL pointer");
    rc = tc_pw_fwd_packed(x_packed, w, bias, y, N, C, OH, OW, F, ws, ws_bytes, as_stream(stream));
    DK_REQUIRE(rc != DK_ERR_UNSUPPORTED, "dk_pwconv_fwd_packed: tensor-core path unavailable for this call");
    if (rc == DK_OK) ++g_tc_calls;
    return rc;
}

int dk_pwconv_dgrad_packed(const float *dy_packed, const float *w, float *dx, int N, int C, int OH, int OW, int F, int stride,
                           void *ws, size_t ws_bytes, dk_stream_t stream) {
    int rc = conv_check("dk_pwconv_dgrad_packed", N, C, OH * stride, OW * stride, F, 1, 1, stride, 0);
    if (rc) return rc;
    DK_REQUIRE(dy_packed && w && dx, "dk_pwconv_dgrad_packed: NULL pointer");
    rc = tc_pw_dgrad_packed(dy_packed, w, dx, N, C, OH, OW, F, stride, ws, ws_bytes, as_stream(stream));
    DK_REQUIRE(rc != DK_ERR_UNSUPPORTED, "dk_pwconv_dgrad_packed: tensor-core path unavailable for this call");
    if (rc == DK_OK) ++g_tc_calls;
    return rc;
}

int dk_pwconv_wgrad_packed(const float *dy, int dy_is_packed, const float *x, int x_is_packed, const float *w, float *dw,
                           float *dbias, float l2, int N, int C, int H, int W, int F, int stride, void *ws, size_t ws_bytes,
                           dk_stream_t stream) {
    int rc = conv_check("dk_pwconv_wgrad_packed", N, C, H, W, F, 1, 1, stride, 0);
    if (rc) return rc;
    DK_REQUIRE(dy && x && w && dw, "dk_pwconv_wgrad_packed: NULL pointer");
    DK_REQUIRE(dbias == nullptr || !dy_is_packed, "dk_pwconv_wgrad_packed: the bias gradient needs the unpacked dY");
    const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
    if (dbias) {
        rc = dk_bias_grad(dy, dbias, N, F, OH * OW, nullptr, 0, stream);
        if (rc) return rc;
    }
    rc = tc_pw_wgrad_packed(dy, dy_is_packed, x, x_is_packed, w, dw, l2, N, C, H, W, F, stride, ws, ws_bytes, as_stream(stream));
    DK_REQUIRE(rc != DK_ERR_UNSUPPORTED, "dk_pwconv_wgrad_packed: tensor-core path unavailable for this call");
    if (rc == DK_OK) ++g_tc_calls;
    return rc;
}

size_t dk_dense_ws_bytes(int B, int in_dim, int out_dim) {
    if (B <= 0 || in_dim <= 0 || out_dim <= 0) return 0;
    return max_sz(simt_dense_ws_bytes(B, in_dim, out_dim), tc_dense_ws_bytes(B, in_dim, out_dim)) + 256;
}

int dk_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim,
                 void *ws, size_t ws_bytes, dk_stream_t stream) {
    DK_REQUIRE(B > 0 && in_dim > 0 && out_dim > 0 && x && w && y, "dk_dense_fwd: bad arguments");
    DK_TRY_TC(tc_dense_fwd(x, w, bias, y, B, in_dim, out_dim, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_dense_fwd(x, w, bias, y, B, in_dim, out_dim, as_stream(stream));
}

int dk_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, float l2, int B,
                 int in_dim, int out_dim, void *ws, size_t ws_bytes, dk_stream_t stream) {
    DK_REQUIRE(B > 0 && in_dim > 0 && out_dim > 0 && dy && x && w && dx && dw, "dk_dense_bwd: bad arguments");
    if (dbias) {
        /* sum over the batch axis: dy viewed as [N=B, F=out, HW=1] */
        int rc = dk_bias_grad(dy, dbias, B, out_dim, 1, nullptr, 0, stream);
        if (rc) return rc;
    }
    DK_TRY_TC(tc_dense_bwd(dy, x, w, dx, dw, l2, B, in_dim, out_dim, ws, ws_bytes, as_stream(stream)));
    ++g_simt_calls;
    return simt_dense_bwd(dy, x, w, dx, dw, l2, B, in_dim, out_dim, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
