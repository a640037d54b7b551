/*
 * dk_oracle.c -- CPU restatement of the reference's native (Cython/OpenMP) kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path; it is never
 * linked into, imported by or called from the product package (dorknet_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity status: PINNED.  Every function here is checked against the reference's own
 * compiled kernels (oracle/_ref, built from /root/reference/layers/*.pyx by
 * oracle/build_ref.py) in tests/test_oracle.py, and against the committed golden
 * vectors in tests/golden/ (generated from the live reference by
 * tests/golden/make_golden.py).
 *
 * Conventions: float32, contiguous NCHW, "Xp" means an already zero-padded input.
 * Output-size rule (reference layers/im2col.pyx:18-21): OHf = (Hp - kh)/s + 1 as a
 * real number, OH = floor(OHf).  Backward buffers have s*(OHf-1)+kh = Hp rows
 * (im2col.pyx:212-213), i.e. exactly the padded input, even when OHf is x.5.
 *
 * Build: gcc -O3 -fopenmp -ffast-math -shared -fPIC (same flags as reference setup.py:9).
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline int out_dim(int padded, int k, int s) { return (padded - k) / s + 1; }

/* reference: layers/im2col.pyx:16-36 (im2col_cy).
 * P[(n*OH + oh)*OW + ow, (c*kh + i)*kw + j] = Xp[n, c, oh*s + i, ow*s + j] */
void dk_oracle_im2col(const float *Xp, int N, int C, int Hp, int Wp, int kh, int kw, int s, float *P)
{
    const int OH = out_dim(Hp, kh, s), OW = out_dim(Wp, kw, s);
    const size_t K = (size_t)C * kh * kw;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n)
        for (int oh = 0; oh < OH; ++oh)
            for (int ow = 0; ow < OW; ++ow) {
                float *row = P + ((size_t)(n * OH + oh) * OW + ow) * K;
                for (int c = 0; c < C; ++c)
                    for (int i = 0; i < kh; ++i)
                        for (int j = 0; j < kw; ++j)
                            row[(c * kh + i) * kw + j] =
                                Xp[(((size_t)n * C + c) * Hp + oh * s + i) * Wp + ow * s + j];
            }
}

/* reference: layers/im2col.pyx:209-234 (row2im_cy): col2im scatter-add into a zero buffer
 * of the padded input size, then crop `pad` on each side.  dX is [N, C, Hp-2pad, Wp-2pad]. */
void dk_oracle_row2im(const float *rows, int N, int C, int Hp, int Wp, int kh, int kw, int s, int pad,
                      float *dX, float *scratch_padded /* [N,C,Hp,Wp] */)
{
    const int OH = out_dim(Hp, kh, s), OW = out_dim(Wp, kw, s);
    const size_t K = (size_t)C * kh * kw;
    memset(scratch_padded, 0, sizeof(float) * (size_t)N * C * Hp * Wp);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n)
        for (int oh = 0; oh < OH; ++oh)
            for (int ow = 0; ow < OW; ++ow) {
                const float *row = rows + ((size_t)(n * OH + oh) * OW + ow) * K;
                for (int c = 0; c < C; ++c)
                    for (int i = 0; i < kh; ++i)
                        for (int j = 0; j < kw; ++j)
                            scratch_padded[(((size_t)n * C + c) * Hp + oh * s + i) * Wp + ow * s + j] +=
                                row[(c * kh + i) * kw + j];
            }
    const int H = Hp - 2 * pad, W = Wp - 2 * pad;
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc)
        for (int h = 0; h < H; ++h)
            memcpy(dX + ((size_t)nc * H + h) * W,
                   scratch_padded + ((size_t)nc * Hp + h + pad) * Wp + pad, sizeof(float) * W);
}

/* reference: layers/im2col.pyx:109-139 (depthwise_conv_cy): per-channel cross-correlation,
 * accumulation order kh-major then kw. */
void dk_oracle_depthwise_fwd(const float *Xp, const float *Wt, int N, int C, int Hp, int Wp,
                             int kh, int kw, int s, float *Y)
{
    const int OH = out_dim(Hp, kh, s), OW = out_dim(Wp, kw, s);
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc) {
        const int c = nc % C;
        const float *x = Xp + (size_t)nc * Hp * Wp;
        const float *w = Wt + (size_t)c * kh * kw;
        float *y = Y + (size_t)nc * OH * OW;
        for (int oh = 0; oh < OH; ++oh)
            for (int ow = 0; ow < OW; ++ow) {
                float acc = 0.0f;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j)
                        acc += x[(size_t)(oh * s + i) * Wp + ow * s + j] * w[i * kw + j];
                y[(size_t)oh * OW + ow] = acc;
            }
    }
}

/* reference: layers/im2col.pyx:143-178 (depthwise_backward_direct_cy): fused dX scatter and
 * per-image dW partials dWn[N,C,kh,kw] (the caller sums over N,
 * layers/depthwise_convolution.py:193); dX cropped by `pad`. */
void dk_oracle_depthwise_bwd(const float *dY, const float *Xp, const float *Wt, int N, int C,
                             int Hp, int Wp, int kh, int kw, int s, int pad,
                             float *dX, float *dWn, float *scratch_padded /* [N,C,Hp,Wp] */)
{
    const int OH = out_dim(Hp, kh, s), OW = out_dim(Wp, kw, s);
    memset(scratch_padded, 0, sizeof(float) * (size_t)N * C * Hp * Wp);
    memset(dWn, 0, sizeof(float) * (size_t)N * C * kh * kw);
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc) {
        const int c = nc % C;
        const float *x = Xp + (size_t)nc * Hp * Wp;
        const float *w = Wt + (size_t)c * kh * kw;
        const float *dy = dY + (size_t)nc * OH * OW;
        float *dxp = scratch_padded + (size_t)nc * Hp * Wp;
        float *dw = dWn + (size_t)nc * kh * kw;
        for (int oh = 0; oh < OH; ++oh)
            for (int ow = 0; ow < OW; ++ow) {
                const float g = dy[(size_t)oh * OW + ow];
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) {
                        const size_t xi = (size_t)(oh * s + i) * Wp + ow * s + j;
                        dw[i * kw + j] += g * x[xi];
                        dxp[xi] += g * w[i * kw + j];
                    }
            }
    }
    const int H = Hp - 2 * pad, W = Wp - 2 * pad;
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc)
        for (int h = 0; h < H; ++h)
            memcpy(dX + ((size_t)nc * H + h) * W,
                   scratch_padded + ((size_t)nc * Hp + h + pad) * Wp + pad, sizeof(float) * W);
}

/* reference: layers/batch_norm_stats_cy.pyx:17-46 (channelwise_mean_and_var_4d): two-pass
 * biased variance with float32 accumulators. */
void dk_oracle_bn_stats(const float *A, int N, int C, int H, int W, float *mean, float *var)
{
    const size_t HW = (size_t)H * W;
    const float cnt = (float)((size_t)N * HW);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; ++c) {
        float m = 0.0f;
        for (int n = 0; n < N; ++n) {
            const float *a = A + ((size_t)n * C + c) * HW;
            for (size_t i = 0; i < HW; ++i) m += a[i];
        }
        m /= cnt;
        float v = 0.0f;
        for (int n = 0; n < N; ++n) {
            const float *a = A + ((size_t)n * C + c) * HW;
            for (size_t i = 0; i < HW; ++i) { const float d = a[i] - m; v += d * d; }
        }
        mean[c] = m;
        var[c] = v / cnt;
    }
}

/* reference: layers/relu_cy.pyx:11-107: out = x > 0 ? x : 0; train also writes a float 0/1 mask. */
void dk_oracle_relu_fwd(const float *X, size_t n, float *Y, float *mask)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        const int pos = X[i] > 0.0f;
        Y[i] = pos ? X[i] : 0.0f;
        if (mask) mask[i] = pos ? 1.0f : 0.0f;
    }
}

/* reference: layers/pooling_cy.pyx:10-69 (pool, pool_train): s x s window, stride s, strict '>'
 * so the first maximum in the row-major window scan wins; the int32 mask has a single 1 per
 * window in INPUT geometry.  H and W must be divisible by s (the reference reads out of bounds
 * otherwise). */
void dk_oracle_pool(const float *X, int N, int C, int H, int W, int s, float *Y, int32_t *mask)
{
    const int OH = H / s, OW = W / s;
    if (mask) memset(mask, 0, sizeof(int32_t) * (size_t)N * C * H * W);
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc) {
        const float *x = X + (size_t)nc * H * W;
        for (int p = 0; p < OH; ++p)
            for (int q = 0; q < OW; ++q) {
                const int k = p * s, l = q * s;
                float best = x[(size_t)k * W + l];
                int r = 0, t = 0;
                for (int m = 0; m < s; ++m)
                    for (int n2 = 0; n2 < s; ++n2) {
                        const float v = x[(size_t)(k + m) * W + l + n2];
                        if (v > best) { best = v; r = m; t = n2; }
                    }
                Y[((size_t)nc * OH + p) * OW + q] = best;
                if (mask) mask[((size_t)nc * H + k + r) * W + l + t] = 1;
            }
    }
}

/* reference: layers/pooling_cy.pyx:72-88 (pool_backward). */
void dk_oracle_pool_bwd(const int32_t *mask, const float *dY, int N, int C, int H, int W, int s, float *dX)
{
    const int OH = H / s, OW = W / s;
#pragma omp parallel for schedule(static)
    for (int nc = 0; nc < N * C; ++nc)
        for (int k = 0; k < H; ++k)
            for (int l = 0; l < W; ++l) {
                const size_t i = ((size_t)nc * H + k) * W + l;
                dX[i] = (mask[i] == 1) ? dY[((size_t)nc * OH + k / s) * OW + l / s] : 0.0f;
            }
}
