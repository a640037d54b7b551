// bn_fold.cu -- a training-mode BatchNorm (no ReLU) folded into the pointwise convolution that follows it.
//
// The depthwise-separable unit of the reference is dw3x3 -> BatchNorm -> pointwise -> BatchNorm -> ReLU
// (examples/imagenet_dogs_225_resnet_18_depsep.py:34-70): between the first BatchNorm (layers/batch_norm.py:54-100) and the
// pointwise GEMM (layers/pointwise_convolution.py:46-55) there is no non-linearity, so once the batch statistics are known the
// normalisation is a per-channel affine map xh = scale[c]*x + shift[c] that the GEMM absorbs:
//
//   forward   y = W . xh = (W diag(scale)) . x + W . shift          -> dk_bn_fold_fwd builds W' = W diag(scale), b' = b + W.shift;
//                                                                     the GEMM reads the RAW depthwise output, xh never exists
//   wgrad     dW[f,c] = sum_p dY[f,p] xh[c,p] = scale[c] G[f,c] + shift[c] S[f],   G = dY . x^T (the raw wgrad GEMM), S[f] = sum_p dY[f,p]
//   BatchNorm backward (batch_norm.py:118-174) needs, per channel, sum_p dXh and sum_p dXh*x with dXh = W^T dY:
//             sum_p dXh[c,p]        = sum_f W[f,c] S[f]
//             sum_p dXh[c,p] x[c,p] = sum_f W[f,c] G[f,c]                           -> no pass over the activations at all
//   and then  dx = scale*(dXh - mean(dXh)) - scale*invstd*(x - mu)*mean(dXh*xn) = (W'^T dY) + cb[c]*x + cd[c]
//             -> the dgrad GEMM runs on W' and adds cb*x + cd in its epilogue (dk_pwconv_dgrad_affine).
//
// dk_bn_fold_bwd turns (G, S) into dW (+ l2*W), dgamma, dbeta and the epilogue coefficients cb, cd.  Against the unfused
// sequence this removes the BatchNorm's normalisation pass (read + write of the activation) in forward and its whole backward
// kernel (dY and x read, dx written) -- the dgrad epilogue reads x once instead.
#include "common.cuh"

namespace dk {

// one CTA per filter f.  The tensor core multiplies TF32-TRUNCATED operands (kind::tf32 drops the low 13 mantissa bits of the
// fp32 containers), so the folded weights are stored already truncated, Wt = trunc(W*scale) -- the GEMM result is the same --
// and the bias is built from what the GEMM will really add up, so that the per-channel MEAN of its output stays exact although
// it now multiplies un-centred activations (mean mu, truncation loss rho = mean(x - trunc(x)), from the producer's statistics):
//   mean_p(Wt . trunc(x)) = sum_c Wt[f,c]*(mu[c] - rho[c])   and the normalised mean is   sum_c W[f,c]*(shift[c] + scale[c]*mu[c])
//   =>  b'[f] = b[f] + sum_c ( W[f,c]*shift[c] + (W[f,c]*scale[c] - Wt[f,c])*mu[c] + Wt[f,c]*rho[c] )
// Without the last two terms every output channel carries an offset of ~2^-11 * |W' . mu| (measured: 1e-4 .. 3e-4 of the
// following BatchNorm's running mean in ResNet-18-depsep).  mean / resid NULL: the plain b + W.shift.  Fixed order: deterministic.
__global__ void __launch_bounds__(128)
bn_fold_fwd_kernel(const float *__restrict__ w, const float *__restrict__ bias, const float *__restrict__ scale,
                   const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ resid,
                   float *__restrict__ w_out, float *__restrict__ b_out, int C) {
    __shared__ float red[33];
    const int f = blockIdx.x;
    const float *wr = w + (long long)f * C;
    float *wo = w_out + (long long)f * C;
    float acc = 0.0f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float v = wr[c];
        const float ws = v * __ldg(scale + c);
        const float wt = mean ? __uint_as_float(__float_as_uint(ws) & 0xFFFFE000u) : ws;
        wo[c] = wt;
        acc = fmaf(v, __ldg(shift + c), acc);
        if (mean) acc = fmaf(ws - wt, __ldg(mean + c), acc);
        if (resid) acc = fmaf(wt, __ldg(resid + c), acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) b_out[f] = acc + (bias ? bias[f] : 0.0f);
}

// A CTA owns 32 input channels: thread (tx, ty) = (channel lane, filter lane) walks f = ty, ty + 32, ... (coalesced over c for
// every f), the 32 filter lanes are added in a fixed order through shared memory.  Channel sums in double: sum_gx - mu*sum_g
// cancels.  (The first version ran one thread per channel over all F filters: a latency-bound serial loop on 1-4 CTAs.)
__global__ void __launch_bounds__(1024)
bn_fold_bwd_kernel(const float *__restrict__ g_raw, const float *__restrict__ s_col, const float *__restrict__ w,
                   const float *__restrict__ mean, const float *__restrict__ invstd,
                   const float *__restrict__ scale, const float *__restrict__ shift, float l2, float inv_count,
                   float *__restrict__ dw, float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ cb,
                   float *__restrict__ cd, int F, int C) {
    __shared__ double red_g[32][33], red_gx[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const bool ok = c < C;
    double sum_g = 0.0, sum_gx = 0.0;
    if (ok) {
        for (int f = ty; f < F; f += 32) {
            const float wv = w[(long long)f * C + c];
            if (s_col) sum_g += (double)wv * (double)__ldg(s_col + f);
            sum_gx += (double)wv * (double)g_raw[(long long)f * C + c];
        }
    }
    red_g[ty][tx] = sum_g;
    red_gx[ty][tx] = sum_gx;
    __syncthreads();
    if (ty == 0 && ok) {
        sum_g = 0.0;
        sum_gx = 0.0;
        for (int k = 0; k < 32; ++k) {
            sum_g += red_g[k][tx];
            sum_gx += red_gx[k][tx];
        }
        const double mu = mean[c], is = invstd[c], sc = scale[c];
        const double sg_xhat = is * (sum_gx - mu * sum_g);  // sum_p dXh * (x - mu) * invstd
        dgamma[c] = (float)sg_xhat;
        dbeta[c] = (float)sum_g;
        const double b = -sc * is * sg_xhat * (double)inv_count;
        cb[c] = (float)b;
        cd[c] = (float)(-sc * sum_g * (double)inv_count - b * mu);
    }
    if (!ok) return;
    const float scf = scale[c], shf = shift[c];
    for (int f = ty; f < F; f += 32) {
        const long long i = (long long)f * C + c;
        float v = scf * g_raw[i];
        if (s_col) v = fmaf(shf, __ldg(s_col + f), v);
        if (l2 != 0.0f) v = fmaf(l2, w[i], v);
        dw[i] = v;
    }
}

// out[n,c,p] += cb[c]*x[n,c,p] + cd[c]   (the dgrad epilogue of the SIMT / fallback path)
__global__ void __launch_bounds__(256)
affine_add_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ cb,
                  const float *__restrict__ cd, long long total, int C, int P) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / P) % C);
        out[i] = fmaf(__ldg(cb + c), __ldg(x + i), out[i]) + __ldg(cd + c);
    }
}

int affine_add_launch(float *out, const float *x, const float *cb, const float *cd, int N, int C, int P, cudaStream_t st) {
    const long long total = (long long)N * C * P;
    affine_add_kernel<<<stream_grid(total, 256), 256, 0, st>>>(out, x, cb, cd, total, C, P);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk

using namespace dk;

extern "C" {

int dk_bn_fold_fwd(const float *w, const float *bias, const float *scale, const float *shift, const float *mean,
                   const float *trunc_resid, float *w_out, float *b_out, int F, int C, dk_stream_t stream) {
    DK_REQUIRE(F > 0 && C > 0 && w && scale && shift && w_out && b_out, "dk_bn_fold_fwd: bad arguments");
    DK_REQUIRE(trunc_resid == nullptr || mean != nullptr, "dk_bn_fold_fwd: trunc_resid needs mean");
    if (gemm_backend() != 0) mean = trunc_resid = nullptr;  // fp32 SIMT GEMMs truncate nothing: the plain fold is exact
    bn_fold_fwd_kernel<<<F, 128, 0, as_stream(stream)>>>(w, bias, scale, shift, mean, trunc_resid, w_out, b_out, C);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dk_bn_fold_bwd(const float *g_raw, const float *s_col, const float *w, const float *gamma, const float *save_mean,
                   const float *save_invstd, const float *save_scale, const float *save_shift, float l2, int64_t count,
                   float *dw, float *dgamma, float *dbeta, float *cb, float *cd, int F, int C, dk_stream_t stream) {
    DK_REQUIRE(F > 0 && C > 0 && count > 0 && g_raw && w && gamma && save_mean && save_invstd && save_scale && save_shift &&
                   dw && dgamma && dbeta && cb && cd,
               "dk_bn_fold_bwd: bad arguments");
    bn_fold_bwd_kernel<<<(C + 31) / 32, dim3(32, 32), 0, as_stream(stream)>>>(g_raw, s_col, w, save_mean, save_invstd, save_scale,
                                                                             save_shift, l2, 1.0f / (float)count, dw, dgamma,
                                                                             dbeta, cb, cd, F, C);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // extern "C"
