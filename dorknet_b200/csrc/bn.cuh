// bn.cuh -- pieces shared by the BatchNorm kernels (batchnorm.cu: split statistics / apply / backward passes;
// bn_fused.cu: the cluster kernels that keep a channel slice resident in shared memory between the passes).
#pragma once
#include "common.cuh"

namespace dk {

struct Moments {
    float n, mean, m2;
};

__device__ __forceinline__ Moments merge(const Moments &a, const Moments &b) {
    Moments r;
    r.n = a.n + b.n;
    if (r.n == 0.0f) {
        r.mean = 0.0f;
        r.m2 = 0.0f;
        return r;
    }
    const float delta = b.mean - a.mean;
    const float fb = b.n / r.n;
    r.mean = a.mean + delta * fb;
    r.m2 = a.m2 + b.m2 + delta * delta * a.n * fb;
    return r;
}

struct BnFinalize {
    // mode 0: write mean/var only (dk_bn_stats); mode 1: full training finalise
    int mode;
    float *mean_out, *var_out;
    const float *gamma, *beta;
    float *running_mean, *running_std;
    int first_batch;
    float momentum, eps;
    float *save_mean, *save_invstd, *save_scale, *save_shift;
};

// writes the per-channel results of the statistics (batch_norm.py:66-89); returns scale/shift for the apply pass
__device__ __forceinline__ void bn_finalize_channel(const BnFinalize &fin, int c, float mean, float var, bool write,
                                                    float *scale_out, float *shift_out) {
    if (fin.mode == 0) {
        if (write) {
            fin.mean_out[c] = mean;
            fin.var_out[c] = var;
        }
        *scale_out = 1.0f;
        *shift_out = 0.0f;
        return;
    }
    const float std = sqrtf(var + fin.eps);  // batch_norm.py:69
    const float invstd = 1.0f / std;
    const float scale = fin.gamma[c] * invstd;
    const float shift = fin.beta[c] - mean * scale;
    *scale_out = scale;
    *shift_out = shift;
    if (!write) return;
    fin.save_mean[c] = mean;
    fin.save_invstd[c] = invstd;
    fin.save_scale[c] = scale;
    fin.save_shift[c] = shift;
    if (fin.running_mean) {  // batch_norm.py:76-89 (tracks std, not var)
        if (fin.first_batch) {
            fin.running_mean[c] = mean;
            fin.running_std[c] = std;
        } else {
            const float mo = fin.momentum;
            fin.running_mean[c] = mo * fin.running_mean[c] + (1.0f - mo) * mean;
            fin.running_std[c] = mo * fin.running_std[c] + (1.0f - mo) * std;
        }
    }
}

// The same finalise with the channel's parameters already in registers: the loads are issued when the kernel starts and
// have landed long before the statistics are ready, instead of sitting on the critical path between the two passes.
struct BnChannelParams {
    float gamma, beta, rmean, rstd;
};
__device__ __forceinline__ BnChannelParams bn_load_channel_params(const BnFinalize &fin, int c) {
    BnChannelParams p = {1.0f, 0.0f, 0.0f, 0.0f};
    if (fin.mode != 0) {
        p.gamma = fin.gamma[c];
        p.beta = fin.beta[c];
        if (fin.running_mean && !fin.first_batch) {
            p.rmean = fin.running_mean[c];
            p.rstd = fin.running_std[c];
        }
    }
    return p;
}
__device__ __forceinline__ void bn_finalize_channel(const BnFinalize &fin, const BnChannelParams &p, int c, float mean, float var,
                                                    bool write, float *scale_out, float *shift_out) {
    if (fin.mode == 0) {
        if (write) {
            fin.mean_out[c] = mean;
            fin.var_out[c] = var;
        }
        *scale_out = 1.0f;
        *shift_out = 0.0f;
        return;
    }
    const float std = sqrtf(var + fin.eps);  // batch_norm.py:69
    const float invstd = 1.0f / std;
    const float scale = p.gamma * invstd;
    const float shift = p.beta - mean * scale;
    *scale_out = scale;
    *shift_out = shift;
    if (!write) return;
    fin.save_mean[c] = mean;
    fin.save_invstd[c] = invstd;
    fin.save_scale[c] = scale;
    fin.save_shift[c] = shift;
    if (fin.running_mean) {  // batch_norm.py:76-89 (tracks std, not var)
        if (fin.first_batch) {
            fin.running_mean[c] = mean;
            fin.running_std[c] = std;
        } else {
            const float mo = fin.momentum;
            fin.running_mean[c] = mo * p.rmean + (1.0f - mo) * mean;
            fin.running_std[c] = mo * p.rstd + (1.0f - mo) * std;
        }
    }
}

// bn_fused.cu: return DK_ERR_UNSUPPORTED (no error text) when the channel slices do not fit shared memory
int bn_fused_init();
// add != nullptr: y = relu?(x*scale + shift + add) -- the residual join folded into the normalisation pass
int bn_fused_fwd(const float *x, float *y, const BnFinalize &fin, int relu, int N, int C, int HW, cudaStream_t st,
                 const float *add = nullptr);
// join_out != nullptr: the gradient is first masked by a ResidualBlock's ReLU, g = dy * (join_out > 0), and g is also written
// to join_g (the skip path's gradient); relu must be 0 then
int bn_fused_bwd(const float *dy, const float *x, const float *save_mean, const float *save_invstd, const float *save_scale,
                 const float *save_shift, float *dx, float *dgamma, float *dbeta, int relu, int N, int C, int HW,
                 cudaStream_t st, const float *join_out = nullptr, float *join_g = nullptr);
extern int g_bn_fused_enabled;  // 0: split kernels only; 1: group kernels (small planes) and cluster kernels; 2: cluster kernels only

// bn_group.cu: channel-group kernels for small planes (same contract: DK_ERR_UNSUPPORTED when the shape does not fit)
int bn_group_init();
int bn_group_fwd(const float *x, float *y, const BnFinalize &fin, int relu, int N, int C, int HW, cudaStream_t st,
                 const float *add = nullptr);
int bn_group_bwd(const float *dy, const float *x, const float *save_mean, const float *save_invstd, const float *save_scale,
                 const float *save_shift, float *dx, float *dgamma, float *dbeta, int relu, int N, int C, int HW,
                 cudaStream_t st, const float *join_out = nullptr, float *join_g = nullptr);

}  // namespace dk
