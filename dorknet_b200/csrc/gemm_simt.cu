// gemm_simt.cu -- plain CUDA-core implicit GEMM for conv / pointwise / dense (fwd, dgrad, wgrad).
//
// Role: (1) GPU-side cross-check for the tcgen05 kernels in tests (dk_set_gemm_backend(1));
// (2) the path for shapes the tensor-core kernels do not cover.  No patch matrix is ever
// materialised: operand loaders compute the im2col / col2im index maps on the fly.
//
// C[M, Nn] = sum_k A(m, k) * B(k, n), 64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
#include "common.cuh"
#include "gemm.cuh"

namespace dk {

constexpr int BM = 64, BN = 64, BK = 16, SG_THREADS = 256;

struct ConvShape {
    int N, C, H, W, F, kh, kw, s, p, OH, OW;
};

// ---- operand loaders --------------------------------------------------------------------------
// forward: A = im2col(X) [N*OH*OW, C*kh*kw]
struct FwdA {
    const float *x;
    ConvShape g;
    struct Row { int64_t base; int ih0, iw0; bool ok; };
    __device__ Row row(int64_t m, int64_t M) const {
        Row r;
        r.ok = m < M;
        const int ohw = g.OH * g.OW;
        const int n = (int)(m / ohw), pix = (int)(m - (int64_t)n * ohw);
        const int oh = pix / g.OW, ow = pix - oh * g.OW;
        r.base = (int64_t)n * g.C * g.H * g.W;
        r.ih0 = oh * g.s - g.p;
        r.iw0 = ow * g.s - g.p;
        return r;
    }
    __device__ float load(const Row &r, int k, int K) const {
        if (!r.ok || k >= K) return 0.0f;
        const int kk = g.kh * g.kw;
        const int c = k / kk, t = k - c * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ih = r.ih0 + i, iw = r.iw0 + j;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return 0.0f;
        return __ldg(x + r.base + ((int64_t)c * g.H + ih) * g.W + iw);
    }
};
struct FwdB {  // B(k, f) = W[f, k]
    const float *w;
    struct Col { int64_t base; bool ok; };
    __device__ Col col(int n, int Nn, int K) const { return Col{(int64_t)n * K, n < Nn}; }
    __device__ float load(const Col &c, int k, int K) const { return (c.ok && k < K) ? __ldg(w + c.base + k) : 0.0f; }
};
struct FwdStore {  // Y[n, f, oh, ow] (+ bias[f])
    float *y;
    const float *bias;
    ConvShape g;
    __device__ void store(int64_t m, int n, float v, int) const {
        const int ohw = g.OH * g.OW;
        const int img = (int)(m / ohw), pix = (int)(m - (int64_t)img * ohw);
        y[((int64_t)img * g.F + n) * ohw + pix] = v + (bias ? __ldg(bias + n) : 0.0f);
    }
};

// dgrad: M = N*H*W input pixels, Nn = C, K = F*kh*kw; gather form of col2im
struct DgradA {
    const float *dy;
    ConvShape g;
    struct Row { int64_t base; int hp, wp; bool ok; };
    __device__ Row row(int64_t m, int64_t M) const {
        Row r;
        r.ok = m < M;
        const int hw = g.H * g.W;
        const int n = (int)(m / hw), pix = (int)(m - (int64_t)n * hw);
        const int h = pix / g.W, w = pix - h * g.W;
        r.base = (int64_t)n * g.F * g.OH * g.OW;
        r.hp = h + g.p;
        r.wp = w + g.p;
        return r;
    }
    __device__ float load(const Row &r, int k, int K) const {
        if (!r.ok || k >= K) return 0.0f;
        const int kk = g.kh * g.kw;
        const int f = k / kk, t = k - f * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ti = r.hp - i, tj = r.wp - j;
        if (ti < 0 || tj < 0 || (ti % g.s) != 0 || (tj % g.s) != 0) return 0.0f;
        const int oh = ti / g.s, ow = tj / g.s;
        if (oh >= g.OH || ow >= g.OW) return 0.0f;
        return __ldg(dy + r.base + ((int64_t)f * g.OH + oh) * g.OW + ow);
    }
};
struct DgradB {  // B(k=(f,i,j), c) = W[f, c, i, j]
    const float *w;
    ConvShape g;
    struct Col { int c; bool ok; };
    __device__ Col col(int n, int Nn, int) const { return Col{n, n < Nn}; }
    __device__ float load(const Col &c, int k, int K) const {
        if (!c.ok || k >= K) return 0.0f;
        const int kk = g.kh * g.kw;
        const int f = k / kk, t = k - f * kk;
        return __ldg(w + ((int64_t)f * g.C + c.c) * kk + t);
    }
};
struct DgradStore {
    float *dx;
    ConvShape g;
    __device__ void store(int64_t m, int n, float v, int) const {
        const int hw = g.H * g.W;
        const int img = (int)(m / hw), pix = (int)(m - (int64_t)img * hw);
        dx[((int64_t)img * g.C + n) * hw + pix] = v;
    }
};

// wgrad: M = F, Nn = C*kh*kw, K = N*OH*OW (split over blockIdx.z)
struct WgradA {  // A(f, kk) = dY[n, f, pix]
    const float *dy;
    ConvShape g;
    struct Row { int f; bool ok; };
    __device__ Row row(int64_t m, int64_t M) const { return Row{(int)m, m < M}; }
    __device__ float load(const Row &r, int64_t k, int64_t K) const {
        if (!r.ok || k >= K) return 0.0f;
        const int ohw = g.OH * g.OW;
        const int n = (int)(k / ohw), pix = (int)(k - (int64_t)n * ohw);
        return __ldg(dy + ((int64_t)n * g.F + r.f) * ohw + pix);
    }
};
struct WgradB {  // B(kk, col=(c,i,j)) = Xpad[n, c, oh*s+i, ow*s+j]
    const float *x;
    ConvShape g;
    struct Col { int c, i, j; bool ok; };
    __device__ Col col(int n, int Nn, int64_t) const {
        Col c;
        c.ok = n < Nn;
        const int kk = g.kh * g.kw;
        c.c = n / kk;
        const int t = n - c.c * kk;
        c.i = t / g.kw;
        c.j = t - c.i * g.kw;
        return c;
    }
    __device__ float load(const Col &c, int64_t k, int64_t K) const {
        if (!c.ok || k >= K) return 0.0f;
        const int ohw = g.OH * g.OW;
        const int n = (int)(k / ohw), pix = (int)(k - (int64_t)n * ohw);
        const int oh = pix / g.OW, ow = pix - oh * g.OW;
        const int ih = oh * g.s + c.i - g.p, iw = ow * g.s + c.j - g.p;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) return 0.0f;
        return __ldg(x + (((int64_t)n * g.C + c.c) * g.H + ih) * g.W + iw);
    }
};
struct PartialStore {  // partial[z][m*Nn + n]
    float *partial;
    int64_t mn;
    int Nn;
    __device__ void store(int64_t m, int n, float v, int z) const { partial[(int64_t)z * mn + m * Nn + n] = v; }
};

// dense: generic strided matrices  A(m,k) = a[m*am + k*ak], B(k,n) = b[k*bk + n*bn]
struct MatA {
    const float *a;
    int64_t am, ak;
    struct Row { int64_t base; bool ok; };
    __device__ Row row(int64_t m, int64_t M) const { return Row{m * am, m < M}; }
    __device__ float load(const Row &r, int64_t k, int64_t K) const { return (r.ok && k < K) ? __ldg(a + r.base + k * ak) : 0.0f; }
};
struct MatB {
    const float *b;
    int64_t bk, bn;
    struct Col { int64_t base; bool ok; };
    __device__ Col col(int n, int Nn, int64_t) const { return Col{(int64_t)n * bn, n < Nn}; }
    __device__ float load(const Col &c, int64_t k, int64_t K) const { return (c.ok && k < K) ? __ldg(b + c.base + k * bk) : 0.0f; }
};
struct MatStore {  // out[m*ldo + n] (+ bias[n])
    float *out;
    const float *bias;
    int64_t ldo;
    __device__ void store(int64_t m, int n, float v, int) const { out[m * ldo + n] = v + (bias ? __ldg(bias + n) : 0.0f); }
};

// ---- the kernel ---------------------------------------------------------------------------------
// A_KCONTIG / B_KCONTIG pick which index runs fastest across a warp when loading from global.
template <class AL, class BL, class ST, bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(AL al, BL bl, ST st, int64_t M, int Nn, int64_t K, int64_t k_per_split) {
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
    const int64_t kend = kbeg + k_per_split < K ? kbeg + k_per_split : K;

    // load mapping
    int a_m[4], a_k[4], b_n[4], b_k[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (A_KCONTIG) { a_k[j] = t % BK; a_m[j] = t / BK + (SG_THREADS / BK) * j; }
        else           { a_m[j] = t % BM; a_k[j] = t / BM + (SG_THREADS / BM) * j; }
        if (B_KCONTIG) { b_k[j] = t % BK; b_n[j] = t / BK + (SG_THREADS / BK) * j; }
        else           { b_n[j] = t % BN; b_k[j] = t / BN + (SG_THREADS / BN) * j; }
    }
    typename AL::Row arow[4];
    typename BL::Col bcol[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (A_KCONTIG || j == 0) arow[j] = al.row(m0 + a_m[j], M);
        else arow[j] = arow[0];
        if (B_KCONTIG || j == 0) bcol[j] = bl.col(n0 + b_n[j], Nn, K);
        else bcol[j] = bcol[0];
    }

    const int tx = t % 16, ty = t / 16;  // tx -> m (4 consecutive rows), ty -> n
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[a_k[j]][a_m[j]] = al.load(arow[j], k0 + a_k[j], kend);
            Bs[b_k[j]][b_n[j]] = bl.load(bcol[j], k0 + b_k[j], kend);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + ty * 4 + j;
        if (n >= Nn) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + tx * 4 + i;
            if (m < M) st.store(m, n, acc[i][j], blockIdx.z);
        }
    }
}

// out[i] = sum_z partial[z][i] + l2*w[i]
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w,
                                     float *__restrict__ out, float l2, int64_t mn, int Z) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += nthreads) {
        float s = 0.0f;
        for (int z = 0; z < Z; ++z) s += partial[(int64_t)z * mn + i];
        out[i] = s + (l2 != 0.0f ? l2 * w[i] : 0.0f);
    }
}

// Same reduction with the Z partials spread over 8 thread rows (fixed combination order -> deterministic): the
// split-K GEMMs leave up to 148 partials of a small [M][N] matrix, which a one-thread-per-output loop reads as a
// serial chain of dependent-latency loads.
constexpr int SKR_X = 32;
template <int SKR_Y>
__global__ void __launch_bounds__(SKR_X * SKR_Y)
splitk_reduce2_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ out, float l2,
                      int64_t mn, int Z) {
    __shared__ float red[SKR_Y][SKR_X + 1];
    const int tx = threadIdx.x % SKR_X, ty = threadIdx.x / SKR_X;
    const int64_t i = (int64_t)blockIdx.x * SKR_X + tx;
    float s = 0.0f;
    if (i < mn) {
        int z = ty;
        for (; z + 3 * SKR_Y < Z; z += 4 * SKR_Y) {
            const float a = partial[(int64_t)z * mn + i], b = partial[(int64_t)(z + SKR_Y) * mn + i];
            const float c = partial[(int64_t)(z + 2 * SKR_Y) * mn + i], d = partial[(int64_t)(z + 3 * SKR_Y) * mn + i];
            s += (a + b) + (c + d);
        }
        for (; z < Z; z += SKR_Y) s += partial[(int64_t)z * mn + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < mn) {
        float t = red[0][tx];
#pragma unroll
        for (int y = 1; y < SKR_Y; ++y) t += red[y][tx];
        out[i] = t + (l2 != 0.0f ? l2 * w[i] : 0.0f);
    }
}

void splitk_reduce_launch(const float *partial, const float *w, float *out, float l2, int64_t mn, int Z, cudaStream_t st) {
    // Z >= 64 (a small [M][N] with ~148 partials: the 56x56 / 28x28 pointwise wgrads): 32 thread rows, so that a thread's
    // 4-5 loads are one batch in flight instead of a chain of five
    if (Z >= 64) splitk_reduce2_kernel<32><<<(unsigned)ceil_div(mn, SKR_X), SKR_X * 32, 0, st>>>(partial, w, out, l2, mn, Z);
    else if (Z >= 16) splitk_reduce2_kernel<8><<<(unsigned)ceil_div(mn, SKR_X), SKR_X * 8, 0, st>>>(partial, w, out, l2, mn, Z);
    else splitk_reduce_kernel<<<stream_grid(mn, 256), 256, 0, st>>>(partial, w, out, l2, mn, Z);
}

template <class AL, class BL, class ST, bool AK, bool BK_>
static int launch_gemm(AL al, BL bl, ST st, int64_t M, int Nn, int64_t K, int Z, int64_t k_per_split, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(Nn, BN), (unsigned)Z);
    DK_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "simt gemm: grid too large");
    gemm_simt_kernel<AL, BL, ST, AK, BK_><<<grid, SG_THREADS, 0, s>>>(al, bl, st, M, Nn, K, k_per_split);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// split-K plan for the long reductions of wgrad
static void plan_splitk(int64_t M, int Nn, int64_t K, int *Z, int64_t *k_per_split) {
    const int64_t tiles = ceil_div(M, BM) * ceil_div(Nn, BN);
    int64_t z = ceil_div((int64_t)sm_count() * 2, tiles);
    const int64_t zmax = ceil_div(K, 4 * BK);
    if (z > zmax) z = zmax;
    if (z > 1024) z = 1024;
    if (z < 1) z = 1;
    int64_t per = ceil_div(ceil_div(K, z), BK) * BK;
    *Z = (int)ceil_div(K, per);
    *k_per_split = per;
}

size_t simt_wgrad_ws_bytes(int64_t M, int Nn, int64_t K) {
    int Z;
    int64_t per;
    plan_splitk(M, Nn, K, &Z, &per);
    return (size_t)Z * M * Nn * sizeof(float);
}

static ConvShape mk_shape(int N, int C, int H, int W, int F, int kh, int kw, int s, int p) {
    ConvShape g{N, C, H, W, F, kh, kw, s, p, (H + 2 * p - kh) / s + 1, (W + 2 * p - kw) / s + 1};
    return g;
}

int simt_conv_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                  int kh, int kw, int s, int p, cudaStream_t st) {
    const ConvShape g = mk_shape(N, C, H, W, F, kh, kw, s, p);
    const int64_t M = (int64_t)N * g.OH * g.OW;
    const int K = C * kh * kw;
    return launch_gemm<FwdA, FwdB, FwdStore, false, true>(FwdA{x, g}, FwdB{w}, FwdStore{y, bias, g}, M, F, K, 1, K, st);
}

int simt_conv_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int s, int p, int OH, int OW, cudaStream_t st) {
    ConvShape g = mk_shape(N, C, H, W, F, kh, kw, s, p);
    g.OH = OH;  // the pointwise layer's dx is [OH*s, OW*s], so OH/OW are given, not derived
    g.OW = OW;
    const int64_t M = (int64_t)N * H * W;
    const int K = F * kh * kw;
    return launch_gemm<DgradA, DgradB, DgradStore, false, false>(DgradA{dy, g}, DgradB{w, g}, DgradStore{dx, g}, M, C,
                                                                K, 1, K, st);
}

int simt_conv_wgrad(const float *dy, const float *x, const float *w, float *dw, float l2, int N, int C, int H, int W,
                    int F, int kh, int kw, int s, int p, void *ws, size_t ws_bytes, cudaStream_t st) {
    const ConvShape g = mk_shape(N, C, H, W, F, kh, kw, s, p);
    const int Nn = C * kh * kw;
    const int64_t K = (int64_t)N * g.OH * g.OW;
    int Z;
    int64_t per;
    plan_splitk(F, Nn, K, &Z, &per);
    const int64_t mn = (int64_t)F * Nn;
    if (ws == nullptr || ws_bytes < (size_t)Z * mn * sizeof(float)) {
        set_error("conv wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, (size_t)Z * mn * sizeof(float));
        return DK_ERR_WORKSPACE;
    }
    float *partial = reinterpret_cast<float *>(ws);
    int rc = launch_gemm<WgradA, WgradB, PartialStore, true, true>(WgradA{dy, g}, WgradB{x, g},
                                                                   PartialStore{partial, mn, Nn}, F, Nn, K, Z, per, st);
    if (rc) return rc;
    splitk_reduce_kernel<<<stream_grid(mn, 256), 256, 0, st>>>(partial, w, dw, l2, mn, Z);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int simt_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in_dim, int out_dim,
                   cudaStream_t st) {
    return launch_gemm<MatA, MatB, MatStore, true, false>(MatA{x, in_dim, 1}, MatB{w, out_dim, 1},
                                                          MatStore{y, bias, out_dim}, B, out_dim, in_dim, 1, in_dim, st);
}

int simt_dense_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float l2, int B, int in_dim,
                   int out_dim, void *ws, size_t ws_bytes, cudaStream_t st) {
    // dx[B,in] = dy[B,out] @ w^T : A(m,k) = dy[m*out + k], B(k,n) = w[n*out + k]
    int rc = launch_gemm<MatA, MatB, MatStore, true, true>(MatA{dy, out_dim, 1}, MatB{w, 1, out_dim},
                                                           MatStore{dx, nullptr, in_dim}, B, in_dim, out_dim, 1,
                                                           out_dim, st);
    if (rc) return rc;
    // dw[in,out] = x^T @ dy : A(m,k) = x[k*in + m], B(k,n) = dy[k*out + n]; K = B (split-K)
    int Z;
    int64_t per;
    plan_splitk(in_dim, out_dim, B, &Z, &per);
    const int64_t mn = (int64_t)in_dim * out_dim;
    if (ws == nullptr || ws_bytes < (size_t)Z * mn * sizeof(float)) {
        set_error("dense bwd: workspace too small (%zu < %zu bytes)", ws_bytes, (size_t)Z * mn * sizeof(float));
        return DK_ERR_WORKSPACE;
    }
    float *partial = reinterpret_cast<float *>(ws);
    rc = launch_gemm<MatA, MatB, PartialStore, false, false>(MatA{x, 1, in_dim}, MatB{dy, out_dim, 1},
                                                            PartialStore{partial, mn, out_dim}, in_dim, out_dim, B, Z,
                                                            per, st);
    if (rc) return rc;
    splitk_reduce_kernel<<<stream_grid(mn, 256), 256, 0, st>>>(partial, w, dw, l2, mn, Z);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

size_t simt_dense_ws_bytes(int B, int in_dim, int out_dim) { return simt_wgrad_ws_bytes(in_dim, out_dim, B); }

// debug: the reference's patch matrix, bit-exact index map (layers/im2col.pyx:33-34)
__global__ void im2col_kernel(const float *__restrict__ x, float *__restrict__ P, ConvShape g, int64_t total) {
    const int K = g.C * g.kh * g.kw;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthreads) {
        const int64_t m = idx / K;
        const int k = (int)(idx - m * K);
        const int ohw = g.OH * g.OW;
        const int n = (int)(m / ohw), pix = (int)(m - (int64_t)n * ohw);
        const int oh = pix / g.OW, ow = pix - oh * g.OW;
        const int kk = g.kh * g.kw;
        const int c = k / kk, t = k - c * kk;
        const int i = t / g.kw, j = t - i * g.kw;
        const int ih = oh * g.s + i - g.p, iw = ow * g.s + j - g.p;
        float v = 0.0f;
        if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) v = x[(((int64_t)n * g.C + c) * g.H + ih) * g.W + iw];
        P[idx] = v;
    }
}

int simt_im2col(const float *x, float *P, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st) {
    const ConvShape g = mk_shape(N, C, H, W, 0, kh, kw, s, p);
    const int64_t total = (int64_t)N * g.OH * g.OW * C * kh * kw;
    if (total == 0) return DK_OK;
    im2col_kernel<<<stream_grid(total, 256), 256, 0, st>>>(x, P, g, total);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
