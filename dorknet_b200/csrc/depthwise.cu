// depthwise.cu -- depthwise convolution forward and fused backward (dX + dW + db), NCHW float32.
//
// Not a contraction: K = kh*kw (9) taps per output, so this is an HBM-bound stencil.  A CTA stages
// the input rows it needs for G consecutive (n,c) planes (planes are adjacent in NCHW memory, so the
// global reads are long coalesced runs) into zero-padded shared-memory tiles and every thread
// computes strips of 4 adjacent outputs from shared memory.  Tiny planes (7x7 ... 28x28) are
// batched G = 32 ... 4 per CTA; planes too big for the tile budget are cut into row bands.
//
// Backward is one pass over dY and X: the CTA stages a dY tile (with the halo dX needs) and the X tile
// the filter gradient needs, writes dX, and reduces per-plane dW/db partials with shuffles in a
// fixed order (deterministic, no atomics -- the reference's GPU kernel used N*OH*OW-way atomicAdd
// contention, layers/depthwise_convolution.py:132-134).  A second tiny kernel sums the partials over
// images, as the reference sums its per-image dw (layers/depthwise_convolution.py:193).
//
// Optional input transform relu?(x*scale[c]+shift[c]): a deferred BatchNorm(+ReLU) of the producer
// applied while staging, so the normalised activation is never written to HBM.
#include "bn.cuh"
#include "common.cuh"

namespace dk {

constexpr int DW_THREADS = 256;
constexpr int DW_FWD_TILE_BUDGET = 32 * 1024;
constexpr int DW_BWD_TILE_BUDGET = 48 * 1024;
constexpr int DW_MAX_TAPS = 49;

struct DwGeom {
    int N, C, H, W, kh, kw, s, p, OH, OW;
    int G, tpp, bands;
    int64_t planes;
    // forward tile
    int f_rows_in_max, f_ws;
    // backward tiles
    int PL, PR, dws, xws, b_dy_rows_max, b_x_rows_max;
};

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// row ranges of backward band `b`
struct DwBand {
    int h0, h1;      // dX rows written
    int oo0, oo1;    // output rows owned for dW/db
    int dy_lo, dy_hi;  // dY rows staged (may lie outside [0,OH): zero filled)
    int x_lo, x_hi;    // X rows staged (may lie outside [0,H): zero filled)
};

__host__ __device__ inline DwBand dw_band(const DwGeom &g, int b) {
    DwBand r;
    const int bh = cdiv(g.H, g.bands), bo = cdiv(g.OH, g.bands);
    r.h0 = min(b * bh, g.H);
    r.h1 = min(r.h0 + bh, g.H);
    r.oo0 = min(b * bo, g.OH);
    r.oo1 = min(r.oo0 + bo, g.OH);
    int lo = 1 << 30, hi = -(1 << 30);
    if (r.h0 < r.h1) {
        lo = floor_div(r.h0 + g.p - (g.kh - 1), g.s);
        hi = floor_div(r.h1 - 1 + g.p, g.s) + 1;
    }
    if (r.oo0 < r.oo1) {
        lo = min(lo, r.oo0);
        hi = max(hi, r.oo1);
        r.x_lo = r.oo0 * g.s - g.p;
        r.x_hi = (r.oo1 - 1) * g.s - g.p + g.kh;
    } else {
        r.x_lo = 0;
        r.x_hi = 0;
    }
    if (lo > hi) { lo = 0; hi = 0; }
    r.dy_lo = lo;
    r.dy_hi = hi;
    return r;
}

__device__ __forceinline__ float dw_xform(float v, float sc, float sh, int relu) {
    v = fmaf(v, sc, sh);
    return (relu && !(v > 0.0f)) ? 0.0f : v;
}

// Stage rows [row_lo, row_lo + nrows) x cols [col_lo, col_lo + stride_s) of G planes of a [planes, SH, SW]
// tensor into smem (zero outside the tensor).  XFORM applies the per-channel affine(+relu) to in-bounds data.
template <bool XFORM>
__device__ __forceinline__ void dw_stage(float *__restrict__ tile, int plane_stride, int stride_s,
                                         const float *__restrict__ src, int64_t g0, int64_t planes, int G,
                                         int SH, int SW, int row_lo, int nrows, int col_lo, int C,
                                         const float *__restrict__ in_scale, const float *__restrict__ in_shift,
                                         int in_relu) {
    const int per_plane = nrows * stride_s;
    const int total = G * per_plane;
    for (int idx = threadIdx.x; idx < total; idx += DW_THREADS) {
        const int gl = idx / per_plane;
        const int rem = idx - gl * per_plane;
        const int r = rem / stride_s;
        const int cc = rem - r * stride_s;
        const int ih = row_lo + r, iw = col_lo + cc;
        const int64_t plane = g0 + gl;
        float v = 0.0f;
        if (plane < planes && ih >= 0 && ih < SH && iw >= 0 && iw < SW) {
            v = __ldg(src + (plane * SH + ih) * (int64_t)SW + iw);
            if (XFORM) {
                const int c = (int)(plane % C);
                v = dw_xform(v, __ldg(in_scale + c), __ldg(in_shift + c), in_relu);
            }
        }
        tile[gl * plane_stride + r * stride_s + cc] = v;
    }
}

// ---------------------------------------------------------------------------------------- forward
template <int KH, int KW, int S>
__global__ void __launch_bounds__(DW_THREADS)
dw_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
              float *__restrict__ y, const float *__restrict__ in_scale, const float *__restrict__ in_shift,
              int in_relu, DwGeom g) {
    extern __shared__ float smem[];
    const int kh = KH ? KH : g.kh, kw = KW ? KW : g.kw, s = S ? S : g.s;
    const int taps = kh * kw;
    const int band = blockIdx.y;
    const int64_t g0 = (int64_t)blockIdx.x * g.G;
    const int bro = cdiv(g.OH, g.bands);
    const int oh0 = band * bro;
    const int oh1 = min(oh0 + bro, g.OH);
    if (oh0 >= oh1) return;
    const int rows_out = oh1 - oh0;
    const int rows_in = (rows_out - 1) * s + kh;
    const int ws = g.f_ws;
    const int plane_stride = g.f_rows_in_max * ws;
    float *tile = smem;
    float *wsm = smem + g.G * plane_stride;  // [G][taps + 1]

    if (in_scale)
        dw_stage<true>(tile, plane_stride, ws, x, g0, g.planes, g.G, g.H, g.W, oh0 * s - g.p, rows_in, -g.p, g.C,
                       in_scale, in_shift, in_relu);
    else
        dw_stage<false>(tile, plane_stride, ws, x, g0, g.planes, g.G, g.H, g.W, oh0 * s - g.p, rows_in, -g.p, g.C,
                        nullptr, nullptr, 0);
    for (int idx = threadIdx.x; idx < g.G * (taps + 1); idx += DW_THREADS) {
        const int gl = idx / (taps + 1), t = idx - gl * (taps + 1);
        const int64_t plane = g0 + gl;
        float v = 0.0f;
        if (plane < g.planes) {
            const int c = (int)(plane % g.C);
            v = (t < taps) ? __ldg(w + c * taps + t) : (bias ? __ldg(bias + c) : 0.0f);
        }
        wsm[idx] = v;
    }
    __syncthreads();

    const int gl = threadIdx.x / g.tpp, tl = threadIdx.x - gl * g.tpp;
    const int64_t plane = g0 + gl;
    if (plane >= g.planes) return;
    const float *wp = wsm + gl * (taps + 1);
    const float bv = wp[taps];
    // Vertical strips: a thread owns 4 consecutive output rows of ONE column, consecutive lanes own
    // consecutive columns -> conflict-free shared-memory reads and 128-byte coalesced stores.
    const int nstrips = cdiv(rows_out, 4) * g.OW;
    float *yp = y + plane * (int64_t)g.OH * g.OW;
    const float *tp = tile + gl * plane_stride;
    for (int st = tl; st < nstrips; st += g.tpp) {
        const int rg = st / g.OW, ow = st - rg * g.OW;
        const int r0 = rg << 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (KH && KW && S) {
            constexpr int PROWS = 3 * (S ? S : 1) + (KH ? KH : 1);
            float v[PROWS][KW ? KW : 1];
            const float *tcol = tp + (r0 * S) * ws + ow * S;
#pragma unroll
            for (int a = 0; a < PROWS; ++a)
#pragma unroll
                for (int j = 0; j < KW; ++j) v[a][j] = (r0 * S + a < rows_in) ? tcol[a * ws + j] : 0.0f;
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int i = 0; i < KH; ++i)
#pragma unroll
                    for (int j = 0; j < KW; ++j) acc[o] = fmaf(v[o * S + i][j], wp[i * KW + j], acc[o]);
        } else {
            for (int o = 0; o < 4; ++o) {
                if (r0 + o >= rows_out) break;
                const float *tcol = tp + ((r0 + o) * s) * ws + ow * s;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) acc[o] = fmaf(tcol[i * ws + j], wp[i * kw + j], acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (r0 + o < rows_out) yp[(int64_t)(oh0 + r0 + o) * g.OW + ow] = acc[o] + bv;
    }
}

// --------------------------------------------------------------------------------------- backward
// Sum `v` over the tpp threads of a plane group (fixed order).  Result valid in the group's thread 0.
__device__ __forceinline__ float dw_group_sum(float v, int tpp, float *red /* [DW_THREADS/32] */) {
    const int lim = tpp < 32 ? tpp : 32;
    for (int o = lim >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tpp > 32) {
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) red[wid] = v;
        __syncthreads();
        const int wpg = tpp >> 5;  // warps per group
        const int gl = threadIdx.x / tpp;
        if ((threadIdx.x - gl * tpp) == 0) {
            float t = 0.0f;
            for (int k = 0; k < wpg; ++k) t += red[gl * wpg + k];
            v = t;
        }
    }
    return v;
}

template <int KH, int KW, int S>
__global__ void __launch_bounds__(DW_THREADS)
dw_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ w,
              float *__restrict__ dx, float *__restrict__ partial, const float *__restrict__ in_scale,
              const float *__restrict__ in_shift, int in_relu, const float *__restrict__ dx_add, DwGeom g) {
    extern __shared__ float smem[];
    __shared__ float red[DW_THREADS / 32];
    const int kh = KH ? KH : g.kh, kw = KW ? KW : g.kw, s = S ? S : g.s;
    const int taps = kh * kw;
    const int band = blockIdx.y;
    const int64_t g0 = (int64_t)blockIdx.x * g.G;
    const DwBand bd = dw_band(g, band);
    const int dy_rows = bd.dy_hi - bd.dy_lo, x_rows = bd.x_hi - bd.x_lo;
    const int dplane = g.b_dy_rows_max * g.dws, xplane = g.b_x_rows_max * g.xws;
    float *dys = smem;                         // [G][dy_rows_max][dws]
    float *xs = dys + g.G * dplane;            // [G][x_rows_max][xws]
    float *wsm = xs + g.G * xplane;            // [G][taps]

    dw_stage<false>(dys, dplane, g.dws, dy, g0, g.planes, g.G, g.OH, g.OW, bd.dy_lo, dy_rows, -g.PL, g.C,
                    nullptr, nullptr, 0);
    if (in_scale)
        dw_stage<true>(xs, xplane, g.xws, x, g0, g.planes, g.G, g.H, g.W, bd.x_lo, x_rows, -g.p, g.C, in_scale,
                       in_shift, in_relu);
    else
        dw_stage<false>(xs, xplane, g.xws, x, g0, g.planes, g.G, g.H, g.W, bd.x_lo, x_rows, -g.p, g.C, nullptr,
                        nullptr, 0);
    for (int idx = threadIdx.x; idx < g.G * taps; idx += DW_THREADS) {
        const int gl = idx / taps, t = idx - gl * taps;
        const int64_t plane = g0 + gl;
        wsm[idx] = plane < g.planes ? __ldg(w + (int)(plane % g.C) * taps + t) : 0.0f;
    }
    __syncthreads();

    const int gl = threadIdx.x / g.tpp, tl = threadIdx.x - gl * g.tpp;
    const int64_t plane = g0 + gl;
    const bool live = plane < g.planes;
    const float *wp = wsm + gl * taps;
    const float *dyt = dys + gl * dplane;
    const float *xt = xs + gl * xplane;

    // ---- dX: gather form of the reference's scatter (im2col.pyx:175-177) -------------------------
    // vertical strips again: 4 consecutive input rows of one column per thread.
    if (live) {
        const int rows = bd.h1 - bd.h0;
        const int nstrips = cdiv(rows, 4) * g.W;
        float *dxp = dx + plane * (int64_t)g.H * g.W;
        const float *addp = dx_add ? dx_add + plane * (int64_t)g.H * g.W : nullptr;
        for (int st = tl; st < nstrips; st += g.tpp) {
            const int rg = st / g.W, wc = st - rg * g.W;
            const int hb = bd.h0 + (rg << 2);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (KH && KW && S == 1) {
                // dY rows hb+p-(KH-1) .. hb+p+3, cols wc+p-(KW-1) .. wc+p  (all inside the padded tile)
                constexpr int PROWS = 3 + (KH ? KH : 1);
                float v[PROWS][KW ? KW : 1];
                const int row_top = hb + g.p - (KH - 1) - bd.dy_lo;
                const float *dcol = dyt + row_top * g.dws + g.PL + wc + g.p - (KW - 1);
#pragma unroll
                for (int a = 0; a < PROWS; ++a)
#pragma unroll
                    for (int j = 0; j < KW; ++j) v[a][j] = (row_top + a < dy_rows) ? dcol[a * g.dws + j] : 0.0f;
                // dX[h] = sum_{i,j} dY[h+p-i][w+p-j] w[i][j];  patch row a = o + (KH-1-i), patch col = KW-1-j
#pragma unroll
                for (int o = 0; o < 4; ++o)
#pragma unroll
                    for (int i = 0; i < KH; ++i)
#pragma unroll
                        for (int j = 0; j < KW; ++j)
                            acc[o] = fmaf(v[o + KH - 1 - i][KW - 1 - j], wp[i * KW + j], acc[o]);
            } else {
                for (int o = 0; o < 4; ++o) {
                    const int h = hb + o;
                    if (h >= bd.h1) break;
                    for (int i = 0; i < kh; ++i) {
                        const int ti = h + g.p - i;
                        if ((ti % s) != 0) continue;
                        const float *drow = dyt + (ti / s - bd.dy_lo) * g.dws + g.PL;
                        for (int j = 0; j < kw; ++j) {
                            const int tj = wc + g.p - j;
                            if ((tj % s) == 0) acc[o] = fmaf(drow[tj / s], wp[i * kw + j], acc[o]);
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                if (hb + o < bd.h1) {
                    const int64_t off = (int64_t)(hb + o) * g.W + wc;
                    dxp[off] = acc[o] + (addp ? addp[off] : 0.0f);
                }
            }
        }
    }

    // ---- dW / db partials over the owned output rows ------------------------------------------------
    const int orows = bd.oo1 - bd.oo0;
    const int nstrips = cdiv(orows, 4) * g.OW;
    float *pout = partial + (plane * g.bands + band) * (int64_t)(taps + 1);
    if (KH && KW && S) {
        float aw[(KH ? KH : 1) * (KW ? KW : 1)];
#pragma unroll
        for (int t = 0; t < KH * KW; ++t) aw[t] = 0.0f;
        float ab = 0.0f;
        if (live) {
            for (int st = tl; st < nstrips; st += g.tpp) {
                const int rg = st / g.OW, ow = st - rg * g.OW;
                const int ob = bd.oo0 + (rg << 2);
                constexpr int PROWS = 3 * (S ? S : 1) + (KH ? KH : 1);
                float v[PROWS][KW ? KW : 1];
                const int xr0 = ob * S - g.p - bd.x_lo;
                const float *xcol = xt + xr0 * g.xws + ow * S;
#pragma unroll
                for (int a = 0; a < PROWS; ++a)
#pragma unroll
                    for (int j = 0; j < KW; ++j) v[a][j] = (xr0 + a < x_rows) ? xcol[a * g.xws + j] : 0.0f;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float gv = (ob + o < bd.oo1) ? dyt[(ob + o - bd.dy_lo) * g.dws + g.PL + ow] : 0.0f;
                    ab += gv;
#pragma unroll
                    for (int i = 0; i < KH; ++i)
#pragma unroll
                        for (int j = 0; j < KW; ++j) aw[i * KW + j] = fmaf(gv, v[o * S + i][j], aw[i * KW + j]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < KH * KW; ++t) {
            const float v = dw_group_sum(aw[t], g.tpp, red);
            if (live && tl == 0) pout[t] = v;
        }
        const float vb = dw_group_sum(ab, g.tpp, red);
        if (live && tl == 0) pout[KH * KW] = vb;
    } else {
        // generic filter size: one tap at a time (re-reads shared memory; rare path)
        for (int t = 0; t <= taps; ++t) {
            const int i = t / kw, j = t - i * kw;
            float a = 0.0f;
            if (live) {
                for (int st = tl; st < nstrips; st += g.tpp) {
                    const int rg = st / g.OW, ow = st - rg * g.OW;
                    for (int o = 0; o < 4; ++o) {
                        const int oh = bd.oo0 + (rg << 2) + o;
                        if (oh >= bd.oo1) break;
                        const float gv = dyt[(oh - bd.dy_lo) * g.dws + g.PL + ow];
                        a += (t < taps) ? gv * xt[(oh * s + i - g.p - bd.x_lo) * g.xws + ow * s + j] : gv;
                    }
                }
            }
            const float v = dw_group_sum(a, g.tpp, red);
            if (live && tl == 0) pout[t] = v;
        }
    }
}

// dw[c][t] = sum_{n,band} partial + l2*w ; dbias[c] = sum of the extra slot
__global__ void __launch_bounds__(128)
dw_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ dw,
                 float *__restrict__ dbias, float l2, int N, int C, int bands, int taps) {
    __shared__ float red[33];
    const int c = blockIdx.x;
    const int T = taps + 1;
    const int nb = N * bands;
    for (int t = 0; t < T; ++t) {
        float s = 0.0f;
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            const int n = k / bands, b = k - n * bands;
            s += partial[(((int64_t)n * C + c) * bands + b) * T + t];
        }
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            if (t < taps) dw[c * taps + t] = s + (l2 != 0.0f ? l2 * w[c * taps + t] : 0.0f);
            else if (dbias) dbias[c] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host
static int pow2_floor(int v) {
    int r = 1;
    while (r * 2 <= v) r *= 2;
    return r;
}

static size_t dw_fwd_smem(const DwGeom &g) {
    return ((size_t)g.G * g.f_rows_in_max * g.f_ws + (size_t)g.G * (g.kh * g.kw + 1)) * sizeof(float);
}
static size_t dw_bwd_smem(const DwGeom &g) {
    return ((size_t)g.G * (g.b_dy_rows_max * g.dws + g.b_x_rows_max * g.xws) + (size_t)g.G * g.kh * g.kw) * sizeof(float);
}

static void dw_bwd_tiles(DwGeom &g) {
    g.b_dy_rows_max = 0;
    g.b_x_rows_max = 0;
    for (int b = 0; b < g.bands; ++b) {
        const DwBand bd = dw_band(g, b);
        if (bd.dy_hi - bd.dy_lo > g.b_dy_rows_max) g.b_dy_rows_max = bd.dy_hi - bd.dy_lo;
        if (bd.x_hi - bd.x_lo > g.b_x_rows_max) g.b_x_rows_max = bd.x_hi - bd.x_lo;
    }
}

static int dw_geom(DwGeom &g, int N, int C, int H, int W, int kh, int kw, int s, int p, bool backward,
                   const char *who) {
    DK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && kh > 0 && kw > 0 && s > 0 && p >= 0, "%s: bad shape", who);
    DK_REQUIRE(kh * kw <= DW_MAX_TAPS && kh <= 8 && kw <= 8, "%s: filter %dx%d too large (max 8x8 / 49 taps)", who, kh, kw);
    DK_REQUIRE(H + 2 * p >= kh && W + 2 * p >= kw, "%s: filter larger than padded input", who);
    g.N = N; g.C = C; g.H = H; g.W = W; g.kh = kh; g.kw = kw; g.s = s; g.p = p;
    g.OH = (H + 2 * p - kh) / s + 1;
    g.OW = (W + 2 * p - kw) / s + 1;
    g.planes = (int64_t)N * C;
    g.f_ws = W + 2 * p;
    g.xws = W + 2 * p;
    g.PL = cdiv((kw - 1 - p) > 0 ? (kw - 1 - p) : 0, s);
    const int pr = floor_div(W - 1 + p, s) - (g.OW - 1);
    g.PR = pr > 0 ? pr : 0;
    g.dws = g.PL + g.OW + g.PR;
    const int budget = backward ? DW_BWD_TILE_BUDGET : DW_FWD_TILE_BUDGET;
    // whole plane first
    g.bands = 1;
    g.G = 1;
    g.f_rows_in_max = (g.OH - 1) * s + kh;
    dw_bwd_tiles(g);
    size_t plane_bytes = backward ? dw_bwd_smem(g) : dw_fwd_smem(g);
    if ((int64_t)plane_bytes <= budget) {
        int G = pow2_floor((int)(budget / plane_bytes));
        if (G > 32) G = 32;
        g.G = G;
    } else {
        const int max_bands = backward ? (g.OH < g.H ? g.OH : g.H) : g.OH;
        int bands = 2;
        for (; bands <= max_bands; ++bands) {
            g.bands = bands;
            g.f_rows_in_max = (cdiv(g.OH, bands) - 1) * s + kh;
            dw_bwd_tiles(g);
            if ((int64_t)(backward ? dw_bwd_smem(g) : dw_fwd_smem(g)) <= budget) break;
        }
        DK_REQUIRE(bands <= max_bands, "%s: image rows too wide for the shared-memory tile (W=%d)", who, W);
    }
    g.tpp = DW_THREADS / g.G;
    return DK_OK;
}

// register-window fast path for 3x3 / stride 1 / pad 1 (depthwise_rows.cu)
size_t dw_rows_ws_bytes(int N, int C, int H, int W, int kh, int kw, int s, int p);
size_t dw_rows_fwd_bn_ws_bytes(int N, int C, int H, int W, int kh, int kw, int s, int p);
int dw_rows_fwd_bn(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                   int s, int p, const BnFinalize &fin, float *trunc_resid, void *ws, size_t ws_bytes, cudaStream_t st);
int dw_rows_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                int s, int p, cudaStream_t st);
int dw_rows_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, void *ws, size_t ws_bytes,
                cudaStream_t st);
// planes-in-shared-memory path for 3x3 / pad 1 / stride 1 or 2 (depthwise_planes.cu)
size_t dw_planes_ws_bytes(int N, int C);
int init_dw_planes();
int dw_planes_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                  int s, int p, cudaStream_t st);
int dw_planes_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                  float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, void *ws, size_t ws_bytes,
                  cudaStream_t st);
int dw_chan_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st);
extern int g_dw_chan_enabled;
// depthwise_group.cu: channel-group kernels for 7x7 / 14x14 planes (3x3, stride 1, pad 1)
int init_dw_group();
int dw_group_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                 int s, int p, cudaStream_t st);
int dw_group_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                 float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, cudaStream_t st);
extern int g_dw_group_enabled;
static int g_dw_rows_enabled = 1;
// 0: never, 1: where measured faster than the register-window kernels (dw_use_planes), 2: wherever it applies
static int g_dw_planes_mode = 1;
extern int g_dwr_bwd_rb, g_dwr_bwd_vec_cap;

int init_depthwise() {
    // budgets are within the 48 KB default for forward; backward may slightly exceed with the filter stash
    int rc = init_dw_group();
    if (rc) return rc;
    return init_dw_planes();
}

static bool dw_use_planes(bool backward, int H, int W, int s) {
    if (g_dw_planes_mode == 0) return false;
    if (g_dw_planes_mode == 2) return true;
    (void)backward; (void)H; (void)W;
    return s == 2;  // TODO(measure): stride-1 shapes
}

template <int KH, int KW, int S>
static int dw_launch_fwd(const float *x, const float *w, const float *bias, float *y, const float *sc,
                         const float *sh, int relu, const DwGeom &g, cudaStream_t st) {
    const size_t smem = dw_fwd_smem(g);
    if (smem > 48 * 1024)
        DK_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<KH, KW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(g.planes, g.G), (unsigned)g.bands);
    dw_fwd_kernel<KH, KW, S><<<grid, DW_THREADS, smem, st>>>(x, w, bias, y, sc, sh, relu, g);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

template <int KH, int KW, int S>
static int dw_launch_bwd(const float *dy, const float *x, const float *w, float *dx, float *partial,
                         const float *sc, const float *sh, int relu, const float *dx_add, const DwGeom &g,
                         cudaStream_t st) {
    const size_t smem = dw_bwd_smem(g);
    if (smem > 48 * 1024)
        DK_CUDA(cudaFuncSetAttribute(dw_bwd_kernel<KH, KW, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(g.planes, g.G), (unsigned)g.bands);
    dw_bwd_kernel<KH, KW, S><<<grid, DW_THREADS, smem, st>>>(dy, x, w, dx, partial, sc, sh, relu, dx_add, g);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk

using namespace dk;

extern "C" {

size_t dk_dwconv_ws_bytes(int N, int C, int H, int W, int kh, int kw, int stride, int pad) {
    DwGeom g;
    if (dw_geom(g, N, C, H, W, kh, kw, stride, pad, true, "dk_dwconv_ws_bytes")) return 0;
    const size_t tiles = (size_t)g.planes * g.bands * (kh * kw + 1) * sizeof(float);
    const size_t rows = dw_rows_ws_bytes(N, C, H, W, kh, kw, stride, pad);
    const size_t planes = dw_planes_ws_bytes(N, C);
    const size_t m = tiles > rows ? tiles : rows;
    return m > planes ? m : planes;
}

int dk_dwconv_fwd(const float *x, const float *w, const float *bias, float *y, const float *in_scale,
                  const float *in_shift, int in_relu, int N, int C, int H, int W, int kh, int kw, int stride,
                  int pad, dk_stream_t stream) {
    DwGeom g;
    int rc = dw_geom(g, N, C, H, W, kh, kw, stride, pad, false, "dk_dwconv_fwd");
    if (rc) return rc;
    DK_REQUIRE(x && w && y, "dk_dwconv_fwd: NULL pointer");
    DK_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "dk_dwconv_fwd: in_scale/in_shift must come together");
    cudaStream_t st = as_stream(stream);
    if (in_scale == nullptr) {
        rc = dw_group_fwd(x, w, bias, y, N, C, H, W, kh, kw, stride, pad, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    if (in_scale == nullptr && dw_use_planes(false, H, W, stride)) {
        rc = dw_planes_fwd(x, w, bias, y, N, C, H, W, kh, kw, stride, pad, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    if (g_dw_rows_enabled && in_scale == nullptr) {
        rc = dw_rows_fwd(x, w, bias, y, N, C, H, W, kh, kw, stride, pad, st);
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    if (kh == 3 && kw == 3 && stride == 1) return dw_launch_fwd<3, 3, 1>(x, w, bias, y, in_scale, in_shift, in_relu, g, st);
    if (kh == 3 && kw == 3 && stride == 2) return dw_launch_fwd<3, 3, 2>(x, w, bias, y, in_scale, in_shift, in_relu, g, st);
    return dw_launch_fwd<0, 0, 0>(x, w, bias, y, in_scale, in_shift, in_relu, g, st);
}

size_t dk_dwconv_fwd_bn_ws_bytes(int N, int C, int H, int W, int kh, int kw, int stride, int pad) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || !g_dw_rows_enabled) return 0;
    return dw_rows_fwd_bn_ws_bytes(N, C, H, W, kh, kw, stride, pad);
}

int dk_dwconv_fwd_bn(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                     int stride, int pad, const float *gamma, const float *beta, float *running_mean, float *running_std,
                     int first_batch, float momentum, float eps, float *save_mean, float *save_invstd, float *save_scale,
                     float *save_shift, float *trunc_resid, void *ws, size_t ws_bytes, dk_stream_t stream) {
    DwGeom g;
    int rc = dw_geom(g, N, C, H, W, kh, kw, stride, pad, false, "dk_dwconv_fwd_bn");
    if (rc) return rc;
    DK_REQUIRE(x && w && y && gamma && beta && save_mean && save_invstd && save_scale && save_shift, "dk_dwconv_fwd_bn: NULL pointer");
    DK_REQUIRE((running_mean == nullptr) == (running_std == nullptr), "dk_dwconv_fwd_bn: running stats must come in pairs");
    const size_t need = dk_dwconv_fwd_bn_ws_bytes(N, C, H, W, kh, kw, stride, pad);
    DK_REQUIRE(need > 0, "dk_dwconv_fwd_bn: shape not covered (3x3, stride 1, pad 1, W %% 4 == 0, H*W >= 512): "
                         "call dk_dwconv_fwd + dk_bn_fwd_train instead (dk_dwconv_fwd_bn_ws_bytes returns 0 for it)");
    DK_REQUIRE(ws != nullptr && ws_bytes >= need, "dk_dwconv_fwd_bn: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    BnFinalize fin = {};
    fin.mode = 1;
    fin.gamma = gamma; fin.beta = beta;
    fin.running_mean = running_mean; fin.running_std = running_std;
    fin.first_batch = first_batch; fin.momentum = momentum; fin.eps = eps;
    fin.save_mean = save_mean; fin.save_invstd = save_invstd; fin.save_scale = save_scale; fin.save_shift = save_shift;
    rc = dw_rows_fwd_bn(x, w, bias, y, N, C, H, W, kh, kw, stride, pad, fin, trunc_resid, ws, ws_bytes, as_stream(stream));
    if (rc == DK_ERR_UNSUPPORTED) {
        set_error("dk_dwconv_fwd_bn: x / y must be 16-byte aligned");
        return DK_ERR_INVALID;
    }
    return rc;
}

int dk_dwconv_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias,
                  const float *in_scale, const float *in_shift, int in_relu, const float *dx_add, float l2, int N,
                  int C, int H, int W, int kh, int kw, int stride, int pad, void *ws, size_t ws_bytes,
                  dk_stream_t stream) {
    DwGeom g;
    int rc = dw_geom(g, N, C, H, W, kh, kw, stride, pad, true, "dk_dwconv_bwd");
    if (rc) return rc;
    DK_REQUIRE(dy && x && w && dx && dw, "dk_dwconv_bwd: NULL pointer");
    DK_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "dk_dwconv_bwd: in_scale/in_shift must come together");
    if (in_scale == nullptr) {
        rc = dw_group_bwd(dy, x, w, dx, dw, dbias, dx_add, l2, N, C, H, W, kh, kw, stride, pad, as_stream(stream));
        if (rc != DK_ERR_UNSUPPORTED) return rc;
        rc = dw_chan_bwd(dy, x, w, dx, dw, dbias, dx_add, l2, N, C, H, W, kh, kw, stride, pad, as_stream(stream));
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    if (in_scale == nullptr && dw_use_planes(true, H, W, stride)) {
        rc = dw_planes_bwd(dy, x, w, dx, dw, dbias, dx_add, l2, N, C, H, W, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream));
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    if (g_dw_rows_enabled && in_scale == nullptr) {
        rc = dw_rows_bwd(dy, x, w, dx, dw, dbias, dx_add, l2, N, C, H, W, kh, kw, stride, pad, ws, ws_bytes, as_stream(stream));
        if (rc != DK_ERR_UNSUPPORTED) return rc;
    }
    const size_t need = (size_t)g.planes * g.bands * (kh * kw + 1) * sizeof(float);
    if (ws == nullptr || ws_bytes < need) {
        set_error("dk_dwconv_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return DK_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float *partial = reinterpret_cast<float *>(ws);
    if (kh == 3 && kw == 3 && stride == 1) rc = dw_launch_bwd<3, 3, 1>(dy, x, w, dx, partial, in_scale, in_shift, in_relu, dx_add, g, st);
    else if (kh == 3 && kw == 3 && stride == 2) rc = dw_launch_bwd<3, 3, 2>(dy, x, w, dx, partial, in_scale, in_shift, in_relu, dx_add, g, st);
    else rc = dw_launch_bwd<0, 0, 0>(dy, x, w, dx, partial, in_scale, in_shift, in_relu, dx_add, g, st);
    if (rc) return rc;
    dw_reduce_kernel<<<C, 128, 0, st>>>(partial, w, dw, dbias, l2, N, C, g.bands, kh * kw);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

/* test hook: 0 = always use the shared-memory tile kernels, 1 = default dispatch, 2 = planes-in-smem kernels wherever
   they apply, 3 = register-window kernels without the planes kernels (2 and 3 also switch the per-channel backward
   kernel off), 4 = default dispatch without the per-channel backward kernel, 5 = default dispatch without the channel-group kernels
   of depthwise_group.cu (only mode 1 uses them) */
int dk_dw_debug_set(int enable_rows) {
    if (enable_rows < 10) g_dw_group_enabled = enable_rows == 1 ? 1 : 0;  // 5 = default dispatch without the channel-group kernels
    if (enable_rows == 2 || enable_rows == 3) {
        g_dw_rows_enabled = 1;
        g_dw_planes_mode = enable_rows == 2 ? 2 : 0;
        g_dw_chan_enabled = 0;
    } else if (enable_rows == 4) {
        g_dw_rows_enabled = 1;
        g_dw_planes_mode = 1;
        g_dw_chan_enabled = 0;
    } else if (enable_rows >= 20) g_dwr_bwd_vec_cap = enable_rows - 20;  // 21 / 22 / 24
    else if (enable_rows >= 10) g_dwr_bwd_rb = enable_rows - 10;  // 12 / 14: rows per iteration of the backward kernel
    else {
        g_dw_rows_enabled = enable_rows;
        g_dw_planes_mode = enable_rows ? 1 : 0;
        g_dw_chan_enabled = enable_rows ? 1 : 0;
    }
    return DK_OK;
}

}  // extern "C"
