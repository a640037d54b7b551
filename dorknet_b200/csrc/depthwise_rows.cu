// depthwise_rows.cu -- register-window depthwise 3x3 / stride 1 / pad 1 kernels (the shape of every depthwise
// layer but three in ResNet-18-depsep and of most of the MobileNet stack).
//
// The shared-memory tile kernels in depthwise.cu are instruction-bound (ncu: 77 % issue-active at 12 % of HBM
// peak): they spend ~20 instructions per element on staging and index arithmetic.  Here a thread owns a strip of
// VEC adjacent columns and walks down the rows of its plane keeping a 3-row sliding window in registers: per row it
// issues ONE vector load per input tensor, gets the two halo columns from its neighbours with warp shuffles, and
// does 9 FMAs per output (forward) or 18 (backward: dX and the dW accumulators share the windows).  Four rows are
// loaded ahead of the arithmetic to keep enough bytes in flight.  Consecutive threads own consecutive strips of
// consecutive planes, so global accesses are contiguous runs of whole rows.
//
// Backward is one pass over dY and X (im2col.pyx:143-178 does the same fusion on the CPU): it writes dX and
// per-(plane, band) partial sums of dW / db -- reduced deterministically by dw_reduce_kernel (depthwise.cu), as the
// reference sums its per-image dW (depthwise_convolution.py:193).
#include "bn.cuh"
#include "common.cuh"

namespace dk {

constexpr int DWR_THREADS = 128;
constexpr int DWR_RB = 4;  // rows per loop iteration (loads in flight per thread and per tensor)

template <int VEC> struct VecT;
template <> struct VecT<4> { using T = float4; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<1> { using T = float; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float *p, float (&v)[VEC]) {
    if (VEC == 4) {
        const float4 q = ld_stream4(p);
        v[0] = q.x; v[1] = q.y; v[VEC > 2 ? 2 : 0] = q.z; v[VEC > 3 ? 3 : 0] = q.w;
    } else if (VEC == 2) {
        const float2 q = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = q.x; v[VEC > 1 ? 1 : 0] = q.y;
    } else {
        v[0] = __ldg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float *p, const float (&v)[VEC]) {
    if (VEC == 4) st_stream4(p, make_float4(v[0], v[1], v[VEC > 2 ? 2 : 0], v[VEC > 3 ? 3 : 0]));
    else if (VEC == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[VEC > 1 ? 1 : 0]);
    else *p = v[0];
}

// One input row of a strip with its halo columns: r[0] = left neighbour, r[1..VEC] = own columns, r[VEC+1] = right.
// Loading is split in two so that the global loads of the NEXT row group are in flight while the current one is
// being computed: issue_row() only issues the loads (own columns + the two halo scalars that cannot come from a
// neighbouring lane), finish_row() does the shuffles.  All 32 lanes execute the shuffles; `ok` masks the loads.
template <int VEC>
struct RawRow {
    float v[VEC];
    float hl, hr;  // halo values for lane 0 / lane 31 (neighbour strip lives in another warp)
    bool in;
};

template <int VEC>
__device__ __forceinline__ void issue_row(const float *__restrict__ src, bool ok, int row, int H, int W, int strip,
                                          int SP, int lane, RawRow<VEC> &q) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) q.v[i] = 0.0f;
    q.hl = 0.0f;
    q.hr = 0.0f;
    q.in = ok && row >= 0 && row < H;
    const float *p = src + (long long)row * W + strip * VEC;
    if (q.in) {
        ld_vec<VEC>(p, q.v);
        if (lane == 0 && strip > 0) q.hl = __ldg(p - 1);
        if (lane == 31 && strip < SP - 1) q.hr = __ldg(p + VEC);
    }
}

template <int VEC>
__device__ __forceinline__ void finish_row(const RawRow<VEC> &q, int strip, int SP, int lane, float (&r)[VEC + 2]) {
    float left = __shfl_up_sync(0xffffffffu, q.v[VEC - 1], 1);
    float right = __shfl_down_sync(0xffffffffu, q.v[0], 1);
    if (lane == 0) left = q.hl;
    if (lane == 31) right = q.hr;
    if (strip == 0 || !q.in) left = 0.0f;
    if (strip == SP - 1 || !q.in) right = 0.0f;
    r[0] = left;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r[1 + i] = q.v[i];
    r[VEC + 1] = right;
}

template <int VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ src, bool ok, int row, int H, int W, int strip,
                                         int SP, int lane, float (&r)[VEC + 2]) {
    RawRow<VEC> q;
    issue_row<VEC>(src, ok, row, H, W, strip, SP, lane, q);
    finish_row<VEC>(q, strip, SP, lane, r);
}

// ---------------------------------------------------------------------------------------------- forward
// STATS: the kernel also leaves, per (plane, band) segment and warp piece, the sums BatchNorm needs of the values it just
// produced -- sum y, sum y^2 and sum (y - tf32_truncate(y)) -- so that a BatchNorm that follows (and is folded into the
// pointwise GEMM after it, bn_fold.cu) costs no pass over the activation at all (dw_stats_finalize_kernel turns them into
// mean / invstd / scale / shift).  Same fixed-order segmented reduction as the backward kernel's dW sums.
template <int VEC, bool STATS>
__global__ void __launch_bounds__(DWR_THREADS, 5)  // (5 CTAs per SM as without the statistics: 102 registers)
dw3x3_rows_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                      float *__restrict__ y, long long planes, int C, int H, int W, int bands,
                      float *__restrict__ stat_partial) {
    const int SP = W / VEC;
    const long long gid = (long long)blockIdx.x * DWR_THREADS + threadIdx.x;
    const long long total = planes * bands * SP;
    const bool ok = gid < total;
    const long long g = ok ? gid : total - 1;
    const int strip = (int)(g % SP);
    const long long t = g / SP;
    const int band = (int)(t % bands);
    const long long plane = t / bands;
    const int c = (int)(plane % C);
    const int lane = threadIdx.x & 31;
    const int rows = H / bands;  // host guarantees H % bands == 0
    const int h0 = band * rows;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
    const float bv = bias ? __ldg(bias + c) : 0.0f;
    const float *xp = x + plane * (long long)H * W;
    float *yp = y + plane * (long long)H * W;

    float st_sum = 0.0f, st_sq = 0.0f, st_res = 0.0f;
    float win[DWR_RB + 2][VEC + 2];
    load_row<VEC>(xp, ok, h0 - 1, H, W, strip, SP, lane, win[0]);
    load_row<VEC>(xp, ok, h0, H, W, strip, SP, lane, win[1]);
    RawRow<VEC> nxt[DWR_RB];
#pragma unroll
    for (int r = 0; r < DWR_RB; ++r) issue_row<VEC>(xp, ok && (r < rows), h0 + r + 1, H, W, strip, SP, lane, nxt[r]);
    for (int hb = 0; hb < rows; hb += DWR_RB) {
#pragma unroll
        for (int r = 0; r < DWR_RB; ++r) finish_row<VEC>(nxt[r], strip, SP, lane, win[2 + r]);
        // loads of the next row group fly while this one is computed
#pragma unroll
        for (int r = 0; r < DWR_RB; ++r)
            issue_row<VEC>(xp, ok && (hb + DWR_RB + r < rows), h0 + hb + DWR_RB + r + 1, H, W, strip, SP, lane, nxt[r]);
#pragma unroll
        for (int r = 0; r < DWR_RB; ++r) {
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float a = bv;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) a = fmaf(win[r + i][v + j], k[i * 3 + j], a);
                o[v] = a;
            }
            if (ok && hb + r < rows) {
                st_vec<VEC>(yp + (long long)(h0 + hb + r) * W + strip * VEC, o);
                if (STATS) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        st_sum += o[v];
                        st_sq = fmaf(o[v], o[v], st_sq);
                        st_res += o[v] - __uint_as_float(__float_as_uint(o[v]) & 0xFFFFE000u);  // what kind::tf32 drops
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC + 2; ++v) {
            win[0][v] = win[DWR_RB][v];
            win[1][v] = win[DWR_RB + 1][v];
        }
    }
    if (STATS) {
        float acc[3] = {st_sum, st_sq, st_res};
        const long long key = ok ? t : -1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float v = ok ? acc[i] : 0.0f;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {  // segmented inclusive scan by (plane, band), towards higher lanes
                const float up = __shfl_up_sync(0xffffffffu, v, d);
                const long long kup = __shfl_up_sync(0xffffffffu, key, d);
                if (lane >= d && kup == key) v += up;
            }
            acc[i] = v;
        }
        const long long knext = __shfl_down_sync(0xffffffffu, key, 1);
        if (ok && (lane == 31 || knext != key)) {
            const long long first_gid = t * SP;
            const int piece = (int)((gid >> 5) - (first_gid >> 5));
            const int pieces_max = (SP + 30) / 32 + 1;
            float *dst = stat_partial + ((t * pieces_max) + piece) * 3;
            dst[0] = acc[0];
            dst[1] = acc[1];
            dst[2] = acc[2];
        }
    }
}

// One CTA per channel: adds the (image, band, piece) partial sums of dw3x3_rows_fwd_kernel<.., true> in a fixed order, in
// double (sum y^2 - (sum y)^2 / n cancels), and finalises the BatchNorm statistics exactly like the BatchNorm kernels do
// (bn.cuh: batch_norm.py:66-89); trunc_resid[c] = mean of y - tf32_truncate(y) (bn_fold.cu uses it to keep the folded GEMM's
// output mean exact).
__global__ void __launch_bounds__(128)
dw_stats_finalize_kernel(const float *__restrict__ partial, int N, int C, int bands, int pieces_max, int SP, long long count,
                         BnFinalize fin, float *__restrict__ trunc_resid) {
    __shared__ double red[3][4];
    const int c = blockIdx.x;
    const int per_plane = bands * pieces_max;
    const int total = N * per_plane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int n = k / per_plane, q = k - n * per_plane;
        const int band = q / pieces_max, piece = q - band * pieces_max;
        const long long seg = ((long long)n * C + c) * bands + band;
        const long long first = seg * SP, last = first + SP - 1;
        const int npieces = (int)((last >> 5) - (first >> 5)) + 1;
        if (piece < npieces) {
            const float *src = partial + (seg * pieces_max + piece) * 3;
            a0 += (double)src[0];
            a1 += (double)src[1];
            a2 += (double)src[2];
        }
    }
    // fixed-order tree: xor-shuffles inside each warp (every lane ends with the warp's sum), then the four warp sums in order
    // (the first version had thread 0 add 3 x 128 doubles out of shared memory one after the other: 6 us of an 8.8 us kernel)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = a0;
        red[1][threadIdx.x >> 5] = a1;
        red[2][threadIdx.x >> 5] = a2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a0 = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
        a1 = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
        a2 = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
        const double mean = a0 / (double)count;
        double var = a1 / (double)count - mean * mean;  // biased (batch_norm_stats_cy.pyx:44)
        if (var < 0.0) var = 0.0;
        float sc, sh;
        bn_finalize_channel(fin, c, (float)mean, (float)var, true, &sc, &sh);
        if (trunc_resid) trunc_resid[c] = (float)(a2 / (double)count);
    }
}

// --------------------------------------------------------------------------------------------- backward
// dX[h][w] = sum_{i,j} dY[h+1-i][w+1-j] w[i][j];  dW[i][j] += dY[h][w] X[h+i-1][w+j-1];  db += dY[h][w]
template <int VEC, int RB>
__global__ void __launch_bounds__(DWR_THREADS)
dw3x3_rows_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ x, const float *__restrict__ w,
                      float *__restrict__ dx, float *__restrict__ partial, const float *__restrict__ dx_add,
                      long long planes, int C, int H, int W, int bands) {
    const int SP = W / VEC;
    const long long gid = (long long)blockIdx.x * DWR_THREADS + threadIdx.x;
    const long long total = planes * bands * SP;
    const bool ok = gid < total;
    const long long g = ok ? gid : total - 1;
    const int strip = (int)(g % SP);
    const long long t = g / SP;
    const int band = (int)(t % bands);
    const long long plane = t / bands;
    const int c = (int)(plane % C);
    const int lane = threadIdx.x & 31;
    const int rows = H / bands;
    const int h0 = band * rows;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = __ldg(w + c * 9 + i);
    const long long pbase = plane * (long long)H * W;
    const float *gp = dy + pbase, *xp = x + pbase;
    float *dxp = dx + pbase;
    const float *ap = dx_add ? dx_add + pbase : nullptr;

    float acc[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) acc[i] = 0.0f;
    float gw[RB + 2][VEC + 2], xw[RB + 2][VEC + 2];
    load_row<VEC>(gp, ok, h0 - 1, H, W, strip, SP, lane, gw[0]);
    load_row<VEC>(gp, ok, h0, H, W, strip, SP, lane, gw[1]);
    load_row<VEC>(xp, ok, h0 - 1, H, W, strip, SP, lane, xw[0]);
    load_row<VEC>(xp, ok, h0, H, W, strip, SP, lane, xw[1]);
    RawRow<VEC> ng[RB], nx[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        issue_row<VEC>(gp, ok && (r < rows), h0 + r + 1, H, W, strip, SP, lane, ng[r]);
        issue_row<VEC>(xp, ok && (r < rows), h0 + r + 1, H, W, strip, SP, lane, nx[r]);
    }
    for (int hb = 0; hb < rows; hb += RB) {
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            finish_row<VEC>(ng[r], strip, SP, lane, gw[2 + r]);
            finish_row<VEC>(nx[r], strip, SP, lane, xw[2 + r]);
        }
        // loads of the next row group fly while this one is computed
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const bool more = ok && (hb + RB + r < rows);
            issue_row<VEC>(gp, more, h0 + hb + RB + r + 1, H, W, strip, SP, lane, ng[r]);
            issue_row<VEC>(xp, more, h0 + hb + RB + r + 1, H, W, strip, SP, lane, nx[r]);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const bool live = ok && hb + r < rows;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                // window row r+1 is the centre row h; dY row (h + 1 - i) is window row r + 2 - i, column v + 2 - j
                float a = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) a = fmaf(gw[r + 2 - i][v + 2 - j], k[i * 3 + j], a);
                o[v] = a;
                const float gv = live ? gw[r + 1][v + 1] : 0.0f;
                acc[9] += gv;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc[i * 3 + j] = fmaf(gv, xw[r + i][v + j], acc[i * 3 + j]);
            }
            if (live) {
                const long long off = (long long)(h0 + hb + r) * W + strip * VEC;
                if (ap) {
                    float av[VEC];
                    ld_vec<VEC>(ap + off, av);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) o[v] += av[v];
                }
                st_vec<VEC>(dxp + off, o);
            }
        }
#pragma unroll
        for (int v = 0; v < VEC + 2; ++v) {
            gw[0][v] = gw[RB][v]; gw[1][v] = gw[RB + 1][v];
            xw[0][v] = xw[RB][v]; xw[1][v] = xw[RB + 1][v];
        }
    }
    // Reduce the 10 accumulators over the strips of this (plane, band).  The strips of one (plane, band) are
    // consecutive threads; a segmented shuffle scan in fixed order keeps the result deterministic.
    // key = (plane, band) id of this thread
    const long long key = ok ? t : -1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        float v = ok ? acc[i] : 0.0f;
        // segmented inclusive scan (by key) towards higher lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float up = __shfl_up_sync(0xffffffffu, v, d);
            const long long kup = __shfl_up_sync(0xffffffffu, key, d);
            if (lane >= d && kup == key) v += up;
        }
        acc[i] = v;
    }
    // the last lane of each segment holds the segment total; segments crossing a warp boundary combine with atomics-free
    // two-slot writes: every segment piece writes its own partial row (piece index = 0 for the piece that starts the
    // segment in this warp ... ) -- simpler: each warp-piece adds into a per-piece slot indexed by the warp-local order.
    const long long knext = __shfl_down_sync(0xffffffffu, key, 1);
    const bool seg_end = ok && (lane == 31 || knext != key);
    if (seg_end) {
        // piece id: number of warp boundaries between the first strip of the segment and this lane's strip
        const long long first_gid = t * SP;                       // gid of strip 0 of this (plane, band)
        const int piece = (int)((gid >> 5) - (first_gid >> 5));    // 0 .. pieces-1
        const int pieces_max = (SP + 30) / 32 + 1;
        float *dst = partial + ((t * pieces_max) + piece) * 10;
#pragma unroll
        for (int i = 0; i < 10; ++i) dst[i] = acc[i];
    }
}

// dw[c][t] = sum over (n, band, piece) partial + l2*w ; dbias[c].  Only the pieces a (plane, band) segment really
// has are read (a segment of SP consecutive threads starting at thread seg*SP spans a known number of warps), so the
// workspace needs no clearing.  One pass: every thread accumulates all 10 sums, then 10 block reductions.
__global__ void __launch_bounds__(128)
dw_rows_reduce_kernel(const float *__restrict__ partial, const float *__restrict__ w, float *__restrict__ dw,
                      float *__restrict__ dbias, float l2, int N, int C, int bands, int pieces_max, int SP) {
    __shared__ float red[33];
    const int c = blockIdx.x;
    const int per_plane = bands * pieces_max;
    const int total = N * per_plane;
    float acc[10];
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[t] = 0.0f;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int n = k / per_plane, q = k - n * per_plane;
        const int band = q / pieces_max, piece = q - band * pieces_max;
        const long long seg = ((long long)n * C + c) * bands + band;  // (plane, band) id
        const long long first = seg * SP, last = first + SP - 1;
        const int npieces = (int)((last >> 5) - (first >> 5)) + 1;
        if (piece < npieces) {
            const float *src = partial + (seg * pieces_max + piece) * 10;
#pragma unroll
            for (int t = 0; t < 10; ++t) acc[t] += src[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 10; ++t) {
        const float s = block_sum(acc[t], red);
        if (threadIdx.x == 0) {
            if (t < 9) dw[c * 9 + t] = s + (l2 != 0.0f ? l2 * w[c * 9 + t] : 0.0f);
            else if (dbias) dbias[c] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host
int g_dwr_bwd_vec_cap = 4;  // test knob: cap the vector width of the backward kernel
int g_dwr_bwd_rb = 2;  // rows per iteration of the vec-4 backward kernel (2: 4 CTAs/SM, 4: 2 CTAs/SM)

struct DwRowsPlan {
    int vec, bands, sp, pieces_max;
    long long planes, total;
};

// enough threads to cover the machine ~2x; bands must divide H and keep >= 2*DWR_RB rows each
static int dw_rows_bands(long long planes, int sp, int H) {
    int bands = 1;
    const long long want = (long long)sm_count() * 2048 * 2;
    while (planes * sp * bands < want && bands * 2 <= H / (2 * DWR_RB) && H % (bands * 2) == 0) bands *= 2;
    return bands;
}

static bool dw_rows_plan(DwRowsPlan &pl, const void *a, const void *b, const void *c, const void *d, int N, int C, int H,
                         int W, int kh, int kw, int s, int p) {
    if (kh != 3 || kw != 3 || s != 1 || p != 1 || H < 1 || W < 2) return false;
    int vec = 1;
    const bool al16 = aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d);
    if (W % 4 == 0 && al16) vec = 4;
    else if (W % 2 == 0 && al16) vec = 2;
    pl.vec = vec;
    pl.sp = W / vec;
    pl.planes = (long long)N * C;
    pl.bands = dw_rows_bands(pl.planes, pl.sp, H);
    pl.total = pl.planes * pl.bands * pl.sp;
    pl.pieces_max = (pl.sp + 30) / 32 + 1;
    return true;
}

size_t dw_rows_ws_bytes(int N, int C, int H, int W, int kh, int kw, int s, int p) {
    if (kh != 3 || kw != 3 || s != 1 || p != 1 || H < 1 || W < 2) return 0;
    // the vector width depends on pointer alignment, unknown here: size for the worst of the three variants
    long long worst = 0;
    for (int vec = 1; vec <= 4; vec *= 2) {
        if (W % vec) continue;
        const int sp = W / vec;
        const int pieces = (sp + 30) / 32 + 1;
        const long long n = (long long)N * C * dw_rows_bands((long long)N * C, sp, H) * pieces * 10;
        if (n > worst) worst = n;
    }
    return (size_t)worst * sizeof(float);
}

int dw_rows_fwd(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                int s, int p, cudaStream_t st) {
    DwRowsPlan pl;
    if (!dw_rows_plan(pl, x, y, x, y, N, C, H, W, kh, kw, s, p)) return DK_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)ceil_div(pl.total, DWR_THREADS);
    if (pl.vec == 4) dw3x3_rows_fwd_kernel<4, false><<<grid, DWR_THREADS, 0, st>>>(x, w, bias, y, pl.planes, C, H, W, pl.bands, nullptr);
    else if (pl.vec == 2) dw3x3_rows_fwd_kernel<2, false><<<grid, DWR_THREADS, 0, st>>>(x, w, bias, y, pl.planes, C, H, W, pl.bands, nullptr);
    else dw3x3_rows_fwd_kernel<1, false><<<grid, DWR_THREADS, 0, st>>>(x, w, bias, y, pl.planes, C, H, W, pl.bands, nullptr);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

// forward + the BatchNorm statistics of its output (two launches; the second is C small CTAs)
size_t dw_rows_fwd_bn_ws_bytes(int N, int C, int H, int W, int kh, int kw, int s, int p) {
    // (planes of fewer than 512 pixels belong to the channel-group kernels of depthwise_group.cu)
    if (kh != 3 || kw != 3 || s != 1 || p != 1 || W % 4 != 0 || H * W < 512) return 0;
    const int sp = W / 4;
    return (size_t)N * C * dw_rows_bands((long long)N * C, sp, H) * ((sp + 30) / 32 + 1) * 3 * sizeof(float);
}
int dw_rows_fwd_bn(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int kh, int kw,
                   int s, int p, const BnFinalize &fin, float *trunc_resid, void *ws, size_t ws_bytes, cudaStream_t st) {
    DwRowsPlan pl;
    if (!dw_rows_plan(pl, x, y, x, y, N, C, H, W, kh, kw, s, p) || pl.vec != 4) return DK_ERR_UNSUPPORTED;
    const size_t need = (size_t)pl.planes * pl.bands * pl.pieces_max * 3 * sizeof(float);
    if (ws == nullptr || ws_bytes < need || (long long)N * pl.bands * pl.pieces_max >= (1ll << 31)) return DK_ERR_UNSUPPORTED;
    float *partial = reinterpret_cast<float *>(ws);
    const unsigned grid = (unsigned)ceil_div(pl.total, DWR_THREADS);
    dw3x3_rows_fwd_kernel<4, true><<<grid, DWR_THREADS, 0, st>>>(x, w, bias, y, pl.planes, C, H, W, pl.bands, partial);
    DK_LAUNCH_CHECK();
    dw_stats_finalize_kernel<<<C, 128, 0, st>>>(partial, N, C, pl.bands, pl.pieces_max, pl.sp, (long long)N * H * W, fin, trunc_resid);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

int dw_rows_bwd(const float *dy, const float *x, const float *w, float *dx, float *dw, float *dbias, const float *dx_add,
                float l2, int N, int C, int H, int W, int kh, int kw, int s, int p, void *ws, size_t ws_bytes,
                cudaStream_t st) {
    DwRowsPlan pl;
    if (!dw_rows_plan(pl, dy, x, dx, dx_add, N, C, H, W, kh, kw, s, p)) return DK_ERR_UNSUPPORTED;
    if (pl.vec > g_dwr_bwd_vec_cap) {
        pl.vec = g_dwr_bwd_vec_cap;
        pl.sp = W / pl.vec;
        pl.bands = dw_rows_bands(pl.planes, pl.sp, H);
        pl.total = pl.planes * pl.bands * pl.sp;
        pl.pieces_max = (pl.sp + 30) / 32 + 1;
    }
    const int per_plane = pl.bands * pl.pieces_max;
    const long long nfloats = pl.planes * per_plane * 10;
    if (ws == nullptr || ws_bytes < (size_t)nfloats * sizeof(float)) return DK_ERR_UNSUPPORTED;  // caller falls back
    float *partial = reinterpret_cast<float *>(ws);
    const unsigned grid = (unsigned)ceil_div(pl.total, DWR_THREADS);
    if (pl.vec == 4) {
        if (g_dwr_bwd_rb == 4) dw3x3_rows_bwd_kernel<4, 4><<<grid, DWR_THREADS, 0, st>>>(dy, x, w, dx, partial, dx_add, pl.planes, C, H, W, pl.bands);
        else dw3x3_rows_bwd_kernel<4, 2><<<grid, DWR_THREADS, 0, st>>>(dy, x, w, dx, partial, dx_add, pl.planes, C, H, W, pl.bands);
    } else if (pl.vec == 2) dw3x3_rows_bwd_kernel<2, 4><<<grid, DWR_THREADS, 0, st>>>(dy, x, w, dx, partial, dx_add, pl.planes, C, H, W, pl.bands);
    else dw3x3_rows_bwd_kernel<1, 4><<<grid, DWR_THREADS, 0, st>>>(dy, x, w, dx, partial, dx_add, pl.planes, C, H, W, pl.bands);
    DK_LAUNCH_CHECK();
    dw_rows_reduce_kernel<<<C, 128, 0, st>>>(partial, w, dw, dbias, l2, N, C, pl.bands, pl.pieces_max, pl.sp);
    DK_LAUNCH_CHECK();
    return DK_OK;
}

}  // namespace dk
