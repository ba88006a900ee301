// Kernel 3, fp32 mode -- set-abstraction MLP (gather + relative-position concat + Lin/BN/ReLU x2 + Lin
// + max aggregation) and its backward on CUDA cores with fp32 FMA.
//
// This is the 1e-4 parity mode of the path (SURVEY.md 7.2 "fp32 mode"); the bf16 tcgen05 kernels in
// sa_tc.cu share its data layout.  Reference call sites: /root/reference/pointnet2_regressor.py:18
// (PointConv), :29-30 (GlobalSAModule), arithmetic per SURVEY.md A.3-A.6.
//
// Layout: an SA level works on ROWS.  SLOTS mode: row = m*K + k is neighbour slot k of centroid m
// (valid iff k < cnt[m]); CLOUDS mode: row = source point, segment = batch[row].  Three layer GEMMs
// run over row tiles of 128; the concat [x_j || pos_j - pos_i] is formed in the A-tile loader and the
// E x c3 output of the last layer is reduced to the per-target max in the epilogue, so neither ever
// reaches HBM.  Only the pre-BN activations h1, h2 are stored (train-mode BatchNorm needs batch
// statistics over all rows before the next layer can start; they are also what backward re-reads).
#include <math.h>

#include "common.cuh"

namespace b2pn {
namespace simt {

constexpr int BM = 128;  // rows per tile
constexpr int BK = 16;   // reduction chunk
constexpr int NT = 256;  // threads per CTA

struct RowMap {
    int seg_mode;
    int K;
    const int32_t *nbr;
    const int32_t *cnt;
    const int64_t *batch;
    int64_t rows;
    __device__ __forceinline__ bool valid(int64_t row) const
    {
        if (row >= rows) return false;
        if (seg_mode) return true;
        const int64_t m = row / K;
        return (int)(row - m * K) < cnt[m];
    }
    __device__ __forceinline__ int64_t seg(int64_t row) const { return seg_mode ? batch[row] : row / K; }
    __device__ __forceinline__ int64_t src(int64_t row) const { return seg_mode ? row : (int64_t)nbr[row]; }
    // id stored in arg[]: slot (SLOTS) or source row (CLOUDS)
    __device__ __forceinline__ int id(int64_t row) const { return seg_mode ? (int)row : (int)(row % K); }
};

// ------------------------------------------------------------------------------------------------
//  A-tile loaders: value of column c of one row, zero outside the row's validity / channel range
// ------------------------------------------------------------------------------------------------
struct GatherLoader {  // [x_j || pos_j - pos_i]   (PointNetConv.message)
    RowMap rm;
    const float *x;
    int c_in;
    const float *pos_src;
    const float *pos_dst;
    bool ok;
    const float *xr;
    float d0, d1, d2;
    __device__ __forceinline__ void prepare(int64_t row)
    {
        ok = rm.valid(row);
        xr = nullptr;
        d0 = d1 = d2 = 0.f;
        if (ok) {
            const int64_t s = rm.src(row);
            xr = x ? x + s * c_in : nullptr;
            d0 = pos_src[3 * s + 0];
            d1 = pos_src[3 * s + 1];
            d2 = pos_src[3 * s + 2];
            if (!rm.seg_mode) {
                const int64_t m = row / rm.K;
                d0 = __fsub_rn(d0, pos_dst[3 * m + 0]);
                d1 = __fsub_rn(d1, pos_dst[3 * m + 1]);
                d2 = __fsub_rn(d2, pos_dst[3 * m + 2]);
            }
        }
    }
    __device__ __forceinline__ float load(int c) const
    {
        if (!ok) return 0.f;
        if (c < c_in) return xr[c];
        const int j = c - c_in;
        return j == 0 ? d0 : (j == 1 ? d1 : (j == 2 ? d2 : 0.f));
    }
};

struct BnActLoader {  // act(h * scale + shift): BatchNorm1d + activation applied on load
    RowMap rm;
    const float *h;
    int C;
    const float *scale;
    const float *shift;
    int act;
    bool ok;
    const float *hr;
    __device__ __forceinline__ void prepare(int64_t row)
    {
        ok = rm.valid(row);
        hr = h + row * C;
    }
    __device__ __forceinline__ float load(int c) const
    {
        if (!ok || c >= C) return 0.f;
        const float v = fmaf(hr[c], scale[c], shift[c]);
        return act == B2PN_ACT_RELU ? fmaxf(v, 0.f) : v;
    }
};

struct PlainLoader {
    RowMap rm;
    const float *a;
    int C;
    bool ok;
    const float *ar;
    __device__ __forceinline__ void prepare(int64_t row)
    {
        ok = rm.valid(row);
        ar = a + row * C;
    }
    __device__ __forceinline__ float load(int c) const { return (ok && c < C) ? ar[c] : 0.f; }
};

struct ArgGradLoader {  // gradient of the max aggregation: flows to the arg-max row only (A.5)
    RowMap rm;
    const float *dout;
    const int32_t *arg;
    int C;
    bool ok;
    int64_t base;
    int myid;
    __device__ __forceinline__ void prepare(int64_t row)
    {
        ok = row < rm.rows;
        base = 0;
        myid = -2;
        if (ok) {
            base = rm.seg(row) * C;
            myid = rm.id(row);
        }
    }
    __device__ __forceinline__ float load(int c) const
    {
        if (!ok || c >= C) return 0.f;
        return arg[base + c] == myid ? dout[base + c] : 0.f;
    }
};

template <class L>
struct WithOnes {  // appends a column of ones (valid rows) so the bias gradient falls out of the dW GEMM
    L l;
    int ones_col;
    __device__ __forceinline__ void prepare(int64_t row) { l.prepare(row); }
    __device__ __forceinline__ float load(int c) const { return c == ones_col ? (l.ok ? 1.f : 0.f) : l.load(c); }
};

// B operand: element (k, n) of a [Kdim x N] matrix stored row-major with leading dimension ld
struct BMat {
    const float *p;
    int ld;
    int Kdim, N;
    __device__ __forceinline__ float get(int k, int n) const { return (k < Kdim && n < N) ? __ldg(p + (int64_t)k * ld + n) : 0.f; }
};

// ------------------------------------------------------------------------------------------------
//  Epilogues
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Scratch {
    double rs[8][BN];
    double rq[8][BN];
};

// column sums of two quantities over the tile's rows (double, fixed order) -> partial[tile][2][N]
template <int BN, int TN>
__device__ __forceinline__ void tile_column_sums(double (&s)[TN], double (&q)[TN], Scratch<BN> &sc, double *partial,
                                                 int N, int n0, int tx)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
    }
    if (lane < 16) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            sc.rs[warp][tx * TN + j] = s[j];
            sc.rq[warp][tx * TN + j] = q[j];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < BN; c += NT) {
        double S = 0.0, Q = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            S += sc.rs[w][c];
            Q += sc.rq[w][c];
        }
        if (n0 + c < N) {
            double *pt = partial + (int64_t)blockIdx.x * 2 * N;
            pt[n0 + c] = S;
            pt[N + n0 + c] = Q;
        }
    }
}

struct StoreStatsEp {  // h = acc + bias -> store; per-column sum and sum of squares over valid rows
    RowMap rm;
    float *h;
    const float *bias;
    double *partial;  // [tiles][2][N]
    template <int BN, int TN>
    __device__ __forceinline__ void run(float (&acc)[8][TN], int64_t row0, int n0, int ty, int tx, int N, void *smem)
    {
        Scratch<BN> &sc = *reinterpret_cast<Scratch<BN> *>(smem);
        double s[TN], q[TN];
        float bj[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            s[j] = q[j] = 0.0;
            const int c = n0 + tx * TN + j;
            bj[j] = c < N ? bias[c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + ty * 8 + i;
            const bool ok = rm.valid(row);
            if (row < rm.rows) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int c = n0 + tx * TN + j;
                    const float v = ok ? acc[i][j] + bj[j] : 0.f;
                    if (c < N) h[row * N + c] = v;
                    if (ok) {
                        s[j] += (double)v;
                        q[j] += (double)v * (double)v;
                    }
                }
            }
        }
        tile_column_sums<BN, TN>(s, q, sc, partial, N, n0, tx);
    }
};

struct SlotMaxEp {  // out[m] = max over the valid slots of centroid m (first max wins ties), arg = slot
    RowMap rm;
    float *out;
    int32_t *arg;
    const float *bias;
    template <int BN, int TN>
    __device__ __forceinline__ void run(float (&acc)[8][TN], int64_t row0, int n0, int ty, int tx, int N, void *smem)
    {
        float(*sv)[BN] = reinterpret_cast<float(*)[BN]>(smem);                         // [16][BN]
        int(*sk)[BN] = reinterpret_cast<int(*)[BN]>(reinterpret_cast<float *>(smem) + 16 * BN);  // [16][BN]
        const int K = rm.K;
        const int64_t r0 = row0 + ty * 8;          // 8 rows of one centroid (K is a multiple of 8)
        const int64_t m = r0 / K;
        const int k0 = (int)(r0 - m * K);
        const int64_t n_dst = rm.rows / K;
        const int cm = m < n_dst ? rm.cnt[m] : 0;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = n0 + tx * TN + j;
            const float bj = c < N ? bias[c] : 0.f;
            float best = -INFINITY;
            int bk = -1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float v = acc[i][j] + bj;
                if (k0 + i < cm && (bk < 0 || v > best)) {
                    best = v;
                    bk = k0 + i;
                }
            }
            sv[ty][tx * TN + j] = best;
            sk[ty][tx * TN + j] = bk;
        }
        __syncthreads();
        const int groups = K / 8;            // ty-groups per centroid
        const int cens = BM / K;             // centroids per tile
        for (int t = threadIdx.x; t < cens * BN; t += NT) {
            const int cen = t / BN, col = t - cen * BN;
            const int64_t mm = row0 / K + cen;
            const int c = n0 + col;
            if (mm < n_dst && c < N) {
                float best = 0.f;
                int bk = -1;
                for (int g = 0; g < groups; ++g) {
                    const int kk = sk[cen * groups + g][col];
                    const float v = sv[cen * groups + g][col];
                    if (kk >= 0 && (bk < 0 || v > best)) {
                        best = v;
                        bk = kk;
                    }
                }
                out[mm * N + c] = best;   // targets without neighbours -> 0
                arg[mm * N + c] = bk;
            }
        }
    }
};

__device__ __forceinline__ unsigned f32_orderable(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_orderable(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct CloudMaxEp {  // global_max_pool: 64-bit atomicMax of (orderable value, ~row) -> first max row wins ties
    RowMap rm;
    unsigned long long *keys;  // [n_dst][N], zero-initialised
    const float *bias;
    template <int BN, int TN>
    __device__ __forceinline__ void run(float (&acc)[8][TN], int64_t row0, int n0, int ty, int tx, int N, void *smem)
    {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = n0 + tx * TN + j;
            if (c >= N) continue;
            const float bj = bias[c];
            int64_t curseg = -1;
            unsigned long long best = 0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int64_t row = row0 + ty * 8 + i;
                if (row >= rm.rows) break;
                const int64_t sg = rm.batch[row];
                if (sg != curseg) {
                    if (curseg >= 0) atomicMax(keys + curseg * N + c, best);
                    curseg = sg;
                    best = 0ull;
                }
                const unsigned long long key =
                    ((unsigned long long)f32_orderable(acc[i][j] + bj) << 32) | (unsigned long long)(0xffffffffu - (unsigned)row);
                best = key > best ? key : best;
            }
            if (curseg >= 0) atomicMax(keys + curseg * N + c, best);
        }
    }
};

__global__ void unpack_keys_kernel(const unsigned long long *keys, float *out, int32_t *arg, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    if (k == 0ull) {
        out[i] = 0.f;
        arg[i] = -1;
    } else {
        out[i] = f32_from_orderable((unsigned)(k >> 32));
        arg[i] = (int32_t)(0xffffffffu - (unsigned)(k & 0xffffffffu));
    }
}

struct MaskStoreSumsEp {  // backward through activation: dz = da * [z > 0]; sums for the BatchNorm backward
    RowMap rm;
    float *dz;
    const float *hprev;  // pre-BN activation of the layer being differentiated through
    const float *bn;     // [4][cmax]: mean, rstd, scale, shift
    int cmax;
    int act;
    double *partial;
    template <int BN, int TN>
    __device__ __forceinline__ void run(float (&acc)[8][TN], int64_t row0, int n0, int ty, int tx, int N, void *smem)
    {
        Scratch<BN> &sc = *reinterpret_cast<Scratch<BN> *>(smem);
        double s[TN], q[TN];
        float mean[TN], rstd[TN], scale[TN], shift[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            s[j] = q[j] = 0.0;
            const int c = n0 + tx * TN + j;
            const bool in = c < N;
            mean[j] = in ? bn[c] : 0.f;
            rstd[j] = in ? bn[cmax + c] : 0.f;
            scale[j] = in ? bn[2 * cmax + c] : 0.f;
            shift[j] = in ? bn[3 * cmax + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + ty * 8 + i;
            if (row >= rm.rows) continue;
            const bool ok = rm.valid(row);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = n0 + tx * TN + j;
                if (c >= N) continue;
                float g = 0.f;
                if (ok) {
                    const float hv = hprev[row * N + c];
                    const float z = fmaf(hv, scale[j], shift[j]);
                    g = (act == B2PN_ACT_RELU && !(z > 0.f)) ? 0.f : acc[i][j];
                    const float zhat = (hv - mean[j]) * rstd[j];
                    s[j] += (double)g;
                    q[j] += (double)g * (double)zhat;
                }
                dz[row * N + c] = g;
            }
        }
        tile_column_sums<BN, TN>(s, q, sc, partial, N, n0, tx);
    }
};

struct ScatterEp {  // gradient w.r.t. the gathered source features
    RowMap rm;
    float *dx;
    template <int BN, int TN>
    __device__ __forceinline__ void run(float (&acc)[8][TN], int64_t row0, int n0, int ty, int tx, int N, void *smem)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + ty * 8 + i;
            if (!rm.valid(row)) continue;
            const int64_t s = rm.src(row);
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int c = n0 + tx * TN + j;
                if (c >= N) continue;
                if (rm.seg_mode) dx[s * N + c] = acc[i][j];        // every source row appears exactly once
                else atomicAdd(dx + s * N + c, acc[i][j]);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
//  C[rows x N] = A[rows x Kdim] * B[Kdim x N], A produced by a loader, C consumed by an epilogue
// ------------------------------------------------------------------------------------------------
template <int BN, class AL, class EP>
__global__ void __launch_bounds__(NT) rows_gemm_kernel(AL al, BMat bm, EP ep)
{
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ __align__(16) unsigned char scratch[sizeof(Scratch<BN>)];
    static_assert(sizeof(Scratch<BN>) >= 2 * 16 * BN * 4, "scratch too small for SlotMaxEp");

    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int lr = tid & (BM - 1), lk = (tid >> 7) * 8;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    al.prepare(row0 + lr);
    for (int k0 = 0; k0 < bm.Kdim; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[lk + i][lr] = al.load(k0 + lk + i);
#pragma unroll
        for (int e = tid; e < BK * BN; e += NT) {
            const int kk = e / BN, n = e - kk * BN;
            Bs[kk][n] = bm.get(k0 + kk, n0 + n);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4 *>(&b[j]) = *reinterpret_cast<const float4 *>(&Bs[kk][tx * TN + j]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    ep.template run<BN, TN>(acc, row0, n0, ty, tx, bm.N, scratch);
}

// ------------------------------------------------------------------------------------------------
//  dW[n_out x k_in] partial = sum over a row range of Y[row, :]^T (x) A[row, :]
// ------------------------------------------------------------------------------------------------
template <class YL, class AL>
__global__ void __launch_bounds__(NT) dw_gemm_kernel(YL yl, AL al, int n_out, int k_in, int64_t rows, int64_t rows_per_split,
                                                     float *partial)
{
    __shared__ __align__(16) float Ys[BK][128];
    __shared__ __align__(16) float As[BK][128];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.z * 128;
    const int64_t rbeg = (int64_t)blockIdx.x * rows_per_split;
    const int64_t rend = min(rows, rbeg + rows_per_split);
    const int lkk = tid >> 4, lc = tid & 15;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int64_t r0 = rbeg; r0 < rend; r0 += BK) {
        const int64_t row = r0 + lkk;
        const bool in = row < rend;
        yl.prepare(in ? row : rows);  // rows -> invalid -> zeros
        al.prepare(in ? row : rows);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            Ys[lkk][lc + 16 * i] = yl.load(m0 + lc + 16 * i);
            As[lkk][lc + 16 * i] = al.load(n0 + lc + 16 * i);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[8];
            *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&Ys[kk][ty * 8]);
            *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&Ys[kk][ty * 8 + 4]);
            *reinterpret_cast<float4 *>(&b[0]) = *reinterpret_cast<const float4 *>(&As[kk][tx * 8]);
            *reinterpret_cast<float4 *>(&b[4]) = *reinterpret_cast<const float4 *>(&As[kk][tx * 8 + 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *pt = partial + (int64_t)blockIdx.x * n_out * k_in;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= n_out) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n < k_in) pt[(int64_t)m * k_in + n] = acc[i][j];
        }
    }
}

// dW = sum of the split partials in fixed order; last column of the partials is the bias gradient
__global__ void dw_reduce_kernel(const float *partial, int splits, int n_out, int k_in, float *grad_w, float *grad_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t tot = (int64_t)n_out * k_in;
    if (i >= tot) return;
    double s = 0.0;
    for (int p = 0; p < splits; ++p) s += (double)partial[(int64_t)p * tot + i];
    const int m = (int)(i / k_in), n = (int)(i - (int64_t)m * k_in);
    if (n == k_in - 1) {
        if (grad_b) grad_b[m] = (float)s;
    } else if (grad_w) {
        grad_w[(int64_t)m * (k_in - 1) + n] = (float)s;
    }
}

// ------------------------------------------------------------------------------------------------
//  small kernels: weight transpose, valid-row count, BatchNorm finalisation, BN-backward apply
// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const float *w, int n_out, int k_in, float *wt)  // wt[k][n] = w[n][k]
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out * k_in) return;
    const int n = i / k_in, k = i - n * k_in;
    wt[(int64_t)k * n_out + n] = w[i];
}

__global__ void set_double_kernel(double *p, double v) { *p = v; }

__global__ void count_valid_kernel(const int32_t *cnt, int64_t n, double *out)
{
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)cnt[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        *out = t;
    }
}

// block = 32 channels x 8 tile groups; reduces partial[tiles][2][C] in a fixed order
__device__ __forceinline__ void reduce_partials(const double *partial, int64_t tiles, int C, int c, int g, double &S, double &Q,
                                                double (*sm)[2][32])
{
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int64_t t = g; t < tiles; t += 8) {
            s += partial[t * 2 * C + c];
            q += partial[t * 2 * C + C + c];
        }
    }
    sm[g][0][threadIdx.x & 31] = s;
    sm[g][1][threadIdx.x & 31] = q;
    __syncthreads();
    S = Q = 0.0;
    for (int w = 0; w < 8; ++w) {
        S += sm[w][0][threadIdx.x & 31];
        Q += sm[w][1][threadIdx.x & 31];
    }
}

__global__ void __launch_bounds__(256) bn_fwd_finalize_kernel(const double *partial, int64_t tiles, int C, int cmax,
                                                              const double *count, int training, const float *gamma,
                                                              const float *beta, float *running_mean, float *running_var,
                                                              int64_t *nbt, float eps, float momentum, float *bn)
{
    __shared__ double sm[8][2][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    double mean, var;
    if (training) {
        double S, Q;
        reduce_partials(partial, tiles, C, c, g, S, Q, sm);
        const double E = *count;
        mean = E > 0 ? S / E : 0.0;
        var = E > 0 ? Q / E - mean * mean : 0.0;
        if (var < 0.0) var = 0.0;
        if (g == 0 && c < C) {
            const double unbiased = E > 1.0 ? var * E / (E - 1.0) : var;
            running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
            running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
            if (c == 0 && nbt) *nbt += 1;
        }
    } else {
        mean = c < C ? running_mean[c] : 0.0;
        var = c < C ? running_var[c] : 1.0;
    }
    if (g == 0 && c < C) {
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const double scale = (double)gamma[c] * rstd;
        bn[c] = (float)mean;
        bn[cmax + c] = (float)rstd;
        bn[2 * cmax + c] = (float)scale;
        bn[3 * cmax + c] = (float)((double)beta[c] - mean * scale);
    }
}

// sums S1 = sum dz, S2 = sum dz*zhat -> dbeta, dgamma and the per-channel means the BN backward needs
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const double *partial, int64_t tiles, int C, const double *count,
                                                              int training, float *grad_gamma, float *grad_beta, float *sbar)
{
    __shared__ double sm[8][2][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    double S, Q;
    reduce_partials(partial, tiles, C, c, g, S, Q, sm);
    if (g == 0 && c < C) {
        const double E = *count;
        if (grad_beta) grad_beta[c] = (float)S;
        if (grad_gamma) grad_gamma[c] = (float)Q;
        sbar[c] = (training && E > 0) ? (float)(S / E) : 0.f;
        sbar[C + c] = (training && E > 0) ? (float)(Q / E) : 0.f;
    }
}

// dh = scale * (dz - mean(dz) - zhat * mean(dz*zhat)), in place, zero on invalid rows
__global__ void bn_bwd_apply_kernel(RowMap rm, float *dz, const float *h, int C, const float *bn, int cmax, const float *sbar)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rm.rows * C) return;
    const int64_t row = i / C;
    const int c = (int)(i - row * C);
    float v = 0.f;
    if (rm.valid(row)) {
        const float zhat = (h[i] - bn[c]) * bn[cmax + c];
        v = bn[2 * cmax + c] * (dz[i] - sbar[c] - zhat * sbar[C + c]);
    }
    dz[i] = v;
}

// ------------------------------------------------------------------------------------------------
//  host orchestration
// ------------------------------------------------------------------------------------------------
static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Ws {  // bump allocator over the caller's workspace; identical walk in *_bytes and in the launchers
    char *base;
    int64_t off;
    explicit Ws(void *p) : base((char *)p), off(0) {}
    template <class T>
    T *take(int64_t n)
    {
        T *r = base ? (T *)(base + off) : nullptr;
        off = align_up(off + n * (int64_t)sizeof(T), 256);
        return r;
    }
};

struct Shapes {
    int64_t rows, tiles;
    int c0, c1, c2, c3, cmax;
    int splits;
};

static Shapes shapes_of(const b2pn_sa_args &a)
{
    Shapes s;
    s.rows = a.seg_mode == B2PN_SEG_CLOUDS ? a.n_src : a.n_dst * (int64_t)a.K;
    s.tiles = (s.rows + BM - 1) / BM;
    s.c0 = a.mlp.c[0];
    s.c1 = a.mlp.c[1];
    s.c2 = a.mlp.c[2];
    s.c3 = a.mlp.c[3];
    s.cmax = s.c1 > s.c2 ? s.c1 : s.c2;
    int64_t sp = s.rows / 2048;
    s.splits = (int)(sp < 1 ? 1 : (sp > 296 ? 296 : sp));
    return s;
}

struct FwdWs {
    float *wt[3];
    double *count;
    double *partial;
    unsigned long long *keys;
};
static FwdWs carve_fwd(const b2pn_sa_args &a, const Shapes &s, Ws &ws)
{
    FwdWs f;
    f.wt[0] = ws.take<float>((int64_t)s.c0 * s.c1);
    f.wt[1] = ws.take<float>((int64_t)s.c1 * s.c2);
    f.wt[2] = ws.take<float>((int64_t)s.c2 * s.c3);
    f.count = ws.take<double>(1);
    f.partial = ws.take<double>(s.tiles * 2 * s.cmax);
    f.keys = a.seg_mode == B2PN_SEG_CLOUDS ? ws.take<unsigned long long>(a.n_dst * (int64_t)s.c3) : nullptr;
    return f;
}

struct BwdWs {
    double *count;
    double *partial;
    float *dz1, *dz2;
    float *sbar;  // [2][cmax]
    float *dwp;   // [splits][max n_out][max k_in+1]
};
static BwdWs carve_bwd(const b2pn_sa_args &a, const Shapes &s, Ws &ws)
{
    BwdWs b;
    b.count = ws.take<double>(1);
    b.partial = ws.take<double>(s.tiles * 2 * s.cmax);
    b.dz1 = ws.take<float>(s.rows * s.c1);
    b.dz2 = ws.take<float>(s.rows * s.c2);
    b.sbar = ws.take<float>(2 * s.cmax);
    int64_t m1 = (int64_t)s.c1 * (s.c0 + 1), m2 = (int64_t)s.c2 * (s.c1 + 1), m3 = (int64_t)s.c3 * (s.c2 + 1);
    int64_t mx = m1 > m2 ? m1 : m2;
    mx = mx > m3 ? mx : m3;
    b.dwp = ws.take<float>(mx * s.splits);
    return b;
}

static RowMap rowmap_of(const b2pn_sa_args &a, const Shapes &s)
{
    RowMap rm;
    rm.seg_mode = a.seg_mode;
    rm.K = a.seg_mode == B2PN_SEG_CLOUDS ? 1 : a.K;
    rm.nbr = a.nbr;
    rm.cnt = a.cnt;
    rm.batch = a.batch;
    rm.rows = s.rows;
    return rm;
}

template <class AL, class EP>
static void launch_rows_gemm(const AL &al, const BMat &bm, const EP &ep, int64_t tiles, cudaStream_t st)
{
    if (bm.N <= 64) {
        dim3 grid((unsigned)tiles, (unsigned)((bm.N + 63) / 64));
        rows_gemm_kernel<64, AL, EP><<<grid, NT, 0, st>>>(al, bm, ep);
        note_launch();
    } else {
        dim3 grid((unsigned)tiles, (unsigned)((bm.N + 127) / 128));
        rows_gemm_kernel<128, AL, EP><<<grid, NT, 0, st>>>(al, bm, ep);
        note_launch();
    }
}

template <class YL, class AL>
static void launch_dw(const YL &yl, const AL &al, int n_out, int k_in_plus1, const Shapes &s, float *dwp, float *gw, float *gb,
                      cudaStream_t st)
{
    const int64_t rps = align_up((s.rows + s.splits - 1) / s.splits, BK);
    dim3 grid((unsigned)s.splits, (unsigned)((n_out + 127) / 128), (unsigned)((k_in_plus1 + 127) / 128));
    dw_gemm_kernel<YL, AL><<<grid, NT, 0, st>>>(yl, al, n_out, k_in_plus1, s.rows, rps, dwp);
    note_launch();
    const int64_t tot = (int64_t)n_out * k_in_plus1;
    dw_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dwp, s.splits, n_out, k_in_plus1, gw, gb);
    note_launch();
}

static int check_args(const b2pn_sa_args &a)
{
    if (a.n_src < 0 || a.n_dst < 0 || a.c_in < 0) return B2PN_EINVAL;
    if (a.mlp.c[0] != a.c_in + 3 || a.mlp.c[1] <= 0 || a.mlp.c[2] <= 0 || a.mlp.c[3] <= 0) return B2PN_EINVAL;
    if (a.mlp.act != B2PN_ACT_NONE && a.mlp.act != B2PN_ACT_RELU) return B2PN_ENOTSUP;
    if (a.seg_mode == B2PN_SEG_SLOTS) {
        if (a.K < 8 || a.K > 128 || (a.K % 8) != 0 || (128 % a.K) != 0) return B2PN_ENOTSUP;
        if (a.n_dst > 0 && (!a.nbr || !a.cnt || !a.pos_dst)) return B2PN_EINVAL;
    } else if (a.seg_mode == B2PN_SEG_CLOUDS) {
        if (a.n_src > 0 && !a.batch) return B2PN_EINVAL;
    } else {
        return B2PN_EINVAL;
    }
    if (a.n_src > 0 && (!a.pos_src || (a.c_in > 0 && !a.x))) return B2PN_EINVAL;
    for (int l = 0; l < 3; ++l)
        if (!a.mlp.w[l] || !a.mlp.b[l]) return B2PN_EINVAL;
    for (int l = 0; l < 2; ++l)
        if (!a.mlp.gamma[l] || !a.mlp.beta[l] || !a.mlp.running_mean[l] || !a.mlp.running_var[l]) return B2PN_EINVAL;
    if (!a.out || !a.arg || !a.h1 || !a.h2 || !a.bn) return B2PN_EINVAL;
    return B2PN_OK;
}

int64_t sa_workspace_bytes_f32(const b2pn_sa_args &a, int backward)
{
    const Shapes s = shapes_of(a);
    Ws ws(nullptr);
    if (backward) carve_bwd(a, s, ws);
    else carve_fwd(a, s, ws);
    return ws.off + 256;
}

int sa_forward_f32(const b2pn_sa_args &a, cudaStream_t st)
{
    int rc = check_args(a);
    if (rc) return rc;
    const Shapes s = shapes_of(a);
    if (a.n_dst == 0) return B2PN_OK;
    if (a.workspace_bytes < sa_workspace_bytes_f32(a, 0) || !a.workspace) return B2PN_EINVAL;
    Ws ws(a.workspace);
    const FwdWs f = carve_fwd(a, s, ws);
    const RowMap rm = rowmap_of(a, s);
    const int cdim[4] = {s.c0, s.c1, s.c2, s.c3};
    float *h1 = (float *)a.h1, *h2 = (float *)a.h2;

    for (int l = 0; l < 3; ++l) {
        const int n = cdim[l] * cdim[l + 1];
        transpose_kernel<<<(n + 255) / 256, 256, 0, st>>>(a.mlp.w[l], cdim[l + 1], cdim[l], f.wt[l]);
        note_launch();
    }
    if (a.seg_mode == B2PN_SEG_SLOTS) {
        count_valid_kernel<<<1, 1024, 0, st>>>(a.cnt, a.n_dst, f.count);
        note_launch();
    } else {
        set_double_kernel<<<1, 1, 0, st>>>(f.count, (double)s.rows);
        note_launch();
    }
    if (s.rows > 0) {
        // layer 1: gather + concat + Linear
        GatherLoader gl = {rm, (const float *)a.x, a.c_in, a.pos_src, a.pos_dst};
        StoreStatsEp e1 = {rm, h1, a.mlp.b[0], f.partial};
        launch_rows_gemm(gl, BMat{f.wt[0], s.c1, s.c0, s.c1}, e1, s.tiles, st);
    }
    bn_fwd_finalize_kernel<<<(s.c1 + 31) / 32, 256, 0, st>>>(f.partial, s.tiles, s.c1, s.cmax, f.count, a.training, a.mlp.gamma[0],
                                                             a.mlp.beta[0], a.mlp.running_mean[0], a.mlp.running_var[0],
                                                             a.mlp.num_batches_tracked[0], a.mlp.eps, a.mlp.momentum, a.bn);
    note_launch();
    float *bn1 = a.bn, *bn2 = a.bn + 4 * s.cmax;
    if (s.rows > 0) {
        BnActLoader l2 = {rm, h1, s.c1, bn1 + 2 * s.cmax, bn1 + 3 * s.cmax, a.mlp.act};
        StoreStatsEp e2 = {rm, h2, a.mlp.b[1], f.partial};
        launch_rows_gemm(l2, BMat{f.wt[1], s.c2, s.c1, s.c2}, e2, s.tiles, st);
    }
    bn_fwd_finalize_kernel<<<(s.c2 + 31) / 32, 256, 0, st>>>(f.partial, s.tiles, s.c2, s.cmax, f.count, a.training, a.mlp.gamma[1],
                                                             a.mlp.beta[1], a.mlp.running_mean[1], a.mlp.running_var[1],
                                                             a.mlp.num_batches_tracked[1], a.mlp.eps, a.mlp.momentum, bn2);
    note_launch();
    BnActLoader l3 = {rm, h2, s.c2, bn2 + 2 * s.cmax, bn2 + 3 * s.cmax, a.mlp.act};
    if (a.seg_mode == B2PN_SEG_SLOTS) {
        SlotMaxEp e3 = {rm, a.out, a.arg, a.mlp.b[2]};
        launch_rows_gemm(l3, BMat{f.wt[2], s.c3, s.c2, s.c3}, e3, s.tiles, st);
    } else {
        const int64_t n = a.n_dst * (int64_t)s.c3;
        B2PN_CUDA(cudaMemsetAsync(f.keys, 0, n * sizeof(unsigned long long), st));
        if (s.rows > 0) {
            CloudMaxEp e3 = {rm, f.keys, a.mlp.b[2]};
            launch_rows_gemm(l3, BMat{f.wt[2], s.c3, s.c2, s.c3}, e3, s.tiles, st);
        }
        unpack_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(f.keys, a.out, a.arg, n);
        note_launch();
    }
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}

int sa_backward_f32(const b2pn_sa_args &a, const b2pn_sa_grads &g, cudaStream_t st)
{
    int rc = check_args(a);
    if (rc) return rc;
    if (!g.grad_out) return B2PN_EINVAL;
    const Shapes s = shapes_of(a);
    if (a.n_dst == 0 || s.rows == 0) return B2PN_OK;
    if (a.workspace_bytes < sa_workspace_bytes_f32(a, 1) || !a.workspace) return B2PN_EINVAL;
    Ws ws(a.workspace);
    const BwdWs b = carve_bwd(a, s, ws);
    const RowMap rm = rowmap_of(a, s);
    const float *h1 = (const float *)a.h1, *h2 = (const float *)a.h2;
    float *bn1 = a.bn, *bn2 = a.bn + 4 * s.cmax;

    if (a.seg_mode == B2PN_SEG_SLOTS) {
        count_valid_kernel<<<1, 1024, 0, st>>>(a.cnt, a.n_dst, b.count);
        note_launch();
    } else {
        set_double_kernel<<<1, 1, 0, st>>>(b.count, (double)s.rows);
        note_launch();
    }

    // ---- layer 3: dh3 is the routed output gradient ------------------------------------------------
    ArgGradLoader y3 = {rm, g.grad_out, a.arg, s.c3};
    BnActLoader a2 = {rm, h2, s.c2, bn2 + 2 * s.cmax, bn2 + 3 * s.cmax, a.mlp.act};
    {
        MaskStoreSumsEp ep = {rm, b.dz2, h2, bn2, s.cmax, a.mlp.act, b.partial};
        launch_rows_gemm(y3, BMat{a.mlp.w[2], s.c2, s.c3, s.c2}, ep, s.tiles, st);   // da2 = dh3 * W3
        WithOnes<BnActLoader> a2o = {a2, s.c2};
        launch_dw(y3, a2o, s.c3, s.c2 + 1, s, b.dwp, g.grad_w[2], g.grad_b[2], st);  // dW3 = dh3^T a2
    }
    bn_bwd_finalize_kernel<<<(s.c2 + 31) / 32, 256, 0, st>>>(b.partial, s.tiles, s.c2, b.count, a.training, g.grad_gamma[1],
                                                             g.grad_beta[1], b.sbar);
    note_launch();
    {
        const int64_t n = s.rows * s.c2;
        bn_bwd_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rm, b.dz2, h2, s.c2, bn2, s.cmax, b.sbar);
        note_launch();
    }
    // ---- layer 2 -------------------------------------------------------------------------------------
    PlainLoader y2 = {rm, b.dz2, s.c2};
    BnActLoader a1 = {rm, h1, s.c1, bn1 + 2 * s.cmax, bn1 + 3 * s.cmax, a.mlp.act};
    {
        MaskStoreSumsEp ep = {rm, b.dz1, h1, bn1, s.cmax, a.mlp.act, b.partial};
        launch_rows_gemm(y2, BMat{a.mlp.w[1], s.c1, s.c2, s.c1}, ep, s.tiles, st);
        WithOnes<BnActLoader> a1o = {a1, s.c1};
        launch_dw(y2, a1o, s.c2, s.c1 + 1, s, b.dwp, g.grad_w[1], g.grad_b[1], st);
    }
    bn_bwd_finalize_kernel<<<(s.c1 + 31) / 32, 256, 0, st>>>(b.partial, s.tiles, s.c1, b.count, a.training, g.grad_gamma[0],
                                                             g.grad_beta[0], b.sbar);
    note_launch();
    {
        const int64_t n = s.rows * s.c1;
        bn_bwd_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rm, b.dz1, h1, s.c1, bn1, s.cmax, b.sbar);
        note_launch();
    }
    // ---- layer 1 -------------------------------------------------------------------------------------
    PlainLoader y1 = {rm, b.dz1, s.c1};
    {
        GatherLoader gl = {rm, (const float *)a.x, a.c_in, a.pos_src, a.pos_dst};
        WithOnes<GatherLoader> glo = {gl, s.c0};
        launch_dw(y1, glo, s.c1, s.c0 + 1, s, b.dwp, g.grad_w[0], g.grad_b[0], st);
    }
    if (g.grad_x && a.c_in > 0) {
        ScatterEp ep = {rm, g.grad_x};
        launch_rows_gemm(y1, BMat{a.mlp.w[0], s.c0, s.c1, a.c_in}, ep, s.tiles, st);  // only the feature columns
    }
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}

}  // namespace simt
}  // namespace b2pn
