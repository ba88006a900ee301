// Kernel 2 -- segmented radius ball query with a neighbour cap, fixed-width slots.
//
// Replaces torch_cluster's radius kernel reached from /root/reference/pointnet2_regressor.py:14-15
// (+ the masked_select compaction and its host sync).  Semantics: for every
// query scan the sources of the same cloud in ascending index, keep the first K with
// d2 < r*r (strict), d2 = ((dx*dx+dy*dy)+dz*dz) in separately rounded fp32, dx = src - qry.
//
// B200 design: a CTA owns QPB consecutive queries of one cloud.  The cloud's positions stream
// through shared memory in SoA tiles filled with coalesced float4 loads of the raw xyz stream;
// a warp tests 32 sources per step against TWO queries at once (packed FADD2/FMUL2), turns the
// hits into ordered slots with ballot + popc, and stops scanning a query once it has K hits.
// Output is nbr[M,K] int32 (global source index, -1 pad) + cnt[M]: no data-dependent shapes.
#include "common.cuh"

namespace b2pn {

struct BqParams {
    const float *src;
    const float *qry;
    const int64_t *src_ptr;
    const int64_t *qry_ptr;
    int32_t *nbr;
    int32_t *cnt;
    float r2;
    int K;
};

template <int THREADS, int QPW, int TILE>
__global__ void __launch_bounds__(THREADS) ball_query_kernel(const BqParams p)
{
    static_assert(QPW % 2 == 0 && TILE % 32 == 0, "");
    constexpr int NW = THREADS / 32;
    constexpr int QPB = NW * QPW;
    __shared__ float sx[TILE], sy[TILE], sz[TILE];

    const int b = blockIdx.y;
    const int64_t s0 = p.src_ptr[b];
    const int n = (int)(p.src_ptr[b + 1] - s0);
    const int64_t q0 = p.qry_ptr[b];
    const int mq = (int)(p.qry_ptr[b + 1] - q0);
    const int qbase = blockIdx.x * QPB;
    if (qbase >= mq) return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int K = p.K;
    const float r2 = p.r2;

    // my warp's queries: local ids qw .. qw+QPW-1 (may run past mq -> inactive)
    const int qw = qbase + warp * QPW;
    int c[QPW];
    u64 QX[QPW / 2], QY[QPW / 2], QZ[QPW / 2];
#pragma unroll
    for (int h = 0; h < QPW / 2; ++h) {
        float x[2], y[2], z[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int ql = qw + 2 * h + e;
            const bool ok = ql < mq;
            const float *q = p.qry + 3 * (q0 + (ok ? ql : 0));
            x[e] = __ldg(q + 0);
            y[e] = __ldg(q + 1);
            z[e] = __ldg(q + 2);
            c[2 * h + e] = ok ? 0 : K;  // inactive queries count as full
        }
        QX[h] = pack2(x[0], x[1]);
        QY[h] = pack2(y[0], y[1]);
        QZ[h] = pack2(z[0], z[1]);
    }

    const float *gsrc = p.src + 3 * s0;  // raw xyz stream of this cloud (4-byte aligned only)
    for (int t0 = 0; t0 < n; t0 += TILE) {
        const int tn = min(TILE, n - t0);
        // ---- stage tile: coalesced loads of the raw stream, float4 where 16B-aligned -------------
        {
            const float *g = gsrc + 3 * (int64_t)t0;
            const int nf = 3 * tn;
            const int head = min(nf, (int)(((16u - ((unsigned)(uintptr_t)g & 15u)) & 15u) >> 2));
            const int nvec = (nf - head) >> 2;
            const int tail0 = head + (nvec << 2);
            auto put = [&](int e, float v) {
                const int pt = e / 3, comp = e - 3 * pt;
                (comp == 0 ? sx : (comp == 1 ? sy : sz))[pt] = v;
            };
            if (tid < head) put(tid, __ldg(g + tid));
            const float4 *g4 = reinterpret_cast<const float4 *>(g + head);
            for (int i = tid; i < nvec; i += THREADS) {
                const float4 v = __ldg(g4 + i);
                const int e = head + 4 * i;
                put(e, v.x);
                put(e + 1, v.y);
                put(e + 2, v.z);
                put(e + 3, v.w);
            }
            if (tid < nf - tail0) put(tail0 + tid, __ldg(g + tail0 + tid));
            // pad the last partial 32-chunk with +inf so it can never be a hit
            const int padded = (tn + 31) & ~31;
            for (int i = tn + tid; i < padded; i += THREADS) {
                sx[i] = __int_as_float(0x7f800000);
                sy[i] = __int_as_float(0x7f800000);
                sz[i] = __int_as_float(0x7f800000);
            }
        }
        __syncthreads();

#pragma unroll
        for (int h = 0; h < QPW / 2; ++h) {
            int c0 = c[2 * h], c1 = c[2 * h + 1];
            if (c0 < K || c1 < K) {
                int32_t *row0 = p.nbr + (q0 + qw + 2 * h) * (int64_t)K;
                int32_t *row1 = row0 + K;
                for (int j0 = 0; j0 < tn; j0 += 32) {
                    const float x = sx[j0 + lane], y = sy[j0 + lane], z = sz[j0 + lane];
                    float d0, d1;
                    {
                        const u64 dx = sub2(pack2(x, x), QX[h]), dy = sub2(pack2(y, y), QY[h]),
                                  dz = sub2(pack2(z, z), QZ[h]);
                        const u64 ax = mul2(dx, dx), ay = mul2(dy, dy), az = mul2(dz, dz);
                        d0 = __fadd_rn(__fadd_rn(lo32(ax), lo32(ay)), lo32(az));
                        d1 = __fadd_rn(__fadd_rn(hi32(ax), hi32(ay)), hi32(az));
                    }
                    const int gidx = (int)(s0 + t0 + j0 + lane);
                    const bool h0 = d0 < r2, h1 = d1 < r2;
                    const unsigned b0 = __ballot_sync(0xffffffffu, h0);
                    const unsigned b1 = __ballot_sync(0xffffffffu, h1);
                    if (b0 != 0u && c0 < K) {
                        const int slot = c0 + __popc(b0 & lt_mask);
                        if (h0 && slot < K) row0[slot] = gidx;
                        c0 += __popc(b0);
                    }
                    if (b1 != 0u && c1 < K) {
                        const int slot = c1 + __popc(b1 & lt_mask);
                        if (h1 && slot < K) row1[slot] = gidx;
                        c1 += __popc(b1);
                    }
                    if (c0 >= K && c1 >= K) break;
                }
                c[2 * h] = c0;
                c[2 * h + 1] = c1;
            }
        }
        if (t0 + TILE < n) __syncthreads();  // before the next tile overwrites shared memory
    }

    // ---- counts and -1 padding --------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < QPW; ++i) {
        const int ql = qw + i;
        if (ql < mq) {
            const int ci = min(c[i], K);
            int32_t *row = p.nbr + (q0 + ql) * (int64_t)K;
            for (int s = ci + lane; s < K; s += 32) row[s] = -1;
            if (lane == 0) p.cnt[q0 + ql] = ci;
        }
    }
}

}  // namespace b2pn

extern "C" int b2pn_ball_query_f32(const float *src_pos, const float *qry_pos, const int64_t *src_ptr,
                                   const int64_t *qry_ptr, int32_t B, int64_t max_src, int64_t max_qry, double r,
                                   int32_t K, int32_t *nbr, int32_t *cnt, b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0 || max_src < 0 || max_qry < 0 || K <= 0 || !(r >= 0.0)) return B2PN_EINVAL;
    if (B == 0 || max_qry == 0) return B2PN_OK;
    if (!src_pos || !qry_pos || !src_ptr || !qry_ptr || !nbr || !cnt) return B2PN_EINVAL;
    BqParams p = {src_pos, qry_pos, src_ptr, qry_ptr, nbr, cnt, (float)(r * r), K};
    constexpr int THREADS = 256, QPW = 8, TILE = 4096;
    constexpr int QPB = (THREADS / 32) * QPW;
    // grid.x covers the cloud with the most queries (host scalar, no device read)
    const unsigned gx = (unsigned)((max_qry + QPB - 1) / QPB);
    dim3 grid(gx, (unsigned)B);
    ball_query_kernel<THREADS, QPW, TILE><<<grid, THREADS, 0, (cudaStream_t)stream>>>(p);
    note_launch();
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}

// =================================================================================================
//  Kernel 2b -- the same query through a uniform grid (large clouds).
//
//  The brute-force kernel above tests every source of the cloud against every query (m*n tests, early exit once K
//  hits are found in ascending index order).  At the reference's level-1 radius a query has ~24 neighbours among
//  10 000 points, so almost all of that scan is wasted.  Here the sources of a cloud are binned into cells of edge
//  >= 1.0001 r (counting sort in shared memory, one CTA per cloud); a query only visits the 27 cells around its own,
//  collects the hits of the SAME fp32 distance test into a per-warp list and sorts the list by source index, which
//  restores the canonical "first K by ascending index" result bit for bit whatever order the cells were visited in.
// =================================================================================================
namespace b2pn {

constexpr int GRID_MAX_CELLS = 8192;    // per cloud: 32 KB of counters in shared memory
constexpr int GRID_LIST_CAP = 1024;     // hits a warp can hold before it falls back to the plain scan

struct GridInfo {   // one per cloud, written by the build kernel
    float lo[3];
    float inv[3];   // 1 / cell edge
    int n[3];       // cells per axis
    int pad;
};

struct BqGridParams {
    const float *src;
    const float *qry;
    const int64_t *src_ptr;
    const int64_t *qry_ptr;
    int32_t *nbr;
    int32_t *cnt;
    float r, r2;
    int K;
    GridInfo *info;         // [B]
    int32_t *cell_start;    // [B][GRID_MAX_CELLS + 1]
    int32_t *cell_pts;      // [N] cloud-local source indices grouped by cell
    float4 *cell_rec;       // [N] the same points as records {x, y, z, bits of the cloud-local index}, in cell order: a
                            // query streams its candidates with ONE coalesced load each instead of index -> position gathers
};

__device__ __forceinline__ int grid_cell_of(const GridInfo &g, float x, float y, float z)
{
    int cx = (int)((x - g.lo[0]) * g.inv[0]), cy = (int)((y - g.lo[1]) * g.inv[1]), cz = (int)((z - g.lo[2]) * g.inv[2]);
    cx = min(max(cx, 0), g.n[0] - 1);
    cy = min(max(cy, 0), g.n[1] - 1);
    cz = min(max(cz, 0), g.n[2] - 1);
    return (cz * g.n[1] + cy) * g.n[0] + cx;
}

__global__ void __launch_bounds__(1024) bq_grid_build_kernel(const BqGridParams p)
{
    __shared__ int s_cnt[GRID_MAX_CELLS];
    __shared__ float s_red[6][32];
    __shared__ GridInfo s_g;
    __shared__ int s_warp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t s0 = p.src_ptr[b];
    const int n = (int)(p.src_ptr[b + 1] - s0);
    const float *g = p.src + 3 * s0;
    // bounding box
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += 1024) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(g + 3 * i + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            s_red[a][warp] = mn[a];
            s_red[3 + a][warp] = mx[a];
        }
    }
    for (int i = tid; i < GRID_MAX_CELLS; i += 1024) s_cnt[i] = 0;
    __syncthreads();
    if (tid == 0) {
        GridInfo gi;
        int nc[3];
        const float edge = p.r * 1.0001f;  // strictly larger than r: two points closer than r are at most one cell apart
        for (int a = 0; a < 3; ++a) {
            float lo = INFINITY, hi = -INFINITY;
            for (int w = 0; w < 32; ++w) {
                lo = fminf(lo, s_red[a][w]);
                hi = fmaxf(hi, s_red[3 + a][w]);
            }
            if (!(lo <= hi)) lo = hi = 0.f;
            const float ext = hi - lo;
            int c = edge > 0.f ? (int)fminf(ext / edge, 64.f) : 1;
            nc[a] = c < 1 ? 1 : c;
            gi.lo[a] = lo;
        }
        while ((int64_t)nc[0] * nc[1] * nc[2] > GRID_MAX_CELLS) {  // halve the finest axis: cells only grow
            int a = nc[0] >= nc[1] ? (nc[0] >= nc[2] ? 0 : 2) : (nc[1] >= nc[2] ? 1 : 2);
            nc[a] = (nc[a] + 1) / 2;
        }
        for (int a = 0; a < 3; ++a) {
            float lo = gi.lo[a], hi = -INFINITY;
            for (int w = 0; w < 32; ++w) hi = fmaxf(hi, s_red[3 + a][w]);
            if (!(lo <= hi)) hi = lo;
            const float ext = hi - lo;
            gi.n[a] = nc[a];
            gi.inv[a] = ext > 0.f ? (float)nc[a] / ext : 0.f;  // cell edge = ext / n >= 1.0001 r
        }
        gi.pad = 0;
        s_g = gi;
        p.info[b] = gi;
    }
    __syncthreads();
    const GridInfo gi = s_g;
    const int ncell = gi.n[0] * gi.n[1] * gi.n[2];
    for (int i = tid; i < n; i += 1024) atomicAdd(&s_cnt[grid_cell_of(gi, __ldg(g + 3 * i), __ldg(g + 3 * i + 1), __ldg(g + 3 * i + 2))], 1);
    __syncthreads();
    // exclusive scan of the counters (ncell <= 8192 = 8 per thread)
    const int per = (ncell + 1023) / 1024;
    const int c0 = tid * per;
    int sum = 0;
    for (int j = 0; j < per; ++j)
        if (c0 + j < ncell) sum += s_cnt[c0 + j];
    int v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    int run = (warp > 0 ? s_warp[warp - 1] : 0) + v - sum;
    int32_t *cs = p.cell_start + (int64_t)b * (GRID_MAX_CELLS + 1);
    for (int j = 0; j < per; ++j) {
        if (c0 + j < ncell) {
            const int c = s_cnt[c0 + j];
            cs[c0 + j] = run;
            s_cnt[c0 + j] = run;  // becomes the running insert position
            run += c;
        }
    }
    if (tid == 0) cs[ncell] = n;
    __syncthreads();
    int32_t *pts = p.cell_pts + s0;
    float4 *rec = p.cell_rec + s0;
    for (int i = tid; i < n; i += 1024) {
        const float x = __ldg(g + 3 * i), y = __ldg(g + 3 * i + 1), z = __ldg(g + 3 * i + 2);
        const int c = grid_cell_of(gi, x, y, z);
        const int slot = atomicAdd(&s_cnt[c], 1);  // order inside a cell is arbitrary: the query sorts its hits
        pts[slot] = i;
        rec[slot] = make_float4(x, y, z, __int_as_float(i));
    }
}

// one warp per query
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) bq_grid_query_kernel(const BqGridParams p)
{
    __shared__ int s_list[WARPS][GRID_LIST_CAP];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t s0 = p.src_ptr[b];
    const int n = (int)(p.src_ptr[b + 1] - s0);
    const int64_t q0 = p.qry_ptr[b];
    const int mq = (int)(p.qry_ptr[b + 1] - q0);
    const int ql = blockIdx.x * WARPS + warp;
    if (ql >= mq) return;
    const GridInfo gi = p.info[b];
    const float *gsrc = p.src + 3 * s0;
    const int32_t *cs = p.cell_start + (int64_t)b * (GRID_MAX_CELLS + 1);
    const float qx = __ldg(p.qry + 3 * (q0 + ql)), qy = __ldg(p.qry + 3 * (q0 + ql) + 1), qz = __ldg(p.qry + 3 * (q0 + ql) + 2);
    const unsigned lt_mask = (1u << lane) - 1u;
    int *list = s_list[warp];
    const int K = p.K;
    int h = 0;            // hits collected (warp-uniform)
    bool overflow = false;
    {
        int cx = (int)((qx - gi.lo[0]) * gi.inv[0]), cy = (int)((qy - gi.lo[1]) * gi.inv[1]), cz = (int)((qz - gi.lo[2]) * gi.inv[2]);
        cx = min(max(cx, 0), gi.n[0] - 1);
        cy = min(max(cy, 0), gi.n[1] - 1);
        cz = min(max(cz, 0), gi.n[2] - 1);
        // The nine (z, y) cell rows around the query are nine contiguous ranges of the cell-ordered records (the three
        // x-neighbours are consecutive cells).  Lanes 0..8 hold one range each; the warp then walks the CONCATENATION of
        // the ranges 32 candidates at a time, so a query costs ceil(candidates / 32) coalesced record loads instead of
        // nine (or more) rounds of two dependent gathers (index, then position).
        int beg = 0, len = 0;
        if (lane < 9) {
            const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
            if (z >= 0 && z < gi.n[2] && y >= 0 && y < gi.n[1]) {
                const int x0 = max(cx - 1, 0), x1 = min(cx + 1, gi.n[0] - 1);
                const int c_lo = (z * gi.n[1] + y) * gi.n[0] + x0;
                beg = cs[c_lo];
                len = cs[c_lo + (x1 - x0) + 1] - beg;
            }
        }
        int pre = len;   // inclusive prefix of the range lengths over lanes 0..8
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const int total = __shfl_sync(0xffffffffu, pre, 8);
        // Dense neighbourhood?  The ball holds ~15 % of the 27 cells' points; once that is several times K the
        // ascending scan finds its K hits after a short prefix of the cloud and beats collecting + sorting them all.
        overflow = total > 13 * K;
        const float4 *rec = p.cell_rec + s0;
        for (int t0 = 0; t0 < total && !overflow; t0 += 32) {
            const int t = t0 + lane;
            int j = -1;
#pragma unroll
            for (int rr = 0; rr < 9; ++rr) {   // the range candidate t falls into
                const int hi = __shfl_sync(0xffffffffu, pre, rr), ln = __shfl_sync(0xffffffffu, len, rr),
                          bg = __shfl_sync(0xffffffffu, beg, rr);
                if (j < 0 && t < hi) j = bg + (t - (hi - ln));
            }
            int idx = -1;
            bool hit = false;
            if (t < total) {
                const float4 q = __ldg(rec + j);
                idx = __float_as_int(q.w);
                hit = dist2_scalar(q.x, q.y, q.z, qx, qy, qz) < p.r2;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m != 0u) {
                const int pos = h + __popc(m & lt_mask);
                if (hit && pos < GRID_LIST_CAP) list[pos] = idx;
                h += __popc(m);
                if (h > GRID_LIST_CAP) overflow = true;
            }
        }
    }
    int32_t *row = p.nbr + (q0 + ql) * (int64_t)K;
    if (overflow) {
        // (rare) more hits than the list holds: plain ascending scan with early exit, like Kernel 2
        // four batches of 32 points per round: their loads are in flight together (the early-exit test made every round
        // of the plain loop wait for its own L2 round trip), the ballots keep the ascending order
        int c = 0;
        for (int j0 = 0; j0 < n && c < K; j0 += 128) {
            bool hit[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * 32 + lane;
                hit[u] = false;
                if (j < n) hit[u] = dist2_scalar(__ldg(gsrc + 3 * j), __ldg(gsrc + 3 * j + 1), __ldg(gsrc + 3 * j + 2), qx, qy, qz) < p.r2;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned m = __ballot_sync(0xffffffffu, hit[u]);
                const int slot = c + __popc(m & lt_mask);
                if (hit[u] && slot < K) row[slot] = (int)(s0 + j0 + u * 32 + lane);
                c += __popc(m);
            }
        }
        c = min(c, K);
        for (int s = c + lane; s < K; s += 32) row[s] = -1;
        if (lane == 0) p.cnt[q0 + ql] = c;
        return;
    }
    __syncwarp();
    // sort the h hits by source index (bitonic over the next power of two, padded with INT_MAX)
    int np = 32;
    while (np < h) np <<= 1;
    for (int i = h + lane; i < np; i += 32) list[i] = 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= np; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (np >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const int a = list[i], c = list[l];
                const bool up = (i & k) == 0;
                if ((a > c) == up) {
                    list[i] = c;
                    list[l] = a;
                }
            }
            __syncwarp();
        }
    }
    const int c = min(h, K);
    for (int s = lane; s < K; s += 32) row[s] = s < c ? (int)(s0 + list[s]) : -1;
    if (lane == 0) p.cnt[q0 + ql] = c;
}

}  // namespace b2pn

extern "C" int64_t b2pn_ball_query_workspace_bytes(int32_t B, int64_t n_src_total)
{
    if (B < 0 || n_src_total < 0) return B2PN_EINVAL;
    return (int64_t)B * (sizeof(b2pn::GridInfo) + (b2pn::GRID_MAX_CELLS + 1) * sizeof(int32_t)) +
           n_src_total * (int64_t)(sizeof(int32_t) + sizeof(float4)) + 2048;
}

extern "C" int b2pn_ball_query_grid_f32(const float *src_pos, const float *qry_pos, const int64_t *src_ptr,
                                        const int64_t *qry_ptr, int32_t B, int64_t n_src_total, int64_t max_qry, double r,
                                        int32_t K, int32_t *nbr, int32_t *cnt, void *workspace, int64_t workspace_bytes,
                                        b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0 || n_src_total < 0 || max_qry < 0 || K <= 0 || !(r >= 0.0)) return B2PN_EINVAL;
    if (B == 0 || max_qry == 0) return B2PN_OK;
    if (!src_pos || !qry_pos || !src_ptr || !qry_ptr || !nbr || !cnt || !workspace) return B2PN_EINVAL;
    if (workspace_bytes < b2pn_ball_query_workspace_bytes(B, n_src_total)) return B2PN_EINVAL;
    char *w = (char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    GridInfo *info = (GridInfo *)w;
    w += ((int64_t)B * sizeof(GridInfo) + 255) / 256 * 256;
    int32_t *cell_start = (int32_t *)w;
    w += ((int64_t)B * (GRID_MAX_CELLS + 1) * sizeof(int32_t) + 255) / 256 * 256;
    int32_t *cell_pts = (int32_t *)w;
    w += (n_src_total * (int64_t)sizeof(int32_t) + 255) / 256 * 256;
    float4 *cell_rec = (float4 *)w;
    BqGridParams p = {src_pos, qry_pos, src_ptr, qry_ptr, nbr, cnt, (float)r, (float)(r * r), K, info, cell_start, cell_pts, cell_rec};
    cudaStream_t st = (cudaStream_t)stream;
    bq_grid_build_kernel<<<(unsigned)B, 1024, 0, st>>>(p);
    note_launch();
    constexpr int WARPS = 8;
    dim3 grid((unsigned)((max_qry + WARPS - 1) / WARPS), (unsigned)B);
    bq_grid_query_kernel<WARPS><<<grid, WARPS * 32, 0, st>>>(p);
    note_launch();
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}
