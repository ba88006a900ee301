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
