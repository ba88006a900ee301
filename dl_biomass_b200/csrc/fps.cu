// Kernel 1 -- farthest-point sampling, register-resident, one thread-block cluster per cloud.
//
// Replaces torch_cluster's fps kernel reached from /root/reference/pointnet2_regressor.py:13.
// Semantics: explicit start point, running min-distance, arg-max with ties to
// the lowest point index, d2 = ((dx*dx+dy*dy)+dz*dz) in separately rounded fp32.
//
// B200 design: the cloud never leaves the SM(s) after the first load.  A cluster of CLUSTER CTAs
// owns one cloud; thread g = rank*THREADS + tid keeps points [g*c, g*c+c) (c <= PPT) and their
// running distances in REGISTERS (packed f32x2 for FADD2/FMUL2).  Because ownership is
// contiguous and ascending in g, "lowest index among ties" is "lowest (rank, warp, lane, slot)",
// so each reduction level is one REDUX.MAX on the distance bits plus a ballot/ffs:
//   thread max (FMNMX) -> warp (redux.sync.max.s32) -> CTA (smem, 1 barrier) -> cluster (DSMEM
//   records + barrier.cluster).  The winner's xyz travels with the record, so the next iteration
//   starts without another memory round trip.
#include <limits.h>

#include "common.cuh"

namespace b2pn {

struct FpsParams {
    const float *pos;
    const int64_t *ptr;
    const int64_t *out_ptr;
    const int64_t *start;
    int64_t *out_idx;
    float *out_pos;
    int64_t *out_batch;
    // random_start drawn inside the kernel (torch_cluster.fps default): start_b = floor(u * n_b) with u a
    // counter-based uniform of (seed, rng_state[0], b).  rng_state[0] is a DEVICE call counter so that a CUDA-graph
    // replay draws fresh starts; every block reads it on entry and the last block to leave bumps it (ticket rng_state[1]).
    unsigned long long seed;
    int64_t *rng_state;
};

// start index of cloud b for call number `call`: floor(u * n), u uniform in [0, 1) with 24 random bits
__host__ __device__ __forceinline__ int fps_random_start(unsigned long long seed, long long call, int b, int n)
{
    unsigned long long z = seed * 0x9e3779b97f4a7c15ull + (unsigned long long)call * 0x100000001b3ull + (unsigned long long)b;
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    const float u = (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
    int s = (int)(u * (float)n);
    return s < n ? s : n - 1;
}
// called once per block on every exit path (one thread): the last block publishes the next call number
__device__ __forceinline__ void fps_rng_leave(const FpsParams &p, long long call)
{
    if (p.rng_state == nullptr) return;
    __threadfence();
    const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long *>(p.rng_state + 1), 1ull);
    if (t == (unsigned long long)gridDim.x - 1ull) {
        p.rng_state[1] = 0;
        p.rng_state[0] = call + 1;
        __threadfence();
    }
}

template <int CLUSTER, int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1) fps_kernel(const FpsParams p)
{
    static_assert(PPT % 2 == 0, "points per thread must be even (packed f32x2)");
    constexpr int NW = THREADS / 32;
    const int b = blockIdx.x / CLUSTER;
    const unsigned rank = (CLUSTER > 1) ? cluster_ctarank() : 0u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int64_t p0 = p.ptr[b];
    const int n = (int)(p.ptr[b + 1] - p0);
    const int64_t o0 = p.out_ptr[b];
    const int m = (int)(p.out_ptr[b + 1] - o0);
    const long long call = p.rng_state ? (long long)p.rng_state[0] : 0ll;
    if (n <= 0 || m <= 0) {  // uniform over the cluster
        if (tid == 0) fps_rng_leave(p, call);
        return;
    }

    __shared__ int s_wmax[NW];
    __shared__ __align__(16) unsigned s_rec[2][CLUSTER][8];  // {key, idx, x, y, z, -, -, -}

    // ---- load my points into registers ----------------------------------------------------------
    const int c = (n + CLUSTER * THREADS - 1) / (CLUSTER * THREADS);  // points per thread in use
    const int g = (int)rank * THREADS + tid;
    const int base = g * c;
    u64 X[PPT / 2], Y[PPT / 2], Z[PPT / 2];
    float D[PPT];
    {
        const int lim = min(c, n - base);                       // my valid slots (may be <= 0)
        const float *q = p.pos + 3 * (p0 + (int64_t)base);       // only dereferenced where k < lim
#pragma unroll
        for (int j = 0; j < PPT / 2; ++j) {
            float xs[2], ys[2], zs[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 2 * j + h;
                const bool ok = k < lim;
                xs[h] = ok ? __ldg(q + 3 * k + 0) : 0.f;
                ys[h] = ok ? __ldg(q + 3 * k + 1) : 0.f;
                zs[h] = ok ? __ldg(q + 3 * k + 2) : 0.f;
                D[k] = ok ? __int_as_float(0x7f800000) : -1.f;  // +inf: first update sets dist-to-start
            }
            X[j] = pack2(xs[0], xs[1]);
            Y[j] = pack2(ys[0], ys[1]);
            Z[j] = pack2(zs[0], zs[1]);
        }
    }

    // ---- start point ------------------------------------------------------------------------------
    int cur = 0;
    if (p.start != nullptr) {
        const int64_t s = p.start[b];
        cur = (s >= 0 && s < n) ? (int)s : 0;
    } else if (p.rng_state != nullptr) {
        cur = fps_random_start(p.seed, call, b, n);
    }
    float cx = __ldg(p.pos + 3 * (p0 + cur) + 0);
    float cy = __ldg(p.pos + 3 * (p0 + cur) + 1);
    float cz = __ldg(p.pos + 3 * (p0 + cur) + 2);
    if (rank == 0 && tid == 0) {
        p.out_idx[o0] = p0 + cur;
        if (p.out_pos) {
            p.out_pos[3 * o0 + 0] = cx;
            p.out_pos[3 * o0 + 1] = cy;
            p.out_pos[3 * o0 + 2] = cz;
        }
    }
    if (p.out_batch) {
#pragma unroll 1
        for (int i = g; i < m; i += CLUSTER * THREADS) p.out_batch[o0 + i] = b;
    }

    for (int it = 1; it < m; ++it) {
        // 1. update running distances with the last winner, thread-local max
        const u64 px = pack2(cx, cx), py = pack2(cy, cy), pz = pack2(cz, cz);
        float lmax = -1.f;
#pragma unroll
        for (int j = 0; j < PPT / 2; ++j) {
            float d0, d1;
            dist2_pair(X[j], Y[j], Z[j], px, py, pz, d0, d1);
            D[2 * j] = fminf(D[2 * j], d0);
            D[2 * j + 1] = fminf(D[2 * j + 1], d1);
            lmax = fmaxf(lmax, fmaxf(D[2 * j], D[2 * j + 1]));
        }
        // 2. warp arg-max (distances are >= 0 or the -1 pad, so signed-int order == float order)
        const int lb = __float_as_int(lmax);
        const int wmax = __reduce_max_sync(0xffffffffu, lb);
        const unsigned cand = __ballot_sync(0xffffffffu, lb == wmax);
        const int wlane = __ffs(cand) - 1;
        if (lane == 0) s_wmax[warp] = wmax;
        __syncthreads();
        // 3. CTA arg-max, computed redundantly by every warp
        const int v = (lane < NW) ? s_wmax[lane] : INT_MIN;
        const int cmax = __reduce_max_sync(0xffffffffu, v);
        const int wwarp = __ffs(__ballot_sync(0xffffffffu, v == cmax)) - 1;
        const int par = it & 1;
        if (warp == wwarp) {
            // the winning lane recovers its slot (lowest k with D[k] == cmax) and its coordinates
            int ksel = 0;
            float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
            for (int k = PPT - 1; k >= 0; --k) {
                if (__float_as_int(D[k]) == cmax) {
                    ksel = k;
                    wx = (k & 1) ? hi32(X[k >> 1]) : lo32(X[k >> 1]);
                    wy = (k & 1) ? hi32(Y[k >> 1]) : lo32(Y[k >> 1]);
                    wz = (k & 1) ? hi32(Z[k >> 1]) : lo32(Z[k >> 1]);
                }
            }
            const int widx = base + ksel;
            if (CLUSTER == 1) {
                if (lane == wlane) {
                    *reinterpret_cast<uint4 *>(&s_rec[par][0][0]) =
                        make_uint4((unsigned)cmax, (unsigned)widx, __float_as_uint(wx), __float_as_uint(wy));
                    s_rec[par][0][4] = __float_as_uint(wz);
                    p.out_idx[o0 + it] = p0 + widx;
                    if (p.out_pos) {
                        p.out_pos[3 * (o0 + it) + 0] = wx;
                        p.out_pos[3 * (o0 + it) + 1] = wy;
                        p.out_pos[3 * (o0 + it) + 2] = wz;
                    }
                }
            } else {
                // broadcast the winning lane's record over the warp; lane r stores it into CTA r
                const unsigned ri = __shfl_sync(0xffffffffu, (unsigned)widx, wlane);
                const unsigned rx = __shfl_sync(0xffffffffu, __float_as_uint(wx), wlane);
                const unsigned ry = __shfl_sync(0xffffffffu, __float_as_uint(wy), wlane);
                const unsigned rz = __shfl_sync(0xffffffffu, __float_as_uint(wz), wlane);
                if (lane < CLUSTER) {
                    const unsigned dst = mapa_u32(smem_u32(&s_rec[par][rank][0]), (unsigned)lane);
                    st_cluster_v4(dst, (unsigned)cmax, ri, rx, ry);
                    st_cluster_v2(dst + 16, rz, 0u);
                }
            }
        }
        if (CLUSTER == 1) {
            __syncthreads();
            const uint4 r = *reinterpret_cast<const uint4 *>(&s_rec[par][0][0]);
            cx = __uint_as_float(r.z);
            cy = __uint_as_float(r.w);
            cz = __uint_as_float(s_rec[par][0][4]);
        } else {
            cluster_arrive_release();
            cluster_wait_acquire();
            // 4. cluster arg-max over the CLUSTER records (ascending rank == ascending index)
            const int key = (lane < CLUSTER) ? (int)s_rec[par][lane][0] : INT_MIN;
            const int gmax = __reduce_max_sync(0xffffffffu, key);
            const int wr = __ffs(__ballot_sync(0xffffffffu, key == gmax)) - 1;
            const uint4 r = *reinterpret_cast<const uint4 *>(&s_rec[par][wr][0]);
            cx = __uint_as_float(r.z);
            cy = __uint_as_float(r.w);
            cz = __uint_as_float(s_rec[par][wr][4]);
            if (rank == 0 && tid == 0) {
                p.out_idx[o0 + it] = p0 + (int)r.y;
                if (p.out_pos) {
                    p.out_pos[3 * (o0 + it) + 0] = cx;
                    p.out_pos[3 * (o0 + it) + 1] = cy;
                    p.out_pos[3 * (o0 + it) + 2] = cz;
                }
            }
        }
    }
    if (CLUSTER > 1) {  // nobody may exit while a peer can still write into its shared memory
        cluster_arrive_release();
        cluster_wait_acquire();
    }
    if (tid == 0) fps_rng_leave(p, call);
}

// =================================================================================================
//  Kernel 1b -- the same sampling with SPATIAL PRUNING (one CTA per cloud).
//
//  The points are sorted along a Morton curve inside the kernel (bitonic sort in shared memory), so
//  every warp owns a spatially compact bucket and keeps its bounding box.  A new sample c can only
//  lower the running distance of a point p if d2(p,c) < D[p] <= max D of p's bucket, and for every p of
//  a bucket d2(p,c) >= d2(box,c) IN THE SAME fp32 ARITHMETIC (sub/mul/add are monotone under
//  round-to-nearest and the box distance is built from the same operations), so a bucket with
//  d2(box,c) >= its current max is skipped without changing a single bit of the result.  Late in the
//  sampling a new point touches 2-4 buckets of 32, so the scan shrinks by ~8x and an iteration costs
//  one block barrier: every warp publishes {max, index, xyz} (from registers; recomputed only when
//  the bucket was touched), then every warp reduces the 32 records redundantly.
//  Ties go to the lowest ORIGINAL index at every level (perm[] maps sorted position -> original).
// =================================================================================================
__device__ __forceinline__ unsigned morton_spread10(unsigned v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

template <int THREADS, int MAXS>
__global__ void __launch_bounds__(THREADS, 1) fps_pruned_kernel(const FpsParams p, const int npad)
{
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char fps_dyn[];
    u64 *keys = reinterpret_cast<u64 *>(fps_dyn);  // [npad] during the sort
    __shared__ float s_red[6][NW];
    __shared__ int s_rec[2][5][32];  // per warp {max distance bits, original index, x, y, z}, double-buffered

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = p.ptr[b];
    const int n = (int)(p.ptr[b + 1] - p0);
    const int64_t o0 = p.out_ptr[b];
    const int m = (int)(p.out_ptr[b + 1] - o0);
    if (n <= 0 || m <= 0) return;
    const float *gpos = p.pos + 3 * p0;

    // ---- 1. cloud bounding box -> 10-bit quantisation per axis ----------------------------------------------
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = tid; i < n; i += THREADS) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(gpos + 3 * i + a);
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
    }
    __syncthreads();
    float qlo[3], qsc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = INFINITY, mx = -INFINITY;
        for (int w = 0; w < NW; ++w) {
            mn = fminf(mn, s_red[a][w]);
            mx = fmaxf(mx, s_red[3 + a][w]);
        }
        qlo[a] = mn;
        const float ext = mx - mn;
        qsc[a] = ext > 0.f ? 1023.f / ext : 0.f;
    }
    // ---- 2. Morton keys (the sort order only decides the bucketing, never the result) -----------------------
    for (int i = tid; i < npad; i += THREADS) {
        u64 k = ~0ull;
        if (i < n) {
            unsigned code = 0u;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(gpos + 3 * i + a);
                float q = (v - qlo[a]) * qsc[a];
                q = fminf(fmaxf(q, 0.f), 1023.f);
                code |= morton_spread10((unsigned)q) << a;
            }
            k = ((u64)code << 32) | (u64)(unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    // ---- 3. bitonic sort of the 64-bit keys (unique, so the order is deterministic) --------------------------
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const u64 a = keys[i], c = keys[l];
                const bool up = (i & k) == 0;
                if ((a > c) == up) {
                    keys[i] = c;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
    }
    // ---- 4. buckets: warp w owns S buckets of 32 consecutive points of the Morton order; point (bucket k, lane l)
    //         sits at q = (w*S + k)*32 + l of the SoA arrays below.  Lane k keeps bucket k's box and running max. ---
    const int S = (n + THREADS - 1) / THREADS;  // buckets per warp in use (<= MAXS <= 32)
    const int NP = THREADS * S;
    unsigned myidx[MAXS];
#pragma unroll
    for (int k = 0; k < MAXS; ++k) {
        const int q = (warp * S + k) * 32 + lane;
        myidx[k] = (k < S && q < n) ? (unsigned)(keys[q] & 0xffffffffull) : 0xffffffffu;
    }
    __syncthreads();  // everybody has read its keys: the buffer becomes the point arrays
    float *sx = reinterpret_cast<float *>(fps_dyn);
    float *sy = sx + NP, *sz = sy + NP, *sD = sz + NP;
    unsigned short *sperm = reinterpret_cast<unsigned short *>(sD + NP);  // original index (n <= 65535)
    float blx = INFINITY, bly = INFINITY, blz = INFINITY, bhx = -INFINITY, bhy = -INFINITY, bhz = -INFINITY;
    float bmax = -1.f;
#pragma unroll
    for (int k = 0; k < MAXS; ++k) {
        if (k < S) {
            const bool ok = myidx[k] != 0xffffffffu;
            const float *g = gpos + 3 * (int64_t)(ok ? myidx[k] : 0u);
            const float x = ok ? __ldg(g + 0) : 0.f, y = ok ? __ldg(g + 1) : 0.f, z = ok ? __ldg(g + 2) : 0.f;
            const int q = (warp * S + k) * 32 + lane;
            sx[q] = x;
            sy[q] = y;
            sz[q] = z;
            sD[q] = ok ? __int_as_float(0x7f800000) : -1.f;  // +inf: the first update sets dist-to-start
            sperm[q] = (unsigned short)(ok ? myidx[k] : 0xffffu);
            float lx = ok ? x : INFINITY, ly = ok ? y : INFINITY, lz = ok ? z : INFINITY;
            float hx = ok ? x : -INFINITY, hy = ok ? y : -INFINITY, hz = ok ? z : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
                ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
                lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            }
            if (lane == k) {
                blx = lx; bly = ly; blz = lz; bhx = hx; bhy = hy; bhz = hz;
                bmax = (lx <= hx) ? __int_as_float(0x7f800000) : -1.f;  // empty bucket: never touched
            }
        }
    }

    // ---- start point ---------------------------------------------------------------------------------------------
    int cur = 0;
    if (p.start != nullptr) {
        const int64_t s = p.start[b];
        cur = (s >= 0 && s < n) ? (int)s : 0;
    }
    float cx = __ldg(gpos + 3 * cur + 0), cy = __ldg(gpos + 3 * cur + 1), cz = __ldg(gpos + 3 * cur + 2);
    if (tid == 0) {
        p.out_idx[o0] = p0 + cur;
        if (p.out_pos) {
            p.out_pos[3 * o0 + 0] = cx;
            p.out_pos[3 * o0 + 1] = cy;
            p.out_pos[3 * o0 + 2] = cz;
        }
    }
    if (p.out_batch) {
#pragma unroll 1
        for (int i = tid; i < m; i += THREADS) p.out_batch[o0 + i] = b;
    }
    if (lane < 5) s_rec[0][lane][warp] = lane == 0 ? __float_as_int(-1.f) : (lane == 1 ? 0x7fffffff : 0);
    __syncwarp();

    const int qw = warp * S * 32 + lane;
    for (int it = 1; it < m; ++it) {
        const int par = it & 1;
        // lane k: distance from the new sample to the box of bucket k, in the arithmetic of the point distances
        const float ex = fmaxf(fmaxf(__fsub_rn(blx, cx), __fsub_rn(cx, bhx)), 0.f);
        const float ey = fmaxf(fmaxf(__fsub_rn(bly, cy), __fsub_rn(cy, bhy)), 0.f);
        const float ez = fmaxf(fmaxf(__fsub_rn(blz, cz), __fsub_rn(cz, bhz)), 0.f);
        const float bd = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        unsigned hits = __ballot_sync(0xffffffffu, bd < bmax);  // empty buckets: bmax = -1
        if (hits != 0u) {                                       // warp-uniform
            // touched buckets, two at a time so that their shared-memory round trips overlap
            while (hits != 0u) {
                const int k0 = __ffs(hits) - 1;
                hits &= hits - 1u;
                const bool two = hits != 0u;
                const int k1 = two ? __ffs(hits) - 1 : k0;
                hits &= hits - 1u;  // 0 - 1 & 0 = 0: harmless when already empty
                const int q0 = qw + k0 * 32, q1 = qw + k1 * 32;
                const float x0 = sx[q0], y0 = sy[q0], z0 = sz[q0], D0 = sD[q0];
                const float x1 = sx[q1], y1 = sy[q1], z1 = sz[q1], D1 = sD[q1];
                const float n0 = fminf(D0, dist2_scalar(x0, y0, z0, cx, cy, cz));
                const float n1 = fminf(D1, dist2_scalar(x1, y1, z1, cx, cy, cz));
                sD[q0] = n0;
                if (two) sD[q1] = n1;
                const int m0 = __reduce_max_sync(0xffffffffu, __float_as_int(n0));
                const int m1 = __reduce_max_sync(0xffffffffu, __float_as_int(n1));
                if (lane == k0) bmax = __int_as_float(m0);
                if (lane == k1) bmax = __int_as_float(m1);
            }
            __syncwarp();
            // new warp record: max over my buckets, lowest original index among the points at the max
            const int wmax = __reduce_max_sync(0xffffffffu, __float_as_int(bmax));
            unsigned tb = __ballot_sync(0xffffffffu, __float_as_int(bmax) == wmax);
            unsigned best = 0xffffffffu;
            int bq = 0;
            do {  // one bucket unless distances tie across buckets
                const int k = __ffs(tb) - 1;
                tb &= tb - 1u;
                const int q = qw + k * 32;
                const unsigned cand = (__float_as_int(sD[q]) == wmax) ? (unsigned)sperm[q] : 0xffffffffu;
                if (cand < best) {
                    best = cand;
                    bq = q;
                }
            } while (tb != 0u);
            const unsigned widx = __reduce_min_sync(0xffffffffu, best);
            if (best == widx) {  // exactly one lane (original indices are unique); an empty warp never gets here
                s_rec[par][0][warp] = wmax;
                s_rec[par][1][warp] = (int)widx;
                s_rec[par][2][warp] = __float_as_int(sx[bq]);
                s_rec[par][3][warp] = __float_as_int(sy[bq]);
                s_rec[par][4][warp] = __float_as_int(sz[bq]);
            }
        } else if (lane < 5) {
            s_rec[par][lane][warp] = s_rec[par ^ 1][lane][warp];  // unchanged bucket set: republish
        }
        __syncthreads();
        const bool live = lane < NW;
        const int key = live ? s_rec[par][0][lane] : INT_MIN;
        const unsigned ki = live ? (unsigned)s_rec[par][1][lane] : 0xffffffffu;
        const float rx = live ? __int_as_float(s_rec[par][2][lane]) : 0.f;
        const float ry = live ? __int_as_float(s_rec[par][3][lane]) : 0.f;
        const float rz = live ? __int_as_float(s_rec[par][4][lane]) : 0.f;
        const int gmax = __reduce_max_sync(0xffffffffu, key);
        const unsigned sel = key == gmax ? ki : 0xffffffffu;
        const unsigned gidx = __reduce_min_sync(0xffffffffu, sel);
        const int ww = __ffs(__ballot_sync(0xffffffffu, sel == gidx)) - 1;
        cx = __shfl_sync(0xffffffffu, rx, ww);
        cy = __shfl_sync(0xffffffffu, ry, ww);
        cz = __shfl_sync(0xffffffffu, rz, ww);
        if (tid == 0) {
            p.out_idx[o0 + it] = p0 + (int64_t)gidx;
            if (p.out_pos) {
                p.out_pos[3 * (o0 + it) + 0] = cx;
                p.out_pos[3 * (o0 + it) + 1] = cy;
                p.out_pos[3 * (o0 + it) + 2] = cz;
            }
        }
    }
}

static int next_pow2(int v)
{
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

template <int THREADS, int MAXS>
static int launch_fps_pruned(const FpsParams &p, int B, int64_t max_n, cudaStream_t stream)
{
    auto kern = fps_pruned_kernel<THREADS, MAXS>;
    const int npad = next_pow2((int)max_n);
    const int S = (int)((max_n + THREADS - 1) / THREADS);
    const int after = THREADS * S * 18;  // x, y, z, D (fp32) + original index (u16) per slot
    const int smem = npad * 8 > after ? npad * 8 : after;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)B, THREADS, smem, stream>>>(p, npad);
    note_launch();
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// capacity of the pruned variant: 18 B of shared memory per point
constexpr int64_t PRUNED_MAX_N = 12288;

template <int THREADS>
static int dispatch_pruned_s(const FpsParams &p, int B, int64_t max_n, cudaStream_t stream)
{
    const int S = (int)((max_n + THREADS - 1) / THREADS);
    if (S <= 4) return launch_fps_pruned<THREADS, 4>(p, B, max_n, stream);
    if (S <= 8) return launch_fps_pruned<THREADS, 8>(p, B, max_n, stream);
    if (S <= 12) return launch_fps_pruned<THREADS, 12>(p, B, max_n, stream);
    if (S <= 24) return launch_fps_pruned<THREADS, 24>(p, B, max_n, stream);
    if (S <= 32) return launch_fps_pruned<THREADS, 32>(p, B, max_n, stream);
    return B2PN_ENOTSUP;
}

static int dispatch_pruned(const FpsParams &p, int B, int64_t max_n, int threads, cudaStream_t stream)
{
    if (max_n > PRUNED_MAX_N || max_n > 65535) return B2PN_ENOTSUP;
    if (threads == 0) threads = max_n <= 2048 ? 256 : (max_n <= 4096 ? 512 : 1024);
    if ((int64_t)threads * 32 < max_n) return B2PN_ENOTSUP;
    switch (threads) {
        case 256: return dispatch_pruned_s<256>(p, B, max_n, stream);
        case 512: return dispatch_pruned_s<512>(p, B, max_n, stream);
        case 1024: return dispatch_pruned_s<1024>(p, B, max_n, stream);
    }
    return B2PN_EINVAL;
}

// =================================================================================================
//  Kernel 1c -- the register scan of Kernel 1 with a per-THREAD spatial test in front of it.
//
//  Same layout and arg-max chain as fps_kernel (points and running distances in registers, REDUX per warp,
//  two block barriers), but the points are Morton-sorted first, so the PPT points a thread owns are neighbours in
//  space and a warp's 32 x PPT points form one compact region.  Before the update a thread tests the new sample
//  against the bounding box of its own points with the distance arithmetic of the points themselves (monotone
//  under round-to-nearest, see Kernel 1b); if NO thread of the warp can be affected, the warp skips its whole
//  update and keeps its cached maximum.  Results are bit-identical; ties go to the lowest ORIGINAL index: slots of
//  a thread are ordered by original index, lanes and warps that tie are resolved through perm[].
// =================================================================================================
template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1) fps_sorted_kernel(const FpsParams p, const int npad)
{
    static_assert(PPT % 2 == 0, "points per thread must be even (packed f32x2)");
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char fps_dyn[];
    u64 *keys = reinterpret_cast<u64 *>(fps_dyn);  // [npad] during the sort
    __shared__ float s_red[6][NW];
    __shared__ int s_wmax[32];
    __shared__ unsigned s_cand[32];
    __shared__ __align__(16) unsigned s_rec[2][8];  // {key, original index, x, y, z}

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = p.ptr[b];
    const int n = (int)(p.ptr[b + 1] - p0);
    const int64_t o0 = p.out_ptr[b];
    const int m = (int)(p.out_ptr[b + 1] - o0);
    if (n <= 0 || m <= 0) return;
    const float *gpos = p.pos + 3 * p0;

    // ---- Morton sort (decides the ownership of points only, never the result) -------------------------------
    {
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = tid; i < n; i += THREADS) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(gpos + 3 * i + a);
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
    }
    __syncthreads();
    {
        float qlo[3], qsc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float mn = INFINITY, mx = -INFINITY;
            for (int w = 0; w < NW; ++w) {
                mn = fminf(mn, s_red[a][w]);
                mx = fmaxf(mx, s_red[3 + a][w]);
            }
            qlo[a] = mn;
            const float ext = mx - mn;
            qsc[a] = ext > 0.f ? 1023.f / ext : 0.f;
        }
        for (int i = tid; i < npad; i += THREADS) {
            u64 k = ~0ull;
            if (i < n) {
                unsigned code = 0u;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float v = __ldg(gpos + 3 * i + a);
                    float q = (v - qlo[a]) * qsc[a];
                    q = fminf(fmaxf(q, 0.f), 1023.f);
                    code |= morton_spread10((unsigned)q) << a;
                }
                k = ((u64)code << 32) | (u64)(unsigned)i;
            }
            keys[i] = k;
        }
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const u64 a = keys[i], c = keys[l];
                const bool up = (i & k) == 0;
                if ((a > c) == up) {
                    keys[i] = c;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
    }
    // ---- my points: sorted positions [tid*c, tid*c + c), slots in ascending ORIGINAL index -----------------------
    const int c = (n + THREADS - 1) / THREADS;
    const int base = tid * c;
    const int lim = min(c, n - base);  // may be <= 0
    unsigned myidx[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) myidx[k] = (k < lim) ? (unsigned)(keys[base + k] & 0xffffffffull) : 0xffffffffu;
#pragma unroll
    for (int pass = 0; pass < PPT; ++pass) {  // odd-even transposition sort, static indices only
#pragma unroll
        for (int k = pass & 1; k + 1 < PPT; k += 2) {
            const unsigned lo = min(myidx[k], myidx[k + 1]), hi = max(myidx[k], myidx[k + 1]);
            myidx[k] = lo;
            myidx[k + 1] = hi;
        }
    }
    __syncthreads();  // everybody has read its keys: the buffer becomes perm[]
    unsigned *perm = reinterpret_cast<unsigned *>(fps_dyn);
    u64 X[PPT / 2], Y[PPT / 2], Z[PPT / 2];
    float D[PPT];
    float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
#pragma unroll
    for (int j = 0; j < PPT / 2; ++j) {
        float xs[2], ys[2], zs[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 2 * j + h;
            const bool ok = myidx[k] != 0xffffffffu;
            const float *q = gpos + 3 * (int64_t)(ok ? myidx[k] : 0u);
            xs[h] = ok ? __ldg(q + 0) : 0.f;
            ys[h] = ok ? __ldg(q + 1) : 0.f;
            zs[h] = ok ? __ldg(q + 2) : 0.f;
            D[k] = ok ? __int_as_float(0x7f800000) : -1.f;  // +inf: the first update sets dist-to-start
            if (k < c && base + k < THREADS * c) perm[base + k] = myidx[k];
            if (ok) {
                lox = fminf(lox, xs[h]); hix = fmaxf(hix, xs[h]);
                loy = fminf(loy, ys[h]); hiy = fmaxf(hiy, ys[h]);
                loz = fminf(loz, zs[h]); hiz = fmaxf(hiz, zs[h]);
            }
        }
        X[j] = pack2(xs[0], xs[1]);
        Y[j] = pack2(ys[0], ys[1]);
        Z[j] = pack2(zs[0], zs[1]);
    }
    __syncthreads();  // perm[] complete

    // ---- start point ---------------------------------------------------------------------------------------------
    int cur = 0;
    if (p.start != nullptr) {
        const int64_t s = p.start[b];
        cur = (s >= 0 && s < n) ? (int)s : 0;
    }
    float cx = __ldg(gpos + 3 * cur + 0), cy = __ldg(gpos + 3 * cur + 1), cz = __ldg(gpos + 3 * cur + 2);
    if (tid == 0) {
        p.out_idx[o0] = p0 + cur;
        if (p.out_pos) {
            p.out_pos[3 * o0 + 0] = cx;
            p.out_pos[3 * o0 + 1] = cy;
            p.out_pos[3 * o0 + 2] = cz;
        }
    }
    if (p.out_batch) {
#pragma unroll 1
        for (int i = tid; i < m; i += THREADS) p.out_batch[o0 + i] = b;
    }

    float lmax = lim > 0 ? __int_as_float(0x7f800000) : -1.f;  // my running maximum (cached while the warp skips)
    int wmax = __float_as_int(-1.f);                           // the warp's (cached)
    int wlane = 0;                                             // lane holding it, ties resolved by original index
    bool first = true;

    for (int it = 1; it < m; ++it) {
        // 1. can the new sample lower any of my distances?  box distance in the arithmetic of the point distances
        const float ex = fmaxf(fmaxf(__fsub_rn(lox, cx), __fsub_rn(cx, hix)), 0.f);
        const float ey = fmaxf(fmaxf(__fsub_rn(loy, cy), __fsub_rn(cy, hiy)), 0.f);
        const float ez = fmaxf(fmaxf(__fsub_rn(loz, cz), __fsub_rn(cz, hiz)), 0.f);
        const float bd = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
        if (__any_sync(0xffffffffu, bd < lmax) || first) {  // warp-uniform; empty threads: box = (+inf,-inf) -> NaN/inf
            first = false;
            const u64 px = pack2(cx, cx), py = pack2(cy, cy), pz = pack2(cz, cz);
            float lm = -1.f;
#pragma unroll
            for (int j = 0; j < PPT / 2; ++j) {
                float d0, d1;
                dist2_pair(X[j], Y[j], Z[j], px, py, pz, d0, d1);
                D[2 * j] = fminf(D[2 * j], d0);
                D[2 * j + 1] = fminf(D[2 * j + 1], d1);
                lm = fmaxf(lm, fmaxf(D[2 * j], D[2 * j + 1]));
            }
            lmax = lm;
            const int lb = __float_as_int(lm);
            wmax = __reduce_max_sync(0xffffffffu, lb);
            const unsigned cand = __ballot_sync(0xffffffffu, lb == wmax);
            wlane = __ffs(cand) - 1;
            if (cand & (cand - 1u)) {  // several lanes tie: lowest original index wins (rare)
                unsigned oi = 0xffffffffu;
                if (lb == wmax) {
#pragma unroll
                    for (int k = PPT - 1; k >= 0; --k)
                        if (__float_as_int(D[k]) == wmax) oi = perm[base + k];  // ends at the lowest k = lowest index
                }
                const unsigned best = __reduce_min_sync(0xffffffffu, oi);
                wlane = __ffs(__ballot_sync(0xffffffffu, oi == best)) - 1;
            }
            if (lane == 0) s_wmax[warp] = wmax;
        }
        __syncthreads();
        // 2. block arg-max, computed redundantly by every warp
        const int v = (lane < NW) ? s_wmax[lane] : INT_MIN;
        const int cmax = __reduce_max_sync(0xffffffffu, v);
        const unsigned cw = __ballot_sync(0xffffffffu, v == cmax);
        int wwarp = __ffs(cw) - 1;
        if (cw & (cw - 1u)) {  // several warps tie (block-uniform, rare): compare their candidates' original indices
            if ((cw >> warp) & 1u) {
                if (lane == wlane) {
                    unsigned oi = 0xffffffffu;
#pragma unroll
                    for (int k = PPT - 1; k >= 0; --k)
                        if (__float_as_int(D[k]) == cmax) oi = perm[base + k];
                    s_cand[warp] = oi;
                }
            }
            __syncthreads();
            const unsigned oi = (lane < NW && ((cw >> lane) & 1u)) ? s_cand[lane] : 0xffffffffu;
            const unsigned best = __reduce_min_sync(0xffffffffu, oi);
            wwarp = __ffs(__ballot_sync(0xffffffffu, oi == best)) - 1;
        }
        const int par = it & 1;
        if (warp == wwarp && lane == wlane) {
            int ksel = 0;
            float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
            for (int k = PPT - 1; k >= 0; --k) {
                if (__float_as_int(D[k]) == cmax) {
                    ksel = k;
                    wx = (k & 1) ? hi32(X[k >> 1]) : lo32(X[k >> 1]);
                    wy = (k & 1) ? hi32(Y[k >> 1]) : lo32(Y[k >> 1]);
                    wz = (k & 1) ? hi32(Z[k >> 1]) : lo32(Z[k >> 1]);
                }
            }
            const unsigned widx = perm[base + ksel];
            *reinterpret_cast<uint4 *>(&s_rec[par][0]) =
                make_uint4((unsigned)cmax, widx, __float_as_uint(wx), __float_as_uint(wy));
            s_rec[par][4] = __float_as_uint(wz);
            p.out_idx[o0 + it] = p0 + (int64_t)widx;
            if (p.out_pos) {
                p.out_pos[3 * (o0 + it) + 0] = wx;
                p.out_pos[3 * (o0 + it) + 1] = wy;
                p.out_pos[3 * (o0 + it) + 2] = wz;
            }
        }
        __syncthreads();
        const uint4 r = *reinterpret_cast<const uint4 *>(&s_rec[par][0]);
        cx = __uint_as_float(r.z);
        cy = __uint_as_float(r.w);
        cz = __uint_as_float(s_rec[par][4]);
    }
}

template <int THREADS, int PPT>
static int launch_fps_sorted(const FpsParams &p, int B, int64_t max_n, cudaStream_t stream)
{
    auto kern = fps_sorted_kernel<THREADS, PPT>;
    const int npad = next_pow2((int)max_n);
    const int c = (int)((max_n + THREADS - 1) / THREADS);
    const int smem = npad * 8 > THREADS * c * 4 ? npad * 8 : THREADS * c * 4;  // sort keys, then perm[] in the same bytes
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)B, THREADS, smem, stream>>>(p, npad);
    note_launch();
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// capacity: 16384 sort keys of 8 B in shared memory
constexpr int64_t SORTED_MAX_N = 16384;

template <int THREADS>
static int dispatch_sorted_ppt(const FpsParams &p, int B, int64_t max_n, cudaStream_t stream)
{
    constexpr int MAXP = THREADS >= 1024 ? 8 : (THREADS >= 768 ? 14 : (THREADS >= 640 ? 18 : (THREADS >= 512 ? 20 : 32)));
    const int c = (int)((max_n + THREADS - 1) / THREADS);
    if (c <= 2) return launch_fps_sorted<THREADS, 2>(p, B, max_n, stream);
    if (c <= 4) return launch_fps_sorted<THREADS, 4>(p, B, max_n, stream);
    if (c <= 8) return launch_fps_sorted<THREADS, 8>(p, B, max_n, stream);
    if (MAXP >= 12 && c <= 12) return launch_fps_sorted<THREADS, (MAXP >= 12 ? 12 : 2)>(p, B, max_n, stream);
    if (MAXP >= 14 && c <= 14) return launch_fps_sorted<THREADS, (MAXP >= 14 ? 14 : 2)>(p, B, max_n, stream);
    if (MAXP >= 16 && c <= 16) return launch_fps_sorted<THREADS, (MAXP >= 16 ? 16 : 2)>(p, B, max_n, stream);
    if (MAXP >= 18 && c <= 18) return launch_fps_sorted<THREADS, (MAXP >= 18 ? 18 : 2)>(p, B, max_n, stream);
    if (MAXP >= 20 && c <= 20) return launch_fps_sorted<THREADS, (MAXP >= 20 ? 20 : 2)>(p, B, max_n, stream);
    if (MAXP >= 32 && c <= 32) return launch_fps_sorted<THREADS, (MAXP >= 32 ? 32 : 2)>(p, B, max_n, stream);
    return B2PN_ENOTSUP;
}

static int dispatch_sorted(const FpsParams &p, int B, int64_t max_n, int threads, cudaStream_t stream)
{
    if (max_n > SORTED_MAX_N) return B2PN_ENOTSUP;
    if (threads == 0) threads = max_n <= 2048 ? 256 : (max_n <= 8192 ? 512 : 640);
    switch (threads) {
        case 256: return dispatch_sorted_ppt<256>(p, B, max_n, stream);
        case 512: return dispatch_sorted_ppt<512>(p, B, max_n, stream);
        case 640: return dispatch_sorted_ppt<640>(p, B, max_n, stream);
        case 768: return dispatch_sorted_ppt<768>(p, B, max_n, stream);
        case 1024: return dispatch_sorted_ppt<1024>(p, B, max_n, stream);
    }
    return B2PN_EINVAL;
}

template <int CLUSTER, int THREADS, int PPT>
static int launch_fps(const FpsParams &p, int B, cudaStream_t stream)
{
    auto kern = fps_kernel<CLUSTER, THREADS, PPT>;
    if (CLUSTER > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CLUSTER));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    note_launch();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int CLUSTER, int THREADS>
static int dispatch_ppt(const FpsParams &p, int B, int c, cudaStream_t stream)
{
    // register budget: 65536 / THREADS per thread; 4 registers per point + ~28 of control
    constexpr int MAXP = THREADS >= 1024 ? 8 : (THREADS >= 512 ? 24 : 48);
    if (c <= 2) return launch_fps<CLUSTER, THREADS, 2>(p, B, stream);
    if (c <= 4) return launch_fps<CLUSTER, THREADS, 4>(p, B, stream);
    if (c <= 6) return launch_fps<CLUSTER, THREADS, 6>(p, B, stream);
    if (c <= 8) return launch_fps<CLUSTER, THREADS, 8>(p, B, stream);
    if (MAXP >= 12 && c <= 12) return launch_fps<CLUSTER, THREADS, (MAXP >= 12 ? 12 : 8)>(p, B, stream);
    if (MAXP >= 16 && c <= 16) return launch_fps<CLUSTER, THREADS, (MAXP >= 16 ? 16 : 8)>(p, B, stream);
    if (MAXP >= 20 && c <= 20) return launch_fps<CLUSTER, THREADS, (MAXP >= 20 ? 20 : 8)>(p, B, stream);
    if (MAXP >= 24 && c <= 24) return launch_fps<CLUSTER, THREADS, (MAXP >= 24 ? 24 : 8)>(p, B, stream);
    if (MAXP >= 32 && c <= 32) return launch_fps<CLUSTER, THREADS, (MAXP >= 32 ? 32 : 8)>(p, B, stream);
    if (MAXP >= 40 && c <= 40) return launch_fps<CLUSTER, THREADS, (MAXP >= 40 ? 40 : 8)>(p, B, stream);
    if (MAXP >= 48 && c <= 48) return launch_fps<CLUSTER, THREADS, (MAXP >= 48 ? 48 : 8)>(p, B, stream);
    return B2PN_ENOTSUP;
}

static int max_ppt(int threads) { return threads >= 1024 ? 8 : (threads >= 512 ? 24 : 48); }

template <int CLUSTER>
static int dispatch_threads(const FpsParams &p, int B, int64_t max_n, int threads, cudaStream_t stream)
{
    const int c = (int)((max_n + (int64_t)CLUSTER * threads - 1) / ((int64_t)CLUSTER * threads));
    switch (threads) {
        case 256: return dispatch_ppt<CLUSTER, 256>(p, B, c, stream);
        case 512: return dispatch_ppt<CLUSTER, 512>(p, B, c, stream);
        case 1024: return dispatch_ppt<CLUSTER, 1024>(p, B, c, stream);
    }
    return B2PN_EINVAL;
}

static int dispatch_cluster(const FpsParams &p, int B, int64_t max_n, int cluster, int threads, cudaStream_t stream)
{
    switch (cluster) {
        case 1: return dispatch_threads<1>(p, B, max_n, threads, stream);
        case 2: return dispatch_threads<2>(p, B, max_n, threads, stream);
        case 4: return dispatch_threads<4>(p, B, max_n, threads, stream);
        case 8: return dispatch_threads<8>(p, B, max_n, threads, stream);
        case 16: return dispatch_threads<16>(p, B, max_n, threads, stream);
    }
    return B2PN_EINVAL;
}

}  // namespace b2pn

extern "C" int64_t b2pn_fps_random_start(uint64_t seed, int64_t call, int32_t cloud, int64_t n)
{
    if (n <= 0 || n > 0x7fffffff) return B2PN_EINVAL;
    return (int64_t)b2pn::fps_random_start((unsigned long long)seed, (long long)call, (int)cloud, (int)n);
}

extern "C" int b2pn_fps_f32(const float *pos, const int64_t *ptr, const int64_t *out_ptr, const int64_t *start,
                            int32_t B, int64_t max_n, int64_t *out_idx, float *out_pos, int64_t *out_batch,
                            const b2pn_fps_options *opts, b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0 || max_n < 0) return B2PN_EINVAL;
    int threads = opts ? opts->threads : 0, cluster = opts ? opts->cluster : 0;
    if (!(cluster == -2 || cluster == -1 || cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8 || cluster == 16))
        return B2PN_EINVAL;
    if (!(threads == 0 || threads == 256 || threads == 512 || threads == 640 || threads == 768 || threads == 1024))
        return B2PN_EINVAL;
    if (B == 0 || max_n == 0) return B2PN_OK;
    if (!pos || !ptr || !out_ptr || !out_idx) return B2PN_EINVAL;
    int64_t *rng_state = (opts && !start) ? opts->rng_state : nullptr;
    FpsParams p = {pos, ptr, out_ptr, start, out_idx, out_pos, out_batch, opts ? (unsigned long long)opts->seed : 0ull, rng_state};

    // cluster == -1: the spatially pruned kernel (bit-identical results).  Measured on B200
    // (profiles/r01_fps_experiments.md) it prunes ~75 % of the scan but its serial chain of uniform-datapath
    // reductions (CREDUX / VOTE / FLO at 50-70 cycles each) makes an iteration SLOWER than the plain register scan
    // at 10k points, so it is opt-in and the register scan below stays the default.
    if (cluster < 0 && rng_state) return B2PN_ENOTSUP;  // the opt-in variants take explicit starts only
    if (cluster == -1) return dispatch_pruned(p, B, max_n, threads, (cudaStream_t)stream);
    if (cluster == -2) return dispatch_sorted(p, B, max_n, threads, (cudaStream_t)stream);
    if (threads == 0) threads = 512;
    if (cluster == 0) {
        // Measured on B200 (profiles/r01_fps_sweep.md): barrier.cluster costs more per iteration than
        // the scan it saves up to ~12k points, so use the smallest cluster whose registers hold the
        // largest cloud (512 threads x 24 points per CTA).
        cluster = 16;
        for (int cl = 1; cl <= 16; cl *= 2) {
            if ((int64_t)cl * threads * max_ppt(threads) >= max_n) {
                cluster = cl;
                break;
            }
        }
    }
    if ((int64_t)cluster * threads * max_ppt(threads) < max_n) return B2PN_ENOTSUP;
    return dispatch_cluster(p, B, max_n, cluster, threads, (cudaStream_t)stream);
}
