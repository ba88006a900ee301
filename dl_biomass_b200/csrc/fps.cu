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
};

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
    if (n <= 0 || m <= 0) return;  // uniform over the cluster

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
}

static int g_force_cluster = 0;
static int g_force_threads = 0;

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

extern "C" int b2pn_fps_set_variant(int32_t cluster, int32_t threads)
{
    if (!(cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8 || cluster == 16))
        return B2PN_EINVAL;
    if (!(threads == 0 || threads == 256 || threads == 512 || threads == 1024)) return B2PN_EINVAL;
    b2pn::g_force_cluster = cluster;
    b2pn::g_force_threads = threads;
    return B2PN_OK;
}

extern "C" int b2pn_fps_f32(const float *pos, const int64_t *ptr, const int64_t *out_ptr, const int64_t *start,
                            int32_t B, int64_t max_n, int64_t *out_idx, float *out_pos, int64_t *out_batch,
                            b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0 || max_n < 0) return B2PN_EINVAL;
    if (B == 0 || max_n == 0) return B2PN_OK;
    if (!pos || !ptr || !out_ptr || !out_idx) return B2PN_EINVAL;
    FpsParams p = {pos, ptr, out_ptr, start, out_idx, out_pos, out_batch};

    int threads = g_force_threads, cluster = g_force_cluster;
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
