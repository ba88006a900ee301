// Offline resampler -- farthest-point sampling in float64, the drop-in for the reference's numpy
// `farthest_point_sampling(coords, k)` (/root/reference/downsampling_point_clouds.py:55-92), which prepares the
// 7 168-point training clouds from raw plots (called at :153).  SURVEY.md section 8 row f1.
//
// Semantics, bit for bit those of the numpy code: distances ((cx-x)^2 + (cy-y)^2) + (cz-z)^2 in separately rounded
// float64 (raw UTM coordinates: float32 would not do), running minimum, first arg-max, start at point 0, and a selected
// point never competes again (np.delete) -- only visible when duplicates exhaust the cloud.
//
// One CTA per plot; points and running distances stay in global memory (L2 resident), so many plots are resampled
// concurrently -- the dataset holds thousands of plots, the parallelism is across them.  An iteration streams the plot
// once: 40 B per point at the L2 -> SM bandwidth of one SM.
#include <math.h>

#include "common.cuh"

namespace b2pn {

constexpr int F64_THREADS = 1024;

struct FpsF64Params {
    const double *pos;
    const int64_t *ptr;
    const int64_t *out_ptr;
    const int64_t *start;
    int64_t *out_idx;
    double *dist;  // [N] workspace
};

__global__ void __launch_bounds__(F64_THREADS, 1) fps_f64_kernel(const FpsF64Params p)
{
    __shared__ double s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_cur;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t p0 = p.ptr[b];
    const int n = (int)(p.ptr[b + 1] - p0);
    const int64_t o0 = p.out_ptr[b];
    const int m = (int)(p.out_ptr[b + 1] - o0);
    if (n <= 0 || m <= 0) return;
    const double *g = p.pos + 3 * p0;
    double *dist = p.dist + p0;
    int cur = 0;
    if (p.start != nullptr) {
        const int64_t s = p.start[b];
        cur = (s >= 0 && s < n) ? (int)s : 0;
    }
    for (int j = tid; j < n; j += F64_THREADS) dist[j] = j == cur ? -1.0 : INFINITY;  // -1: selected, out of the race
    if (tid == 0) p.out_idx[o0] = p0 + cur;
    __syncthreads();

    for (int it = 1; it < m; ++it) {
        const double cx = g[3 * cur], cy = g[3 * cur + 1], cz = g[3 * cur + 2];
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int j = tid; j < n; j += F64_THREADS) {  // ascending j: strict '>' keeps my lowest index
            const double dj = dist[j];
            if (dj < 0.0) continue;
            const double dx = __dsub_rn(cx, g[3 * j]), dy = __dsub_rn(cy, g[3 * j + 1]), dz = __dsub_rn(cz, g[3 * j + 2]);
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            double nd = dj;
            if (d < dj) {
                nd = d;
                dist[j] = d;
            }
            if (nd > bv) {
                bv = nd;
                bi = j;
            }
        }
        // arg-max, ties to the lowest index
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (lane == 0) {
            s_v[warp] = bv;
            s_i[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            bv = s_v[lane];
            bi = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) {
                    bv = ov;
                    bi = oi;
                }
            }
            if (lane == 0) {
                s_cur = bi;
                dist[bi] = -1.0;
                p.out_idx[o0 + it] = p0 + bi;
            }
        }
        __syncthreads();
        cur = s_cur;
    }
}

}  // namespace b2pn

extern "C" int b2pn_fps_f64(const double *pos, const int64_t *ptr, const int64_t *out_ptr, const int64_t *start, int32_t B,
                            int64_t *out_idx, double *dist_workspace, b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0) return B2PN_EINVAL;
    if (B == 0) return B2PN_OK;
    if (!pos || !ptr || !out_ptr || !out_idx || !dist_workspace) return B2PN_EINVAL;
    FpsF64Params p = {pos, ptr, out_ptr, start, out_idx, dist_workspace};
    fps_f64_kernel<<<(unsigned)B, F64_THREADS, 0, (cudaStream_t)stream>>>(p);
    note_launch();
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}
