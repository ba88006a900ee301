/*
 * b2pn CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (dl_biomass_b200/) never does.
 *
 * It restates, in plain C, the grouping arithmetic the reference reaches through
 *   /root/reference/pointnet2_regressor.py:13     fps(pos, batch, ratio)
 *   /root/reference/pointnet2_regressor.py:14-15  radius(pos, pos[idx], r, batch, batch[idx], 64)
 * Those calls land in torch_cluster (un-vendored, unpinned; era 1.6.0) -- see SURVEY.md
 * Appendix A.1 / A.2 for the published semantics restated here.  The only in-repo
 * statement of either primitive is the numpy FPS at
 *   /root/reference/downsampling_point_clouds.py:55-92
 * and tests/golden/fps_reference_numpy.npz (made by oracle/gen_golden.py from that
 * function, run in place) pins oracle_fps_f32/f64 against it.  The ball query has no
 * runnable reference here: PARITY UNPINNED for it (canonical rule = torch_cluster's
 * CUDA kernel: first K sources by ascending index with d2 < r2, strict).
 *
 * Arithmetic contract, shared with the CUDA kernels:
 *   d2 = ((dx*dx + dy*dy) + dz*dz), every op separately rounded in fp32, NO FMA
 *   (build with -ffp-contract=off), dx = p_src - p_ref.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#if defined(__FAST_MATH__) || (defined(__FP_FAST_FMAF) && !defined(B2PN_ORACLE_NO_CONTRACT))
#error "build the oracle with -ffp-contract=off -DB2PN_ORACLE_NO_CONTRACT and without -ffast-math (see oracle/Makefile)"
#endif

/* ---- tiny pthread parallel-for (no libgomp in the image) ---------------------------- */
typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t n, chunk; int64_t *next; pthread_mutex_t *mu; } pf_task;

static void *pf_worker(void *arg)
{
    pf_task *t = (pf_task *)arg;
    for (;;) {
        pthread_mutex_lock(t->mu);
        int64_t lo = *t->next;
        *t->next = lo + t->chunk;
        pthread_mutex_unlock(t->mu);
        if (lo >= t->n) break;
        int64_t hi = lo + t->chunk < t->n ? lo + t->chunk : t->n;
        t->fn(lo, hi, t->ctx);
    }
    return NULL;
}

int oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(int64_t n, int64_t chunk, int threads, range_fn fn, void *ctx)
{
    if (threads <= 0) threads = oracle_num_threads();
    if (threads > 256) threads = 256;
    if (threads == 1 || n <= chunk) { fn(0, n, ctx); return; }
    pthread_t th[256];
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    int64_t next = 0;
    pf_task t = { fn, ctx, n, chunk, &next, &mu };
    int started = 0;
    for (int i = 0; i < threads - 1; ++i)
        if (pthread_create(&th[started], NULL, pf_worker, &t) == 0) ++started;
    pf_worker(&t);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

static inline float d2_f32(const float *a, const float *b)
{
    /* torch CPU: ((y - y[a])**2).sum(1) -> sub, mul, left-to-right add (A.1) */
    const float dx = a[0] - b[0];
    const float dy = a[1] - b[1];
    const float dz = a[2] - b[2];
    const float qx = dx * dx;
    const float qy = dy * dy;
    const float qz = dz * dz;
    const float s = qx + qy;
    return s + qz;
}

/* number of samples for a cloud of n points: ceil(float32(n) * float32(ratio))  (A.1) */
int64_t oracle_fps_num_samples(int64_t n, float ratio)
{
    const float prod = (float)n * ratio;
    return (int64_t)ceilf(prod);
}

/*
 * Farthest-point sampling, torch_cluster fps_cpu semantics with an explicit start.
 *   pos      [N,3] fp32 row-major, clouds concatenated
 *   ptr      [B+1] cloud offsets into pos
 *   out_ptr  [B+1] offsets into out_idx (out_ptr[b+1]-out_ptr[b] = samples of cloud b)
 *   start    [B]   cloud-local start index (NULL -> 0)
 *   out_idx  [M]   GLOBAL indices, selection order, first = start point
 * dist is initialised from the start point; argmax ties -> lowest index.
 */
typedef struct { const float *pos; const int64_t *ptr, *out_ptr, *start; int64_t *out_idx; } fps_ctx;

static void fps_range(int64_t b0, int64_t b1, void *vctx)
{
    const fps_ctx *c_ = (const fps_ctx *)vctx;
    const float *pos = c_->pos; const int64_t *ptr = c_->ptr, *out_ptr = c_->out_ptr, *start = c_->start;
    int64_t *out_idx = c_->out_idx;
    for (int64_t b = b0; b < b1; ++b) {
        const int64_t n = ptr[b + 1] - ptr[b];
        const int64_t m = out_ptr[b + 1] - out_ptr[b];
        if (n <= 0 || m <= 0) continue;
        const float *p = pos + 3 * ptr[b];
        int64_t *out = out_idx + out_ptr[b];
        float *dist = (float *)malloc(sizeof(float) * (size_t)n);
        int64_t cur = start ? start[b] : 0;
        if (cur < 0 || cur >= n) cur = 0;
        out[0] = ptr[b] + cur;
        for (int64_t j = 0; j < n; ++j) dist[j] = d2_f32(p + 3 * j, p + 3 * cur);
        for (int64_t i = 1; i < m; ++i) {
            int64_t best = 0;
            float bestd = dist[0];
            for (int64_t j = 1; j < n; ++j)
                if (dist[j] > bestd) { bestd = dist[j]; best = j; }
            cur = best;
            out[i] = ptr[b] + cur;
            const float *c = p + 3 * cur;
            for (int64_t j = 0; j < n; ++j) {
                float d = d2_f32(p + 3 * j, c);
                if (d < dist[j]) dist[j] = d;
            }
        }
        free(dist);
    }
}

int oracle_fps_f32(const float *pos, const int64_t *ptr, const int64_t *out_ptr,
                   const int64_t *start, int32_t B, int64_t *out_idx, int32_t threads)
{
    if (!pos || !ptr || !out_ptr || !out_idx || B < 0) return -1;
    fps_ctx c = { pos, ptr, out_ptr, start, out_idx };
    parallel_for(B, 1, threads, fps_range, &c);   /* clouds in parallel, like at::parallel_for */
    return 0;
}

/* float64 twin of the reference's numpy farthest_point_sampling (/root/reference/downsampling_point_clouds.py:55-92):
 * distances ((cx-x)^2 + (cy-y)^2) + (cz-z)^2 in separately rounded float64, running minimum, first arg-max, and -- like
 * the reference's np.delete(idx, selected) -- a selected point never competes again (only visible when duplicates
 * exhaust the cloud: the next pick is then the first UNSELECTED index, not the first index with distance 0). */
int oracle_fps_f64(const double *pos, int64_t n, int64_t m, int64_t start, int64_t *out)
{
    if (!pos || !out || n <= 0 || m <= 0 || m > n) return -1;
    double *dist = (double *)malloc(sizeof(double) * (size_t)n);
    int64_t cur = start;
    out[0] = cur;
    for (int64_t j = 0; j < n; ++j) dist[j] = INFINITY;
    dist[cur] = -1.0; /* selected */
    for (int64_t i = 1; i < m; ++i) {
        const double *c = pos + 3 * cur;
        int64_t best = -1;
        double bestd = -1.0;
        for (int64_t j = 0; j < n; ++j) {
            if (dist[j] < 0.0) continue;
            const double dx = c[0] - pos[3 * j], dy = c[1] - pos[3 * j + 1], dz = c[2] - pos[3 * j + 2];
            const double qx = dx * dx, qy = dy * dy, qz = dz * dz;
            const double s = qx + qy;
            const double d = s + qz;
            if (d < dist[j]) dist[j] = d;
            if (dist[j] > bestd) { bestd = dist[j]; best = j; }
        }
        cur = best;
        out[i] = cur;
        dist[cur] = -1.0;
    }
    free(dist);
    return 0;
}

/*
 * Ball query, torch_cluster radius_cuda semantics (A.2): for each query (centroid) scan the
 * sources of the same cloud in ascending index, keep the first K with d2 < r2 (strict).
 *   nbr [M,K] int32  GLOBAL source indices, slots >= cnt filled with -1
 *   cnt [M]   int32
 * r2 = (float)((double)r * (double)r).
 */
typedef struct { const float *src, *qry; int64_t s0, s1, q0; float r2; int32_t K; int32_t *nbr, *cnt; } bq_ctx;

static void bq_range(int64_t lo, int64_t hi, void *vctx)
{
    const bq_ctx *c_ = (const bq_ctx *)vctx;
    for (int64_t q = c_->q0 + lo; q < c_->q0 + hi; ++q) {
        int32_t c = 0;
        int32_t *row = c_->nbr + (size_t)q * c_->K;
        for (int64_t j = c_->s0; j < c_->s1 && c < c_->K; ++j)
            if (d2_f32(c_->src + 3 * j, c_->qry + 3 * q) < c_->r2) row[c++] = (int32_t)j;
        c_->cnt[q] = c;
        for (int32_t k = c; k < c_->K; ++k) row[k] = -1;
    }
}

int oracle_ball_query_f32(const float *src, const float *qry, const int64_t *src_ptr,
                          const int64_t *qry_ptr, int32_t B, double r, int32_t K,
                          int32_t *nbr, int32_t *cnt, int32_t threads)
{
    if (!src || !qry || !src_ptr || !qry_ptr || !nbr || !cnt || B < 0 || K <= 0) return -1;
    const float r2 = (float)(r * r);
    for (int32_t b = 0; b < B; ++b) {
        bq_ctx c = { src, qry, src_ptr[b], src_ptr[b + 1], qry_ptr[b], r2, K, nbr, cnt };
        parallel_for(qry_ptr[b + 1] - qry_ptr[b], 64, threads, bq_range, &c);
    }
    return 0;
}
