// Training-time augmentation on the device, written straight into the ragged batch layout the model takes.
//
// Restates /root/reference/augmentation.py:54-122 as used by AugmentPointCloudsInFiles.__getitem__ (:287-289):
//   point_removal  (:73-89)   keep a uniformly random subset of round(0.9 n)..n points, in random order
//   random_noise   (:92-122)  jitter every kept point by +-N(0, sd) (coordinates and attributes), pick 0..round(0.1 n')
//                             distinct jittered points and APPEND them after the kept ones
//   rotate_points  (:54-70)   coords @ [[c,-s,0],[s,c,0],[0,0,1]]  ->  x' = x c + y s,  y' = y c - x s,  z' = z
// The reference does this per sample in numpy on the loader's CPU workers; here one CTA handles one cloud of a
// resident cloud cache.  The scalar draws (subset sizes, sd and sign, angle) come from the host -- they fix the batch
// layout -- the per-point draws from a counter-based generator (splitmix64 finaliser of (seed, cloud uid, stream,
// counter), restated in oracle/augment_ref.py), so a batch is a pure function of (cache, parameters, seed):
//   stream 0: sort key of source point i      -> random order; the first n_keep are kept
//   stream 1: sort key of kept position j     -> the first n_dup positions are duplicated
//   stream 2: normal deviate (Box-Muller) of kept position j, component d (x, y, z, then attributes)
#include "common.cuh"

namespace b2pn {

constexpr int AUG_THREADS = 1024;
constexpr int AUG_CHUNK = 64;  // clouds per launch: their parameters travel in the kernel's parameter space

struct AugChunk {
    b2pn_augment_cloud c[AUG_CHUNK];
};

__host__ __device__ __forceinline__ uint64_t aug_mix(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}
__host__ __device__ __forceinline__ uint64_t aug_stream(uint64_t seed, uint64_t uid, uint64_t stream)
{
    const uint64_t a = aug_mix(seed + 0x9E3779B97F4A7C15ull * (uid + 1));
    return aug_mix(a + 0xD1B54A32D192ED03ull * (stream + 1));
}
__host__ __device__ __forceinline__ uint64_t aug_draw(uint64_t stream_state, uint64_t ctr)
{
    return aug_mix(stream_state + 0x9E3779B97F4A7C15ull * (ctr + 1));
}

__device__ __forceinline__ float aug_normal(uint64_t r)
{
    const float u1 = (float)((uint32_t)(r >> 40) + 1u) * 5.9604644775390625e-08f;   // (0, 1]
    const float u2 = (float)((uint32_t)(r >> 16) & 0xFFFFFFu) * 5.9604644775390625e-08f;  // [0, 1)
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// ascending bitonic sort of keys[0..n_pad) (n_pad a power of two) by the whole block
__device__ void aug_sort(uint64_t *keys, int n_pad)
{
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
                const int l = i | j;
                const uint64_t a = keys[i], b = keys[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int aug_pow2(int n)
{
    int p = 2;
    while (p < n) p <<= 1;
    return p;
}

__global__ void __launch_bounds__(AUG_THREADS) augment_kernel(const float *__restrict__ pos, const float *__restrict__ x, int F,
                                                              uint64_t seed, int batch_base, float *__restrict__ out_pos,
                                                              float *__restrict__ out_x, int64_t *__restrict__ out_batch,
                                                              int32_t *__restrict__ out_src, const __grid_constant__ AugChunk ch)
{
    extern __shared__ __align__(16) unsigned char aug_smem[];
    const b2pn_augment_cloud &c = ch.c[blockIdx.x];
    const int n = c.n_src, nk = c.n_keep, nd = c.n_dup;
    if (n <= 0 || nk <= 0) return;
    const int n_pad = aug_pow2(n);
    uint64_t *keys = reinterpret_cast<uint64_t *>(aug_smem);
    int32_t *perm = reinterpret_cast<int32_t *>(keys + n_pad);
    const float *sp = pos + c.src_off * 3;
    const float *sx = x ? x + c.src_off * F : nullptr;
    float *op = out_pos + c.out_off * 3;
    float *ox = out_x ? out_x + c.out_off * F : nullptr;
    const float ca = c.cos_a, sa = c.sin_a;

    // ---- point_removal: random order, first n_keep
    const uint64_t s0 = aug_stream(seed, c.uid, 0);
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
        keys[i] = i < n ? ((aug_draw(s0, (uint64_t)i) >> 32) << 32) | (uint32_t)i : ~0ull;
    __syncthreads();
    aug_sort(keys, n_pad);
    for (int j = threadIdx.x; j < nk; j += blockDim.x) perm[j] = (int32_t)(uint32_t)keys[j];
    __syncthreads();

    // ---- kept points: rotate and write
    for (int j = threadIdx.x; j < nk; j += blockDim.x) {
        const int p = perm[j];
        const float px = sp[3 * p], py = sp[3 * p + 1], pz = sp[3 * p + 2];
        op[3 * j] = __fadd_rn(__fmul_rn(px, ca), __fmul_rn(py, sa));
        op[3 * j + 1] = __fsub_rn(__fmul_rn(py, ca), __fmul_rn(px, sa));
        op[3 * j + 2] = pz;
        for (int f = 0; f < F; ++f) ox[(int64_t)j * F + f] = sx[(int64_t)p * F + f];
        if (out_batch) out_batch[c.out_off + j] = batch_base + (int)blockIdx.x;
        if (out_src) out_src[c.out_off + j] = p;
    }
    if (nd <= 0) return;

    // ---- random_noise: choose n_dup kept positions, jitter, append
    const int k_pad = aug_pow2(nk);
    const uint64_t s1 = aug_stream(seed, c.uid, 1), s2 = aug_stream(seed, c.uid, 2);
    __syncthreads();
    for (int j = threadIdx.x; j < k_pad; j += blockDim.x)
        keys[j] = j < nk ? ((aug_draw(s1, (uint64_t)j) >> 32) << 32) | (uint32_t)j : ~0ull;
    __syncthreads();
    aug_sort(keys, k_pad);
    const int comps = 3 + F;
    const float sd = c.noise_sd;  // signed: + adds the deviates, - subtracts them (augmentation.py:97-111)
    for (int t = threadIdx.x; t < nd; t += blockDim.x) {
        const int jj = (int)(uint32_t)keys[t];
        const int p = perm[jj];
        const uint64_t base = (uint64_t)jj * (uint64_t)comps;
        const float px = __fadd_rn(sp[3 * p], __fmul_rn(sd, aug_normal(aug_draw(s2, base))));
        const float py = __fadd_rn(sp[3 * p + 1], __fmul_rn(sd, aug_normal(aug_draw(s2, base + 1))));
        const float pz = __fadd_rn(sp[3 * p + 2], __fmul_rn(sd, aug_normal(aug_draw(s2, base + 2))));
        const int64_t o = (int64_t)nk + t;
        op[3 * o] = __fadd_rn(__fmul_rn(px, ca), __fmul_rn(py, sa));
        op[3 * o + 1] = __fsub_rn(__fmul_rn(py, ca), __fmul_rn(px, sa));
        op[3 * o + 2] = pz;
        for (int f = 0; f < F; ++f)
            ox[o * F + f] = __fadd_rn(sx[(int64_t)p * F + f], __fmul_rn(sd, aug_normal(aug_draw(s2, base + 3 + f))));
        if (out_batch) out_batch[c.out_off + o] = batch_base + (int)blockIdx.x;
        if (out_src) out_src[c.out_off + o] = p;
    }
}

}  // namespace b2pn

extern "C" int32_t b2pn_augment_max_points(void) { return B2PN_AUG_MAX_POINTS; }

extern "C" uint64_t b2pn_augment_draw(uint64_t seed, uint64_t uid, uint32_t stream, uint64_t counter)
{
    return b2pn::aug_draw(b2pn::aug_stream(seed, uid, stream), counter);
}

extern "C" int b2pn_augment_batch(const float *pos, const float *x, int32_t F, const b2pn_augment_cloud *clouds, int32_t B,
                                  uint64_t seed, float *out_pos, float *out_x, int64_t *out_batch, int32_t *out_src,
                                  b2pn_stream_t stream)
{
    using namespace b2pn;
    if (B < 0 || F < 0) return B2PN_EINVAL;
    if (B == 0) return B2PN_OK;
    if (!clouds || !pos || !out_pos || (F > 0 && (!x || !out_x))) return B2PN_EINVAL;
    for (int b = 0; b < B; ++b) {
        const b2pn_augment_cloud &c = clouds[b];
        if (c.n_src < 0 || c.n_keep < 0 || c.n_dup < 0 || c.n_keep > c.n_src || c.n_dup > c.n_keep || c.src_off < 0 ||
            c.out_off < 0)
            return B2PN_EINVAL;
        if (c.n_src > B2PN_AUG_MAX_POINTS) return B2PN_ENOTSUP;
    }
    {   // the attribute is per DEVICE: remember it per device (idempotent: a race just sets it twice)
        static bool attr_set[64] = {};
        int dev = 0;
        B2PN_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            B2PN_CUDA(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2PN_AUG_MAX_POINTS * 12));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    for (int b0 = 0; b0 < B; b0 += AUG_CHUNK) {
        AugChunk ch;
        const int nb = B - b0 < AUG_CHUNK ? B - b0 : AUG_CHUNK;
        int mx = 2;
        for (int i = 0; i < nb; ++i) {
            ch.c[i] = clouds[b0 + i];
            int p = 2;
            while (p < ch.c[i].n_src) p <<= 1;
            mx = p > mx ? p : mx;
        }
        augment_kernel<<<nb, AUG_THREADS, (size_t)mx * 12, (cudaStream_t)stream>>>(pos, x, F, seed, b0, out_pos, out_x, out_batch,
                                                                                   out_src, ch);
        note_launch();
    }
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}
