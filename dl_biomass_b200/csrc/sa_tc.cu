// Kernel 3, bf16 mode -- set-abstraction MLP on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Orientation: every layer computes the TRANSPOSED product  D^T[channel, row] = W[channel, k] * A^T[k, row]
// so that output channels sit on the 128 TMEM lanes and rows (neighbour slots / points) on the TMEM
// columns.  One epilogue thread then owns one channel: BatchNorm statistics are two registers per
// thread, the per-centroid max over its K slots is a per-thread scan of K consecutive accumulator
// columns (no shuffles, no atomics), and activations are stored FEATURE-MAJOR (hT[channel][row]) with
// 16-byte vector stores.  Feature-major activations are exactly the MN-major B operand of the next
// layer and the K-major operands of the dW GEMMs, so no transposes are ever needed.
//
// Pipeline (persistent CTAs, 9 warps): warps 4-7 build the B tile in shared memory (gather + concat,
// or BatchNorm + ReLU on load) and fetch the pre-packed weight chunk with a bulk copy (TMA unit);
// warp 8 issues tcgen05.mma; warps 0-3 drain TMEM and run the epilogue.  K is streamed in chunks of
// 64 through a ring of shared-memory stages; two TMEM accumulator buffers overlap epilogue and MMA.
#include <math.h>

#include "tc_common.cuh"

namespace b2pn {
namespace tc {

constexpr int R = 128;          // rows per tile = UMMA N
constexpr int KC = 64;          // reduction chunk = one 128-byte line of bf16
constexpr int NUM_EPI = 128;    // warps 0-3
constexpr int NUM_LOAD = 128;   // warps 4-7
constexpr int NT = 288;         // + warp 8: MMA issuer / TMEM owner
constexpr int B_BYTES = R * LINE_BYTES;  // 16 KB: [128 row lines x 64 k] or [2 row blocks][64 k lines x 64 rows]

struct RowMapTC {
    int seg_mode, K;
    const int32_t *nbr;
    const int32_t *cnt;
    const int64_t *batch;
    int64_t rows;    // logical rows
    int64_t n_dst;
    __device__ __forceinline__ bool valid(int64_t row) const
    {
        if (row >= rows) return false;
        if (seg_mode) return true;
        const int64_t m = row / K;
        return (int)(row - m * K) < cnt[m];
    }
    // number of valid rows among the 8 rows starting at row8 (row8 % 8 == 0, K % 8 == 0)
    __device__ __forceinline__ int valid8(int64_t row8) const
    {
        if (row8 >= rows) return 0;
        if (seg_mode) return (int)min((int64_t)8, rows - row8);
        const int64_t m = row8 / K;
        const int k0 = (int)(row8 - m * K);
        return max(0, min(8, cnt[m] - k0));
    }
};

struct GemmParams {
    const uint8_t *a_packed;  // [m_group][k_chunk][MT*128 lines][128 B], swizzled bf16
    int num_kc;
    int64_t num_tiles;
};

// =================================================================================================
//  B-tile loaders (128 threads).  produce() fills one 16 KB chunk for reduction chunk kc.
// =================================================================================================
struct GatherLoaderTC {  // K-major B: line = row of the tile, 64 k per line: [x_j || pos_j - pos_i || 0]
    static constexpr bool B_MN = false;
    RowMapTC rm;
    const __nv_bfloat16 *x;  // [n_src, c_in] bf16 row-major (may be null when c_in == 0)
    int c_in;
    const float *pos_src;
    const float *pos_dst;
    // per-thread, per-tile state
    bool ok;
    int64_t src;
    float d0, d1, d2;
    __device__ __forceinline__ void begin_tile(int64_t tile, int lt)
    {
        const int64_t row = tile * R + lt;
        ok = rm.valid(row);
        src = 0;
        d0 = d1 = d2 = 0.f;
        if (ok) {
            src = rm.seg_mode ? row : (int64_t)rm.nbr[row];
            d0 = pos_src[3 * src + 0];
            d1 = pos_src[3 * src + 1];
            d2 = pos_src[3 * src + 2];
            if (!rm.seg_mode) {
                const int64_t m = row / rm.K;
                d0 = __fsub_rn(d0, pos_dst[3 * m + 0]);
                d1 = __fsub_rn(d1, pos_dst[3 * m + 1]);
                d2 = __fsub_rn(d2, pos_dst[3 * m + 2]);
            }
        }
    }
    __device__ __forceinline__ float elem(int k) const
    {
        if (k < c_in) return __bfloat162float(x[src * c_in + k]);
        const int j = k - c_in;
        return j == 0 ? d0 : (j == 1 ? d1 : (j == 2 ? d2 : 0.f));
    }
    __device__ __forceinline__ void produce(uint8_t *B, int kc, int lt) const
    {
        const int k0 = kc * KC;
        const bool vec_ok = (c_in & 7) == 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int kk = k0 + c * 8;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (ok) {
                if (vec_ok && kk + 8 <= c_in) {
                    v = __ldg(reinterpret_cast<const uint4 *>(x + src * c_in + kk));
                } else if (kk < c_in + 3) {
                    v.x = pack_bf16x2(elem(kk + 0), elem(kk + 1));
                    v.y = pack_bf16x2(elem(kk + 2), elem(kk + 3));
                    v.z = pack_bf16x2(elem(kk + 4), elem(kk + 5));
                    v.w = pack_bf16x2(elem(kk + 6), elem(kk + 7));
                }
            }
            *reinterpret_cast<uint4 *>(B + line_chunk_off(lt, c)) = v;
        }
    }
    // descriptor of the 16-wide k step ks inside the chunk: K-major, SBO = 8 lines
    static __device__ __forceinline__ uint64_t b_desc(uint32_t b_saddr, int ks) { return smem_desc_sw128(b_saddr + ks * 32, 16, ATOM_BYTES); }
};

// feature-major source [C][ld] -> MN-major B tile: [2 row blocks of 64][64 channel lines]
// thread lt handles channel line (lt >> 1) of the chunk and row block (lt & 1)
template <int MODE>  // 0: plain (masked), 1: BN + act (masked with 0), 2: BN + act, invalid slots duplicate slot 0
struct FeatLoaderTC {
    static constexpr bool B_MN = true;
    RowMapTC rm;
    const __nv_bfloat16 *h;  // [C][ld]
    int C;
    int64_t ld;
    const float *scale;
    const float *shift;
    int act;
    int64_t row0;
    __device__ __forceinline__ void begin_tile(int64_t tile, int lt) { row0 = tile * R + (lt & 1) * 64; }
    __device__ __forceinline__ void produce(uint8_t *B, int kc, int lt) const
    {
        const int cl = lt >> 1, nb = lt & 1;
        const int ch = kc * KC + cl;
        uint8_t *dst = B + nb * (64 * LINE_BYTES);
        if (ch >= C) {
#pragma unroll
            for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4 *>(dst + line_chunk_off(cl, g)) = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
        const __nv_bfloat16 *srcp = h + (int64_t)ch * ld + row0;
        float sc = 1.f, sh = 0.f;
        if (MODE != 0) {
            sc = scale[ch];
            sh = shift[ch];
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int64_t r8 = row0 + g * 8;
            const int nv = rm.valid8(r8);
            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
            if (nv > 0 || MODE == 2) raw = __ldg(reinterpret_cast<const uint4 *>(srcp + g * 8));
            float f[8] = {bf16_lo(raw.x), bf16_hi(raw.x), bf16_lo(raw.y), bf16_hi(raw.y),
                          bf16_lo(raw.z), bf16_hi(raw.z), bf16_lo(raw.w), bf16_hi(raw.w)};
            float fill = 0.f;
            if (MODE == 2 && nv < 8 && !rm.seg_mode && r8 < rm.rows) {  // value of slot 0 of this centroid
                const int64_t m = r8 / rm.K;
                fill = fmaf(__bfloat162float(h[(int64_t)ch * ld + m * rm.K]), sc, sh);
                if (act == B2PN_ACT_RELU) fill = fmaxf(fill, 0.f);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float v = f[e];
                if (MODE != 0) {
                    v = fmaf(v, sc, sh);
                    if (act == B2PN_ACT_RELU) v = fmaxf(v, 0.f);
                }
                f[e] = e < nv ? v : fill;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4 *>(dst + line_chunk_off(cl, g)) = o;
        }
    }
    // MN-major: 16 k lines per step; LBO = stride between the two 64-row blocks, SBO = 8 lines
    static __device__ __forceinline__ uint64_t b_desc(uint32_t b_saddr, int ks)
    {
        return smem_desc_sw128(b_saddr + ks * (16 * LINE_BYTES), 64 * LINE_BYTES, ATOM_BYTES);
    }
};

// =================================================================================================
//  Epilogues (128 threads, thread et owns TMEM lane et = output channel within the M tile)
// =================================================================================================
struct StoreF32Ep {  // self-test: out[ch][row] = acc
    float *out;
    int C;
    int64_t ld;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void tile(uint32_t taddr, int64_t tile, int ch)
    {
#pragma unroll 1
        for (int cc = 0; cc < R / 32; ++cc) {
            float v[32];
            tmem_ld32(taddr + cc * 32, v);
            if (ch < C) {
                float4 *dst = reinterpret_cast<float4 *>(out + (int64_t)ch * ld + tile * R + cc * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    }
    __device__ __forceinline__ void finish(int mt_count, int ch_base) {}
};

template <int MT>
struct StoreStatsEpTC {  // hT[ch][row] = bf16(acc + bias); per-channel sum / sum of squares of acc
    __nv_bfloat16 *h;
    int C;
    int64_t ld;
    const float *bias;
    double *partial;  // [gridDim.x][2][cpad]
    int cpad;
    double S[MT], Q[MT];
    __device__ __forceinline__ void begin()
    {
#pragma unroll
        for (int i = 0; i < MT; ++i) S[i] = Q[i] = 0.0;
    }
    __device__ __forceinline__ void tile_mt(uint32_t taddr, int64_t tile, int ch, int mt)
    {
        const float b = ch < C ? bias[ch] : 0.f;
        float s = 0.f, q = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < R / 32; ++cc) {
            float v[32];
            tmem_ld32(taddr + cc * 32, v);
            if (ch < C) {
                uint4 *dst = reinterpret_cast<uint4 *>(h + (int64_t)ch * ld + tile * R + cc * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 o;
                    const float *p = v + 8 * j;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        s += p[e];
                        q = fmaf(p[e], p[e], q);
                    }
                    o.x = pack_bf16x2(p[0] + b, p[1] + b);
                    o.y = pack_bf16x2(p[2] + b, p[3] + b);
                    o.z = pack_bf16x2(p[4] + b, p[5] + b);
                    o.w = pack_bf16x2(p[6] + b, p[7] + b);
                    dst[j] = o;
                }
            }
        }
        S[mt] += (double)s;
        Q[mt] += (double)q;
    }
    __device__ __forceinline__ void finish_mt(int ch, int mt)
    {
        if (ch < C) {
            double *pt = partial + (int64_t)blockIdx.x * 2 * cpad;
            pt[ch] = S[mt];
            pt[cpad + ch] = Q[mt];
        }
    }
};

template <int KS>  // slots per centroid: 16, 32, 64 or 128
struct SlotMaxEpTC {  // out[m][ch] = max over the K slots (invalid slots were filled with slot 0), arg = first max slot
    __nv_bfloat16 *out;  // [n_dst][C] bf16 row-major
    int32_t *arg;
    int C;
    const float *bias;
    const int32_t *cnt;
    int64_t n_dst;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void tile_mt(uint32_t taddr, int64_t tile, int ch, int mt)
    {
        const float b = ch < C ? bias[ch] : 0.f;
        float best = -INFINITY;
        int bk = 0;
#pragma unroll 1
        for (int cc = 0; cc < R / 32; ++cc) {
            float v[32];
            tmem_ld32(taddr + cc * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = cc * 32 + j;
                const int slot = col % KS;
                if (slot == 0) {
                    best = -INFINITY;
                    bk = 0;
                }
                const float x = v[j] + b;
                if (x > best) {
                    best = x;
                    bk = slot;
                }
                if (slot == KS - 1) {
                    const int64_t m = tile * (R / KS) + col / KS;
                    if (ch < C && m < n_dst) {
                        const bool any = cnt[m] > 0;
                        out[m * C + ch] = __float2bfloat16(any ? best : 0.f);
                        arg[m * C + ch] = any ? bk : -1;
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void finish_mt(int ch, int mt) {}
};

__device__ __forceinline__ unsigned f32_orderable_tc(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct CloudMaxEpTC {  // global_max_pool over sorted cloud ids: 64-bit atomicMax keys, unpacked by a tiny kernel
    unsigned long long *keys;  // [n_dst][C], zero-initialised
    int C;
    const float *bias;
    const int64_t *batch;
    int64_t rows;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void tile_mt(uint32_t taddr, int64_t tile, int ch, int mt)
    {
        const float b = ch < C ? bias[ch] : 0.f;
        int64_t curseg = -1;
        unsigned long long best = 0ull;
#pragma unroll 1
        for (int cc = 0; cc < R / 32; ++cc) {
            float v[32];
            tmem_ld32(taddr + cc * 32, v);
            if (ch < C) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int64_t row = tile * R + cc * 32 + j;
                    if (row < rows) {
                        const int64_t sg = batch[row];
                        if (sg != curseg) {
                            if (curseg >= 0) atomicMax(keys + curseg * C + ch, best);
                            curseg = sg;
                            best = 0ull;
                        }
                        const unsigned long long key = ((unsigned long long)f32_orderable_tc(v[j] + b) << 32) |
                                                       (unsigned long long)(0xffffffffu - (unsigned)row);
                        best = key > best ? key : best;
                    }
                }
            }
        }
        if (ch < C && curseg >= 0) atomicMax(keys + curseg * C + ch, best);
    }
    __device__ __forceinline__ void finish_mt(int ch, int mt) {}
};

struct StoreF32EpW {  // adapter giving StoreF32Ep the tile_mt / finish_mt interface
    StoreF32Ep e;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void tile_mt(uint32_t taddr, int64_t tile, int ch, int mt) { e.tile(taddr, tile, ch); }
    __device__ __forceinline__ void finish_mt(int ch, int mt) {}
};

// =================================================================================================
//  The kernel
// =================================================================================================
template <int MT>
struct SmemPlan {
    static constexpr int A_BYTES = MT * 128 * LINE_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = MT == 1 ? 5 : 4;
    static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;  // barriers + alignment slack
};

template <int MT, class BL, class EP>
__global__ void __launch_bounds__(NT, 1) tc_rows_gemm_kernel(const GemmParams gp, BL bl, EP ep)
{
    using P = SmemPlan<MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + P::BAR_OFF);
    uint64_t *empty = full + P::STAGES;
    uint64_t *tfull = empty + P::STAGES;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mg = blockIdx.y;
    constexpr int TCOLS = MT * R * 2;  // two accumulator buffers

    if (tid == 0) {
        for (int s = 0; s < P::STAGES; ++s) {
            mbar_init(&full[s], NUM_LOAD);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], NUM_EPI);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<TCOLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ loaders
        const int lt = tid - NUM_EPI;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < gp.num_tiles; tile += gridDim.x) {
            bl.begin_tile(tile, lt);
            for (int kc = 0; kc < gp.num_kc; ++kc, ++it) {
                const int s = it % P::STAGES;
                const uint32_t ph = (it / P::STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                uint8_t *A = smem + s * P::STAGE_BYTES;
                uint8_t *B = A + P::A_BYTES;
                if (lt == 0) {
                    mbar_expect_tx(&full[s], P::A_BYTES);
                    bulk_g2s(A, gp.a_packed + ((int64_t)mg * gp.num_kc + kc) * P::A_BYTES, P::A_BYTES, &full[s]);
                }
                bl.produce(B, kc, lt);
                fence_proxy_async_smem();
                mbar_arrive(&full[s]);
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t IDESC = idesc_bf16(128, R, false, BL::B_MN);
            uint32_t it = 0, tl = 0;
            for (int64_t tile = blockIdx.x; tile < gp.num_tiles; tile += gridDim.x, ++tl) {
                const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
                mbar_wait(&tempty[acc], aph ^ 1u);
                tc_fence_after();
                for (int kc = 0; kc < gp.num_kc; ++kc, ++it) {
                    const int s = it % P::STAGES;
                    const uint32_t ph = (it / P::STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_s = smem_u32(smem + s * P::STAGE_BYTES);
                    const uint32_t b_s = a_s + P::A_BYTES;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                        for (int ks = 0; ks < KC / 16; ++ks) {
                            const uint64_t ad = smem_desc_sw128(a_s + mt * (128 * LINE_BYTES) + ks * 32, 16, ATOM_BYTES);
                            const uint64_t bd = BL::b_desc(b_s, ks);
                            umma_bf16(tmem_base + acc * (MT * R) + mt * R, ad, bd, IDESC, (kc | ks) != 0);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0-3)
        ep.begin();
        uint32_t tl = 0;
        const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
        for (int64_t tile = blockIdx.x; tile < gp.num_tiles; tile += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int ch = (mg * MT + mt) * 128 + tid;
                ep.tile_mt(tmem_base + lane_base + acc * (MT * R) + mt * R, tile, ch, mt);
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) ep.finish_mt((mg * MT + mt) * 128 + tid, mt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<TCOLS>(tmem_base);
}

// =================================================================================================
//  small kernels
// =================================================================================================
// bf16 operand image of a weight matrix: element (m, k) = w[m*sm + k*sk] for m < M, k < Kd, else 0.
// Layout: [m_group][k_chunk][MT*128 lines][128 B] with the 128B swizzle applied per line.
__global__ void pack_weights_kernel(const float *w, int M, int Kd, int64_t sm, int64_t sk, int MT, int num_mg, int num_kc,
                                    uint8_t *img)
{
    const int lines = MT * 128;
    const int64_t total = (int64_t)num_mg * num_kc * lines * 8;  // one thread per 16-byte chunk
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i & 7);
    const int64_t t = i >> 3;
    const int line = (int)(t % lines);
    const int64_t t2 = t / lines;
    const int kc = (int)(t2 % num_kc);
    const int mgi = (int)(t2 / num_kc);
    const int m = mgi * lines + line;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kc * KC + c * 8 + e;
        f[e] = (m < M && k < Kd) ? w[(int64_t)m * sm + (int64_t)k * sk] : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    o.z = pack_bf16x2(f[4], f[5]);
    o.w = pack_bf16x2(f[6], f[7]);
    uint8_t *dst = img + (((int64_t)mgi * num_kc + kc) * lines) * LINE_BYTES + line_chunk_off(line, c);
    *reinterpret_cast<uint4 *>(dst) = o;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Packed {
    uint8_t *img;
    int MT, num_mg, num_kc;
    int64_t bytes;
};
static Packed plan_pack(int M, int Kd)
{
    Packed p;
    const int mpad = (int)align_up(M, 128);
    p.MT = mpad >= 256 ? 2 : 1;
    p.num_mg = mpad / (p.MT * 128);
    if (p.num_mg * p.MT * 128 < mpad) p.num_mg += 1;
    p.num_kc = (Kd + KC - 1) / KC;
    p.bytes = (int64_t)p.num_mg * p.num_kc * p.MT * 128 * LINE_BYTES;
    p.img = nullptr;
    return p;
}
static void launch_pack(const float *w, int M, int Kd, int64_t sm, int64_t sk, const Packed &p, cudaStream_t st)
{
    const int64_t total = p.bytes / 16;
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, M, Kd, sm, sk, p.MT, p.num_mg, p.num_kc, p.img);
    note_launch();
}

template <int MT, class BL, class EP>
static int launch_gemm(const Packed &pk, int64_t tiles, const BL &bl, const EP &ep, cudaStream_t st)
{
    using P = SmemPlan<MT>;
    auto kern = tc_rows_gemm_kernel<MT, BL, EP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, P::TOTAL);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int gx = sms / pk.num_mg;
    if (gx < 1) gx = 1;
    if ((int64_t)gx > tiles) gx = (int)tiles;
    GemmParams gp = {pk.img, pk.num_kc, tiles};
    dim3 grid((unsigned)gx, (unsigned)pk.num_mg);
    kern<<<grid, NT, P::TOTAL, st>>>(gp, bl, ep);
    note_launch();
    e = cudaPeekAtLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// grid.x used by launch_gemm (the statistics partials are indexed by blockIdx.x)
static int grid_x_for(const Packed &pk, int64_t tiles)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int gx = sms / pk.num_mg;
    if (gx < 1) gx = 1;
    if ((int64_t)gx > tiles) gx = (int)tiles;
    return gx;
}

// -------------------------------------------------------------------------------------------------
//  self-test: out[m][row] = sum_k w[m][k] * b(row, k) with b given row-major [rows][k] bf16 (mode 0, K-major
//  B tiles through the gather loader without positions) or feature-major [k][ld] bf16 (mode 1, MN-major)
// -------------------------------------------------------------------------------------------------
int tc_gemm_selftest(const float *w, int m_out, int k, const void *b, int mode, int64_t rows, int64_t ld, const float *zeros3,
                     float *out, int64_t ld_out, void *workspace, int64_t workspace_bytes, cudaStream_t st)
{
    if (!w || !b || !out || !workspace || !zeros3 || m_out <= 0 || k <= 0 || rows <= 0) return B2PN_EINVAL;
    Packed pk = plan_pack(m_out, k);
    if (workspace_bytes < pk.bytes + 1024) return B2PN_EINVAL;
    pk.img = (uint8_t *)align_up((int64_t)(uintptr_t)workspace, 1024);
    launch_pack(w, m_out, k, k, 1, pk, st);
    const int64_t tiles = (rows + R - 1) / R;
    RowMapTC rm = {B2PN_SEG_CLOUDS, 1, nullptr, nullptr, nullptr, rows, 0};
    StoreF32EpW ep = {{out, m_out, ld_out}};
    if (mode == 0) {
        // all k columns are features (c_in = k); the appended pos_j - 0 columns multiply zero weights
        GatherLoaderTC gl = {rm, (const __nv_bfloat16 *)b, k, zeros3, nullptr};
        return pk.MT == 1 ? launch_gemm<1>(pk, tiles, gl, ep, st) : launch_gemm<2>(pk, tiles, gl, ep, st);
    }
    FeatLoaderTC<0> fl = {rm, (const __nv_bfloat16 *)b, k, ld, nullptr, nullptr, 0};
    return pk.MT == 1 ? launch_gemm<1>(pk, tiles, fl, ep, st) : launch_gemm<2>(pk, tiles, fl, ep, st);
}

}  // namespace tc
}  // namespace b2pn
